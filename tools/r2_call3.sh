#!/bin/bash
# round 2, GPU call 3: lean depthwise kernel (ring4) parity + A/B, GEMM per-role wait counters for the mid-K shapes
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
run python -m pytest tests/test_kernels_gpu.py -q -x -k "dw"
run python -m pytest tests/test_c_shim.py -q -m gpu -k reference
run python -m pytest tests/test_model_gpu.py -q -x
echo "=== bench (ring4)"; timeout 600 python bench.py --no-cudnn-ref --no-train --no-cpu-baseline --dump-kernels gpurun_out/r2_kernels_ring4.tsv > gpurun_out/r2_bench_ring4.json 2> gpurun_out/r2_bench_ring4.err; tail -2 gpurun_out/r2_bench_ring4.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_bench_ring4.json')); print(d['value'], d['ms_per_step'], d['per_kind_ms'], d['roofline_dw'])
PY
grep dwconv gpurun_out/r2_kernels_ring4.tsv | sort -t$'\t' -k7 | awk -F'\t' '{a[$7]+=$2; n[$7]++; b[$7]=$4} END{for(k in a) printf "%s\tn=%d\tms=%.3f\tGB/s=%.0f\n", k, n[k], a[k], b[k]*n[k]/a[k]/1e3*1e3/1e3}' | sort
echo "=== gemm waits (DBG build)"
SEGB200_LIB=$PWD/segmentron_b200/libsegb200_dbg.so TAIL=60 run python tools/gemm_waits.py 0
