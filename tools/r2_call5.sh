#!/bin/bash
# round 2, GPU call 5: persistent dw kernel parity + bench, fused SyncBN protocol test, MMA probe, fold diff diagnostic
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
TAIL=6 run python -m pytest tests/test_kernels_gpu.py -q -x
TAIL=15 run python -m pytest tests/test_syncbn_gpu.py -q -x
TAIL=6 run python -m pytest tests/test_c_shim.py -q -m gpu -k reference
TAIL=30 run python tools/fold_diff.py
TAIL=12 run python tools/mma_probe.py
echo "=== bench"; timeout 600 python bench.py --no-cudnn-ref --no-train --no-cpu-baseline --dump-kernels gpurun_out/r2_kernels_c5.tsv > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err; tail -2 gpurun_out/r2_bench_c5.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_bench_c5.json')); print(d['value'], d['ms_per_step'], d['per_kind_ms'], d['roofline_dw']['frac'], d['roofline']['frac'], d['roofline_all_gemm']['frac'])
PY
grep dwconv gpurun_out/r2_kernels_c5.tsv | awk -F'\t' '{a[$7]+=$2; n[$7]++; b[$7]+=$4} END{for(k in a) printf "%.3f ms\tn=%d\t%.0f GB/s\t%s\n", a[k], n[k], b[k]/a[k], k}' | sort -rn
TAIL=8 run python -m pytest tests/test_model_gpu.py -q
