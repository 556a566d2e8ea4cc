#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
TAIL=8 run python -m pytest tests/test_kernels_gpu.py -q -x
SEGB200_LIB=$PWD/segmentron_b200/libsegb200_dbg.so TAIL=40 run python tools/gemm_waits.py 0
echo "=== bench"; timeout 600 python bench.py --no-cudnn-ref --no-train --no-cpu-baseline --dump-kernels gpurun_out/r2_kernels_c9.tsv > gpurun_out/r2_bench_c9.json 2> gpurun_out/r2_bench_c9.err; tail -2 gpurun_out/r2_bench_c9.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_bench_c9.json')); print(d['value'], d['ms_per_step'], d['per_kind_ms'], d['roofline_dw']['frac'], d['roofline']['frac'], d['roofline_all_gemm']['frac'])
PY
grep conv_gemm gpurun_out/r2_kernels_c9.tsv | awk -F'\t' '{a[$7]+=$2; n[$7]++; f[$7]+=$3} END{for(k in a) printf "%.3f ms\tn=%d\t%.0f TF/s\t%s\n", a[k], n[k], f[k]/a[k], k}' | sort -rn | head -14
TAIL=8 run python -m pytest tests/test_model_gpu.py -q -x
