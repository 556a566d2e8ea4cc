#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
TAIL=8 run python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -x
echo "=== bench"; timeout 600 python bench.py --no-cudnn-ref --no-train --no-cpu-baseline --dump-kernels gpurun_out/r2_kernels_c11.tsv > gpurun_out/r2_bench_c11.json 2> gpurun_out/r2_bench_c11.err; tail -2 gpurun_out/r2_bench_c11.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_bench_c11.json')); print(d['value'], d['ms_per_step'], d['per_kind_ms'], d['roofline_dw']['frac'], d['roofline']['frac'], d['roofline_all_gemm']['frac'])
PY
grep conv_gemm gpurun_out/r2_kernels_c11.tsv | awk -F'\t' '{a[$7]+=$2; n[$7]++; f[$7]+=$3} END{for(k in a) printf "%.3f ms\tn=%d\t%.0f TF/s\t%s\n", a[k], n[k], f[k]/a[k], k}' | sort -rn | head -8
TAIL=12 TMO=900 run python tools/bench_configs.py c5 c3 c1 --no-ref --kinds
