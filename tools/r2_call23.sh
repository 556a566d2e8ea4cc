#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv_gemm or stem" 2>&1 | tail -8
for d in 0 1; do
  SEGB200_OPTS="gemm_dual=$d" timeout 600 python bench.py --steps 20 --warmup 5 --no-cudnn-ref --no-train --no-cpu-baseline --dump-kernels gpurun_out/r2_kernels_dual$d.tsv > gpurun_out/c23_bench_dual$d.json 2> gpurun_out/c23_bench_dual$d.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c23_bench_dual$d.json').read().strip().splitlines()[-1])
print('dual=$d', round(d['value'],1), round(d['ms_per_step'],3), d.get('per_kind_ms'), round(d['roofline']['frac'],3), round(d['roofline_all_gemm']['frac'],3))
PY
done
