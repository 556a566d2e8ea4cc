#!/bin/bash
mkdir -p gpurun_out
for o in "--opt:gemm_kgroup=1" "--opt:gemm_kgroup=2" "--opt:gemm_kgroup=4" "--opt:gemm_kgroup=6 --opt:gemm_kgroup_kb=48" "--opt:gemm_kgroup=9 --opt:gemm_kgroup_kb=40" "--opt:gemm_kgroup=9 --opt:gemm_kgroup_kb=64"; do
  echo "=== $o"; timeout 600 python tools/bench_configs.py c5 c3 --no-ref $o 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['segb200_img_s'],1), 'img/s', round(d['segb200_ms'],3), 'ms')"
done
