#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "dwconv" 2>&1 | tail -3
timeout 300 python tools/dw_sweep.py --cols2 2> gpurun_out/c30_sweep.err | tee gpurun_out/r2_dw_sweep_cw5.jsonl | cut -c1-160
for o in "dw_cw5=0" "dw_cw5=1"; do
  SEGB200_OPTS="$o" timeout 600 python bench.py --steps 30 --warmup 5 --no-cudnn-ref --no-train --no-cpu-baseline > gpurun_out/c30_bench_$o.json 2> gpurun_out/c30_bench_$o.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c30_bench_$o.json').read().strip().splitlines()[-1])
print('$o', round(d['value'],1), round(d['ms_per_step'],3), d['per_kind_ms'], round(d['roofline_dw']['frac'],3), d['clocks']['sm_mhz'])
PY
done
