#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "dwconv" 2>&1 | tail -8
timeout 300 python tools/dw_sweep.py --cols2 > gpurun_out/r2_dw_sweep_cols2.jsonl 2> gpurun_out/c20_sweep.err; cat gpurun_out/r2_dw_sweep_cols2.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-cudnn-ref --no-train > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c20_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('per_kind_ms'), d['roofline_dw']['frac'])
PY
