#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_metric_gpu.py -q -k "dwconv or upsample or resize or metric" 2>&1 | tail -8
timeout 300 python tools/dw_sweep.py --cols2 > gpurun_out/r2_dw_sweep_cols2_dil.jsonl 2> gpurun_out/c21_sweep.err; cut -c1-200 gpurun_out/r2_dw_sweep_cols2_dil.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-cudnn-ref --no-train --dump-kernels gpurun_out/r2_kernels_c21.tsv > gpurun_out/c21_bench.json 2> gpurun_out/c21_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c21_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('per_kind_ms'), d['roofline_dw']['frac'], d['e2e']['value'])
PY
