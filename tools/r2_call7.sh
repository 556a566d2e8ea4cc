#!/bin/bash
# round 2, GPU call 7: dw knob sweep, MMA issue-protocol probe, PSPNet + fixed tests, bench
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
TAIL=12 run python tools/mma_probe.py
TAIL=60 run python tools/dw_sweep.py
TAIL=8 run python -m pytest tests/test_model_gpu.py tests/test_c_shim.py -q -m gpu -s -k "pspnet or ccnet or reference"
echo "=== bench"; timeout 600 python bench.py --no-cudnn-ref --no-train --no-cpu-baseline > gpurun_out/r2_bench_c7.json 2> gpurun_out/r2_bench_c7.err; tail -2 gpurun_out/r2_bench_c7.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_bench_c7.json')); print(d['value'], d['ms_per_step'], d['per_kind_ms'], d['roofline_dw']['frac'], d['roofline']['frac'], d['roofline_all_gemm']['frac'])
PY
