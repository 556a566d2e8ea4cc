#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
TAIL=10 run python -m pytest tests/test_modules_gpu.py -q -x -k "attention or pam"
TAIL=12 run python -m pytest tests/test_model_gpu.py tests/test_fullsize_parity_gpu.py -q -s -k "danet or C4"
TAIL=6 run python tools/bench_configs.py c4 --no-ref --kinds
