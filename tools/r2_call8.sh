#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
TAIL=14 run python tools/mma_probe.py
TAIL=25 run python -m pytest tests/test_kernels_gpu.py -q
TAIL=25 run python -m pytest tests/test_model_gpu.py -q -s -k "danet or pspnet"
TAIL=25 run python -m pytest tests/test_modules_gpu.py tests/test_train_kernels_gpu.py -q
