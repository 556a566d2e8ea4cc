#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
TAIL=8 run python -m pytest tests/test_kernels_gpu.py -q -x -k "2cta or conv"
TAIL=20 run python tools/gemm_2cta_probe.py
