#!/bin/bash
# first-contact GPU script: isolate each test group in its own process so one trap does not hide the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { echo "=== $*"; timeout 600 "$@" 2>&1 | tail -40; echo "=== exit $?"; }
run python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "pool or layout or errors" 
run python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dwconv"
run python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "test_conv_gemm and pw_flat_64_128 and bf16"
run python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "test_conv_gemm and pw_flat"
run python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "test_conv_gemm and (c3_ or pw_s2)"
run python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "stem"
run python -m pytest tests/test_model_gpu.py -q -m gpu -s -k "65x129"
run python -c "import __graft_entry__ as g; g.smoke()"
