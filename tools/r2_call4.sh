#!/bin/bash
# round 2, GPU call 4: CCNet fp16 parity A/B (host vs device folding, repeated), GEMM resource decomposition (DBG modes), ncu of the
# lean depthwise kernel
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
for i in 1 2; do TAIL=4 run python -m pytest tests/test_model_gpu.py -q -s -k "ccnet and f16"; done
SEGB200_DEVICE_FOLD=1 TAIL=4 run python -m pytest tests/test_model_gpu.py -q -s -k "ccnet and f16"
TAIL=6 run python -m pytest tests/test_c_shim.py -q -m gpu -k reference
echo "=== gemm waits (DBG build), modes 0..3"
SEGB200_LIB=$PWD/segmentron_b200/libsegb200_dbg.so TAIL=80 run python tools/gemm_waits.py 0 1 2 3
echo "=== ncu dw"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -o gpurun_out/r2_prof_dw python tools/prof_kernels.py dw > gpurun_out/r2_prof_dw.log 2>&1; tail -3 gpurun_out/r2_prof_dw.log
