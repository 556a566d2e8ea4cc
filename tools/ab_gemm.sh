#!/bin/bash
# A/B the conv_gemm kernel across library builds (SEGB200_LIB override)
for lib in segmentron_b200/libsegb200.so build/lib_s8.so build/lib_s32_nodbg.so build/lib_s8_nodbg.so build/lib_r1c.so; do
  echo "=== $lib"
  SEGB200_LIB=$PWD/$lib python tools/gemm_waits.py 0 2>&1 | grep -E "^pw|^c3" 
done
