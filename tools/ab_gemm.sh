#!/bin/bash
# A/B the conv_gemm kernel across library builds (SEGB200_LIB override)
for lib in "$@"; do
  echo "=== $lib"
  SEGB200_LIB=$PWD/$lib python tools/gemm_waits.py 0 2>&1 | grep -E "^pw|^c3|Error|error" 
done
