#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "nonlocal" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_model_gpu.py -q -x -s -k "ocnet or danet_resnet101_64x96 or pspnet" 2>&1 | grep -v Warning | tail -25
timeout 600 python -m pytest tests/test_modules_gpu.py -q -x -k "dropin_modules or pam" 2>&1 | tail -4
