#!/bin/bash
# Round-2 first contact (one gpurun call, ~6 GPU-minutes): everything that was written after round 1's GPU budget ran out, each group
# in its own process so that one failure / trap does not hide the rest.  Un-gate (drop the SEGB200_TEST_ALL conditions) what passes.
#   gpurun --timeout 900 -- 'bash tools/r2_first_contact.sh > gpurun_out/r2_first_contact.log 2>&1'
mkdir -p gpurun_out
export SEGB200_TEST_ALL=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
run() { echo "=== $*"; timeout 300 "$@" 2>&1 | tail -25; echo "=== exit ${PIPESTATUS[0]}"; }
run python -m pytest tests/test_modules_gpu.py -q -k "errors_and_cache"                      # gated part: _ASPP batch-1 ValueError
run python -m pytest tests/test_train_model_gpu.py -q -s -k "ccnet"                         # CCNet training plan, every launch
run python -m pytest tests/test_train_model_gpu.py -q -s -k "danet"                         # DANet training plan, every launch (PAM / CAM training kernels)
run python -m pytest tests/test_train_model_gpu.py -q -s -k "hrnet"                         # HRNet training plan, every launch (+ upsample_add_bwd)
run python -m pytest tests/test_modules_gpu.py -q -k "cam_module_backward or pam_module_backward"   # CAM / PAM training drop-ins
run python -m pytest tests/test_metric_gpu.py -q                                            # device-resident pixAcc / mIoU counts
run python -m pytest tests/test_evaluate_gpu.py -q                                          # multi-scale + flip driver
run python -m pytest tests/test_train_kernels_gpu.py -q -k "wgrad_v2 or upsample_add"                     # opt-in sliding-window depthwise weight gradient
run python -m pytest tests -q -m gpu -x                                                     # the whole suite with everything un-gated
