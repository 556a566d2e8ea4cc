#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "dwconv" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_train_model_gpu.py -q -x -k "xception65_65x129 or vs_oracle_larger or every_launch" 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
