#!/bin/bash
mkdir -p gpurun_out
echo "=== bench_configs (with the cuDNN reference leg)"; timeout 1200 python tools/bench_configs.py c3 c4 c5 c1 --kinds 2>gpurun_out/c13_cfg.err | grep '^{' | tee gpurun_out/r2_other_configs.jsonl | cut -c1-900
echo "=== bench_train resnet101 (1 GPU, with reference leg)"; timeout 900 python tools/bench_train.py --steps 8 2>gpurun_out/c13_tr.err | tail -1 | tee gpurun_out/r2_bench_train_resnet101.json | cut -c1-1500
echo "=== bench_train xception65"; timeout 900 python tools/bench_train.py --steps 6 --backbone xception65 --no-ref 2>>gpurun_out/c13_tr.err | tail -1 | tee gpurun_out/r2_bench_train_xception65.json | cut -c1-900
