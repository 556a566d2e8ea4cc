#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
echo "=== bench.py --gpus 8"; timeout 1200 $TR bench.py --gpus 8 --no-cpu-baseline --no-cudnn-ref 2>gpurun_out/c15.err | tail -1 > gpurun_out/r2_bench_8gpu.json; tail -4 gpurun_out/c15.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_8gpu.json')); print({k: d.get(k) for k in ('value','ms_per_step','n_gpus','train','e2e')})
PY
echo "=== train 8 GPUs, NCCL per-layer form (A/B)"; SEGB200_NO_FUSED_SYNCBN=1 timeout 600 $TR tools/bench_train.py --steps 8 --no-ref 2>>gpurun_out/c15.err | tail -1 | tee gpurun_out/r2_train_8gpu_nccl.json | cut -c1-400
