#!/bin/bash
# round 2, GPU call 6 (2 GPUs): data-parallel training with the fused SyncBatchNorm exchange: same-data equivalence with 1 GPU,
# A/B against per-layer NCCL all-reduces, bench.py --gpus 2 (both records)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "=== 1 GPU, same data"; timeout 600 python tools/bench_train.py --same-data --no-dropout --steps 4 --no-ref 2>gpurun_out/c6_a.err | tail -1 | tee gpurun_out/r2_train_same_data_1gpu.json | cut -c1-700
echo "=== 2 GPUs, same data, fused exchange"; timeout 600 $TR tools/bench_train.py --same-data --no-dropout --steps 4 --no-ref 2>gpurun_out/c6_b.err | tail -1 | tee gpurun_out/r2_train_same_data_2gpu_fused.json | cut -c1-700; tail -3 gpurun_out/c6_b.err
echo "=== 2 GPUs, fused exchange"; timeout 600 $TR tools/bench_train.py --steps 10 --no-ref 2>gpurun_out/c6_c.err | tail -1 | tee gpurun_out/r2_train_2gpu_fused.json | cut -c1-700; tail -3 gpurun_out/c6_c.err
echo "=== 2 GPUs, NCCL per-layer all-reduces"; SEGB200_NO_FUSED_SYNCBN=1 timeout 600 $TR tools/bench_train.py --steps 10 --no-ref 2>gpurun_out/c6_d.err | tail -1 | tee gpurun_out/r2_train_2gpu_nccl.json | cut -c1-700
echo "=== bench.py --gpus 2"; timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline 2>gpurun_out/c6_e.err | tail -1 > gpurun_out/r2_bench_2gpu.json; tail -3 gpurun_out/c6_e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu.json')); print({k: d.get(k) for k in ('value','ms_per_step','n_gpus','train')})
PY
