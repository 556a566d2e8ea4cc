#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu_final.json 2> gpurun_out/r2_bench_2gpu_final.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_final.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['n_gpus'], round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1)); print('train', json.dumps(d.get('train'))[:600]); print(d['clocks'])
PY
tail -3 gpurun_out/r2_bench_2gpu_final.err | cut -c1-300
