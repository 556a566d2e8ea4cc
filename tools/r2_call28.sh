#!/bin/bash
mkdir -p gpurun_out
for o in "pdl=0" "pdl=1"; do
  SEGB200_OPTS="$o" timeout 600 python bench.py --steps 30 --warmup 5 --no-cudnn-ref --no-train --no-cpu-baseline > gpurun_out/c28_bench_$o.json 2> gpurun_out/c28_bench_$o.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c28_bench_$o.json').read().strip().splitlines()[-1])
print('$o', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'logits', round(d['e2e_logits']['value'],1))
PY
done
for g in "" "--graph"; do for o in "pdl=0" "pdl=1"; do SEGB200_OPTS="$o" timeout 600 python tools/bench_train.py --steps 8 --no-ref $g 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train $g $o', d['segb200_ms_per_step'], d['segb200_img_s'])"; done; done
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
