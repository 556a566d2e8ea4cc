#!/bin/bash
# full GPU validation: the whole -m gpu suite, smoke(), and a short bench
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-1200} "$@" 2>&1 | tail -${TAIL:-25}; echo "=== exit ${PIPESTATUS[0]}"; }
TAIL=30 run python -m pytest tests -q -m gpu
TAIL=6 run python __graft_entry__.py smoke
echo "=== train bench eager vs graph"; for g in "" "--graph"; do timeout 600 python tools/bench_train.py --steps 8 --no-ref $g 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$g', d['segb200_ms_per_step'], d['segb200_img_s'])"; done
