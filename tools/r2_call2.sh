#!/bin/bash
# round 2, GPU call 2: new tests (segmentron._C shim, reference in the loop, full-size parity), smoke kernel capture, bench with the
# real reference as cudnn_ref + the training sub-record, DANet multi-seed statistics
mkdir -p gpurun_out
run() { echo "=== $*"; timeout ${TMO:-600} "$@" 2>&1 | tail -${TAIL:-40}; echo "=== exit ${PIPESTATUS[0]}"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
run python -m pytest tests/test_c_shim.py -q -s -m gpu
TAIL=60 run python -m pytest tests/test_fullsize_parity_gpu.py -q -s
TAIL=60 TMO=900 run python -m pytest tests/test_reference_loop_gpu.py -q -s
run python __graft_entry__.py smoke
echo "=== ncu launch list of smoke()"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_ncu.log 2>&1
grep -o '"[a-zA-Z_0-9:<>, ]*kernel[a-zA-Z_0-9:<>, ]*"' gpurun_out/r2_smoke_launches.csv | sed 's/<.*//' | sort | uniq -c | sort -rn | head -30
TAIL=5 run python tools/danet_diag.py --seeds 6
echo "=== bench"
timeout 900 python bench.py > gpurun_out/r2_bench_call2.json 2> gpurun_out/r2_bench_call2.err; echo "exit $?"; tail -3 gpurun_out/r2_bench_call2.err; cat gpurun_out/r2_bench_call2.json
