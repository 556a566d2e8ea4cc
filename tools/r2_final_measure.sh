#!/bin/bash
# final measurement of the round: headline bench (all records), ncu launch list of a step, ncu --set full of the top kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "=== bench.py (default: value, e2e, e2e_logits, roofline, cpu_baseline, cudnn_ref from the real reference, train sub-record)"
timeout 1200 python bench.py --dump-kernels gpurun_out/r2_kernels_per_launch.tsv > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_final.err; echo "exit $?"; tail -2 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_1gpu_final.json'))
for k in ('value','ms_per_step','e2e','e2e_logits','roofline','roofline_all_gemm','roofline_dw','per_kind_ms','cudnn_ref','vs_cudnn_ref','train','clocks','cpu_baseline','gpu_launches'):
    print(k, json.dumps(d.get(k))[:700])
PY
echo "=== bench.py --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-600 gpurun_out/r2_bench_reference_arm.json
echo "=== ncu launch list of one step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cudnn-ref --no-train --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1; tail -1 gpurun_out/r2_ncu_bench.log | cut -c1-200
echo "=== ncu --set full of the depthwise and GEMM kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv|conv_gemm|bilinear" -o gpurun_out/r2_prof_full python tools/prof_kernels.py > gpurun_out/r2_prof_full.log 2>&1; tail -2 gpurun_out/r2_prof_full.log
echo "=== the other configs (inference), engine only"
timeout 900 python tools/bench_configs.py c3 c4 c5 c1 --no-ref 2>/dev/null | tee gpurun_out/r2_other_configs_final.jsonl | cut -c1-300
echo "=== kernel capture of smoke()"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_ncu.log 2>&1; grep -c segb200 gpurun_out/r2_smoke_launches.csv
echo "=== full GPU suite + smoke"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
