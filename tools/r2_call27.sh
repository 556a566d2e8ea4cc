#!/bin/bash
mkdir -p gpurun_out
bash tools/r2_validate.sh 2>&1 | grep -v "Warning\|warnings.warn\|Did you mean\|how-to/assert" | tee gpurun_out/r2_validate.log | tail -45
timeout 900 python bench.py --steps 20 --warmup 5 --no-train --dump-kernels gpurun_out/r2_kernels_c27.tsv > gpurun_out/c27_bench.json 2> gpurun_out/c27_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c27_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), d.get('per_kind_ms'), 'e2e', round(d['e2e']['value'],1), 'logits', round(d['e2e_logits']['value'],1))
print('roofline', round(d['roofline']['frac'],3), round(d['roofline_all_gemm']['frac'],3), round(d['roofline_dw']['frac'],3))
print('cudnn_ref', {k:(round(v,2) if isinstance(v,float) else v) for k,v in (d.get('cudnn_ref') or {}).items() if not isinstance(v,dict)})
PY
