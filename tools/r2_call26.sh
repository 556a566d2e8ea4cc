#!/bin/bash
timeout 900 python -m pytest tests/test_model_gpu.py -q -s -k "ocnet" 2>&1 | grep -E "^\[|passed|failed|Error" | tail -12
timeout 600 python - <<'PY' 2>&1 | tail -5
import torch, time, sys
sys.path.insert(0, '.')
from oracle import segref as R
from segmentron_b200.engine import OCNetB200
P = R.build_params("ocnet_resnet50", 0)
x = torch.randn(4, 3, 1024, 2048, generator=torch.Generator().manual_seed(1)).cuda()
eng = OCNetB200(P.state_dict(), dtype=torch.bfloat16, want_argmax=True)
for _ in range(3): eng(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"[OCNet base/ResNet50 bf16 4x1024x2048, N=8192 tokens] {ms:.2f} ms/step = {4e3/ms:.1f} img/s")
rows = eng.plan_for(x)["plan"].run_timed(); rows = eng.plan_for(x)["plan"].run_timed()
agg = {}
for m, t in rows: agg[m["kind"]] = agg.get(m["kind"], 0.0) + t
print({k: round(v, 3) for k, v in agg.items()})
# the reference's own bf16 forward (oracle port, torch ops through cuDNN/cuBLAS) on the same GPU
Pd = P.to("cuda", torch.bfloat16); xb = x.to(torch.bfloat16)
with torch.no_grad():
    for _ in range(2): R.forward("ocnet_resnet50", Pd, xb)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3): R.forward("ocnet_resnet50", Pd, xb)
    torch.cuda.synchronize(); tr = (time.time() - t0) / 3 * 1e3
print(f"[torch bf16 (cuDNN/cuBLAS, materialised N x N)] {tr:.2f} ms/step = {4e3/tr:.1f} img/s -> x{tr/ms:.2f}")
PY
