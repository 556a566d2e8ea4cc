"""Diagnostic: build the CCNet fp16 plan with host-side and with device-side folding / packing and report every derived tensor
(packed weights, folded scale / shift) that differs, with the size of the difference."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import segref as R                      # noqa: E402
from segmentron_b200 import engine as E             # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "ccnet_resnet101"
dtype = torch.float16
fx = torch.load(os.path.join(ROOT, "tests", "golden", "ccnet_resnet101_65x97.pt"))
P = R.build_params(fx["model"], fx["seed"])
x = torch.randn(*fx["shape"], generator=torch.Generator().manual_seed(fx["input_seed"])).cuda()
plans = []
for dev_fold in (False, True):
    E._DEVICE_FOLD = dev_fold
    eng = E.CCNetB200(P.state_dict(), dtype=dtype, cuda_graph=False)
    y = eng(x).float().clone()
    pl = eng.plan_for(x)["plan"]
    plans.append((pl, y))
(pa, ya), (pb, yb) = plans
print("outputs: rel diff host-fold vs device-fold", float((ya - yb).norm() / yb.norm()), "vs golden:",
      float((ya.cpu() - fx["y_ref"]).norm() / fx["y_ref"].norm()), float((yb.cpu() - fx["y_ref"]).norm() / fx["y_ref"].norm()))
ka = [t for t in pa.keep if t.dtype in (torch.float32, dtype) and t.numel() < 5e7]
kb = [t for t in pb.keep if t.dtype in (torch.float32, dtype) and t.numel() < 5e7]
print(len(ka), len(kb))
nd = 0
for i, (a, b) in enumerate(zip(ka, kb)):
    if a.shape != b.shape or a.dtype != b.dtype:
        print("shape/dtype mismatch at", i, a.shape, b.shape, a.dtype, b.dtype); break
    if a.data_ptr() == b.data_ptr():
        continue
    d = (a.float() - b.float()).abs()
    if float(d.max()) > 0 and not (a.numel() > 1e5 and a.dim() == 4 and a.shape[1] > 4):      # skip activations
        nd += 1
        if nd <= 25:
            print(f"#{i} {tuple(a.shape)} {a.dtype}: max abs diff {float(d.max()):.3e}, max |b| {float(b.float().abs().max()):.3e}, n diff {int((d > 0).sum())}")
print("differing derived tensors:", nd)
