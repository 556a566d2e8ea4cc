"""Two full-size training iterations (config 3) for profilers: the first warms up (plan build, module load), the second is the
one to read.  `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<engine kernels> --csv --log-file ...`
lists every launch; the second half of the list is one steady-state step.  Prints the number of engine launches per step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from oracle import segref as R  # noqa: E402  (parameter generator only)
from segmentron_b200.train import DeepLabV3PlusTrainerB200  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1025, 2049)
P = R.build_params("deeplabv3plus_resnet101", 0)
g = torch.Generator().manual_seed(1024)
x = torch.randn(b, 3, h, w, generator=g).cuda()
t = torch.randint(-1, 19, (b, h, w), generator=g).cuda()
tr = DeepLabV3PlusTrainerB200(P.state_dict(), dtype=torch.bfloat16)
for i in range(2):
    loss = tr.step(x, t)
    torch.cuda.synchronize()
    print(f"step {i}: loss {float(loss):.4f}", flush=True)
pl = tr.plan_for(x.shape)["plan"]
n_engine = sum(1 for s in pl.fwd + pl.bwd if s.kind not in ("zero", "allreduce")) + 1 + 2 + 2      # + ce_finalize, 2 gather_cast, 2 sgd
print(f"engine kernel launches per step: {n_engine}", flush=True)
