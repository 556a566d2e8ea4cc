"""DANet parity statistics over several seeds (the CAM softmax amplifies 16-bit noise, so one fixture is one noisy sample):
rel-L2 against the fp32 oracle of (a) this engine and (b) the reference's own 16-bit forward (oracle port, same torch ops),
bf16 and fp16, 64x96 and 128x192 inputs.  One JSON line.   python tools/danet_diag.py [--seeds 8]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=8)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from oracle import segref as R
    from segmentron_b200.engine import DANetB200
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    model = "danet_resnet101"
    out = {}
    for dtype in (torch.bfloat16, torch.float16):
        for shape in ((1, 3, 64, 96), (1, 3, 128, 192)):
            rows = []
            for seed in range(a.seeds):
                P = R.build_params(model, 100 + seed)
                x = torch.randn(*shape, generator=torch.Generator().manual_seed(200 + seed)).cuda()
                with torch.no_grad():
                    y32 = R.forward(model, P.to("cuda"), x).float()
                    y16 = R.forward(model, P.to("cuda", dtype), x.to(dtype)).float()
                y = DANetB200(P.state_dict(), dtype=dtype)(x).float()
                rel = lambda u: float((u - y32).norm() / y32.norm())       # noqa: E731
                rows.append((rel(y), rel(y16)))
            ours = [r[0] for r in rows]
            ref = [r[1] for r in rows if r[1] == r[1]]
            key = f"{str(dtype).replace('torch.', '')}_{shape[2]}x{shape[3]}"
            out[key] = {"ours": [round(v, 5) for v in ours], "ref16": [round(r[1], 5) if r[1] == r[1] else None for r in rows],
                        "ours_mean": sum(ours) / len(ours), "ref16_mean": (sum(ref) / len(ref)) if ref else None,
                        "ref16_nan": len(rows) - len(ref)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
