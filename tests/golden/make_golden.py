"""Generate the golden fixtures that pin oracle/segref.py to the REAL reference.

Run in the build container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

For every case it (1) builds the oracle's seeded parameter dict, (2) builds the reference
nn.Module from its own YAML config / module class and ``load_state_dict(strict=True)``s that
dict (so key names and shapes are proven to be the reference's), (3) runs the reference on a
seeded input, (4) asserts the oracle reproduces it, and (5) stores the reference output
(fp16-compressed where large) in ``tests/golden/<case>.pt``.  One reference config per process
(the reference's ``cfg`` is a frozen global), hence the subprocess fan-out.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

MODEL_CASES = {
    # case: (oracle model name, reference yaml, input shape, seed)
    "dlv3p_xception65_65x129": ("deeplabv3plus_xception65", "cityscapes_deeplabv3_plus.yaml", (1, 3, 65, 129), 0),
    "dlv3p_xception65_97x161_b2": ("deeplabv3plus_xception65", "cityscapes_deeplabv3_plus.yaml", (2, 3, 97, 161), 1),
    "dlv3p_mobilenetv2_64x128": ("deeplabv3plus_mobilenet_v2", "cityscapes_deeplabv3_plus_mobilenet.yaml", (1, 3, 64, 128), 2),
    "dlv3p_resnet101_65x129": ("deeplabv3plus_resnet101", "cityscapes_deeplabv3_plus_resnet.yaml", (1, 3, 65, 129), 4),
    "danet_resnet101_64x96": ("danet_resnet101", "cityscapes_danet_resnet.yaml", (1, 3, 64, 96), 5),
    "ccnet_resnet101_65x97": ("ccnet_resnet101", "cityscapes_ccnet_resnet.yaml", (1, 3, 65, 97), 6),
    "hrnet_w18s_128x192": ("hrnet_w18_small_v1", "cityscapes_hrnet_w18_small_v1.yaml", (2, 3, 128, 192), 7),
}


def run_model_case(case):
    import numpy as np
    np.int = int
    import torch
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    from oracle import segref as R
    name, yaml_file, shape, seed = MODEL_CASES[case]
    if name == "ccnet_resnet101":
        # The reference's CCNet needs the pybind module `segmentron._C`, which cannot be built (THC headers, App. B5) and
        # is commented out of models/__init__.py.  Provide a stand-in whose two forward functions are the oracle's
        # ca_weight / ca_map (themselves pinned to a scalar transcription of ca_cuda.cu in run_module_cases), so the
        # fixture pins the reference's MODEL structure (_RCCAModule, recurrence, cat order, bottleneck) around them.
        import types
        import segmentron
        fake = types.ModuleType("segmentron._C")
        fake.ca_forward = lambda t, f: R.ca_weight(t, f)
        fake.ca_map_forward = lambda w, g: R.ca_map(w, g)
        sys.modules["segmentron._C"] = fake
        segmentron._C = fake
        import segmentron.models.ccnet  # noqa: F401  (registers "CCNet")
    from segmentron.config import cfg
    from segmentron.models.model_zoo import get_segmentation_model
    cfg.update_from_file(os.path.join(REF, "configs", yaml_file))
    cfg.PHASE = "test"
    cfg.check_and_freeze()
    model = get_segmentation_model().eval()
    # tools/eval.py:50-53 : BN eps override for the encoder
    if cfg.MODEL.BN_EPS_FOR_ENCODER:
        for m in model.encoder.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eps = cfg.MODEL.BN_EPS_FOR_ENCODER
    P = R.build_params(name, seed)
    missing = model.load_state_dict(P.state_dict(), strict=True)
    if name == "danet_resnet101":                       # three outputs (sasc, sa, sc): check all, store the first
        g_ = torch.Generator().manual_seed(1000 + seed)
        x_ = torch.randn(*shape, generator=g_)
        with torch.no_grad():
            yr, yo = model(x_), R.forward(name, P, x_, all=True)
        for a_, b_ in zip(yr, yo):
            assert float((a_ - b_).abs().max() / a_.abs().max()) < 1e-5
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(*shape, generator=g)
    with torch.no_grad():
        y_ref = model(x)[0]
        y_orc = R.forward(name, P, x)
    err = float((y_ref - y_orc).abs().max() / y_ref.abs().max())
    assert err < 1e-5, f"oracle != reference for {case}: {err}"
    out = dict(case=case, model=name, seed=seed, input_seed=1000 + seed, shape=shape,
               y_ref=y_ref.contiguous(), argmax=y_ref.argmax(1).to(torch.uint8),
               oracle_vs_ref_maxrel=err, n_params=len(P.t))
    torch.save(out, os.path.join(HERE, case + ".pt"))
    print(f"{case}: oracle vs reference max-rel {err:.2e}; saved {tuple(y_ref.shape)}")


def run_module_cases():
    """Module-level fixtures: PAM, CAM, PyramidPooling from the reference classes; criss-cross
    attention from a line-by-line python transcription of ca_cuda.cu's index map (the CUDA
    extension cannot be built: THC headers are gone)."""
    import numpy as np
    np.int = int
    import torch
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    from oracle import segref as R
    from segmentron.modules.module import PAM_Module, CAM_Module, PyramidPooling
    out = {}
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 64, 12, 20, generator=g)
    # PAM
    P = R.Params(11)
    with torch.no_grad():
        y_o = R.pam(P, x, "pam")
        m = PAM_Module(64).eval()
        m.load_state_dict({k[len("pam."):]: v for k, v in P.state_dict().items()}, strict=True)
        y_r = m(x)
    assert float((y_r - y_o).abs().max()) < 1e-4
    out["pam"] = dict(x=x, y=y_r, seed=11)
    # CAM
    P = R.Params(12)
    with torch.no_grad():
        y_o = R.cam(P, x * 0.2, "cam")
        m = CAM_Module(64).eval()
        m.load_state_dict({k[len("cam."):]: v for k, v in P.state_dict().items()}, strict=True)
        y_r = m(x * 0.2)
    assert float((y_r - y_o).abs().max()) < 1e-4
    out["cam"] = dict(x=x * 0.2, y=y_r, seed=12)
    # PyramidPooling (PSPNet() itself cannot be constructed: App. B2)
    P = R.Params(13)
    xp = torch.randn(1, 64, 17, 33, generator=g)
    with torch.no_grad():
        y_o = R.pyramid_pooling(P, xp, "psp")
        m = PyramidPooling(64).eval()
        m.load_state_dict({k[len("psp."):]: v for k, v in P.state_dict().items()}, strict=True)
        y_r = m(xp)
    assert float((y_r - y_o).abs().max()) < 1e-4
    out["psp"] = dict(x=xp, y=y_r, seed=13)
    # criss-cross: scalar loops written from ca_cuda.cu:8-36 and :94-120
    n, c, h, w = 1, 4, 5, 7
    t = torch.randn(n, c, h, w, generator=g); f = torch.randn(n, c, h, w, generator=g)
    wgt = torch.zeros(n, h + w - 1, h, w)
    for b in range(n):
        for y in range(h):
            for xx in range(w):
                for z in range(h + w - 1):
                    for pl in range(c):
                        if z < w:
                            wgt[b, z, y, xx] += t[b, pl, y, xx] * f[b, pl, y, z]
                        else:
                            i = z - w
                            j = i if i < y else i + 1
                            wgt[b, w + i, y, xx] += t[b, pl, y, xx] * f[b, pl, j, xx]
    assert float((wgt - R.ca_weight(t, f)).abs().max()) < 1e-5
    gv = torch.randn(n, 6, h, w, generator=g)
    att = torch.softmax(wgt, 1)
    o = torch.zeros(n, 6, h, w)
    for b in range(n):
        for pl in range(6):
            for y in range(h):
                for xx in range(w):
                    for i in range(w):
                        o[b, pl, y, xx] += gv[b, pl, y, i] * att[b, i, y, xx]
                    for i in range(h):
                        if i == y:
                            continue
                        j = i if i < y else i - 1
                        o[b, pl, y, xx] += gv[b, pl, i, xx] * att[b, w + j, y, xx]
    assert float((o - R.ca_map(att, gv)).abs().max()) < 1e-5
    out["cca"] = dict(t=t, f=f, weight=wgt, g=gv, att=att, out=o)
    torch.save(out, os.path.join(HERE, "modules.pt"))
    print("modules: pam/cam/psp/cca OK")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        (run_module_cases() if sys.argv[1] == "modules" else run_model_case(sys.argv[1]))
    else:
        env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
        for c in list(MODEL_CASES) + ["modules"]:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), c], env=env)
