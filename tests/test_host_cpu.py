"""CPU-side checks: the C-ABI library loads and exports every symbol include/segb200.h declares, host-side
weight packing is exact, and the product refuses to run without CUDA (no fallback)."""
import os
import re
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from segmentron_b200 import lib
    return lib


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "segb200.h")).read()
    declared = set(re.findall(r"\b(segb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = built.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/segb200.h but not exported"
    assert declared == set(built.SYMBOLS), declared ^ set(built.SYMBOLS)
    assert lib.segb200_version() == 100
    assert [lib.segb200_conv_kblock(c) for c in (16, 24, 32, 48, 64, 728)] == [16, 16, 32, 32, 64, 64]


def test_argument_errors_without_gpu(built):
    import ctypes as C
    lib = built.load()
    assert lib.segb200_conv_gemm(None, None) < 0
    assert b"null" in lib.segb200_last_error()
    a = built.DwArgs()
    a.x = a.wgt = a.y = 16
    a.c, a.x_ld, a.y_ld, a.dtype = 60, 60, 60, 0
    assert lib.segb200_dwconv3x3(C.byref(a), None) == -4


def test_no_cpu_fallback(built):
    from segmentron_b200 import ops
    from segmentron_b200.engine import DeepLabV3PlusB200
    x = torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.dwconv3x3(x, x, x)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            DeepLabV3PlusB200({})


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "segmentron_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_bn_fold_and_weight_packing():
    from segmentron_b200 import fold
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 24, 5, 7, generator=g)
    w, b, m = torch.rand(24, generator=g) + 0.5, torch.randn(24, generator=g), torch.randn(24, generator=g)
    v = torch.rand(24, generator=g) + 0.5
    s, t = fold.bn_fold(w, b, m, v, 1e-3)
    ref = F.batch_norm(x, m, v, w, b, False, 0.0, 1e-3)
    assert torch.allclose(x * s[None, :, None, None] + t[None, :, None, None], ref, atol=1e-5)
    wt = torch.randn(19, 40, 3, 3, generator=g)
    pk = fold.pack_conv_weight(wt, torch.float32)
    assert pk.shape == (24, 9, 64)
    assert torch.equal(pk[:19, 4, :40], wt[:, :, 1, 1]) and float(pk[19:].abs().sum()) == 0 and float(pk[:, :, 40:].abs().sum()) == 0
    dw = torch.randn(16, 1, 3, 3, generator=g)
    pdw = fold.pack_dw_weight(dw, torch.full((16,), 2.0))
    assert torch.allclose(pdw[5], dw[:, 0, 1, 2] * 2)


@pytest.mark.parametrize("k,pad", [(3, 1), (7, 3), (5, 2)])
def test_stem_space_to_depth_is_exact(k, pad):
    """stride-2 kxk conv == stride-1 TxT conv over the space-to-depth tensor (what pack_s2d + conv_gemm run)."""
    from segmentron_b200 import fold
    g = torch.Generator().manual_seed(1)
    n, c, h, w = 2, 3, 13, 18
    x = torch.randn(n, c, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(8, c, k, k, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, wt, None, 2, pad)
    pk, T, pad2, ld = fold.pack_stem_s2d(wt, pad, torch.float64)
    hs, ws = (h + 1) // 2, (w + 1) // 2
    xp = F.pad(x, (0, 2 * ws - w, 0, 2 * hs - h))
    s2d = torch.zeros(n, ld, hs, ws, dtype=torch.float64)
    for py in range(2):
        for px in range(2):
            s2d[:, (py * 2 + px) * c:(py * 2 + px + 1) * c] = xp[:, :, py::2, px::2]
    w2 = pk.reshape(8, T, T, ld).permute(0, 3, 1, 2)
    ho, wo = ref.shape[2:]
    full = F.conv2d(F.pad(s2d, (pad2, T, pad2, T)), w2)
    assert torch.allclose(full[:, :, :ho, :wo], ref, atol=1e-5)   # packing goes through fp32


def test_param_store_index_tables_match_the_standalone_packers():
    """ParamStore regenerates every 16-bit GEMM operand from the fp32 masters with ONE index-table gather per step; the tables
    must reproduce fold.pack_conv_weight (forward), train_ops.pack_dgrad_weight (data gradient: transposed, taps reversed),
    fold.pack_stem_s2d (space-to-depth stem) and the depthwise [9][C] / flipped layouts that the per-kernel GPU tests use."""
    import torch
    from segmentron_b200 import fold, ops, train_ops as T
    from segmentron_b200.train import ParamStore
    ops._PLAN_DRY_RUN = True
    try:
        g = torch.Generator().manual_seed(0)
        sd = {"encoder.conv1.weight": torch.randn(64, 3, 7, 7, generator=g),
              "encoder.a.weight": torch.randn(48, 304, 1, 1, generator=g),
              "encoder.b.weight": torch.randn(128, 64, 3, 3, generator=g),
              "head.c.depthwise.weight": torch.randn(72, 1, 3, 3, generator=g),
              "head.d.weight": torch.randn(19, 256, 1, 1, generator=g), "head.d.bias": torch.randn(19, generator=g)}
        S = ParamStore(sd, "cpu", torch.float32, stem="encoder.conv1.weight")
        idx16, idx32 = S.idx16.long(), S.idx32.long()
        S.w16.copy_(torch.where(idx16 >= 0, S.master[idx16.clamp(min=0)], torch.zeros(())))
        S.w32.copy_(torch.where(idx32 >= 0, S.master[idx32.clamp(min=0)], torch.zeros(())))
        for k in ("encoder.a.weight", "encoder.b.weight", "head.d.weight"):
            assert torch.equal(S.packed(k, "fwd"), fold.pack_conv_weight(sd[k], torch.float32)), k
            assert torch.equal(S.packed(k, "dgrad"), T.pack_dgrad_weight(sd[k], torch.float32)), k
        stem, t, pad2, ld = fold.pack_stem_s2d(sd["encoder.conv1.weight"], 3, torch.float32)
        assert (S.pk["encoder.conv1.weight"]["T"], S.pk["encoder.conv1.weight"]["pad"], ld) == (t, pad2, 16)
        assert torch.equal(S.packed("encoder.conv1.weight", "fwd"), stem)
        dw = sd["head.c.depthwise.weight"].reshape(72, 9)
        assert torch.equal(S.packed("head.c.depthwise.weight", "fwd"), dw.t())
        assert torch.equal(S.packed("head.c.depthwise.weight", "flip"), dw.flip(1).t())
        out = S.state_dict()
        assert all(torch.equal(out[k], sd[k]) for k in sd)
    finally:
        ops._PLAN_DRY_RUN = False


def test_composite_dropins_training_wiring(monkeypatch):
    """InvertedResidual / _ASPP in training mode are compositions of the differentiable units (train_modules.py).  The units
    themselves are GPU kernels (tested with -m gpu); here they are substituted by their torch definitions so that the WIRING --
    unit order, activations, leading ReLU, skip, concat order, pooling branch, state_dict names -- is checked against the oracle
    (basic.py:139-163, module.py:62-77) in fp32 on the CPU: output, input gradient and every parameter gradient."""
    import torch
    import torch.nn.functional as F
    from oracle import segref as R
    from segmentron_b200 import modules as M, train_modules as TM

    def unit(xh, conv, bn, act, pre_relu=False):
        t = xh.permute(0, 3, 1, 2)
        if pre_relu:
            t = F.relu(t)
        t = F.conv2d(t, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        t = F.batch_norm(t, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
        t = F.relu(t) if act == "relu" else F.relu6(t) if act == "relu6" else t
        return t.permute(0, 2, 3, 1)

    class Gap:
        apply = staticmethod(lambda t: t.mean((1, 2), keepdim=True))

    class Bcast:
        apply = staticmethod(lambda v, h, w: v.expand(-1, h, w, -1))

    monkeypatch.setattr(M, "_train_conv_bn_act", unit)
    monkeypatch.setattr(M, "_train_enter", lambda x, m: x.permute(0, 2, 3, 1).contiguous())
    monkeypatch.setattr(TM, "GlobalAvgPoolFunction", Gap)
    monkeypatch.setattr(TM, "BroadcastFunction", Bcast)

    def aspp(P, t):
        P.dropout_masks["m.dropout"] = torch.ones(1)
        return R.aspp(P, t, "m", 32, 16)

    cases = {
        "ir_skip": (lambda: M.InvertedResidual(16, 16, 1, 6), lambda P, t: R.inverted_residual(P, t, "m", 16, 1, 6)),
        "ir_s2": (lambda: M.InvertedResidual(16, 24, 2, 6), lambda P, t: R.inverted_residual(P, t, "m", 24, 2, 6)),
        "ir_t1_d2": (lambda: M.InvertedResidual(16, 8, 1, 1, dilation=2), lambda P, t: R.inverted_residual(P, t, "m", 8, 1, 1, 2)),
        "aspp": (lambda: M._ASPP(16, 32, output_stride=16), aspp),
    }
    for name, (make, oracle_fn) in cases.items():
        x = torch.randn(2, 16, 21, 19, generator=torch.Generator().manual_seed(5))
        P = R.Params(7)
        with torch.no_grad():
            oracle_fn(P, x)
        keys = [k for k in P.t if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
        for k in keys:
            P.t[k] = P.t[k].detach().requires_grad_(True)
        sd0 = {k[2:]: v.detach().clone() for k, v in P.state_dict().items()}
        P.training = True
        xr = x.clone().requires_grad_(True)
        ref = oracle_fn(P, xr)
        dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(6))
        ref.backward(dy)
        m = make()
        m.load_state_dict(sd0, strict=True)
        m.train()
        if hasattr(m, "dropout"):
            m.dropout.p = 0.0
        xg = x.clone().requires_grad_(True)
        y = m(xg)
        y.backward(dy)
        assert torch.allclose(y, ref, atol=1e-5, rtol=1e-5), name
        assert torch.allclose(xg.grad, xr.grad, atol=1e-5, rtol=1e-4), name
        for k in keys:
            g = dict(m.named_parameters())[k[2:]].grad
            assert g is not None and torch.allclose(g, P.t[k].grad, atol=1e-4, rtol=1e-4), (name, k)


def test_metric_dropin_api_without_gpu(built):
    """segmentron_b200.metric.SegmentationMetric mirrors score.py:11-81 (constructor, update / get / reset, attribute names); with
    no update it returns the reference's (0.0, 0.0); CPU tensors raise (no host implementation in the product); the C entry
    points reject null pointers and class counts beyond 64 without touching a GPU."""
    import ctypes as C
    import torch
    from segmentron_b200 import lib as L
    from segmentron_b200.metric import SegmentationMetric
    m = SegmentationMetric(19, False)
    assert m.get() == (0.0, 0.0) and (m.total_correct, m.total_label) == (0, 0)
    assert tuple(m.total_inter.shape) == tuple(m.total_union.shape) == (19,)
    pix, miou, iou = m.get(return_category_iou=True)
    assert iou.shape == (19,)
    with pytest.raises(RuntimeError, match="not implemented on the CPU"):
        m.update(torch.zeros(1, 19, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        SegmentationMetric(65, False)
    m.reset()
    lib = L.load()
    buf = (C.c_ulonglong * 64)()
    assert lib.segb200_seg_metric(None, L.F32, None, 1, 19, 4, 4, None, None) < 0
    assert lib.segb200_seg_metric(buf, L.F32, buf, 1, 65, 4, 4, buf, None) < 0
    assert b"nclass" in lib.segb200_last_error()
    assert lib.segb200_seg_metric_lowres(buf, L.F32, 24, 4, 4, 1, L.F32, buf, 1, 19, 8, 8, buf, None) < 0     # 16-bit logits only
    assert lib.segb200_seg_metric_lowres(buf, L.BF16, 16, 4, 4, 1, L.F32, buf, 1, 19, 8, 8, buf, None) < 0    # x_ld < round_up(19, 8)
    assert lib.segb200_seg_metric_accumulate(None, 19, None, None, None, None) < 0


def test_evaluate_driver_size_rules_and_errors(built):
    """segmentron_b200.evaluate: the host-side size rules equal the oracle's (which is pinned to SegBaseModel.evaluate,
    segbase.py:53-68,93, by tests/golden/evaluate_cases.pt) over a sweep of shapes / scales / crops; CPU images raise; the C entry
    points validate their arguments without a GPU."""
    import ctypes as C
    import itertools
    import torch
    from oracle import evalref as E
    from segmentron_b200 import evaluate as V, lib as L
    for (h, w), scale in itertools.product([(1024, 2048), (1025, 2049), (57, 31), (33, 65), (480, 480), (769, 769)],
                                           [0.5, 0.75, 1.0, 1.25, 1.5, 1.75, 2.0]):
        assert V.scaled_size(h, w, scale) == E.scaled_size(h, w, scale)
        hh, ww = V.scaled_size(h, w, scale)
        for crop in (None, (max(h, w), max(h, w)), (h, w), (h + 7, w + 64)):
            assert V.padded_size(hh, ww, crop, scale) == E.padded_size(hh, ww, crop, scale)
    assert V._to_tuple(769) == (769, 769) and V._to_tuple([512, 1024]) == (512, 1024)
    with pytest.raises(RuntimeError, match="not implemented on the CPU"):
        V.evaluate(lambda x: x, torch.zeros(1, 3, 8, 8))
    lib = L.load()
    buf = (C.c_float * 16)()
    assert lib.segb200_eval_prepare(None, None, 1, 3, 8, 8, 8, 8, 8, 8, 0, None) < 0
    assert lib.segb200_eval_prepare(buf, buf, 1, 3, 8, 8, 12, 12, 8, 8, 0, None) < 0          # padded size smaller than the resize
    assert lib.segb200_eval_accumulate(buf, buf, 7, 1, 3, 8, 8, 8, 8, 8, 8, 0, 0, None) < 0   # bad dtype
    assert lib.segb200_eval_accumulate(buf, buf, L.F32, 1, 3, 8, 8, 9, 8, 8, 8, 0, 0, None) < 0


def test_ctypes_signatures_match_the_header():
    """lib.SYMBOLS (the ctypes argtypes) against the prototypes of include/segb200.h, parameter by parameter: pointer / int /
    long long / float / double classes must agree (every .cu includes the header, so the compiler already checks the definitions
    against it; this closes the loop on the Python side)."""
    import ctypes as C
    import re
    from segmentron_b200 import lib as L
    src = open(os.path.join(ROOT, "include", "segb200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = dict(re.findall(r"\b(?:int|const char\s*\*)\s+(segb200_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert set(protos) == set(L.SYMBOLS), set(protos) ^ set(L.SYMBOLS)

    def cls_of_c(param):
        p = " ".join(param.split())
        if p in ("void", ""):
            return None
        if "*" in p:
            return "ptr"
        if "long long" in p:
            return "ll"
        if "double" in p:
            return "f64"
        if "float" in p:
            return "f32"
        assert "int" in p, p
        return "int"

    def cls_of_ctypes(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int: "int", C.c_int32: "int", C.c_longlong: "ll", C.c_float: "f32", C.c_double: "f64"}[t]
    for name, params in protos.items():
        want = [c for c in (cls_of_c(p) for p in params.split(",")) if c is not None]
        got = [cls_of_ctypes(t) for t in L.SYMBOLS[name][1]]
        assert want == got, (name, want, got)


def test_metric_wrapper_call_sites_match_the_signatures(monkeypatch):
    """segmentron_b200.metric's three C calls (update, update_lowres, accumulate): argument tuples checked against lib.SYMBOLS by a
    recording stand-in library (the wrappers themselves need CUDA tensors, so the device check is bypassed for this test only)."""
    import ctypes as C
    import torch
    from segmentron_b200 import lib as L, metric as MM, ops
    calls = []

    class Rec:
        def __getattr__(self, name):
            sig = L.SYMBOLS[name][1]

            def fn(*args):
                assert len(args) == len(sig), (name, len(args), len(sig))
                for pos, (a, t) in enumerate(zip(args, sig)):
                    if t is C.c_void_p:
                        assert a is None or isinstance(a, C.c_void_p), (name, pos, type(a))
                    else:
                        assert isinstance(a, (int, float)) and not isinstance(a, bool), (name, pos, a)
                calls.append(name)
                return 0
            return fn
    monkeypatch.setattr(L, "load", lambda: Rec())
    monkeypatch.setattr(ops, "_PLAN_DRY_RUN", True)
    monkeypatch.setattr(MM, "_stream", lambda: None)

    def check(self, t, labels, what):
        if self._counts is None:
            self._alloc(t.device)
        return labels.long().contiguous()
    monkeypatch.setattr(MM.SegmentationMetric, "_check", check)
    m = MM.SegmentationMetric(19, False)
    m.update(torch.zeros(2, 19, 8, 8), torch.zeros(2, 8, 8, dtype=torch.long))
    m.update_lowres(torch.zeros(2, 4, 4, 24)[..., :19], torch.zeros(2, 16, 16, dtype=torch.long), out_dtype=torch.float32)
    assert calls == ["segb200_seg_metric", "segb200_seg_metric_accumulate", "segb200_seg_metric_lowres", "segb200_seg_metric_accumulate"]


def test_input_normalize_is_torchvision_bit_for_bit(built):
    """segmentron_b200.data.normalize == transforms.ToTensor() + transforms.Normalize (the reference's input transform,
    tools/train.py:36-39): the kernel's formula, transcribed with numpy float32 IEEE operations in the kernel's order, equals
    torchvision on random uint8 images EXACTLY (all 256 byte values x the Cityscapes mean/std); CPU tensors raise."""
    import numpy as np
    import torch
    from torchvision import transforms
    from segmentron_b200 import data as D
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (37, 53, 3), generator=g, dtype=torch.uint8)
    img[0, :, 0] = torch.arange(53) % 256
    img[1:6, :51, 1] = torch.arange(255).reshape(5, 51).to(torch.uint8)
    for mean, std in (([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]), ([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])):
        from PIL import Image
        ref = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean, std)])(Image.fromarray(img.numpy()))
        u = img.numpy().astype(np.float32)
        got = ((u / np.float32(255.0)) - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)      # kernel order, fp32 IEEE
        assert got.dtype == np.float32 and np.array_equal(got.transpose(2, 0, 1), ref.numpy())
    with pytest.raises(RuntimeError, match="not implemented on the CPU"):
        D.normalize(img[None], [0.5] * 3, [0.5] * 3)


def test_launcher_local_rank_shim_keeps_the_command_line_parseable(monkeypatch):
    """under torchrun the launcher injects --local_rank=N.  It must not split an option from its value
    nor land in the reference's `opts` REMAINDER (utils/options.py:25-26): the result is parsed with the reference's own parser when
    the staged copy is present, else with the same argparse definition."""
    import argparse
    import importlib
    from segmentron_b200 import launch
    monkeypatch.setenv("LOCAL_RANK", "3")
    argv = ["tools/eval.py", "--config-file", "configs/x.yaml", "--log-iter", "5", "TEST.BATCH_SIZE", "2"]
    monkeypatch.setattr(sys, "argv", list(argv))
    launch._shims()
    assert sys.argv[0] == "tools/eval.py" and sys.argv[1] == "--local_rank=3"
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "segmentron")):
        spec = importlib.util.spec_from_file_location("ref_options", os.path.join(ref, "segmentron", "utils", "options.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        args = mod.parse_args()
    else:
        p = argparse.ArgumentParser()
        p.add_argument("--config-file")
        p.add_argument("--local_rank", type=int, default=0)
        p.add_argument("--log-iter", type=int, default=10)
        p.add_argument("opts", nargs=argparse.REMAINDER)
        args = p.parse_args()
    assert args.local_rank == 3 and args.config_file == "configs/x.yaml" and args.log_iter == 5
    assert args.opts == ["TEST.BATCH_SIZE", "2"]
    # an explicit --local-rank (what torch.distributed.run passes) is rewritten, not duplicated
    monkeypatch.setattr(sys, "argv", ["tools/train.py", "--local-rank=1", "--config-file", "c.yaml"])
    launch._shims()
    assert sys.argv == ["tools/train.py", "--local_rank=1", "--config-file", "c.yaml"]


def test_tuning_knobs_and_env_plumbing(built):
    """segb200_set_option: every A/B knob named in the sources / DESIGN.md is accepted and an unknown name is an error (no GPU needed:
    the knobs are host-side state); the defaults the measurements selected stay what the docs say (the opt-in experiments are OFF);
    SEGB200_OPTS applies knobs at library load time in a fresh process and rejects unknown ones loudly."""
    import subprocess
    lib = built.load()
    src = open(os.path.join(ROOT, "segmentron_b200", "csrc", "conv_gemm.cu")).read()
    knobs = re.findall(r'!strcmp\(name, "([a-z0-9_]+)"\)', src)
    assert {"gemm_dual", "gemm_dual_subk", "pdl", "dw_cols2", "dw_cw5", "bilinear_out_v1", "gemm_kgroup", "gemm_2cta"} <= set(knobs)
    assert lib.segb200_set_option(b"no_such_knob", 1) != 0 and b"unknown option" in lib.segb200_last_error()
    # defaults: experiments measured slower are off, the kept ones on (sources are the single place these live)
    for f, pat in (("conv_gemm.cu", r"static int g_dual = 0;"), ("conv_gemm.cu", r"static int g_2cta = 0;"), ("api.cu", r"static int g_pdl = 0;"),
                   ("dwconv.cu", r"static int g_dw_cols2 = 1;"), ("dwconv.cu", r"static int g_dw_cw5 = 0;"), ("misc.cu", r"static int g_bilinear_out_v1 = 0;")):
        assert re.search(pat, open(os.path.join(ROOT, "segmentron_b200", "csrc", f)).read()), (f, pat)
    code = "import sys; sys.path.insert(0, %r); from segmentron_b200 import lib; lib.load(); print('loaded')" % ROOT
    ok = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SEGB200_OPTS="gemm_dual=0,dw_cols2=1"), capture_output=True, text=True)
    assert ok.returncode == 0 and "loaded" in ok.stdout, ok.stderr[-400:]
    bad = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SEGB200_OPTS="no_such_knob=1"), capture_output=True, text=True)
    assert bad.returncode != 0 and "unknown option" in bad.stderr
