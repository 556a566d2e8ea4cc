"""``segmentron._C`` drop-in: the four functions of the reference's pybind module (segmentron/modules/csrc/vision.cpp:6-11) over
the C ABI (``segb200_ca_forward`` / ``_backward`` / ``_map_forward`` / ``_map_backward``, csrc/ca_nchw.cu).

The reference's extension cannot be built on a current torch (``THC/THC.h`` is gone), so
``segmentron/modules/cc_attention.py`` -- which does ``from segmentron import _C`` at import time -- cannot even be imported and
CCNet is commented out of ``segmentron/models/__init__.py:11``.  ``install()`` puts this module object into ``sys.modules`` as
``segmentron._C`` BEFORE that import, after which the reference's own autograd Functions ``_CAWeight`` / ``_CAMap``
(cc_attention.py:11-45), its ``CrissCrossAttention`` and ``segmentron.models.ccnet`` work unchanged.

Semantics of the reference's functions (csrc/criss_cross_attention/ca.h:25-72, ca_cuda.cu:188-312): inputs are made contiguous,
outputs are fresh tensors of the input dtype, work is enqueued on the current CUDA stream, a non-CUDA tensor raises
``RuntimeError("Not implemented on the CPU")``.
"""
import ctypes as C
import sys
import types

import torch

from . import lib as L

_DT = {torch.float32: L.F32, torch.float16: L.F16, torch.bfloat16: L.BF16}


def _prep(name, *tensors):
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU")            # ca.h:34,46,58,70
        if t.dim() != 4:
            raise RuntimeError(f"{name}: expected 4-D NCHW tensors")
    dt = tensors[0].dtype
    if dt not in _DT:
        raise RuntimeError(f"{name}: unsupported dtype {dt}")
    if any(t.dtype != dt or t.device != tensors[0].device for t in tensors):
        raise RuntimeError(f"{name}: all tensors must share dtype and device")
    return [t.detach().contiguous() for t in tensors], _DT[dt]           # ca_cuda.cu:205-207 makes them contiguous too


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


def _check_pair(name, a, b, wshape=None):
    n, c, h, w = b.shape
    if wshape is not None:
        if tuple(a.shape) != (n, h + w - 1, h, w):
            raise RuntimeError(f"{name}: weight must be [N, H+W-1, H, W] = {(n, h + w - 1, h, w)}, got {tuple(a.shape)}")
    elif a.shape != b.shape:
        raise RuntimeError(f"{name}: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    return n, c, h, w


def ca_forward(t, f):
    """energy of every pixel against its row and column: [N,C,H,W] x [N,C,H,W] -> [N,H+W-1,H,W]"""
    (t, f), dt = _prep("ca_forward", t, f)
    n, c, h, w = _check_pair("ca_forward", t, f)
    weight = torch.empty(n, h + w - 1, h, w, dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        L.check(L.load().segb200_ca_forward(_p(t), _p(f), _p(weight), n, c, h, w, dt, _stream(t)), "ca_forward")
    return weight


def ca_backward(dw, t, f):
    (dw, t, f), dt = _prep("ca_backward", dw, t, f)
    n, c, h, w = _check_pair("ca_backward", t, f)
    _check_pair("ca_backward", dw, t, wshape=True)
    dt_, df = torch.empty_like(t), torch.empty_like(f)
    with torch.cuda.device(t.device):
        L.check(L.load().segb200_ca_backward(_p(dw), _p(t), _p(f), _p(dt_), _p(df), n, c, h, w, dt, _stream(t)), "ca_backward")
    return dt_, df


def ca_map_forward(weight, g):
    (weight, g), dt = _prep("ca_map_forward", weight, g)
    n, c, h, w = _check_pair("ca_map_forward", weight, g, wshape=True)
    out = torch.empty_like(g)
    with torch.cuda.device(g.device):
        L.check(L.load().segb200_ca_map_forward(_p(weight), _p(g), _p(out), n, c, h, w, dt, _stream(g)), "ca_map_forward")
    return out


def ca_map_backward(dout, weight, g):
    (dout, weight, g), dt = _prep("ca_map_backward", dout, weight, g)
    n, c, h, w = _check_pair("ca_map_backward", dout, g)
    _check_pair("ca_map_backward", weight, g, wshape=True)
    dw, dg = torch.empty_like(weight), torch.empty_like(g)
    with torch.cuda.device(g.device):
        L.check(L.load().segb200_ca_map_backward(_p(dout), _p(weight), _p(g), _p(dw), _p(dg), n, c, h, w, dt, _stream(g)),
                "ca_map_backward")
    return dw, dg


def make_module():
    m = types.ModuleType("segmentron._C")
    m.__doc__ = "segb200 drop-in for the reference's pybind module (vision.cpp:6-11)"
    m.ca_forward, m.ca_backward, m.ca_map_forward, m.ca_map_backward = ca_forward, ca_backward, ca_map_forward, ca_map_backward
    return m


def install(register_ccnet=True):
    """Make ``from segmentron import _C`` resolve to this shim; optionally import ``segmentron.models.ccnet`` so that the
    ``CCNet`` model registers itself (models/__init__.py:11 leaves the import commented out).  Returns the module object."""
    import segmentron
    if "segmentron._C" in sys.modules and getattr(sys.modules["segmentron._C"], "ca_forward", None) is not None \
            and not getattr(sys.modules["segmentron._C"], "__doc__", "").startswith("segb200"):
        return sys.modules["segmentron._C"]                # a real extension is present: leave it alone
    m = make_module()
    sys.modules["segmentron._C"] = m
    segmentron._C = m
    if register_ccnet:
        import segmentron.modules.cc_attention  # noqa: F401
        from segmentron.models.model_zoo import MODEL_REGISTRY
        if "CCNet" not in MODEL_REGISTRY.get_list():
            import segmentron.models.ccnet  # noqa: F401
    return m
