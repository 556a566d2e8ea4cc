"""Installing the B200 modules behind the reference's own registry / YAML / tools.

Two mechanisms, both leave /root/reference untouched:

* ``install()``            -- after ``import segmentron``: rebind ``SeparableConv2d``, ``_ConvBNReLU`` ... in EVERY
  ``segmentron.*`` module namespace that holds the reference class (``from ..modules import X`` copies the name into each
  consumer's globals, so patching ``segmentron.modules`` alone would swap 3 of 68 instances).  Models built afterwards
  by ``get_segmentation_model()`` are made of the drop-in modules; ``state_dict`` keys are unchanged.
* ``convert_to_b200(model)`` -- recursive in-place swap on an already built model (the idiom of
  ``FrozenBatchNorm2d.convert_frozen_batchnorm``, modules/batch_norm.py:74-104).  Parameters/buffers are shared, not
  copied.
* ``accelerate(model)``    -- whole-model fast path: replace ``DeepLabV3Plus.forward`` by the fused execution plan
  (``engine.DeepLabV3PlusB200``) built from ``model.state_dict()``; eval mode only.
"""
import sys

import torch
import torch.nn as nn

from . import modules as M


def install(verbose=False):
    """Rebind the reference class names to the drop-in classes in all loaded ``segmentron.*`` namespaces."""
    if "segmentron" not in sys.modules:
        import segmentron  # noqa: F401  (the reference must be importable: PYTHONPATH=/path/to/SegmenTron)
    ref_mods = [sys.modules["segmentron.modules.basic"], sys.modules["segmentron.modules.module"]]
    if "segmentron.modules.cc_attention" in sys.modules:      # only importable with a segmentron._C
        ref_mods.append(sys.modules["segmentron.modules.cc_attention"])
    originals = {}
    for name in M.REPLACEMENTS:
        for rm in ref_mods:
            if hasattr(rm, name) and getattr(rm, name).__module__.startswith("segmentron."):
                originals[name] = getattr(rm, name)
    count = 0
    for modname, mod in list(sys.modules.items()):
        if mod is None or not modname.startswith("segmentron"):
            continue
        for name, orig in originals.items():
            if getattr(mod, name, None) is orig:
                setattr(mod, name, M.REPLACEMENTS[name])
                count += 1
                if verbose:
                    print(f"[segb200] {modname}.{name} -> segmentron_b200.modules.{name}")
    return count


def install_metric(verbose=False):
    """Opt-in: rebind ``SegmentationMetric`` (segmentron/utils/score.py:11) to the device-resident drop-in
    (``segmentron_b200.metric.SegmentationMetric``: one kernel per update, no host synchronisation before ``get()``) in every
    loaded ``segmentron.*`` namespace.  Call before ``tools/eval.py`` / ``tools/train.py`` import the name."""
    from .metric import SegmentationMetric
    if "segmentron.utils.score" not in sys.modules:
        import segmentron.utils.score  # noqa: F401
    orig = sys.modules["segmentron.utils.score"].SegmentationMetric
    count = 0
    for modname, mod in list(sys.modules.items()):
        if mod is not None and modname.startswith("segmentron") and getattr(mod, "SegmentationMetric", None) is orig:
            mod.SegmentationMetric = SegmentationMetric
            count += 1
            if verbose:
                print(f"[segb200] {modname}.SegmentationMetric -> segmentron_b200.metric.SegmentationMetric")
    return count


def install_evaluate():
    """Opt-in: route ``SegBaseModel.evaluate`` (segmentron/models/segbase.py:44-79, called by tools/eval.py for every batch) through
    ``segmentron_b200.evaluate.evaluate`` -- same scales / flip / crop rules read from the reference's cfg.TEST, one model call per
    scale on the stacked [image; mirrored image] batch, two kernels instead of nine torch ops per scale."""
    from segmentron.config import cfg
    from segmentron.models.segbase import SegBaseModel
    from .evaluate import evaluate

    def _evaluate(self, image):
        return evaluate(self.forward, image, cfg.TEST.SCALES, cfg.TEST.FLIP, cfg.TEST.CROP_SIZE)
    SegBaseModel.evaluate = _evaluate
    return SegBaseModel


def _adopt(cls, ref):
    """Build a drop-in instance that shares ``ref``'s sub-modules, parameters and buffers."""
    new = cls.__new__(cls)
    nn.Module.__init__(new)
    new._modules = ref._modules
    new._parameters = ref._parameters
    new._buffers = ref._buffers
    new.training = ref.training
    name = cls.__name__
    if name == "SeparableConv2d":
        new.relu_first = "relu" in ref.block._modules
        new._c_dw, new._c_pw = M._Cache(), M._Cache()
    elif name == "_ConvBNReLU":
        new._act = "relu6" if isinstance(ref.relu, nn.ReLU6) else "relu"
        new._cache = M._Cache()
    elif name == "_ConvBN":
        new._cache = M._Cache()
    elif name == "InvertedResidual":
        new.use_res_connect = ref.use_res_connect
        new._cache = M._Cache()
    elif name == "_ASPP":
        new._c0, new._cp, new._cproj = M._Cache(), M._Cache(), M._Cache()
    elif name in ("CrissCrossAttention", "PAM_Module"):
        new._cache = M._Cache()
    elif name == "PyramidPooling":
        def _size(p):
            s = p.output_size
            return s[0] if isinstance(s, (tuple, list)) else s
        new.sizes = tuple(_size(p) for p in ref.avgpools)
    return new


def convert_to_b200(module):
    """Recursively replace reference L1 modules inside ``module`` by their B200 drop-ins (in place); returns the
    (possibly new) root."""
    for name, child in list(module.named_children()):
        module._modules[name] = convert_to_b200(child)
    cls_name = type(module).__name__
    if cls_name in M.REPLACEMENTS and type(module).__module__.startswith("segmentron."):
        return _adopt(M.REPLACEMENTS[cls_name], module)
    return module


def accelerate(model, dtype=torch.bfloat16, cuda_graph=True, want_argmax=False):
    """Route ``model.forward`` of a reference ``DeepLabV3Plus`` through the fused whole-model plan.

    Reads everything from the model itself: weights from ``state_dict()``, backbone / head switches from its
    attributes, the encoder's BN eps (possibly mutated by tools/eval.py:50-53) at call time."""
    from .engine import DeepLabV3PlusB200
    if type(model).__name__ != "DeepLabV3Plus":
        raise RuntimeError(f"segb200.accelerate: no whole-model plan for {type(model).__name__} yet")
    if getattr(model, "aux", False):
        raise RuntimeError("segb200.accelerate: aux head not supported by the whole-model plan")
    try:
        from segmentron.config import cfg
        output_stride = cfg.MODEL.OUTPUT_STRIDE
    except Exception:
        output_stride = 16
    state = {}

    def forward(x):
        if model.training:
            raise RuntimeError("segb200: the whole-model plan is inference-only; call model.eval()")
        eps = next(m.eps for m in model.encoder.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm))
        key = (eps, tuple((p.data_ptr(), p._version) for p in model.parameters()))
        if state.get("key") != key:
            state["eng"] = DeepLabV3PlusB200(model.state_dict(), backbone=model.backbone, nclass=model.nclass,
                                             output_stride=output_stride, eps_encoder=eps,
                                             use_aspp=getattr(model.head, "use_aspp", True),
                                             use_decoder=getattr(model.head, "use_decoder", True), dtype=dtype,
                                             out_dtype=x.dtype if x.dtype in (torch.float16, torch.bfloat16, torch.float32) else dtype,
                                             cuda_graph=cuda_graph, want_argmax=want_argmax)
            state["key"] = key
        return (state["eng"](x).clone(),)

    model.forward = forward
    model._segb200_state = state
    return model
