"""Run an UNMODIFIED reference script (tools/train.py, tools/eval.py, tools/demo.py) on the B200 modules:

    PYTHONPATH=/path/to/SegmenTron python -m segmentron_b200.launch [--accelerate] /path/to/SegmenTron/tools/eval.py \\
        --config-file configs/cityscapes_deeplabv3_plus.yaml

Installs, outside the reference tree, the compatibility shims the reference needs on a current stack
(``np.int``, a stub ``thop``, ``--local-rank`` -> ``--local_rank``), imports ``segmentron``, installs the
``segmentron._C`` shim (``c_shim``: the four criss-cross functions of vision.cpp:6-11 over the C ABI, which re-enables CCNet), rebinds the
L1 classes (``patch.install``) and, with ``--accelerate``, wraps ``get_segmentation_model`` so DeepLabV3+ models run the
fused whole-model plan; then executes the script as ``__main__``.
"""
import os
import runpy
import sys
import types


def _shims():
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int                                               # backbones/hrnet.py:291
    if "thop" not in sys.modules:                                  # utils/visualize.py:8 imports it unconditionally
        thop = types.ModuleType("thop")

        def profile(*a, **k):
            raise RuntimeError("thop is not installed (segb200 stub)")
        thop.profile = profile
        sys.modules["thop"] = thop
    sys.argv = [a.replace("--local-rank", "--local_rank") if a.startswith("--local-rank") else a for a in sys.argv]
    if "LOCAL_RANK" in os.environ and not any(a.startswith("--local_rank") for a in sys.argv[1:]):
        # index 1 = directly after the script name: anywhere later could split an option from its value, and anything after the
        # first positional lands in the reference's `opts` REMAINDER (utils/options.py:25-26)
        sys.argv.insert(1, f"--local_rank={os.environ['LOCAL_RANK']}")


def main():
    argv = sys.argv[1:]
    accel = False
    if argv and argv[0] == "--accelerate":
        accel, argv = True, argv[1:]
    if not argv:
        print(__doc__)
        return 2
    script = argv[0]
    sys.argv = [script] + argv[1:]
    _shims()
    import segmentron  # noqa: F401
    from . import c_shim, patch
    # the reference's only native module: `segmentron._C` cannot be built on a current torch; install the
    # C-ABI shim BEFORE anything imports segmentron.modules.cc_attention, and let CCNet register itself (models/__init__.py:11)
    c_shim.install(register_ccnet=True)
    n = patch.install()
    print(f"[segb200] rebound {n} class references in segmentron.* namespaces", file=sys.stderr)
    if accel:
        from segmentron.models import model_zoo
        orig = model_zoo.get_segmentation_model

        def get_segmentation_model(*a, **k):
            model = orig(*a, **k)
            try:
                return patch.accelerate(model)
            except RuntimeError as e:
                print(f"[segb200] whole-model plan unavailable ({e}); using module-level drop-ins", file=sys.stderr)
                return model
        model_zoo.get_segmentation_model = get_segmentation_model
        for modname, mod in list(sys.modules.items()):
            if modname.startswith("segmentron") and getattr(mod, "get_segmentation_model", None) is orig:
                mod.get_segmentation_model = get_segmentation_model
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
