"""Harness around the UNMODIFIED reference staged under ``baseline/_ref`` (``tools/make_baseline_ref.py``).

Test / benchmark infrastructure, not product code: the product (``segmentron_b200``) never imports this file.  It provides

* ``enter()``                       -- put the staged reference (and the ``thop`` stub) on ``sys.path`` and apply the in-process
                                       compatibility shims of SURVEY.md App. B (``np.int``);
* ``build_model(yaml, ...)``        -- ``get_segmentation_model()`` (segmentron/models/model_zoo.py:17-24) for one of the reference's
                                       own YAML configs, eval-mode BN eps override applied like tools/eval.py:50-53; several configs
                                       per process are possible because the frozen global ``cfg`` is snapshotted and restored;
* ``make_run_dir(path, ...)``       -- a writable run directory: ``tools/`` and ``configs/`` symlinked to the staged reference (the
                                       scripts derive their dataset root from their own location) and a synthetic
                                       Cityscapes tree (``leftImg8bit`` / ``gtFine`` PNGs with raw label ids, cityscapes.py:49-57);
* ``run_script(run_dir, script, argv, through_launch)`` -- run ``tools/train.py`` / ``tools/eval.py`` as a subprocess, either STOCK
                                       (plain reference, cuDNN) or through ``python -m segmentron_b200.launch`` (the drop-ins);
* ``parse_train_losses`` / ``parse_eval_result`` -- read the numbers the reference logs.
"""
import copy
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
STUBS = os.path.join(REF, "_stubs")

_cfg_snapshot = None


def available():
    return os.path.isdir(os.path.join(REF, "segmentron")) and os.path.isfile(os.path.join(REF, "MANIFEST.json"))


def enter():
    """Make ``import segmentron`` resolve to the staged reference in THIS process."""
    if not available():
        raise FileNotFoundError("baseline/_ref is not staged (run tools/make_baseline_ref.py in the build container)")
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int                                                     # backbones/hrnet.py:291
    for p in (STUBS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)


def _copy_cfg(c):
    new = type(c)()
    for k, v in c.items():
        dict.__setitem__(new, k, _copy_cfg(v) if isinstance(v, dict) else copy.deepcopy(v))
    return new


def _reset_cfg():
    """The reference's ``cfg`` is a process global that ``check_and_freeze`` prunes and freezes (config.py:99-103): keep a
    pristine copy from before the first freeze and put it back before every build."""
    global _cfg_snapshot
    from segmentron.config import cfg
    if _cfg_snapshot is None:
        if cfg.is_immutable():
            raise RuntimeError("cfg was frozen before ref_harness saw it")
        _cfg_snapshot = _copy_cfg(cfg)
    cfg.set_immutable(False)
    dict.clear(cfg)
    for k, v in _copy_cfg(_cfg_snapshot).items():
        dict.__setitem__(cfg, k, v)
    return cfg


def build_model(yaml_file, opts=(), phase="test", eval_eps=True, install_c=False):
    """The reference's own model for one of its YAML configs (random init, CPU, eval mode).  ``install_c``: install the
    segb200 ``segmentron._C`` shim first and register CCNet (models/__init__.py:11 leaves it commented out)."""
    enter()
    import segmentron  # noqa: F401
    if install_c:
        from segmentron_b200 import c_shim
        c_shim.install()
        import segmentron.models.ccnet  # noqa: F401
    import torch
    from segmentron.models.model_zoo import get_segmentation_model
    cfg = _reset_cfg()
    cfg.update_from_file(os.path.join(REF, "configs", yaml_file))
    cfg.update_from_list(list(opts))
    cfg.PHASE = phase
    cfg.check_and_freeze()
    model = get_segmentation_model()
    if phase == "test":
        model.eval()
    if eval_eps and hasattr(model, "encoder") and cfg.MODEL.BN_EPS_FOR_ENCODER:      # tools/eval.py:50-53
        for m in model.encoder.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.SyncBatchNorm)):
                m.eps = cfg.MODEL.BN_EPS_FOR_ENCODER
    return model


def randomise_bn(model, seed=0):
    """Default init makes eval-mode BatchNorm the identity and the attention gammas zero: give the running
    statistics / affine parameters seeded non-trivial values so that a comparison is not vacuous."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                n = m.num_features
                m.running_mean.copy_(0.1 * torch.randn(n, generator=g))
                m.running_var.copy_(0.5 + torch.rand(n, generator=g))
                m.weight.copy_(0.8 + 0.4 * torch.rand(n, generator=g))
                m.bias.copy_(0.1 * torch.randn(n, generator=g))
        for name, p in model.named_parameters():
            if name.endswith("gamma") and p.numel() == 1:
                p.fill_(0.5)
    return model


# ------------------------------------------------------------------------------------------------------------------
# running the reference's own scripts
# ------------------------------------------------------------------------------------------------------------------
_LABEL_IDS = [7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33, 0]      # 19 classes + one ignored id


def make_run_dir(path, n_train=8, n_val=4, h=128, w=256, seed=0):
    """Writable run directory for the unmodified scripts + a synthetic Cityscapes tree (blocky label maps, noisy images)."""
    import numpy as np
    from PIL import Image
    os.makedirs(os.path.join(path, "tools"), exist_ok=True)
    for s in ("train.py", "eval.py"):
        dst = os.path.join(path, "tools", s)
        if not os.path.lexists(dst):
            os.symlink(os.path.join(REF, "tools", s), dst)
    if not os.path.lexists(os.path.join(path, "configs")):
        os.symlink(os.path.join(REF, "configs"), os.path.join(path, "configs"))
    rng = np.random.RandomState(seed)
    for split, n in (("train", n_train), ("val", n_val)):
        idir = os.path.join(path, "datasets", "cityscapes", "leftImg8bit", split, "synth")
        mdir = os.path.join(path, "datasets", "cityscapes", "gtFine", split, "synth")
        os.makedirs(idir, exist_ok=True)
        os.makedirs(mdir, exist_ok=True)
        for i in range(n):
            blocks = rng.randint(0, len(_LABEL_IDS), size=((h + 15) // 16, (w + 15) // 16))
            lab = np.asarray(_LABEL_IDS, dtype=np.uint8)[blocks].repeat(16, 0).repeat(16, 1)[:h, :w]
            img = (lab[..., None].astype(np.int32) * np.array([5, 3, 7]) % 200 + rng.randint(0, 56, size=(h, w, 3))).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(idir, f"synth_{i:06d}_000019_leftImg8bit.png"))
            Image.fromarray(lab).save(os.path.join(mdir, f"synth_{i:06d}_000019_gtFine_labelIds.png"))
    return path


def run_script(run_dir, script, argv, through_launch, launch_flags=(), timeout=900, env_extra=None):
    """Run ``tools/<script>`` from ``run_dir``.  STOCK: the reference alone on PYTHONPATH, shims applied by a two-line
    bootstrap.  Through the launcher: ``python -m segmentron_b200.launch tools/<script> ...`` (the documented invocation)."""
    env = dict(os.environ)
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env.pop("RANK", None), env.pop("WORLD_SIZE", None), env.pop("LOCAL_RANK", None)
    if through_launch:
        env["PYTHONPATH"] = os.pathsep.join([REF, ROOT, env.get("PYTHONPATH", "")])
        cmd = [sys.executable, "-m", "segmentron_b200.launch", *launch_flags, os.path.join("tools", script), *argv]
    else:
        env["PYTHONPATH"] = os.pathsep.join([REF, STUBS, env.get("PYTHONPATH", "")])
        boot = ("import sys, runpy, numpy as np; np.int = int; "
                f"sys.argv = [{os.path.join('tools', script)!r}] + sys.argv[1:]; "
                f"runpy.run_path({os.path.join('tools', script)!r}, run_name='__main__')")
        cmd = [sys.executable, "-c", boot, *argv]
    if env_extra:
        env.update(env_extra)
    p = subprocess.run(cmd, cwd=run_dir, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    return p.returncode, p.stdout


def parse_train_losses(log):
    """Losses the reference logs every ``--log-iter`` iterations (tools/train.py:155-161)."""
    return [float(m) for m in re.findall(r"\|\| Loss: ([0-9.]+) \|\|", log)]


def parse_eval_result(log):
    """(pixAcc %, mIoU %) of the 'End validation' line (tools/eval.py:94-95)."""
    m = re.search(r"End validation pixAcc: ([0-9.]+), mIoU: ([0-9.]+)", log)
    return (float(m.group(1)), float(m.group(2))) if m else None
