"""Running the REAL reference (mrcfps/WESUP) -- TEST INFRASTRUCTURE ONLY.

The reference is pure Python and cannot be pip-installed (no setup.py); `stage()` copies its tree from
/root/reference to baseline/_ref/ in the build container (git-ignored: no reference source enters the
history; not gpurun-ignored: the copy travels to the GPU box, where /root/reference does not exist).
`import_reference()` then imports its `models` package UNMODIFIED behind import stubs for the third-party
packages this image lacks (scikit-image, albumentations, matplotlib, fire) -- none of them is on the hot
path except `skimage.segmentation.slic`, which is served by the C restatement oracle/slic_ref.c -- and with
torchvision's vgg16 patched to random init (there is no network for the ImageNet weights).

Used by: tests (drop-in test: the reference's own train.py / infer_tile.py driving wesup_b200.models),
bench.py's cpu_baseline / --impl reference legs (kind "reference").  Never by the product package.
"""
from __future__ import annotations

import shutil
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
STAGED = ROOT / "baseline" / "_ref"
SOURCE = Path("/root/reference")


def stage(force: bool = False) -> Path | None:
    """Copy the reference tree to baseline/_ref (build container only).  Returns the staged path or None."""
    if not SOURCE.exists():
        return STAGED if (STAGED / "models" / "wesup.py").exists() else None
    if force or not (STAGED / "models" / "wesup.py").exists():
        if STAGED.exists():
            shutil.rmtree(STAGED)
        shutil.copytree(SOURCE, STAGED, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc", "*.pth", "*.png", "*.jpg"))
    return STAGED


def reference_root() -> Path | None:
    return STAGED if (STAGED / "models" / "wesup.py").exists() else None


def _slic_via_c_restatement(image, n_segments=100, compactness=10.0, **_kw):
    from . import slic as oslic
    return oslic.slic(image, int(n_segments), float(compactness))


def install_stubs(with_albumentations: bool = True) -> None:
    """Empty stand-ins for the third-party imports of the reference that this image lacks."""
    import numpy as np
    from PIL import Image
    names = ["skimage", "skimage.segmentation", "skimage.io", "skimage.morphology", "skimage.transform", "skimage.measure",
             "matplotlib", "matplotlib.pyplot", "fire"]
    if with_albumentations:
        names.append("albumentations")
    for name in names:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)

    def absent(*_a, **_k):
        raise RuntimeError("stubbed third-party function (not on the hot path)")
    seg = sys.modules["skimage.segmentation"]
    seg.slic = _slic_via_c_restatement
    seg.find_boundaries = absent
    sys.modules["skimage.io"].imread = lambda p, *a, **k: np.asarray(Image.open(str(p)))
    sys.modules["skimage.io"].imsave = absent
    sys.modules["skimage.morphology"].dilation = absent
    sys.modules["skimage.morphology"].opening = absent
    sys.modules["skimage.transform"].resize = absent
    sys.modules["skimage.measure"].label = absent
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["fire"].Fire = absent


def import_reference(root: Path | None = None):
    """The reference's `models` package (unmodified code) ready to run on the CPU or on CUDA."""
    root = Path(root) if root is not None else (reference_root() or (SOURCE if SOURCE.exists() else None))
    if root is None:
        raise ImportError("the reference is not staged under baseline/_ref (run __graft_entry__.build() in the build container)")
    install_stubs()
    import torchvision
    if not getattr(torchvision.models.vgg16, "_wesup_patched", False):
        orig = torchvision.models.vgg16

        def vgg16(pretrained=False, **kw):          # random init: no network for the ImageNet weights
            return orig(weights=None)
        vgg16._wesup_patched = True
        torchvision.models.vgg16 = vgg16
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    import models as ref_models                      # noqa: E402  (the real reference)
    if not str(Path(ref_models.__file__).resolve()).startswith(str(Path(root).resolve())):
        raise ImportError(f"`models` resolved to {ref_models.__file__}, not to the reference under {root}")
    return ref_models


def reference_sgd(trainer):
    """The optimizer `WESUPTrainer.get_default_optimizer` builds (models/wesup.py:445-455); the function itself passes
    `verbose=` to ReduceLROnPlateau, which torch 2.11 rejects, and it discards the scheduler anyway."""
    import torch
    return torch.optim.SGD(filter(lambda p: p.requires_grad, trainer.model.parameters()), lr=5e-5,
                           momentum=trainer.kwargs.get("momentum"), weight_decay=trainer.kwargs.get("weight_decay"))
