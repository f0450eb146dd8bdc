"""The drop-in claim, executed: the reference's OWN, UNMODIFIED
`train.py` (`fit`) and `infer_tile.py` (`predict`) -- imported from the staged copy of the reference under
baseline/_ref/ -- drive `wesup_b200.models` through the `sys.modules['models']` alias, on a small synthetic
image folder.  Runs in a subprocess so the alias never leaks into the other tests."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]

SCRIPT = r"""
import json, os, sys
sys.path.insert(0, {root!r})
from oracle import reference_harness as RH
ref_root = RH.reference_root()
assert ref_root is not None, "baseline/_ref is not staged"
RH.install_stubs(with_albumentations=False)      # utils.data then fails to import: the trainer falls back to its PIL readers
sys.modules["fire"].Fire = lambda fn: None
import wesup_b200.models as models, wesup_b200.models.wesup as wesup, wesup_b200.models.base as base
sys.modules["models"] = models
sys.modules["models.wesup"] = wesup
sys.modules["models.base"] = base
sys.path.insert(0, str(ref_root))
import train as ref_train, infer_tile as ref_tile                  # the reference's files, unmodified
for mod in (ref_train, ref_tile):
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(str(ref_root))), mod.__file__
assert ref_train.initialize_trainer is models.initialize_trainer
import torch
from wesup_b200 import _lib
n0 = _lib.load().wesup_kernel_launches()
ref_train.fit({data!r}, model="wesup", epochs=2, pretrained=False, num_workers=0, cuda_graph={graph})
records = sorted(os.listdir(os.environ["RECORD_ROOT"]))
rec = os.path.join(os.environ["RECORD_ROOT"], records[-1])
import pandas as pd
hist = pd.read_csv(os.path.join(rec, "history.csv"))
trainer = models.initialize_trainer("wesup", device="cuda", pretrained=False)
ckpts = sorted(os.listdir(os.path.join(rec, "checkpoints")))
trainer.load_checkpoint(os.path.join(rec, "checkpoints", ckpts[-1]))
trainer.model.eval()
img_path = sorted(os.listdir(os.path.join({data!r}, "val", "images")))[0]
pred = ref_tile.predict(trainer, os.path.join({data!r}, "val", "images", img_path), 64, device="cuda")
print("RESULT " + json.dumps({{"columns": list(hist.columns), "rows": len(hist), "loss": [float(v) for v in hist["loss"]],
                              "ckpts": ckpts, "pred_shape": list(pred.shape), "pred_values": sorted(set(pred.reshape(-1).tolist()))[:4],
                              "native_launches": int(_lib.load().wesup_kernel_launches() - n0)}}))
"""


def make_dataset(root: Path):
    from PIL import Image
    from wesup_b200 import synth
    for split, n in (("train", 3), ("val", 1)):
        for sub in ("images", "masks", "points"):
            (root / split / sub).mkdir(parents=True, exist_ok=True)
        for i in range(n):
            h, w = (240, 280) if split == "train" else (120, 150)
            img, gland = synth.he_like_image(h, w, seed=300 + i)
            Image.fromarray(img).save(root / split / "images" / f"im{i}.png")
            Image.fromarray((gland * 255).astype(np.uint8)).save(root / split / "masks" / f"im{i}.png")
            rng = np.random.default_rng(i)
            pts = [(int(x), int(y), int(gland[y, x])) for y, x in zip(rng.integers(0, h, 40), rng.integers(0, w, 40))]
            np.savetxt(root / split / "points" / f"im{i}.csv", np.array(pts), fmt="%d", delimiter=",")


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_reference_train_and_infer_tile_drive_the_package_unmodified(tmp_path, graph):
    if not (ROOT / "baseline" / "_ref" / "train.py").exists():
        pytest.skip("baseline/_ref not staged (build() stages it in the build container)")
    data = tmp_path / "data"
    make_dataset(data)
    script = tmp_path / "dropin.py"
    script.write_text(SCRIPT.format(root=str(ROOT), data=str(data), graph=graph))
    env = dict(os.environ, RECORD_ROOT=str(tmp_path / "records"), HOME=str(tmp_path))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert res["rows"] == 2 and all(np.isfinite(res["loss"])), res                       # two epochs of history
    assert {"loss", "accuracy", "dice"} <= set(res["columns"]), res["columns"]           # the reference's metric functions ran
    assert res["ckpts"] == ["ckpt.0002.pth"], res["ckpts"]
    assert res["pred_shape"] == [120, 150] and set(res["pred_values"]) <= {0.0, 0.5, 1.0}, res
    assert res["native_launches"] > 0                                                     # the CUDA library did the work


def test_reference_is_staged_and_importable_when_present():
    """CPU: the staged reference imports behind the stubs and exposes the surface the package mirrors."""
    from oracle import reference_harness as RH
    if RH.reference_root() is None and not RH.SOURCE.exists():
        pytest.skip("no reference tree in this environment")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import reference_harness as RH\n"
            "m = RH.import_reference()\n"
            "import inspect\n"
            "print(sorted(n for n in ('WESUP','WESUPTrainer','WESUPConfig','initialize_trainer') if hasattr(m, n)))\n") % str(ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "['WESUP', 'WESUPConfig', 'WESUPTrainer', 'initialize_trainer']" in r.stdout
