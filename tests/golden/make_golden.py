"""Generate tests/golden/*.npz by running the REAL reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference imports packages that are not installed here (skimage, fire,
albumentations, matplotlib) and downloads VGG16 weights; neither matters for the
hot-path math, so empty stub modules are installed in sys.modules and
torchvision.models.vgg16 is patched to random init.  The
reference code itself runs unmodified.  `skimage.segmentation.slic` cannot run,
so label maps come from wesup_b200.synth.perturbed_grid_segments.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def import_reference():
    for name in ["skimage", "skimage.segmentation", "skimage.io", "skimage.morphology",
                 "skimage.transform", "skimage.measure", "albumentations", "matplotlib",
                 "matplotlib.pyplot", "fire"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sk = sys.modules["skimage.segmentation"]
    sk.slic = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("slic stub"))
    def _absent(*a, **k):
        raise RuntimeError("stubbed third-party function")
    sk.find_boundaries = _absent
    sys.modules["skimage.io"].imread = _absent
    sys.modules["skimage.io"].imsave = _absent
    sys.modules["skimage.morphology"].dilation = _absent
    sys.modules["skimage.morphology"].opening = _absent
    sys.modules["skimage.transform"].resize = _absent
    sys.modules["skimage.measure"].label = _absent
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["fire"].Fire = _absent
    import torchvision
    orig = torchvision.models.vgg16
    torchvision.models.vgg16 = lambda pretrained=False, **k: orig(weights=None)
    sys.path.insert(0, str(REF))
    import models.wesup as ref          # noqa: E402  (the real reference)
    return ref


def main():
    sys.path.insert(0, str(ROOT))
    from oracle.wesup_ref import seeded_init_
    from wesup_b200 import synth
    ref = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(4)

    # ---- KAT: 4x4 map with a two-class superpixel -------------
    seg = torch.tensor([[0, 0, 1, 1], [0, 0, 1, 1], [2, 2, 3, 3], [2, 2, 3, 3]])
    mask = torch.zeros(2, 4, 4, dtype=torch.long)
    mask[0, 0, 0] = 1
    mask[1, 0, 1] = 1            # superpixel 0: one pixel of each class -> multi-hot
    mask[1, 2, 0] = 1            # superpixel 2: class 1
    maps, labels = ref._preprocess_superpixels(seg, mask)
    np.savez(OUT / "kat_preprocess_4x4.npz", segments=seg.numpy(), mask=mask.numpy(),
             sp_maps=maps.numpy(), sp_labels=labels.numpy())

    # ---- preprocess on ragged maps, with point masks / full masks / none ---
    cases = {}
    for i, (h, w, cell, ratio) in enumerate([(40, 56, 8, 0.01), (33, 47, 6, 0.02), (64, 64, 9, 0.003)]):
        seg_np = synth.perturbed_grid_segments(h, w, cell, seed=10 + i)
        _, gland = synth.he_like_image(h, w, seed=1000 + i)
        pm = synth.point_mask(gland, seed=2000 + i, ratio=ratio)
        maps, labels = ref._preprocess_superpixels(torch.from_numpy(seg_np), pm)
        cases[f"seg{i}"] = seg_np
        cases[f"mask{i}"] = pm.numpy()
        cases[f"order{i}"] = np.array([int(seg_np.reshape(-1)[m.reshape(-1).argmax()]) for m in maps])
        cases[f"counts{i}"] = np.array([int((m > 0).sum()) for m in maps])
        cases[f"labels{i}"] = labels.numpy()
        full = synth.pixel_mask(gland)
        maps_f, labels_f = ref._preprocess_superpixels(torch.from_numpy(seg_np), full)
        cases[f"labels_full{i}"] = labels_f.numpy()
        cases[f"order_full{i}"] = np.array([int(seg_np.reshape(-1)[m.reshape(-1).argmax()]) for m in maps_f])
        maps_n, labels_n = ref._preprocess_superpixels(torch.from_numpy(seg_np), None)
        cases[f"order_none{i}"] = np.array([int(seg_np.reshape(-1)[m.reshape(-1).argmax()]) for m in maps_n])
        assert labels_n.dim() == 0
    np.savez(OUT / "preprocess_cases.npz", **cases)

    # ---- label propagation -------------------------------------------------
    lp = {}
    g = torch.Generator().manual_seed(7)
    for i, (n, n_l, scale) in enumerate([(40, 6, 0.06), (257, 33, 0.06), (130, 64, 0.05), (64, 3, 0.2)]):
        f = (torch.randn(n, 32, generator=g) * scale).abs()          # post-ReLU features are >= 0
        y_l = torch.zeros(n_l, 2)
        y_l[torch.arange(n_l), torch.randint(0, 2, (n_l,), generator=g)] = 1
        if n_l > 2:
            y_l[1] = 1.0                                              # one multi-hot row
        if i == 0:
            f[n_l + 2] = f[3]                                         # exact duplicate of labeled row 3
            f[4] = f[3]                                               # tie between labeled rows 3 and 4
        for thr in (0.8, 0.95):
            y_u = ref._label_propagate(f, y_l, threshold=thr)
            lp[f"yu{i}_{int(thr * 100)}"] = y_u.numpy()
        lp[f"f{i}"] = f.numpy()
        lp[f"yl{i}"] = y_l.numpy()
    np.savez(OUT / "label_propagate_cases.npz", **lp)

    # ---- cross entropy -----------------------------------------------------
    ce = {}
    y_hat = torch.softmax(torch.randn(12, 2, generator=g), dim=1)
    y_hat[0] = torch.tensor([1.0, 0.0])                               # exercises the clamp
    y_true = torch.zeros(12, 2)
    y_true[:5, 0] = 1
    y_true[5, :] = 1                                                  # multi-hot row
    ce["y_hat"], ce["y_true"] = y_hat.numpy(), y_true.numpy()
    ce["loss"] = ref._cross_entropy(y_hat, y_true).numpy()
    ce["loss_none"] = ref._cross_entropy(y_hat, torch.zeros(12, 2)).numpy()
    np.savez(OUT / "cross_entropy_cases.npz", **ce)

    # ---- full forward + loss + backward ------------------------------------
    h, w = 48, 40
    img_u8, gland = synth.he_like_image(h, w, seed=1003)
    x = synth.to_tensor(img_u8).unsqueeze(0)
    seg_np = synth.perturbed_grid_segments(h, w, 8, seed=13)
    pm = synth.point_mask(gland, seed=2003, ratio=0.004)
    model = ref.WESUP()
    seeded_init_(model, seed=3)
    trainer = ref.WESUPTrainer(model, device="cpu")
    sp_maps, sp_labels = ref._preprocess_superpixels(torch.from_numpy(seg_np), pm)
    pred = model((x, sp_maps))
    sp_features = model.sp_features.detach().clone()
    sp_pred = model.sp_pred.detach().clone()
    feats = model.feature_maps.detach().clone()
    metrics = {}
    loss = trainer.compute_loss(pred, (None, sp_labels), metrics=metrics)
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    picks = ["backbone.0.weight", "backbone.28.bias", "side_conv0.weight", "side_conv448.weight",
             "side_conv1856.bias", "fc_layers.0.weight", "fc_layers.4.bias", "classifier.0.weight"]
    out = dict(img_u8=img_u8, segments=seg_np, point_mask=pm.numpy(), pred=pred.detach().numpy(),
               sp_features=sp_features.numpy(), sp_pred=sp_pred.numpy(), loss=loss.detach().numpy(),
               sp_labels=sp_labels.numpy(),
               pooled_probe=torch.mm(sp_maps.view(sp_maps.size(0), -1),
                                     feats.view(feats.size(0), -1).t()).numpy(),
               feats_probe=feats[:, ::7, ::5].numpy(),
               labeled_sp_ratio=np.float64(metrics["labeled_sp_ratio"]),
               propagated_labels=np.float64(metrics["propagated_labels"]),
               propagate_loss=np.float64(metrics["propagate_loss"]))
    for k in picks:
        out["gradnorm_" + k] = grads[k].norm().numpy()
        out["gradhead_" + k] = grads[k].flatten()[:16].numpy()
    np.savez(OUT / "forward_loss_backward_48x40.npz", **out)

    # ---- pixel inference ---------------------------------------------------
    pix = ref.WESUPPixelInference()
    seeded_init_(pix, seed=3)
    with torch.no_grad():
        pp = pix(x[:, :, :32, :32])
    np.savez(OUT / "pixel_inference_32x32.npz", img_u8=img_u8[:32, :32], pred=pp.numpy())
    # ---- tile split / merge (infer_tile.py:23-91) -----------------------------
    import infer_tile as ref_tile       # the real reference (fire / skimage.io are stubbed)
    rng = np.random.default_rng(5)
    tl = {}
    for i, (h, w, p) in enumerate([(50, 37, 16), (40, 40, 20), (33, 70, 24)]):
        img = rng.integers(0, 255, (h, w, 3), dtype=np.uint8)
        patches = ref_tile.divide_image_to_patches(img, p)
        preds = rng.random((len(patches), p, p))
        tl[f"img{i}"], tl[f"patch{i}"], tl[f"patches{i}"], tl[f"preds{i}"] = img, np.int64(p), patches, preds
        tl[f"combined{i}"] = ref_tile.combine_patches_to_image(preds, h, w)
        tl[f"coords{i}"] = np.array(list(ref_tile._get_top_left_coordinates(h, w, p)))
    np.savez_compressed(OUT / "tiles_cases.npz", **tl)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
