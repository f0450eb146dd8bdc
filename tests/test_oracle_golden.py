"""Pin the oracle (oracle/wesup_ref.py) against outputs of the real reference
(tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import wesup_ref as O
from wesup_b200 import synth


def _order_counts(maps, seg):
    flat = seg.reshape(-1)
    order = np.array([int(flat[m.reshape(-1).argmax()]) for m in maps])
    counts = np.array([int((m > 0).sum()) for m in maps])
    return order, counts


def test_kat_4x4(golden):
    g = golden("kat_preprocess_4x4.npz")
    maps, labels, order = O.preprocess_superpixels(torch.from_numpy(g["segments"]),
                                                   torch.from_numpy(g["mask"]))
    assert order.tolist() == [0, 2, 1, 3]
    assert labels.tolist() == [[1.0, 1.0], [0.0, 1.0]]
    np.testing.assert_array_equal(maps.numpy(), g["sp_maps"])
    np.testing.assert_array_equal(labels.numpy(), g["sp_labels"])


def test_preprocess_cases(golden):
    g = golden("preprocess_cases.npz")
    for i in range(3):
        seg = torch.from_numpy(g[f"seg{i}"])
        maps, labels, order = O.preprocess_superpixels(seg, torch.from_numpy(g[f"mask{i}"]))
        np.testing.assert_array_equal(order.numpy(), g[f"order{i}"])
        np.testing.assert_array_equal(labels.numpy(), g[f"labels{i}"])
        o2, c2 = _order_counts(maps.numpy(), g[f"seg{i}"])
        np.testing.assert_array_equal(o2, g[f"order{i}"])
        np.testing.assert_array_equal(c2, g[f"counts{i}"])
        _, gland = synth.he_like_image(*g[f"seg{i}"].shape, seed=1000 + i)
        _, labels_f, order_f = O.preprocess_superpixels(seg, synth.pixel_mask(gland))
        np.testing.assert_array_equal(order_f.numpy(), g[f"order_full{i}"])
        np.testing.assert_array_equal(labels_f.numpy(), g[f"labels_full{i}"])
        _, labels_n, order_n = O.preprocess_superpixels(seg, None)
        assert labels_n is None
        np.testing.assert_array_equal(order_n.numpy(), g[f"order_none{i}"])


def test_label_propagate_cases(golden):
    g = golden("label_propagate_cases.npz")
    for i in range(4):
        f, yl = torch.from_numpy(g[f"f{i}"]), torch.from_numpy(g[f"yl{i}"])
        for thr in (0.8, 0.95):
            yu = O.label_propagate(f, yl, threshold=thr)
            np.testing.assert_array_equal(yu.numpy(), g[f"yu{i}_{int(thr * 100)}"])


def test_cross_entropy_cases(golden):
    g = golden("cross_entropy_cases.npz")
    yh, yt = torch.from_numpy(g["y_hat"]), torch.from_numpy(g["y_true"])
    np.testing.assert_allclose(O.cross_entropy(yh, yt).numpy(), g["loss"], rtol=1e-6)
    assert float(O.cross_entropy(yh, torch.zeros_like(yt))) == float(g["loss_none"]) == 0.0


def test_forward_loss_backward(golden):
    g = golden("forward_loss_backward_48x40.npz")
    torch.set_num_threads(4)
    model = O.seeded_init_(O.RefWESUP(), seed=3)
    x = synth.to_tensor(g["img_u8"]).unsqueeze(0)
    maps, labels, _ = O.preprocess_superpixels(torch.from_numpy(g["segments"]),
                                               torch.from_numpy(g["point_mask"]))
    np.testing.assert_array_equal(labels.numpy(), g["sp_labels"])
    pred = model((x, maps))
    np.testing.assert_allclose(model.sp_features.detach().numpy(), g["sp_features"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(model.sp_pred.detach().numpy(), g["sp_pred"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(pred.detach().numpy(), g["pred"], rtol=1e-4, atol=1e-6)
    metrics = {}
    loss = O.compute_loss(model.sp_pred, model.sp_features, labels, metrics=metrics)
    np.testing.assert_allclose(float(loss.detach()), float(g["loss"]), rtol=1e-5)
    assert metrics["labeled_sp_ratio"] == float(g["labeled_sp_ratio"])
    assert metrics["propagated_labels"] == float(g["propagated_labels"])
    loss.backward()
    grads = dict(model.named_parameters())
    for key in g.files:
        if key.startswith("gradnorm_"):
            name = key[len("gradnorm_"):]
            np.testing.assert_allclose(float(grads[name].grad.norm()), float(g[key]), rtol=1e-4)
            np.testing.assert_allclose(grads[name].grad.flatten()[:16].numpy(), g["gradhead_" + name],
                                       rtol=1e-3, atol=1e-6)


def test_state_dict_keys_match_reference_layout():
    keys = list(O.RefWESUP().state_dict())
    offsets = [0, 32, 64, 128, 192, 320, 448, 576, 832, 1088, 1344, 1600, 1856]
    for off in offsets:
        assert f"side_conv{off}.weight" in keys
    for idx in (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28):
        assert f"backbone.{idx}.weight" in keys
    assert {"fc_layers.0.weight", "fc_layers.2.weight", "fc_layers.4.weight", "classifier.0.weight"} <= set(keys)
    assert O.RefWESUP().fm_channels_sum == 2112


def test_pixel_inference(golden):
    g = golden("pixel_inference_32x32.npz")
    model = O.seeded_init_(O.RefWESUP(), seed=3)
    with torch.no_grad():
        out = model.forward_pixels(synth.to_tensor(g["img_u8"]).unsqueeze(0))
    np.testing.assert_allclose(out.numpy(), g["pred"], rtol=1e-4, atol=1e-6)


# ---------------------------------------------------------------------------
# the scalable forms used by the GlaS / CRAG GPU tests must equal their dense twins
# ---------------------------------------------------------------------------
def test_count_form_of_preprocessing_equals_the_dense_form(golden):
    g = golden("preprocess_cases.npz")
    cases = [(g[f"seg{i}"], g[f"mask{i}"]) for i in range(3)]
    k = golden("kat_preprocess_4x4.npz")
    cases.append((k["segments"], k["mask"]))
    rng = np.random.default_rng(3)
    seg = synth.perturbed_grid_segments(61, 47, 9, seed=4)
    mask = np.zeros((3, 61, 47), np.int64)
    pick = rng.random((61, 47)) < 0.1
    cls = rng.integers(0, 3, (61, 47))
    for c in range(3):
        mask[c][pick & (cls == c)] = 1
    cases.append((seg, mask))
    for seg, mask in cases:
        seg_t, mask_t = torch.from_numpy(seg), torch.from_numpy(mask)
        maps, labels, order = O.preprocess_superpixels(seg_t, mask_t)
        order_c, labels_c, counts_c = O.superpixel_order_and_labels_counts(seg_t, mask_t)
        assert order_c.tolist() == order.tolist()
        assert torch.equal(labels_c, labels)
        assert counts_c.tolist() == (maps > 0).sum(dim=(1, 2)).tolist()
        order_n, labels_n, _ = O.superpixel_order_and_labels_counts(seg_t, None)
        assert order_n.tolist() == O.superpixel_order_and_labels(seg_t, None)[0].tolist() and labels_n is None


def test_sparse_hypercolumn_and_pooling_equal_interpolate_and_dense_mm():
    g = torch.Generator().manual_seed(2)
    for h, w in ((48, 40), (37, 51), (75, 94)):                       # 75 -> 37 -> 18 -> 9 -> 4: floor-division levels
        levels = [torch.randn(1, c, h >> s, w >> s, generator=g) for c, s in zip((8, 8, 12, 16, 16), (0, 1, 2, 3, 4))]
        dense = O.hypercolumn_from_sides(levels, (h, w))               # (C,H,W)
        px = torch.randperm(h * w, generator=g)[:200]
        sparse = O.hypercolumn_at_pixels(levels, (h, w), px)
        ref = dense.reshape(dense.size(0), -1).t()[px].double()
        assert float((sparse - ref).abs().max()) < 1e-5 * float(ref.abs().max())
        seg = torch.from_numpy(synth.perturbed_grid_segments(h, w, 8, seed=h))
        order = torch.unique(seg)
        pooled = O.pool_dense(O.dense_sp_maps(seg, order), dense)
        rows = torch.arange(0, order.numel(), 3)
        sparse_pooled = O.pooled_rows_sparse(levels, (h, w), seg.reshape(-1), rows)
        assert float((sparse_pooled - pooled[rows].double()).abs().max()) < 1e-5 * float(pooled.abs().max())


def test_block_form_of_label_propagation_equals_the_dense_form(golden):
    gl = golden("label_propagate_cases.npz")
    cases = [(torch.from_numpy(gl[f"f{i}"]), torch.from_numpy(gl[f"yl{i}"])) for i in range(4)]
    g = torch.Generator().manual_seed(1)
    f = (torch.randn(700, 32, generator=g) * 0.06).abs()
    yl = torch.zeros(90, 2)
    yl[torch.arange(90), torch.randint(0, 2, (90,), generator=g)] = 1
    cases.append((f, yl))
    for f, yl in cases:
        for thr in (0.8, 0.95):
            y_u, src, best = O.label_propagate(f, yl, thr, return_aux=True)
            y_b, src_b, best_b = O.label_propagate_block(f, yl, thr, chunk=97)
            assert torch.equal(y_u, y_b) and torch.equal(src, src_b)
            assert torch.allclose(best, best_b, rtol=1e-6, atol=0)
