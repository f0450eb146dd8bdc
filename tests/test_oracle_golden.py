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
