"""GPU parity of the Python surface (WESUP / WESUPTrainer / WESUPPixelInference)
against the golden vectors minted from the real reference and against the CPU
oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import wesup_ref as O                      # noqa: E402
from wesup_b200 import ops, synth                            # noqa: E402
from wesup_b200.models import WESUP, WESUPPixelInference, WESUPTrainer, initialize_trainer  # noqa: E402
from wesup_b200.models.wesup import _cross_entropy, _label_propagate, _preprocess_superpixels  # noqa: E402
from wesup_b200.ops import SuperpixelMaps              # noqa: E402
from wesup_b200.utils import is_empty_tensor           # noqa: E402

DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_fp32():
    """The oracle is CPU fp32; keep cuDNN/cuBLAS out of TF32 for parity runs."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def build(cls=WESUP, **kw):
    model = cls(pretrained=False, **kw)
    O.seeded_init_(model, seed=3)
    return model.to(DEV)


@pytest.mark.parametrize("layout,fused,materialize,pool_first,footprints",
                         [("hwc", True, True, False, True), ("hwc", False, True, False, True), ("chw", False, True, False, True),
                          ("hwc", True, False, False, True), ("hwc", True, False, True, True),
                          ("hwc", True, False, False, False), ("hwc", True, False, True, False)])
def test_forward_loss_backward_matches_reference(golden, layout, fused, materialize, pool_first, footprints):
    g = golden("forward_loss_backward_48x40.npz")
    model = build(hc_layout=layout, fused_backward=fused, materialize_hypercolumn=materialize, pool_first=pool_first,
                  footprints=footprints)
    trainer = WESUPTrainer(model, device=DEV)
    x = synth.to_tensor(g["img_u8"]).unsqueeze(0).to(DEV)
    sp_maps, sp_labels = _preprocess_superpixels(torch.from_numpy(g["segments"]).to(DEV),
                                                 torch.from_numpy(g["point_mask"]).to(DEV))
    assert isinstance(sp_maps, SuperpixelMaps) and tuple(sp_maps.size()) == (30, 48, 40)
    np.testing.assert_array_equal(sp_labels.cpu().numpy(), g["sp_labels"])
    pred = model((x, sp_maps))
    assert pred.shape == (1, 48, 40) and pred.dtype == torch.float32
    if materialize:
        assert model.feature_maps.shape == (2112, 48, 40)
        probe = model.feature_maps[:, ::7, ::5].detach().cpu()
        ref_probe = torch.from_numpy(g["feats_probe"])
        assert float((probe - ref_probe).norm() / ref_probe.norm()) < 1e-5            # 1e-4 relative is the north-star bar
        np.testing.assert_allclose(probe.numpy(), g["feats_probe"], rtol=1e-4, atol=1e-4)   # cuDNN conv algorithms differ in the last bits
    else:
        assert model.feature_maps is None                                              # nothing of size H*W*C exists
    np.testing.assert_allclose(model.sp_features.detach().cpu().numpy(), g["sp_features"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(model.sp_pred.detach().cpu().numpy(), g["sp_pred"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(pred.cpu().numpy(), g["pred"], rtol=1e-4, atol=1e-6)
    metrics = {}
    loss = trainer.compute_loss(pred, (None, sp_labels), metrics=metrics)
    np.testing.assert_allclose(float(loss.detach()), float(g["loss"]), rtol=1e-4)      # north-star tolerance
    assert metrics["labeled_sp_ratio"] == float(g["labeled_sp_ratio"])
    assert metrics["propagated_labels"] == float(g["propagated_labels"])
    assert model.sp_pred is None
    loss.backward()
    grads = dict(model.named_parameters())
    for key in g.files:
        if key.startswith("gradnorm_"):
            name = key[len("gradnorm_"):]
            np.testing.assert_allclose(float(grads[name].grad.norm()), float(g[key]), rtol=1e-3)
            np.testing.assert_allclose(grads[name].grad.flatten()[:16].cpu().numpy(), g["gradhead_" + name],
                                       rtol=2e-3, atol=1e-5)
    post, tgt = trainer.postprocess(pred, (torch.zeros(1, 2, 48, 40, device=DEV), sp_labels))
    assert post.dtype == torch.long and tgt.shape == (1, 48, 40)


def test_dense_sp_maps_give_the_same_forward(golden):
    g = golden("forward_loss_backward_48x40.npz")
    model = build()
    x = synth.to_tensor(g["img_u8"]).unsqueeze(0).to(DEV)
    dense, _ = _preprocess_superpixels(torch.from_numpy(g["segments"]).to(DEV),
                                       torch.from_numpy(g["point_mask"]).to(DEV), dense=True)
    assert torch.is_tensor(dense) and dense.shape == (30, 48, 40)
    with torch.no_grad():
        pred = model((x, dense))
    np.testing.assert_allclose(pred.cpu().numpy(), g["pred"], rtol=1e-4, atol=1e-6)


def test_bf16_hypercolumn_within_1e2(golden):
    g = golden("forward_loss_backward_48x40.npz")
    model = build(hc_dtype=torch.bfloat16)
    x = synth.to_tensor(g["img_u8"]).unsqueeze(0).to(DEV)
    sp_maps, sp_labels = _preprocess_superpixels(torch.from_numpy(g["segments"]).to(DEV),
                                                 torch.from_numpy(g["point_mask"]).to(DEV))
    with torch.no_grad():
        model((x, sp_maps))
    ref = torch.from_numpy(g["sp_features"])
    got = model.sp_features.cpu()
    assert float((got - ref).norm() / ref.norm()) < 1e-2


def test_pixel_inference_matches_reference(golden):
    g = golden("pixel_inference_32x32.npz")
    model = build(WESUPPixelInference)
    x = synth.to_tensor(g["img_u8"]).unsqueeze(0).to(DEV)
    with torch.no_grad():
        out = model(x)
    assert out.shape == (32, 32, 2)
    np.testing.assert_allclose(out.cpu().numpy(), g["pred"], rtol=1e-4, atol=1e-6)
    # same state_dict as WESUP (pixel_infer_tile.py:38-39 of the reference)
    model.load_state_dict(build(WESUP).state_dict())
    # opt-in bf16 hypercolumn + bf16 tensor-core MLP: class probabilities within 1e-2
    model16 = build(WESUPPixelInference, hc_dtype=torch.bfloat16)
    with torch.no_grad():
        out16 = model16(x)
    assert out16.dtype == torch.float32 and out16.shape == (32, 32, 2)
    assert float((out16.cpu() - torch.from_numpy(g["pred"])).abs().max()) < 1e-2


def test_module_level_functions(golden):
    g = golden("cross_entropy_cases.npz")
    yh, yt = torch.from_numpy(g["y_hat"]).to(DEV), torch.from_numpy(g["y_true"]).to(DEV)
    np.testing.assert_allclose(float(_cross_entropy(yh, yt)), float(g["loss"]), rtol=1e-6)
    assert float(_cross_entropy(yh, torch.zeros_like(yt))) == 0.0
    lp = golden("label_propagate_cases.npz")
    y_u = _label_propagate(torch.from_numpy(lp["f1"]).to(DEV), torch.from_numpy(lp["yl1"]).to(DEV), threshold=0.8)
    np.testing.assert_array_equal(y_u.cpu().numpy(), lp["yu1_80"])
    seg = torch.from_numpy(synth.perturbed_grid_segments(20, 20, 5, seed=1)).to(DEV)
    sp, labels = _preprocess_superpixels(seg)
    assert is_empty_tensor(labels)


def test_trainer_preprocess_and_train_iteration_end_to_end():
    """SLIC -> stats -> forward -> loss -> backward -> SGD step through the
    trainer API (models/base.py:184-211 of the reference), and agreement of the
    whole step with the CPU oracle given the same label map and weights."""
    torch.manual_seed(0)
    trainer = initialize_trainer("wesup", device=DEV, pretrained=False)
    O.seeded_init_(trainer.model, seed=5)
    img, pixel_mask, point_mask = synth.sample(96, 112, index=3, ratio=2e-3)
    (x, sp), (pm, sp_labels) = trainer.preprocess(img, pixel_mask, point_mask)
    assert x.is_cuda and isinstance(sp, SuperpixelMaps)
    assert sp.n_labeled == sp_labels.size(0) > 0
    # oracle on the GPU-produced label map (SLIC parity is tested separately)
    seg = torch.empty(96 * 112, dtype=torch.long)
    seg[:] = sp.order.cpu().long()[sp.row_labels.cpu().long()]
    ref_model = O.seeded_init_(O.RefWESUP(), seed=5)
    maps, labels, order = O.preprocess_superpixels(seg.view(96, 112), point_mask[0])
    assert order.tolist() == sp.order.cpu().tolist()
    assert torch.equal(labels, sp_labels.cpu())
    ref_pred = ref_model((img, maps))
    ref_loss = O.compute_loss(ref_model.sp_pred, ref_model.sp_features, labels)
    ref_loss.backward()
    trainer.optimizer, _ = trainer.get_default_optimizer()
    trainer.optimizer.zero_grad()
    pred = trainer.model((x, sp))
    loss = trainer.compute_loss(pred, (pm, sp_labels), metrics={})
    loss.backward()
    np.testing.assert_allclose(pred.cpu().numpy(), ref_pred.detach().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(float(loss.detach()), float(ref_loss.detach()), rtol=1e-4)
    ref_grads = dict(ref_model.named_parameters())
    for name, p in trainer.model.named_parameters():
        r = ref_grads[name].grad
        err = float((p.grad.cpu() - r).norm() / (r.norm() + 1e-12))
        assert err < 2e-3, (name, err)
    before = trainer.model.classifier[0].weight.detach().clone()
    trainer.optimizer.step()
    assert not torch.equal(before, trainer.model.classifier[0].weight.detach())
    # the public iteration entry point runs as well
    from wesup_b200.utils.metrics import accuracy, dice
    trainer.metric_funcs = [accuracy, dice]
    trainer.train_one_iteration("train", img, pixel_mask, point_mask)
    assert "loss" not in trainer.tracker.history       # default metrics_lag=1: scalars are read one iteration later
    trainer.flush_metrics()
    assert "loss" in trainer.tracker.history and "dice" in trainer.tracker.history


def test_preprocess_input_arity():
    trainer = initialize_trainer("wesup", device=DEV, pretrained=False)
    img, pixel_mask, point_mask = synth.sample(64, 64, index=1, ratio=1e-2)
    (_, sp1), (pm1, l1) = trainer.preprocess(img)
    assert is_empty_tensor(pm1) and is_empty_tensor(l1)
    (_, sp2), (_, l2) = trainer.preprocess(img, pixel_mask)
    assert l2.size(0) == sp2.n                       # full mask: every superpixel labeled
    with pytest.raises(ValueError):
        trainer.preprocess(img, pixel_mask, point_mask, img)
    with pytest.raises(ValueError):
        initialize_trainer("mild")
    with pytest.raises(RuntimeError):
        trainer.compute_loss(None, (None, l2))       # no forward pass yet


def test_prefetched_preprocess_is_identical_and_single_use():
    trainer = initialize_trainer("wesup", device=DEV, pretrained=False)
    a = synth.sample(96, 112, index=5, ratio=2e-3)
    b = synth.sample(96, 112, index=6, ratio=2e-3)
    (xa, spa), (pma, la) = trainer.preprocess(*a)
    trainer.prefetch(*a)
    assert len(trainer._prefetched) == 1
    (xb, spb), (pmb, lb) = trainer.preprocess(*b)            # different tensors: the staged result must not be used
    assert spb.n != 0 and len(trainer._prefetched) == 1
    trainer.prefetch(*b)                                     # a second staged entry does not evict the first
    (x2, sp2), (pm2, l2) = trainer.preprocess(*a)            # same tensors: picks the staged result up
    assert len(trainer._prefetched) == 1
    assert sp2.n == spa.n and sp2.n_labeled == spa.n_labeled
    for name in ("order", "row_labels", "counts", "seg_offsets", "seg_pixels"):
        assert torch.equal(getattr(sp2, name), getattr(spa, name)), name
    assert torch.equal(l2, la) and torch.equal(x2, xa)
    # a full iteration through the prefetch path
    trainer.optimizer, _ = trainer.get_default_optimizer()
    trainer.train_one_iteration("train", *b)
    trainer.flush_metrics()
    assert len(trainer._prefetched) == 0
    assert np.isfinite(trainer.tracker.history["loss"][-1])


def test_metrics_lag_zero_reads_in_the_same_iteration_and_nan_raises():
    from wesup_b200.utils.metrics import accuracy, dice
    trainer = initialize_trainer("wesup", device=DEV, pretrained=False, metrics_lag=0)
    trainer.optimizer, _ = trainer.get_default_optimizer()
    trainer.metric_funcs = [accuracy, dice]
    b = synth.sample(96, 112, index=6, ratio=2e-3)
    trainer.train_one_iteration("train", *b)
    assert len(trainer.tracker.history["loss"]) == 1 and np.isfinite(trainer.tracker.history["loss"][-1])
    with torch.no_grad():
        trainer.model.classifier[0].weight.fill_(float("nan"))
    with pytest.raises(ValueError, match="Loss is nan!"):
        trainer.train_one_iteration("train", *b)
    # lagged mode: the same error surfaces when the scalars are read
    lagged = initialize_trainer("wesup", device=DEV, pretrained=False)
    lagged.optimizer, _ = lagged.get_default_optimizer()
    with torch.no_grad():
        lagged.model.classifier[0].weight.fill_(float("nan"))
    lagged.train_one_iteration("train", *b)
    with pytest.raises(ValueError, match="Loss is nan!"):
        lagged.flush_metrics()


def test_nan_loss_does_not_touch_the_weights_with_lagged_metrics():
    """metrics_lag=1 reads the loss one iteration late; the fused optimizer step is guarded on the device, so the
    weights and the momentum of a NaN iteration stay what they were (the reference raises before backward)."""
    trainer = initialize_trainer("wesup", device=DEV, pretrained=False)
    trainer.optimizer, _ = trainer.get_default_optimizer()
    b = synth.sample(96, 112, index=6, ratio=2e-3)
    trainer.train_one_iteration("train", *b)                      # a regular step (creates the momentum buffers)
    trainer.flush_metrics()
    before = [p.detach().clone() for p in trainer.model.parameters()]
    moment = [trainer.optimizer.state[p]["momentum_buffer"].clone() for p in trainer.model.parameters()]
    real = trainer.xentropy
    trainer.xentropy = lambda *a, **k: real(*a, **k) * float("nan")
    trainer.train_one_iteration("train", *b)                      # NaN loss, NaN gradients: the step must be skipped
    trainer.xentropy = real
    with pytest.raises(ValueError, match="Loss is nan!"):
        trainer.flush_metrics()
    for p, q in zip(trainer.model.parameters(), before):
        assert torch.equal(p.detach(), q)
    for p, m in zip(trainer.model.parameters(), moment):
        assert torch.equal(trainer.optimizer.state[p]["momentum_buffer"], m)
    trainer.train_one_iteration("train", *b)                      # and training goes on from intact weights
    trainer.flush_metrics()
    assert np.isfinite(trainer.tracker.history["loss"][-1])
    assert any(not torch.equal(p.detach(), q) for p, q in zip(trainer.model.parameters(), before))


def test_cuda_graph_iteration_matches_the_eager_iteration():
    """cuda_graph=True: one captured graph per input shape (SLIC -> ... -> SGD step), replayed for images with
    different superpixel / labeled counts; losses, metrics and parameters follow the eager trainer's."""
    from wesup_b200.utils.metrics import accuracy, dice
    data = [synth.sample(96, 112, index=i, ratio=2e-3) for i in (3, 4, 5)]

    def run(graph):
        trainer = initialize_trainer("wesup", device=DEV, pretrained=False, materialize_hypercolumn=False,
                                     cuda_graph=graph, cuda_graph_after=1)
        O.seeded_init_(trainer.model, seed=11)
        trainer.optimizer, _ = trainer.get_default_optimizer()
        trainer.optimizer.param_groups[0]["lr"] = 1e-3          # large enough for the updates to matter within 6 steps
        trainer.metric_funcs = [accuracy, dice]
        for i in range(6):
            trainer.train_one_iteration("train", *data[i % 3])
        trainer.flush_metrics()
        return trainer

    eager, graphed = run(False), run(True)
    assert 1 <= len(graphed._graphs) <= 3                          # one graph per (shape, 64-row capacity bucket)
    he, hg = eager.tracker.history, graphed.tracker.history
    assert set(he) == set(hg)
    for key in he:
        assert len(he[key]) == len(hg[key]) == 6
        # the two trainers run SLIC separately (its fp64 colour sums are atomics: a near-tie pixel may land in the
        # neighbouring superpixel on one of them), cuDNN's wgrad reductions are not run-to-run deterministic, and
        # the padded rows change fp32 summation order; all of it compounds over the steps
        np.testing.assert_allclose(hg[key], he[key], rtol=5e-3, atol=1e-5, err_msg=key)
        np.testing.assert_allclose(hg[key][:2], he[key][:2], rtol=5e-4, atol=1e-6, err_msg=key)
    assert len(set(np.round(he["labeled_sp_ratio"], 6))) > 1        # the replayed images really differ in their counts
    for (name, a), (_, b) in zip(eager.model.named_parameters(), graphed.model.named_parameters()):
        assert float((a.detach() - b.detach()).norm() / (a.detach().norm() + 1e-12)) < 1e-3, name


def test_label_propagate_static_equals_the_sliced_call():
    g = torch.Generator().manual_seed(5)
    n_max, n, n_l = 300, 211, 17
    f = (torch.randn(n_max, 32, generator=g) * 0.06).to(DEV)
    y = torch.zeros(n_max, 2)
    y[torch.arange(n_l), torch.randint(0, 2, (n_l,), generator=g)] = 1
    y = y.to(DEV)
    counts = torch.tensor([n, n_l], dtype=torch.int32, device=DEV)
    full = ops.label_propagate_static(f, y, counts, 0.8)
    ref = ops.label_propagate(f[:n], y[:n_l], 0.8, algo="exact")
    assert torch.equal(full[n_l:n], ref)
    assert float(full[:n_l].abs().sum()) == 0 and float(full[n:].abs().sum()) == 0
    assert float(ref.sum()) > 0
