"""Parity at the BASELINE shapes that are not 464x464: GlaS 522x775, CRAG
1516x1512 (odd level sizes 189 -> 94), the 2048^2 microbench, label propagation at N = 11 460 /
n_l = 229 and the 8000 / 4000 stress shape.  The dense oracle cannot be held at these sizes
(dense sp_maps: 105 GB at CRAG), so the checks use the oracle's scalable forms -- each pinned to
its dense twin by tests/test_oracle_golden.py -- on sampled rows / cells, plus size-independent
properties (adjoint identity, row sums).  The end-to-end step at GlaS size is compared with the
oracle's dense formulation run on the same GPU in fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import wesup_ref as O                      # noqa: E402
from wesup_b200 import ops, synth                      # noqa: E402
from wesup_b200.ops import SuperpixelMaps              # noqa: E402

DEV = "cuda"
FULL_C = [64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512]       # backbone levels (pool-first path)
SIDE_C = [32, 32, 64, 64, 128, 128, 128, 256, 256, 256, 256, 256, 256]
SHAPES = {"glas": (522, 775), "crag": (1516, 1512)}


def level_sizes(h, w):
    out = []
    for n_convs in (2, 2, 3, 3, 3):
        out += [(h, w)] * n_convs
        h, w = h // 2, w // 2
    return out


def slic_maps(h, w, index):
    img, pixel_mask, point_mask = synth.sample(h, w, index=index)
    labels, n = ops.slic(img.to(DEV), int(h * w / 200), 40)
    return img, labels, int(n.item()), point_mask[0]


@pytest.mark.parametrize("shape", ["glas", "crag"])
def test_sp_stats_bit_exact_at_baseline_shapes(shape):
    h, w = SHAPES[shape]
    _, labels, n, mask = slic_maps(h, w, 31)
    for m in (mask, synth.pixel_mask(synth.he_like_image(h, w, seed=1031)[1]), None):
        order, sp_labels, counts = O.superpixel_order_and_labels_counts(labels.cpu(), m)
        sp = SuperpixelMaps.from_labels(labels, None if m is None else m.to(DEV), n_sp=n)
        assert torch.equal(sp.order.cpu().long(), order)
        assert torch.equal(sp.counts.cpu().long(), counts)
        rank = torch.empty(n, dtype=torch.long)
        rank[order] = torch.arange(n)
        assert torch.equal(sp.label_map.cpu().long(), rank[labels.cpu().long()])
        offs = sp.seg_offsets.cpu().long()
        assert torch.equal(offs[1:] - offs[:-1], counts) and int(offs[-1]) == h * w
        px = sp.seg_pixels.cpu().long()
        assert torch.equal(sp.row_labels.cpu().long()[px], torch.repeat_interleave(torch.arange(n), counts))
        if m is None:
            assert sp.sp_labels is None
        else:
            assert sp.n_labeled == sp_labels.size(0) and torch.equal(sp.sp_labels.cpu(), sp_labels)


def make_levels(h, w, channels, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [torch.randn(1, c, hh, ww, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
            for c, (hh, ww) in zip(channels, level_sizes(h, w))]


def level_grad_at_cells(level_hw, size, row_labels, inv_count, g_level, cells):
    """Reference gradient of one low-resolution level at the given (i, j) cells, fp64: the adjoint of
    upsample + mean, sum over the pixels that tap the cell of tap weight * g[row(pixel)] / |S_row|."""
    H, W = size
    h, w = level_hw
    y0, y1, wy0, wy1 = O.bilinear_taps(torch.arange(H), h, H)
    x0, x1, wx0, wx1 = O.bilinear_taps(torch.arange(W), w, W)
    rl = row_labels.view(H, W)
    out = []
    for i, j in cells:
        wy = wy0.double() * (y0 == i) + wy1.double() * (y1 == i)
        wx = wx0.double() * (x0 == j) + wx1.double() * (x1 == j)
        ys, xs = torch.nonzero(wy).flatten(), torch.nonzero(wx).flatten()
        rows = rl[ys.to(DEV)][:, xs.to(DEV)].long()                         # (ny, nx)
        wgt = (wy[ys].unsqueeze(1) * wx[xs].unsqueeze(0)).to(DEV)
        contrib = g_level.double()[rows] * inv_count[rows].unsqueeze(-1) * wgt.unsqueeze(-1)
        out.append(contrib.sum(dim=(0, 1)))
    return torch.stack(out)


@pytest.mark.parametrize("shape,channels", [("glas", FULL_C), ("glas", SIDE_C), ("crag", FULL_C)])
def test_default_path_pooling_fwd_bwd_at_baseline_shapes(shape, channels):
    h, w = SHAPES[shape]
    _, labels, n, mask = slic_maps(h, w, 32)
    sp = SuperpixelMaps.from_labels(labels, mask.to(DEV), n_sp=n)
    levels = [lv.requires_grad_(True) for lv in make_levels(h, w, channels, seed=h)]
    sizes = level_sizes(h, w)
    assert sizes[-1] == ((32, 48) if shape == "glas" else (94, 94))
    fp = ops.build_footprints(sp, sizes, with_bwd=True)
    pooled, _ = ops.hypercolumn_pool(levels, (h, w), sp, materialize=False, footprints=fp)
    pooled_inkernel, _ = ops.hypercolumn_pool([lv.detach() for lv in levels], (h, w), sp, materialize=False, footprints=None)
    # forward: 64 sampled superpixels against the fp64 sparse reference (first, last, random rows)
    g = torch.Generator().manual_seed(n)
    rows = torch.cat([torch.tensor([0, n - 1]), torch.randperm(n, generator=g)[:62]])
    ref = O.pooled_rows_sparse([lv.detach() for lv in levels], (h, w), sp.row_labels.long(), rows)
    scale = float(ref.abs().max())
    assert float((pooled.detach()[rows.to(DEV)].double() - ref).abs().max()) < 1e-4 * scale
    assert float((pooled_inkernel[rows.to(DEV)].double() - ref).abs().max()) < 1e-4 * scale
    assert float((pooled.detach() - pooled_inkernel).abs().max()) < 1e-4 * scale
    # backward: adjoint identity over the whole tensor + sampled cells of every distinct resolution
    gp = torch.randn(pooled.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(7))
    pooled.backward(gp)
    lhs = float((pooled.detach().double() * gp.double()).sum())
    rhs = float(sum((lv.detach().double() * lv.grad.double()).sum() for lv in levels))
    assert abs(lhs - rhs) < 1e-6 * (abs(lhs) + float(pooled.detach().abs().double().sum()) * 1e-3)
    inv_count = 1.0 / sp.counts.double()
    coff = np.cumsum([0] + list(channels))
    for li in (0, 2, 4, 7, 12):
        hh, ww = sizes[li]
        cells = [(0, 0), (hh - 1, ww - 1), (hh // 2, ww // 3), (1, ww - 2)]
        ref_g = level_grad_at_cells((hh, ww), (h, w), sp.row_labels, inv_count, gp[:, coff[li]:coff[li + 1]], cells)
        got = torch.stack([levels[li].grad[0, :, i, j] for i, j in cells]).double()
        assert float((got - ref_g).abs().max()) < 1e-4 * float(ref_g.abs().max() + 1e-12), f"level {li}"


def test_level_sizes_match_real_conv_shapes_on_the_floor_division_chain():
    from wesup_b200.models.wesup import WESUP
    model = WESUP(pretrained=False).to(DEV)
    for h, w in ((1516, 1512), (522, 775), (75, 94)):
        with torch.no_grad():
            x = torch.zeros(1, 3, h, w, device=DEV).contiguous(memory_format=torch.channels_last)
            real = []
            for layer in model.backbone:
                x = layer(x)
                if isinstance(layer, torch.nn.Conv2d):
                    real.append((x.size(2), x.size(3)))
        assert model._level_sizes(h, w) == real == level_sizes(h, w)
    assert level_sizes(1516, 1512)[-1] == (94, 94) and level_sizes(1516, 1512)[7] == (189, 189)


def check_lp_against_block_oracle(f, y_l, thr, algo):
    y_ref, src_ref, sim_ref = O.label_propagate_block(f, y_l, thr)
    y_u, src, sim = ops.label_propagate(f.to(DEV), y_l.to(DEV), thr, return_aux=True, algo=algo)
    y_u, src, sim = y_u.cpu(), src.cpu().long(), sim.cpu()
    # rows whose best two similarities (or similarity and threshold) are within 4 fp32 ulp are order-dependent in
    # the reference itself (einsum order is backend-defined): they must pick one of the tied answers
    n_l = y_l.size(0)
    same = src == src_ref
    for u in torch.nonzero(~same).flatten().tolist():
        d_mine = float(((f[n_l + u].double() - f[src[u]].double()) ** 2).sum())
        d_ref = float(((f[n_l + u].double() - f[src_ref[u]].double()) ** 2).sum())
        assert abs(np.exp(-d_mine) - np.exp(-d_ref)) < 4 * 1.2e-7, f"row {u}: src {int(src[u])} vs {int(src_ref[u])}"
    assert float(same.float().mean()) >= 0.999
    assert torch.allclose(sim, sim_ref, rtol=1e-5, atol=1e-7)
    clear = same & ((sim_ref - thr).abs() > 4 * 1.2e-7)
    assert torch.equal(y_u[clear], y_ref[clear])
    return y_u, src


@pytest.mark.parametrize("n,n_l", [(11460, 229), (8000, 4000), (20971, 419), (2022, 40)])
@pytest.mark.parametrize("algo", ["auto", "exact", "tc"])
def test_label_propagation_at_baseline_sizes_directly_vs_oracle(n, n_l, algo):
    g = torch.Generator().manual_seed(n + n_l)
    f = (torch.randn(n, 32, generator=g) * 0.06).abs()
    y_l = torch.zeros(n_l, 2)
    y_l[torch.arange(n_l), torch.randint(0, 2, (n_l,), generator=g)] = 1
    y_l[::7, :] = 1                                                          # some multi-hot labeled rows
    y_u, _ = check_lp_against_block_oracle(f, y_l, 0.8, algo)
    assert 0 < float(y_u.sum()) < 2 * (n - n_l)                              # the threshold discriminates


@pytest.mark.parametrize("n,n_l,cap", [(1076, 21, 1088), (2022, 40, 2048), (11460, 229, 11520), (300, 300, 320), (64, 0, 64)])
def test_device_count_label_propagation_directly_vs_oracle(n, n_l, cap):
    """`wesup_label_propagate_dev` (the kernel inside the CUDA-graph training step) against the oracle, not against
    another CUDA kernel: rows [n_l, n) carry the pseudo-labels, every other row of the capacity buffer is zero."""
    g = torch.Generator().manual_seed(n * 3 + n_l)
    f = torch.zeros(cap, 32)
    f[:n] = (torch.randn(n, 32, generator=g) * 0.06).abs()
    f[n:] = 0.123                                                            # garbage beyond the true count must be ignored
    y_full = torch.zeros(cap, 2)
    if n_l:
        y_full[torch.arange(n_l), torch.randint(0, 2, (n_l,), generator=g)] = 1
    y_full[n_l:] = 0.5                                                       # rows >= n_labeled of y_l must be ignored
    counts = torch.tensor([n, n_l], dtype=torch.int32, device=DEV)
    out = ops.label_propagate_static(f.to(DEV), y_full.to(DEV), counts, 0.8).cpu()
    assert out.shape == (cap, 2)
    assert float(out[:n_l].abs().sum()) == 0.0 and float(out[n:].abs().sum()) == 0.0
    if n_l == 0 or n_l == n:
        assert float(out.abs().sum()) == 0.0
        return
    y_ref, src_ref, sim_ref = O.label_propagate_block(f[:n], y_full[:n_l], 0.8)
    rows = out[n_l:n]
    d2 = torch.cdist(f[n_l:n].double(), f[:n_l].double()) ** 2
    top2 = torch.exp(-d2).topk(min(2, n_l), dim=1).values
    clear = ((top2[:, 0] - top2[:, -1] > 4 * 1.2e-7 * top2[:, 0]) if n_l > 1 else torch.ones(n - n_l, dtype=torch.bool)) \
        & ((top2[:, 0] - 0.8).abs() > 4 * 1.2e-7)
    assert float(clear.float().mean()) > 0.99
    assert torch.equal(rows[clear], y_ref[clear])


def test_training_step_at_glas_size_matches_the_dense_oracle_on_the_gpu():
    """One full step (forward, loss with propagation, backward) on a 522x775 image: the product path (pool-first,
    footprints) against the oracle's dense formulation (dense sp_maps 3.3 GB + hypercolumn 3.4 GB + mm) executed by
    torch on the same GPU in fp32 with TF32 off."""
    from wesup_b200.models import initialize_trainer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    h, w = SHAPES["glas"]
    trainer = initialize_trainer("wesup", device=DEV, pretrained=False, materialize_hypercolumn=False)
    O.seeded_init_(trainer.model, seed=11)
    img, pixel_mask, point_mask = synth.sample(h, w, index=5, ratio=1e-3)
    (x, sp), (pm, sp_labels) = trainer.preprocess(img, pixel_mask, point_mask)
    labels = sp.order.long()[sp.row_labels.long()].view(h, w)
    ref_model = O.seeded_init_(O.RefWESUP(), seed=11).to(DEV)
    order, ref_labels, _ = O.superpixel_order_and_labels_counts(labels.cpu(), point_mask[0])
    assert torch.equal(order, sp.order.cpu().long()) and torch.equal(ref_labels, sp_labels.cpu())
    maps = O.dense_sp_maps(labels, order.to(DEV))
    ref_pred = ref_model((x, maps))
    ref_feats = ref_model.sp_features
    ref_sp_pred = ref_model.sp_pred
    n_l = ref_labels.size(0)
    # the oracle's own propagation needs the (N,N,D) affinity (0.5 GB at N=2022): fine on the GPU
    y_u_ref = O.label_propagate(ref_feats.cpu(), ref_labels, 0.8)
    ref_loss_total = O.cross_entropy(ref_sp_pred[:n_l], ref_labels.to(DEV)) + 0.5 * O.cross_entropy(ref_sp_pred[n_l:], y_u_ref.to(DEV))
    ref_loss_total.backward()
    del maps
    pred = trainer.model((x, sp))
    metrics = {}
    loss = trainer.compute_loss(pred, (pm, sp_labels), metrics=metrics)
    loss.backward()
    rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))   # noqa: E731
    assert rel(trainer.model.sp_features.detach(), ref_feats.detach()) < 1e-4
    assert rel(pred.detach(), ref_pred.detach().to(DEV)) < 1e-4
    assert abs(float(loss) - float(ref_loss_total)) <= 1e-4 * abs(float(ref_loss_total)), (float(loss), float(ref_loss_total))
    for name in ("fc_layers.0.weight", "classifier.0.weight", "side_conv0.weight", "backbone.0.weight", "backbone.28.weight"):
        got = dict(trainer.model.named_parameters())[name].grad
        want = dict(ref_model.named_parameters())[name].grad
        assert rel(got, want) < 2e-3, name
    torch.backends.cudnn.allow_tf32 = True
