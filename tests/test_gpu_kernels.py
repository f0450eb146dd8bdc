"""GPU parity tests: every kernel behind the C ABI against the CPU oracle
(oracle/wesup_ref.py, oracle/slic_ref.c) and the golden vectors minted from the
real reference.  Tolerances follow BASELINE.json's north star: index/assignment
results bit-exact, pooled features and loss 1e-4 relative in fp32 (1e-2 bf16),
SLIC >= 99 % pixel agreement."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import wesup_ref as O                      # noqa: E402
from oracle import slic as oslic                       # noqa: E402
from wesup_b200 import ops, synth                      # noqa: E402
from wesup_b200.ops import SuperpixelMaps              # noqa: E402

DEV = "cuda"
VGG_C = [32, 32, 64, 64, 128, 128, 128, 256, 256, 256, 256, 256, 256]
VGG_SHIFT = [0, 0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4]


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a.detach() - b.detach()).norm() / (b.detach().norm() + 1e-30))


def make_sides(h, w, seed=0, channels=VGG_C, shifts=VGG_SHIFT):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(1, c, h >> s, w >> s, generator=g) for c, s in zip(channels, shifts)]


# ---------------------------------------------------------------------------
# sp_stats (a2)
# ---------------------------------------------------------------------------
def check_stats(seg_np, mask_t):
    seg = torch.from_numpy(seg_np)
    maps, labels, order = O.preprocess_superpixels(seg, mask_t)
    sp = SuperpixelMaps.from_labels(seg.to(DEV), None if mask_t is None else mask_t.to(DEV))
    assert sp.order.cpu().tolist() == order.tolist()                      # bit-exact ordering
    counts = (maps > 0).sum(dim=(1, 2))
    assert sp.counts.cpu().tolist() == counts.tolist()
    owner = maps.argmax(dim=0)
    assert torch.equal(sp.label_map.cpu().long(), owner)
    offs = sp.seg_offsets.cpu().long()
    assert offs[0] == 0 and offs[-1] == seg.numel()
    assert torch.equal(offs[1:] - offs[:-1], counts)
    px = sp.seg_pixels.cpu().long()
    flat_owner = owner.reshape(-1)
    for k in range(sp.n):
        mine = px[offs[k]:offs[k + 1]]
        assert torch.all(flat_owner[mine] == k)
        assert torch.all(mine[1:] > mine[:-1])                            # ascending => deterministic sums
    if mask_t is None:
        assert sp.sp_labels is None
    else:
        assert sp.n_labeled == labels.size(0)
        assert torch.equal(sp.sp_labels.cpu(), labels)                     # bit-exact multi-hot labels
    return sp


def test_stats_kat_and_golden(golden):
    g = golden("kat_preprocess_4x4.npz")
    sp = check_stats(g["segments"], torch.from_numpy(g["mask"]))
    assert sp.order.cpu().tolist() == [0, 2, 1, 3]
    assert sp.sp_labels.cpu().tolist() == [[1.0, 1.0], [0.0, 1.0]]
    np.testing.assert_allclose(sp.to_dense().cpu().numpy(), g["sp_maps"], rtol=0, atol=0)
    c = golden("preprocess_cases.npz")
    for i in range(3):
        sp = check_stats(c[f"seg{i}"], torch.from_numpy(c[f"mask{i}"]))
        assert sp.order.cpu().tolist() == c[f"order{i}"].tolist()
        assert sp.counts.cpu().tolist() == c[f"counts{i}"].tolist()
        np.testing.assert_array_equal(sp.sp_labels.cpu().numpy(), c[f"labels{i}"])
        sp_n = check_stats(c[f"seg{i}"], None)
        assert sp_n.order.cpu().tolist() == c[f"order_none{i}"].tolist()


@pytest.mark.parametrize("h,w,n", [(1, 7, 3), (5, 1, 2), (37, 53, 40), (64, 64, 1000), (128, 96, 7)])
def test_stats_random_label_maps(h, w, n):
    rng = np.random.default_rng(h * 1000 + w)
    seg = rng.integers(0, n, size=(h, w))
    _, seg = np.unique(seg, return_inverse=True)        # contiguous ids, every id present
    seg = seg.reshape(h, w)
    k = int(seg.max()) + 1
    if k < 2:
        pytest.skip("reference needs N >= 2 (squeeze at models/wesup.py:58)")
    mask = torch.zeros(3, h, w, dtype=torch.int64)
    pick = rng.random((h, w)) < 0.2
    cls = rng.integers(0, 3, size=(h, w))
    for c in range(3):
        mask[c][torch.from_numpy(pick & (cls == c))] = 1
    check_stats(seg, mask)
    check_stats(seg, None)


def test_stats_full_mask_and_single_pixel_superpixels():
    seg = np.arange(6 * 9).reshape(6, 9)                # every superpixel is one pixel
    _, gland = synth.he_like_image(6, 9, seed=5)
    check_stats(seg, synth.pixel_mask(gland))


# ---------------------------------------------------------------------------
# pooling (a4) + paint (a6)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["hwc", "chw"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_pool_fwd_bwd_vs_dense_mm(layout, dtype, tol):
    h, w, c = 45, 61, 132
    seg = synth.perturbed_grid_segments(h, w, 7, seed=3)
    maps, _, _ = O.preprocess_superpixels(torch.from_numpy(seg), None)
    g = torch.Generator().manual_seed(1)
    feat_chw = torch.randn(c, h, w, generator=g)
    ref_in = feat_chw.to(dtype).float().clone().requires_grad_(True)   # oracle sees the same rounded inputs
    ref = O.pool_dense(maps, ref_in)
    grad = torch.randn(ref.shape, generator=g)
    ref.backward(grad)
    sp = SuperpixelMaps.from_labels(torch.from_numpy(seg).to(DEV))
    if layout == "hwc":
        x = feat_chw.permute(1, 2, 0).reshape(h * w, c).to(DEV, dtype).contiguous().requires_grad_(True)
    else:
        x = feat_chw.to(DEV, dtype).contiguous().requires_grad_(True)
    out = ops.sp_pool(x, sp, layout=layout)
    assert out.dtype == torch.float32 and out.shape == ref.shape
    assert rel_err(out.cpu(), ref.detach()) < max(tol, 2e-6)
    out.backward(grad.to(DEV))
    gx = x.grad.float().cpu()
    gx = gx.view(h, w, c).permute(2, 0, 1) if layout == "hwc" else gx
    assert rel_err(gx, ref_in.grad) < tol
    # run twice: the gather is deterministic, results must be bit-identical
    out2 = ops.sp_pool(x.detach(), sp, layout=layout)
    assert torch.equal(out2, out.detach())


def test_pool_large_properties():
    """Full 464x464 x 2112 shape: size-independent properties instead of the dense
    oracle (the reference's one-hot mm would need 0.93 GB + 0.98 TFLOP on the CPU)."""
    h = w = 464
    c = 2112
    seg = synth.perturbed_grid_segments(h, w, 14, seed=11)
    sp = SuperpixelMaps.from_labels(torch.from_numpy(seg).to(DEV))
    feat = torch.randn(h * w, c, device=DEV)
    pooled = ops.sp_pool(feat, sp)
    # (1) count-weighted sum of means == global sum (linearity / partition of unity)
    total = (pooled.double() * sp.counts.double().unsqueeze(1)).sum(0)
    assert rel_err(total, feat.double().sum(0)) < 1e-6
    # (2) a feature that is constant inside every superpixel pools to itself, exactly
    const = sp.row_labels.float().unsqueeze(1).expand(-1, 8).contiguous()
    pc = ops.sp_pool(const, sp)
    assert rel_err(pc, torch.arange(sp.n, device=DEV, dtype=torch.float32).unsqueeze(1).expand(-1, 8)) < 1e-6
    # (3) spot-check 16 rows against a direct mean
    lab = sp.row_labels.long()
    for k in np.random.default_rng(0).integers(0, sp.n, 16):
        ref = feat[lab == int(k)].double().mean(0)
        assert rel_err(pooled[int(k)], ref) < 1e-6
    # (4) adjoint identity <pool(x), g> == <x, pool^T(g)>
    x = feat[:, :64].contiguous().requires_grad_(True)
    gp = torch.randn(sp.n, 64, device=DEV)
    y = ops.sp_pool(x, sp)
    y.backward(gp)
    lhs = (y.detach().double() * gp.double()).sum()
    rhs = (x.detach().double() * x.grad.double()).sum()
    scale = float((y.detach().double() * gp.double()).abs().sum())      # fp32 rounding scales with the terms, not their sum
    assert abs(float(lhs - rhs)) < 1e-6 * scale + 1e-6


def test_paint_matches_reference_loop():
    h, w = 40, 56
    seg = synth.perturbed_grid_segments(h, w, 8, seed=10)
    maps, _, _ = O.preprocess_superpixels(torch.from_numpy(seg), None)
    pred = torch.softmax(torch.randn(maps.size(0), 2, generator=torch.Generator().manual_seed(2)), dim=1)
    ref = O.paint_dense(maps, pred)
    sp = SuperpixelMaps.from_labels(torch.from_numpy(seg).to(DEV))
    out = ops.paint(sp, pred.to(DEV), cls=1)
    assert out.shape == ref.shape
    assert torch.equal(out.cpu(), ref)                   # pure gather: bit-exact


def test_dense_sp_maps_are_accepted():
    seg = synth.perturbed_grid_segments(24, 30, 6, seed=4)
    maps, _, _ = O.preprocess_superpixels(torch.from_numpy(seg), None)
    sp = SuperpixelMaps.from_dense(maps.to(DEV))
    assert sp.n == maps.size(0)
    assert torch.equal(sp.label_map.cpu().long(), maps.argmax(0))


# ---------------------------------------------------------------------------
# hypercolumn (a3)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["hwc", "chw"])
@pytest.mark.parametrize("h,w", [(48, 40), (37, 51), (16, 16)])
def test_hypercolumn_fwd_bwd_vs_interpolate(layout, h, w):
    sides = make_sides(h, w, seed=h)
    ref_in = [s.clone().requires_grad_(True) for s in sides]
    ref = O.hypercolumn_from_sides(ref_in, (h, w))                       # (C,H,W)
    grad = torch.randn(ref.shape, generator=torch.Generator().manual_seed(9))
    ref.backward(grad)
    xs = [s.to(DEV).requires_grad_(True) for s in sides]
    out = ops.hypercolumn(xs, (h, w), layout=layout)
    got = out.detach().cpu().view(h, w, -1).permute(2, 0, 1) if layout == "hwc" else out.detach().cpu()
    assert rel_err(got, ref.detach()) < 1e-6
    np.testing.assert_allclose(got.numpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-5)
    # identity levels are exact copies
    assert torch.equal(got[:32], sides[0][0])
    g_dev = grad.permute(1, 2, 0).reshape(h * w, -1).contiguous() if layout == "hwc" else grad
    out.backward(g_dev.to(DEV))
    for x, r in zip(xs, ref_in):
        assert rel_err(x.grad.cpu(), r.grad) < 1e-5


def test_hypercolumn_bf16_and_golden_probe(golden):
    g = golden("forward_loss_backward_48x40.npz")
    model = O.seeded_init_(O.RefWESUP(), seed=3)
    x = synth.to_tensor(g["img_u8"]).unsqueeze(0)
    with torch.no_grad():
        sides = model.side_outputs(x)
    out = ops.hypercolumn([s.to(DEV) for s in sides], (48, 40))
    got = out.cpu().view(48, 40, -1).permute(2, 0, 1)
    np.testing.assert_allclose(got[:, ::7, ::5].numpy(), g["feats_probe"], rtol=1e-4, atol=1e-5)
    out16 = ops.hypercolumn([s.to(DEV) for s in sides], (48, 40), dtype=torch.bfloat16)
    assert out16.dtype == torch.bfloat16
    assert rel_err(out16.float().cpu(), out.cpu()) < 1e-2


def test_hypercolumn_adjoint_identity_full_size():
    h = w = 464
    xs = [s.to(DEV).requires_grad_(True) for s in make_sides(h, w, seed=1)]
    out = ops.hypercolumn(xs, (h, w))
    assert out.shape == (h * w, 2112)
    g = torch.randn_like(out)
    out.backward(g)
    lhs = (out.detach().double() * g.double()).sum()
    rhs = sum((x.detach().double() * x.grad.double()).sum() for x in xs)
    assert abs(float(lhs - rhs)) < 1e-6 * abs(float(lhs)) + 1e-3
    # partition of unity: constant sides upsample to the same constant
    ones = [torch.full_like(x, 2.5) for x in xs]
    const = ops.hypercolumn(ones, (h, w))
    assert float((const - 2.5).abs().max()) < 1e-5


@pytest.mark.parametrize("h,w,n", [(48, 40, 30), (37, 51, 7), (16, 16, 256), (131, 97, 60)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_pool_hypercolumn_backward_vs_autograd_of_dense_reference(h, w, n, dtype):
    """d(side outputs) of `mm(sp_maps, cat(interpolate(side)))` from the fused
    kernel vs torch autograd of the reference's dense formulation on the CPU, and
    vs the unfused pool_bwd -> hypercolumn_bwd pair."""
    gen = torch.Generator().manual_seed(h * w + n)
    sides = make_sides(h, w, seed=w)
    seg = torch.randint(0, n, (h, w), generator=gen)
    seg.view(-1)[:n] = torch.arange(n)                                    # every id present
    maps, _, _ = O.preprocess_superpixels(seg, None)
    ref_in = [s.clone().requires_grad_(True) for s in sides]
    feats = O.hypercolumn_from_sides(ref_in, (h, w))
    pooled_ref = O.pool_dense(maps, feats)
    gp = torch.randn(pooled_ref.shape, generator=gen)
    pooled_ref.backward(gp)
    sp = SuperpixelMaps.from_labels(seg.to(DEV))
    xs = [s.to(DEV).requires_grad_(True) for s in sides]
    pooled, hc = ops.hypercolumn_pool(xs, (h, w), sp, dtype=dtype)
    assert hc.shape == (h * w, 2112) and hc.dtype == dtype and not hc.requires_grad
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(pooled.cpu(), pooled_ref) < tol
    pooled.backward(gp.to(DEV))
    for x, r in zip(xs, ref_in):
        assert rel_err(x.grad.cpu(), r.grad) < 1e-5                       # backward never touches the bf16 tensor
    if dtype == torch.float32:
        ys = [s.to(DEV).requires_grad_(True) for s in sides]
        ops.sp_pool(ops.hypercolumn(ys, (h, w)), sp).backward(gp.to(DEV))
        for x, y in zip(xs, ys):
            assert rel_err(x.grad, y.grad) < 1e-6


@pytest.mark.parametrize("h,w,n", [(48, 40, 30), (37, 51, 7), (131, 97, 60), (464, 464, 1076)])
def test_fused_forward_equals_hypercolumn_then_pool(h, w, n):
    """Pooling straight from the side outputs == kernel (a) then kernel (b) (same arithmetic, same order)."""
    gen = torch.Generator().manual_seed(h + w + n)
    seg = torch.from_numpy(synth.perturbed_grid_segments(h, w, max(2, int((h * w / n) ** 0.5)), seed=n)).long()
    sp = SuperpixelMaps.from_labels(seg.to(DEV))
    xs = [s.to(DEV).requires_grad_(True) for s in make_sides(h, w, seed=n)]
    ys = [x.detach().clone().requires_grad_(True) for x in xs]
    pooled_a, feats = ops.hypercolumn_pool(xs, (h, w), sp)
    pooled_b, none = ops.hypercolumn_pool(ys, (h, w), sp, materialize=False)
    assert none is None and feats is not None
    assert rel_err(pooled_b, pooled_a) < 1e-6
    np.testing.assert_allclose(pooled_b.detach().cpu().numpy(), pooled_a.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    g = torch.randn(pooled_a.shape, generator=gen).to(DEV)
    pooled_a.backward(g)
    pooled_b.backward(g)
    for x, y in zip(xs, ys):
        assert torch.equal(x.grad, y.grad)                                   # same fused backward either way


FULL_C = [64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512]      # backbone conv outputs ("pool first")


def _levels_call(name, *args):
    from wesup_b200 import _lib
    _lib.check(getattr(_lib.load(), name)(*args), name)


@pytest.mark.parametrize("h,w,n,channels", [(48, 40, 30, VGG_C), (131, 97, 60, VGG_C), (96, 112, 50, FULL_C),
                                            (464, 464, 1076, FULL_C), (200, 180, 3, VGG_C)])
def test_footprint_kernels_match_the_per_pixel_walk_kernels(h, w, n, channels):
    """wesup_levels_pool_fwd/bwd (weights aggregated per low-res cell) against the independent per-pixel
    walk formulation of the same operators, on SLIC-like grids, on a few huge superpixels (bounding box
    larger than the shared grid => per-pixel path inside the kernel) and with the 4224 backbone channels."""
    from wesup_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    seg = torch.from_numpy(synth.perturbed_grid_segments(h, w, max(2, int((h * w / n) ** 0.5)), seed=n)).long()
    sp = SuperpixelMaps.from_labels(seg.to(DEV))
    sides = [s.to(DEV).permute(0, 2, 3, 1).contiguous() for s in make_sides(h, w, seed=n, channels=channels)]
    C, hs, ws_ = [s.size(3) for s in sides], [s.size(1) for s in sides], [s.size(2) for s in sides]
    ca, ha, wa = _lib.int_array(C), _lib.int_array(hs), _lib.int_array(ws_)
    ptrs = _lib.ptr_array([s.data_ptr() for s in sides])
    ctot = sum(C)
    a = torch.empty(sp.n, ctot, device=DEV)
    b = torch.empty(sp.n, ctot, device=DEV)
    _levels_call("wesup_levels_pool_fwd", ptrs, ca, ha, wa, len(C), h, w, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),
                 sp.n, a.data_ptr(), st)
    _levels_call("wesup_hypercolumn_pool_fwd_walk", ptrs, ca, ha, wa, len(C), h, w, sp.seg_offsets.data_ptr(),
                 sp.seg_pixels.data_ptr(), sp.n, b.data_ptr(), st)
    tol = 1e-6 if h * w / sp.n < 2000 else 1e-5          # fp32 sums over 12 000-pixel superpixels in two different orders
    assert rel_err(a, b) < tol
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=2e-6)
    gp = torch.randn(sp.n, ctot, device=DEV)
    ga = [torch.full_like(s, float("nan")) for s in sides]          # every element must be overwritten
    gb = [torch.empty_like(s) for s in sides]
    wsl = torch.empty(lib.wesup_levels_pool_bwd_workspace_bytes(ca, ha, wa, len(C), h, w), dtype=torch.uint8, device=DEV)
    _levels_call("wesup_levels_pool_bwd", gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(), ca, ha, wa, len(C), h, w,
                 sp.n, _lib.ptr_array([t.data_ptr() for t in ga]), wsl.data_ptr(), st)
    wsb = torch.empty(lib.wesup_sp_pool_hypercolumn_bwd_workspace_bytes(ca, ha, wa, len(C), h, w, sp.n), dtype=torch.uint8, device=DEV)
    _levels_call("wesup_sp_pool_hypercolumn_bwd_walk", gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(), ca, ha, wa,
                 len(C), h, w, sp.n, _lib.ptr_array([t.data_ptr() for t in gb]), wsb.data_ptr(), st)
    for x, y in zip(ga, gb):
        assert torch.isfinite(x).all()
        assert rel_err(x, y) < tol
    # adjoint identity of the footprint pair itself
    lhs = (a.double() * gp.double()).sum()
    rhs = sum((s.double() * g.double()).sum() for s, g in zip(sides, ga))
    scale = float((a.double() * gp.double()).abs().sum())
    assert abs(float(lhs - rhs)) < 1e-6 * scale + 1e-6
    # run-to-run determinism (fixed summation order, no floating-point atomics)
    a2 = torch.empty_like(a)
    _levels_call("wesup_levels_pool_fwd", ptrs, ca, ha, wa, len(C), h, w, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),
                 sp.n, a2.data_ptr(), st)
    assert torch.equal(a, a2)


def _scattered_segments(h, w, n, seed):
    """Every pixel draws its superpixel at random: footprints meet > 64 superpixels (slow path of the
    backward list builder) and bounding boxes cover the image (per-pixel entries in the forward lists)."""
    g = torch.Generator().manual_seed(seed)
    seg = torch.randint(0, n, (h, w), generator=g)
    seg.view(-1)[:n] = torch.arange(n)
    return seg


@pytest.mark.parametrize("h,w,n,channels,kind", [(48, 40, 30, VGG_C, "grid"), (131, 97, 60, VGG_C, "grid"),
                                                 (96, 112, 50, FULL_C, "grid"), (464, 464, 1076, FULL_C, "grid"),
                                                 (200, 180, 3, VGG_C, "grid"), (80, 72, 700, VGG_C, "scattered"),
                                                 (37, 51, 7, [8, 20, 36], "grid")])
def test_precomputed_footprint_pooling_matches_the_in_kernel_footprint_kernels(h, w, n, channels, kind):
    """wesup_footprint_build + wesup_levels_pool_fwd_fp/bwd_fp (lists built once per label map) against
    wesup_levels_pool_fwd/bwd (same weights rebuilt inside the kernels): equal up to fp32 summation order,
    adjoint of each other, bit-reproducible, and rows beyond the true superpixel count pool to zero."""
    from wesup_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    if kind == "grid":
        seg = torch.from_numpy(synth.perturbed_grid_segments(h, w, max(2, int((h * w / n) ** 0.5)), seed=n)).long()
    else:
        seg = _scattered_segments(h, w, n, seed=n)
    sp = SuperpixelMaps.from_labels(seg.to(DEV))
    shifts = VGG_SHIFT if len(channels) == 13 else [0, 1, 3]
    sides = [s.to(DEV).permute(0, 2, 3, 1).contiguous() for s in make_sides(h, w, seed=n, channels=channels, shifts=shifts)]
    C, hs, ws_ = [s.size(3) for s in sides], [s.size(1) for s in sides], [s.size(2) for s in sides]
    ca, ha, wa = _lib.int_array(C), _lib.int_array(hs), _lib.int_array(ws_)
    ptrs = _lib.ptr_array([s.data_ptr() for s in sides])
    ctot, nl = sum(C), len(C)
    # fixed-capacity call: 5 extra rows with empty CSR segments, as the CUDA-graph iteration uses
    cap = sp.n + 5
    offs = torch.full((cap + 1,), h * w, dtype=torch.int32, device=DEV)
    offs[:sp.n + 1] = sp.seg_offsets
    counts = torch.zeros(cap, dtype=torch.int32, device=DEV)
    counts[:sp.n] = sp.counts
    nbytes = lib.wesup_footprint_bytes(ha, wa, nl, h, w, cap)
    assert nbytes > 0
    fp = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _levels_call("wesup_footprint_build", ha, wa, nl, h, w, cap, offs.data_ptr(), sp.seg_pixels.data_ptr(),
                 sp.row_labels.data_ptr(), counts.data_ptr(), 1, fp.data_ptr(), st)
    a = torch.full((cap, ctot), float("nan"), device=DEV)
    _levels_call("wesup_levels_pool_fwd_fp", ptrs, ca, ha, wa, nl, h, w, offs.data_ptr(), sp.seg_pixels.data_ptr(), cap,
                 fp.data_ptr(), a.data_ptr(), st)
    b = torch.empty(sp.n, ctot, device=DEV)
    _levels_call("wesup_levels_pool_fwd", ptrs, ca, ha, wa, nl, h, w, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),
                 sp.n, b.data_ptr(), st)
    assert torch.isfinite(a).all() and float(a[sp.n:].abs().max()) == 0.0
    tol = 1e-6 if h * w / sp.n < 2000 else 1e-5
    assert rel_err(a[:sp.n], b) < tol
    # the library has two forward kernels over the lists (whole cells per warp for 32..512-channel levels, 128-channel
    # chunks otherwise); WESUP_FP_FWD=chunks forces the second one: same sums in another order
    import os
    os.environ["WESUP_FP_FWD"] = "chunks"
    try:
        a_chunks = torch.full((cap, ctot), float("nan"), device=DEV)
        _levels_call("wesup_levels_pool_fwd_fp", ptrs, ca, ha, wa, nl, h, w, offs.data_ptr(), sp.seg_pixels.data_ptr(), cap,
                     fp.data_ptr(), a_chunks.data_ptr(), st)
    finally:
        os.environ.pop("WESUP_FP_FWD")
    assert torch.isfinite(a_chunks).all() and rel_err(a_chunks, a) < tol
    # large superpixels take several warps per (superpixel, level) list (shared-memory reduction in warp order)
    for split in ("4", "16"):
        os.environ["WESUP_FP_FWD_SPLIT"] = split
        try:
            a_split = torch.full((cap, ctot), float("nan"), device=DEV)
            _levels_call("wesup_levels_pool_fwd_fp", ptrs, ca, ha, wa, nl, h, w, offs.data_ptr(), sp.seg_pixels.data_ptr(), cap,
                         fp.data_ptr(), a_split.data_ptr(), st)
        finally:
            os.environ.pop("WESUP_FP_FWD_SPLIT")
        assert torch.isfinite(a_split).all() and float(a_split[sp.n:].abs().max()) == 0.0 and rel_err(a_split, a) < tol
    np.testing.assert_allclose(a[:sp.n].cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=2e-6)
    gp = torch.randn(cap, ctot, device=DEV)
    ga = [torch.full_like(s, float("nan")) for s in sides]          # every element must be overwritten
    gb = [torch.empty_like(s) for s in sides]
    _levels_call("wesup_levels_pool_bwd_fp", gp.data_ptr(), sp.row_labels.data_ptr(), counts.data_ptr(), ca, ha, wa, nl, h, w, cap,
                 fp.data_ptr(), _lib.ptr_array([t.data_ptr() for t in ga]), st)
    wsl = torch.empty(lib.wesup_levels_pool_bwd_workspace_bytes(ca, ha, wa, nl, h, w), dtype=torch.uint8, device=DEV)
    _levels_call("wesup_levels_pool_bwd", gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(), ca, ha, wa, nl, h, w,
                 sp.n, _lib.ptr_array([t.data_ptr() for t in gb]), wsl.data_ptr(), st)
    for x, y in zip(ga, gb):
        assert torch.isfinite(x).all()
        assert rel_err(x, y) < tol
    os.environ["WESUP_FP_BWD"] = "chunks"                    # the second backward kernel over the lists, likewise
    try:
        gc = [torch.full_like(s, float("nan")) for s in sides]
        _levels_call("wesup_levels_pool_bwd_fp", gp.data_ptr(), sp.row_labels.data_ptr(), counts.data_ptr(), ca, ha, wa, nl, h, w,
                     cap, fp.data_ptr(), _lib.ptr_array([t.data_ptr() for t in gc]), st)
    finally:
        os.environ.pop("WESUP_FP_BWD")
    for x, y in zip(gc, ga):
        assert torch.isfinite(x).all() and rel_err(x, y) < tol
    os.environ["WESUP_FP_BWD"] = "split"                     # default = one merged launch; "split" = cells + identity kernels
    try:
        gs = [torch.full_like(s, float("nan")) for s in sides]
        _levels_call("wesup_levels_pool_bwd_fp", gp.data_ptr(), sp.row_labels.data_ptr(), counts.data_ptr(), ca, ha, wa, nl, h, w,
                     cap, fp.data_ptr(), _lib.ptr_array([t.data_ptr() for t in gs]), st)
    finally:
        os.environ.pop("WESUP_FP_BWD")
    for x, y in zip(gs, ga):
        assert torch.equal(x, y)                             # same arithmetic per element, only the launch shape differs
    lhs = (a.double() * gp.double()).sum()
    rhs = sum((s_.double() * g.double()).sum() for s_, g in zip(sides, ga))
    scale = float((a.double() * gp.double()).abs().sum())
    assert abs(float(lhs - rhs)) < 1e-6 * scale + 1e-6
    # a rebuilt blob may place its lists elsewhere (atomic cursor); the results must not change by a bit
    fp2 = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _levels_call("wesup_footprint_build", ha, wa, nl, h, w, cap, offs.data_ptr(), sp.seg_pixels.data_ptr(),
                 sp.row_labels.data_ptr(), counts.data_ptr(), 1, fp2.data_ptr(), st)
    a2 = torch.empty_like(a)
    _levels_call("wesup_levels_pool_fwd_fp", ptrs, ca, ha, wa, nl, h, w, offs.data_ptr(), sp.seg_pixels.data_ptr(), cap,
                 fp2.data_ptr(), a2.data_ptr(), st)
    ga2 = [torch.empty_like(s) for s in sides]
    _levels_call("wesup_levels_pool_bwd_fp", gp.data_ptr(), sp.row_labels.data_ptr(), counts.data_ptr(), ca, ha, wa, nl, h, w, cap,
                 fp2.data_ptr(), _lib.ptr_array([t.data_ptr() for t in ga2]), st)
    assert torch.equal(a, a2)
    for x, y in zip(ga, ga2):
        assert torch.equal(x, y)


@pytest.mark.parametrize("h,w,n", [(37, 51, 7), (131, 97, 60), (464, 464, 1076)])
def test_hypercolumn_pool_over_footprints_vs_dense_reference(h, w, n):
    """`ops.hypercolumn_pool(..., footprints=build_footprints(...))` with the build forked onto a side
    stream: forward and autograd gradients against the reference's dense formulation on the CPU."""
    gen = torch.Generator().manual_seed(h * w + n)
    sides = make_sides(h, w, seed=w)
    seg = torch.from_numpy(synth.perturbed_grid_segments(h, w, max(2, int((h * w / n) ** 0.5)), seed=n)).long()
    sp = SuperpixelMaps.from_labels(seg.to(DEV))
    xs = [s.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for s in sides]
    side_stream = torch.cuda.Stream()
    fp = ops.build_footprints(sp, [(s.size(2), s.size(3)) for s in sides], with_bwd=True, stream=side_stream)
    pooled, none = ops.hypercolumn_pool(xs, (h, w), sp, materialize=False, footprints=fp)
    assert none is None
    gp = torch.randn(pooled.shape, generator=gen)
    pooled.backward(gp.to(DEV))
    if h * w <= 20000:
        maps, _, _ = O.preprocess_superpixels(seg, None)
        ref_in = [s.clone().requires_grad_(True) for s in sides]
        pooled_ref = O.pool_dense(maps, O.hypercolumn_from_sides(ref_in, (h, w)))
        pooled_ref.backward(gp)
        assert rel_err(pooled.cpu(), pooled_ref) < 1e-5
        for x, r in zip(xs, ref_in):
            assert rel_err(x.grad.cpu(), r.grad) < 1e-5
    ys = [s.to(DEV).requires_grad_(True) for s in sides]
    pooled_b, _ = ops.hypercolumn_pool(ys, (h, w), sp)                 # kernel (a) then kernel (b), fused backward
    pooled_b.backward(gp.to(DEV))
    assert rel_err(pooled, pooled_b) < 1e-6
    for x, y in zip(xs, ys):
        assert rel_err(x.grad, y.grad) < 1e-6
    with pytest.raises(ValueError):
        other = SuperpixelMaps.from_labels((seg // 2).to(DEV))
        ops.hypercolumn_pool(xs, (h, w), other, materialize=False, footprints=fp)


def test_fused_backward_full_size_adjoint_identity():
    """<pool(hyper(x)), g> == <x, fused_bwd(g)> at 464^2 with a SLIC-like grid of superpixels."""
    h = w = 464
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    seg = (yy // 14) * ((w + 13) // 14) + xx // 14
    sp = SuperpixelMaps.from_labels(seg.to(DEV))
    xs = [s.to(DEV).requires_grad_(True) for s in make_sides(h, w, seed=2)]
    pooled, _ = ops.hypercolumn_pool(xs, (h, w), sp)
    g = torch.randn_like(pooled)
    pooled.backward(g)
    lhs = (pooled.detach().double() * g.double()).sum()
    rhs = sum((x.detach().double() * x.grad.double()).sum() for x in xs)
    assert abs(float(lhs - rhs)) < 1e-5 * abs(float(lhs)) + 1e-3


# ---------------------------------------------------------------------------
# label propagation (a7)
# ---------------------------------------------------------------------------
def check_propagation(f, y_l, thr):
    """Bit-exact in src index and propagated/not decision wherever the decision
    is numerically unambiguous; rows whose top-2 similarities (or whose
    similarity and the threshold) are closer than 4 fp32 ulp may legitimately
    differ between two fp32 evaluation orders (the reference's own einsum order
    is backend-dependent) and are only required to pick one of the tied answers."""
    y_u_ref, src_ref, sim_ref = O.label_propagate(f, y_l, thr, return_aux=True)
    y_u, src, sim = ops.label_propagate(f.to(DEV), y_l.to(DEV), thr, return_aux=True)
    y_u, src, sim = y_u.cpu(), src.cpu().long(), sim.cpu()
    n_l = y_l.size(0)
    d2 = torch.cdist(f[n_l:].double(), f[:n_l].double()) ** 2
    w_all = torch.exp(-d2)
    top2 = w_all.topk(min(2, n_l), dim=1).values
    eps = 4 * 1.2e-7
    ambiguous_src = (top2[:, 0] - top2[:, -1] < eps * top2[:, 0]) if n_l > 1 else torch.zeros(len(src), dtype=torch.bool)
    ambiguous_thr = (top2[:, 0] - thr).abs() < eps
    clear = ~ambiguous_src
    assert torch.equal(src[clear], src_ref[clear])
    tied = torch.nonzero(ambiguous_src).flatten()
    for u in tied.tolist():
        assert w_all[u, src[u]] >= top2[u, 0] - eps
    ok_rows = clear & ~ambiguous_thr
    assert torch.equal(y_u[ok_rows], y_u_ref[ok_rows])
    assert torch.allclose(sim, sim_ref, rtol=1e-5, atol=1e-7)
    return int(clear.sum()), len(src)


def test_label_propagate_golden(golden):
    g = golden("label_propagate_cases.npz")
    for i in range(4):
        f, yl = torch.from_numpy(g[f"f{i}"]), torch.from_numpy(g[f"yl{i}"])
        for thr in (0.8, 0.95):
            y_u = ops.label_propagate(f.to(DEV), yl.to(DEV), thr).cpu()
            np.testing.assert_array_equal(y_u.numpy(), g[f"yu{i}_{int(thr * 100)}"])
            check_propagation(f, yl, thr)


@pytest.mark.parametrize("n,n_l,scale", [(300, 1, 0.06), (1076, 21, 0.06), (2022, 40, 0.06), (1500, 750, 0.05), (4000, 2000, 0.06)])
def test_label_propagate_vs_oracle(n, n_l, scale):
    g = torch.Generator().manual_seed(n)
    f = (torch.randn(n, 32, generator=g) * scale).abs()
    y_l = torch.zeros(n_l, 2)
    y_l[torch.arange(n_l), torch.randint(0, 2, (n_l,), generator=g)] = 1
    clear, total = check_propagation(f, y_l, 0.8)
    assert clear >= 0.99 * total


def test_label_propagate_edge_cases():
    f = torch.zeros(10, 32)
    f[:, 0] = torch.arange(10).float() * 0.1
    y_l = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    # identical features: tie -> lowest labeled index
    same = torch.ones(6, 32) * 0.3
    y_u, src, sim = ops.label_propagate(same.to(DEV), y_l.to(DEV), 0.8, return_aux=True)
    assert src.cpu().tolist() == [0, 0, 0] and torch.all(sim.cpu() == 1.0)
    assert y_u.cpu().tolist() == [[1.0, 0.0]] * 3
    # sim == fp32(thr) is NOT propagated (strict >)
    thr = float(torch.tensor(1.0))
    y_u = ops.label_propagate(same.to(DEV), y_l.to(DEV), thr)
    assert float(y_u.sum()) == 0.0
    # multi-hot labeled rows are copied as they are
    f2 = torch.zeros(5, 32)
    f2[3:, 1] = 0.01
    f2[2, 1] = 0.01
    y_u = ops.label_propagate(f2.to(DEV), y_l.to(DEV), 0.5)
    assert y_u.cpu().tolist() == [[1.0, 1.0], [1.0, 1.0]]
    # no unlabeled rows -> empty result
    assert ops.label_propagate(f2[:3].to(DEV), y_l.to(DEV), 0.5).shape == (0, 2)


# ---- tensor-core path (tcgen05): must be bit-identical to the CUDA-core path ----
TC_KAPPA = 2.0 ** -16


def assert_tc_equals_exact(f, y_l, thr):
    a = ops.label_propagate(f.to(DEV), y_l.to(DEV), thr, return_aux=True, algo="exact")
    y_u, src, sim, stats = ops.label_propagate(f.to(DEV), y_l.to(DEV), thr, return_aux=True, algo="tc", return_stats=True)
    assert torch.equal(src, a[1]), f"src differs in {(src != a[1]).sum().item()} rows"
    assert torch.equal(sim, a[2])                                         # same bits, not just close
    assert torch.equal(y_u, a[0])
    assert stats["max_err_ratio"] < 0.5 * TC_KAPPA, stats                 # the filter's error bound holds with margin
    assert stats["exact_evals"] >= src.numel()
    return stats


@pytest.mark.parametrize("n,n_l,scale", [(300, 1, 0.06), (1076, 21, 0.06), (2022, 40, 0.06), (1500, 750, 0.05),
                                         (4000, 2000, 0.06), (8000, 4000, 0.06), (513, 129, 0.3), (700, 257, 1.5)])
def test_label_propagate_tc_is_bit_identical(n, n_l, scale):
    g = torch.Generator().manual_seed(n + n_l)
    f = (torch.randn(n, 32, generator=g) * scale).abs()
    y_l = torch.zeros(n_l, 2)
    y_l[torch.arange(n_l), torch.randint(0, 2, (n_l,), generator=g)] = 1
    stats = assert_tc_equals_exact(f, y_l, 0.8)
    if scale <= 0.06 and n_l >= 40:
        # separated features: the tensor-core filter leaves only a few exact re-evaluations per row
        assert stats["exact_evals"] < 0.25 * stats["pairs"], stats


def test_label_propagate_tc_collapsed_and_tied_features(golden):
    """Worst cases for the filter: every labeled row is a candidate (random-init
    collapse, SURVEY.md section 8d caveat), exact ties, duplicated labeled rows."""
    y_l = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    same = torch.ones(400, 32) * 0.3
    y3 = y_l.repeat(50, 1)[:150]
    y_u, src, sim = ops.label_propagate(same.to(DEV), y3.to(DEV), 0.8, return_aux=True, algo="tc")
    assert torch.all(src == 0) and torch.all(sim == 1.0)
    g = torch.Generator().manual_seed(5)
    base = torch.rand(1, 32, generator=g) * 3.0
    near = base + torch.randn(1200, 32, generator=g) * 1e-4              # collapsed cloud, large norms
    yl = torch.zeros(300, 2); yl[:, 0] = 1
    assert_tc_equals_exact(near.abs(), yl, 0.8)
    dup = (torch.randn(200, 32, generator=g) * 0.06).abs()
    dup = torch.cat([dup, dup, dup[:77]])                                 # every unlabeled row has an exact twin (two of them)
    yl = torch.zeros(200, 2); yl[:, 1] = 1
    y_u, src, sim = ops.label_propagate(dup.to(DEV), yl.to(DEV), 0.8, return_aux=True, algo="tc")
    assert src.cpu().tolist() == list(range(200)) + list(range(77)) and torch.all(sim == 1.0)
    gl = golden("label_propagate_cases.npz")
    for i in range(4):
        f, yl = torch.from_numpy(gl[f"f{i}"]), torch.from_numpy(gl[f"yl{i}"])
        if f.size(1) != 32 or yl.size(0) >= f.size(0):
            continue
        for thr in (0.8, 0.95):
            y_u = ops.label_propagate(f.to(DEV), yl.to(DEV), thr, algo="tc").cpu()
            np.testing.assert_array_equal(y_u.numpy(), gl[f"yu{i}_{int(thr * 100)}"])


# ---------------------------------------------------------------------------
# SLIC (a1)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("h,w,seed", [(96, 128, 1), (131, 97, 2), (464, 464, 1000)])
def test_slic_agreement_with_cpu_restatement(h, w, seed):
    img_u8, _ = synth.he_like_image(h, w, seed=seed)
    img = img_u8.astype(np.float32) / 255.0
    n_segments = int(h * w / 200)
    ref_labels, ref_raw, _, ref_n = oslic.slic(img, n_segments, 40, return_aux=True)
    x = torch.from_numpy(img.transpose(2, 0, 1).copy()).to(DEV)
    raw, k = ops.slic(x, n_segments, 40, enforce_connectivity=False)
    raw_agree = float((raw.cpu().numpy() == ref_raw).mean())
    assert raw_agree >= 0.99, f"k-means assignment agreement {raw_agree:.4f}"
    labels, n = ops.slic(x, n_segments, 40)
    labels = labels.cpu().numpy()
    # both sides number labels 0..n-1 in raster order of first pixel, so ids are directly comparable
    agree = float((labels == ref_labels).mean())
    assert agree >= 0.99, f"label agreement {agree:.4f}"
    n = int(n.item())
    assert labels.min() == 0 and labels.max() == n - 1
    assert len(np.unique(labels)) == n                     # contiguous ids, as the reference requires
    assert abs(n - ref_n) <= max(2, 0.01 * ref_n)
    first = [np.argmax(labels.reshape(-1) == i) for i in range(n)]
    assert first == sorted(first)                          # raster-order numbering


def test_slic_connectivity_matches_sequential_on_kmeans_output():
    """Feed the CPU oracle's raw k-means assignment through both connectivity
    implementations: isolates the CCL/merge kernels from fp differences."""
    h, w = 200, 232
    img_u8, _ = synth.he_like_image(h, w, seed=77)
    img = img_u8.astype(np.float32) / 255.0
    n_segments = int(h * w / 200)
    x = torch.from_numpy(img.transpose(2, 0, 1).copy()).to(DEV)
    raw, _ = ops.slic(x, n_segments, 40, enforce_connectivity=False)
    seg_size = h * w / n_segments
    ref, ref_n = oslic.connectivity(raw.cpu().numpy(), int(0.5 * seg_size), int(3 * seg_size))
    labels, n = ops.slic(x, n_segments, 40)
    assert float((labels.cpu().numpy() == ref).mean()) >= 0.999
    assert int(n.item()) == ref_n


# ---------------------------------------------------------------------------
# bias gradient of the channels_last convolutions
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("rows,c", [(1, 4), (37, 8), (215296, 64), (53824, 128), (3364, 512), (1000, 96), (841, 1024)])
def test_colsum_vs_fp64_sum(rows, c):
    g = torch.Generator().manual_seed(rows + c)
    x = torch.randn(rows, c, generator=g).to(DEV)
    out = ops.colsum(x)
    ref = x.double().sum(0)
    assert out.shape == (c,)
    assert float((out.double() - ref).abs().max()) <= 1e-5 * float(x.double().abs().sum(0).max()) + 1e-6
    assert torch.equal(out, ops.colsum(x))                          # fixed summation order
    with pytest.raises(ValueError):
        ops.colsum(torch.zeros(5, 6, device=DEV))                   # C % 4 != 0


@pytest.mark.parametrize("cin,cout,k,h,w", [(3, 64, 3, 48, 40), (64, 128, 3, 37, 51), (256, 128, 1, 29, 29)])
def test_conv2d_channels_last_matches_autograd_of_conv2d(cin, cout, k, h, w):
    """Same forward as nn.Conv2d; input and weight gradients from the same aten op; the bias gradient (wesup_colsum)
    equals autograd's grad.sum((0,2,3)) up to fp32 summation order."""
    torch.manual_seed(cin + cout)
    conv = torch.nn.Conv2d(cin, cout, k, padding=k // 2).to(DEV)
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    x = torch.randn(1, cin, h, w, device=DEV).contiguous(memory_format=torch.channels_last)
    g = torch.randn(1, cout, h, w, device=DEV).contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = conv(xa)
    ya.backward(g)
    ref = [xa.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()]
    conv.zero_grad(set_to_none=True)
    yb = ops.conv2d_channels_last(xb, conv)
    assert rel_err(yb, ya) < 1e-5
    yb.backward(g)
    assert rel_err(xb.grad, ref[0]) < 1e-4 and rel_err(conv.weight.grad, ref[1]) < 1e-4          # same aten op; cuDNN may pick another algorithm
    assert rel_err(conv.bias.grad, ref[2]) < 1e-5
