"""SLIC (/root/reference/models/wesup.py:471-478).

scikit-image is absent from the image, so parity with the real package is UNPINNED.  What is
pinned instead, in this order:
  * hand-derived known answers (tests/golden/slic_kats.npz, derivations in make_slic_kats.py);
  * two independent CPU restatements (oracle/slic_ref.c and oracle/slic_np.py) that must agree;
  * the CUDA path against the C restatement: bit-exact for the connectivity stage (integer work),
    >= 99 % pixel agreement for the full pipeline (north star), bit-exact batch == single.
CPU tests are unmarked; GPU tests carry `@pytest.mark.gpu` and call through the C ABI.
"""
import numpy as np
import pytest
import torch

from oracle import slic as oslic
from oracle import slic_np
from wesup_b200 import synth

gpu = pytest.mark.gpu
DEV = "cuda"


def kat_uniform(golden):
    g = golden("slic_kats.npz")
    for key in g.files:
        if key.startswith("uniform") and key.endswith("_labels"):
            exp = g[key]
            n_segments = int(g[key.replace("_labels", "_n_segments")])
            img = np.empty(exp.shape + (3,), np.float32)
            img[...] = (0.5, 0.3, 0.7)
            yield key, img, n_segments, exp


def kat_conn(golden):
    g = golden("slic_kats.npz")
    for key in g.files:
        if key.startswith("conn") and key.endswith("_seg"):
            mn, mx, n = (int(v) for v in g[key.replace("_seg", "_sizes")])
            yield key, g[key], mn, mx, n, g[key.replace("_seg", "_expected")]


def flat_two_colour(h, w, seed):
    """Piecewise-constant gland / stroma image: with a low compactness whole regions collapse into
    single clusters (raw components above max_size) and many clusters end up empty (NaN centres)."""
    _, gland = synth.he_like_image(h, w, seed=seed)
    img = np.where(gland[..., None] == 1, synth.HAEMATOXYLIN, synth.EOSIN)
    return (np.round(img * 255.0) / 255.0).astype(np.float32)


def blob_label_map(h, w, seed, n_blobs, speckle):
    """Adversarial input for the connectivity stage: a few huge regions (far above any max_size),
    thin stripes, and single-pixel speckle (far below any min_size)."""
    rng = np.random.default_rng(seed)
    cy, cx = rng.uniform(0, h, n_blobs), rng.uniform(0, w, n_blobs)
    yy, xx = np.mgrid[:h, :w]
    seg = np.argmin((yy[..., None] - cy) ** 2 + (xx[..., None] - cx) ** 2, axis=-1).astype(np.int32)
    seg[h // 3, :] = n_blobs                       # one-pixel-wide stripe across everything
    seg[:, w // 2] = n_blobs + 1
    noise = rng.random((h, w)) < speckle
    seg[noise] = rng.integers(0, n_blobs + 2, noise.sum())
    return seg


# ---------------------------------------------------------------------------
# CPU: the oracle against the hand-derived vectors and against its twin
# ---------------------------------------------------------------------------
def test_oracles_reproduce_hand_derived_uniform_images(golden):
    for key, img, n_segments, exp in kat_uniform(golden):
        for impl in (oslic.slic, slic_np.slic):
            np.testing.assert_array_equal(impl(img, n_segments, 40), exp, err_msg=f"{key} {impl.__module__}")


def test_oracles_reproduce_hand_derived_connectivity(golden):
    for key, seg, mn, mx, n, exp in kat_conn(golden):
        out, n_out = oslic.connectivity(seg, mn, mx)
        np.testing.assert_array_equal(out, exp, err_msg=key)
        assert n_out == n, key
        out2, n2 = slic_np.enforce_connectivity(seg, mn, mx)
        np.testing.assert_array_equal(out2, exp, err_msg=key)
        assert n2 == n, key


@pytest.mark.parametrize("h,w,seed,compactness,flat", [(64, 80, 5, 40, False), (96, 128, 1, 1.0, True), (131, 97, 2, 0.5, False),
                                                       (232, 200, 3, 0.1, True), (120, 90, 9, 10, False)])
def test_the_two_cpu_restatements_agree(h, w, seed, compactness, flat):
    img = flat_two_colour(h, w, seed) if flat else synth.he_like_image(h, w, seed=seed)[0].astype(np.float32) / 255.0
    n_segments = int(h * w / 200)
    a = oslic.slic(img, n_segments, compactness)
    b = slic_np.slic(img, n_segments, compactness)
    np.testing.assert_array_equal(a, b)


def test_the_two_connectivity_restatements_agree_on_adversarial_maps():
    for seed, (h, w) in enumerate([(40, 56), (33, 71)]):
        seg = blob_label_map(h, w, seed, 5, 0.03)
        for mn, mx in ((3, 20), (10, 200), (1, 1), (50, 50)):
            a, na = oslic.connectivity(seg, mn, mx)
            b, nb = slic_np.enforce_connectivity(seg, mn, mx)
            np.testing.assert_array_equal(a, b)
            assert na == nb


def test_abi_exports_slic_symbols():
    from wesup_b200 import _lib
    lib = _lib.load()
    for name in ("wesup_slic", "wesup_slic_batch", "wesup_enforce_connectivity", "wesup_slic_batch_workspace_bytes",
                 "wesup_enforce_connectivity_workspace_bytes"):
        assert hasattr(lib, name)
    assert lib.wesup_slic_workspace_bytes(464, 464, 1076) == lib.wesup_slic_batch_workspace_bytes(1, 464, 464, 1076)
    assert lib.wesup_slic_batch_workspace_bytes(4, 464, 464, 1076) > 3 * lib.wesup_slic_workspace_bytes(464, 464, 1076)
    assert lib.wesup_slic_workspace_bytes(1, 500, 2) == 0                # strip thinner than one grid step: out of contract


# ---------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------
def gpu_slic(img_hwc, n_segments, compactness, **kw):
    from wesup_b200 import ops
    x = torch.from_numpy(np.ascontiguousarray(img_hwc.transpose(2, 0, 1))).to(DEV)
    labels, n = ops.slic(x, n_segments, compactness, **kw)
    return labels.cpu().numpy(), int(n.item())


@gpu
def test_gpu_reproduces_hand_derived_vectors(golden):
    from wesup_b200 import ops
    for key, img, n_segments, exp in kat_uniform(golden):
        labels, n = gpu_slic(img, n_segments, 40)
        np.testing.assert_array_equal(labels, exp, err_msg=key)
        assert n == exp.max() + 1
    for key, seg, mn, mx, n, exp in kat_conn(golden):
        out, n_out = ops.enforce_connectivity(torch.from_numpy(seg).to(DEV), mn, mx)
        np.testing.assert_array_equal(out.cpu().numpy(), exp, err_msg=key)
        assert int(n_out.item()) == n, key


@gpu
@pytest.mark.parametrize("h,w,seed,n_blobs,speckle", [(40, 56, 0, 5, 0.03), (200, 232, 1, 9, 0.01), (464, 464, 2, 30, 0.002),
                                                      (97, 131, 3, 2, 0.0), (16, 16, 4, 1, 0.0), (522, 775, 5, 40, 0.005)])
def test_connectivity_is_bit_exact_on_adversarial_maps(h, w, seed, n_blobs, speckle):
    """Huge regions (max_size cut, many cuts per region), stripes, speckle (min_size merges, chains of merges):
    integer work, so the result must equal the sequential algorithm's exactly."""
    from wesup_b200 import ops
    seg = blob_label_map(h, w, seed, n_blobs, speckle)
    seg_size = 200.0
    for mn, mx in ((int(0.5 * seg_size), int(3 * seg_size)), (3, 20), (1, 1), (40, 40), (0, 7), (300, 1000), (129, 129)):
        ref, ref_n = oslic.connectivity(seg, mn, mx)
        out, n = ops.enforce_connectivity(torch.from_numpy(seg).to(DEV), mn, mx)
        out = out.cpu().numpy()
        assert int(n.item()) == ref_n, (mn, mx)
        bad = int((out != ref).sum())
        assert bad == 0, f"min_size={mn} max_size={mx}: {bad} pixels differ"


@gpu
def test_connectivity_rejects_min_above_max_and_batches():
    from wesup_b200 import ops
    seg = torch.from_numpy(blob_label_map(64, 64, 7, 4, 0.02)).to(DEV)
    with pytest.raises(RuntimeError):
        ops.enforce_connectivity(seg, 10, 5)
    segs = torch.stack([torch.from_numpy(blob_label_map(64, 80, s, 4, 0.02)) for s in range(5)]).to(DEV)
    out, n = ops.enforce_connectivity(segs, 20, 120)
    for i in range(5):
        one, n1 = ops.enforce_connectivity(segs[i], 20, 120)
        assert torch.equal(out[i], one) and int(n[i]) == int(n1[0])


@gpu
@pytest.mark.parametrize("h,w,seed,compactness", [(96, 128, 1, 1.0), (232, 200, 3, 0.1), (464, 464, 11, 1.0), (150, 310, 5, 0.5)])
def test_slic_with_max_size_cut_and_empty_clusters(h, w, seed, compactness):
    """Flat two-colour images at low compactness: raw components above max_size = int(3*H*W/n_segments) (the cut the
    round-1 kernels lacked) and NaN centres of emptied clusters."""
    img = flat_two_colour(h, w, seed)
    n_segments = int(h * w / 200)
    ref, ref_raw, cent, ref_n = oslic.slic(img, n_segments, compactness, return_aux=True)
    assert np.isnan(cent[:, 0]).any()                                   # the case really has empty clusters
    raw, _ = gpu_slic(img, n_segments, compactness, enforce_connectivity=False)
    assert float((raw == ref_raw).mean()) >= 0.99
    labels, n = gpu_slic(img, n_segments, compactness)
    assert float((labels == ref).mean()) >= 0.99, float((labels == ref).mean())
    assert abs(n - ref_n) <= max(2, 0.01 * ref_n)
    # connectivity of the GPU's own raw assignment, isolated from fp differences: exact
    seg_size = h * w / n_segments
    seq, seq_n = oslic.connectivity(raw, int(0.5 * seg_size), int(3 * seg_size))
    np.testing.assert_array_equal(labels, seq)
    assert n == seq_n
    sizes = np.bincount(labels.ravel())
    assert sizes.max() <= int(3 * seg_size) + int(0.5 * seg_size) * 8        # cut pieces (+ merged crumbs), not whole regions


@gpu
@pytest.mark.parametrize("h,w,seed", [(522, 775, 21), (1516, 1512, 22)])
def test_slic_agreement_at_glas_and_crag_shapes(h, w, seed):
    img_u8, _ = synth.he_like_image(h, w, seed=seed)
    img = img_u8.astype(np.float32) / 255.0
    n_segments = int(h * w / 200)
    ref, ref_raw, _, ref_n = oslic.slic(img, n_segments, 40, return_aux=True)
    raw, _ = gpu_slic(img, n_segments, 40, enforce_connectivity=False)
    assert float((raw == ref_raw).mean()) >= 0.99
    labels, n = gpu_slic(img, n_segments, 40)
    agree = float((labels == ref).mean())
    assert agree >= 0.99, agree
    assert abs(n - ref_n) <= max(2, 0.01 * ref_n)
    assert len(np.unique(labels)) == n and labels.min() == 0 and labels.max() == n - 1


@gpu
def test_slic_batch_equals_single_calls_bit_for_bit_and_is_reproducible():
    from wesup_b200 import ops
    h, w = 200, 232
    imgs = torch.stack([synth.to_tensor(synth.he_like_image(h, w, seed=40 + i)[0]) for i in range(5)]).to(DEV)
    n_segments = int(h * w / 200)
    labels, n = ops.slic_batch(imgs, n_segments, 40)
    again, n2 = ops.slic_batch(imgs, n_segments, 40)
    assert torch.equal(labels, again) and torch.equal(n, n2)                 # integer / fixed-point sums: no run-to-run noise
    for i in range(5):
        one, n1 = ops.slic(imgs[i], n_segments, 40)
        assert torch.equal(labels[i], one), f"image {i}: {(labels[i] != one).sum().item()} pixels differ"
        assert int(n[i]) == int(n1[0])
    raw_b, k_b = ops.slic_batch(imgs, n_segments, 40, enforce_connectivity=False)
    raw_1, k_1 = ops.slic(imgs[3], n_segments, 40, enforce_connectivity=False)
    assert torch.equal(raw_b[3], raw_1) and int(k_b[3]) == int(k_1[0])


@gpu
def test_slic_on_a_non_default_device_stream():
    """every ABI call must go to the stream of the device that owns the tensors."""
    from wesup_b200 import ops
    h, w = 96, 128
    x = synth.to_tensor(synth.he_like_image(h, w, seed=3)[0]).to(DEV)
    ref, n_ref = ops.slic(x, int(h * w / 200), 40)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        out, n = ops.slic(x, int(h * w / 200), 40)
    side.synchronize()
    assert torch.equal(out, ref) and torch.equal(n, n_ref)
    if torch.cuda.device_count() > 1:
        x1 = x.to("cuda:1")
        out1, n1 = ops.slic(x1, int(h * w / 200), 40)               # current device stays cuda:0
        assert out1.device == x1.device and torch.equal(out1.cpu(), ref.cpu())
