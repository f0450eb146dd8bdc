"""Tiled inference (SURVEY.md section 8 rows e-tiles, f3): the batched device pipeline of wesup_b200.tiles against
the reference's per-tile loop (/root/reference/infer_tile.py:105-116, pixel_infer_tile.py:45-57) executed tile by
tile through the same trainer / model, and the device-side merge against the golden merge vectors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from wesup_b200 import synth, tiles                      # noqa: E402
from wesup_b200.models import initialize_trainer         # noqa: E402
from wesup_b200.models.wesup import WESUPPixelInference  # noqa: E402

DEV = "cuda"


def per_tile_reference(trainer, img, patch):
    """The reference's loop body, one tile at a time: postprocess(model(preprocess(tile)))."""
    outs = []
    with torch.no_grad():
        for p in tiles.divide_image_to_patches(img, patch):
            x = synth.to_tensor(p).unsqueeze(0).to(DEV)
            input_, _ = trainer.preprocess(x)
            outs.append(trainer.postprocess(trainer.model(input_))[0].to(torch.uint8).cpu().numpy())
    return np.stack(outs)


@pytest.mark.parametrize("size,patch,batch,graph", [((288, 384), 96, 4, True), ((200, 250), 96, 4, False), ((192, 192), 96, 16, True)])
def test_superpixel_tile_engine_equals_the_per_tile_loop(size, patch, batch, graph):
    torch.manual_seed(0)
    trainer = initialize_trainer("wesup", device=DEV, pretrained=False, materialize_hypercolumn=False)
    trainer.model.eval()
    img, _ = synth.he_like_image(*size, seed=5)
    ref_stack = per_tile_reference(trainer, img, patch)
    ref = tiles.combine_patches_to_image(ref_stack.astype(np.float64), *size)
    engine = tiles.SuperpixelTileEngine(trainer, batch=batch, use_graph=graph)
    for rep in range(3):                                    # eager warm-up, capture, replay
        out = tiles.predict_tiles(engine, img, patch, DEV)
        assert out.shape == size
        # VGG16 runs at batch size `batch` instead of 1 (cuDNN may pick another algorithm): class maps agree except for
        # superpixels whose probability sits within rounding of 0.5
        assert float((out == ref).mean()) >= 0.995, (rep, float((out == ref).mean()))
    on_dev = tiles.predict_tiles(engine, torch.from_numpy(img).to(DEV), patch, DEV)
    np.testing.assert_array_equal(on_dev, out)              # device-resident slide == host slide
    assert engine.launches > 0


def test_superpixel_tile_engine_is_exact_when_the_network_is_deterministic_in_batch():
    """With all weights zero except the classifier bias every superpixel gets the same probability: the engine's
    plumbing (SLIC batch == single, statistics into static buffers, capacity rows, paint, merge) must then reproduce the
    per-tile loop exactly, and the label maps the two paths paint from must be identical."""
    from wesup_b200 import ops
    torch.manual_seed(1)
    trainer = initialize_trainer("wesup", device=DEV, pretrained=False, materialize_hypercolumn=False)
    trainer.model.eval()
    img, _ = synth.he_like_image(192, 288, seed=8)
    patches = tiles.divide_image_to_patches(img, 96)
    x = torch.from_numpy(patches).to(DEV).permute(0, 3, 1, 2).float().div(255.0).contiguous()
    labels_b, n_b = ops.slic_batch(x, int(96 * 96 / 200), 40)
    for t in range(x.size(0)):
        one, n1 = ops.slic(x[t], int(96 * 96 / 200), 40)
        assert torch.equal(labels_b[t], one) and int(n_b[t]) == int(n1[0])
    buf = ops.StaticSuperpixelBuffers(96, 96, 96 * 96 // 23 + 1, DEV)
    ops.sp_stats_into(labels_b[0], buf)
    n = int(n_b[0])
    sp = ops.SuperpixelMaps.from_labels(labels_b[0], None, n_sp=n)
    cap = -(-n // 64) * 64
    view = buf.view(cap)
    assert torch.equal(view.order[:n], sp.order) and torch.equal(view.counts[:n], sp.counts)
    assert torch.equal(view.row_labels, sp.row_labels) and torch.equal(view.seg_pixels, sp.seg_pixels)
    assert torch.equal(view.seg_offsets[:n + 1], sp.seg_offsets) and bool((view.seg_offsets[n:] == 96 * 96).all())
    assert bool((view.counts[n:] == 0).all())


def test_device_merge_equals_the_golden_host_merge(golden):
    g = golden("tiles_cases.npz")
    for i in range(3):
        img, p = g[f"img{i}"], int(g[f"patch{i}"])
        h, w, _ = img.shape
        merged = tiles.combine_on_device(torch.from_numpy(g[f"preds{i}"]).to(DEV), h, w).cpu().numpy()
        np.testing.assert_array_equal(np.squeeze(merged), g[f"combined{i}"])        # same float64 operation order: bit-exact
    rng = np.random.default_rng(0)
    stack = rng.integers(0, 2, (12, 5, 5)).astype(np.uint8)
    np.testing.assert_array_equal(tiles.combine_on_device(torch.from_numpy(stack).to(DEV), 15, 20).cpu().numpy(),
                                  tiles.combine_patches_to_image(stack.astype(np.float64), 15, 20).astype(np.uint8))


@pytest.mark.parametrize("hc_dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_pixel_tile_engine_equals_the_per_tile_loop(hc_dtype, tol):
    torch.manual_seed(0)
    model = WESUPPixelInference(pretrained=False, hc_dtype=hc_dtype).to(DEV).eval()
    img, _ = synth.he_like_image(160, 240, seed=6)
    outs = []
    with torch.no_grad():
        for p in tiles.divide_image_to_patches(img, 80):
            outs.append(model(synth.to_tensor(p).unsqueeze(0).to(DEV))[..., 1].cpu().numpy())
    ref = tiles.combine_patches_to_image(np.stack(outs).astype(np.float64), 160, 240)
    engine = tiles.PixelTileEngine(model, batch=4)
    for _ in range(3):
        out = tiles.predict_tiles(engine, img, 80, DEV)
        assert out.shape == (160, 240)
        assert float(np.abs(out - ref).max()) <= tol, float(np.abs(out - ref).max())


def test_plain_step_functions_still_work():
    img, _ = synth.he_like_image(64, 96, seed=2)
    out = tiles.predict_tiles(lambda x: x[0, 0] * 2.0, img, 32, DEV)
    np.testing.assert_allclose(out, img[..., 0].astype(np.float32) / 255.0 * 2.0, rtol=1e-6)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("h,w", [(48, 40), (37, 51), (400, 400)])
def test_upsample_sum_kernel_vs_interpolate(dtype, tol, h, w):
    """out = relu(bias + sum_g bilinear(align_corners=True)(z_g)) against F.interpolate on fp32 copies of the terms."""
    import torch.nn.functional as F
    from wesup_b200 import ops
    g = torch.Generator().manual_seed(h * w)
    c = 64 if h < 400 else 1024
    sizes = [(h, w), (h // 2, w // 2), (h // 4, w // 4), (h // 8, w // 8), (max(h // 16, 1), max(w // 16, 1))]
    terms = [torch.randn(hh, ww, c, generator=g).to(DEV).to(dtype) for hh, ww in sizes]
    bias = torch.randn(c, generator=g).to(DEV)
    for n_terms in (5, 1, 3):
        use = terms[:n_terms] if n_terms != 3 else terms[1:4]          # also: no full-resolution term
        ref = bias.view(1, -1, 1, 1)
        for t in use:
            ref = ref + F.interpolate(t.float().permute(2, 0, 1).unsqueeze(0), (h, w), mode="bilinear", align_corners=True)
        ref = torch.relu(ref)[0].permute(1, 2, 0).reshape(h * w, c)
        out = ops.upsample_sum(use, (h, w), bias=bias, relu=True)
        assert out.dtype == dtype and out.shape == (h * w, c)
        err = float((out.float() - ref).abs().max()) / float(ref.abs().max())
        assert err <= tol, (n_terms, err)
    lin = ops.upsample_sum(terms[:2], (h, w))                           # no bias, no activation
    ref = terms[0].float().reshape(h * w, c) + F.interpolate(terms[1].float().permute(2, 0, 1).unsqueeze(0), (h, w), mode="bilinear",
                                                             align_corners=True)[0].permute(1, 2, 0).reshape(h * w, c)
    assert float((lin.float() - ref).abs().max()) / float(ref.abs().max()) <= tol


@pytest.mark.parametrize("hc_dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_pixel_model_without_hypercolumn_equals_the_reference_order_of_operations(hc_dtype, tol):
    """project_first (GEMMs at the levels' resolution + upsample_sum) against the reference's order (hypercolumn, then
    the 2112-wide Linear) through the same weights, and against the oracle's dense formulation in fp32."""
    from oracle import wesup_ref as O
    torch.backends.cudnn.allow_tf32 = False            # the side convolutions of the reference order would otherwise run in TF32
    torch.manual_seed(3)
    model = WESUPPixelInference(pretrained=False, hc_dtype=hc_dtype).to(DEV).eval()
    O.seeded_init_(model, seed=9)
    classic = WESUPPixelInference(pretrained=False, hc_dtype=hc_dtype, project_first=False).to(DEV).eval()
    classic.load_state_dict(model.state_dict())
    x = synth.to_tensor(synth.he_like_image(75, 94, seed=4)[0]).unsqueeze(0).to(DEV)       # floor-division level sizes
    with torch.no_grad():
        a, b = model(x), classic(x)
        ref = O.seeded_init_(O.RefWESUP(), seed=9).forward_pixels(x.cpu())
    assert a.shape == (75, 94, 2) and model.feature_maps is None
    assert float((a - b).abs().max()) <= tol
    assert float((a.cpu() - ref).abs().max()) <= tol
    xb = torch.cat([x, x.flip(-1)])
    with torch.no_grad():
        ab = model.forward_batch(xb)
    assert float((ab[0] - a).abs().max()) <= tol and float((ab[1] - model(x.flip(-1))).abs().max()) <= tol
    torch.backends.cudnn.allow_tf32 = True
