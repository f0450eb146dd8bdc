"""CPU-side tests: the C-ABI library loads and exports every symbol the header
declares, argument validation returns error codes without touching a GPU, the
product path refuses to run without CUDA, and the host logic (sharding,
gradient averaging over gloo, tile split/merge, synthetic data) works."""
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from wesup_b200 import _lib
    header = (ROOT / "include" / "wesup_b200.h").read_text()
    declared = set(re.findall(r"\b(wesup_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.wesup_abi_version() == _lib.ABI_VERSION
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name


def test_argument_validation_needs_no_gpu():
    from wesup_b200 import _lib
    lib = _lib.load()
    assert lib.wesup_sp_paint(None, None, 10, 2, 1, None, None) == -1
    assert b"null pointer" in lib.wesup_last_error()
    assert lib.wesup_sp_pool_fwd(None, 0, 1, None, None, 10, 8, 2, None, None) == -1
    assert lib.wesup_slic_workspace_bytes(464, 464, 1076) > 464 * 464 * 24
    assert lib.wesup_slic_workspace_bytes(4, 4000, 10) == 0            # degenerate grid
    assert lib.wesup_sp_stats_workspace_bytes(464, 464, 1076, 2) >= 1076 * 4 * 8
    with pytest.raises(_lib.WesupNativeError):
        _lib.check(-1, "demo")


def test_product_path_has_no_cpu_fallback():
    from wesup_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.SuperpixelMaps.from_labels(torch.zeros(4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.label_propagate(torch.zeros(4, 32), torch.zeros(2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.slic(torch.zeros(3, 32, 32), 5)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.hypercolumn([torch.zeros(1, 4, 8, 8)], (8, 8))
    src = "\n".join(p.read_text() for p in (ROOT / "wesup_b200").rglob("*.py"))
    assert "import oracle" not in src and "from oracle" not in src    # the oracle is never on the product path


def test_surface_and_state_dict_keys():
    from wesup_b200.models import WESUP, WESUPConfig, WESUPPixelInference, initialize_trainer
    from oracle.wesup_ref import RefWESUP
    m = WESUP(pretrained=False)
    assert list(m.state_dict()) == list(RefWESUP().state_dict())
    assert list(WESUPPixelInference(pretrained=False).state_dict()) == list(m.state_dict())
    cfg = WESUPConfig().to_dict()
    assert cfg["sp_area"] == 200 and cfg["sp_compactness"] == 40 and cfg["propagate_threshold"] == 0.8
    assert cfg["propagate_weight"] == 0.5 and cfg["epsilon"] == 1e-7 and cfg["batch_size"] == 1
    trainer = initialize_trainer("wesup", device="cpu", pretrained=False)
    opt, sched = trainer.get_default_optimizer()
    assert sched is None and opt.defaults["lr"] == 5e-5 and opt.defaults["momentum"] == 0.9
    with pytest.raises(ValueError):
        initialize_trainer("mild")


def test_cross_entropy_host_semantics(golden):
    from wesup_b200.models.wesup import _cross_entropy
    g = golden("cross_entropy_cases.npz")
    yh, yt = torch.from_numpy(g["y_hat"]), torch.from_numpy(g["y_true"])
    np.testing.assert_allclose(float(_cross_entropy(yh, yt)), float(g["loss"]), rtol=1e-6)
    assert float(_cross_entropy(yh, torch.zeros_like(yt))) == 0.0


def test_shard_range_partitions():
    from wesup_b200.parallel import shard_range
    for n in (0, 1, 7, 2500):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_tiles_split_and_merge_roundtrip():
    from wesup_b200 import tiles
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, (130, 95, 3), dtype=np.uint8)
    coords = tiles.top_left_coordinates(130, 95, 40)
    assert len(coords) == 4 * 3 and coords[0] == (0, 0) and coords[-1] == (90, 55)
    patches = tiles.divide_image_to_patches(img, 40)
    assert patches.shape == (12, 40, 40, 3) and patches.dtype == np.uint8
    merged = tiles.combine_patches_to_image(patches.astype(np.float64), 130, 95)
    np.testing.assert_allclose(merged, img)                              # overlaps average identical values
    exact = tiles.top_left_coordinates(20000, 20000, 400)
    assert len(exact) == 2500 and exact[1] == (0, 400)


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from wesup_b200.parallel import GradientAllReduce, init_from_env, shard_range, gather_tiles
rank, world, _ = init_from_env("gloo")
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
sync = GradientAllReduce(model)
with torch.no_grad():
    for p in model.parameters():
        p.add_(rank)                      # de-synchronise, then broadcast must repair it
sync.broadcast_parameters()
x = torch.arange(12, dtype=torch.float32).view(2, 6) + rank
model(x).sum().backward()
local = [p.grad.clone() for p in model.parameters()]
sync.average_gradients()
gathered = [None] * world
dist.all_gather_object(gathered, [g.tolist() for g in local])
mean = [sum(torch.tensor(g[i]) for g in gathered) / world for i in range(len(local))]
ok = all(torch.allclose(p.grad, m, atol=1e-6) for p, m in zip(model.parameters(), mean))
# overlapped path: tiny buckets, all-reduces started from the backward hooks, finish() waits for them
torch.manual_seed(1)
model2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.ReLU(), torch.nn.Linear(3, 2))
sync2 = GradientAllReduce(model2, bucket_mb=1e-4)
sync2.broadcast_parameters()
sync2.enable_overlap()
ok = ok and len(sync2.buckets) >= 3 and sync2.buckets[0][1] == len(sync2.params) and sync2.buckets[-1][0] == 0
ok = ok and sorted(i for lo, hi in sync2.buckets for i in range(lo, hi)) == list(range(len(sync2.params)))
for it in range(2):
    sync2.zero_grad()
    model2(x * (it + 1)).sum().backward()
    started = len(sync2._works)
    sync2.finish()
    ref = [torch.zeros_like(p) for p in model2.parameters()]
    for r in range(world):
        m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.ReLU(), torch.nn.Linear(3, 2))
        m.load_state_dict(model2.state_dict())
        xr = (torch.arange(12, dtype=torch.float32).view(2, 6) + r) * (it + 1)
        m(xr).sum().backward()
        for a, p in zip(ref, m.parameters()):
            a += p.grad / world
    ok = ok and started == len(sync2.buckets)
    ok = ok and all(torch.allclose(p.grad, a, atol=1e-5) for p, a in zip(model2.parameters(), ref))
sync2.suspended = True
sync2.zero_grad()
model2(x).sum().backward()
sync2.finish()
ok = ok and len(sync2._works) == 0          # suspended: purely local
sync2.suspended = False
lo, hi = shard_range(7, rank, world)
tiles = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(-1, 2, 2).contiguous()
full = gather_tiles(tiles, 7, rank, world)
if rank == 0:
    ok = ok and full.shape == (7, 2, 2) and full[:, 0, 0].tolist() == list(range(7))
print("RANK", rank, "OK" if ok else "FAIL")
dist.destroy_process_group()
"""


def test_gradient_allreduce_and_tile_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "RANK 0 OK" in r.stdout and "RANK 1 OK" in r.stdout, r.stdout + r.stderr[-2000:]


def test_synthetic_data_contract():
    from wesup_b200 import synth
    img, pixel_mask, point_mask = synth.sample(64, 80, index=0, ratio=1e-3)
    assert img.shape == (1, 3, 64, 80) and img.dtype == torch.float32 and 0 <= float(img.min()) and float(img.max()) <= 1
    assert pixel_mask.shape == (1, 2, 64, 80) and pixel_mask.dtype == torch.int64
    assert torch.all(pixel_mask.sum(1) == 1)
    assert int(point_mask.sum()) == max(2, int(64 * 80 * 1e-3)) and int(point_mask.sum(1).max()) == 1
    again = synth.sample(64, 80, index=0, ratio=1e-3)
    assert all(torch.equal(a, b) for a, b in zip((img, pixel_mask, point_mask), again))
    seg = synth.perturbed_grid_segments(40, 40, 8, seed=0)
    assert seg.min() == 0 and len(np.unique(seg)) == seg.max() + 1


def test_tiles_match_reference_golden(golden):
    from wesup_b200 import tiles
    g = golden("tiles_cases.npz")
    for i in range(3):
        img, p = g[f"img{i}"], int(g[f"patch{i}"])
        h, w, _ = img.shape
        assert np.array_equal(np.array(tiles.top_left_coordinates(h, w, p)), g[f"coords{i}"])
        assert np.array_equal(tiles.divide_image_to_patches(img, p), g[f"patches{i}"])
        merged = tiles.combine_patches_to_image(g[f"preds{i}"], h, w)
        assert np.array_equal(merged, g[f"combined{i}"])          # same arithmetic order => bit-exact


def test_cli_parses_like_fire():
    from wesup_b200 import cli
    args, kw = cli.parse(["data/glas", "--epochs", "3", "--scales=(0.5,1)", "--smoke", "--model", "wesup", "--lr", "5e-5"])
    assert args == ["data/glas"]
    assert kw == {"epochs": 3, "scales": (0.5, 1), "smoke": True, "model": "wesup", "lr": 5e-5}
    assert cli.run(lambda a, b=1, **k: (a, b, k), ["x", "--b", "2", "--c-d", "q"]) == ("x", 2, {"c_d": "q"})


def test_disjoint_tile_merge_equals_running_mean():
    from wesup_b200 import tiles
    rng = np.random.default_rng(0)
    for shape in [(12, 5, 5, 2), (12, 5, 5)]:
        patches = rng.random(shape)
        assert tiles.tiles_are_disjoint(15, 20, 5) and not tiles.tiles_are_disjoint(16, 20, 5)
        np.testing.assert_array_equal(tiles.combine_patches_to_image(patches, 15, 20), tiles.combine_disjoint(patches, 15, 20))


def test_folder_datasets_follow_the_reference_contract(tmp_path):
    from PIL import Image
    from wesup_b200.utils import is_empty_tensor
    from wesup_b200.utils.data import PointDataset, SegmentationDataset
    root = tmp_path / "train"
    for sub in ("images", "masks", "points"):
        (root / sub).mkdir(parents=True)
    rng = np.random.default_rng(1)
    Image.fromarray(rng.integers(0, 255, (20, 30, 3), dtype=np.uint8)).save(root / "images" / "a.png")
    Image.fromarray(((rng.random((20, 30)) > 0.5) * 255).astype(np.uint8)).save(root / "masks" / "a.png")
    (root / "points" / "a.csv").write_text("3,4,1\n10,12,0\n")
    img, mask = SegmentationDataset(root, train=False)[0]
    assert img.shape == (3, 20, 30) and img.dtype == torch.float32 and 0 <= float(img.min()) and float(img.max()) <= 1
    assert mask.shape == (2, 20, 30) and mask.dtype == torch.int64 and torch.all(mask.sum(0) == 1)
    img, pixel_mask, point_mask = PointDataset(root, train=False)[0]
    assert point_mask.shape == (2, 20, 30) and int(point_mask.sum()) == 2
    assert int(point_mask[1, 4, 3]) == 1 and int(point_mask[0, 12, 10]) == 1
    half = SegmentationDataset(root, train=False, rescale_factor=0.5)[0][0]
    assert half.shape == (3, 10, 15)
    (root / "masks" / "a.png").unlink(); (root / "masks").rmdir()
    assert is_empty_tensor(SegmentationDataset(root, train=False)[0][1])


def test_level_sizes_follow_the_backbone_arithmetic():
    """`WESUP._level_sizes` (what the footprint build is planned with, before the backbone runs) against the
    shapes the VGG16 convolutions actually produce (/root/reference/models/wesup.py:205-210 hooks the same layers)."""
    import torch
    from torch import nn
    from wesup_b200.models import WESUP
    model = WESUP(pretrained=False)
    for h, w in ((37, 51), (96, 112), (400, 400)):
        x, seen = torch.zeros(1, 3, h, w), []
        with torch.no_grad():
            for layer in model.backbone:
                x = layer(x)
                if isinstance(layer, nn.Conv2d):
                    seen.append((x.size(2), x.size(3)))
        assert model._level_sizes(h, w) == seen


def test_gradient_bucket_plan_of_the_wesup_model():
    """Three buckets by time of readiness: every parameter exactly once, tail of the parameter list
    first, a small head bucket (first backbone convolutions: the last gradients of backward) of its own, and the flat
    staging views laid out like their parameters (the fused optimizer wants matching strides)."""
    import torch
    from wesup_b200.models.wesup import WESUP
    from wesup_b200.parallel import GradientAllReduce
    model = WESUP(pretrained=False)
    model.backbone.to(memory_format=torch.channels_last)
    sync = GradientAllReduce(model)
    covered = sorted(i for lo, hi in sync.buckets for i in range(lo, hi))
    assert covered == list(range(len(sync.params)))
    assert sync.buckets[0][1] == len(sync.params) and sync.buckets[-1][0] == 0
    mb = [sum(p.numel() for p in sync.params[lo:hi]) * 4 / 2**20 for lo, hi in sync.buckets]
    assert len(mb) == 3 and mb[-1] <= 8.0 and all(m <= 40.0 for m in mb) and abs(sum(mb) - 72.0) < 4.0
    for (lo, hi), (idx, flat, views) in zip(sync.buckets, sync._small):
        assert idx == list(range(lo, hi)) and flat.numel() == sum(sync.params[i].numel() for i in idx)
        for i, v in zip(idx, views):
            assert v.shape == sync.params[i].shape and v.stride() == sync.params[i].stride()
    # the multi-tensor copy into the staging views keeps the values (channels_last weights included)
    for p in sync.params:
        p.grad = torch.randn_like(p)
    before = [p.grad.clone() for p in sync.params]
    idx, flat, views = sync._small[-1]
    torch._foreach_copy_(views, [sync.params[i].grad for i in idx])
    for i, v in zip(idx, views):
        assert torch.equal(v, before[i])
    assert torch.isfinite(sync.probe())


def test_folded_first_layer_equals_the_reference_order_of_operations():
    """Host logic of pixel-wise inference without the hypercolumn (models/wesup.py:246-261, :392-400): side conv ->
    bilinear upsample (align_corners) -> concat -> Linear(2112, 1024) equals, by linearity, the sum over level
    resolutions of the upsampled products with the folded weights plus the folded bias.  Checked on the CPU in fp64
    with F.interpolate standing in for the CUDA kernel `wesup_upsample_sum`."""
    import torch
    import torch.nn.functional as F
    from wesup_b200.models.wesup import WESUPPixelInference
    torch.manual_seed(3)
    model = WESUPPixelInference(pretrained=False).double().eval()
    x = torch.rand(1, 3, 40, 56, dtype=torch.float64)
    with torch.no_grad():
        outs = model._backbone_levels(x)
        size = (x.size(2), x.size(3))
        # the reference's order: 13 side convolutions, upsample each, concatenate, first Linear
        sides = [getattr(model, n)(o) for n, o in zip(model._side_names, outs)]
        hyper = torch.cat([F.interpolate(s, size, mode="bilinear", align_corners=True) for s in sides], dim=1)
        ref = model.fc_layers[0](hyper[0].permute(1, 2, 0).reshape(-1, hyper.size(1)))
        # the folded order: one product per resolution at the level's own size, upsample, sum, folded bias
        groups, bias = model._folded_first_layer(outs, torch.float64)
        assert len(groups) == 5 and [w.size(1) for _, w, _ in groups] == [128, 256, 768, 1536, 1536]
        total = bias.view(1, -1, 1, 1).double()
        for (gh, gw), w_g, levels in groups:
            rows = torch.cat([o.permute(0, 2, 3, 1).reshape(-1, o.size(1)) for o in levels], dim=1)
            z = F.linear(rows, w_g.double()).view(1, gh, gw, -1).permute(0, 3, 1, 2)
            total = total + (z if (gh, gw) == size else F.interpolate(z, size, mode="bilinear", align_corners=True))
        got = total[0].permute(1, 2, 0).reshape(-1, total.size(1))
    assert got.shape == ref.shape == (40 * 56, 1024)
    # the folding itself runs in fp32 (the weights' dtype): agreement to fp32 rounding of the folded weights
    assert float((got - ref).abs().max()) < 1e-6 * max(1.0, float(ref.abs().max()))
