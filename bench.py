#!/usr/bin/env python
"""Benchmarks of the WESUP hot path (BASELINE.json).  One JSON line on rank 0 per invocation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]
                    [--workload train|tiles_sp|tiles_pixel|micro] [--shape 464|glas|crag]

train (default = the headline metric, BASELINE.json `metric`: train img/s at 464x464)
    one "step" = `--images-per-step` full training iterations (GPU SLIC -> superpixel stats -> VGG16 -> superpixel
    means -> MLP -> label propagation -> loss -> backward -> [bucketed gradient all-reduce, overlapped] -> SGD step)
    on synthetic H&E-like images with 1e-4 point labels and random-init weights, batch-1 SGD as in the reference.
    `value` is device-resident throughput; `e2e` goes through the public trainer API with pinned HOST inputs (H2D +
    loss D2H inside the timed region).  `--shape glas|crag` runs BASELINE configs 2 / 4 (522x775, 1516x1512).
tiles_sp / tiles_pixel (BASELINE config 5)
    one step = one pass of infer_tile / pixel_infer_tile over a synthetic `--slide`-pixel whole-slide image in
    `--patch`-pixel tiles, sharded over the ranks; `value` = tiles/s with the slide resident on the device, `e2e` =
    through `infer_tile.predict` / `pixel_infer_tile.predict` with the slide in host memory (tile cut, H2D, merge,
    D2H inside the timed region).
micro (BASELINE config 3)
    the pooling and label-propagation kernels alone at H=W in {464,1024,2048}, N in {500..8000} superpixels.
--impl reference
    the reference's own code (staged copy under baseline/_ref, unmodified, behind import stubs; SLIC = the C
    restatement because scikit-image is not installable) on the host cores, same workload, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SHAPES = {"464": (464, 464), "glas": (522, 775), "crag": (1516, 1512)}
H = W = 464
C_HYPER = 2112
VGG_C = [32, 32, 64, 64, 128, 128, 128, 256, 256, 256, 256, 256, 256]
VGG_SHIFT = [0, 0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4]


def metric_name(workload, h, w):
    if workload == "train":
        return (f"train img/s at {h}x{w} (WESUP weakly-supervised step: GPU SLIC + VGG16 + superpixel stage + loss + "
                "backward + SGD)")
    if workload == "tiles_sp":
        return "tiled inference tiles/s, superpixel-wise (infer_tile: GPU SLIC + VGG16 + superpixel means + MLP + paint per tile)"
    if workload == "tiles_pixel":
        return "tiled inference tiles/s, pixel-wise (pixel_infer_tile: VGG16 + hypercolumn + per-pixel MLP per tile)"
    return "superpixel-stage kernel microbenchmark (pooling + label propagation)"


def level_hw(h, w):
    out = []
    for n in (2, 2, 3, 3, 3):
        out += [(h, w)] * n
        h, w = h // 2, w // 2
    return out


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu, self.t_begin = [], None, gpu_index, None

    def mark_begin(self):
        """The timed region starts now: only samples taken from here on count (the sampler itself is started
        earlier -- nvidia-smi needs up to a second to deliver its first line, longer than a short timed region)."""
        self.t_begin = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if not any(t >= (self.t_begin or 0.0) for t, _ in self.rows):
            time.sleep(0.25)                        # a region shorter than the sampling period: take the sample that follows it
        self.proc.terminate()
        inside = [r for t, r in self.rows if self.t_begin is None or t >= self.t_begin]
        if not inside and self.rows:
            inside = [self.rows[-1][1]]             # nothing landed inside: the sample nearest to the region
        sm, mx, reasons = [], None, set()
        for r in inside:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# the reference itself (cpu_baseline leg, --impl reference, gpu_eager_baseline): the ONLY users of oracle/ here
# ---------------------------------------------------------------------------
def make_reference_trainer(device):
    """The reference's own WESUPTrainer (unmodified code from baseline/_ref behind import stubs) when the staged copy
    is present, else the oracle port.  Returns (step(img, pixel_mask, point_mask) -> loss, kind)."""
    import torch
    from oracle import reference_harness as RH
    torch.manual_seed(0)
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    if RH.reference_root() is not None:
        ref = RH.import_reference()
        trainer = ref.initialize_trainer("wesup", device=device)
        trainer.optimizer = RH.reference_sgd(trainer)
        trainer.model.train()
        trainer.tracker.train()

        def step(img, pixel_mask, point_mask):
            trainer.train_one_iteration("train", img, pixel_mask, point_mask)      # models/base.py:184-211, as it is
            return trainer.tracker.history["loss"][-1]
        return step, trainer, "reference"
    from oracle import slic as oslic
    from oracle import wesup_ref as O
    model = O.seeded_init_(O.RefWESUP(), seed=0).to(device)
    sgd = torch.optim.SGD(model.parameters(), lr=5e-5, momentum=0.9, weight_decay=1e-3)

    def step(img, pixel_mask, point_mask):
        h, w = img.shape[-2:]
        seg = oslic.slic(img[0].permute(1, 2, 0).cpu().numpy(), int(h * w / 200), 40)
        maps, labels, _ = O.preprocess_superpixels(torch.from_numpy(seg).to(device), point_mask[0].to(device))
        sgd.zero_grad()
        model((img.to(device), maps))
        loss = O.compute_loss(model.sp_pred, model.sp_features, labels, propagate_threshold=0.8)
        loss.backward()
        sgd.step()
        return float(loss.detach())
    return step, None, "port"


REF_NOTE = ("the reference's own WESUPTrainer.train_one_iteration (models/base.py:184-211) from the staged, unmodified copy under "
            "baseline/_ref; skimage.segmentation.slic is served by the C restatement oracle/slic_ref.c (scikit-image is not "
            "installable in this image); vgg16 random init; fp32")


def cpu_baseline_train(h, w, reps=1):
    import torch
    from wesup_b200 import synth
    step, _, kind = make_reference_trainer("cpu")
    step(*synth.sample(h, w, index=1))                 # warm-up
    t0 = time.perf_counter()
    for i in range(reps):
        step(*synth.sample(h, w, index=i))
    dt = (time.perf_counter() - t0) / reps
    return {"value": 1.0 / dt, "unit": "img/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{reps} training iteration(s) on one {h}x{w} image after 1 warm-up iteration: " + (REF_NOTE if kind == "reference" else
                      "oracle port (C-SLIC + dense sp_maps + VGG16 + dense mm pooling + loss + backward + SGD, fp32)"),
            "host_cpus": os.cpu_count(), "seconds": dt}


def gpu_eager_baseline(dev, shapes):
    """The reference's own op sequence executed by PyTorch eager ON THE B200 (SURVEY.md
    section 2.1: "the Blackwell kernel to beat"): dense (N,H,W) sp_maps, 13 x interpolate + cat, dense mm pooling,
    (N,N,D) affinity, per-superpixel Python loops, CPU SLIC with its GPU->CPU->GPU hop -- `device='cuda'`."""
    import torch
    from wesup_b200 import synth
    out = {}
    for name in shapes:
        h, w = SHAPES[name]
        try:
            step, trainer, kind = make_reference_trainer(str(dev))
            data = [tuple(t.to(dev) for t in synth.sample(h, w, index=i)) for i in range(2)]
            step(*data[1])
            torch.cuda.synchronize(dev)
            reps = 3
            t0 = time.perf_counter()
            for i in range(reps):
                step(*data[i % 2])
            torch.cuda.synchronize(dev)
            dt = (time.perf_counter() - t0) / reps
            # the same step with the CPU SLIC taken out of the timed region (the reference op sequence alone)
            slic_s = None
            if trainer is not None:
                from oracle import slic as oslic
                img_np = data[0][0][0].permute(1, 2, 0).cpu().numpy()
                t1 = time.perf_counter()
                oslic.slic(img_np, int(h * w / 200), 40)
                slic_s = time.perf_counter() - t1
            out[f"{h}x{w}"] = {"value": 1.0 / dt, "unit": "img/s", "seconds": dt, "kind": kind, "device": torch.cuda.get_device_name(dev),
                               "cpu_slic_seconds_inside": slic_s,
                               "img_per_s_without_cpu_slic": (1.0 / (dt - slic_s)) if slic_s and dt > slic_s else None,
                               "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
            del step, trainer, data
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001  (a baseline leg must never take the bench line down)
            out[f"{h}x{w}"] = {"unavailable": f"{type(ex).__name__}: {str(ex)[:160]}"}
            torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                   # CPU arm: rank 0 alone does the work
    import torch
    from wesup_b200 import synth
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm runs on rank 0 alone and may use all host threads
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    h, w = SHAPES[args.shape]
    if args.workload in ("tiles_sp", "tiles_pixel"):
        return run_reference_tiles(args)
    step, _, kind = make_reference_trainer("cpu")
    data = [synth.sample(h, w, index=i) for i in range(2)]
    for i in range(args.warmup):
        step(*data[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(*data[i % 2])
    dt = time.perf_counter() - t0
    value = args.steps / dt
    cores = torch.get_num_threads()
    sample = f"1 image ({h}x{w}) per step: " + (REF_NOTE if kind == "reference" else "oracle port of the reference step, fp32")
    line = {"impl": "reference", "metric": metric_name("train", h, w), "value": value, "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"WESUP train step {h}x{w}, batch-1 SGD, 1e-4 point labels, random-init VGG16",
                       "images_per_step": 1, "device": "cpu"},
            "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": kind, "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_reference_tiles(args):
    """The reference's infer_tile.predict / pixel-wise loop on the CPU over a bounded sample of the slide."""
    import numpy as np
    import torch
    from oracle import reference_harness as RH
    torch.set_num_threads(os.cpu_count() or 1)
    n_side = 2                                                    # a 2x2 corner of the slide per step: bounded sample
    slide = synthetic_slide(args.patch * n_side)
    ref = RH.import_reference()
    kind = "reference"
    if args.workload == "tiles_sp":
        import infer_tile as ref_tile                             # the staged reference's file (sys.path[0] = baseline/_ref)
        assert "baseline" in os.path.realpath(ref_tile.__file__)
        trainer = ref.initialize_trainer("wesup", device="cpu")
        trainer.model.eval()
        import tempfile
        from PIL import Image
        path = os.path.join(tempfile.mkdtemp(prefix="wesup_bench_"), "corner.png")
        Image.fromarray(slide).save(path)                         # predict() reads an image file

        def one_pass():
            with torch.no_grad():
                return ref_tile.predict(trainer, path, args.patch, device="cpu")
    else:
        from models.wesup import WESUPPixelInference
        import infer_tile as ref_tile
        model = WESUPPixelInference().eval()

        def one_pass():                                           # pixel_infer_tile.py:45-57
            import torchvision.transforms.functional as TF
            from PIL import Image
            preds = []
            with torch.no_grad():
                for patch in ref_tile.divide_image_to_patches(slide, args.patch):
                    preds.append(model(TF.to_tensor(Image.fromarray(patch)).unsqueeze(0))[..., 1].numpy())
            return ref_tile.combine_patches_to_image(np.stack(preds), slide.shape[0], slide.shape[1])
    for _ in range(max(0, min(args.warmup, 1))):
        one_pass()
    steps = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = (time.perf_counter() - t0) / steps
    value = n_side * n_side / dt
    cores = torch.get_num_threads()
    sample = (f"{n_side * n_side} tiles of {args.patch} px per step (a corner of the slide) through the reference's own "
              f"{'infer_tile.predict' if args.workload == 'tiles_sp' else 'pixel_infer_tile loop'} on the CPU; " + REF_NOTE)
    line = {"impl": "reference", "metric": metric_name(args.workload, 0, 0), "value": value, "unit": "tiles/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {args.patch}-px tiles, bounded sample of {n_side * n_side} tiles", "device": "cpu"},
            "cpu_baseline": {"value": value, "unit": "tiles/s", "cores": cores, "kind": kind, "sample": sample, "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# own arm: kernel microbenchmarks
# ---------------------------------------------------------------------------
QUICK = bool(int(os.environ.get("WESUP_BENCH_QUICK", "0")))     # 1 launch per kernel (ncu --set full captures)


class L2Flush:
    """Evicts the 126 MB L2 between timed launches: a 256 MB write, then a 256 MB read of a second
    buffer.  The write alone would leave ~126 MB of DIRTY lines behind, whose write-back would then
    be charged (as extra DRAM traffic, ~20 us) to whatever kernel is timed next; the read pass forces
    that write-back before the timed region and leaves clean lines."""

    def __init__(self, dev, nbytes=256 << 20):
        import torch
        self.w = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.r = torch.zeros(nbytes // 4, dtype=torch.int32, device=dev)

    def __call__(self):
        self.w.zero_()
        self.r.sum()


def time_kernel(fn, iters, flush):
    """Average device time of `fn` in ms over `iters` launches, CUDA events on the
    current stream, an L2 flush (`L2Flush`) before every launch."""
    import torch
    if QUICK:
        iters = 1
    for _ in range(1 if QUICK else 3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        total += s.elapsed_time(e)
    return total / iters


def pooling_kernels(dev, h, w, sp, flush, lib, channels, tag, out, with_build=True):
    """The default-path pooling operators (footprint formulation) on feature levels of the given channel counts."""
    import torch
    from wesup_b200 import ops
    ia = ops._lib.int_array
    sizes = level_hw(h, w)
    g = torch.Generator(device=dev).manual_seed(1)
    lv_mem = [torch.randn(hh, ww, c, device=dev, generator=g) for c, (hh, ww) in zip(channels, sizes)]
    lv_bytes = sum(t.numel() * 4 for t in lv_mem)
    ctot, n_sp, hw = sum(channels), sp.n, h * w
    lca, ha, wa = ia(channels), ia([s[0] for s in sizes]), ia([s[1] for s in sizes])
    lptrs = ops._lib.ptr_array([t.data_ptr() for t in lv_mem])
    st = torch.cuda.current_stream().cuda_stream
    pooled = torch.empty(n_sp, ctot, device=dev)
    gl = [torch.empty_like(t) for t in lv_mem]
    gptrs = ops._lib.ptr_array([t.data_ptr() for t in gl])
    gp = torch.randn(n_sp, ctot, device=dev, generator=g)
    b = lv_bytes + hw * 4 + n_sp * ctot * 4 + n_sp * 4
    ms = time_kernel(lambda: lib.wesup_levels_pool_fwd(lptrs, lca, ha, wa, 13, h, w, sp.seg_offsets.data_ptr(),
                                                       sp.seg_pixels.data_ptr(), n_sp, pooled.data_ptr(), st), 10, flush)
    out[f"levels_pool_fwd_{tag}"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6}
    lws = torch.empty(lib.wesup_levels_pool_bwd_workspace_bytes(lca, ha, wa, 13, h, w), dtype=torch.uint8, device=dev)
    ms = time_kernel(lambda: lib.wesup_levels_pool_bwd(gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(),
                                                       lca, ha, wa, 13, h, w, n_sp, gptrs, lws.data_ptr(), st), 10, flush)
    out[f"levels_pool_bwd_{tag}"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6}
    fpb = torch.empty(lib.wesup_footprint_bytes(ha, wa, 13, h, w, n_sp), dtype=torch.uint8, device=dev)
    build = lambda: lib.wesup_footprint_build(ha, wa, 13, h, w, n_sp, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),  # noqa: E731
                                              sp.row_labels.data_ptr(), sp.counts.data_ptr(), 1, fpb.data_ptr(), st)
    if with_build:
        ms = time_kernel(build, 10, flush)
        out["footprint_build"] = {"ms": ms, "note": "per image, label map only; forked beside the backbone"}
    build()
    ms = time_kernel(lambda: lib.wesup_levels_pool_fwd_fp(lptrs, lca, ha, wa, 13, h, w, sp.seg_offsets.data_ptr(),
                                                          sp.seg_pixels.data_ptr(), n_sp, fpb.data_ptr(), pooled.data_ptr(), st), 10, flush)
    out[f"fp_pool_fwd_{tag}"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6}
    ms = time_kernel(lambda: lib.wesup_levels_pool_bwd_fp(gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(),
                                                          lca, ha, wa, 13, h, w, n_sp, fpb.data_ptr(), gptrs, st), 10, flush)
    out[f"fp_pool_bwd_{tag}"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6}


def label_propagation_kernels(dev, flush, lib, shapes, out):
    import torch
    from wesup_b200 import ops
    st = torch.cuda.current_stream().cuda_stream
    for tag, n, n_l in shapes:
        f = (torch.randn(n, 32, device=dev) * 0.06).abs()
        y_l = torch.zeros(n_l, 2, device=dev); y_l[:, 0] = 1
        n_u = n - n_l
        y_u = torch.zeros(n_u, 2, device=dev)
        src = torch.zeros(n_u, dtype=torch.int32, device=dev)
        sim = torch.zeros(n_u, device=dev)
        lws = torch.empty(lib.wesup_label_propagate_workspace_bytes(n, 32, n_l), dtype=torch.uint8, device=dev)
        for algo in ("exact", "tc"):
            cfn = getattr(lib, ops._LP_ALGOS[algo])
            ms = time_kernel(lambda: cfn(f.data_ptr(), n, 32, n_l, y_l.data_ptr(), 2, 0.8, y_u.data_ptr(), src.data_ptr(),
                                         sim.data_ptr(), lws.data_ptr(), st), 10, flush)
            entry = {"ms": ms, "n": n, "n_l": n_l, "useful_flops": 2.0 * n_u * n_l * 32,
                     "useful_tflops": 2.0 * n_u * n_l * 32 / ms / 1e9}
            if algo == "tc":
                entry["tensor_flops"] = 2.0 * n_u * n_l * 104          # split-TF32: 13 MMAs of K = 8 (hi.hi, hi.lo, lo.hi, norm step)
                entry["tensor_tflops"] = entry["tensor_flops"] / ms / 1e9
                st_ = ops.label_propagate(f, y_l, 0.8, algo="tc", return_stats=True)[-1]
                entry["exact_reevaluations_per_row"] = st_["exact_evals"] / max(n_u, 1)
            out[f"label_propagate_{algo}_{tag}"] = entry


def kernel_rooflines(dev, peak_gbs, h=None, w=None, materialized=True):
    """Microbench of every superpixel-stage kernel at the workload's shape, timed
    alone with CUDA events; algorithmic bytes per SURVEY.md section 8d / DESIGN.md."""
    import torch
    from wesup_b200 import ops, synth
    from wesup_b200.ops import SuperpixelMaps
    h, w = h or H, w or W
    flush = L2Flush(dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    sizes = level_hw(h, w)
    hw = h * w
    out = {}
    img, _, point_mask = synth.sample(h, w, index=0)
    x = img.to(dev)
    labels, n = ops.slic(x, int(hw / 200), 40)
    n_sp = int(n.item())
    sp = SuperpixelMaps.from_labels(labels, point_mask[0].to(dev), n_sp=n_sp)
    lib = ops._lib.load()
    st = torch.cuda.current_stream().cuda_stream
    ia = ops._lib.int_array
    if materialized:
        sides = [torch.randn(1, c, hh, ww, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
                 for c, (hh, ww) in zip(VGG_C, sizes)]
        side_bytes = sum(s.numel() * 4 for s in sides)
        for dtype, es, tag in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
            feats = ops.hypercolumn(sides, (h, w), dtype=dtype)
            ms = time_kernel(lambda: ops.hypercolumn(sides, (h, w), dtype=dtype), 10, flush)
            b = side_bytes + C_HYPER * hw * es
            out[f"hypercolumn_fwd_{tag}"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6}
            ms = time_kernel(lambda: ops.sp_pool(feats, sp), 10, flush)
            b = C_HYPER * hw * es + hw * 4 + n_sp * C_HYPER * 4 + n_sp * 4
            out[f"sp_pool_fwd_{tag}"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6}
            gp = torch.randn(n_sp, C_HYPER, device=dev)
            gf = torch.empty_like(feats)
            code = ops._DTYPES[dtype]
            ms = time_kernel(lambda: lib.wesup_sp_pool_bwd(gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(), hw,
                                                           C_HYPER, n_sp, gf.data_ptr(), code, 1, st), 10, flush)
            b = n_sp * C_HYPER * 4 + hw * 4 + C_HYPER * hw * es
            out[f"sp_pool_bwd_{tag}"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6}
            gsides = [torch.empty((1, s.size(2), s.size(3), s.size(1)), device=dev) for s in sides]
            Cs, hs, ws_ = [s.size(1) for s in sides], [s.size(2) for s in sides], [s.size(3) for s in sides]
            ptrs = ops._lib.ptr_array([t.data_ptr() for t in gsides])
            hws = torch.empty(lib.wesup_hypercolumn_bwd_workspace_bytes(ia(Cs), ia(hs), ia(ws_), 13, h, w), dtype=torch.uint8, device=dev)
            ms = time_kernel(lambda: lib.wesup_hypercolumn_bwd(gf.data_ptr(), code, 1, ia(Cs), ia(hs), ia(ws_), 13, h, w, ptrs,
                                                               hws.data_ptr(), st), 5, flush)
            b = side_bytes + C_HYPER * hw * es
            out[f"hypercolumn_bwd_{tag}"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6}
            del feats, gf, gsides
        del sides
    # the default path: pooling straight from the levels, on the 13 side outputs (2112 ch) and on the 13 backbone outputs (4224)
    pooling_kernels(dev, h, w, sp, flush, lib, VGG_C, "side2112", out, with_build=False)
    pooling_kernels(dev, h, w, sp, flush, lib, [2 * c for c in VGG_C], "backbone4224", out, with_build=True)
    # bias gradients of the 13 backbone convolutions (channels_last conv gradients): wesup_colsum
    grads = [torch.randn(hh * ww, 2 * c, generator=g).to(dev) for c, (hh, ww) in zip(VGG_C, sizes)]
    ms = time_kernel(lambda: [ops.colsum(t) for t in grads], 10, flush)
    b = sum(t.numel() * 4 for t in grads)
    out["conv_bias_grad_colsum_x13"] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6,
                                        "note": "13 eager calls through Python (2 allocations + 2 launches each): host-bound here; inside the "
                                                "training graph the same launches replace ATen reductions"}
    del grads
    # SLIC: one image, and a batch of four in the same launches
    for bsz in (1, 4):
        xs = torch.stack([synth.sample(h, w, index=i)[0][0] for i in range(bsz)]).to(dev).contiguous()
        ws = torch.empty(lib.wesup_slic_batch_workspace_bytes(bsz, h, w, int(hw / 200)), dtype=torch.uint8, device=dev)
        lab_out = torch.empty((bsz, h, w), dtype=torch.int32, device=dev)
        n_out = torch.zeros(bsz, dtype=torch.int32, device=dev)
        ms = time_kernel(lambda: lib.wesup_slic_batch(xs.data_ptr(), 0, bsz, h, w, int(hw / 200), 40.0, 10, 1, lab_out.data_ptr(),
                                                      n_out.data_ptr(), ws.data_ptr(), st), 10, flush)
        b = bsz * hw * 360
        out["slic" if bsz == 1 else f"slic_batch{bsz}"] = {"ms": ms, "ms_per_image": ms / bsz, "bytes": b, "gbs": b / ms / 1e6, "launches": 3}
        del xs, ws
    ms = time_kernel(lambda: SuperpixelMaps.from_labels(labels, point_mask[0].to(dev), n_sp=n_sp), 10, flush)
    out["sp_stats"] = {"ms": ms, "bytes": hw * 5 + n_sp * 12, "gbs": (hw * 5 + n_sp * 12) / ms / 1e6}
    # (c) label propagation: realistic (1e-4 point labels) and the microbench stress shape, both paths
    label_propagation_kernels(dev, flush, lib, (("realistic", n_sp, max(sp.n_labeled, 1)), ("stress", 8000, 4000)), out)
    for v in out.values():
        if "gbs" in v:
            v["frac_of_hbm_peak"] = v["gbs"] / peak_gbs
    return out


# ---------------------------------------------------------------------------
# own arm: training step
# ---------------------------------------------------------------------------
def end_process(world):
    """Leave without tearing NCCL down collectively: ranks finish at different times (rank 0 goes on to the kernel
    microbenchmarks and the baselines) and CUDA graphs that captured NCCL kernels outlive the trainer through reference
    cycles; a communicator destroyed under them can block interpreter shutdown.  Everything is flushed; the process ends."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def init_dist():
    import torch
    from wesup_b200 import parallel
    rank, world, local = parallel.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path for the own arm)"
    torch.cuda.set_device(local)
    return rank, world, local, torch.device("cuda", local)


def timed_region(dev, world, step_fn, steps):
    """EXACTLY `steps` calls of step_fn between barrier + synchronize on both sides; CUDA events; max over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        step_fn(i)
    e.record()
    barrier()
    ms = torch.tensor([s.elapsed_time(e)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def run_train(args):
    import torch
    import torch.distributed as dist
    from wesup_b200 import _lib, synth
    from wesup_b200.models import initialize_trainer
    from wesup_b200.utils.metrics import accuracy, dice

    h, w = SHAPES[args.shape]
    rank, world, local, dev = init_dist()
    torch.manual_seed(0)
    lib = _lib.load()
    trainer = initialize_trainer("wesup", device=dev, pretrained=False, materialize_hypercolumn=args.materialize,
                                 pool_first=not args.no_pool_first, cuda_graph=not args.no_graph,
                                 footprints=not args.no_footprints, cudnn_benchmark=not args.no_cudnn_benchmark)
    trainer.optimizer, _ = trainer.get_default_optimizer()
    trainer.metric_funcs = [accuracy, dice]
    if (world > 1 or args.dp_at_1) and not args.no_dp:
        trainer.enable_data_parallel(overlap=not args.no_overlap, bucket_mb=args.bucket_mb)
    ips = args.images_per_step
    pool = max(ips, 4)
    host = [synth.sample(h, w, index=rank * pool + i) for i in range(pool)]
    host = [tuple(t.pin_memory() for t in s) for s in host]
    resident = [tuple(t.to(dev) for t in s) for s in host]
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    last_step = [-1]

    def run_step(samples, i):
        # the trainer's epoch loop does the same: preprocessing (H2D copy, GPU SLIC, superpixel
        # statistics) of image k+1 is enqueued on a side stream before image k trains
        for j in range(ips):
            k = i * ips + j
            if not args.no_prefetch:
                trainer.prefetch(*samples[(k + 1) % pool])
            trainer.train_one_iteration("train", *samples[k % pool])
        if i == last_step[0]:
            trainer.flush_metrics()          # the last iteration's loss is read inside the timed region too

    def timed(samples, steps):
        last_step[0] = steps - 1
        l0 = lib.wesup_kernel_launches() + getattr(trainer, "replayed_launches", 0)
        ms = timed_region(dev, world, lambda i: run_step(samples, i), steps)
        # own kernels launched in the timed region: calls through the C ABI (preprocessing) + the ones each graph replay re-issues
        return ms, lib.wesup_kernel_launches() + getattr(trainer, "replayed_launches", 0) - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        run_step(resident, i)
    sampler.mark_begin()
    ms_total, launches = timed(resident, args.steps)
    for i in range(max(1, args.warmup // 2)):
        run_step(host, i)
    ms_e2e, _ = timed(host, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if args.trace:
        # diagnosis only (after the timed regions): kernel timeline of two steps from the CUPTI activity records
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(2):
                run_step(resident, i)
            trainer.flush_metrics()
            torch.cuda.synchronize(dev)
        rows = sorted(((e.time_range.start, e.time_range.end - e.time_range.start, e.name[:90]) for e in prof.events()
                       if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda r_: r_[0])
        t0 = rows[0][0] if rows else 0
        Path(args.trace).mkdir(parents=True, exist_ok=True)
        with open(Path(args.trace) / f"trace_rank{rank}.tsv", "w") as fh:
            for st_, du, nm in rows:
                fh.write(f"{st_ - t0:.1f}\t{du:.1f}\t{nm}\n")
    value = world * ips * args.steps / (ms_total / 1e3)
    e2e_value = world * ips * args.steps / (ms_e2e / 1e3)
    # peak over the whole run: includes the multi-GB workspaces cuDNN's autotuner (cudnn.benchmark) tries during the
    # warm-up steps; what the caching allocator holds when the timed regions end is reported next to it
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2**30
    reserved_end = torch.cuda.memory_reserved(dev) / 2**30
    if rank != 0:
        return end_process(world)
    del trainer, resident
    torch.cuda.empty_cache()
    peak, peak_src = peaks()
    kernels = kernel_rooflines(dev, peak, h, w, materialized=(args.shape == "464")) if not args.skip_kernels else {}
    roofline = None
    if kernels:
        # the dominant pooling launch of the step that was timed (BASELINE metric: "sp-pool+propagation GB/s vs HBM peak")
        per_image_ms = ms_total / args.steps / ips
        if args.materialize:
            key, name, label = "hypercolumn_fwd_f32", "hyper_fwd_bulk_kernel", "hyper_fwd_bulk_kernel<float,4> (wesup_hypercolumn_fwd)"
        else:
            tag = "side2112" if args.no_pool_first else "backbone4224"
            pre = "levels_pool" if args.no_footprints else "fp_pool"
            key = max((f"{pre}_fwd_{tag}", f"{pre}_bwd_{tag}"), key=lambda k_: kernels[k_]["ms"])
            if args.no_footprints:
                name, label = ("levels_pool_fwd_kernel", "levels_pool_fwd_kernel (wesup_levels_pool_fwd)") if "fwd" in key else \
                    ("levels_pool_bwd_kernel", "levels_pool_bwd_kernel<V> x 5 resolution groups (wesup_levels_pool_bwd)")
            else:
                name, label = ("fp_pool_fwd_cells_kernel", "fp_pool_fwd_cells_kernel (wesup_levels_pool_fwd_fp)") if "fwd" in key else \
                    ("fp_pool_bwd", "fp_pool_bwd kernels (wesup_levels_pool_bwd_fp)")
        k = kernels[key]
        traffic, traffic_src = None, None
        tfile = ROOT / "profiles" / "roofline_traffic.json"
        if tfile.exists() and args.shape == "464":
            tj = json.loads(tfile.read_text())
            hits = [v["traffic_bytes"] for n_, v in tj.items() if isinstance(v, dict) and name in n_]
            traffic = sum(hits) if hits else None
            traffic_src = f"ncu --set full capture of this kernel at this shape, {tj.get('_source', 'profiles/')} (not measured in this run)"
        roofline = {"kernel": label, "bound": "hbm", "achieved": k["gbs"],
                    "peak": peak, "peak_source": peak_src + ": a read+write copy; write-only streams measure 7.4 TB/s on this pool",
                    "unit": "GB/s", "frac": k["gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": k["bytes"], "ms_per_launch": k["ms"],
                    "launches_per_image": 1, "share_of_step": k["ms"] / per_image_ms,
                    "largest_own_kernel_of_the_step": {"kernel": "slic_kmeans_kernel + slic_connect_kernel (wesup_slic, one image ahead on the side stream)",
                                                       "ms_per_image": kernels["slic"]["ms"], "gbs": kernels["slic"]["gbs"],
                                                       "frac": kernels["slic"]["gbs"] / peak, "bound": "latency / issue (fp64 distance arithmetic), not HBM",
                                                       "share_of_step": kernels["slic"]["ms"] / per_image_ms}}
    cpu = cpu_baseline_train(h, w) if not args.skip_cpu else None
    eager = gpu_eager_baseline(dev, [args.shape] if args.shape != "464" else ["464", "glas"]) if not args.skip_eager else None
    if eager:
        own = {f"{h}x{w}": value / world}
        for k_, v in eager.items():
            if "value" in v and k_ in own:
                v["speedup_of_this_repo_on_the_same_gpu"] = own[k_] / v["value"]
    line = {"metric": metric_name("train", h, w), "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"WESUP train step {h}x{w}, batch-1 SGD, 1e-4 point labels, random-init VGG16",
                       "images_per_step": ips, "images_per_step_all_ranks": ips * world, "hypercolumn": "f32 pixel-major (H*W,2112)" if args.materialize else
                       ("not materialised: superpixel means from the 13 side outputs" if args.no_pool_first else
                        "not materialised: superpixel means from the 13 backbone levels (4224 ch), side convs on the N pooled rows"),
                       "footprints": "rebuilt inside the pooling kernels" if args.no_footprints else
                       "precomputed per image (wesup_footprint_build, forked beside the backbone)",
                       "iteration": "eager" if args.no_graph else "one CUDA graph per image shape (VGG16 .. gradient all-reduce .. SGD step) replayed; GPU SLIC + superpixel statistics run one image ahead on a side stream",
                       "parallelism": ("replicas (diagnosis)" if args.no_dp else f"dp{world}") + ("" if world == 1 or args.no_dp else (", blocking all-reduce" if args.no_overlap else ", bucketed all-reduce overlapped with backward, captured in the graph")),
                       "l2": "activations of one image (0.25 GB of backbone levels + cuDNN workspaces) >> 126 MB L2; "
                       "kernel microbenches flush L2 (256 MB write, then 256 MB read so no dirty lines remain) before every launch",
                       "cudnn_tf32": bool(torch.backends.cudnn.allow_tf32), "cudnn_benchmark": bool(torch.backends.cudnn.benchmark)},
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d * ips, "d2h_bytes_per_step": 4 * ips,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "gpu_eager_baseline": eager, "peak_mem_gb": peak_mem, "mem_reserved_gb_at_end": reserved_end,
            "kernels": kernels}
    print(json.dumps(line), flush=True)
    end_process(world)


# ---------------------------------------------------------------------------
# own arm: tiled inference over a synthetic whole-slide image
# ---------------------------------------------------------------------------
def synthetic_slide(size, base=2000):
    """H&E-like slide built by tiling a `base`-pixel synthetic image (mirrored so seams are continuous)."""
    import numpy as np
    from wesup_b200 import synth
    base = min(base, size)
    img, _ = synth.he_like_image(base, base, seed=77)
    reps = -(-size // base)
    row = np.concatenate([img if i % 2 == 0 else img[:, ::-1] for i in range(reps)], axis=1)
    full = np.concatenate([row if i % 2 == 0 else row[::-1] for i in range(reps)], axis=0)
    return np.ascontiguousarray(full[:size, :size])


def run_tiles(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import infer_tile
    import pixel_infer_tile
    from wesup_b200 import tiles
    from wesup_b200.models import initialize_trainer
    from wesup_b200.models.wesup import WESUPPixelInference

    rank, world, local, dev = init_dist()
    torch.manual_seed(0)
    mode = "sp" if args.workload == "tiles_sp" else "pixel"
    hc_dtype = torch.bfloat16 if args.hc_dtype == "bf16" else torch.float32
    slide = synthetic_slide(args.slide)
    n_tiles = len(tiles.top_left_coordinates(args.slide, args.slide, args.patch))
    if mode == "sp":
        trainer = initialize_trainer("wesup", device=dev, pretrained=False, materialize_hypercolumn=False, cuda_graph=not args.no_graph,
                                     tile_batch=args.tile_batch)
        trainer.model.eval()

        def api(img):          # the public call: infer_tile.predict(trainer, image, patch, ...)
            return infer_tile.predict(trainer, img, args.patch, device=dev, rank=rank, world_size=world)
        engine_of = lambda: trainer._tile_engine           # noqa: E731
    else:
        model = WESUPPixelInference(pretrained=False, hc_dtype=hc_dtype).to(dev).eval()
        engine = tiles.PixelTileEngine(model, batch=args.tile_batch, use_graph=not args.no_graph)

        def api(img):          # the public call: pixel_infer_tile.predict(model, image, patch, ...)
            return pixel_infer_tile.predict(model, img, args.patch, dev, rank, world, engine=engine)
        engine_of = lambda: engine                          # noqa: E731
    # warm-up on a corner of the slide (cuDNN autotune, graph capture, allocator), then W full passes
    corner = args.patch * max(4, int(np.ceil(np.sqrt(4 * args.tile_batch))))
    warm = slide[:corner, :corner]
    if mode == "sp":
        _sp_engine(trainer, args)
    for _ in range(3):
        tiles.predict_tiles(engine_of(), warm, args.patch, dev, 0, 1)
    slide_dev = torch.from_numpy(slide).to(dev)
    resident = lambda i: tiles.predict_tiles(engine_of(), slide_dev, args.patch, dev, rank, world, return_device=True)   # noqa: E731
    for _ in range(args.warmup):
        resident(0)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = engine_of().launches
    ms_total = timed_region(dev, world, resident, args.steps)
    launches = engine_of().launches - l0
    merged = [None]

    def host_pass(i):
        merged[0] = api(slide)
    host_pass(0)
    t0 = time.perf_counter()
    ms_e2e = timed_region(dev, world, host_pass, args.steps)
    wall_e2e = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = n_tiles * args.steps / (ms_total / 1e3)
    e2e_value = n_tiles * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        return end_process(world)
    assert merged[0].shape[:2] == (args.slide, args.slide)
    out_bytes = merged[0].size * merged[0].itemsize
    line = {"metric": metric_name(args.workload, 0, 0), "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if (mode == "sp" or args.hc_dtype == "fp32") else "bf16 (hypercolumn + MLP GEMMs), f32 backbone",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {args.slide}x{args.slide} synthetic slide, {n_tiles} tiles of {args.patch} px, "
                                   f"batches of {args.tile_batch} tiles, contiguous stripes of the tile list per rank",
                       "slide": args.slide, "patch": args.patch, "tiles": n_tiles, "tile_batch": args.tile_batch,
                       "parallelism": f"tile-parallel x{world}, gather of finished tiles to rank 0", "graph": not args.no_graph,
                       "l2": "a batch of tiles (VGG16 activations of 16 tiles: > 1 GB) >> 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": int(slide.size), "d2h_bytes_per_step": int(out_bytes),
                    "ms_per_step": ms_e2e / args.steps, "wall_s_per_slide": wall_e2e,
                    "api": "infer_tile.predict" if mode == "sp" else "pixel_infer_tile.predict"},
            "seconds_per_slide_device_resident": ms_total / args.steps / 1e3,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": None, "cpu_baseline": None,
            "positive_fraction": float(np.mean(merged[0] > 0.5))}
    print(json.dumps(line), flush=True)
    end_process(world)


def _sp_engine(trainer, args):
    from wesup_b200 import tiles
    if getattr(trainer, "_tile_engine", None) is None:
        trainer._tile_engine = tiles.SuperpixelTileEngine(trainer, batch=args.tile_batch, use_graph=not args.no_graph)
    return trainer._tile_engine


# ---------------------------------------------------------------------------
# own arm: BASELINE config 3 sweep
# ---------------------------------------------------------------------------
def run_micro(args):
    import torch
    from wesup_b200 import ops, synth
    from wesup_b200.ops import SuperpixelMaps
    rank, world, local, dev = init_dist()
    if rank != 0:
        return
    peak, peak_src = peaks()
    flush = L2Flush(dev)
    lib = ops._lib.load()
    table = {}
    for size in (464, 1024, 2048):
        img, _, point_mask = synth.sample(size, size, index=0)
        x = img.to(dev)
        for n_target in (500, 1000, 2000, 4000, 8000):
            labels, n = ops.slic(x, n_target, 40)
            n_sp = int(n.item())
            sp = SuperpixelMaps.from_labels(labels, point_mask[0].to(dev), n_sp=n_sp)
            out = {}
            pooling_kernels(dev, size, size, sp, flush, lib, [2 * c for c in VGG_C], "backbone4224", out, with_build=True)
            label_propagation_kernels(dev, flush, lib, (("realistic", n_sp, max(2, int(size * size * 1e-4))), ("stress", n_sp, n_sp // 2)), out)
            for v in out.values():
                if "gbs" in v:
                    v["frac_of_hbm_peak"] = v["gbs"] / peak
            table[f"{size}x{size}_N{n_target}"] = {"n_superpixels": n_sp, **{k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                                                                               for k, v in out.items() if not k.startswith("levels_pool")}}
            del sp, labels
            torch.cuda.empty_cache()
    best = max(v["fp_pool_fwd_backbone4224"]["gbs"] for v in table.values())
    print(json.dumps({"metric": metric_name("micro", 0, 0), "value": best, "unit": "GB/s", "n_gpus": 1, "steps": 10, "warmup": 3,
                      "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "BASELINE config 3: H=W in {464,1024,2048}, N in {500..8000} (SLIC n_segments), 4224-channel backbone levels; "
                                             "value = best forward pooling GB/s of the sweep", "l2": "flushed before every launch"},
                      "roofline": {"bound": "hbm", "achieved": best, "peak": peak, "unit": "GB/s", "frac": best / peak, "traffic": None, "peak_source": peak_src},
                      "table": table}), flush=True)


def main():
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ.pop("NCCL_DEBUG", None)       # VERSION / WARN print a banner to stdout, which carries the one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "tiles_sp", "tiles_pixel", "micro"])
    ap.add_argument("--shape", default="464", choices=list(SHAPES), help="train: image shape (464 = headline; glas 522x775; crag 1516x1512)")
    ap.add_argument("--images-per-step", type=int, default=4)
    ap.add_argument("--slide", type=int, default=20000, help="tiles_*: edge of the synthetic whole-slide image")
    ap.add_argument("--patch", type=int, default=400, help="tiles_*: tile edge")
    ap.add_argument("--tile-batch", type=int, default=None, help="tiles_*: tiles per batch (default 16 superpixel-wise, 8 pixel-wise)")
    ap.add_argument("--hc-dtype", choices=["fp32", "bf16"], default="bf16", help="tiles_pixel: hypercolumn / MLP GEMM precision")
    ap.add_argument("--materialize", action="store_true",
                    help="classic path: kernel (a) writes the (H*W,2112) hypercolumn, kernel (b) pools it (default: fused, "
                         "superpixel means straight from the backbone levels, side convs on the pooled rows)")
    ap.add_argument("--no-pool-first", action="store_true",
                    help="fused path over the 13 side outputs (side convs on H*W pixels) instead of pool-first")
    ap.add_argument("--no-footprints", action="store_true",
                    help="pooling kernels rebuild the superpixel footprints internally (default: built once per image on a side stream)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true",
                    help="leave torch.backends.cudnn.benchmark off (default here: on -- cuDNN times its algorithms during the eager "
                         "iterations that precede the graph capture)")
    ap.add_argument("--no-graph", action="store_true",
                    help="eager iterations (default: one CUDA graph per image shape, captured after two eager iterations)")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: blocking gradient all-reduce after backward instead of overlapped buckets")
    ap.add_argument("--no-dp", action="store_true", help="N>1 diagnosis: independent replicas, no gradient exchange (NOT data-parallel training)")
    ap.add_argument("--dp-at-1", action="store_true", help="N=1 diagnosis: gradients as views of the flat all-reduce buffer, no collective")
    ap.add_argument("--bucket-mb", type=float, default=40.0, help="N>1: gradient bucket size")
    ap.add_argument("--no-prefetch", action="store_true", help="preprocess inline instead of one image ahead on a side stream")
    ap.add_argument("--trace", default=None, help="train: directory for a kernel timeline (start us, duration us, name) of two extra steps")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-eager", action="store_true", help="omit the gpu_eager_baseline leg")
    ap.add_argument("--skip-kernels", action="store_true", help="omit the per-kernel roofline microbench")
    args = ap.parse_args()
    tiles_wl = args.workload.startswith("tiles")
    if args.steps is None:
        args.steps = 2 if tiles_wl else 10
    if args.warmup is None:
        args.warmup = 1 if tiles_wl else 3
    if args.tile_batch is None:
        args.tile_batch = 16 if args.workload == "tiles_sp" else 8
    global H, W
    H, W = SHAPES[args.shape]
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    {"train": run_train, "tiles_sp": run_tiles, "tiles_pixel": run_tiles, "micro": run_micro}[args.workload](args)


if __name__ == "__main__":
    main()
