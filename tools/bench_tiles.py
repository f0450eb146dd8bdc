#!/usr/bin/env python
"""Tiled inference throughput on a synthetic whole-slide image (BASELINE.json config 5:
400-px patches of a 20k x 20k slide partitioned across the GPUs of one node).

    python tools/bench_tiles.py [--size 20000] [--patch 400] [--mode sp|pixel] [--hc-dtype fp32|bf16]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_tiles.py --size 20000

Each rank owns a contiguous stripe of the row-major tile list (wesup_b200.parallel.shard_range),
runs uint8 tile -> device -> fp32 -> [GPU SLIC + stats ->] VGG16 -> hypercolumn -> [pooling -> MLP ->
paint | per-pixel MLP] -> uint8/fp32 prediction, and rank 0 gathers and merges the finished tiles.
Prints one JSON line on rank 0: tiles/s over all ranks (max-over-ranks device time) and wall time
including the gather + merge.
"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from wesup_b200 import parallel, synth, tiles  # noqa: E402
from wesup_b200.models import initialize_trainer  # noqa: E402
from wesup_b200.models.wesup import WESUPPixelInference  # noqa: E402


def synthetic_slide(size, base=2000):
    """H&E-like slide built by tiling a `base`-pixel synthetic image (mirrored so seams are continuous)."""
    base = min(base, size)
    img, _ = synth.he_like_image(base, base, seed=77)
    reps = -(-size // base)
    row = np.concatenate([img if i % 2 == 0 else img[:, ::-1] for i in range(reps)], axis=1)
    full = np.concatenate([row if i % 2 == 0 else row[::-1] for i in range(reps)], axis=0)
    return np.ascontiguousarray(full[:size, :size])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4000)
    ap.add_argument("--patch", type=int, default=400)
    ap.add_argument("--mode", choices=["sp", "pixel"], default="sp")
    ap.add_argument("--hc-dtype", choices=["fp32", "bf16"], default="fp32")
    ap.add_argument("--no-graph", action="store_true", help="eager tile steps (default: CUDA graphs)")
    ap.add_argument("--no-footprints", action="store_true", help="sp mode: pooling kernels rebuild the footprints internally")
    ap.add_argument("--materialize", action="store_true", help="sp mode: write the (H*W,2112) hypercolumn, then pool it")
    args = ap.parse_args()
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    hc_dtype = torch.bfloat16 if args.hc_dtype == "bf16" else torch.float32
    slide = synthetic_slide(args.size)
    n_tiles = len(tiles.top_left_coordinates(args.size, args.size, args.patch))
    if args.mode == "sp":
        trainer = initialize_trainer("wesup", device=dev, pretrained=False, hc_dtype=hc_dtype,
                                     materialize_hypercolumn=args.materialize, cuda_graph=not args.no_graph,
                                     footprints=not args.no_footprints)
        trainer.model.eval()
        step, prefetch = trainer.predict_labels, trainer.prefetch
        out_dtype = torch.uint8
    else:
        model = WESUPPixelInference(pretrained=False, hc_dtype=hc_dtype).to(dev).eval()

        def eager(x):
            with torch.no_grad():
                return model(x)[..., 1]
        step, prefetch = (eager if args.no_graph else tiles.GraphedStep(eager)), None
        out_dtype = None
    # warm-up on a few tiles (cuDNN autotune, allocator)
    warm = slide[: args.patch * 2, : args.patch * 2]
    tiles.predict_tiles(step, warm, args.patch, dev, 0, 1, out_dtype=out_dtype, prefetch=prefetch)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    merged = tiles.predict_tiles(step, slide, args.patch, dev, rank, world, out_dtype=out_dtype, prefetch=prefetch)
    e.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = torch.tensor([s.elapsed_time(e)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank == 0:
        assert merged.shape[:2] == (args.size, args.size)
        print(json.dumps({"metric": f"tiled inference ({args.mode}) tiles/s", "value": n_tiles / (float(ms.item()) / 1e3),
                          "unit": "tiles/s", "n_gpus": world, "tiles": n_tiles, "patch": args.patch, "slide": args.size,
                          "device_ms_max_over_ranks": float(ms.item()), "wall_s_incl_gather_merge": wall,
                          "hc_dtype": args.hc_dtype, "graph": not args.no_graph, "positive_fraction": float(np.mean(merged > 0.5))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
