#!/usr/bin/env python
"""Pixel-wise tiled inference with the reference's interface
(/root/reference/pixel_infer_tile.py:18-60):

    python pixel_infer_tile.py <data_root> -c <checkpoint> [-p 300] [-o <output>]

`WESUPPixelInference` loads the same `model_state_dict` as `WESUP`.  Additive:
under torchrun the tiles are partitioned across ranks (see infer_tile.py).
"""
import argparse
import os
from pathlib import Path

import numpy as np
import torch
from PIL import Image

from wesup_b200 import parallel
from wesup_b200.models.wesup import WESUPPixelInference
from wesup_b200.tiles import PixelTileEngine, predict_tiles
from wesup_b200.utils.data import imread


def predict(model, img, patch_size, device, rank=0, world_size=1, engine=None):
    """Class-1 probability (H,W) of `img` (rank 0; None elsewhere), tiles in batches (one CUDA graph per batch shape)."""
    return predict_tiles(engine or PixelTileEngine(model), img, patch_size, device, rank, world_size)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("data_root")
    parser.add_argument("-c", "--checkpoint", required=True)
    parser.add_argument("-p", "--patch-size", type=int, default=300)
    parser.add_argument("-o", "--output")
    parser.add_argument("--hc-dtype", choices=["fp32", "bf16"], default="fp32")
    args = parser.parse_args(argv)
    rank, world, local = parallel.init_from_env()
    data_root = Path(args.data_root).expanduser()
    ckpt_path = Path(args.checkpoint).expanduser()
    device = f"cuda:{local}"
    output_dir = Path(args.output).expanduser() if args.output else \
        ckpt_path.parent.parent / f"results-pixel-tile-{args.patch_size}" / data_root.name
    if rank == 0:
        os.makedirs(output_dir, exist_ok=True)
    model = WESUPPixelInference(pretrained=False,
                                hc_dtype=torch.bfloat16 if args.hc_dtype == "bf16" else torch.float32).to(device)
    model.load_state_dict(torch.load(ckpt_path, map_location=device)["model_state_dict"])
    model.eval()
    if rank == 0:
        print("Making inference ...")
    engine = PixelTileEngine(model)    # every full tile has the same shape: one CUDA graph per batch, replayed
    for img_path in sorted((data_root / "images").iterdir()):
        final = predict(model, imread(img_path), args.patch_size, device, rank, world, engine=engine)
        if rank == 0:
            Image.fromarray(np.round(final).astype("uint8") * 255).save(output_dir / img_path.name)


if __name__ == "__main__":
    main()
