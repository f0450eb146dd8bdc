#!/usr/bin/env python
"""Superpixel-wise tiled inference with the reference's interface
(/root/reference/infer_tile.py:23-182): `divide_image_to_patches`,
`combine_patches_to_image`, `predict`, `save_predictions`, `infer`,
`main(data_dir, model_type='wesup', patch_size=464, checkpoint=None, output_dir=None, device=None)`.

Additive: under torchrun the tiles of every image are partitioned across the
ranks (one process per GPU, contiguous stripes) and only finished tile
predictions are gathered to rank 0, which merges and saves them.
"""
import os
import os.path as osp
from pathlib import Path

import numpy as np
import torch
from PIL import Image

from wesup_b200 import cli, parallel
from wesup_b200.models import initialize_trainer
from wesup_b200.tiles import (SuperpixelTileEngine, combine_patches_to_image, divide_image_to_patches,  # noqa: F401  (re-exported)
                              predict_tiles)
from wesup_b200.utils.data import imread


def predict(trainer, img_path, patch_size, device="cuda", rank=0, world_size=1):
    """(H,W) prediction for one image (rank 0; None on the other ranks)."""
    img = imread(img_path) if not isinstance(img_path, np.ndarray) else img_path

    # batches of tiles: one batched GPU SLIC + VGG16 at batch size per step, preprocessing one batch ahead on a side
    # stream, predictions merged on the device (wesup_b200.tiles)
    engine = getattr(trainer, "_tile_engine", None)
    if engine is None:
        engine = trainer._tile_engine = SuperpixelTileEngine(trainer, batch=int(trainer.kwargs.get("tile_batch", 16)),
                                                             use_graph=bool(trainer.kwargs.get("cuda_graph", True)))
    return predict_tiles(engine, img, patch_size, device, rank, world_size)


def save_predictions(predictions, img_paths, output_dir="predictions"):
    print(f"\nSaving prediction to {output_dir} ...")
    os.makedirs(output_dir, exist_ok=True)
    for pred, img_path in zip(predictions, img_paths):
        Image.fromarray(pred.astype("uint8") * 255).save(osp.join(output_dir, osp.basename(img_path)))


def infer(trainer, data_dir, patch_size, output_dir=None, device="cuda", rank=0, world_size=1):
    data_dir = Path(data_dir).expanduser()
    img_paths = sorted((data_dir / "images").iterdir())
    if rank == 0:
        print(f"Predicting {len(img_paths)} images from {data_dir} ...")
    trainer.model.eval()
    predictions = [predict(trainer, p, patch_size, device=device, rank=rank, world_size=world_size) for p in img_paths]
    if output_dir is not None and rank == 0:
        save_predictions(predictions, img_paths, output_dir)
    return predictions


def main(data_dir, model_type="wesup", patch_size=464, checkpoint=None, output_dir=None, device=None, **kwargs):
    # the reference defaults model_type to 'mild', which its own factory rejects
    rank, world, local = parallel.init_from_env()
    if output_dir is None and checkpoint is not None:
        output_dir = Path(checkpoint).expanduser().parent.parent / "results"
        output_dir.mkdir(exist_ok=True)
    device = device or (f"cuda:{local}" if world > 1 else "cuda")
    # inference never needs the (H*W,2112) tensor: superpixel means straight from the backbone levels,
    # one CUDA graph per tile shape (both can be overridden from the command line)
    kwargs = {"materialize_hypercolumn": False, "cuda_graph": True, **kwargs}
    if checkpoint is not None:
        kwargs.setdefault("pretrained", False)       # every weight comes from the checkpoint: no ImageNet download
    trainer = initialize_trainer(model_type, device=device, **kwargs)
    if checkpoint is not None:
        trainer.load_checkpoint(checkpoint)
    return infer(trainer, data_dir, patch_size, output_dir, device=device, rank=rank, world_size=world)


if __name__ == "__main__":
    cli.run(main)
