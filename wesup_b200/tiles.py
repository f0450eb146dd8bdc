"""Tile split / merge for tiled inference, with the reference's geometry
(/root/reference/infer_tile.py:23-91): tile origins are
np.linspace(0, size - patch, ceil(size / patch), dtype=int) per axis, tiles may
overlap, and overlapping predictions are combined by a running mean in float64
in tile order."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np


def top_left_coordinates(height: int, width: int, patch_size: int) -> List[Tuple[int, int]]:
    tops = np.linspace(0, height - patch_size, math.ceil(height / patch_size), dtype=int)
    lefts = np.linspace(0, width - patch_size, math.ceil(width / patch_size), dtype=int)
    return [(int(t), int(l)) for t in tops for l in lefts]


def divide_image_to_patches(img: np.ndarray, patch_size: int) -> np.ndarray:
    """(H,W,3) -> (N,patch,patch,3) uint8, row-major over tile origins."""
    if img.ndim != 3 or img.shape[-1] != 3:
        raise AssertionError("img must be (H, W, 3)")
    h, w, _ = img.shape
    return np.stack([img[t:t + patch_size, l:l + patch_size] for t, l in top_left_coordinates(h, w, patch_size)]
                    ).astype("uint8")


def combine_patches_to_image(patches: np.ndarray, target_height: int, target_width: int) -> np.ndarray:
    """(N,p,p[,C]) predictions -> (H,W[,C]); where tiles overlap the result is the
    running mean of the tiles seen so far (same arithmetic order as the reference)."""
    patch_size = patches.shape[1]
    if patches.ndim == 3:
        patches = patches[..., None]
    acc = np.zeros((target_height, target_width, patches.shape[-1]), np.float64)
    seen = np.zeros((target_height, target_width, 1), np.float64)
    for tile, (t, l) in zip(patches, top_left_coordinates(target_height, target_width, patch_size)):
        win = (slice(t, t + patch_size), slice(l, l + patch_size))
        acc[win] = (acc[win] * seen[win] + tile) / (seen[win] + 1)
        seen[win] += 1.0
    return np.squeeze(acc)
