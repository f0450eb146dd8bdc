"""Synthetic H&E-like images and sparse point labels.

The reference has no data generator; its input contract comes from
utils/data.py:136-142,501-508 (img fp32 (3,H,W) in [0,1] via TF.to_tensor of a
uint8 image; masks int64 (C,H,W) one-hot-or-zero) and the point density from
scripts/generate_points.py:48-78 (ratio 1e-4).  Everything here is host-side
numpy; tensors are returned on the CPU so callers decide about pinning / H2D.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.ndimage import gaussian_filter

EOSIN = np.array([0.91, 0.65, 0.80])
HAEMATOXYLIN = np.array([0.45, 0.30, 0.60])


def he_like_image(height: int, width: int, seed: int):
    """Returns (img uint8 (H,W,3), gland (H,W) in {0,1})."""
    rng = np.random.default_rng(seed)
    low = gaussian_filter(rng.standard_normal((height, width)), sigma=12.0, mode="reflect")
    gland = (low > np.median(low)).astype(np.int64)
    texture = gaussian_filter(rng.standard_normal((height, width)), sigma=3.0, mode="reflect")
    texture = texture / (np.abs(texture).max() + 1e-12)
    img = np.where(gland[..., None] == 1, HAEMATOXYLIN, EOSIN)
    img = img + 0.08 * texture[..., None] + rng.normal(0.0, 0.03, (height, width, 3))
    img = np.clip(img, 0.0, 1.0)
    return np.round(img * 255.0).astype(np.uint8), gland


def to_tensor(img_u8: np.ndarray) -> torch.Tensor:
    """What torchvision's TF.to_tensor yields for a uint8 HWC image."""
    return torch.from_numpy(np.ascontiguousarray(img_u8.transpose(2, 0, 1))).float().div(255.0)


def point_mask(gland: np.ndarray, seed: int, ratio: float = 1e-4, n_classes: int = 2):
    """K = max(2, int(H*W*ratio)) single-pixel labels; class = gland value."""
    h, w = gland.shape
    k = max(2, int(h * w * ratio))
    rng = np.random.default_rng(seed)
    flat = rng.choice(h * w, size=k, replace=False)
    mask = np.zeros((n_classes, h, w), np.int64)
    ys, xs = np.unravel_index(flat, (h, w))
    mask[gland[ys, xs], ys, xs] = 1
    return torch.from_numpy(mask)


def pixel_mask(gland: np.ndarray, n_classes: int = 2) -> torch.Tensor:
    return torch.from_numpy(np.stack([(gland == c) for c in range(n_classes)]).astype(np.int64))


def sample(height: int, width: int, index: int = 0, ratio: float = 1e-4):
    """One training datum as the DataLoader would deliver it (batch dim 1):
    img (1,3,H,W) fp32, pixel_mask (1,2,H,W) int64, point_mask (1,2,H,W) int64."""
    img_u8, gland = he_like_image(height, width, seed=1000 + index)
    return (to_tensor(img_u8).unsqueeze(0), pixel_mask(gland).unsqueeze(0),
            point_mask(gland, seed=2000 + index, ratio=ratio).unsqueeze(0))


def perturbed_grid_segments(height: int, width: int, cell: int, seed: int) -> np.ndarray:
    """A 0-based contiguous superpixel-like label map without running SLIC:
    Voronoi cells of a jittered grid.  Used where a test needs ragged,
    non-rectangular superpixels of a known count."""
    rng = np.random.default_rng(seed)
    ys = np.arange(cell // 2, height, cell)
    xs = np.arange(cell // 2, width, cell)
    cy, cx = np.meshgrid(ys, xs, indexing="ij")
    cy = cy.ravel() + rng.uniform(-cell / 3, cell / 3, cy.size)
    cx = cx.ravel() + rng.uniform(-cell / 3, cell / 3, cx.size)
    yy, xx = np.mgrid[:height, :width]
    best = np.full((height, width), np.inf)
    lab = np.zeros((height, width), np.int64)
    for k in range(cy.size):
        d = (yy - cy[k]) ** 2 + (xx - cx[k]) ** 2
        win = d < best
        best[win] = d[win]
        lab[win] = k
    _, inv = np.unique(lab, return_inverse=True)
    return inv.reshape(height, width).astype(np.int64)
