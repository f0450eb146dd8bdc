"""Minimal image-folder readers with the *output contract* of the reference's
data layer (/root/reference/utils/data.py:33-160,430-528): `img` fp32 (3,H,W)
in [0,1], `pixel_mask` / `point_mask` int64 (C,H,W) one-hot-or-zero, the 0-dim
sentinel when a mask is absent.  Augmentation (albumentations) and the dataset
preparation scripts are CPU tooling outside the hot path (SURVEY.md section 2,
rows 13 and 17); pass the reference's own `utils.data` datasets to the trainer
when they are needed -- `WESUPTrainer.get_default_dataset` prefers them when
importable.  Only PIL + numpy are used here.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

from . import empty_tensor


def imread(path) -> np.ndarray:
    return np.asarray(Image.open(str(path)))


def to_tensor(img: np.ndarray) -> torch.Tensor:
    """uint8 (H,W,3) -> fp32 (3,H,W) in [0,1]; what TF.to_tensor yields (utils/data.py:136)."""
    if img.ndim == 2:
        img = np.repeat(img[..., None], 3, axis=-1)
    img = np.ascontiguousarray(img[..., :3])
    return torch.from_numpy(img).permute(2, 0, 1).float().div_(255.0)


def one_hot_mask(mask: np.ndarray, n_classes: int) -> torch.Tensor:
    """(H,W) class ids -> (C,H,W) int64 planes (utils/data.py:140-142)."""
    planes = np.stack([(mask == i) for i in range(n_classes)]).astype("int64")
    return torch.from_numpy(planes)


def _resize(img: np.ndarray, size, nearest: bool) -> np.ndarray:
    h, w = size
    pil = Image.fromarray(img)
    return np.asarray(pil.resize((w, h), Image.NEAREST if nearest else Image.BILINEAR))


class SegmentationDataset(Dataset):
    """`<root>/images/*` (+ optional `<root>/masks/*`): returns `(img, mask)`.
    Supports `rescale_factor`, `multiscale_range` and `target_size` like the
    reference; no photometric / elastic augmentation."""

    def __init__(self, root_dir, target_size=None, rescale_factor=None, multiscale_range=None, train=True,
                 proportion=1, n_classes=2, seed=0, **_ignored):
        self.root_dir = Path(root_dir).expanduser()
        self.img_paths = sorted((self.root_dir / "images").iterdir())
        masks = self.root_dir / "masks"
        self.mask_paths = sorted(masks.iterdir()) if masks.exists() else None
        self.target_size, self.rescale_factor, self.multiscale_range = target_size, rescale_factor, multiscale_range
        self.train, self.proportion, self.n_classes = train, proportion, n_classes
        self.picked = np.arange(len(self.img_paths))
        if proportion < 1:
            rng = np.random.RandomState(seed)
            rng.shuffle(self.picked)
            self.picked = np.sort(self.picked[: len(self)])

    def __len__(self):
        return int(self.proportion * len(self.img_paths))

    def _target(self, height, width):
        if self.target_size is not None:
            return tuple(self.target_size)
        factor = None
        if self.train and self.multiscale_range is not None:
            factor = np.random.uniform(*self.multiscale_range)
        elif self.rescale_factor is not None:
            factor = self.rescale_factor
        if factor is None:
            return height, width
        return int(np.ceil(factor * height)), int(np.ceil(factor * width))

    def _load(self, idx):
        idx = int(self.picked[idx])
        img = imread(self.img_paths[idx])
        mask = imread(self.mask_paths[idx]) if self.mask_paths is not None else None
        size = self._target(*img.shape[:2])
        if size != img.shape[:2]:
            img = _resize(img, size, nearest=False)
            if mask is not None:
                mask = _resize(mask, size, nearest=True)
        return img, mask

    def __getitem__(self, idx):
        img, mask = self._load(idx)
        if mask is not None and mask.ndim == 3:
            mask = mask[..., 0]
        if mask is not None and mask.max() > self.n_classes - 1:
            mask = (mask > 0).astype("uint8")            # 0/255 PNG masks
        return to_tensor(img), (one_hot_mask(mask, self.n_classes) if mask is not None else empty_tensor())


class PointDataset(SegmentationDataset):
    """`<root>/points/<stem>.csv` rows `x,y,class` -> single-pixel point mask
    (radius 0, utils/data.py:501-508).  Returns `(img, pixel_mask, point_mask)`."""

    def __init__(self, root_dir, **kwargs):
        super().__init__(root_dir, **kwargs)
        self.point_dir = self.root_dir / "points"

    def __getitem__(self, idx):
        img, mask = self._load(idx)
        path = self.img_paths[int(self.picked[idx])]
        orig_h, orig_w = imread(path).shape[:2]
        h, w = img.shape[:2]
        pts = np.loadtxt(self.point_dir / f"{path.stem}.csv", delimiter=",", ndmin=2, dtype=np.float64)
        point_mask = np.zeros((self.n_classes, h, w), dtype="int64")
        for x, y, cls in pts[:, :3]:
            yy = min(h - 1, int(y * h / orig_h))
            xx = min(w - 1, int(x * w / orig_w))
            point_mask[int(cls), yy, xx] = 1
        if mask is not None and mask.ndim == 3:
            mask = mask[..., 0]
        if mask is not None and mask.max() > self.n_classes - 1:
            mask = (mask > 0).astype("uint8")
        pixel_mask = one_hot_mask(mask, self.n_classes) if mask is not None else empty_tensor()
        return to_tensor(img), pixel_mask, torch.from_numpy(point_mask)
