"""Second, independent CPU restatement of scikit-image SLIC -- TEST INFRASTRUCTURE ONLY.

Written from SURVEY.md Appendix B without looking at oracle/slic_ref.c, in a different
style on purpose (vectorised numpy over all pixels x one centre, a deque-based breadth-first
search for connectivity) so that the two restatements pin each other: tests require them to
agree on small images.  Targets the semantics of scikit-image 0.15-0.18 (`slic(image,
n_segments, compactness)` with max_iter=10, sigma=0, convert2lab, enforce_connectivity,
min_size_factor=0.5, max_size_factor=3, 0-based labels), the releases contemporary with the
reference (/root/reference/requirements.txt:10 pins no version; /root/reference/models/wesup.py:34-47
indexes superpixels from 0, which rules out >= 0.19's start_label=1 default).  PARITY UNPINNED against
the real package (absent from this image).  Slow: small images only.
"""
from __future__ import annotations

from collections import deque

import numpy as np


def rgb2lab(rgb: np.ndarray) -> np.ndarray:
    """skimage.color.rgb2lab (D65, 2 degree observer) for float rgb in [0,1], (H,W,3) -> float64."""
    v = rgb.astype(np.float64)
    lin = np.where(v > 0.04045, np.power((v + 0.055) / 1.055, 2.4), v / 12.92)
    m = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
    xyz = np.stack([lin[..., 0] * m[r, 0] + lin[..., 1] * m[r, 1] + lin[..., 2] * m[r, 2] for r in range(3)], axis=-1)
    xyz = xyz / np.array([0.95047, 1.0, 1.08883])
    f = np.where(xyz > 0.008856, np.cbrt(xyz), 7.787 * xyz + 16.0 / 116.0)
    return np.stack([116.0 * f[..., 1] - 16.0, 500.0 * (f[..., 0] - f[..., 1]), 200.0 * (f[..., 1] - f[..., 2])], axis=-1)


def seed_grid(h: int, w: int, n_segments: int):
    """skimage.util.regular_grid on the (1,H,W) volume: in-plane step and first seed."""
    s = np.sqrt(h * w / n_segments)
    step = int(np.round(s))            # half to even, like np.round
    return max(step, 1), int(np.floor(s / 2.0))


def kmeans(lab_scaled: np.ndarray, n_segments: int, max_iter: int = 10):
    h, w, _ = lab_scaled.shape
    step, start = seed_grid(h, w, n_segments)
    ys, xs = np.arange(start, h, step), np.arange(start, w, step)
    cent = np.zeros((len(ys) * len(xs), 5))
    cent[:, 0] = np.repeat(ys, len(xs))
    cent[:, 1] = np.tile(xs, len(ys))
    yy, xx = np.mgrid[:h, :w]
    weight = 1.0 / float(np.float32(step) * np.float32(step))
    nearest = np.zeros((h, w), np.int64)
    for _ in range(max_iter):
        dist = np.full((h, w), np.finfo(np.float64).max)
        for k, (cy, cx, cl, ca, cb) in enumerate(cent):
            if np.isnan(cy) or np.isnan(cx):
                continue
            y0, y1 = int(max(cy - 2 * step, 0)), int(min(cy + 2 * step + 1, h))
            x0, x1 = int(max(cx - 2 * step, 0)), int(min(cx + 2 * step + 1, w))
            if y0 >= y1 or x0 >= x1:
                continue
            win = (slice(y0, y1), slice(x0, x1))
            dy = (cy - yy[win]) * (cy - yy[win])
            dx = (cx - xx[win]) * (cx - xx[win])
            px = lab_scaled[win]
            t0, t1, t2 = px[..., 0] - cl, px[..., 1] - ca, px[..., 2] - cb
            d = (dy + dx) * weight + ((t0 * t0 + t1 * t1) + t2 * t2)
            better = dist[win] > d
            dist[win] = np.where(better, d, dist[win])
            nearest[win] = np.where(better, k, nearest[win])
        flat = nearest.ravel()
        n = np.bincount(flat, minlength=len(cent)).astype(np.float64)
        sums = [np.bincount(flat, weights=v.ravel(), minlength=len(cent)) for v in
                (yy.astype(np.float64), xx.astype(np.float64), lab_scaled[..., 0], lab_scaled[..., 1], lab_scaled[..., 2])]
        with np.errstate(invalid="ignore", divide="ignore"):
            cent = np.stack(sums, axis=1) / n[:, None]
    return nearest


def enforce_connectivity(seg: np.ndarray, min_size: int, max_size: int):
    """_enforce_label_connectivity_cython: raster scan, 4-connected breadth-first search capped at max_size
    (neighbour order +x, -x, +y, -y), pieces below min_size take the label of the last labelled neighbour seen."""
    h, w = seg.shape
    out = -np.ones((h, w), np.int64)
    nxt = 0
    for y in range(h):
        for x in range(w):
            if out[y, x] >= 0:
                continue
            lab, adjacent = seg[y, x], 0
            out[y, x] = nxt
            members, queue = [(y, x)], deque([(y, x)])
            while queue and len(members) < max_size:
                cy, cx = queue.popleft()
                for dy, dx in ((0, 1), (0, -1), (1, 0), (-1, 0)):
                    qy, qx = cy + dy, cx + dx
                    if not (0 <= qy < h and 0 <= qx < w):
                        continue
                    if seg[qy, qx] == lab and out[qy, qx] == -1:
                        out[qy, qx] = nxt
                        members.append((qy, qx))
                        queue.append((qy, qx))
                        if len(members) >= max_size:
                            break
                    elif out[qy, qx] >= 0 and out[qy, qx] != nxt:
                        adjacent = out[qy, qx]
            if len(members) < min_size:
                for my, mx in members:
                    out[my, mx] = adjacent
            else:
                nxt += 1
    return out, nxt


def slic(img_hwc: np.ndarray, n_segments: int, compactness: float = 10.0, max_iter: int = 10,
         enforce: bool = True):
    h, w, _ = img_hwc.shape
    lab = rgb2lab(img_hwc.astype(np.float32)) * (1.0 / compactness)
    nearest = kmeans(lab, n_segments, max_iter)
    if not enforce:
        return nearest
    seg_size = h * w / n_segments
    out, _ = enforce_connectivity(nearest, int(0.5 * seg_size), int(3.0 * seg_size))
    return out
