"""Mint tests/golden/slic_kats.npz: SLIC known-answer tests DERIVED BY HAND from the algorithm
description, not from any of the implementations under test.

    python tests/golden/make_slic_kats.py

scikit-image (the reference's SLIC, /root/reference/models/wesup.py:471-476) is not installable in
this image, so no output of the real package can be recorded; these vectors pin what can be pinned
without it.  Targeted semantics: scikit-image 0.15-0.18 `slic(image, n_segments, compactness)` with
its defaults of that era (max_iter=10, sigma=0, convert2lab, enforce_connectivity=True,
min_size_factor=0.5, max_size_factor=3, labels from 0).

1. uniform_*: an image of ONE colour.  Every colour term is the same for all centres, so k-means is
   purely spatial; the seeds form a product grid and stay one (the centroid of a rectangle), hence
   the 2-D result is the product of two 1-D Lloyd iterations "pixel -> nearest centre inside the
   centre's [c-2S, c+2S] window, ties to the lowest index; centre -> mean of its pixels", run here
   in exact rational arithmetic (fractions.Fraction).  All cells are rectangles of about S x S
   pixels >= min_size, so connectivity enforcement keeps them and numbers them in raster order of
   their first pixel = grid order.
2. conn_*: tiny label maps whose connectivity-enforcement result is worked out in the comments
   (raster scan; 4-connected breadth-first search with neighbour order +x, -x, +y, -y, capped at
   max_size; a piece below min_size takes the label of the last already-labelled neighbour seen,
   initially 0, and does not consume a label).
"""
from __future__ import annotations

from fractions import Fraction
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent / "slic_kats.npz"


def lloyd_1d(size: int, step: int, start: int, max_iter: int = 10):
    """Final assignment (pixel -> centre index along this axis) of the 1-D iteration, exact arithmetic."""
    cent = [Fraction(c) for c in range(start, size, step)]
    assign = [0] * size
    for _ in range(max_iter):
        best = [None] * size
        for j, c in enumerate(cent):
            if c is None:                     # empty cluster (0/0): never wins again
                continue
            lo = max(c - 2 * step, 0)
            hi = min(c + 2 * step + 1, size)
            for x in range(int(lo), int(hi)):          # int() truncates like the C cast
                d = (c - x) * (c - x)
                if best[x] is None or d < best[x]:     # strict: ties keep the lowest index
                    best[x] = d
                    assign[x] = j
        cent = []
        for j in range(len(range(start, size, step))):
            members = [x for x in range(size) if assign[x] == j]
            cent.append(Fraction(sum(members), len(members)) if members else None)
    return np.array(assign)


def uniform_case(h: int, w: int):
    n_segments = int(h * w / 200)
    s = np.sqrt(h * w / n_segments)
    step, start = int(np.round(s)), int(np.floor(s / 2.0))
    ay, ax = lloyd_1d(h, step, start), lloyd_1d(w, step, start)
    nx = len(range(start, w, step))
    raw = ay[:, None] * nx + ax[None, :]
    # every cell must survive connectivity enforcement unchanged for the closed form to hold
    seg = h * w / n_segments
    sizes = np.bincount(raw.ravel())
    assert sizes.min() >= int(0.5 * seg) and sizes.max() <= int(3 * seg), (h, w, sizes.min(), sizes.max())
    return n_segments, raw.astype(np.int32)


CONN_CASES = {
    # name: (seg, min_size, max_size, expected, n_labels)
    # strip of six 0s and two 1s, max_size 4: the search from x=0 takes x=1,2,3 and stops at the cap (piece 0); the
    # next unvisited pixel x=4 takes x=5 (piece 1; its -x neighbour is already labelled 0, irrelevant: size 2 >= 2);
    # x=6 takes x=7 (piece 2)
    "cap_strip": ([[0, 0, 0, 0, 0, 0, 1, 1]], 2, 4, [[0, 0, 0, 0, 1, 1, 2, 2]], 3),
    # the lone 7 is below min_size; the last labelled neighbour it saw is x=2 (label 0) -> becomes 0 and consumes no
    # label; the 5s to its right are a new piece, label 1
    "merge_single": ([[5, 5, 5, 7, 5, 5, 5, 5]], 2, 8, [[0, 0, 0, 0, 1, 1, 1, 1]], 2),
    # top-left pixel is small and sees no labelled neighbour: it takes the initial `adjacent` = 0 and consumes no
    # label; the rest becomes label 0 as well (its neighbour (0,0) carries out == next_label, which is skipped)
    "merge_no_neighbour": ([[9, 1, 1], [1, 1, 1], [1, 1, 1]], 2, 100, [[0, 0, 0], [0, 0, 0], [0, 0, 0]], 1),
    # 3x3 of one label, cap 4: from (0,0): +x (0,1), +y (1,0); then from (0,1): +x (0,2) -> cap.  Piece 0 =
    # {(0,0),(0,1),(1,0),(0,2)}.  Next unvisited in raster order is (1,1): +x (1,2), +y (2,1); from (1,2): +y (2,2)
    # -> cap.  Piece 1.  (2,0) is alone: piece 2 (min_size 1 keeps it)
    "cap_square": ([[4, 4, 4], [4, 4, 4], [4, 4, 4]], 1, 4, [[0, 0, 0], [0, 1, 1], [2, 1, 1]], 3),
    # same with min_size 2: (2,0) is small; neighbours in order +x (2,1) -> label 1, -y (1,0) -> label 0: the LAST one
    # seen wins -> 0
    "cap_square_merge": ([[4, 4, 4], [4, 4, 4], [4, 4, 4]], 2, 4, [[0, 0, 0], [0, 1, 1], [0, 1, 1]], 2),
    # two small pieces in a row before the first kept one: both take 0 (the second sees out == next_label == 0 on its
    # left, which is skipped, and keeps the initial 0); the 3s become label 0 too
    "merge_chain": ([[1, 2, 3, 3, 3, 3]], 2, 10, [[0, 0, 0, 0, 0, 0]], 1),
    # a small piece between two kept ones takes the label of the LAST neighbour seen in +x,-x,+y,-y order while the
    # search runs: for the single 8 at (1,1): +x (1,2) unlabelled, -x (1,0) label 0, +y (2,1) unlabelled, -y (0,1)
    # label 0 -> 0.  The 6s (bottom/right, 4 pixels, connected through (1,2)-(2,2)-(2,1)-(2,0)) are piece 1
    "merge_last_seen": ([[3, 3, 3], [3, 8, 6], [6, 6, 6]], 2, 100, [[0, 0, 0], [0, 0, 1], [1, 1, 1]], 2),
}


def main():
    out = {}
    for h, w in ((64, 80), (97, 131), (150, 150)):
        n_segments, raw = uniform_case(h, w)
        out[f"uniform_{h}x{w}_labels"] = raw
        out[f"uniform_{h}x{w}_n_segments"] = np.array(n_segments)
    for name, (seg, mn, mx, exp, n) in CONN_CASES.items():
        out[f"conn_{name}_seg"] = np.array(seg, np.int32)
        out[f"conn_{name}_sizes"] = np.array([mn, mx, n])
        out[f"conn_{name}_expected"] = np.array(exp, np.int32)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {sorted(out)}")


if __name__ == "__main__":
    main()
