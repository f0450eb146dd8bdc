"""CPU oracle for the WESUP superpixel stage -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch-CPU fp32 ops, the algorithm of the reference
hot path (``/root/reference/models/wesup.py``).  It exists so that the CUDA
kernels in ``wesup_b200/csrc`` can be checked against an independent
implementation.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does.

Pinning: the reference ships no golden vectors, so the
oracle is pinned against outputs of the *real* reference imported in the build
container -- see ``tests/golden/make_golden.py`` (generator, committed) and
``tests/golden/*.npz`` (its outputs).  ``tests/test_oracle_golden.py`` replays
them.  The SLIC part lives in ``oracle/slic_ref.c`` and is "parity unpinned"
(scikit-image is not installable here).

Every function cites the reference lines it follows.  The oracle deliberately
keeps the reference's dense formulation (one-hot ``(N,H,W)`` maps, dense ``mm``,
full ``(N,N,D)`` affinity) because that *is* the algorithm being restated and it
is also what the CPU baseline times.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

# VGG16 `features` conv plan: (out_channels, followed_by_pool).  The reference
# takes torchvision's vgg16().features (models/wesup.py:199); the oracle builds
# the same Sequential so state_dict keys line up (backbone.0, backbone.2, ...).
VGG16_PLAN = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M",
              512, 512, 512, "M", 512, 512, 512, "M")


def vgg16_features() -> nn.Sequential:
    layers, cin = [], 3
    for item in VGG16_PLAN:
        if item == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers.append(nn.Conv2d(cin, item, kernel_size=3, padding=1))
            layers.append(nn.ReLU(inplace=True))
            cin = item
    return nn.Sequential(*layers)


# ---------------------------------------------------------------------------
# a2: superpixel preprocessing (models/wesup.py:18-63)
# ---------------------------------------------------------------------------
def superpixel_order_and_labels(segments: torch.Tensor, mask: torch.Tensor | None,
                                epsilon: float = 1e-7):
    """Return (order, sp_labels).

    order[k] = original superpixel id that becomes row k (labeled ids ascending,
    then unlabeled ids ascending -- models/wesup.py:45-47); sp_labels is the
    quantised (N_l, C) multi-hot matrix (models/wesup.py:50-52) or None when no
    mask was given (the reference returns its 0-dim sentinel, :54).
    """
    n_sp = int(segments.max()) + 1
    if mask is None or mask.dim() == 0:
        return torch.unique(segments), None
    rows = []
    for k in range(n_sp):                                   # :39-42
        inside = (segments == k).long()
        per_class = (mask * inside).float()
        rows.append(per_class.sum(dim=(1, 2)) / (per_class.sum() + epsilon))
    dist = torch.stack(rows)                                # (n_sp, C)
    total = dist.sum(dim=-1)
    labeled = torch.nonzero(total > 0).flatten()            # :45
    unlabeled = torch.nonzero(total == 0).flatten()         # :46
    order = torch.cat([labeled, unlabeled])                 # :47
    picked = dist[labeled]
    sp_labels = (picked == picked.max(dim=-1, keepdim=True)[0]).float()  # :50-52
    return order, sp_labels


def dense_sp_maps(segments: torch.Tensor, order: torch.Tensor) -> torch.Tensor:
    """(N,H,W) fp32 maps, each summing to one (models/wesup.py:57-61)."""
    maps = (segments.unsqueeze(0) == order.view(-1, 1, 1)).float()
    return maps / maps.sum(dim=(1, 2), keepdim=True)


def preprocess_superpixels(segments, mask=None, epsilon=1e-7):
    order, sp_labels = superpixel_order_and_labels(segments, mask, epsilon)
    return dense_sp_maps(segments, order), sp_labels, order


# ---------------------------------------------------------------------------
# a3: hypercolumn (models/wesup.py:246-261)
# ---------------------------------------------------------------------------
def hypercolumn_from_sides(sides, size) -> torch.Tensor:
    """Bilinear(align_corners=True) upsample of every side output to `size` and
    channel concatenation in layer order -> (C_total, H, W)."""
    ups = [F.interpolate(s, size, mode="bilinear", align_corners=True).squeeze(0)
           for s in sides]
    return torch.cat(ups, dim=0)


# ---------------------------------------------------------------------------
# a4 / a6: pooling and painting (models/wesup.py:284-285, 295-304)
# ---------------------------------------------------------------------------
def pool_dense(sp_maps: torch.Tensor, feats_chw: torch.Tensor) -> torch.Tensor:
    n = sp_maps.size(0)
    return torch.mm(sp_maps.reshape(n, -1), feats_chw.reshape(feats_chw.size(0), -1).t())


def paint_dense(sp_maps: torch.Tensor, sp_pred: torch.Tensor) -> torch.Tensor:
    """Per-pixel prediction of the superpixel that owns the pixel; returns the
    class-1 plane with a leading batch dim, (1,H,W)."""
    owner = sp_maps.argmax(dim=0)                           # :295
    canvas = torch.zeros(*owner.shape, sp_pred.size(1), device=sp_pred.device)
    for k in range(int(owner.max()) + 1):                   # :301-302
        canvas[owner == k] = sp_pred[k]
    return canvas.unsqueeze(0)[..., 1]


# ---------------------------------------------------------------------------
# a7: label propagation (models/wesup.py:99-139)
# ---------------------------------------------------------------------------
def label_propagate(features: torch.Tensor, y_l: torch.Tensor, threshold: float = 0.95,
                    return_aux: bool = False):
    f = features.detach()
    y_l = y_l.detach()
    n_l = y_l.size(0)
    n_u = f.size(0) - n_l
    diff = f - f.unsqueeze(1)                               # (N,N,D)  :122
    affinity = torch.exp(-torch.einsum("ijk,ijk->ij", diff, diff))  # :121
    block = affinity[n_l:, :n_l]                            # :126
    best, src = block.max(dim=1)                            # :130
    y_u = torch.zeros(n_u, y_l.size(1))
    take = best > threshold                                 # :136 (strict)
    y_u[take] = y_l[src[take]]
    if return_aux:
        return y_u, src, best
    return y_u


# ---------------------------------------------------------------------------
# Scalable forms of the same definitions, for the BASELINE shapes where the dense
# formulation cannot be held (CRAG 1516x1512: dense sp_maps = 105 GB, hypercolumn
# = 19 GB, (N,N,D) affinity = 17 GB).  Each is checked against its dense twin above
# at small sizes by tests/test_oracle_golden.py before the GPU tests rely on it.
# ---------------------------------------------------------------------------
def superpixel_order_and_labels_counts(segments: torch.Tensor, mask: torch.Tensor | None):
    """`superpixel_order_and_labels` from per-superpixel integer class counts (SURVEY.md Appendix
    A.2: all classes of a superpixel share one denominator, so `dist > 0` and `dist == rowmax`
    are decided by the counts).  Returns (order, sp_labels, counts-per-row)."""
    seg = segments.reshape(-1).long()
    n_sp = int(seg.max()) + 1
    sizes = torch.bincount(seg, minlength=n_sp)
    if mask is None or mask.dim() == 0:
        order = torch.unique(seg)
        return order, None, sizes[order]
    m = mask.reshape(mask.size(0), -1).long()
    per_class = torch.stack([torch.bincount(seg, weights=m[c].double(), minlength=n_sp).long()
                             for c in range(m.size(0))], dim=1)              # (n_sp, C)
    total = per_class.sum(dim=1)
    labeled = torch.nonzero(total > 0).flatten()
    unlabeled = torch.nonzero(total == 0).flatten()
    order = torch.cat([labeled, unlabeled])
    picked = per_class[labeled]
    sp_labels = (picked == picked.max(dim=1, keepdim=True)[0]).float()
    return order, sp_labels, sizes[order]


def bilinear_taps(dst: torch.Tensor, in_size: int, out_size: int):
    """(i0, i1, w0, w1) of F.interpolate(mode='bilinear', align_corners=True) along one axis, in
    the arithmetic ATen uses for fp32 tensors: scale and source coordinate in fp32
    (models/wesup.py:254-255)."""
    scale = torch.tensor((in_size - 1) / (out_size - 1) if out_size > 1 else 0.0, dtype=torch.float32)
    src = scale * dst.to(torch.float32)
    i0 = src.floor().long().clamp_(max=in_size - 1)
    i1 = i0 + (i0 < in_size - 1).long()
    w1 = src - i0.to(torch.float32)
    return i0, i1, (1.0 - w1), w1


def hypercolumn_at_pixels(levels, size, pixels: torch.Tensor) -> torch.Tensor:
    """Rows of the (H*W, C_total) hypercolumn for the given flat pixel ids only, fp64 accumulation
    of fp32 tap weights; `levels` are (1,C_l,h_l,w_l) tensors (any device)."""
    H, W = size
    y, x = pixels // W, pixels % W
    cols = []
    for lv in levels:
        _, c, h, w = lv.shape
        y0, y1, wy0, wy1 = bilinear_taps(y.cpu(), h, H)
        x0, x1, wx0, wx1 = bilinear_taps(x.cpu(), w, W)
        dev = lv.device
        y0, y1, x0, x1 = (t.to(dev) for t in (y0, y1, x0, x1))
        wy0, wy1, wx0, wx1 = (t.to(dev).double().unsqueeze(1) for t in (wy0, wy1, wx0, wx1))
        f = lv[0].permute(1, 2, 0)                                     # (h, w, C) view
        top = wx0 * f[y0, x0].double() + wx1 * f[y0, x1].double()
        bot = wx0 * f[y1, x0].double() + wx1 * f[y1, x1].double()
        cols.append(wy0 * top + wy1 * bot)
    return torch.cat(cols, dim=1)


def pooled_rows_sparse(levels, size, row_of_pixel: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """Superpixel means (models/wesup.py:284-285) of the hypercolumn for the selected rows only:
    (len(rows), C_total) fp64.  `row_of_pixel` is the flat (H*W) row index of every pixel."""
    out = []
    for r in rows.tolist():
        px = torch.nonzero(row_of_pixel == r).flatten()
        out.append(hypercolumn_at_pixels(levels, size, px).mean(dim=0))
    return torch.stack(out)


def label_propagate_block(features: torch.Tensor, y_l: torch.Tensor, threshold: float = 0.95, chunk: int = 512):
    """`label_propagate` evaluating only the (n_u, n_l) block it uses, `chunk` unlabeled rows at a
    time, with the same per-pair fp32 arithmetic (difference, einsum over the feature axis, exp)."""
    f = features.detach()
    n_l = y_l.size(0)
    lab = f[:n_l]
    best_all, src_all = [], []
    for lo in range(n_l, f.size(0), chunk):
        diff = lab.unsqueeze(0) - f[lo:lo + chunk].unsqueeze(1)        # (chunk, n_l, D): f_j - f_i as at :122
        aff = torch.exp(-torch.einsum("ijk,ijk->ij", diff, diff))
        best, src = aff.max(dim=1)
        best_all.append(best)
        src_all.append(src)
    best = torch.cat(best_all) if best_all else torch.zeros(0)
    src = torch.cat(src_all) if src_all else torch.zeros(0, dtype=torch.long)
    y_u = torch.zeros(best.numel(), y_l.size(1))
    take = best > threshold
    y_u[take] = y_l.detach()[src[take]]
    return y_u, src, best


# ---------------------------------------------------------------------------
# a8 / a9: loss (models/wesup.py:66-96, 492-531)
# ---------------------------------------------------------------------------
def cross_entropy(y_hat, y_true, class_weights=None, epsilon=1e-7):
    y_hat = torch.clamp(y_hat, min=epsilon, max=1 - epsilon)
    n_rows = torch.sum(y_true.sum(dim=1) > 0).float()
    if n_rows.item() == 0:
        return torch.tensor(0.0)
    ce = -y_true * torch.log(y_hat)
    if class_weights is not None:
        ce = ce * class_weights.unsqueeze(0).float()
    return ce.sum() / n_rows


def compute_loss(sp_pred, sp_features, sp_labels, enable_propagation=True,
                 propagate_threshold=0.8, propagate_weight=0.5, metrics=None):
    total, n_l = sp_pred.size(0), sp_labels.size(0)
    if n_l < total:
        loss = cross_entropy(sp_pred[:n_l], sp_labels)
        if enable_propagation:
            y_u = label_propagate(sp_features, sp_labels, propagate_threshold)
            p_loss = cross_entropy(sp_pred[n_l:], y_u)
            loss = loss + propagate_weight * p_loss
            if metrics is not None:
                metrics["propagated_labels"] = y_u.sum().item()
                metrics["propagate_loss"] = float(p_loss)
        if metrics is not None:
            metrics["labeled_sp_ratio"] = n_l / total
        return loss
    return cross_entropy(sp_pred, sp_labels)


# ---------------------------------------------------------------------------
# the module (models/wesup.py:182-304) -- same parameter names as the reference
# ---------------------------------------------------------------------------
class RefWESUP(nn.Module):
    def __init__(self, n_classes: int = 2, D: int = 32):
        super().__init__()
        self.backbone = vgg16_features()
        self.side_offsets = []
        total = 0
        for layer in self.backbone:
            if isinstance(layer, nn.Conv2d):
                half = layer.out_channels // 2
                setattr(self, f"side_conv{total}", nn.Conv2d(layer.out_channels, half, 1))
                self.side_offsets.append(total)
                total += half
        self.fm_channels_sum = total
        self.fc_layers = nn.Sequential(
            nn.Linear(total, 1024), nn.ReLU(),
            nn.Linear(1024, 1024), nn.ReLU(),
            nn.Linear(1024, D), nn.ReLU())
        self.classifier = nn.Sequential(nn.Linear(D, n_classes), nn.Softmax(dim=1))
        self.sp_features = None
        self.sp_pred = None

    def side_outputs(self, x):
        """1x1 side conv on every *pre-ReLU* conv output (hook fires on the
        Conv2d module, models/wesup.py:205-210,253)."""
        outs, it = [], iter(self.side_offsets)
        for layer in self.backbone:
            x = layer(x)
            if isinstance(layer, nn.Conv2d):
                outs.append(getattr(self, f"side_conv{next(it)}")(x.clone()))
        return outs

    def hypercolumn(self, x):
        return hypercolumn_from_sides(self.side_outputs(x), x.shape[-2:])

    def forward(self, inputs):
        x, sp_maps = inputs
        feats = self.hypercolumn(x)
        pooled = pool_dense(sp_maps, feats)
        self.sp_features = self.fc_layers(pooled)
        self.sp_pred = self.classifier(self.sp_features)
        return paint_dense(sp_maps, self.sp_pred)

    def forward_pixels(self, x):
        """WESUPPixelInference.forward (models/wesup.py:382-400)."""
        h, w = x.shape[-2:]
        feats = self.hypercolumn(x)
        out = self.classifier(self.fc_layers(feats.reshape(feats.size(0), -1).t()))
        return out.view(h, w, -1)


def seeded_init_(module: nn.Module, seed: int = 0) -> nn.Module:
    """Deterministic, name-keyed He-style init used by goldens and tests so the
    reference module, the oracle and the product module share weights without
    shipping a 75 MB checkpoint.  Not part of the reference."""
    sd = module.state_dict()
    for i, name in enumerate(sorted(sd)):
        t = sd[name]
        g = torch.Generator().manual_seed(seed * 1000 + i)
        if t.dim() >= 2:
            fan_in = t[0].numel()
            vals = torch.randn(t.shape, generator=g) * (2.0 / fan_in) ** 0.5
        else:
            vals = torch.randn(t.shape, generator=g) * 0.05
        t.copy_(vals)
    return module
