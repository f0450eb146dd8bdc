"""B200-native mirror of the reference's `models/wesup.py`.

Same public surface (/root/reference/models/wesup.py): `WESUPConfig`, `WESUP`
(`forward((x, sp_maps))`, attributes `sp_features` / `sp_pred` / `feature_maps`,
identical `state_dict` keys), `WESUPPixelInference`, `WESUPTrainer`
(`preprocess` / `compute_loss` / `postprocess` / ...), and the module-level
`_preprocess_superpixels`, `_cross_entropy`, `_label_propagate`.  Underneath,
the superpixel stage runs on the hand-written CUDA kernels behind the C ABI
(`wesup_b200.ops`); VGG16 convolutions and the small MLP stay on
PyTorch/cuDNN/cuBLAS as in the reference.  There is no CPU path.
"""
from __future__ import annotations

import os.path as osp
import warnings
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..ops import SuperpixelMaps
from ..utils import empty_tensor, is_empty_tensor
from .base import BaseConfig, BaseTrainer

# torchvision's vgg16().features layout (reference: models/wesup.py:199); built
# here so that checkpoint keys are `backbone.{0,2,5,...}.{weight,bias}`.
_VGG16 = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M")


def _vgg16_features(pretrained: bool, allow_random_init: bool = False) -> nn.Sequential:
    layers, cin = [], 3
    for item in _VGG16:
        if item == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, item, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            cin = item
    features = nn.Sequential(*layers)
    if pretrained:
        try:            # ImageNet weights as in the reference (models/wesup.py:198)
            from torchvision import models as tvm
            features.load_state_dict(tvm.vgg16(weights=tvm.VGG16_Weights.IMAGENET1K_V1).features.state_dict())
        except Exception as ex:  # noqa: BLE001  (URLError / OSError / missing torchvision)
            # the reference fails here; training from scratch with the same command line gives very different
            # results, so random init has to be asked for (pretrained=False, or allow_random_init=True)
            if not allow_random_init:
                raise RuntimeError(f"VGG16 ImageNet weights unavailable ({type(ex).__name__}: {ex}); pass pretrained=False "
                                   "or allow_random_init=True to train from random initialisation") from ex
            warnings.warn(f"VGG16 ImageNet weights unavailable ({type(ex).__name__}); using random init (allow_random_init)")
    return features


# ---------------------------------------------------------------------------
# module-level functions (reference: models/wesup.py:18-139)
# ---------------------------------------------------------------------------
def _preprocess_superpixels(segments, mask=None, epsilon=1e-7, dense=False, n_sp=None, n_sp_dev=None):
    """Superpixel rows + labels from a SLIC label map (reference :18-63).

    Returns `(sp_maps, sp_labels)`.  `sp_maps` is a compact `SuperpixelMaps`
    (label map + counts + CSR) instead of the dense `(N,H,W)` tensor; pass
    `dense=True` for the reference tensor.  Row order, the labeled/unlabeled
    split and the multi-hot quantisation are bit-identical to the reference.
    `epsilon` only guards a division whose result is compared with zero / with
    the row maximum in the reference, so integer class counts are equivalent.
    `n_sp` / `n_sp_dev`: see `SuperpixelMaps.from_labels` (one host sync in total).
    """
    if mask is not None and is_empty_tensor(mask):
        mask = None
    sp = SuperpixelMaps.from_labels(segments, mask, n_sp=n_sp, n_sp_dev=n_sp_dev)
    sp_labels = sp.sp_labels if mask is not None else empty_tensor().to(segments.device)
    return (sp.to_dense() if dense else sp), sp_labels


def _cross_entropy(y_hat, y_true, class_weights=None, epsilon=1e-7):
    """Semi-supervised cross entropy (reference :66-96).  Rows whose label is
    all-zero are ignored; multi-hot rows count once in the denominator.  Unlike
    the reference this never syncs with the host: with no labeled row the
    numerator is exactly zero, so dividing by max(count, 1) returns the same 0."""
    y_hat = torch.clamp(y_hat, min=epsilon, max=1 - epsilon)
    labeled = (y_true.sum(dim=1) > 0).sum().float()
    ce = -y_true * torch.log(y_hat)
    if class_weights is not None:
        ce = ce * class_weights.unsqueeze(0).float()
    return ce.sum() / torch.clamp(labeled, min=1.0)


def _label_propagate(features, y_l, threshold=0.95):
    """Nearest-labeled-neighbour label propagation (reference :99-139) as one
    fused kernel; returns the (n_u, C) pseudo-label matrix."""
    return ops.label_propagate(features, y_l, threshold)


class WESUPConfig(BaseConfig):
    """Field-for-field the reference's configuration (models/wesup.py:142-179)."""
    rescale_factor = 0.5
    multiscale_range = (0.3, 0.4)
    n_classes = 2
    class_weights = (3, 1)          # never applied by the reference either (:434)
    sp_area = 200
    sp_compactness = 40
    enable_propagation = True
    propagate_threshold = 0.8
    propagate_weight = 0.5
    momentum = 0.9
    weight_decay = 0.001
    freeze_backbone = False
    batch_size = 1
    epochs = 300


class WESUP(nn.Module):
    """Reference: models/wesup.py:182-304.

    Extra keyword arguments (all optional, none stored in the state_dict):
      pretrained    load ImageNet VGG16 weights (default True; raises when they cannot be fetched, like the reference,
                    unless allow_random_init=True)
      hc_dtype      torch.float32 (default) or torch.bfloat16 hypercolumn storage
      hc_layout     'hwc' (default, pixel-major) or 'chw' (the reference's layout)
      fused_backward  True (default): backward of pooling + hypercolumn is one fused
                    kernel working from the pooled gradient (hwc layout only)
      materialize_hypercolumn  False (default, the benched path): the superpixel means are pooled
                    straight from the levels, the (H*W,2112) tensor is never written and
                    `feature_maps` (internal to the reference's forward, models/wesup.py:246-285)
                    stays None -- same numbers, no 1.8 GB round trip.  True: kernel (a)
                    writes the tensor that `feature_maps` exposes, kernel (b) pools it
      pool_first    True (default; only with materialize_hypercolumn=False): the superpixel means
                    are taken from the 13 backbone conv outputs (4224 channels) and the 1x1 side
                    convolutions then run on the N pooled rows instead of on H*W pixels -- a mean
                    and a 1x1 convolution commute, so `sp_features`/`sp_pred`/loss/gradients are the
                    reference's up to fp32 rounding (second half)
      fast_bias_grad  True (default; hwc layout): the backbone convolutions stay `F.conv2d` / cuDNN, but their
                    bias gradient (autograd's `grad.sum((0,2,3))`, a generic ATen reduction: 0.37 ms per 464^2
                    image) comes from the streaming column-sum kernel `wesup_colsum`
      footprints    True (default; fused paths only, training): the aggregated bilinear weights of
                    every superpixel (the sparse counterpart of the dense `sp_maps`) are built once
                    per forward on a side stream while the backbone runs, and the forward and
                    backward pooling kernels stream over those lists.  False, or under no_grad
                    (measured on tiled inference: 242 vs 199 tiles/s): the kernels rebuild them
                    internally.
    """

    def __init__(self, n_classes=2, D=32, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        self.backbone = _vgg16_features(kwargs.get("pretrained", True), bool(kwargs.get("allow_random_init", False)))
        self.fm_channels_sum = 0
        self._side_names = []
        for layer in self.backbone:
            if isinstance(layer, nn.Conv2d):
                name = f"side_conv{self.fm_channels_sum}"
                setattr(self, name, nn.Conv2d(layer.out_channels, layer.out_channels // 2, 1))
                self._side_names.append(name)
                self.fm_channels_sum += layer.out_channels // 2
        self.fc_layers = nn.Sequential(
            nn.Linear(self.fm_channels_sum, 1024), nn.ReLU(),
            nn.Linear(1024, 1024), nn.ReLU(),
            nn.Linear(1024, D), nn.ReLU())
        self.classifier = nn.Sequential(
            nn.Linear(D, self.kwargs.get("n_classes", n_classes)), nn.Softmax(dim=1))
        self.hc_dtype = kwargs.get("hc_dtype", torch.float32)
        self.hc_layout = kwargs.get("hc_layout", "hwc")
        self.fused_backward = bool(kwargs.get("fused_backward", True))
        self.materialize_hypercolumn = bool(kwargs.get("materialize_hypercolumn", False))
        self.pool_first = bool(kwargs.get("pool_first", True))
        self.use_footprints = bool(kwargs.get("footprints", True))
        self.fast_bias_grad = bool(kwargs.get("fast_bias_grad", True))
        self._fp_stream = None
        if self.hc_layout == "hwc":
            # the convolutions run channels_last: keep their weights (and therefore weight gradients
            # and momentum buffers) in that memory format too, so no per-iteration layout copies appear
            # in backward.  Shapes, values and state_dict keys are unaffected.
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
        self.feature_maps = None
        self.fm_size = None
        self.sp_features = None
        self.sp_pred = None

    # -- backbone + 1x1 side convs (cuDNN; the reported baseline) --------------
    def _conv(self, layer, x):
        if self.fast_bias_grad and self.hc_layout == "hwc" and torch.is_grad_enabled():
            return ops.conv2d_channels_last(x, layer)
        return layer(x)

    def _side_outputs(self, x):
        """Side conv on every PRE-ReLU conv output (the reference hooks the Conv2d
        modules, :205-210,253).  Runs channels_last so the side outputs are already
        pixel-major in memory for the hypercolumn kernel."""
        if self.hc_layout == "hwc":
            x = x.contiguous(memory_format=torch.channels_last)
        sides, names = [], iter(self._side_names)
        for layer in self.backbone:
            if isinstance(layer, nn.Conv2d):
                x = self._conv(layer, x)
                sides.append(getattr(self, next(names))(x))
            elif isinstance(layer, nn.ReLU):
                x = F.relu(x)            # out of place: the side conv saved the pre-ReLU tensor
            else:
                x = layer(x)
        return sides

    def _level_sizes(self, height, width):
        """(h, w) of every backbone conv output for an input of the given size (module arithmetic
        only: known before the backbone runs)."""
        def out(n, m):
            k, st, p, d = (v if isinstance(v, int) else v[0] for v in (m.kernel_size, m.stride, m.padding, m.dilation))
            return (n + 2 * p - d * (k - 1) - 1) // st + 1
        sizes = []
        for layer in self.backbone:
            if isinstance(layer, nn.Conv2d):
                height, width = out(height, layer), out(width, layer)
                sizes.append((height, width))
            elif isinstance(layer, nn.MaxPool2d):
                height, width = out(height, layer), out(width, layer)
        return sizes

    def _start_footprints(self, x, sp):
        """Fork the footprint build onto a side stream: it depends on the label map only and
        overlaps the backbone; `ops.hypercolumn_pool` joins it."""
        if not self.use_footprints or not torch.is_grad_enabled():
            return None                                   # forward-only callers (tiled inference): the in-kernel path, no build, no fork
        if self._fp_stream is None or self._fp_stream.device != x.device:
            self._fp_stream = torch.cuda.Stream(device=x.device)
        return ops.build_footprints(sp, self._level_sizes(x.size(2), x.size(3)), with_bwd=True, stream=self._fp_stream)

    def _pooled_first(self, x, sp):
        """Superpixel means of the PRE-ReLU backbone conv outputs (one fused kernel over the 13
        levels), then every side conv (reference :253, a 1x1 convolution + bias) as a small GEMM
        on the pooled rows: mean_S(W f + b) == W mean_S(f) + b."""
        fp = self._start_footprints(x, sp)
        x = x.contiguous(memory_format=torch.channels_last)
        outs = []
        for layer in self.backbone:
            if isinstance(layer, nn.Conv2d):
                x = self._conv(layer, x)
                outs.append(x)
            elif isinstance(layer, nn.ReLU):
                x = F.relu(x)            # out of place: the pooling kernel reads the pre-ReLU tensor afterwards
            else:
                x = layer(x)
        pooled, _ = ops.hypercolumn_pool(outs, self.fm_size, sp, materialize=False, footprints=fp)
        cols = []
        for name, part in zip(self._side_names, pooled.split([o.size(1) for o in outs], dim=1)):
            conv = getattr(self, name)
            cols.append(F.linear(part, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias))
        return torch.cat(cols, dim=1)

    def _hypercolumn(self, x):
        self.fm_size = (x.size(2), x.size(3))
        feats = ops.hypercolumn(self._side_outputs(x), self.fm_size, dtype=self.hc_dtype, layout=self.hc_layout)
        # same attribute as the reference, exposed in its (C,H,W) shape without a copy
        self.feature_maps = feats if self.hc_layout == "chw" else feats.t().view(-1, *self.fm_size)
        return feats

    def forward(self, x):
        """x = (image (1,3,H,W), sp_maps); returns class-1 probability (1,H,W)."""
        x, sp_maps = x
        sp = sp_maps if isinstance(sp_maps, SuperpixelMaps) else SuperpixelMaps.from_dense(sp_maps)
        if self.fused_backward and self.hc_layout == "hwc" and self.pool_first and not self.materialize_hypercolumn:
            self.fm_size = (x.size(2), x.size(3))
            self.feature_maps = None
            pooled = self._pooled_first(x, sp)
        elif self.fused_backward and self.hc_layout == "hwc":
            self.fm_size = (x.size(2), x.size(3))
            fp = None if self.materialize_hypercolumn else self._start_footprints(x, sp)
            pooled, feats = ops.hypercolumn_pool(self._side_outputs(x), self.fm_size, sp, dtype=self.hc_dtype,
                                                 materialize=self.materialize_hypercolumn, footprints=fp)
            self.feature_maps = None if feats is None else feats.t().view(-1, *self.fm_size)
        else:
            feats = self._hypercolumn(x)
            pooled = ops.sp_pool(feats, sp, layout=self.hc_layout)
        x = self.fc_layers(pooled)
        self.sp_features = x
        self.sp_pred = self.classifier(x)
        return ops.paint(sp, self.sp_pred, cls=1)


class WESUPPixelInference(WESUP):
    """Pixel-wise inference (reference: models/wesup.py:307-400): the hypercolumn
    goes straight through the MLP, one row per pixel.  Loads the same
    state_dict as `WESUP`.  The pixel-major hypercolumn is already the
    (H*W, 2112) operand the first Linear wants, so the reference's `x.t()`
    (:398) disappears."""

    def __init__(self, n_classes=2, D=32, **kwargs):
        kwargs = {**kwargs, "hc_layout": "hwc"}
        super().__init__(n_classes=n_classes, D=D, **kwargs)

    def forward(self, x):
        """x (1,3,H,W) -> (H,W,C) class probabilities (reference :382-400)."""
        return self.forward_batch(x)[0]

    def _backbone_levels(self, x):
        """The 13 PRE-ReLU backbone conv outputs (channels_last), what the reference's hooks see (:205-210)."""
        x = x.contiguous(memory_format=torch.channels_last)
        outs = []
        for layer in self.backbone:
            if isinstance(layer, nn.Conv2d):
                x = layer(x)
                outs.append(x)
            elif isinstance(layer, nn.ReLU):
                x = F.relu(x)
            else:
                x = layer(x)
        return outs

    def _folded_first_layer(self, outs, dtype):
        """Per distinct level resolution g: W'_g = cat_{l in g}(W1[:, slice_l] @ Wside_l)  (1024, sum of the backbone
        channels of the resolution), and the total bias b' = b1 + sum_l W1[:, slice_l] @ bside_l.  Recomputed from the
        current parameters on every call (13 small GEMMs, < 1 GFLOP), so it can never go stale."""
        w1, b1 = self.fc_layers[0].weight, self.fc_layers[0].bias
        groups, bias, off = [], b1.float(), 0
        for name, o in zip(self._side_names, outs):
            conv = getattr(self, name)
            half = conv.out_channels
            w1_l = w1[:, off:off + half].float()
            folded = w1_l @ conv.weight.view(half, conv.in_channels).float()          # (1024, C_l)
            bias = bias + w1_l @ conv.bias.float()
            size = (o.size(2), o.size(3))
            if groups and groups[-1][0] == size:
                groups[-1][1].append(folded)
                groups[-1][2].append(o)
            else:
                groups.append((size, [folded], [o]))
            off += half
        return [(size, torch.cat(ws, dim=1).to(dtype), lv) for size, ws, lv in groups], bias

    def forward_batch(self, x):
        """`forward` for a batch of same-sized tiles `(B,3,H,W)` -> `(B,H,W,C)`.  The first MLP layer is taken through
        the linear chain side conv -> upsample -> concat -> Linear(2112,1024) analytically (csrc/upsample_sum.cu): one
        library GEMM per level resolution on (B*h*w, C) rows, then ONE kernel per tile that upsamples, sums, adds the
        bias and applies the ReLU -- the (H*W,2112) hypercolumn is never formed and the layer costs 88 GFLOP per
        400-px tile instead of 692.  `project_first=False` keeps the reference's order of operations
        (hypercolumn kernel, then the 2112-wide GEMM)."""
        b, _, height, width = x.shape
        self.fm_size = (height, width)
        hw = height * width
        if not bool(self.kwargs.get("project_first", True)):
            sides = self._side_outputs(x)
            feats = torch.empty((b * hw, self.fm_channels_sum), dtype=self.hc_dtype, device=x.device)
            for t in range(b):
                ops.hypercolumn_into([s[t:t + 1] for s in sides], self.fm_size, feats[t * hw:(t + 1) * hw])
            self.feature_maps = feats.t().view(-1, height, width) if b == 1 else None
            h1, rest = feats, self.fc_layers
        else:
            self.feature_maps = None
            outs = self._backbone_levels(x)
            dtype = self.hc_dtype
            groups, bias = self._folded_first_layer(outs, dtype)
            terms = []
            for (gh, gw), w_g, levels in groups:
                rows = torch.empty((b * gh * gw, w_g.size(1)), dtype=dtype, device=x.device)
                off = 0
                for o in levels:                                        # (B,C,h,w) channels_last == (B*h*w, C) rows: one converting copy
                    rows[:, off:off + o.size(1)].copy_(o.permute(0, 2, 3, 1).reshape(-1, o.size(1)))
                    off += o.size(1)
                terms.append(F.linear(rows, w_g).view(b, gh, gw, -1))   # Z_g at the level's own resolution
            h1 = torch.empty((b * hw, terms[0].size(-1)), dtype=dtype, device=x.device)
            for t in range(b):
                ops.upsample_sum([z[t] for z in terms], self.fm_size, bias=bias, relu=True, out=h1[t * hw:(t + 1) * hw])
            rest = self.fc_layers[2:]
        if h1.dtype == torch.bfloat16:
            # bf16 tensor-core GEMMs with the bias + ReLU in the GEMM epilogue (cuBLASLt), softmax in fp32
            for layer in rest:
                if isinstance(layer, nn.Linear):
                    h1 = torch._addmm_activation(layer.bias.to(torch.bfloat16), h1, layer.weight.to(torch.bfloat16).t())
            logits = F.linear(h1, self.classifier[0].weight.to(torch.bfloat16), self.classifier[0].bias.to(torch.bfloat16))
            out = torch.softmax(logits.float(), dim=1)
        else:
            out = self.classifier(rest(h1))
        return out.view(b, height, width, -1)


class WESUPTrainer(BaseTrainer):
    """Reference: models/wesup.py:403-547."""

    def __init__(self, model, **kwargs):
        config = WESUPConfig()
        if config.freeze_backbone:
            for param in model.backbone.parameters():
                param.requires_grad = False
        kwargs = {**config.to_dict(), **kwargs}
        super().__init__(model, **kwargs)
        self.xentropy = partial(_cross_entropy)
        if kwargs.get("cudnn_benchmark", False):
            # optional: let cuDNN time its convolution algorithms per shape (the eager iterations that precede a
            # graph capture fill its cache); same arithmetic class, the reference leaves the torch default (off)
            torch.backends.cudnn.benchmark = True

    def get_default_dataset(self, root_dir, train=True, proportion=1.0):
        # PNG/CSV readers and albumentations augmentation are CPU I/O outside this
        # path: the reference's utils.data is used when
        # it is on sys.path, else the plain readers of wesup_b200.utils.data.
        try:
            from utils.data import Digest2019PointDataset, SegmentationDataset
        except ImportError:
            from ..utils.data import PointDataset as Digest2019PointDataset, SegmentationDataset
        if train:
            if osp.exists(osp.join(root_dir, "points")):
                return Digest2019PointDataset(root_dir, proportion=proportion,
                                              multiscale_range=self.kwargs.get("multiscale_range"))
            return SegmentationDataset(root_dir, proportion=proportion,
                                       multiscale_range=self.kwargs.get("multiscale_range"))
        return SegmentationDataset(root_dir, rescale_factor=self.kwargs.get("rescale_factor"), train=False)

    def get_default_optimizer(self):
        params = [p for p in self.model.parameters() if p.requires_grad]
        # same update rule as the reference's SGD; torch's single-kernel ("fused") implementation when
        # the parameters live on the GPU (kwarg fused_optimizer=False restores the default one)
        fused = bool(self.kwargs.get("fused_optimizer", True)) and all(p.is_cuda for p in params)
        optimizer = torch.optim.SGD(params, lr=5e-5, momentum=self.kwargs.get("momentum"),
                                    weight_decay=self.kwargs.get("weight_decay"), **({"fused": True} if fused else {}))
        return optimizer, None      # the reference builds a scheduler and discards it (:452-455)

    def segment(self, img, sync=True):
        """GPU SLIC with the reference's parameters (:471-476).  Returns the int32
        label map and the number of labels -- as a host int (one sync, like the
        reference's `.cpu()` round trip but without moving the image) or, with
        `sync=False`, as `(device scalar, host upper bound)`."""
        n_segments = int(img.size(-2) * img.size(-1) / self.kwargs.get("sp_area"))
        labels, n_labels = ops.slic(img, n_segments=n_segments, compactness=self.kwargs.get("sp_compactness"))
        if sync:
            return labels, int(n_labels.item())
        # every kept component has >= min_size = int(0.5 * HW / n_segments) pixels (skimage's rule)
        hw = img.size(-2) * img.size(-1)
        bound = hw // max(int(0.5 * hw / n_segments), 1) + 1
        return labels, (n_labels, bound)

    def _enqueue_preprocess(self, *data):
        """Everything `preprocess` does up to (not including) its single host wait."""
        data = [datum.to(self.device, non_blocking=True) for datum in data]
        if len(data) == 3:
            img, pixel_mask, point_mask = data
        elif len(data) == 2:
            img, pixel_mask = data
            point_mask = empty_tensor()
        elif len(data) == 1:
            img, = data
            point_mask = empty_tensor()
            pixel_mask = empty_tensor()
        else:
            raise ValueError("Invalid input data for WESUP")
        segments, (n_dev, bound) = self.segment(img, sync=False)
        if point_mask is not None and not is_empty_tensor(point_mask):
            mask = point_mask.squeeze(0) if point_mask.dim() == 4 else point_mask.squeeze()
        elif pixel_mask is not None and not is_empty_tensor(pixel_mask):
            mask = pixel_mask.squeeze(0) if pixel_mask.dim() == 4 else pixel_mask.squeeze()
        else:
            mask = None
        # SLIC -> statistics without a host round trip in between: ONE host wait per image
        pending = SuperpixelMaps.from_labels(segments, None if mask is None else mask, n_sp=bound, n_sp_dev=n_dev, defer=True)
        return img, pixel_mask, mask is not None, pending

    @staticmethod
    def _finish_preprocess(img, pixel_mask, has_mask, pending):
        sp_maps = pending.finish()
        sp_labels = sp_maps.sp_labels if has_mask else empty_tensor().to(img.device)
        return (img, sp_maps), (pixel_mask, sp_labels)

    def preprocess(self, *data):
        staged_map = getattr(self, "_prefetched", None)
        hit = staged_map.pop(tuple(id(t) for t in data), None) if staged_map else None
        if hit is not None and all(a is b for a, b in zip(hit[0], data)):
            _, staged, done = hit
            torch.cuda.current_stream(self.device).wait_event(done)
            return self._finish_preprocess(*staged)
        return self._finish_preprocess(*self._enqueue_preprocess(*data))

    def prefetch(self, *data):
        """Run `preprocess(*data)` (H2D copy, GPU SLIC, superpixel statistics) on a side
        stream, ahead of the training stream.  A later `preprocess` call with these same
        tensor objects picks the result up; its host scalars are usually already there."""
        if not torch.cuda.is_available():
            return
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
            self._prefetched = {}
        key = tuple(id(t) for t in data)
        if key in self._prefetched:
            return
        while len(self._prefetched) >= 4:                  # bounded look-ahead
            self._prefetched.pop(next(iter(self._prefetched)))
        main = torch.cuda.current_stream(self.device)
        side = self._side_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            staged = self._enqueue_preprocess(*data)
            done = torch.cuda.Event()
            done.record(side)
        img, pixel_mask, _, pending = staged
        for t in [img, pixel_mask] + pending.tensors():      # allocated on `side`, consumed on `main`
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(main)
        self._prefetched[key] = (tuple(data), staged, done)

    # ---- CUDA-graph iteration ---------------------------------------------
    # The eager iteration is host-bound on a B200 (~380 launches per image).  With `cuda_graph=True`
    # everything after preprocessing -- VGG16, superpixel stage, loss, backward, SGD step, metrics --
    # is captured ONCE per (input shape, superpixel capacity) and replayed.  Preprocessing (H2D copy,
    # GPU SLIC, superpixel statistics) keeps running one image ahead on the side stream, which is
    # also where the host learns the image's superpixel count N without stalling.  The capacity is
    # N rounded up to a multiple of max(64, 2^ceil(log2(N/32))) (64 at 464^2, 512 at CRAG size -- every capacity is a
    # graph of its own holding a full set of activations, 30 GB at CRAG size, so images of one dataset should land
    # on ONE capacity; r2 measured 135 GB with four): the graph's per-superpixel buffers have that many rows, the
    # rows beyond N are empty superpixels (zero features, zero labels, zero gradient), and label
    # propagation reads N and the labeled count from device memory.  Loss, metrics and updates equal
    # the eager iteration's up to fp32 summation order.  Hyper-parameters are baked into a graph;
    # it is re-captured when the learning rate changes.
    GRAPH_ROW_QUANTUM = 64

    def _static_iteration(self, st, step=True):
        sp = st["sp"]
        if self.grad_sync is not None:
            self.grad_sync.zero_grad()
        pred = self.model((st["img"], sp))
        sp_pred, sp_features, y_l = self.model.sp_pred, self.model.sp_features, sp.sp_labels_full
        counts_dev = st["counts_dev"]
        metrics = {"labeled_sp_ratio": counts_dev[1].float() / counts_dev[0].float()}
        loss = self.xentropy(sp_pred, y_l)                      # all-zero rows are ignored by the loss itself
        if self.kwargs.get("enable_propagation"):
            y_u = ops.label_propagate_static(sp_features, y_l, counts_dev, self.kwargs.get("propagate_threshold"))
            propagate_loss = self.xentropy(sp_pred, y_u)
            loss = loss + self.kwargs.get("propagate_weight") * propagate_loss
            metrics["propagated_labels"] = y_u.sum()
            metrics["propagate_loss"] = propagate_loss.detach()
        self.model.sp_pred = None
        metrics["loss"] = loss.detach()
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()               # the bucketed all-reduces started by the backward hooks (captured too)
        if step:
            self.arm_step_guard(loss)             # the update is skipped on the device when the loss is not finite
            self.optimizer.step()
        self._defer_scalars = True
        try:
            labels, target = self.postprocess(pred, (st["pixel_mask"], None))
            metrics.update(self.evaluate(labels, target))
        finally:
            self._defer_scalars = False
        keys = list(metrics)
        return keys, torch.stack([metrics[k].detach().float().reshape(()) for k in keys])

    @staticmethod
    def _load_static(st, img, pixel_mask, sp):
        """Copy one preprocessed image into a graph's input buffers (device-to-device, a few KB..MB)."""
        n, hw = sp.n, sp.height * sp.width
        st["img"].copy_(img, non_blocking=True)
        if st["pixel_mask"] is not None:
            st["pixel_mask"].copy_(pixel_mask, non_blocking=True)
        dst = st["sp"]
        dst.row_labels.copy_(sp.row_labels, non_blocking=True)
        dst.seg_pixels.copy_(sp.seg_pixels, non_blocking=True)
        dst.counts.zero_()
        dst.counts[:n].copy_(sp.counts, non_blocking=True)
        dst.seg_offsets.fill_(hw)
        dst.seg_offsets[:n + 1].copy_(sp.seg_offsets, non_blocking=True)
        if dst.sp_labels_full is not None:
            dst.sp_labels_full.zero_()
            dst.sp_labels_full[:n].copy_(sp.sp_labels_full, non_blocking=True)
        st["counts_dev"].copy_(sp.counts_dev, non_blocking=True)

    def _capture(self, img, pixel_mask, sp, cap, pool):
        dev, i32 = img.device, dict(dtype=torch.int32, device=img.device)
        hw = sp.height * sp.width
        static_sp = SuperpixelMaps(sp.height, sp.width, cap, None, torch.empty(hw, **i32), torch.empty(cap, **i32),
                                   torch.empty(cap + 1, **i32), torch.empty(hw, **i32),
                                   torch.empty(cap, sp.sp_labels_full.size(1), device=dev), None)
        st = {"img": torch.empty_like(img), "pixel_mask": torch.empty_like(pixel_mask), "sp": static_sp,
              "counts_dev": torch.empty(2, **i32)}
        self._load_static(st, img, pixel_mask, sp)
        self.flush_metrics()
        # autograd caches one AccumulateGrad node per parameter, bound to the stream of the forward that
        # created it, for as long as a graph that uses it is alive: drop the eager iteration's graph (the
        # module caches sp_features) so that warm-up and capture, both on `side`, create fresh ones
        self.model.sp_features = self.model.sp_pred = self.model.feature_maps = None
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        if self.grad_sync is not None:
            self.grad_sync.suspended = True                 # the warm-up run is local: ranks capture at different times
        try:
            with torch.cuda.stream(side):                   # warm-up of the static path: no parameter update, no collective
                self._static_iteration(st, step=False)
        finally:
            if self.grad_sync is not None:
                self.grad_sync.suspended = False
        torch.cuda.current_stream(dev).wait_stream(side)
        self.optimizer.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        lib = ops._lib.load()
        n0 = lib.wesup_kernel_launches()
        # with data parallelism the NCCL watchdog thread polls events while we capture: thread-local capture mode
        opts = {"capture_error_mode": "thread_local"} if self.grad_sync is not None else {}
        if pool is not None:
            opts["pool"] = pool
        with torch.cuda.graph(graph, stream=side, **opts):
            keys, out = self._static_iteration(st, step=True)
        # library launches recorded into the graph: every replay re-issues them without passing through the C ABI
        return {"graph": graph, "st": st, "keys": keys, "out": out, "lr": self.optimizer.param_groups[0]["lr"],
                "launches": int(lib.wesup_kernel_launches() - n0)}

    def train_one_iteration(self, phase, *data):
        use_graph = (phase == "train" and self.kwargs.get("cuda_graph", False) and len(data) in (2, 3)
                     and torch.cuda.is_available() and self.optimizer is not None
                     and all(torch.is_tensor(d) and not is_empty_tensor(d) for d in data))
        if not use_graph:
            return super().train_one_iteration(phase, *data)
        if not hasattr(self, "_graphs"):
            self._graphs, self._graph_seen, self._graph_failed, self._graph_pool = {}, {}, set(), None
        input_, target = self.preprocess(*data)
        (img, sp), (pixel_mask, _) = input_, target
        shape_key = tuple((tuple(d.shape), d.dtype) for d in data)
        q = self.GRAPH_ROW_QUANTUM
        while q * 32 < sp.n:
            q *= 2
        cap = -(-sp.n // q) * q
        key = (shape_key, cap)
        seen = self._graph_seen.get(shape_key, 0)
        self._graph_seen[shape_key] = seen + 1
        entry = self._graphs.get(key)
        if entry is not None and entry["lr"] != self.optimizer.param_groups[0]["lr"]:
            entry = None                                    # baked hyper-parameter changed: capture again
        # the first iterations of a shape run eagerly (cuDNN/cuBLAS set-up, momentum buffers)
        eager = (key in self._graph_failed or sp.counts_dev is None or sp.sp_labels_full is None
                 or seen < int(self.kwargs.get("cuda_graph_after", 2)) or len(self.optimizer.state) == 0)
        if not eager and entry is None:
            try:
                entry = self._graphs[key] = self._capture(img, pixel_mask, sp, cap, self._graph_pool)
                if self._graph_pool is None:
                    self._graph_pool = entry["graph"].pool()      # later graphs replay one at a time: share the memory
            except Exception as ex:  # noqa: BLE001  (capture is an optimisation: never lose the iteration to it)
                warnings.warn(f"CUDA-graph capture of the training iteration failed ({type(ex).__name__}: {ex}); "
                              "this shape keeps running eagerly")
                self._graph_failed.add(key)
                self._graphs.pop(key, None)
                self.model.sp_features = self.model.sp_pred = None
                self.optimizer.zero_grad(set_to_none=True)
                torch.cuda.synchronize()
                eager = True
        if eager:
            return self._run_iteration(phase, input_, target)
        self._load_static(entry["st"], img, pixel_mask, sp)
        entry["graph"].replay()
        self.replayed_launches = getattr(self, "replayed_launches", 0) + entry["launches"]
        self._submit_scalars(dict(zip(entry["keys"], entry["out"].unbind(0))), phase)
        self.flush_metrics(keep=max(int(self.kwargs.get("metrics_lag", 1) or 0), 0))

    # ---- inference on one image / tile ------------------------------------------------------
    def predict_labels(self, img):
        """Class map (H,W) uint8 of one image `(1,3,H,W)` -- what infer.py / infer_tile.py compute as
        `postprocess(model(preprocess(img)))` (/root/reference/infer_tile.py:111-116).  Picks up a
        `prefetch(img)` issued earlier; with `cuda_graph=True` the network part (VGG16 -> superpixel
        means -> MLP -> paint -> round) replays one graph per (tile shape, 64-row superpixel capacity)."""
        with torch.no_grad():
            (x, sp), _ = self.preprocess(img)
            if not (self.kwargs.get("cuda_graph", False) and sp.counts_dev is not None):
                return self.postprocess(self.model((x, sp)))[0].to(torch.uint8)
            if not hasattr(self, "_infer_graphs"):
                self._infer_graphs, self._infer_seen, self._infer_pool = {}, {}, None
            q = self.GRAPH_ROW_QUANTUM
            cap = -(-sp.n // q) * q
            key = (tuple(x.shape), cap, self.model.training)
            entry = self._infer_graphs.get(key)
            if entry is None:
                seen = self._infer_seen.get(key[0], 0)
                self._infer_seen[key[0]] = seen + 1
                if seen < int(self.kwargs.get("cuda_graph_after", 2)):
                    return self.postprocess(self.model((x, sp)))[0].to(torch.uint8)
                entry = self._infer_graphs[key] = self._capture_infer(x, sp, cap)
            self._load_static(entry["st"], x, None, sp)
            entry["graph"].replay()
            return entry["out"].clone()

    def _capture_infer(self, x, sp, cap):
        dev, i32 = x.device, dict(dtype=torch.int32, device=x.device)
        hw = sp.height * sp.width
        static_sp = SuperpixelMaps(sp.height, sp.width, cap, None, torch.empty(hw, **i32), torch.empty(cap, **i32),
                                   torch.empty(cap + 1, **i32), torch.empty(hw, **i32), None, None)
        st = {"img": torch.empty_like(x), "pixel_mask": None, "sp": static_sp, "counts_dev": torch.empty(2, **i32)}
        self._load_static(st, x, None, sp)
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.postprocess(self.model((st["img"], static_sp)))
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side, **({"pool": self._infer_pool} if self._infer_pool is not None else {})):
            out = self.postprocess(self.model((st["img"], static_sp)))[0].to(torch.uint8)
        if self._infer_pool is None:
            self._infer_pool = graph.pool()
        return {"graph": graph, "st": st, "out": out}

    def compute_loss(self, pred, target, metrics=None):
        _, sp_labels = target
        sp_features = self.model.sp_features
        sp_pred = self.model.sp_pred
        if sp_pred is None:
            raise RuntimeError("You must run a forward pass before computing loss.")
        total_num = sp_pred.size(0)
        labeled_num = sp_labels.size(0)
        if labeled_num < total_num:          # weakly-supervised mode
            loss = self.xentropy(sp_pred[:labeled_num], sp_labels)
            if self.kwargs.get("enable_propagation"):
                propagated_labels = _label_propagate(sp_features, sp_labels,
                                                     threshold=self.kwargs.get("propagate_threshold"))
                propagate_loss = self.xentropy(sp_pred[labeled_num:], propagated_labels)
                loss = loss + self.kwargs.get("propagate_weight") * propagate_loss
            if metrics is not None and isinstance(metrics, dict):
                metrics["labeled_sp_ratio"] = labeled_num / total_num
                if self.kwargs.get("enable_propagation"):
                    # inside train_one_iteration the scalars stay on the device and are read
                    # in one transfer at the end of the iteration; direct callers get floats
                    lazy = getattr(self, "_defer_scalars", False)
                    metrics["propagated_labels"] = propagated_labels.sum() if lazy else propagated_labels.sum().item()
                    metrics["propagate_loss"] = propagate_loss.detach() if lazy else propagate_loss.item()
        else:                                # fully-supervised mode
            loss = self.xentropy(sp_pred, sp_labels)
        self.model.sp_pred = None            # clear outdated prediction (:529)
        return loss

    def postprocess(self, pred, target=None):
        pred = pred.round().long()
        if target is not None:
            return pred, target[0].argmax(dim=1)
        return pred

    def post_epoch_hook(self, epoch):
        if self.scheduler is not None:
            labeled_loss = np.mean(self.tracker.history["loss"])
            if "propagate_loss" in self.tracker.history:
                labeled_loss -= np.mean(self.tracker.history["propagate_loss"])
            self.scheduler.step(labeled_loss)
