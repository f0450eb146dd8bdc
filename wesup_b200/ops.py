"""Torch-facing operators of the superpixel stage.

Each function/autograd.Function here is a thin shim over one C-ABI entry point
of libwesup_b200.so (include/wesup_b200.h): it allocates outputs and workspaces
with torch's caching allocator, passes raw device pointers plus the current
CUDA stream, and turns non-zero return codes into RuntimeError.  No math
happens in Python and there is no fallback path: CPU tensors are rejected.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import BF16, CHW, F32, HWC, check

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16}
_LAYOUTS = {"chw": CHW, "hwc": HWC}


def _stream(t: torch.Tensor) -> int:
    """The current stream of the device `t` lives on (not of the current device)."""
    return torch.cuda.current_stream(t.device).cuda_stream


def _call(name: str, t: torch.Tensor, *args) -> None:
    """One C-ABI call on the device that owns `t`: that device is made current for the call (the
    library launches on the current device and queries its occupancy) and the call goes onto
    that device's current stream -- a tensor on cuda:1 never meets cuda:0's stream."""
    lib = _lib.load()
    with torch.cuda.device(t.device):
        check(getattr(lib, name)(*args, _stream(t)), name)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"wesup_b200: `{name}` must be a CUDA tensor (the superpixel stage has no CPU path)")


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------
# compact superpixel representation (replaces the dense (N,H,W) sp_maps)
# ---------------------------------------------------------------------------
class SuperpixelMaps:
    """What `preprocess` hands to `WESUP.forward` in place of the reference's dense
    `sp_maps (N,H,W)` (/root/reference/models/wesup.py:57-61): the relabelled int32
    label map plus per-row counts and a CSR of pixel ids.  Row k of every
    downstream tensor (pooled features, sp_features, sp_pred, y_u) refers to
    original superpixel `order[k]`, exactly as row k of the dense maps did.

    Quacks enough like the dense tensor for the reference's call sites:
    `.size()`, `.shape`, `.dim()`, `.device`, `.to()`; `.to_dense()` rebuilds the
    reference tensor (tests / legacy callers only)."""

    def __init__(self, height, width, n, order, row_labels, counts, seg_offsets, seg_pixels,
                 sp_labels=None, n_labeled_dev=None):
        self.height, self.width, self.n = int(height), int(width), int(n)
        self.order, self.row_labels, self.counts = order, row_labels, counts
        self.seg_offsets, self.seg_pixels = seg_offsets, seg_pixels
        self.sp_labels_full = sp_labels            # (n, n_cls) incl. zero rows, or None
        self.counts_dev = None                     # int32 {n, n_labeled} on the device when known there
        self._n_labeled_dev = n_labeled_dev
        self._n_labeled: Optional[int] = None

    # -- tensor-like surface -------------------------------------------------
    def size(self, dim: Optional[int] = None):
        s = torch.Size((self.n, self.height, self.width))
        return s if dim is None else s[dim]

    @property
    def shape(self):
        return self.size()

    def dim(self) -> int:
        return 3

    @property
    def device(self):
        return self.row_labels.device

    def to(self, *args, **kwargs):
        dev = torch.device(args[0]) if args and not isinstance(args[0], torch.dtype) else kwargs.get("device")
        if dev is None or dev == self.device:
            return self
        raise RuntimeError("SuperpixelMaps lives on the GPU that built it")

    @property
    def n_labeled(self) -> int:
        """Host copy of the labeled-row count (one D2H sync, cached)."""
        if self._n_labeled is None:
            self._n_labeled = 0 if self._n_labeled_dev is None else int(self._n_labeled_dev.item())
        return self._n_labeled

    @property
    def sp_labels(self) -> Optional[torch.Tensor]:
        if self.sp_labels_full is None:
            return None
        return self.sp_labels_full[: self.n_labeled]

    @property
    def label_map(self) -> torch.Tensor:
        return self.row_labels.view(self.height, self.width)

    def to_dense(self) -> torch.Tensor:
        rows = torch.arange(self.n, device=self.device, dtype=torch.int32).view(-1, 1, 1)
        maps = (self.label_map.unsqueeze(0) == rows).float()
        return maps / maps.sum(dim=(1, 2), keepdim=True)

    # -- constructors ----------------------------------------------------------
    @staticmethod
    def from_labels(labels: torch.Tensor, mask: Optional[torch.Tensor] = None, n_sp: Optional[int] = None,
                    n_sp_dev: Optional[torch.Tensor] = None, defer: bool = False) -> "SuperpixelMaps":
        """labels: (H,W) integer ids in [0,n_sp); mask: (C,H,W) int64 one-hot-or-zero or None.

        `n_sp_dev` (a 1-element device tensor holding the true number of ids, e.g. the
        `n_labels` output of `ops.slic`) turns `n_sp` into an UPPER BOUND: the statistics
        run over the bound (absent ids are empty superpixels that sort to the very end of
        the row order), then the true count and the labeled count come back in ONE
        device-to-host read and every per-row array is trimmed to the true count.
        With `defer=True` that read is only *started* (async copy into pinned memory) and a
        `PendingSuperpixelMaps` is returned; `.finish()` completes it later, so a caller can
        run this on a side stream one image ahead without ever stalling the main stream."""
        _require_cuda(labels, "segments")
        if labels.dim() != 2:
            raise ValueError("segments must be (H, W)")
        h, w = labels.shape
        lab32 = labels.to(torch.int32).contiguous()
        if n_sp is None:
            n_sp = int(lab32.max().item()) + 1          # the reference does the same sync (models/wesup.py:41)
        dev = labels.device
        n_cls = 0
        mask64 = None
        if mask is not None and mask.dim() != 0:
            if mask.dim() != 3 or mask.shape[1:] != labels.shape:
                raise ValueError("mask must be (C, H, W) matching segments")
            mask64 = mask.to(device=dev, dtype=torch.int64).contiguous()
            n_cls = mask64.size(0)
        i32 = dict(dtype=torch.int32, device=dev)
        # one allocation for the three per-row int arrays, one for the two per-pixel ones
        rows = torch.empty(3 * n_sp + 1, **i32)
        order, counts, seg_offsets = rows[:n_sp], rows[n_sp:2 * n_sp], rows[2 * n_sp:]
        px = torch.empty(2 * h * w, **i32)
        row_labels, seg_pixels = px[:h * w], px[h * w:]
        n_labeled = torch.zeros(1, **i32)
        sp_labels = torch.empty(n_sp, n_cls, dtype=torch.float32, device=dev) if n_cls else None
        lib = _lib.load()
        ws = _ws(lib.wesup_sp_stats_workspace_bytes(h, w, n_sp, n_cls), dev)
        _call("wesup_sp_stats", lab32, lab32.data_ptr(), mask64.data_ptr() if mask64 is not None else None, h, w, n_cls, n_sp,
                                 order.data_ptr(), row_labels.data_ptr(), counts.data_ptr(), seg_offsets.data_ptr(),
                                 seg_pixels.data_ptr(), sp_labels.data_ptr() if sp_labels is not None else None,
                                 n_labeled.data_ptr(), ws.data_ptr())
        sp = SuperpixelMaps(h, w, n_sp, order, row_labels, counts, seg_offsets, seg_pixels, sp_labels,
                            n_labeled if n_cls else None)
        if n_sp_dev is None:
            return sp
        pending = PendingSuperpixelMaps(sp, torch.cat([n_sp_dev.reshape(1).to(torch.int32), n_labeled]), bool(n_cls))
        return pending if defer else pending.finish()

    @staticmethod
    def from_dense(sp_maps: torch.Tensor) -> "SuperpixelMaps":
        """Legacy callers that still pass the dense (N,H,W) tensor: the owner of a
        pixel is argmax over rows, as at /root/reference/models/wesup.py:295."""
        _require_cuda(sp_maps, "sp_maps")
        owner = sp_maps.argmax(dim=0)
        return SuperpixelMaps.from_labels(owner, None, n_sp=sp_maps.size(0))


class PendingSuperpixelMaps:
    """`SuperpixelMaps` built over an upper bound of the superpixel count whose two
    host scalars (true count, labeled count) are still in flight."""

    def __init__(self, sp: SuperpixelMaps, scalars_dev: torch.Tensor, has_labels: bool):
        self.sp, self.has_labels, self.scalars_dev = sp, has_labels, scalars_dev
        self.host = torch.empty(2, dtype=torch.int32, pin_memory=True)
        self.host.copy_(scalars_dev, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record(torch.cuda.current_stream(scalars_dev.device))

    def tensors(self):
        sp = self.sp
        return [t for t in (sp.order, sp.row_labels, sp.counts, sp.seg_offsets, sp.seg_pixels, sp.sp_labels_full,
                            sp._n_labeled_dev, self.scalars_dev) if t is not None]

    def finish(self) -> SuperpixelMaps:
        self.event.synchronize()                               # the one host wait per image
        n_true, n_labeled = self.host.tolist()
        sp = self.sp
        if n_true > sp.n:
            raise RuntimeError(f"superpixel count {n_true} exceeds the bound {sp.n} the statistics ran with")
        sp.order, sp.counts, sp.seg_offsets = sp.order[:n_true], sp.counts[:n_true], sp.seg_offsets[:n_true + 1]
        if sp.sp_labels_full is not None:
            sp.sp_labels_full = sp.sp_labels_full[:n_true]
        sp.n = n_true
        sp._n_labeled = int(n_labeled) if self.has_labels else 0
        sp.counts_dev = self.scalars_dev            # int32 {n_true, n_labeled} on the device
        return sp


# ---------------------------------------------------------------------------
# precomputed pooling footprints (the sparse counterpart of the dense sp_maps)
# ---------------------------------------------------------------------------
class Footprints:
    """Device blob written by `wesup_footprint_build` for one label map and one list of
    level sizes: per superpixel the low-resolution cells it touches with their aggregated
    bilinear weights (forward lists) and, `with_bwd`, the transpose per cell.  It plays the
    role of the reference's `sp_maps` (/root/reference/models/wesup.py:57-61) for the fused
    upsample + pooling operator and, like it, depends on the image only -- never on the
    network weights -- so it is built once per image, off the critical path."""

    def __init__(self, blob, hs, ws, height, width, n, with_bwd):
        self.blob, self.hs, self.ws = blob, tuple(hs), tuple(ws)
        self.height, self.width, self.n, self.with_bwd = int(height), int(width), int(n), bool(with_bwd)
        self._pending: Optional[torch.cuda.Stream] = None

    def matches(self, hs, ws, height, width, n) -> bool:
        return (self.hs, self.ws, self.height, self.width, self.n) == (tuple(hs), tuple(ws), int(height), int(width), int(n))

    def join(self) -> "Footprints":
        """Make the current stream wait for a build that was forked onto a side stream."""
        if self._pending is not None:
            torch.cuda.current_stream(self.blob.device).wait_stream(self._pending)
            self._pending = None
        return self


def build_footprints(sp: SuperpixelMaps, level_sizes: Sequence[Tuple[int, int]], with_bwd: bool = True,
                     stream: Optional[torch.cuda.Stream] = None) -> Footprints:
    """Build the pooling footprints of `sp` for feature levels of the given (h, w) sizes.
    With `stream`, the build is forked from the current stream onto it (it then overlaps
    whatever the caller enqueues next, e.g. the backbone) and `Footprints.join()` -- called by
    `hypercolumn_pool` -- brings it back; the fork/join pair also captures into a CUDA graph."""
    lib = _lib.load()
    _require_cuda(sp.row_labels, "sp_maps")
    hs, ws = [int(h) for h, _ in level_sizes], [int(w) for _, w in level_sizes]
    ha, wa = _lib.int_array(hs), _lib.int_array(ws)
    nbytes = lib.wesup_footprint_bytes(ha, wa, len(hs), sp.height, sp.width, sp.n)
    if nbytes == 0:
        raise ValueError(f"cannot plan pooling footprints for level sizes {list(zip(hs, ws))} on a {sp.height}x{sp.width} map")
    dev = sp.row_labels.device
    fp = Footprints(_ws(nbytes, dev), hs, ws, sp.height, sp.width, sp.n, with_bwd)

    def launch():
        _call("wesup_footprint_build", sp.row_labels, ha, wa, len(hs), sp.height, sp.width, sp.n, sp.seg_offsets.data_ptr(),
                                        sp.seg_pixels.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(),
                                        int(bool(with_bwd)), fp.blob.data_ptr())

    if stream is None:
        launch()
        return fp
    stream.wait_stream(torch.cuda.current_stream(dev))
    if not torch.cuda.is_current_stream_capturing():
        fp.blob.record_stream(stream)
    with torch.cuda.stream(stream):
        launch()
    fp._pending = stream
    return fp


# ---------------------------------------------------------------------------
# (a) hypercolumn
# ---------------------------------------------------------------------------
def _as_hwc(t: torch.Tensor) -> torch.Tensor:
    """(1,C,h,w) -> memory (h,w,C); free for channels_last tensors."""
    return t.permute(0, 2, 3, 1).contiguous()


class _Hypercolumn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, size, dtype, layout, *sides):
        lib = _lib.load()
        H, W = size
        for s in sides:
            _require_cuda(s, "side output")
            if s.dim() != 4 or s.size(0) != 1 or s.dtype != torch.float32:
                raise ValueError("side outputs must be fp32 (1,C,h,w)")
        C = [s.size(1) for s in sides]
        h = [s.size(2) for s in sides]
        w = [s.size(3) for s in sides]
        mem = [_as_hwc(s) if layout == HWC else s.contiguous() for s in sides]
        ctot = sum(C)
        dev = sides[0].device
        out = torch.empty((H * W, ctot) if layout == HWC else (ctot, H, W), dtype=dtype, device=dev)
        _call("wesup_hypercolumn_fwd", out, _lib.ptr_array([m.data_ptr() for m in mem]), _lib.int_array(C), _lib.int_array(h),
                                        _lib.int_array(w), len(sides), H, W, out.data_ptr(), _DTYPES[dtype], layout)
        ctx.geom = (C, h, w, H, W, layout)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        C, h, w, H, W, layout = ctx.geom
        grad_out = grad_out.contiguous()
        dev = grad_out.device
        if layout == HWC:
            mem = [torch.empty((1, hh, ww, cc), dtype=torch.float32, device=dev) for cc, hh, ww in zip(C, h, w)]
        else:
            mem = [torch.empty((1, cc, hh, ww), dtype=torch.float32, device=dev) for cc, hh, ww in zip(C, h, w)]
        ca, ha, wa = _lib.int_array(C), _lib.int_array(h), _lib.int_array(w)
        ws = _ws(lib.wesup_hypercolumn_bwd_workspace_bytes(ca, ha, wa, len(C), H, W), dev) if layout == HWC else None
        _call("wesup_hypercolumn_bwd", grad_out, grad_out.data_ptr(), _DTYPES[grad_out.dtype], layout, ca, ha, wa, len(C), H, W,
                                        _lib.ptr_array([m.data_ptr() for m in mem]),
                                        ws.data_ptr() if ws is not None else None)
        grads = [m.permute(0, 3, 1, 2) if layout == HWC else m for m in mem]
        return (None, None, None, *grads)


def hypercolumn(sides: Sequence[torch.Tensor], size: Tuple[int, int], dtype=torch.float32, layout: str = "hwc") -> torch.Tensor:
    """Fused bilinear(align_corners=True) upsample + channel concat of the side
    outputs (/root/reference/models/wesup.py:254-261).  Returns (H*W, C) for
    layout 'hwc' or (C, H, W) for 'chw'."""
    return _Hypercolumn.apply((int(size[0]), int(size[1])), dtype, _LAYOUTS[layout], *sides)


# ---------------------------------------------------------------------------
# (b) pooling
# ---------------------------------------------------------------------------
class _SpPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, sp: SuperpixelMaps, layout):
        lib = _lib.load()
        _require_cuda(feat, "features")
        feat = feat.contiguous()
        hw = sp.height * sp.width
        if layout == HWC:
            if feat.dim() != 2 or feat.size(0) != hw:
                raise ValueError(f"features must be (H*W={hw}, C) for the hwc layout, got {tuple(feat.shape)}")
            c = feat.size(1)
        else:
            if feat.dim() != 3 or feat.size(1) * feat.size(2) != hw:
                raise ValueError(f"features must be (C, H, W) for the chw layout, got {tuple(feat.shape)}")
            c = feat.size(0)
        pooled = torch.empty((sp.n, c), dtype=torch.float32, device=feat.device)
        _call("wesup_sp_pool_fwd", pooled, feat.data_ptr(), _DTYPES[feat.dtype], layout, sp.seg_offsets.data_ptr(),
                                    sp.seg_pixels.data_ptr(), hw, c, sp.n, pooled.data_ptr())
        ctx.sp, ctx.layout, ctx.meta = sp, layout, (feat.shape, feat.dtype, c)
        return pooled

    @staticmethod
    def backward(ctx, grad_pooled):
        lib = _lib.load()
        sp = ctx.sp
        shape, dtype, c = ctx.meta
        grad_pooled = grad_pooled.contiguous().float()
        grad_feat = torch.empty(shape, dtype=dtype, device=grad_pooled.device)
        _call("wesup_sp_pool_bwd", grad_pooled, grad_pooled.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(),
                                    sp.height * sp.width, c, sp.n, grad_feat.data_ptr(), _DTYPES[dtype], ctx.layout)
        return grad_feat, None, None


def sp_pool(feat: torch.Tensor, sp: SuperpixelMaps, layout: str = "hwc") -> torch.Tensor:
    """Per-superpixel mean of the pixel features -> (N, C) fp32; replaces
    torch.mm(sp_maps, x.t()) (/root/reference/models/wesup.py:284-285)."""
    return _SpPool.apply(feat, sp, _LAYOUTS[layout])


class _HypercolumnPool(torch.autograd.Function):
    """(a) then (b) as one differentiable op over the side outputs (or the backbone levels).
    Forward: the two north-star kernels (the hypercolumn is written once, then pooled) when a
    `dtype` is given, else the fused operator that pools straight from the levels -- over
    precomputed `Footprints` when given, with the lists rebuilt in-kernel otherwise.
    Backward: always the fused adjoint of both from the pooled gradient (over the same
    footprints when they carry the backward lists), so the (H*W, C) gradient is never
    materialised and nothing of that size is kept for backward (both ops are linear)."""

    @staticmethod
    def forward(ctx, sp: SuperpixelMaps, size, dtype, fp, *sides):
        lib = _lib.load()
        H, W = size
        for s in sides:
            _require_cuda(s, "side output")
            if s.dim() != 4 or s.size(0) != 1 or s.dtype != torch.float32:
                raise ValueError("side outputs must be fp32 (1,C,h,w)")
        if sp.height != H or sp.width != W:
            raise ValueError(f"superpixel map is {sp.height}x{sp.width}, image is {H}x{W}")
        C = [s.size(1) for s in sides]
        h = [s.size(2) for s in sides]
        w = [s.size(3) for s in sides]
        mem = [_as_hwc(s) for s in sides]
        ctot = sum(C)
        dev = sides[0].device
        pooled = torch.empty((sp.n, ctot), dtype=torch.float32, device=dev)
        ptrs, ca, ha, wa = _lib.ptr_array([m.data_ptr() for m in mem]), _lib.int_array(C), _lib.int_array(h), _lib.int_array(w)
        if fp is not None and not fp.matches(h, w, H, W, sp.n):
            raise ValueError("pooling footprints were built for another label map or other level sizes")
        ctx.sp, ctx.geom, ctx.fp = sp, (C, h, w, H, W), fp
        if dtype is None and fp is not None:
            # fully fused over precomputed footprints: a prologue-free streaming gather
            fp.join()
            _call("wesup_levels_pool_fwd_fp", pooled, ptrs, ca, ha, wa, len(sides), H, W, sp.seg_offsets.data_ptr(),
                                               sp.seg_pixels.data_ptr(), sp.n, fp.blob.data_ptr(), pooled.data_ptr())
            feats = torch.empty(0, dtype=torch.float32, device=dev)
        elif dtype is None:
            # fully fused: superpixel means straight from the side outputs, no (H*W, C) tensor
            _call("wesup_hypercolumn_pool_fwd", pooled, ptrs, ca, ha, wa, len(sides), H, W, sp.seg_offsets.data_ptr(),
                                                 sp.seg_pixels.data_ptr(), sp.n, pooled.data_ptr())
            feats = torch.empty(0, dtype=torch.float32, device=dev)
        else:
            feats = torch.empty((H * W, ctot), dtype=dtype, device=dev)
            _call("wesup_hypercolumn_fwd", feats, ptrs, ca, ha, wa, len(sides), H, W, feats.data_ptr(), _DTYPES[dtype], HWC)
            _call("wesup_sp_pool_fwd", pooled, feats.data_ptr(), _DTYPES[dtype], HWC, sp.seg_offsets.data_ptr(),
                                        sp.seg_pixels.data_ptr(), H * W, ctot, sp.n, pooled.data_ptr())
        ctx.mark_non_differentiable(feats)
        return pooled, feats

    @staticmethod
    def backward(ctx, grad_pooled, _grad_feats):
        lib = _lib.load()
        sp = ctx.sp
        C, h, w, H, W = ctx.geom
        grad_pooled = grad_pooled.contiguous().float()
        dev = grad_pooled.device
        mem = [torch.empty((1, hh, ww, cc), dtype=torch.float32, device=dev) for cc, hh, ww in zip(C, h, w)]
        ca, ha, wa = _lib.int_array(C), _lib.int_array(h), _lib.int_array(w)
        fp = ctx.fp
        if fp is not None and fp.with_bwd:
            fp.join()
            _call("wesup_levels_pool_bwd_fp", grad_pooled, grad_pooled.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(),
                                               ca, ha, wa, len(C), H, W, sp.n, fp.blob.data_ptr(),
                                               _lib.ptr_array([m.data_ptr() for m in mem]))
        else:
            ws = _ws(lib.wesup_sp_pool_hypercolumn_bwd_workspace_bytes(ca, ha, wa, len(C), H, W, sp.n), dev)
            _call("wesup_sp_pool_hypercolumn_bwd", grad_pooled, grad_pooled.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(),
                                                    ca, ha, wa, len(C), H, W, sp.n,
                                                    _lib.ptr_array([m.data_ptr() for m in mem]), ws.data_ptr())
        return (None, None, None, None, *[m.permute(0, 3, 1, 2) for m in mem])


def hypercolumn_pool(sides: Sequence[torch.Tensor], size: Tuple[int, int], sp: SuperpixelMaps, dtype=torch.float32,
                     materialize: bool = True, footprints: Optional[Footprints] = None):
    """Hypercolumn (a) + superpixel mean pooling (b) with the fused backward.
    Returns (pooled (N,C) fp32, feats).  materialize=True (default): `feats` is the
    (H*W,C) `dtype` hypercolumn written by kernel (a) (non-differentiable output).
    materialize=False: one fused forward kernel pools straight from the side outputs,
    `feats` is None and nothing of size H*W*C is ever allocated; with `footprints`
    (`build_footprints`) the forward and the backward stream over the precomputed lists."""
    pooled, feats = _HypercolumnPool.apply(sp, (int(size[0]), int(size[1])), dtype if materialize else None, footprints,
                                           *sides)
    return pooled, (feats if materialize else None)


def paint(sp: SuperpixelMaps, sp_pred: torch.Tensor, cls: int = 1) -> torch.Tensor:
    """pred[p] = sp_pred[row(p), cls] -> (1,H,W); replaces the argmax + per-superpixel
    index_put loop (/root/reference/models/wesup.py:295-304)."""
    lib = _lib.load()
    sp_pred = sp_pred.detach().contiguous().float()
    _require_cuda(sp_pred, "sp_pred")
    out = torch.empty((1, sp.height, sp.width), dtype=torch.float32, device=sp_pred.device)
    _call("wesup_sp_paint", out, sp.row_labels.data_ptr(), sp_pred.data_ptr(), sp.height * sp.width, sp_pred.size(1), cls,
                             out.data_ptr())
    return out


# ---------------------------------------------------------------------------
# convolution whose bias gradient is a streaming column sum
# ---------------------------------------------------------------------------
def colsum(x2d: torch.Tensor) -> torch.Tensor:
    """out[c] = sum_p x2d[p, c] for a contiguous (rows, C) fp32 CUDA matrix (C % 4 == 0)."""
    lib = _lib.load()
    _require_cuda(x2d, "x2d")
    rows, c = x2d.shape
    nbytes = lib.wesup_colsum_workspace_bytes(rows, c)
    if nbytes == 0:
        raise ValueError(f"wesup_colsum does not support a ({rows}, {c}) matrix")
    out = torch.empty(c, dtype=torch.float32, device=x2d.device)
    ws = _ws(nbytes, x2d.device)
    _call("wesup_colsum", x2d, x2d.data_ptr(), rows, c, out.data_ptr(), ws.data_ptr())
    return out


class _Conv2dChannelsLast(torch.autograd.Function):
    """`F.conv2d` (cuDNN, unchanged) whose backward asks `aten::convolution_backward` for the input and weight
    gradients only and takes the bias gradient -- autograd's `grad_out.sum((0, 2, 3))`, a generic ATen reduction
    that reads channels_last gradients at a tenth of the HBM rate -- from `wesup_colsum`."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups):
        ctx.save_for_backward(x, weight)
        ctx.conf = (stride, padding, dilation, groups)
        return torch.nn.functional.conv2d(x, weight, bias, stride, padding, dilation, groups)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conf
        grad_out = grad_out.contiguous(memory_format=torch.channels_last)
        gx, gw, _ = torch.ops.aten.convolution_backward(grad_out, x, weight, None, stride, padding, dilation, False, [0, 0], groups,
                                                        [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        gb = None
        if ctx.needs_input_grad[2]:
            gb = colsum(grad_out.permute(0, 2, 3, 1).reshape(-1, grad_out.size(1)))
        return gx, gw, gb, None, None, None, None


def conv2d_channels_last(x: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    """`conv(x)` for a channels_last fp32 CUDA input, bias gradient through `wesup_colsum`; any other case
    (no bias, padding modes other than zeros, channel counts the kernel does not take) is plain `conv(x)`."""
    if (conv.bias is None or conv.padding_mode != "zeros" or not x.is_cuda or x.dtype != torch.float32
            or conv.out_channels % 4 != 0 or conv.out_channels > 1024 or isinstance(conv.padding, str)):
        return conv(x)
    return _Conv2dChannelsLast.apply(x, conv.weight, conv.bias, list(conv.stride), list(conv.padding), list(conv.dilation),
                                     conv.groups)


# ---------------------------------------------------------------------------
# (c) label propagation
# ---------------------------------------------------------------------------
_LP_ALGOS = {"auto": "wesup_label_propagate", "exact": "wesup_label_propagate_exact", "tc": "wesup_label_propagate_tc"}


def label_propagate(features: torch.Tensor, y_l: torch.Tensor, threshold: float = 0.95, return_aux: bool = False,
                    algo: str = "auto", return_stats: bool = False):
    """Fused distance -> similarity -> arg-max -> threshold -> label copy
    (/root/reference/models/wesup.py:99-139).  Returns y_u (n_u, C) [, src, max_sim].
    algo: 'auto' (library dispatch), 'exact' (CUDA cores) or 'tc' (tcgen05 filter +
    exact re-evaluation; D must be 32); all three give bit-identical results.
    return_stats (tc only): also return {'exact_evals', 'max_err_ratio'} (syncs)."""
    lib = _lib.load()
    features = features.detach().contiguous().float()
    y_l = y_l.detach().contiguous().float()
    _require_cuda(features, "features")
    n, d = features.shape
    n_l, n_cls = y_l.shape
    n_u = n - n_l
    dev = features.device
    y_u = torch.zeros((n_u, n_cls), dtype=torch.float32, device=dev)
    src = torch.zeros(n_u, dtype=torch.int32, device=dev)
    sim = torch.zeros(n_u, dtype=torch.float32, device=dev)
    stats = None
    if n_u > 0 and n_l > 0:
        ws = _ws(lib.wesup_label_propagate_workspace_bytes(n, d, n_l), dev)
        name = _LP_ALGOS[algo]
        _call(name, features, features.data_ptr(), n, d, n_l, y_l.data_ptr(), n_cls, float(threshold),
              y_u.data_ptr(), src.data_ptr(), sim.data_ptr(), ws.data_ptr())
        if return_stats and algo == "tc":
            import ctypes
            import struct
            torch.cuda.current_stream(dev).synchronize()
            out = (ctypes.c_ulonglong * 2)()
            check(lib.wesup_label_propagate_tc_stats(ws.data_ptr(), n, n_l, out), "wesup_label_propagate_tc_stats")
            stats = {"exact_evals": int(out[0]), "pairs": n_u * n_l,
                     "max_err_ratio": struct.unpack("f", struct.pack("I", int(out[1]) & 0xFFFFFFFF))[0]}
    result = (y_u, src, sim) if return_aux else (y_u,)
    if return_stats:
        result = result + (stats,)
    return result if len(result) > 1 else result[0]


def label_propagate_static(features: torch.Tensor, y_l_full: torch.Tensor, counts_dev: torch.Tensor,
                           threshold: float = 0.95) -> torch.Tensor:
    """`label_propagate` for fixed-capacity buffers whose true row counts live on the device
    (`counts_dev` = int32 {n_rows, n_labeled}): returns y_full (n_max, C) with the propagated
    labels in rows [n_labeled, n_rows) and zeros elsewhere.  No host scalar is needed, so the
    call can be captured in a CUDA graph and replayed for other images."""
    lib = _lib.load()
    features = features.detach().contiguous().float()
    y_l_full = y_l_full.detach().contiguous().float()
    _require_cuda(features, "features")
    n_max, d = features.shape
    if y_l_full.size(0) != n_max or counts_dev.dtype != torch.int32 or counts_dev.numel() < 2:
        raise ValueError("y_l_full must have n_max rows and counts_dev must be int32 {n_rows, n_labeled}")
    y_full = torch.empty((n_max, y_l_full.size(1)), dtype=torch.float32, device=features.device)
    _call("wesup_label_propagate_dev", features, features.data_ptr(), n_max, d, counts_dev.data_ptr(), y_l_full.data_ptr(),
                                        y_l_full.size(1), float(threshold), y_full.data_ptr())
    return y_full


# ---------------------------------------------------------------------------
# (d) SLIC
# ---------------------------------------------------------------------------
def slic(img: torch.Tensor, n_segments: int, compactness: float = 10.0, max_iter: int = 10,
         enforce_connectivity: bool = True):
    """GPU SLIC with scikit-image's semantics (/root/reference/models/wesup.py:471-476).
    img: (3,H,W) or (1,3,H,W) fp32 in [0,1] on the GPU.  Returns (labels int32 (H,W),
    n_labels int32 device scalar tensor of shape (1,))."""
    lib = _lib.load()
    _require_cuda(img, "img")
    if img.dim() == 4:
        if img.size(0) != 1:
            raise ValueError("SLIC takes one image at a time (the reference is batch-1)")
        img = img[0]
    if img.dim() != 3 or img.size(0) != 3:
        raise ValueError("img must be (3,H,W)")
    img = img.contiguous().float()
    _, h, w = img.shape
    nbytes = lib.wesup_slic_workspace_bytes(h, w, int(n_segments))
    if nbytes == 0:
        raise ValueError(f"degenerate SLIC configuration: {h}x{w} image, n_segments={n_segments}")
    dev = img.device
    ws = _ws(nbytes, dev)
    labels = torch.empty((h, w), dtype=torch.int32, device=dev)
    n_labels = torch.zeros(1, dtype=torch.int32, device=dev)
    _call("wesup_slic", img, img.data_ptr(), CHW, h, w, int(n_segments), float(compactness), int(max_iter),
                         int(bool(enforce_connectivity)), labels.data_ptr(), n_labels.data_ptr(), ws.data_ptr())
    return labels, n_labels


def slic_batch(imgs: torch.Tensor, n_segments: int, compactness: float = 10.0, max_iter: int = 10,
               enforce_connectivity: bool = True):
    """`slic` for B same-sized images `(B,3,H,W)` in the same launches (one Lab conversion, one persistent
    k-means kernel, one persistent connectivity kernel): returns (labels int32 (B,H,W), n_labels int32 (B,)).
    Every image's result equals the single-image call bit for bit."""
    lib = _lib.load()
    _require_cuda(imgs, "imgs")
    if imgs.dim() != 4 or imgs.size(1) != 3:
        raise ValueError("imgs must be (B,3,H,W)")
    imgs = imgs.contiguous().float()
    b, _, h, w = imgs.shape
    nbytes = lib.wesup_slic_batch_workspace_bytes(b, h, w, int(n_segments))
    if nbytes == 0:
        raise ValueError(f"degenerate SLIC configuration: {b} x {h}x{w} images, n_segments={n_segments}")
    dev = imgs.device
    ws = _ws(nbytes, dev)
    labels = torch.empty((b, h, w), dtype=torch.int32, device=dev)
    n_labels = torch.zeros(b, dtype=torch.int32, device=dev)
    _call("wesup_slic_batch", imgs, imgs.data_ptr(), CHW, b, h, w, int(n_segments), float(compactness), int(max_iter),
          int(bool(enforce_connectivity)), labels.data_ptr(), n_labels.data_ptr(), ws.data_ptr())
    return labels, n_labels


def enforce_connectivity(seg: torch.Tensor, min_size: int, max_size: int):
    """skimage's connectivity enforcement alone (the last stage of `slic`) on int label maps `(H,W)` or
    `(B,H,W)`: 4-connected pieces in raster order, search capped at `max_size`, pieces below `min_size`
    merged into the last labelled neighbour seen.  Returns (labels int32 like seg, n_labels int32 (B,))."""
    lib = _lib.load()
    _require_cuda(seg, "seg")
    batched = seg.dim() == 3
    seg3 = (seg if batched else seg.unsqueeze(0)).to(torch.int32).contiguous()
    b, h, w = seg3.shape
    dev = seg.device
    ws = _ws(lib.wesup_enforce_connectivity_workspace_bytes(b, h, w), dev)
    labels = torch.empty_like(seg3)
    n_labels = torch.zeros(b, dtype=torch.int32, device=dev)
    _call("wesup_enforce_connectivity", seg3, seg3.data_ptr(), b, h, w, int(min_size), int(max_size), labels.data_ptr(),
          n_labels.data_ptr(), ws.data_ptr())
    return (labels if batched else labels[0]), n_labels


# ---------------------------------------------------------------------------
# forward-only helpers writing into caller-owned buffers (tiled inference: everything below runs inside one
# CUDA graph per tile batch, so nothing here may allocate or read a host scalar)
# ---------------------------------------------------------------------------
class StaticSuperpixelBuffers:
    """Fixed-capacity device buffers for the superpixel statistics of ONE image of a known size: `sp_stats_into`
    refills them for every new label map, kernels captured in a CUDA graph keep reading the same addresses.
    `bound` = upper bound of the superpixel count (every kept SLIC piece has >= min_size pixels); ids that do not
    occur are empty superpixels at the end of the row order, so the first `cap >= N` rows are a complete description
    for any cap (seg_offsets[k] == H*W for every k >= N).  With `flat` the arrays are carved out of a caller-owned
    int32 tensor at `offset` (several images in one tensor: one copy moves them all)."""

    def __init__(self, height: int, width: int, bound: int, device, flat: Optional[torch.Tensor] = None, offset: int = 0):
        self.height, self.width, self.bound = int(height), int(width), int(bound)
        hw = height * width
        r4 = lambda n: (n + 3) // 4 * 4                                    # noqa: E731  (16-byte aligned sub-arrays)
        need = 2 * r4(bound) + r4(bound + 1) + 2 * r4(hw)
        if flat is None:
            flat, offset = torch.empty(need, dtype=torch.int32, device=device), 0
        if flat.dtype != torch.int32 or offset % 4 != 0 or offset + need > flat.numel():
            raise ValueError("flat must be an int32 tensor with room for the buffers at a 16-byte aligned offset")
        o = offset
        self.order = flat[o:o + bound]; o += r4(bound)
        self.counts = flat[o:o + bound]; o += r4(bound)
        self.seg_offsets = flat[o:o + bound + 1]; o += r4(bound + 1)
        self.row_labels = flat[o:o + hw]; o += r4(hw)
        self.seg_pixels = flat[o:o + hw]
        self.n_labeled = torch.zeros(1, dtype=torch.int32, device=flat.device)

    def view(self, cap: int) -> SuperpixelMaps:
        """The first `cap` rows as a `SuperpixelMaps` (no copy)."""
        return SuperpixelMaps(self.height, self.width, cap, self.order[:cap], self.row_labels, self.counts[:cap],
                              self.seg_offsets[:cap + 1], self.seg_pixels, None, None)


def sp_stats_into(labels: torch.Tensor, buf: StaticSuperpixelBuffers, ws: Optional[torch.Tensor] = None) -> None:
    """`SuperpixelMaps.from_labels(labels, None, n_sp=buf.bound)` into `buf` (labels: (H,W) int32 contiguous).
    `ws`: workspace of wesup_sp_stats_workspace_bytes(H, W, bound, 0) bytes (allocated here when omitted)."""
    if ws is None:
        ws = _ws(_lib.load().wesup_sp_stats_workspace_bytes(buf.height, buf.width, buf.bound, 0), labels.device)
    _call("wesup_sp_stats", labels, labels.data_ptr(), None, buf.height, buf.width, 0, buf.bound, buf.order.data_ptr(),
          buf.row_labels.data_ptr(), buf.counts.data_ptr(), buf.seg_offsets.data_ptr(), buf.seg_pixels.data_ptr(), None,
          buf.n_labeled.data_ptr(), ws.data_ptr())


def slic_batch_into(imgs: torch.Tensor, n_segments: int, compactness: float, labels: torch.Tensor, n_labels: torch.Tensor,
                    ws: torch.Tensor, max_iter: int = 10) -> None:
    """`slic_batch` into caller-owned outputs and workspace (no allocation): imgs (B,3,H,W) fp32 contiguous,
    labels (B,H,W) int32, n_labels (B,) int32, ws >= wesup_slic_batch_workspace_bytes(B,H,W,n_segments) bytes."""
    b, _, h, w = imgs.shape
    if not imgs.is_contiguous() or imgs.dtype != torch.float32 or labels.dtype != torch.int32 or not labels.is_contiguous():
        raise ValueError("imgs must be contiguous fp32 (B,3,H,W) and labels contiguous int32 (B,H,W)")
    _call("wesup_slic_batch", imgs, imgs.data_ptr(), CHW, b, h, w, int(n_segments), float(compactness), int(max_iter), 1,
          labels.data_ptr(), n_labels.data_ptr(), ws.data_ptr())


def levels_pool_fwd_into(levels: Sequence[torch.Tensor], size: Tuple[int, int], sp: SuperpixelMaps, out: torch.Tensor) -> None:
    """Superpixel means straight from the feature levels (in-kernel footprints, no preprocessing) written into
    `out (sp.n, sum C)`; `levels` are (1,C,h,w) fp32 tensors in channels_last memory (or batch slices of one)."""
    C = [s.size(1) for s in levels]
    h = [s.size(2) for s in levels]
    w = [s.size(3) for s in levels]
    mem = [_as_hwc(s) for s in levels]
    if out.shape != (sp.n, sum(C)) or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("out must be a contiguous fp32 (N, sum C) tensor")
    _call("wesup_levels_pool_fwd", out, _lib.ptr_array([m.data_ptr() for m in mem]), _lib.int_array(C), _lib.int_array(h),
          _lib.int_array(w), len(levels), int(size[0]), int(size[1]), sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(), sp.n,
          out.data_ptr())


def hypercolumn_into(sides: Sequence[torch.Tensor], size: Tuple[int, int], out: torch.Tensor) -> None:
    """Kernel (a) into a caller-owned pixel-major `(H*W, sum C)` fp32 / bf16 buffer (forward only): `sides` are
    (1,C,h,w) fp32 tensors in channels_last memory (or batch slices of one)."""
    H, W = int(size[0]), int(size[1])
    C = [s.size(1) for s in sides]
    mem = [_as_hwc(s) for s in sides]
    if out.shape != (H * W, sum(C)) or not out.is_contiguous():
        raise ValueError("out must be a contiguous (H*W, sum C) tensor")
    _call("wesup_hypercolumn_fwd", out, _lib.ptr_array([m.data_ptr() for m in mem]), _lib.int_array(C),
          _lib.int_array([s.size(2) for s in sides]), _lib.int_array([s.size(3) for s in sides]), len(sides), H, W,
          out.data_ptr(), _DTYPES[out.dtype], HWC)


def upsample_sum(terms: Sequence[torch.Tensor], size: Tuple[int, int], bias: Optional[torch.Tensor] = None, relu: bool = False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[p,:] = act(bias + sum_g bilinear(align_corners=True)(terms[g])[p,:]) -> (H*W, C) (forward only).
    `terms`: pixel-major (h_g, w_g, C) tensors of one dtype (fp32 or bf16), at most 5; a term of size (H, W) is
    added as it is.  See csrc/upsample_sum.cu: with terms = per-resolution products of the backbone levels with the
    folded first-layer weights this is ReLU(Linear1(hypercolumn)) of the pixel-wise model without the hypercolumn."""
    H, W = int(size[0]), int(size[1])
    t0 = terms[0]
    _require_cuda(t0, "terms")
    C = t0.size(-1)
    for t in terms:
        if t.dim() != 3 or t.size(-1) != C or t.dtype != t0.dtype or not t.is_contiguous():
            raise ValueError("terms must be contiguous (h, w, C) tensors of one dtype and channel count")
    if out is None:
        out = torch.empty((H * W, C), dtype=t0.dtype, device=t0.device)
    elif out.shape != (H * W, C) or out.dtype != t0.dtype or not out.is_contiguous():
        raise ValueError("out must be a contiguous (H*W, C) tensor of the terms' dtype")
    b = None if bias is None else bias.detach().float().contiguous()
    _call("wesup_upsample_sum", out, _lib.ptr_array([t.data_ptr() for t in terms]), _lib.int_array([t.size(0) for t in terms]),
          _lib.int_array([t.size(1) for t in terms]), len(terms), H, W, C, _DTYPES[t0.dtype], None if b is None else b.data_ptr(),
          int(bool(relu)), out.data_ptr())
    return out


def paint_into(sp: SuperpixelMaps, sp_pred: torch.Tensor, out: torch.Tensor, cls: int = 1) -> None:
    """`paint` into a caller-owned fp32 (H,W) buffer."""
    _call("wesup_sp_paint", out, sp.row_labels.data_ptr(), sp_pred.data_ptr(), sp.height * sp.width, sp_pred.size(1), cls,
          out.data_ptr())
