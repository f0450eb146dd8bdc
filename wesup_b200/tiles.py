"""Tiled inference: tile split / merge with the reference's geometry
(/root/reference/infer_tile.py:23-91: tile origins are np.linspace(0, size - patch, ceil(size / patch),
dtype=int) per axis, tiles may overlap, overlapping predictions are combined by a running mean in float64
in tile order) and the batched device pipeline that replaces the reference's per-tile Python loop
(/root/reference/infer_tile.py:105-116, /root/reference/pixel_infer_tile.py:45-57).

Device pipeline: tiles travel as uint8, `batch` tiles at a time; per batch ONE
uint8->fp32 conversion, ONE batched GPU SLIC (superpixel-wise mode), VGG16 at batch size `batch`, the
superpixel stage per tile, all inside one CUDA graph per (batch, superpixel capacity); predictions are written
as uint8 / fp32 into a device-resident stack and merged on the device (in place when tiles are disjoint, else
the reference's running mean in the same float64 operation order); one device-to-host copy at the end.
Preprocessing of batch k+1 (conversion, SLIC, superpixel statistics) runs on a side stream while batch k is in
the network; the host learns the superpixel counts of a batch from one small asynchronous read that has long
completed when it is needed.  Ranks own contiguous stripes of the row-major tile list and only finished
predictions travel between them (no data-path collective).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np

_NP_DTYPES = {}


def _fill_np_dtypes():
    import torch
    _NP_DTYPES.update({torch.uint8: np.uint8, torch.float32: np.float32, torch.float64: np.float64, torch.float16: np.float16,
                       torch.int32: np.int32, torch.int64: np.int64, torch.bool: np.bool_})


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


def tiles_are_disjoint(height: int, width: int, patch_size: int) -> bool:
    return height % patch_size == 0 and width % patch_size == 0


def combine_disjoint(patches, target_height: int, target_width: int):
    """Fast path of `combine_patches_to_image` when no two tiles overlap: the running mean of one sample is the
    sample itself, so tiles are written in place (bit-identical result, no float64 (H,W,C+1) scratch).  Works on
    numpy arrays and on torch tensors (device-side merge: one permuting copy)."""
    p = patches.shape[1]
    ny, nx = target_height // p, target_width // p
    tail = tuple(patches.shape[3:])
    grid = patches.reshape(ny, nx, p, p, *tail)
    axes = (0, 2, 1, 3) + tuple(range(4, 4 + len(tail)))
    if isinstance(patches, np.ndarray):
        return np.squeeze(grid.transpose(axes).reshape(target_height, target_width, *tail))
    return grid.permute(*axes).reshape(target_height, target_width, *tail)


def combine_on_device(stack, target_height: int, target_width: int):
    """`combine_patches_to_image` on the device holding `stack` (n,p,p[,C]): in place when the tiles are disjoint,
    else the reference's running mean, tile by tile in tile order, in float64 with the reference's operation order
    ((acc * seen + tile) / (seen + 1)) -- IEEE multiplication, addition and division round identically on both sides,
    so the result equals the host merge bit for bit.  Returns a float64 tensor (or `stack`'s dtype when disjoint)."""
    import torch
    p = stack.shape[1]
    if tiles_are_disjoint(target_height, target_width, p) and stack.shape[0] == (target_height // p) * (target_width // p):
        return combine_disjoint(stack, target_height, target_width)
    tiles3 = stack if stack.dim() == 4 else stack.unsqueeze(-1)
    acc = torch.zeros((target_height, target_width, tiles3.shape[-1]), dtype=torch.float64, device=stack.device)
    seen = torch.zeros((target_height, target_width, 1), dtype=torch.float64, device=stack.device)
    for tile, (t, l) in zip(tiles3, top_left_coordinates(target_height, target_width, p)):
        a, s = acc[t:t + p, l:l + p], seen[t:t + p, l:l + p]
        a.copy_((a * s + tile.to(torch.float64)) / (s + 1))
        s.add_(1.0)
    return acc.squeeze(-1) if stack.dim() == 3 else acc


# ---------------------------------------------------------------------------
# tile engines: `enqueue(tiles_u8)` starts the preprocessing of a batch, `finish(token)` runs the network
# ---------------------------------------------------------------------------
class _Engine:
    """Common plumbing of the two engines: fixed-shape staging / active buffers, one CUDA graph per variant of the
    network part, eager execution for odd batch sizes."""

    out_dtype = None

    def __init__(self, device, batch: int, use_graph: bool = True, warm: int = 1):
        import torch
        self.device = torch.device(device)
        self.batch, self.use_graph, self.warm = int(batch), bool(use_graph), int(warm)
        self.side = torch.cuda.Stream(device=self.device)
        self.capture_stream = torch.cuda.Stream(device=self.device)
        self.graphs, self.graph_pool, self.seen = {}, None, {}
        self.shape = None
        self.launches = 0                      # own-library launches issued or replayed (bench: gpu_launches)

    def _graphed(self, key, fn):
        """Run `fn()` (which reads and writes only engine-owned buffers) as a CUDA graph keyed by `key`: the first
        `warm` calls run eagerly, the next is captured, later ones replay."""
        import torch
        from . import _lib
        entry = self.graphs.get(key)
        if entry is not None:
            entry[0].replay()
            self.launches += entry[1]
            return
        seen = self.seen.get(key, 0)
        self.seen[key] = seen + 1
        lib = _lib.load()
        if not self.use_graph or seen < self.warm:
            n0 = lib.wesup_kernel_launches()
            fn()
            self.launches += int(lib.wesup_kernel_launches() - n0)
            return
        main = torch.cuda.current_stream(self.device)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        opts = {"pool": self.graph_pool} if self.graph_pool is not None else {}
        n0 = lib.wesup_kernel_launches()
        with torch.cuda.graph(graph, stream=self.capture_stream, **opts):
            fn()
        n = int(lib.wesup_kernel_launches() - n0)
        if self.graph_pool is None:
            self.graph_pool = graph.pool()
        self.graphs[key] = (graph, n)
        main.wait_stream(self.capture_stream)
        graph.replay()
        self.launches += n


class SuperpixelTileEngine(_Engine):
    """Superpixel-wise tile inference (what /root/reference/infer_tile.py:111-116 computes per tile:
    `postprocess(model(preprocess(tile)))`), `batch` tiles per step."""

    def __init__(self, trainer, batch: int = 16, use_graph: bool = True):
        import torch
        super().__init__(trainer.device, batch, use_graph)
        self.trainer, self.model = trainer, trainer.model
        self.out_dtype = torch.uint8
        self.parity = 0

    def _setup(self, h: int, w: int):
        import torch
        from . import _lib, ops
        dev, B = self.device, self.batch
        self.shape = (h, w)
        self.n_segments = int(h * w / self.trainer.kwargs.get("sp_area"))
        hw = h * w
        self.bound = hw // max(int(0.5 * hw / self.n_segments), 1) + 1        # every kept SLIC piece has >= min_size pixels
        per_tile = _round4(self.bound) * 2 + _round4(self.bound + 1) + 2 * _round4(hw)
        lib = _lib.load()

        def buffer_set():
            flat = torch.empty(B * per_tile, dtype=torch.int32, device=dev)
            bufs = [ops.StaticSuperpixelBuffers(h, w, self.bound, dev, flat=flat, offset=t * per_tile) for t in range(B)]
            return {"flat": flat, "bufs": bufs, "img": torch.empty((B, 3, h, w), dtype=torch.float32, device=dev)}
        self.staging = [buffer_set(), buffer_set()]
        self.active = buffer_set()
        for s in self.staging:
            s["labels"] = torch.empty((B, h, w), dtype=torch.int32, device=dev)
            s["n_dev"] = torch.zeros(B, dtype=torch.int32, device=dev)
            s["n_host"] = torch.empty(B, dtype=torch.int32, pin_memory=True)
            s["slic_ws"] = torch.empty(lib.wesup_slic_batch_workspace_bytes(B, h, w, self.n_segments), dtype=torch.uint8, device=dev)
            s["ready"], s["free"] = torch.cuda.Event(), torch.cuda.Event()
            s["free"].record(torch.cuda.current_stream(dev))
        self.out = torch.empty((B, h, w), dtype=torch.uint8, device=dev)
        self.pred = torch.empty((B, h, w), dtype=torch.float32, device=dev)
        self.stats_ws = torch.empty(lib.wesup_sp_stats_workspace_bytes(h, w, self.bound, 0), dtype=torch.uint8, device=dev)

    def enqueue(self, tiles_u8):
        """Start preprocessing (uint8 -> fp32, batched GPU SLIC, superpixel statistics, async read of the counts) of
        `tiles_u8 (b,h,w,3)` uint8 on the device, on the side stream.  Returns a token for `finish`."""
        import torch
        from . import ops
        b, h, w, _ = tiles_u8.shape
        if self.shape != (h, w):
            if self.shape is not None:
                self.graphs, self.seen = {}, {}
            self._setup(h, w)
        if b > self.batch:
            raise ValueError(f"batch of {b} tiles exceeds the engine's batch size {self.batch}")
        s = self.staging[self.parity]
        self.parity ^= 1
        main, side = torch.cuda.current_stream(self.device), self.side
        side.wait_stream(main)                     # the tiles were produced on the main stream
        side.wait_event(s["free"])                 # ... and this staging set has been copied out
        with torch.cuda.stream(side):
            img = s["img"][:b]
            img.copy_(tiles_u8.permute(0, 3, 1, 2))                 # uint8 -> fp32 ...
            img.div_(255.0)                                         # ... / 255 == TF.to_tensor
            ops.slic_batch_into(img, self.n_segments, self.trainer.kwargs.get("sp_compactness"), s["labels"][:b],
                                s["n_dev"][:b], s["slic_ws"])
            for t in range(b):
                ops.sp_stats_into(s["labels"][t], s["bufs"][t], ws=self.stats_ws)
            s["n_host"][:b].copy_(s["n_dev"][:b], non_blocking=True)
            s["ready"].record(side)
        tiles_u8.record_stream(side)
        self.launches += 3 + 5 * b
        return (s, b)

    def _forward(self, b: int, cap: int):
        """Network part for the first `b` tiles of the active set, `cap` rows per tile: VGG16 at batch b, superpixel
        means of the 13 backbone levels per tile (in-kernel footprints), side convolutions + MLP + classifier on all
        b*cap rows at once, paint + round per tile."""
        import torch
        import torch.nn as nn
        import torch.nn.functional as F
        from . import ops
        m, act = self.model, self.active
        h, w = self.shape
        with torch.no_grad():
            x = act["img"][:b].contiguous(memory_format=torch.channels_last)
            outs = []
            for layer in m.backbone:
                if isinstance(layer, nn.Conv2d):
                    x = layer(x)
                    outs.append(x)
                elif isinstance(layer, nn.ReLU):
                    x = F.relu(x)
                else:
                    x = layer(x)
            ctot = sum(o.size(1) for o in outs)
            pooled = torch.empty((b * cap, ctot), dtype=torch.float32, device=self.device)
            views = [act["bufs"][t].view(cap) for t in range(b)]
            for t in range(b):
                ops.levels_pool_fwd_into([o[t:t + 1] for o in outs], (h, w), views[t], pooled[t * cap:(t + 1) * cap])
            cols = []
            for name, part in zip(m._side_names, pooled.split([o.size(1) for o in outs], dim=1)):
                conv = getattr(m, name)
                cols.append(F.linear(part, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias))
            sp_pred = m.classifier(m.fc_layers(torch.cat(cols, dim=1))).contiguous()
            for t in range(b):
                ops.paint_into(views[t], sp_pred[t * cap:(t + 1) * cap], self.pred[t], cls=1)
            self.out[:b].copy_(self.trainer.postprocess(self.pred[:b]))

    def finish(self, token):
        """Run the network on a batch whose preprocessing was enqueued earlier.  Returns the (b,h,w) uint8 class maps
        -- a view of an engine-owned buffer, valid until the next `finish`."""
        import torch
        s, b = token
        s["ready"].synchronize()                                    # the one host wait per batch (enqueued a batch ago)
        counts = s["n_host"][:b].tolist()
        cap = -(-max(counts) // 64) * 64
        if cap > self.bound:
            raise RuntimeError(f"superpixel count {max(counts)} exceeds the bound {self.bound}")
        main = torch.cuda.current_stream(self.device)
        main.wait_event(s["ready"])
        self.active["flat"].copy_(s["flat"], non_blocking=True)
        self.active["img"][:b].copy_(s["img"][:b], non_blocking=True)
        s["free"].record(main)
        if b == self.batch:
            self._graphed((b, cap, self.model.training), lambda: self._forward(b, cap))
        else:
            from . import _lib
            n0 = _lib.load().wesup_kernel_launches()
            self._forward(b, cap)
            self.launches += int(_lib.load().wesup_kernel_launches() - n0)
        return self.out[:b]


class PixelTileEngine(_Engine):
    """Pixel-wise tile inference (/root/reference/pixel_infer_tile.py:50-54: class-1 probability of
    `WESUPPixelInference.forward` per tile), `batch` tiles per step."""

    def __init__(self, model, batch: int = 8, use_graph: bool = True):
        import torch
        super().__init__(next(model.parameters()).device, batch, use_graph)
        self.model = model
        self.out_dtype = torch.float32

    def enqueue(self, tiles_u8):
        import torch
        b, h, w, _ = tiles_u8.shape
        if self.shape != (h, w):
            self.graphs, self.seen, self.shape = {}, {}, (h, w)
            self.img = torch.empty((self.batch, 3, h, w), dtype=torch.float32, device=self.device)
            self.out = torch.empty((self.batch, h, w), dtype=torch.float32, device=self.device)
        return (tiles_u8, b)

    def _forward(self, b: int):
        import torch
        with torch.no_grad():
            self.out[:b].copy_(self.model.forward_batch(self.img[:b])[..., 1])

    def finish(self, token):
        tiles_u8, b = token
        img = self.img[:b]
        img.copy_(tiles_u8.permute(0, 3, 1, 2))
        img.div_(255.0)
        if b == self.batch:
            self._graphed((b, self.model.training), lambda: self._forward(b))
        else:
            self._forward(b)
        return self.out[:b]


class FunctionEngine(_Engine):
    """Adapter for a plain per-tile function `step((1,3,h,w) fp32) -> (h,w[,C])` (the reference's loop body)."""

    def __init__(self, step, device, out_dtype=None):
        super().__init__(device, 1, use_graph=False)
        self.step, self.out_dtype = step, out_dtype

    def enqueue(self, tiles_u8):
        return (tiles_u8, tiles_u8.shape[0])

    def finish(self, token):
        import torch
        tiles_u8, b = token
        outs = [self.step(tiles_u8[k].permute(2, 0, 1).float().div_(255.0).unsqueeze(0).contiguous()) for k in range(b)]
        out = torch.stack(outs)
        return out if self.out_dtype is None else out.to(self.out_dtype)


def _round4(n: int) -> int:
    return (n + 3) // 4 * 4


# ---------------------------------------------------------------------------
# the driver
# ---------------------------------------------------------------------------
def predict_tiles(engine, img, patch_size: int, device=None, rank: int = 0, world_size: int = 1, group=None,
                  chunk: int = 256, return_device: bool = False):
    """Run `engine` over the tiles of `img` this rank owns (a contiguous block of the row-major tile list, so a
    rank's output is a stripe of the slide), gather the finished tiles to rank 0 in tile order and merge them there
    on the device.  `img`: (H,W,3) uint8 numpy array on the host -- tiles are cut `chunk` at a time into one of two
    pinned staging buffers and travel in one asynchronous copy per chunk -- or a uint8 CUDA tensor already resident
    on the device (tiles are then gathered there).  Returns the merged (H,W[,C]) array on rank 0 (numpy, or the
    device tensor with `return_device`) and None elsewhere.  A callable `engine` is wrapped in `FunctionEngine`."""
    import torch
    from .parallel import gather_tiles, shard_range
    if not _NP_DTYPES:
        _fill_np_dtypes()
    if not hasattr(engine, "enqueue"):
        engine = FunctionEngine(engine, device or "cuda")
    device = engine.device
    height, width = int(img.shape[0]), int(img.shape[1])
    coords = top_left_coordinates(height, width, patch_size)
    lo, hi = shard_range(len(coords), rank, world_size)
    mine = coords[lo:hi]
    on_device = torch.is_tensor(img) and img.is_cuda
    ph, pw = min(patch_size, height), min(patch_size, width)
    B = engine.batch
    local = None
    # Host images travel as BANDS of rows, not as tiles: the `ph` image rows a row of tiles covers are one contiguous
    # block of the array -- one memcpy into one of two pinned staging buffers and one asynchronous copy per band (24 MB
    # for a 20 000-px slide) -- and the tiles are cut out of the band on the device.  (r2: cutting every tile on the
    # host, 2500 strided numpy copies per slide, cost 0.25 s of the 2.2 s an end-to-end slide took on one GPU and half
    # of the 0.48 s on eight.)
    stages, copied, bands = None, [None, None], {}
    n_uploads = 0
    if not on_device and mine:
        stages = [torch.empty((ph, width, 3), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    cols = torch.arange(pw, device=device).view(1, -1)

    def band_of(top):
        nonlocal n_uploads
        if top in bands:
            return bands[top]
        slot = n_uploads & 1
        n_uploads += 1
        if copied[slot] is not None:
            copied[slot].synchronize()                # the copy that last read this buffer has completed
        np.copyto(stages[slot].numpy(), img[top:top + ph, :, :3])
        dev_band = stages[slot].to(device, non_blocking=True)
        copied[slot] = torch.cuda.Event()
        copied[slot].record(torch.cuda.current_stream(device))
        if len(bands) >= 2:                           # a chunk of tiles straddles at most a few bands; keep the last two
            bands.pop(next(iter(bands)))
        bands[top] = dev_band
        return dev_band

    # Single process, disjoint tiles, numpy result: finished BANDS of the result go back to the host while the later
    # tiles are still in the network (copy stream, one chunk behind), instead of one device-to-host copy of the whole
    # merged slide at the end -- the copy and the first-touch page faults of the fresh result array (0.25 s per GB)
    # then hide behind the GPU work.
    p_tile = min(patch_size, height)
    stream_out = (not return_device and world_size == 1 and ph == pw == patch_size and tiles_are_disjoint(height, width, patch_size)
                  and len(coords) == (height // patch_size) * (width // patch_size))
    nx_tiles = width // patch_size if stream_out else 0
    out_np, out_t, bands_out, copy_stream, chunk_done = None, None, 0, None, None

    def drain_bands(upto_tiles):
        """Copy the complete bands among the first `upto_tiles` tiles (all enqueued before `chunk_done` was recorded)."""
        nonlocal out_np, out_t, bands_out
        b1 = upto_tiles // nx_tiles
        if b1 <= bands_out or local is None:
            return
        tail = tuple(local.shape[3:])
        if out_np is None:
            out_np = np.empty((height, width, *tail), dtype=_NP_DTYPES[local.dtype])
            out_t = torch.from_numpy(out_np)
        copy_stream.wait_event(chunk_done)
        with torch.cuda.stream(copy_stream):
            grid = local[bands_out * nx_tiles:b1 * nx_tiles].reshape(b1 - bands_out, nx_tiles, p_tile, p_tile, *tail)
            axes = (0, 2, 1, 3) + tuple(range(4, 4 + len(tail)))
            block = grid.permute(*axes).reshape((b1 - bands_out) * p_tile, width, *tail).contiguous()
            out_t[bands_out * p_tile:b1 * p_tile].copy_(block)            # pageable target: returns when the rows have landed
        bands_out = b1

    if stream_out:
        copy_stream = torch.cuda.Stream(device=device)
    done = 0
    for ci, c0 in enumerate(range(0, len(mine), chunk)):
        part = mine[c0:c0 + chunk]
        if on_device:
            tops = torch.tensor([t for t, _ in part], device=device).view(-1, 1, 1) + torch.arange(ph, device=device).view(1, -1, 1)
            lefts = torch.tensor([l for _, l in part], device=device).view(-1, 1, 1) + torch.arange(pw, device=device).view(1, 1, -1)
            dev_u8 = img[tops, lefts][..., :3].contiguous()
        else:
            pieces, k = [], 0
            while k < len(part):                      # runs of tiles sharing their top row = pieces of one band
                top, k1 = part[k][0], k
                while k1 < len(part) and part[k1][0] == top:
                    k1 += 1
                lefts = torch.tensor([l for _, l in part[k:k1]], device=device).view(-1, 1) + cols      # (n, pw)
                pieces.append(band_of(top)[:, lefts].permute(1, 0, 2, 3))                              # (n, ph, pw, 3)
                k = k1
            dev_u8 = (pieces[0] if len(pieces) == 1 else torch.cat(pieces)).contiguous()
        batches = [dev_u8[i:i + B] for i in range(0, len(part), B)]
        token = engine.enqueue(batches[0])
        for j in range(len(batches)):
            nxt = engine.enqueue(batches[j + 1]) if j + 1 < len(batches) else None      # preprocessing runs one batch ahead
            out = engine.finish(token)
            if local is None:
                local = torch.empty((len(mine), *out.shape[1:]), dtype=engine.out_dtype or out.dtype, device=device)
            local[done:done + out.shape[0]].copy_(out)
            done += out.shape[0]
            token = nxt
        if stream_out:
            if chunk_done is not None:
                drain_bands(c0)                       # tiles before this chunk: finished while this chunk was being enqueued
            chunk_done = torch.cuda.Event()
            chunk_done.record(torch.cuda.current_stream(device))
    if stream_out and local is not None:
        drain_bands(len(mine))
        copy_stream.synchronize()
        return out_np
    if local is None:
        local = torch.zeros((0, ph, pw), dtype=engine.out_dtype or torch.float32, device=device)
    stack = gather_tiles(local, len(coords), rank, world_size, group=group)
    if stack is None:
        return None
    merged = combine_on_device(stack, height, width)
    return merged if return_device else merged.cpu().numpy()
