"""Data parallelism for the two places the path shards:
image batches for training (one exchange step: gradient all-reduce over
NCCL/NVLink) and tiles for tiled inference (no data-path collective, only a
gather of finished tiles).  One process per GPU, `torch.distributed` plumbing.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise the default process group from torchrun's environment.
    Returns (rank, world_size, local_rank); a no-op single-process answer when
    WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


class GradientAllReduce:
    """Gradient averaging over ranks, overlapped with backward.

    All parameter gradients live as views into a single flat fp32 buffer (18.87 M parameters =
    75.5 MB for WESUP), laid out in `model.parameters()` order.  The buffer is cut into a few
    contiguous BUCKETS; backward produces gradients in roughly reverse parameter order, so the
    bucket at the tail of the buffer (classifier, MLP, side convolutions) is complete first and
    the one at the head (first backbone convolutions) last.  A post-accumulate hook on every
    parameter counts arrivals per bucket and, when a bucket is complete, starts ONE asynchronous
    all-reduce of its slice (NCCL: ReduceOp.AVG, so no separate scaling pass) on the process
    group's communication stream while backward keeps running on the compute stream; `finish()`
    makes the compute stream wait for all of them before the optimizer step.  Hooks and
    collectives are plain stream work: they are captured into the training CUDA graph, so a
    replayed iteration needs no Python at all between backward and the SGD step.  Every parameter
    receives a gradient every step on this path, so no unused-parameter
    handling is needed."""

    def __init__(self, model: torch.nn.Module, process_group=None, bucket_mb: float = 24.0):
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.offsets = []
        offset = 0
        for p in self.params:
            self.offsets.append(offset)
            p.grad = self._view(p, offset)
            offset += p.numel()
        backend = dist.get_backend(process_group) if dist.is_initialized() else None
        self._avg = backend == "nccl"                      # gloo has no AVG: sum, then scale
        # buckets: contiguous parameter ranges, filled from the tail of the buffer (first to be ready)
        limit = int(bucket_mb * (1 << 20) / self.flat.element_size())
        self.buckets = []                                  # (first param index, last param index + 1), tail first
        hi = len(self.params)
        while hi > 0:
            lo, size = hi, 0
            while lo > 0 and (size == 0 or size + self.params[lo - 1].numel() <= limit):
                lo -= 1
                size += self.params[lo].numel()
            self.buckets.append((lo, hi))
            hi = lo
        self._bucket_of = {}
        for bi, (lo, hi) in enumerate(self.buckets):
            for i in range(lo, hi):
                self._bucket_of[i] = bi
        self._arrived = [0] * len(self.buckets)
        self._works = []
        self._hooks = []
        self.overlap = False
        self.suspended = False                             # True: hooks and finish() issue no collective (graph warm-up runs)

    def _view(self, p, offset):
        """The parameter's slot of the flat buffer, with the parameter's own strides (channels_last
        convolution weights keep matching gradient strides, so the optimizer's multi-tensor path applies)."""
        dense = p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last) if p.dim() == 4 else p.is_contiguous()
        if dense:
            return torch.as_strided(self.flat, p.size(), p.stride(), offset)
        return self.flat[offset:offset + p.numel()].view_as(p)

    def bucket_slice(self, bi: int) -> torch.Tensor:
        lo, hi = self.buckets[bi]
        end = self.offsets[hi] if hi < len(self.params) else self.flat.numel()
        return self.flat[self.offsets[lo]:end]

    def zero_grad(self):
        """Use instead of optimizer.zero_grad(set_to_none=True), which would drop the views."""
        self.flat.zero_()
        self._arrived = [0] * len(self.buckets)
        self._works = []

    def rebind(self):
        """Re-attach any gradient that was replaced (e.g. by optimizer.zero_grad(set_to_none=True)) to its
        slot of the flat buffer.  Not needed on the iteration path, which never drops the views."""
        for p, offset in zip(self.params, self.offsets):
            view = self._view(p, offset)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def _reduce(self, t: torch.Tensor, async_op: bool):
        if self._avg:
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if not async_op:
            t.div_(self.world_size)
        return work

    # -- overlapped path ---------------------------------------------------------------------
    def enable_overlap(self):
        """Install the per-parameter hooks that start a bucket's all-reduce as soon as backward has
        produced its last gradient."""
        if self.overlap:
            return
        self.overlap = True
        for i, p in enumerate(self.params):
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    def _make_hook(self, index: int):
        bi = self._bucket_of[index]
        lo, hi = self.buckets[bi]

        def hook(_param):
            if self.suspended:
                return
            self._arrived[bi] += 1
            if self._arrived[bi] == hi - lo and self.world_size > 1:
                self._works.append((bi, self._reduce(self.bucket_slice(bi), async_op=True)))
        return hook

    def finish(self):
        """After backward: the compute stream waits for every bucket (any bucket whose hook did not
        fire -- e.g. hooks disabled -- is reduced here)."""
        if self.world_size == 1 or self.suspended:
            return
        started = {bi for bi, _ in self._works}
        for bi in range(len(self.buckets)):
            if bi not in started:
                self._works.append((bi, self._reduce(self.bucket_slice(bi), async_op=True)))
        for bi, work in self._works:
            if work is not None:
                work.wait()
            if not self._avg:
                self.bucket_slice(bi).div_(self.world_size)
        self._works = []
        self._arrived = [0] * len(self.buckets)

    def average_gradients(self):
        """Blocking form (no overlap): one all-reduce per bucket after backward."""
        self.rebind()
        self.finish()

    def broadcast_parameters(self, src: int = 0):
        if self.world_size > 1:
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous block partition: rank r owns [lo, hi).  Blocks differ by at most
    one item; contiguous so each rank's tiles form a stripe of the slide."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_tiles(local_tiles: torch.Tensor, n_total: int, rank: int, world_size: int, group=None):
    """Gather per-rank stacks of finished tiles (n_local, ...) to rank 0 in tile
    order.  Returns the (n_total, ...) stack on rank 0 and None elsewhere."""
    if world_size == 1:
        return local_tiles
    counts = [shard_range(n_total, r, world_size) for r in range(world_size)]
    biggest = max(hi - lo for lo, hi in counts)
    pad = torch.zeros((biggest, *local_tiles.shape[1:]), dtype=local_tiles.dtype, device=local_tiles.device)
    pad[: local_tiles.size(0)] = local_tiles
    bufs = [torch.empty_like(pad) for _ in range(world_size)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(counts)], dim=0)
