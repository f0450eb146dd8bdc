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
    """Gradient averaging over ranks, overlapped with backward, without a staging copy.

    The parameters are cut into a few BUCKETS of consecutive parameters (18.87 M parameters = 75.5 MB for
    WESUP).  Backward produces gradients in roughly reverse parameter order, so the bucket at the tail
    (classifier, MLP, side convolutions) is complete first and the one at the head (first backbone
    convolutions) last.  A post-accumulate hook on every parameter counts arrivals per bucket and, when a bucket
    is complete, starts ONE coalesced asynchronous all-reduce over the bucket's gradient tensors where autograd
    left them (NCCL: one grouped launch, ReduceOp.AVG, so neither a flat staging buffer nor a scaling pass: a
    round-1 flat buffer cost a 75 MB memset plus one accumulate-add per parameter, 0.18 ms per image) on the
    process group's communication stream while backward keeps running on the compute stream; `finish()` makes the
    compute stream wait for all of them before the optimizer step.  Hooks and collectives are plain stream work:
    they are captured into the training CUDA graph, so a replayed iteration needs no Python between backward and
    the SGD step.  Every parameter receives a gradient every step on this path, so no
    unused-parameter handling is needed."""

    def __init__(self, model: torch.nn.Module, process_group=None, bucket_mb: float = 10.0):
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        backend = dist.get_backend(process_group) if dist.is_initialized() else None
        self._avg = backend == "nccl"                      # gloo has no AVG: sum, then scale
        # buckets: consecutive parameter ranges, cut from the tail of the parameter list (first to be ready)
        limit = int(bucket_mb * (1 << 20) / self.params[0].element_size())
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

    def bucket_grads(self, bi: int):
        lo, hi = self.buckets[bi]
        return [p.grad for p in self.params[lo:hi] if p.grad is not None]

    def zero_grad(self):
        """Start of an iteration: gradients are dropped (autograd then hands over its own buffers without an
        accumulate pass) and the arrival counters reset."""
        for p in self.params:
            p.grad = None
        self._arrived = [0] * len(self.buckets)
        self._works = []

    def _reduce(self, tensors, async_op: bool):
        if not tensors:
            return None
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        with dist._coalescing_manager(group=self.group, async_ops=async_op) as cm:
            for t in tensors:
                dist.all_reduce(t, op=op, group=self.group)
        return cm if async_op else None

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
                self._works.append((bi, self._reduce(self.bucket_grads(bi), async_op=True)))
        return hook

    def finish(self):
        """After backward: the compute stream waits for every bucket (any bucket whose hook did not
        fire -- e.g. hooks not installed -- is reduced here)."""
        if self.world_size == 1 or self.suspended:
            return
        started = {bi for bi, _ in self._works}
        for bi in range(len(self.buckets)):
            if bi not in started:
                self._works.append((bi, self._reduce(self.bucket_grads(bi), async_op=True)))
        for bi, work in self._works:
            if work is not None:
                work.wait()
            if not self._avg:
                grads = self.bucket_grads(bi)
                if grads:
                    torch._foreach_div_(grads, float(self.world_size))
        self._works = []
        self._arrived = [0] * len(self.buckets)

    def average_gradients(self):
        """Blocking form (no overlap): one coalesced all-reduce per bucket after backward."""
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
