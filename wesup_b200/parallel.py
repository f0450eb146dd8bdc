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
    unused-parameter handling is needed.

    A bucket's gradients are gathered into ONE persistent flat tensor (a multi-tensor copy, 75 MB per iteration in
    total = 25 us of HBM time) and reduced there with a single all-reduce; afterwards the parameters' `.grad` are views
    of the flat tensor, so nothing is copied back.  Measured on 2 x B200 (kernel timeline, gpurun_out/t8): reduced
    in place, tensor by tensor in a coalesced launch, NCCL treats every tensor as its own operation and picks the
    low-latency protocol for each (`AllReduce_Sum_f32_RING_LL`, 147 GB/s effective, 0.5 ms of NCCL kernels per
    iteration competing with backward for SMs); one 10 MB message per bucket takes the bandwidth protocol.
    `flat_numel` is the size below which a gradient is staged (default: all of them)."""

    def __init__(self, model: torch.nn.Module, process_group=None, bucket_mb: float = 40.0, flat_numel: int = 1 << 62,
                 head_mb: float = 8.0):
        small_numel = flat_numel
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        backend = dist.get_backend(process_group) if dist.is_initialized() else None
        self._avg = backend == "nccl"                      # gloo has no AVG: sum, then scale
        # buckets: consecutive parameter ranges, cut from the tail of the parameter list (first to be ready).  The HEAD of
        # the list (first backbone convolutions) gets a small bucket of its own: its gradients are the last to arrive and
        # nothing overlaps their reduction.  Few large buckets otherwise: measured on 2 x B200 with one-element
        # all-reduces, every collective of an iteration costs ~0.03 ms whatever it carries (the ranks meet in it and the
        # waiting NCCL blocks hold SMs), 8 buckets 0.87 ms per 4-image step against 0.48 ms for 2.
        esize = self.params[0].element_size()
        limit = int(bucket_mb * (1 << 20) / esize)
        head_limit, head_hi, size = int(head_mb * (1 << 20) / esize), 0, 0
        while head_hi < len(self.params) - 1 and size + self.params[head_hi].numel() <= head_limit:
            size += self.params[head_hi].numel()
            head_hi += 1
        self.buckets = []                                  # (first param index, last param index + 1), tail first
        hi = len(self.params)
        while hi > head_hi:
            lo, size = hi, 0
            while lo > head_hi and (size == 0 or size + self.params[lo - 1].numel() <= limit):
                lo -= 1
                size += self.params[lo].numel()
            self.buckets.append((lo, hi))
            hi = lo
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
        # per bucket: indices of the small parameters and their flat staging tensor (persistent: CUDA-graph safe)
        self._small = []
        for lo, hi in self.buckets:
            idx = [i for i in range(lo, hi) if self.params[i].numel() < small_numel]
            n = sum(self.params[i].numel() for i in idx)
            flat = torch.zeros(n, dtype=self.params[lo].dtype, device=self.params[lo].device) if len(idx) > 1 else None
            views, off = [], 0
            if flat is not None:
                for i in idx:
                    # same strides as the parameter (channels_last convolution weights): the fused optimizer kernel
                    # wants parameter, gradient and momentum laid out alike, and the copy in stays a plain memcpy
                    views.append(flat[off:off + self.params[i].numel()].as_strided(self.params[i].size(), self.params[i].stride()))
                    off += self.params[i].numel()
            self._small.append((idx if flat is not None else [], flat, views))
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
        if os.environ.get("WESUP_DP_DEBUG_TINY"):          # timing diagnosis only (wrong gradients): one element per tensor,
            tensors = [t.view(-1)[:1] for t in tensors]    # i.e. the ranks' lockstep without the bytes
        with dist._coalescing_manager(group=self.group, async_ops=async_op) as cm:
            for t in tensors:
                dist.all_reduce(t, op=op, group=self.group)
        return cm if async_op else None

    def _reduce_bucket(self, bi: int, async_op: bool = True):
        """Start the bucket's all-reduce: large gradients where they are, the small ones through the flat tensor."""
        lo, hi = self.buckets[bi]
        idx, flat, views = self._small[bi]
        staged = set(idx)
        tensors = [self.params[i].grad for i in range(lo, hi) if i not in staged and self.params[i].grad is not None]
        if flat is not None:
            srcs = [self.params[i].grad for i in idx]
            if all(g is not None for g in srcs):
                torch._foreach_copy_(views, srcs)
                tensors.append(flat)
            else:                                          # a parameter without a gradient: fall back to in-place
                tensors += [g for g in srcs if g is not None]
                staged = set()
        return (bi, self._reduce(tensors, async_op), bool(staged))

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
                self._works.append(self._reduce_bucket(bi))
        return hook

    def finish(self):
        """After backward: the compute stream waits for every bucket (any bucket whose hook did not
        fire -- e.g. hooks not installed -- is reduced here)."""
        if self.world_size == 1 or self.suspended:
            return
        started = {bi for bi, _, _ in self._works}
        for bi in range(len(self.buckets)):
            if bi not in started:
                self._works.append(self._reduce_bucket(bi))
        for bi, work, staged in self._works:
            if work is not None:
                work.wait()
            if staged:                                     # the reduced gradients stay in the flat tensor
                idx, _flat, views = self._small[bi]
                for i, v in zip(idx, views):
                    self.params[i].grad = v
            if not self._avg:
                grads = self.bucket_grads(bi)
                if grads:
                    torch._foreach_div_(grads, float(self.world_size))
        self._works = []
        self._arrived = [0] * len(self.buckets)

    def probe(self):
        """0-dim tensor that is non-finite iff the reduced gradient of the head bucket is (any rank's NaN ends up in
        every rank's average): what the trainer's device-side step guard looks at."""
        lo, hi = self.buckets[-1]
        _idx, flat, _views = self._small[-1]
        if flat is not None:
            return flat.sum()
        return torch.stack([p.grad.sum() for p in self.params[lo:hi] if p.grad is not None]).sum()

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
