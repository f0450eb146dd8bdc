"""Trainer scaffolding with the reference's interface
(/root/reference/models/base.py:16-360): `BaseConfig.to_dict`, and a
`BaseTrainer` whose template methods (`preprocess`, `compute_loss`,
`postprocess`, `post_epoch_hook`, `get_default_dataset/optimizer`) and loop
(`train`, `train_one_epoch`, `train_one_iteration`, `evaluate`,
`load/save_checkpoint`) behave the same way, so train.py-style drivers work
unchanged.  The loop is host bookkeeping, not a kernel target; the one additive
feature is data parallelism (`wesup_b200.parallel.GradientAllReduce`): when a
process group is initialised, gradients are averaged over ranks between
`backward()` and `optimizer.step()`.
"""
from __future__ import annotations

import logging
import os
import time
from abc import ABC, abstractmethod
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

from ..utils import record, underline
from ..utils.history import HistoryTracker


class BaseConfig:
    batch_size = 1
    epochs = 10
    epsilon = 1e-7

    def to_dict(self):
        return {name: getattr(self, name) for name in dir(self)
                if not name.startswith("_") and name != "to_dict"}

    def __str__(self):
        return "\n".join(f"{k:<32s}{v}" for k, v in self.to_dict().items())


class BaseTrainer(ABC):
    def __init__(self, model, **kwargs):
        self.device = kwargs.get("device", "cuda" if torch.cuda.is_available() else "cpu")
        self.model = model.to(self.device)
        self.kwargs = kwargs
        self.logger = kwargs.get("logger")
        if not self.logger:
            self.logger = logging.getLogger("Train")
            self.logger.setLevel(logging.DEBUG)
            if not self.logger.handlers:
                self.logger.addHandler(logging.StreamHandler())
        self.initial_epoch = 1
        self.record_dir = None
        self.tracker = HistoryTracker()
        self.dataloaders = None
        self.optimizer, self.scheduler = None, None
        self.metric_funcs = []
        self.grad_sync = None            # set by enable_data_parallel()

    # ---- template methods ---------------------------------------------------
    @abstractmethod
    def get_default_dataset(self, root_dir, train=True, proportion=1.0):
        ...

    def get_default_optimizer(self):
        return torch.optim.SGD(self.model.parameters(), lr=1e-3), None

    def preprocess(self, *data):
        return [datum.to(self.device) for datum in data]

    def prefetch(self, *data):
        """Hook: start preprocessing `data` ahead of time (no-op by default)."""

    @abstractmethod
    def compute_loss(self, pred, target, metrics=None):
        ...

    def postprocess(self, pred, target=None):
        return pred if target is None else (pred, target)

    def post_epoch_hook(self, epoch):
        pass

    # ---- data parallelism (additive) ---------------------------------------
    def enable_data_parallel(self, process_group=None, overlap=True, bucket_mb=40.0):
        """Average gradients over the ranks of `process_group` every iteration: bucketed all-reduces
        started from autograd hooks while backward is still running (`overlap=False`: after backward)."""
        from ..parallel import GradientAllReduce
        self.grad_sync = GradientAllReduce(self.model, process_group, bucket_mb=bucket_mb)
        self.grad_sync.broadcast_parameters()
        if overlap:
            self.grad_sync.enable_overlap()
        return self.grad_sync

    # ---- checkpoints (key names are part of the surface) --------------------
    def load_checkpoint(self, ckpt_path=None):
        if ckpt_path is None:
            self.record_dir = Path(record.prepare_record_dir())
            return
        self.record_dir = Path(ckpt_path).parent.parent
        self.logger.info(f"Loading checkpoint from {ckpt_path}.")
        ckpt = torch.load(ckpt_path, map_location=self.device)
        self.initial_epoch = ckpt["epoch"] + 1
        self.model.load_state_dict(ckpt["model_state_dict"])
        if self.optimizer is not None and "optimizer_state_dict" in ckpt:
            self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if self.scheduler is not None and "scheduler_state_dict" in ckpt:
            self.scheduler.load_state_dict(ckpt["scheduler_state_dict"])

    def save_checkpoint(self, ckpt_path, **extra):
        ckpt = {"model_state_dict": self.model.state_dict(),
                "optimizer_state_dict": self.optimizer.state_dict(), **extra}
        if self.scheduler is not None:
            ckpt["scheduler_state_dict"] = self.scheduler.state_dict()
        torch.save(ckpt, ckpt_path)

    # ---- the loop -------------------------------------------------------------
    def train_one_iteration(self, phase, *data):
        """Same sequence as the reference (models/base.py:184-211): preprocess, zero_grad,
        forward, loss, NaN check, backward, step, postprocess, metrics.  Every scalar of the
        iteration (loss, loss-side metrics, accuracy/dice) stays on the device and travels to
        the host in ONE asynchronous transfer, which the host reads `metrics_lag` iterations
        later (default 1: iteration k's numbers are read after iteration k+1 has been
        enqueued), so the GPU never idles behind a `.item()`.  A NaN loss raises
        ValueError('Loss is nan!') when its scalars are read -- uncaught by the epoch loop, as
        in the reference; `metrics_lag=0` restores the reference's same-iteration check."""
        input_, target = self.preprocess(*data)
        self._run_iteration(phase, input_, target)

    def arm_step_guard(self, loss):
        """Device-side guard of the parameter update: with lagged metrics the host sees a NaN loss one iteration late
        (and inside a CUDA graph the SGD step is baked in), so the fused optimizer kernel is told ON THE DEVICE to skip
        the step when the loss -- or, with data parallelism, the averaged gradient of the last bucket, which a NaN on
        any rank poisons -- is not finite (`found_inf`, the switch torch's GradScaler uses).  The weights and the
        momentum stay intact and `ValueError('Loss is nan!')` still fires when the scalars reach the host.  Only the
        fused optimizer honours it; the plain one keeps the reference's behaviour (use metrics_lag=0 there)."""
        opt = self.optimizer
        if opt is None or not loss.is_cuda or not opt.defaults.get("fused", False):
            return
        if getattr(self, "_found_inf", None) is None or self._found_inf.device != loss.device:
            self._found_inf = torch.zeros((), dtype=torch.float32, device=loss.device)
        opt.found_inf = self._found_inf
        bad = ~torch.isfinite(loss.detach())
        if self.grad_sync is not None and self.grad_sync.world_size > 1:
            bad = bad | ~torch.isfinite(self.grad_sync.probe())
        self._found_inf.copy_(bad.float())

    def _run_iteration(self, phase, input_, target):
        """Everything of `train_one_iteration` after `preprocess`."""
        if self.grad_sync is not None:
            self.grad_sync.zero_grad()       # drops the gradients and resets the bucket arrival counters
        else:
            self.optimizer.zero_grad()
        metrics = {}
        self._defer_scalars = True
        try:
            with torch.set_grad_enabled(phase == "train"):
                pred = self.model(input_)
                if phase == "train":
                    loss = self.compute_loss(pred, target, metrics=metrics)
                    metrics["loss"] = loss.detach()
                    loss.backward()
                    if self.grad_sync is not None:
                        self.grad_sync.finish()          # buckets were started from the backward hooks; wait for them
                    self.arm_step_guard(loss)
                    self.optimizer.step()
            pred, target = self.postprocess(pred, target)
            metrics.update(self.evaluate(pred, target))
        finally:
            self._defer_scalars = False
        self._submit_scalars(metrics, phase)
        self.flush_metrics(keep=max(int(self.kwargs.get("metrics_lag", 1) or 0), 0))

    def _submit_scalars(self, metrics, phase):
        """Start the single device-to-host transfer of this iteration's 0-dim tensors."""
        keys = [k for k, v in metrics.items() if torch.is_tensor(v)]
        host = event = None
        if keys:
            stacked = torch.stack([metrics[k].detach().float().reshape(()) for k in keys])
            if stacked.is_cuda:
                host = torch.empty(len(keys), dtype=torch.float32, pin_memory=True)
                host.copy_(stacked, non_blocking=True)
                event = torch.cuda.Event()
                event.record(torch.cuda.current_stream(stacked.device))
            else:
                host = stacked
        if not hasattr(self, "_pending_scalars"):
            self._pending_scalars = []
        self._pending_scalars.append((metrics, keys, host, event, phase))

    def flush_metrics(self, keep=0):
        """Read the scalars of all but the `keep` most recent iterations and hand them to the
        tracker (in iteration order).  Called with keep=0 at the end of every phase."""
        pending = getattr(self, "_pending_scalars", None)
        while pending and len(pending) > keep:
            metrics, keys, host, event, phase = pending.pop(0)
            if event is not None:
                event.synchronize()
            if keys:
                metrics = {**metrics, **dict(zip(keys, host.tolist()))}
            if phase == "train" and metrics["loss"] != metrics["loss"]:
                pending.clear()
                raise ValueError("Loss is nan!")
            self.tracker.step(metrics)

    @staticmethod
    def _read_scalars(metrics):
        """Replace every 0-dim device tensor in `metrics` by its float with one transfer."""
        keys = [k for k, v in metrics.items() if torch.is_tensor(v)]
        if keys:
            vals = torch.stack([metrics[k].detach().float().reshape(()) for k in keys]).tolist()
            metrics = {**metrics, **dict(zip(keys, vals))}
        return metrics

    def train_one_epoch(self, no_val=False):
        for phase in (["train"] if no_val else ["train", "val"]):
            self.logger.info(f"{phase.capitalize()} phase:")
            start = time.time()
            if phase == "train":
                self.model.train()
                self.tracker.train()
            else:
                self.model.eval()
                self.tracker.eval()
            batches = iter(self.dataloaders[phase])
            data = next(batches, None)
            while data is not None:
                upcoming = next(batches, None)
                try:
                    if upcoming is not None and self.kwargs.get("prefetch", True):
                        self.prefetch(*upcoming)         # next image's preprocessing overlaps this iteration
                    self.train_one_iteration(phase, *data)
                except RuntimeError as ex:       # same policy as the reference (base.py:234-237)
                    self.logger.exception(ex)
                    if self.grad_sync is not None and self.grad_sync.world_size > 1:
                        # a rank that skips an iteration skips its collectives and the other ranks would block in
                        # theirs: with data parallelism the error ends the job instead of being swallowed
                        raise
                data = upcoming
            self.flush_metrics()
            self.logger.info(f"Took {time.time() - start:.2f}s.")
            self.logger.info(self.tracker.log())

    def train(self, data_root, **kwargs):
        self.kwargs = {**self.kwargs, **kwargs}
        self.optimizer, self.scheduler = self.get_default_optimizer()
        self.load_checkpoint(self.kwargs.get("checkpoint"))
        self.logger.addHandler(logging.FileHandler(self.record_dir / "train.log"))
        plain = {k: v for k, v in self.kwargs.items() if isinstance(v, (int, float, str, tuple))}
        record.save_params(self.record_dir, plain)
        self.logger.info(str(plain) + "\n")
        self.tracker.save_path = self.record_dir / "history.csv"
        data_root = Path(data_root)
        train_path, val_path = data_root / "train", data_root / "val"
        train_dataset = self.get_default_dataset(train_path, proportion=self.kwargs.get("proportion", 1))
        if hasattr(train_dataset, "summary"):
            train_dataset.summary(logger=self.logger)
        workers = self.kwargs.get("num_workers", os.cpu_count())
        sampler = None
        if self.grad_sync is not None:
            sampler = torch.utils.data.distributed.DistributedSampler(
                train_dataset, num_replicas=self.grad_sync.world_size, rank=self.grad_sync.rank, shuffle=True)
        self.dataloaders = {"train": torch.utils.data.DataLoader(
            train_dataset, batch_size=self.kwargs.get("batch_size"), shuffle=sampler is None,
            sampler=sampler, num_workers=workers)}
        if val_path.exists():
            val_dataset = self.get_default_dataset(val_path, train=False)
            self.dataloaders["val"] = torch.utils.data.DataLoader(val_dataset, batch_size=1, num_workers=workers)
        self.logger.info(underline("\nTraining Stage", "="))
        self.metric_funcs = self.kwargs.get("metrics") or []
        total_epochs = self.kwargs.get("epochs") + self.initial_epoch - 1
        for epoch in range(self.initial_epoch, total_epochs + 1):
            self.logger.info(underline(f"\nEpoch {epoch}/{total_epochs}", "-"))
            if sampler is not None:
                sampler.set_epoch(epoch)
            self.tracker.start_new_epoch(self.optimizer.param_groups[0]["lr"])
            self.train_one_epoch(no_val=not val_path.exists())
            self.post_epoch_hook(epoch)
            if self.grad_sync is None or self.grad_sync.rank == 0:
                self.tracker.save()
                ckpt_dir = self.record_dir / "checkpoints"
                ckpt_dir.mkdir(exist_ok=True)
                self.save_checkpoint(ckpt_dir / f"ckpt.{epoch:04d}.pth", epoch=epoch)
                for old in sorted(ckpt_dir.glob("*.pth"))[:-1]:
                    os.remove(old)
        if self.grad_sync is None or self.grad_sync.rank == 0:
            self.logger.info(self.tracker.report())

    def evaluate(self, pred, target=None, verbose=False):
        if target is None:
            return {}
        lazy = getattr(self, "_defer_scalars", False)
        scores = defaultdict(list)
        for P, G in zip(pred, target):
            for func in self.metric_funcs:
                deferred = getattr(func, "deferred", None) if lazy else None
                scores[func.__name__].append(deferred(P, G) if deferred is not None else func(P, G))
        out = {}
        for k, v in scores.items():
            out[k] = torch.stack(v).mean() if torch.is_tensor(v[0]) else np.mean(v)
        return out
