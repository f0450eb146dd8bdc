"""Per-iteration metric bookkeeping with the interface the trainer loop uses
(/root/reference/utils/history.py:11-64).  Host-only; not on the hot path."""
from __future__ import annotations

import csv
import os
from collections import defaultdict


class HistoryTracker:
    def __init__(self, save_path=None):
        self.history = defaultdict(list)
        self.learning_rate = None
        self.save_path = save_path
        self.is_train = True

    def start_new_epoch(self, lr):
        self.history.clear()
        self.learning_rate = lr

    def train(self):
        self.is_train = True

    def eval(self):
        self.is_train = False

    def step(self, metrics):
        prefix = "" if self.is_train else "val_"
        for key, value in metrics.items():
            self.history[prefix + key].append(value)
        return ", ".join(f"{prefix}{k} = {v:.4f}" for k, v in metrics.items())

    def _means(self):
        return {k: (sum(v) / len(v) if v else 0) for k, v in sorted(self.history.items())}

    def log(self):
        wanted = {k: v for k, v in self._means().items() if k.startswith("val_") != self.is_train}
        return ", ".join(f"average {k} = {v:.4f}" for k, v in wanted.items()).capitalize()

    def save(self):
        if self.save_path is None:
            raise RuntimeError("cannot save history without setting save_path.")
        means = self._means()
        fresh = not os.path.exists(self.save_path)
        with open(self.save_path, "a", newline="") as fp:
            writer = csv.writer(fp)
            if fresh:
                writer.writerow(list(means) + ["lr"])
            writer.writerow(list(means.values()) + [self.learning_rate])

    def report(self, last_n_epochs=5):
        with open(self.save_path) as fp:
            rows = list(csv.DictReader(fp))[-last_n_epochs:]
        lines = []
        for key in (rows[0].keys() if rows else []):
            if key in ("lr", "loss", "val_loss"):
                continue
            vals = [float(r[key]) for r in rows if r.get(key) not in (None, "")]
            if vals:
                lines.append(f"{key:20s} {sum(vals) / len(vals):.4f}")
        return "\nTraining Summary (Avg over last 5 epochs)\n" + "=" * 41 + "\n" + "\n".join(lines)
