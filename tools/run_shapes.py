#!/usr/bin/env python
"""Training iterations at the BASELINE.json image shapes (464^2, GlaS 522x775, CRAG 1516x1512): a few eager and
CUDA-graph iterations each, device time per image and peak memory.  python tools/run_shapes.py"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from wesup_b200 import synth  # noqa: E402
from wesup_b200.models import initialize_trainer  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for name, (h, w) in (("464x464", (464, 464)), ("glas_522x775", (522, 775)), ("crag_1516x1512", (1516, 1512))):
        torch.manual_seed(0)
        trainer = initialize_trainer("wesup", device=dev, pretrained=False, materialize_hypercolumn=False, cuda_graph=True)
        trainer.optimizer, _ = trainer.get_default_optimizer()
        data = [tuple(t.to(dev) for t in synth.sample(h, w, index=i)) for i in range(3)]
        torch.cuda.reset_peak_memory_stats(dev)
        for i in range(4):                                   # eager iterations + capture
            trainer.train_one_iteration("train", *data[i % 3])
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 6
        s.record()
        for i in range(n):
            trainer.prefetch(*data[(i + 1) % 3])
            trainer.train_one_iteration("train", *data[i % 3])
        trainer.flush_metrics()
        e.record()
        torch.cuda.synchronize()
        hist = trainer.tracker.history
        out[name] = {"ms_per_image": s.elapsed_time(e) / n, "img_per_s": n / (s.elapsed_time(e) / 1e3),
                     "last_loss": float(hist["loss"][-1]), "graphs": len(getattr(trainer, "_graphs", {})),
                     "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
        del trainer, data
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
