#!/usr/bin/env python
"""Time label propagation (exact CUDA-core kernel vs tcgen05 filter kernel) at the realistic and stress shapes."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import bench  # noqa: E402
from wesup_b200 import ops  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    out = {}
    bench.label_propagation_kernels(dev, bench.L2Flush(dev), ops._lib.load(),
                                    (("realistic_464", 1087, 21), ("glas", 2022, 40), ("crag", 11460, 229), ("mid", 2000, 1000), ("stress", 8000, 4000)), out)
    for k, v in out.items():
        print(k, json.dumps({a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}), flush=True)
