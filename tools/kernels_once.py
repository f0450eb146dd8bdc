#!/usr/bin/env python
"""Launch every superpixel-stage kernel once at the bench shape (464x464) so an
`ncu --set full` capture stays short:  WESUP_BENCH_QUICK=1 python tools/kernels_once.py"""
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("WESUP_BENCH_QUICK", "1")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import bench  # noqa: E402

if __name__ == "__main__":
    peak, _ = bench.peaks()
    print(json.dumps(bench.kernel_rooflines(torch.device("cuda", 0), peak)))
