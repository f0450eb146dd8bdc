#!/usr/bin/env python
"""Tiled inference throughput on the synthetic whole-slide image: moved into bench.py so that the driver sees it.

    python bench.py --workload tiles_sp    [--slide 20000] [--patch 400] [--tile-batch 16] [--gpus N]
    python bench.py --workload tiles_pixel [--hc-dtype bf16|fp32]
"""
import runpy
import sys
from pathlib import Path

if __name__ == "__main__":
    root = Path(__file__).resolve().parents[1]
    if not any(a.startswith("--workload") for a in sys.argv[1:]):
        sys.argv[1:1] = ["--workload", "tiles_sp"]
    sys.argv[0] = str(root / "bench.py")
    runpy.run_path(str(root / "bench.py"), run_name="__main__")
