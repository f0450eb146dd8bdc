#!/usr/bin/env python
"""Time wesup_upsample_sum (first pixel-MLP layer without the hypercolumn) at a 400-px tile, C = 1024."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import bench  # noqa: E402
from wesup_b200 import ops  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    h = w = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    c = 1024
    peak, _ = bench.peaks()
    flush = bench.L2Flush(dev)
    import os
    for dtype, es, v4 in ((torch.bfloat16, 2, "0"), (torch.bfloat16, 2, "1"), (torch.float32, 4, "0")):
        os.environ["WESUP_UPS_V4"] = v4
        sizes = [(h, w), (h // 2, w // 2), (h // 4, w // 4), (h // 8, w // 8), (h // 16, w // 16)]
        terms = [torch.randn(hh, ww, c, device=dev).to(dtype) for hh, ww in sizes]
        bias = torch.randn(c, device=dev)
        out = torch.empty(h * w, c, dtype=dtype, device=dev)
        ms = bench.time_kernel(lambda: ops.upsample_sum(terms, (h, w), bias=bias, relu=True, out=out), 10, flush)
        b = sum(t.numel() for t in terms) * es + out.numel() * es
        print(json.dumps({"kernel": "upsample_sum", "dtype": str(dtype), "channels_per_thread": 4 if (v4 == "1" or es == 4) else 8, "H": h, "W": w, "C": c, "ms": round(ms, 4), "algorithmic_bytes": b,
                          "gbs": round(b / ms / 1e6, 1), "frac_of_hbm_peak": round(b / ms / 1e6 / peak, 3)}), flush=True)
