#!/usr/bin/env python
"""Time the hypercolumn forward kernel alone (fp32 + bf16) at H x W; used for A/B
runs of kernel variants selected by WESUP_HC_FWD / WESUP_HC_SEG."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import bench  # noqa: E402
from wesup_b200 import ops  # noqa: E402

if __name__ == "__main__":
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 464
    W = int(sys.argv[2]) if len(sys.argv) > 2 else H
    dev = torch.device("cuda", 0)
    flush = bench.L2Flush(dev)
    g = torch.Generator().manual_seed(0)
    sides = [torch.randn(1, c, H >> s, W >> s, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
             for c, s in zip(bench.VGG_C, bench.VGG_SHIFT)]
    side_bytes = sum(s.numel() * 4 for s in sides)
    lib = ops._lib.load()
    mem = [s.permute(0, 2, 3, 1).contiguous() for s in sides]
    ptrs = ops._lib.ptr_array([m.data_ptr() for m in mem])
    ia = ops._lib.int_array
    Cs, hs, ws = [s.size(1) for s in sides], [s.size(2) for s in sides], [s.size(3) for s in sides]
    st = torch.cuda.current_stream().cuda_stream
    res = {"variant": os.environ.get("WESUP_HC_FWD", "default"), "seg": os.environ.get("WESUP_HC_SEG", "default"), "H": H, "W": W}
    ref = None
    for dtype, es, code, tag in ((torch.float32, 4, 0, "f32"), (torch.bfloat16, 2, 1, "bf16")):
        out = torch.empty((H * W, 2112), dtype=dtype, device=dev)
        fn = lambda: ops.check(lib.wesup_hypercolumn_fwd(ptrs, ia(Cs), ia(hs), ia(ws), 13, H, W, out.data_ptr(), code, 1, st), "hc")
        ms = bench.time_kernel(fn, 20, flush)
        b = side_bytes + 2112 * H * W * es
        res[tag] = {"ms": round(ms, 4), "gbs": round(b / ms / 1e6, 1), "frac": round(b / ms / 1e6 / bench.peaks()[0], 4)}
        res[tag + "_checksum"] = float(out.float().double().sum())
    print(json.dumps(res))
