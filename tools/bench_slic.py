#!/usr/bin/env python
"""Time GPU SLIC (wesup_slic through the C ABI, buffers preallocated) at H x W."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import bench  # noqa: E402
from wesup_b200 import ops, synth  # noqa: E402

if __name__ == "__main__":
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 464
    W = int(sys.argv[2]) if len(sys.argv) > 2 else H
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda", 0)
    img, _, _ = synth.sample(H, W, index=0)
    x = img[0].to(dev).contiguous()
    lib = ops._lib.load()
    n_seg = int(H * W / 200)
    ws = torch.empty(lib.wesup_slic_workspace_bytes(H, W, n_seg), dtype=torch.uint8, device=dev)
    labels = torch.empty((H, W), dtype=torch.int32, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: ops.check(lib.wesup_slic(x.data_ptr(), 0, H, W, n_seg, 40.0, 10, 1, labels.data_ptr(), n.data_ptr(), ws.data_ptr(), st), "slic")
    flush = bench.L2Flush(dev)
    ms = bench.time_kernel(fn, reps, flush)
    # back-to-back (no flush, launch-overhead hidden by queueing)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); e.synchronize()
    print(json.dumps({"H": H, "W": W, "n_labels": int(n.item()), "ms_isolated": round(ms, 4),
                      "ms_back_to_back": round(s.elapsed_time(e) / reps, 4), "bytes_per_px": 360,
                      "gbs_back_to_back": round(H * W * 360 / (s.elapsed_time(e) / reps) / 1e6, 1)}))
