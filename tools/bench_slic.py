#!/usr/bin/env python
"""Time GPU SLIC (wesup_slic_batch through the C ABI, buffers preallocated) at H x W for several batch sizes.

    python tools/bench_slic.py [H [W [reps]]]          one JSON line per batch size
"""
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
    lib = ops._lib.load()
    n_seg = int(H * W / 200)
    peak, _ = bench.peaks()
    for B in ((1, 2, 4, 8) if H * W <= 1 << 20 else (1, 2)):
        x = torch.stack([synth.sample(H, W, index=i)[0][0] for i in range(B)]).to(dev).contiguous()
        ws = torch.empty(lib.wesup_slic_batch_workspace_bytes(B, H, W, n_seg), dtype=torch.uint8, device=dev)
        labels = torch.empty((B, H, W), dtype=torch.int32, device=dev)
        n = torch.zeros(B, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        fn = lambda: ops.check(lib.wesup_slic_batch(x.data_ptr(), 0, B, H, W, n_seg, 40.0, 10, 1, labels.data_ptr(),  # noqa: E731
                                                    n.data_ptr(), ws.data_ptr(), st), "slic")
        flush = bench.L2Flush(dev)
        ms = bench.time_kernel(fn, reps, flush)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record(); e.synchronize()
        b2b = s.elapsed_time(e) / reps
        print(json.dumps({"H": H, "W": W, "batch": B, "n_labels": n.tolist(), "ms_isolated": round(ms, 4),
                          "ms_back_to_back": round(b2b, 4), "ms_per_image": round(b2b / B, 4), "bytes_per_px": 360,
                          "gbs": round(B * H * W * 360 / b2b / 1e6, 1), "frac_of_hbm_peak": round(B * H * W * 360 / b2b / 1e6 / peak, 4)}),
              flush=True)
