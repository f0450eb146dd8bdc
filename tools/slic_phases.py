#!/usr/bin/env python
"""Where GPU SLIC spends its time: wesup_slic_batch timed for max_iter in {0,1,2,5,10} with and without connectivity."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from wesup_b200 import ops, synth  # noqa: E402

if __name__ == "__main__":
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 464
    W = int(sys.argv[2]) if len(sys.argv) > 2 else H
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    dev = torch.device("cuda", 0)
    lib = ops._lib.load()
    n_seg = int(H * W / 200)
    x = torch.stack([synth.sample(H, W, index=i)[0][0] for i in range(B)]).to(dev).contiguous()
    ws = torch.empty(lib.wesup_slic_batch_workspace_bytes(B, H, W, n_seg), dtype=torch.uint8, device=dev)
    labels = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    n = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for enforce in (0, 1):
        for it in (1, 10):
            fn = lambda: ops.check(lib.wesup_slic_batch(x.data_ptr(), 0, B, H, W, n_seg, 40.0, it, enforce, labels.data_ptr(),  # noqa: E731
                                                        n.data_ptr(), ws.data_ptr(), st), "slic")
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                fn()
            e.record(); e.synchronize()
            out[f"enforce{enforce}_iter{it}_us"] = round(1e3 * s.elapsed_time(e) / 20, 1)
    import ctypes
    fn = lambda: ops.check(lib.wesup_slic_batch(x.data_ptr(), 0, B, H, W, n_seg, 40.0, 10, 1, labels.data_ptr(),  # noqa: E731
                                                n.data_ptr(), ws.data_ptr(), st), "slic")
    fn()
    t = (ctypes.c_ulonglong * 64)()
    ops.check(lib.wesup_slic_debug_times(ws.data_ptr(), B, H, W, n_seg, t), "times")
    km = [t[31]] + [t[i] for i in range(1, 20)]
    cc = [t[63]] + [t[32 + i] for i in range(1, 6)] + [t[62]]
    out["kmeans_phase_us"] = [round((b - a) / 1e3, 1) for a, b in zip(km, km[1:])]
    out["connect_phase_us"] = dict(zip(["C1_local", "C2_borders", "C3_flatten", "C4_split+count", "C5_scan+C6_small", "C7_relabel"],
                                       [round((b - a) / 1e3, 1) for a, b in zip(cc, cc[1:])]))
    out["kmeans_total_us"] = round((t[19] - t[31]) / 1e3, 1)
    out["connect_total_us"] = round((t[62] - t[63]) / 1e3, 1)
    out["kmeans_last_barrier_to_connect_start_us"] = round((t[63] - t[19]) / 1e3, 1)
    print(json.dumps({"H": H, "W": W, "B": B, **out}))
