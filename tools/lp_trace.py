#!/usr/bin/env python
"""Per-role timeline of CTA (0,0) of label_propagate_tc_kernel (library built with -DWESUP_TC_TRACE): SM clock stamps."""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from wesup_b200 import ops
lib = ctypes.CDLL(str(Path(__file__).resolve().parents[1] / "wesup_b200" / "libwesup_b200.so"))
dev = torch.device("cuda", 0)
n, n_l = int(sys.argv[1]) if len(sys.argv) > 1 else 8000, int(sys.argv[2]) if len(sys.argv) > 2 else 4000
f = (torch.randn(n, 32, device=dev) * 0.06).abs()
y_l = torch.zeros(n_l, 2, device=dev); y_l[:, 0] = 1
for _ in range(3):
    ops.label_propagate(f, y_l, 0.8, algo="tc")
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
lib.wesup_debug_tc_trace(buf)
t0 = buf[0]
rel = lambda i: (buf[i] - t0) / 1.965e3 if buf[i] else None   # us at 1965 MHz
print("start 0; producer past griddep wait", rel(1), "| epilogue loop end", rel(2), "| drain end", rel(3), "| ticket", rel(4))
for t in range(10):
    if not buf[32 + t]:
        break
    print(f"tile {t}: load issued {rel(32+t):7.2f}  full seen {rel(64+t):7.2f}  mma issued {rel(96+t):7.2f}  tfull seen {rel(128+t):7.2f}  pass1 {rel(160+t):7.2f}  released {rel(192+t):7.2f}")
