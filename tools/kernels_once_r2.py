#!/usr/bin/env python
"""Launch the round-2 kernels of interest a few times at their bench shapes so one `ncu --set full` capture stays
short: label propagation (tcgen05, stress + CRAG shapes), upsample_sum (bf16, 400-px tile), SLIC (464^2, batch 1 and 4),
the footprint pooling kernels (464^2, 4224 channels).   ncu -k regex:... python tools/kernels_once_r2.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import bench  # noqa: E402
from wesup_b200 import ops, synth  # noqa: E402
from wesup_b200.ops import SuperpixelMaps  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    lib = ops._lib.load()
    for n, n_l in ((8000, 4000), (11460, 229)):
        f = (torch.randn(n, 32, device=dev) * 0.06).abs()
        y_l = torch.zeros(n_l, 2, device=dev); y_l[:, 0] = 1
        for _ in range(2):
            ops.label_propagate(f, y_l, 0.8, algo="tc")
    h = w = 400
    terms = [torch.randn(hh, ww, 1024, device=dev).to(torch.bfloat16) for hh, ww in ((400, 400), (200, 200), (100, 100), (50, 50), (25, 25))]
    bias = torch.randn(1024, device=dev)
    for _ in range(2):
        ops.upsample_sum(terms, (h, w), bias=bias, relu=True)
    for b in (1, 4):
        xs = torch.stack([synth.sample(464, 464, index=i)[0][0] for i in range(b)]).to(dev)
        for _ in range(2):
            ops.slic_batch(xs, 1076, 40)
    img, _, pm = synth.sample(464, 464, index=0)
    labels, n = ops.slic(img.to(dev), 1076, 40)
    sp = SuperpixelMaps.from_labels(labels, pm[0].to(dev), n_sp=int(n.item()))
    out = {}
    bench.QUICK = True
    bench.pooling_kernels(dev, 464, 464, sp, bench.L2Flush(dev), lib, [2 * c for c in bench.VGG_C], "backbone4224", out)
    torch.cuda.synchronize()
    print("done")
