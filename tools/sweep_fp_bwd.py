#!/usr/bin/env python
"""A/B of the backward pooling launch shapes on one box (WESUP_FP_X experiment knobs of wesup_levels_pool_bwd_fp)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import bench  # noqa: E402
from wesup_b200 import _lib, ops, synth  # noqa: E402
from wesup_b200.ops import SuperpixelMaps  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    H = W = bench.H
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    flush = bench.L2Flush(dev)
    img, _, point_mask = synth.sample(H, W, index=0)
    labels, n = ops.slic(img.to(dev), int(H * W / 200), 40)
    n_sp = int(n.item())
    sp = SuperpixelMaps.from_labels(labels, point_mask[0].to(dev), n_sp=n_sp)
    g = torch.Generator().manual_seed(0)
    lv = [torch.randn(H >> s, W >> s, 2 * c, generator=g).to(dev) for c, s in zip(bench.VGG_C, bench.VGG_SHIFT)]
    ia = _lib.int_array
    ca, ha, wa = ia([t.size(2) for t in lv]), ia([t.size(0) for t in lv]), ia([t.size(1) for t in lv])
    ctot = sum(t.size(2) for t in lv)
    fp = torch.empty(lib.wesup_footprint_bytes(ha, wa, 13, H, W, n_sp), dtype=torch.uint8, device=dev)
    lib.wesup_footprint_build(ha, wa, 13, H, W, n_sp, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),
                              sp.row_labels.data_ptr(), sp.counts.data_ptr(), 1, fp.data_ptr(), st)
    gl = [torch.empty_like(t) for t in lv]
    gptrs = _lib.ptr_array([t.data_ptr() for t in gl])
    gp = torch.randn(n_sp, ctot, device=dev)
    bwd = lambda: lib.wesup_levels_pool_bwd_fp(gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(), ca, ha, wa,  # noqa: E731
                                               13, H, W, n_sp, fp.data_ptr(), gptrs, st)
    # WESUP_FP_X = "order,minb": order 0 identity blocks spread between the list blocks (default), 1 first, 2 last;
    # minb = blocks per SM the kernel is compiled for (6: 80 registers, 8: 64 registers with spills)
    variants = sys.argv[1:] or ["0,6", "1,6", "2,6", "0,8", "2,8"]
    for rep in range(2):
        for v in variants:
            os.environ["WESUP_FP_X"] = v
            ms = bench.time_kernel(bwd, 20, flush)
            print(json.dumps({"x": v, "bwd_us": round(ms * 1e3, 2)}))


if __name__ == "__main__":
    main()
