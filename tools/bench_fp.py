#!/usr/bin/env python
"""Footprint pooling kernels alone at the bench shape (464x464, 4224 backbone channels):
build / fwd / bwd times, L2 flushed before every launch.  python tools/bench_fp.py"""
import json
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
    ptrs = _lib.ptr_array([t.data_ptr() for t in lv])
    ctot = sum(t.size(2) for t in lv)
    nbytes = sum(t.numel() * 4 for t in lv) + H * W * 4 + n_sp * ctot * 4 + n_sp * 4
    fp = torch.empty(lib.wesup_footprint_bytes(ha, wa, 13, H, W, n_sp), dtype=torch.uint8, device=dev)
    build = lambda: lib.wesup_footprint_build(ha, wa, 13, H, W, n_sp, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),  # noqa: E731
                                              sp.row_labels.data_ptr(), sp.counts.data_ptr(), 1, fp.data_ptr(), st)
    out = {"n_sp": n_sp, "bytes": nbytes, "build_ms": bench.time_kernel(build, 10, flush)}
    pooled = torch.empty(n_sp, ctot, device=dev)
    gl = [torch.empty_like(t) for t in lv]
    gptrs = _lib.ptr_array([t.data_ptr() for t in gl])
    gp = torch.randn(n_sp, ctot, device=dev)
    fwd = lambda: lib.wesup_levels_pool_fwd_fp(ptrs, ca, ha, wa, 13, H, W, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),  # noqa: E731
                                               n_sp, fp.data_ptr(), pooled.data_ptr(), st)
    bwd = lambda: lib.wesup_levels_pool_bwd_fp(gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(), ca, ha, wa,  # noqa: E731
                                               13, H, W, n_sp, fp.data_ptr(), gptrs, st)

    def timed(fn):
        ms = bench.time_kernel(fn, 20, flush)
        return {"ms": ms, "gbs": nbytes / ms / 1e6}

    import os
    out["fwd"] = timed(fwd)
    out["bwd"] = timed(bwd)
    merged = [t.clone() for t in gl]
    os.environ["WESUP_FP_BWD"] = "split"              # cells + identity kernels as two launches on two streams
    for t in gl:
        t.zero_()
    out["bwd_split_launches"] = timed(bwd)
    os.environ.pop("WESUP_FP_BWD")
    torch.cuda.synchronize()
    out["bwd_merged_equals_split"] = all(torch.equal(a, b) for a, b in zip(merged, gl))
    if not bench.QUICK:
        os.environ["WESUP_FP_FWD"] = "chunks"
        out["fwd_chunk_kernel"] = timed(fwd)
        os.environ.pop("WESUP_FP_FWD")
        os.environ["WESUP_FP_BWD"] = "chunks"
        out["bwd_chunk_kernel"] = timed(bwd)
        os.environ.pop("WESUP_FP_BWD")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
