#!/usr/bin/env python
"""Footprint pooling kernels per level resolution at the bench shape (464x464, backbone channels): which levels
cost what.  Each subset of levels gets its own footprint blob.  python tools/bench_fp_levels.py"""
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
    H = W = int(sys.argv[1]) if len(sys.argv) > 1 else bench.H
    n_seg = int(sys.argv[2]) if len(sys.argv) > 2 else int(H * W / 200)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    flush = bench.L2Flush(dev)
    img, _, point_mask = synth.sample(H, W, index=0)
    labels, n = ops.slic(img.to(dev), n_seg, 40)
    n_sp = int(n.item())
    sp = SuperpixelMaps.from_labels(labels, point_mask[0].to(dev), n_sp=n_sp)
    g = torch.Generator().manual_seed(0)
    all_lv = [torch.randn(H >> s, W >> s, 2 * c, generator=g).to(dev) for c, s in zip(bench.VGG_C, bench.VGG_SHIFT)]
    shifts = sorted(set(bench.VGG_SHIFT))
    subsets = {f"shift{s}": [t for t, sh in zip(all_lv, bench.VGG_SHIFT) if sh == s] for s in shifts}
    subsets["all"] = all_lv
    subsets["non_identity"] = [t for t, sh in zip(all_lv, bench.VGG_SHIFT) if sh != 0]
    ia = _lib.int_array
    for name, lv in subsets.items():
        nl = len(lv)
        ca, ha, wa = ia([t.size(2) for t in lv]), ia([t.size(0) for t in lv]), ia([t.size(1) for t in lv])
        ptrs = _lib.ptr_array([t.data_ptr() for t in lv])
        ctot = sum(t.size(2) for t in lv)
        wbytes = sum(t.numel() * 4 for t in lv)
        fp = torch.empty(max(lib.wesup_footprint_bytes(ha, wa, nl, H, W, n_sp), 256), dtype=torch.uint8, device=dev)
        lib.wesup_footprint_build(ha, wa, nl, H, W, n_sp, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),
                                  sp.row_labels.data_ptr(), sp.counts.data_ptr(), 1, fp.data_ptr(), st)
        pooled = torch.empty(n_sp, ctot, device=dev)
        gl = [torch.empty_like(t) for t in lv]
        gptrs = _lib.ptr_array([t.data_ptr() for t in gl])
        gp = torch.randn(n_sp, ctot, device=dev)
        fwd = lambda: lib.wesup_levels_pool_fwd_fp(ptrs, ca, ha, wa, nl, H, W, sp.seg_offsets.data_ptr(), sp.seg_pixels.data_ptr(),  # noqa: E731
                                                   n_sp, fp.data_ptr(), pooled.data_ptr(), st)
        bwd = lambda: lib.wesup_levels_pool_bwd_fp(gp.data_ptr(), sp.row_labels.data_ptr(), sp.counts.data_ptr(), ca, ha, wa,  # noqa: E731
                                                   nl, H, W, n_sp, fp.data_ptr(), gptrs, st)
        f_ms, b_ms = bench.time_kernel(fwd, 20, flush), bench.time_kernel(bwd, 20, flush)
        print(json.dumps({"H": H, "n_sp": n_sp, "levels": name, "n": nl, "shape": list(lv[0].shape), "level_mb": round(wbytes / 1e6, 1),
                          "fwd_us": round(f_ms * 1e3, 1), "fwd_gbs": round(wbytes / f_ms / 1e6), "bwd_us": round(b_ms * 1e3, 1),
                          "bwd_gbs": round(wbytes / b_ms / 1e6)}))


if __name__ == "__main__":
    main()
