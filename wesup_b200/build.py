"""Build recipe for libwesup_b200.so (hand-written CUDA, sm_100a only).

    python -m wesup_b200.build [--force] [--verbose]

The library is built IN-TREE (wesup_b200/libwesup_b200.so) so that it travels
to the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "csrc" / "_obj"
LIB = PKG / "libwesup_b200.so"
SOURCES = ["abi.cu", "hypercolumn.cu", "sp_pool.cu", "label_propagate.cu", "label_propagate_tc.cu", "slic.cu", "fused_bwd.cu", "fused_fwd.cu", "pool_levels.cu", "footprint.cu", "colsum.cu", "upsample_sum.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "wesup_b200.h", Path(__file__)]
    sources = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    nvcc = _nvcc()

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_tools() -> Path:
    """tools/l2bw.cu -> tools/_l2bw: the stand-alone L2 / HBM gather microbenchmark quoted in DESIGN.md section 3
    (a measurement tool, not part of the library)."""
    src, out = PKG.parent / "tools" / "l2bw.cu", PKG.parent / "tools" / "_l2bw"
    if src.exists() and _stale(out, [src]):
        r = subprocess.run([_nvcc(), "-O3", "-gencode", "arch=compute_100a,code=sm_100a", str(src), "-o", str(out)],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
