#!/usr/bin/env python
"""Hottest SASS instructions of one kernel in an ncu report (pc-sampling), with the stall reason
columns that carry samples:  python tools/ncu_hot.py rep.ncu-rep 'levels_pool_bwd_kernel<2>' [top]"""
import csv
import io
import subprocess
import sys


def main(rep, needle, top=25):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        if needle not in b["name"]:
            continue
        hdr = b["rows"][0]
        col = {h: i for i, h in enumerate(hdr)}
        body = [r for r in b["rows"][1:] if len(r) == len(hdr)]
        tot = sum(int(r[col["# Samples"]] or 0) for r in body)
        inst = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
        print(f"## {b['name'][:100]}\nsamples {tot}, warp instructions {inst}, SASS lines {len(body)}\n")
        stall_cols = [h for h in hdr if h.startswith("stall_") or "Stall" in h and "Sampling" not in h]
        order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
        for i in sorted(order):
            r = body[i]
            stalls = sorted(((int(r[col[h]] or 0), h) for h in hdr[col["# Samples"] + 1:] if h.startswith("stall") and (r[col[h]] or "0").isdigit()),
                            reverse=True)[:3]
            print(f"{i:5d} {int(r[col['# Samples']]):6d} smp {int(r[col['Instructions Executed']] or 0):9d} ex  {r[col['Source']].strip()[:70]:70s} "
                  + " ".join(f"{h[6:]}={v}" for v, h in stalls if v))
        print()
        break


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
