#!/usr/bin/env python
"""Summarise an ncu report (read with `ncu -i <rep> --page raw --csv`) into the
handful of numbers the roofline discussion uses: duration, DRAM bytes, achieved
DRAM throughput, occupancy, hit rates, issue activity and the top stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/rNN_<kernel>.md

One section per (kernel name, grid size): the first captured launch of every distinct launch shape.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor pipe instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main(rep, needle=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    seen = set()
    print(f"# ncu summary of `{rep}`\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if needle and needle not in name:
            continue
        grid = r[col["launch__grid_size"]] if "launch__grid_size" in col else ""
        if (name, grid) in seen:                            # one section per kernel AND launch shape
            continue
        seen.add((name, grid))
        print(f"## `{name[:120]}`" + (f" — grid {grid}" if grid else "") + "\n")
        print("| metric | value |\n|---|---|")
        for key, label in KEYS:
            if key in col and r[col[key]] not in ("", "n/a"):
                print(f"| {label} | {r[col[key]]} {units[col[key]]} |")
        stalls = []
        for h, i in col.items():
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    stalls.append((float(r[i].replace(",", "")), h.split("stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        top = sorted(stalls, reverse=True)[:5]
        print("| top stall reasons (pc samples) | " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in top) + " |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
