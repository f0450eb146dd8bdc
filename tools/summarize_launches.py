#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel total device time, share of the captured region and launch count.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
"""
import collections
import csv
import sys


def main(path, top=45, marker=None, periods=2):
    """`marker`: keep only the launches between the (periods+1)-th last and the last launch of the kernel whose name
    contains it (e.g. fp_pool_fwd_cells_kernel, once per training iteration): `periods` steady-state iterations out of
    a capture of the whole process (cuDNN autotuning, graph warm-ups and the like are cut away)."""
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if marker:
        idx = [i for i, r in enumerate(rows) if marker in r["Kernel Name"]]
        if len(idx) > periods:
            rows = rows[idx[-periods - 1]:idx[-1]]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in rows:
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = row["Kernel Name"]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    ours = sum(v for k, v in tot.items() if "wesup::" in k)
    print(f"# launch list summary: {path}" + (f" ({periods} periods of `{marker}`)" if marker else "") + "\n")
    print(f"launches: {sum(cnt.values())}; summed device time {total / 1e3:.3f} ms "
          f"(per-launch times are cold-cache and serialised under ncu: compare SHARES);")
    print(f"wesup:: kernels: {ours / 1e3:.3f} ms = {100 * ours / total:.1f} % of the captured region\n")
    print("| us | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f"| {v:.1f} | {100 * v / total:.1f} % | {cnt[k]} | `{k[:110]}` |")


if __name__ == "__main__":
    main(sys.argv[1], marker=sys.argv[2] if len(sys.argv) > 2 else None, periods=int(sys.argv[3]) if len(sys.argv) > 3 else 2)
