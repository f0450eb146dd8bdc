#!/usr/bin/env python
"""Extract per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, bytes per
launch, averaged over the captured launches) and duration from an `ncu --set full` report and
write profiles/roofline_traffic.json, which bench.py reads for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/prof_rNN.ncu-rep profiles/roofline_traffic.json
"""
import collections
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    acc = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_write.sum"]]]
        ms = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * TIME[units[col["gpu__time_duration.sum"]]]
        a = acc[name]
        a[0] += rd; a[1] += wr; a[2] += ms; a[3] += 1
    result = {name: {"dram_read_bytes": a[0] / a[3], "dram_write_bytes": a[1] / a[3], "traffic_bytes": (a[0] + a[1]) / a[3],
                     "ncu_ms": a[2] / a[3], "launches_captured": a[3]} for name, a in acc.items()}
    result["_source"] = rep
    with open(out, "w") as f:
        json.dump(result, f, indent=1, sort_keys=True)
    for name, v in sorted(result.items()):
        if name != "_source":
            print(f"{v['traffic_bytes'] / 1e6:10.1f} MB  {v['ncu_ms'] * 1e3:9.1f} us  x{v['launches_captured']:<3d} {name[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
