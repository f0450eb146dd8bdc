#!/usr/bin/env python
"""Registers / spills / static shared memory of every kernel of the library (nvcc -Xptxas -v, sm_100a), one line per kernel:
python tools/ptxas_table.py > profiles/rNN_ptxas_registers.txt"""
import re
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wesup_b200 import build as B  # noqa: E402

rows = []
for src in B.SOURCES:
    cmd = [B._nvcc(), *B.NVCC_FLAGS, "-Xptxas=-v", "-c", str(B.CSRC / src), "-o", "/dev/null"]
    err = subprocess.run(cmd, capture_output=True, text=True).stderr
    name = None
    spill = "0"
    for line in err.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            continue
        m = re.search(r"(\d+) bytes spill stores", line)
        if m:
            spill = m.group(1)
        m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", line)
        if m and name:
            rows.append((src, re.sub(r"\(.*", "", name)[:80], int(m.group(1)), int(spill), int(m.group(3) or 0)))
            name, spill = None, "0"
print(f"{'file':24s} {'kernel':80s} {'regs':>5s} {'spill B':>8s} {'static smem B':>14s}")
for r in sorted(rows):
    print(f"{r[0]:24s} {r[1]:80s} {r[2]:5d} {r[3]:8d} {r[4]:14d}")
