#!/usr/bin/env python
"""Per-instruction view of an ncu report (SASS page): executed count, stall samples; prints the hottest instructions and
a histogram of executed instructions by opcode.  Usage: python tools/ncu_sass.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_ex = sum(int(r[iex]) for r in rows[1:])
tot_s = sum(int(r[isamp]) for r in rows[1:])
print(f"# {lines[0][:160]}")
print(f"# executed warp instructions {tot_ex}, stall samples {tot_s}")
ops = Counter()
for r in rows[1:]:
    op = r[isrc].split()[0]
    if op.startswith("@"):
        op = r[isrc].split()[1]
    ops[op.split(".")[0]] += int(r[iex])
print("# by opcode:", ", ".join(f"{k} {v / tot_ex:.1%}" for k, v in ops.most_common(18)))
print(f"# {'idx':>4} {'exec%':>6} {'samp%':>6}  sass")
for i, r in enumerate(rows[1:]):
    ex, sm = int(r[iex]), int(r[isamp])
    if sm / max(tot_s, 1) >= 0.012 or (top < 0):
        print(f"  {i:4d} {ex / tot_ex:6.2%} {sm / tot_s:6.2%}  {r[isrc].strip()}")
