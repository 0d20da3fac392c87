#!/usr/bin/env python
"""Per-instruction stall REASONS of an ncu report (source page): for a range of instruction indices prints executed
share, samples and the split by reason; and totals by reason for the kernel.
Usage: python tools/ncu_stalls.py report.ncu-rep [first last]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else -1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in reasons}
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = {h: 0 for h in reasons}
tot_s = 0
tot_ex = sum(int(r[iex]) for r in rows[1:])
for r in rows[1:]:
    tot_s += int(r[isamp])
    for h in reasons:
        tot[h] += int(r[idx[h]] or 0)
print("# totals:", ", ".join(f"{h[6:]} {v / max(tot_s, 1):.1%}" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
if hi >= 0:
    for i, r in enumerate(rows[1:]):
        if lo <= i <= hi:
            parts = sorted(((int(r[idx[h]] or 0), h[6:]) for h in reasons), reverse=True)
            desc = " ".join(f"{n}:{c / tot_s:.2%}" for c, n in parts[:3] if c)
            print(f"{i:5d} {int(r[iex]) / tot_ex:6.2%} {int(r[isamp]) / tot_s:6.2%}  {r[isrc].strip():50s} {desc}")
