#!/usr/bin/env python
"""Print the headline and the per-config figures of one bench.py JSON line as a table."""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` closes the pipe early

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
e = d.get("e2e") or {}
print("default  ms/step %.3f  value %.3e  %s frac %.3f  e2e %s (%.3f of the D2H ceiling %s GB/s)" % (
    d["ms_per_step"], d["value"], r.get("bound"), r.get("frac", float("nan")), e.get("value") and "%.3e" % e["value"],
    e.get("frac_of_d2h_ceiling", float("nan")), e.get("d2h_ceiling_GBps_per_gpu") and "%.1f" % e["d2h_ceiling_GBps_per_gpu"]))
for k, c in (d.get("configs") or {}).items():
    if "error" in c:
        print("%-36s ERROR %s" % (k, c["error"]))
        continue
    rf = c["roofline"]
    print("%-36s ms/step %9.4f  value %.3e %-8s %s frac %.3f%s" % (
        k, c["ms_per_step"], c["value"], c["unit"], rf["bound"], rf["frac"],
        "  hbm %.3f" % rf["hbm_frac"] if "hbm_frac" in rf else ""))
for key in ("gather", "strong"):
    if d.get(key):
        print(key, json.dumps(d[key]))
