#!/usr/bin/env python
"""profiles/rNN_traffic.json from the committed ncu summaries: DRAM bytes (read + write) of the dominant kernel per unit.
Usage: python tools/traffic_from_ncu.py r02"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
# workload -> (summary, units in the captured launch, unit)
CAPTURES = {
    "tabulate_p5_tet": ("dop_p5tet", 2_000_000, "point"),
    "tabulate_q4_hex": ("tp_q4hex", 2_000_000, "point"),  # --units 10M = 5 launches of 2M
    "tabulate_p1_tri": ("small_p1tri", 1_000_000, "point"),
    "push_forward_n1curl3": ("map_rows1", 20_000_000, "cell"),
    "pull_back_n1curl3": ("map_rows1", 20_000_000, "cell"),
    "push_forward_n1curl3_rows45": ("map_rows45", 1_000_000, "cell"),
    "pull_back_n1curl3_rows45": ("map_rows45", 1_000_000, "cell"),
    "push_forward_bcast_n1curl3_rows45": ("map_bcast_rows45", 2_000_000, "cell"),
    "t_apply_rt4": ("dof_matrix_rt4", 4_000_000, "cell"),
    "t_apply_rt4_block3": ("dof_matrix_rt4_block3", 2_000_000, "cell"),
    "permute_p5_tet": ("dof_perm_p5tet", 4_000_000, "cell"),
    "interpolate_n1curl3": ("interp_n1curl3", 4_000_000, "cell"),
    "geometry_tet": ("geometry_tet", 4_000_000, "cell"),
    "geometry_hex": ("geometry_hex", 1_000_000, "cell"),
    "interpolate_pulled_n1curl3": ("interp_pulled_n1curl3", 4_000_000, "cell"),
    "physical_basis_n1curl3": ("physical_basis_n1curl3", 2_000_000, "cell"),
    "physical_basis_rt4": ("physical_basis_rt4", 2_000_000, "cell"),
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def grab(text, metric):
    m = re.search(r"^\s*" + re.escape(metric) + r"\s+([0-9.]+)\s+(\w+)\s*$", text, re.M)
    return float(m.group(1)) * SCALE[m.group(2)]


out = {"_comment": "DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel, per unit, from the "
                   f"`ncu --set full` captures summarised in profiles/{rnd}_ncu_*.txt (tools/ncu_r02.sh). bench.py multiplies "
                   "by the units one launch processes to fill roofline.traffic."}
for wl, (name, units, unit) in CAPTURES.items():
    path = os.path.join(ROOT, "profiles", f"{rnd}_ncu_{name}.txt")
    if not os.path.exists(path):
        continue
    t = open(path).read()
    rd, wr = grab(t, "dram__bytes_read.sum"), grab(t, "dram__bytes_write.sum")
    kern = re.search(r"^kernel: (.*)$", t, re.M).group(1)[:80]
    out[wl] = {"bytes_per_unit": round((rd + wr) / units, 1), "unit": unit,
               "capture": f"{rnd}_ncu_{name}.txt ({units} {unit}s: {rd / 1e9:.3f} GB read + {wr / 1e9:.3f} GB written; {kern})"}
json.dump(out, open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w"), indent=1)
print(json.dumps({k: v["bytes_per_unit"] for k, v in out.items() if k != "_comment"}, indent=1))
