#!/usr/bin/env python
"""Write profiles/README.md: one table row per committed bench line (profiles/rNN_bench_*.json)."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{rnd}_bench_*.json"))):
    name = os.path.basename(f)[len(rnd) + 7:-5]
    try:
        d = json.load(open(f))
    except ValueError:
        continue
    r = d.get("roofline") or {}
    e2e = (d.get("e2e") or {}).get("value")
    cpu = (d.get("cpu_baseline") or {})
    hbm = r.get("hbm", r if r.get("bound") == "hbm" else {})
    rows.append((name, d.get("n_gpus"), d.get("ms_per_step"), d.get("value"), d.get("unit"), r.get("bound"), r.get("frac"),
                 hbm.get("frac"), e2e, cpu.get("value"), cpu.get("cores"), cpu.get("kind")))
fmt = lambda v, f="%.3g": "–" if v is None else f % v
out = [f"# profiles — round {rnd}", "",
       "Bench lines (`bench.py`, B200, CUDA events on the launch stream, inputs larger than L2) and ncu summaries",
       "(`tools/ncu_summary.py` over `ncu --set full --clock-control none` captures of the dominant kernel of each workload).",
       "`roofline frac` = algorithmic flops (tensor) or bytes (hbm) per launch / measured launch time / measured peak",
       "(FP64 DMMA 37.0 TFLOP/s, HBM copy 6542 GB/s); `e2e` = the same call with host buffers (H2D + kernels + D2H timed);",
       "`cpu` = the compiled reference (or, for interpolate, numpy dgemm with the reference's matrix) on the box's host cores.", "",
       "| workload | GPUs | ms/step | value | bound | roofline frac | HBM frac | e2e | cpu baseline |",
       "|---|---:|---:|---:|---|---:|---:|---:|---:|"]
for n, g, ms, v, u, b, fr, hf, e2e, cpu, cores, kind in rows:
    out.append(f"| {n} | {g} | {fmt(ms, '%.4g')} | {fmt(v)} {u or ''} | {b or '–'} | {fmt(fr, '%.3f')} | {fmt(hf, '%.3f')} | {fmt(e2e)} | "
               f"{fmt(cpu)}{'' if cpu is None else f' ({cores} cores, {kind})'} |")
out += ["", "ncu summaries: " + ", ".join(sorted(os.path.basename(f) for f in glob.glob(os.path.join(ROOT, "profiles", f"{rnd}_ncu_*.txt")))), "",
        f"Launch list of the judged command: `{rnd}_ncu_launches_bench_default.csv`; measured DRAM traffic per unit: `{rnd}_traffic.json`;",
        "FP64 / HBM / PCIe peaks: `r01_peaks_fp64_hbm_pcie.json`; shared DMMA/DFMA pipe: `r01_mixed_fp64.json`; cuBLAS DGEMM and the",
        "library-path comparator: `r02_library_comparator.json`; SASS opcode histogram: `r02_sass_opcodes.txt`; per-kernel stall",
        f"reasons: `{rnd}_stalls_*.txt`; compute-sanitizer status and the guard-zone substitute: `r02_sanitizer.txt`.",
        "Rows other than `default_*` / `reference` are the `configs` entries of the default line (`r02_bench_default_n1.json`),",
        "one file each; round-1 files (`r01_*`) are kept for comparison.", ""]
open(os.path.join(ROOT, "profiles", "README.md"), "w").write("\n".join(out))
print("\n".join(out))
