#!/usr/bin/env bash
# One GPU-box pass: parity tests, smoke, the judged bench line (+ reference arm), every other workload, and the
# ncu launch list of the judged command.  Outputs under gpurun_out/ (copy what should be kept into profiles/).
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for w in tabulate_p1_tri tabulate_q4_hex push_forward_n1curl3 pull_back_n1curl3 push_forward_n1curl3_rows45 push_forward_bcast_n1curl3_rows45 t_apply_rt4 permute_p5_tet interpolate_n1curl3; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 ${EXTRA:-} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline") or {}
    e2e = (d.get("e2e") or {}).get("value")
    cpu = (d.get("cpu_baseline") or {}).get("value")
    print(f.split("bench_")[1][:-5], "ms/step %.3f" % d.get("ms_per_step", float("nan")), "value %.3e" % d.get("value", float("nan")),
          r.get("bound"), "frac %.3f" % r.get("frac", float("nan")), "e2e", e2e and "%.3e" % e2e, "cpu", cpu and "%.3e" % cpu)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
