set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tabulate or smoke or pipeline or interpolation" > gpurun_out/pytest_dop.log 2>&1; tail -n 4 gpurun_out/pytest_dop.log
timeout 600 python bench.py --steps 10 --warmup 3 --device-only --no-configs --no-cpu-baseline > gpurun_out/bench_dop.json 2> gpurun_out/bench_dop.err; tail -n 3 gpurun_out/bench_dop.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_dop.json'))
print(d['ms_per_step'], d['roofline']['frac'], d['clocks'])
PY
