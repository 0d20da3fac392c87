set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tabulate or smoke or pipeline or interpolation" > gpurun_out/pytest_dop.log 2>&1; tail -n 2 gpurun_out/pytest_dop.log
BX_DOF_TUNE="direct=1" timeout 900 python -m pytest tests -m gpu -x -q -k "T_apply or transform or operators or t_apply" > gpurun_out/pytest_direct.log 2>&1; tail -n 2 gpurun_out/pytest_direct.log
for t in "direct=0" "direct=1" "direct=0" "direct=1"; do
  BX_DOF_TUNE="$t" python bench.py --workload t_apply_rt4 --steps 10 --warmup 3 --device-only --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$t', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
done
python bench.py --steps 10 --warmup 3 --device-only --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
