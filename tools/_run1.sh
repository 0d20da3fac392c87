set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "register or small or tabulate" > gpurun_out/pytest_small.log 2>&1; tail -3 gpurun_out/pytest_small.log
for r in 1 2 4; do
  BX_SMALL_R=$r timeout 300 python bench.py --workload tabulate_p1_tri --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('R=$r', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done
