set -u
for rep in 1 2; do
for v in "BX_SMALL_DIRECT=0" "BX_SMALL_DIRECT=1" "BX_SMALL_DIRECT=1 BX_SMALL_R=1" "BX_SMALL_DIRECT=1 BX_SMALL_R=4" "BX_SMALL_R=1" "BX_SMALL_R=4"; do
env $v python bench.py --workload tabulate_p1_tri --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['roofline']['frac'])"
done; done
python tools/small_floor.py
