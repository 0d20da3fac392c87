set -u
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/_cmp_bcast.py
for w in push_forward_bcast_n1curl3_rows45; do
python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['ms_per_step'], d['roofline']['frac'])"
done
