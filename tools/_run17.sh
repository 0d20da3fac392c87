set -u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_pipeline.py -m gpu -x -q -k "push or pull or map or pipeline" 2>&1 | tail -3
for k in 1 0; do
for w in push_forward_n1curl3_rows45 pull_back_n1curl3_rows45; do
BX_MAP_WARP=$k python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warp=$k $w', d['ms_per_step'], d['roofline']['frac'])"
done; done
