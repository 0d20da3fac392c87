set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_geometry.py tests/test_pipeline.py -m gpu -x -q 2>&1 | tail -3
for w in physical_basis_n1curl3 physical_basis_rt4; do
for t in "flat=1" "flat=1,nw=4" "flat=2" "flat=1" "flat=2"; do
echo "== $w $t"
BX_FUSED_TUNE="$t" timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['config'].get('fused_equals_three_calls_bitwise'))"
done; done
