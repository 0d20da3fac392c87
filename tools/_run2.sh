run() { BX_DOF_TUNE="$2" timeout 200 python bench.py --workload $1 --units 10000000 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1','$2', round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"; }
for t in "" "pad=1" "skip=1" "skip=1,pad=1" "pad=2"; do run t_apply_rt4 "$t"; done
for t in "" "pad=1" "skip=1" "skip=1,pad=1" "cpt=16" "cpt=64" "nbuf=3" "nbuf=4,cpt=16" "pad=1,nbuf=3"; do run t_apply_rt4_block3 "$t"; done
