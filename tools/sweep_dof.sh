# development aid: sweep the BX_DOF_TUNE knobs of the DOF-transformation kernels on a GPU box
run() { BX_DOF_TUNE="$2" timeout 200 python bench.py --workload $1 --units 10000000 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1','$2', round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"; }
for t in "" "promo=2" "nbuf=4" "nbuf=2" "cpt=32,nbuf=4" "cpt=128,nbuf=2" "skip=1" "skip=1,promo=2" "skip=1,cpt=32,nbuf=4" "skip=1,cpt=128,nbuf=2"; do run t_apply_rt4 "$t"; done
for t in "" "promo=2" "nbuf=4" "cpt=64,nbuf=4" "cpt=256,nbuf=2" "skip=1" "skip=1,promo=2" "skip=1,cpt=256,nbuf=2"; do run permute_p5_tet "$t"; done
