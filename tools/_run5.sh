timeout 900 python -m pytest tests/test_geometry.py -m gpu -x -q 2>&1 | tail -3
run() { BX_FUSED_TUNE="$2" timeout 300 python bench.py --workload $1 --steps 10 --warmup 3 --device-only --no-cpu-baseline 2>gpurun_out/err_$1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', '$2', round(d['ms_per_step'],4), 'ms', '%.3e'%d['value'], d['unit'], 'frac', round(r['frac'],3))"; }
for t in "" "nw=8" ; do run physical_basis_rt4 "$t"; done
for t in "" ; do run physical_basis_n1curl3 "$t"; done
