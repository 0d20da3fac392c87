for w in geometry_tet physical_basis_n1curl3 physical_basis_rt4 physical_basis_n1curl3_4pts t_apply_rt4_block3 tabulate_p1_tri; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --device-only --no-cpu-baseline 2>gpurun_out/err_$w.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', round(d['ms_per_step'],4), 'ms', '%.3e'%d['value'], d['unit'], 'frac', round(r['frac'],3), r.get('line_granular',{}).get('frac'), r.get('frac_of_copy_same_bytes'))"
done
