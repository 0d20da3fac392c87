set -u
mkdir -p gpurun_out
for t in 0 1; do
BX_DOF_TUNE="direct=$t" timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:dof_ -s 2 -c 1 --csv --log-file gpurun_out/dof_direct$t.csv python bench.py --workload t_apply_rt4 --units 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > /dev/null 2>&1
done
