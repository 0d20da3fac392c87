#!/usr/bin/env bash
# Round-2 ncu captures: one --set full capture of the dominant kernel of every workload (one launch each, reduced
# sizes where the full size only repeats tiles), then the launch list of the judged default command.
set -u
mkdir -p gpurun_out
cap() { # name workload kernel-regex units
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -o gpurun_out/r02_$1 -f \
    python bench.py --workload $2 ${4:+--units $4} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-200
}
cap dop_p5tet tabulate_p5_tet tabulate_dop 2000000
cap small_p1tri tabulate_p1_tri tabulate_small
cap tp_q4hex tabulate_q4_hex tabulate_tp_staged 10000000
cap map_rows1 push_forward_n1curl3 map_kernel
cap map_rows45 push_forward_n1curl3_rows45 map_warp_kernel 1000000
cap map_bcast_rows45 push_forward_bcast_n1curl3_rows45 physical_basis 2000000
cap dof_matrix_rt4 t_apply_rt4 dof_ 4000000
cap dof_matrix_rt4_block3 t_apply_rt4_block3 dof_ 2000000
cap dof_perm_p5tet permute_p5_tet dof_ 4000000
cap interp_n1curl3 interpolate_n1curl3 interpolate_kernel 4000000
cap geometry_tet geometry_tet geometry_affine 4000000
cap geometry_hex geometry_hex geometry_warp 1000000
cap interp_pulled_n1curl3 interpolate_pulled_n1curl3 interpolate_pulled 4000000
cap physical_basis_n1curl3 physical_basis_n1curl3 physical_basis 2000000
cap physical_basis_rt4 physical_basis_rt4 physical_basis 2000000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_default.csv \
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
# condense on the box (the reports are too large to travel: 64 MiB limit on gpurun_out/)
for r in gpurun_out/r02_*.ncu-rep; do
  n=$(basename "$r" .ncu-rep)
  python tools/ncu_summary.py "$r" > "gpurun_out/${n/r02_/r02_ncu_}.txt" 2>/dev/null
  python tools/ncu_stalls.py "$r" > "gpurun_out/${n/r02_/r02_stalls_}.txt" 2>/dev/null
done
for r in gpurun_out/r02_*.ncu-rep; do
  rm -f "$r"
done
du -sh gpurun_out
