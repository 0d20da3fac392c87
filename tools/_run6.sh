cap() { # name workload kernel-regex units
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -o gpurun_out/r02_$1 -f \
    python bench.py --workload $2 ${4:+--units $4} --steps 1 --warmup 3 --device-only --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap fused_rt4 physical_basis_rt4 physical_basis_kernel 2000000
