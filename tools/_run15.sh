set -u
AB_WORKLOADS="physical_basis_n1curl3 physical_basis_rt4" ROUNDS=2 bash tools/ab_run.sh
timeout 400 ncu --set full --clock-control none --import-source on -k regex:physical_basis -s 2 -c 1 -o gpurun_out/r02_physical_basis_rt4_flat -f \
    python bench.py --workload physical_basis_rt4 --units 2000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02_physical_basis_rt4_flat.ncu-rep > gpurun_out/r02_ncu_physical_basis_rt4.txt
python tools/ncu_stalls.py gpurun_out/r02_physical_basis_rt4_flat.ncu-rep > gpurun_out/r02_stalls_physical_basis_rt4.txt
rm -f gpurun_out/r02_physical_basis_rt4.ncu-rep
