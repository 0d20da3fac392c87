set -u
mkdir -p gpurun_out
cap() { # name workload kernel-regex units
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -o gpurun_out/r02_$1 -f \
    python bench.py --workload $2 ${4:+--units $4} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/ncu_$1.log 2>&1
  tail -n 1 gpurun_out/ncu_$1.log
}
cap dop_p5tet tabulate_p5_tet tabulate_dop 2000000
