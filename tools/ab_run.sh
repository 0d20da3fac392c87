#!/usr/bin/env bash
# On the GPU box: time every build/ab/*.so, interleaved, ROUNDS times, on the workloads of AB_WORKLOADS (default: the
# headline workload).
set -u
mkdir -p gpurun_out
cp basix_b200/libbasix_b200.so /tmp/lib_orig.so
for r in $(seq 1 ${ROUNDS:-2}); do
  for so in build/ab/*.so; do
    cp $so basix_b200/libbasix_b200.so
    for w in ${AB_WORKLOADS:-tabulate_p5_tet}; do
    python bench.py --workload $w --steps 10 --warmup 3 --device-only --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$so', '$w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['reasons'])"
    done
  done
done
cp /tmp/lib_orig.so basix_b200/libbasix_b200.so
