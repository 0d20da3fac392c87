#!/usr/bin/env bash
# On the GPU box: time every build/ab/*.so on the default workload, interleaved, ROUNDS times.
set -u
mkdir -p gpurun_out
cp basix_b200/libbasix_b200.so /tmp/lib_orig.so
for r in $(seq 1 ${ROUNDS:-2}); do
  for so in build/ab/*.so; do
    cp $so basix_b200/libbasix_b200.so
    python bench.py --steps 10 --warmup 3 --device-only --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$so', round(d['ms_per_step'],4), d['clocks']['reasons'])"
  done
done
cp /tmp/lib_orig.so basix_b200/libbasix_b200.so
