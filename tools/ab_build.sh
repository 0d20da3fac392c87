#!/usr/bin/env bash
# Development A/B builds of the derivative-operator kernel: tools/ab_build.sh NAME "-DBX_DOP_V=1 ..." -> build/ab/NAME.so
# (only the BASELINE element is instantiated: BX_DOP_DEV)
set -eu
name=$1; flags=$2
mkdir -p build/ab build/abobj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v \
  -DBX_DOP_DEV $flags -c basix_b200/csrc/tabulate_dop.cu -o build/abobj/$name.o 2> build/abobj/$name.log
grep -A2 "Li12EdE" build/abobj/$name.log | grep -i "used" || true
objs=$(ls build/obj/*.o | grep -v tabulate_dop.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/$name.so $objs build/abobj/$name.o -cudart static
