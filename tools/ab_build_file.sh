#!/usr/bin/env bash
# Development A/B build of ONE source file with extra flags: tools/ab_build_file.sh NAME FILE.cu "-DX=1 ..." [-fmad=false]
# -> build/ab/NAME.so (all other objects from build/obj).  Time the variants on the GPU box with tools/ab_run.sh
# (AB_BENCH="--workload ... " selects the workload).
set -eu
name=$1; file=$2; flags=$3; extra=${4:-}
base=$(basename "$file" .cu)
mkdir -p build/ab build/abobj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v \
  $extra $flags -c "$file" -o build/abobj/$name.o 2> build/abobj/$name.log
objs=$(ls build/obj/*.o | grep -v "/${AB_REPLACES:-$base}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/$name.so $objs build/abobj/$name.o -cudart static
echo "built build/ab/$name.so"
