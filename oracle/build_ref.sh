#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- never part of the product path.
#
# Builds oracle/_ref/libbasix_ref.so: the UNMODIFIED reference C++ core
# (/root/reference/cpp/basix/*.cpp, compiled where it lies, nothing copied)
# plus our own extern "C" shim (oracle/ref_shim.cpp) so Python tests and the
# bench's cpu_baseline leg can drive it through ctypes.
#
# Recipe follows SURVEY.md section 8c: g++ -std=c++20 -O2 -ffp-contract=off,
# BLAS/LAPACK symbols renamed to the LP64 OpenBLAS that ships inside scipy
# (the image has no system BLAS).  Outputs only into oracle/_ref/ (git-ignored,
# but NOT gpurun-ignored, so the .so travels to the GPU box where
# /root/reference does not exist).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${BASIX_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
if [ ! -d "$REF/cpp/basix" ]; then
  echo "build_ref: $REF/cpp/basix not present; keeping any prebuilt $OUT/libbasix_ref.so" >&2
  exit 0
fi
PY="${PYTHON:-python}"
SCIPY_LIBS="$($PY - <<'EOF'
import os, scipy
print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs"))
EOF
)"
OPENBLAS="$(ls "$SCIPY_LIBS"/libscipy_openblas*.so | head -1)"
mkdir -p "$OBJ/basix"
printf '#pragma once\n#define BASIX_VERSION 0.12.0.0\n' > "$OBJ/basix/version.h"

DEFS="-DMDSPAN_USE_PAREN_OPERATOR=1 -DMDSPAN_USE_BRACKET_OPERATOR=0"
for r in gemm gesv getrf syevd; do
  DEFS="$DEFS -Dd${r}_=scipy_d${r}_ -Ds${r}_=scipy_s${r}_"
done
CXXFLAGS="-std=c++20 -O2 -ffp-contract=off -fPIC $DEFS -I$REF/cpp -I$OBJ"

# newest input (reference sources + shim) vs the library: skip if up to date
LIB="$OUT/libbasix_ref.so"
if [ -f "$LIB" ] && [ -z "$(find "$REF/cpp/basix" "$HERE/ref_shim.cpp" "$HERE/build_ref.sh" -newer "$LIB" -name '*.*' | head -1)" ]; then
  echo "build_ref: $LIB up to date"
  exit 0
fi

pids=()
for src in "$REF"/cpp/basix/*.cpp; do
  o="$OBJ/$(basename "${src%.cpp}").o"
  if [ ! -f "$o" ] || [ "$src" -nt "$o" ]; then
    g++ $CXXFLAGS -c "$src" -o "$o" &
    pids+=($!)
  fi
done
g++ $CXXFLAGS -c "$HERE/ref_shim.cpp" -o "$OBJ/ref_shim.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done

g++ -shared -o "$LIB" "$OBJ"/*.o "$OPENBLAS" -lpthread \
    -Wl,-rpath,"$SCIPY_LIBS" -Wl,-rpath,'$ORIGIN'
echo "build_ref: built $LIB (BLAS: $OPENBLAS)"
