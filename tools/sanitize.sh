#!/usr/bin/env bash
# compute-sanitizer passes over the GPU parity suite (race / memory checking of the kernels with DMA warps,
# mbarriers and in-place shared staging).  Full-size tests are left out (the tools slow kernels down 10-100x); every
# kernel of the library is reached by the small-size parity tests.  Summaries land in gpurun_out/sanitizer_*.txt.
set -u
mkdir -p gpurun_out
SMALL="tests/test_gpu_parity.py tests/test_geometry.py tests/test_interpolation.py tests/test_subentity_closure.py tests/test_pipeline.py"
run() { # tool  timeout  pytest-args...
  local tool=$1 tmo=$2
  shift 2
  local log=gpurun_out/sanitizer_$tool.log
  timeout "$tmo" compute-sanitizer --tool "$tool" --target-processes all --print-limit 20 --error-exitcode 86 \
    python -m pytest -m gpu -q -x -p no:cacheprovider "$@" > "$log" 2>&1
  local rc=$?
  {
    echo "# compute-sanitizer --tool $tool  python -m pytest -m gpu -q -x $*"
    echo "exit code: $rc   (86 = the tool reported errors, 124 = timeout)"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" "$log" | tail -8
    grep -E "^========= (Invalid|Race|Barrier|Uninit|Program hit|Error|Warning|Hazard|ERROR|WARN)" "$log" | sed -E 's/0x[0-9a-f]+/0x../g; s/thread \([0-9,]+\)/thread (..)/g; s/block \([0-9,]+\)/block (..)/g' | sort | uniq -c | sort -rn | head -20
    echo "--- first report ---"
    grep -m1 -A14 -E "^========= (Invalid|Race|Barrier|Uninit|Program hit|Error|Warning|Hazard|ERROR|WARN)" "$log" | cut -c1-220
  } > gpurun_out/sanitizer_$tool.txt
  head -c 60000 "$log" > "$log.head"; tail -c 60000 "$log" > "$log.tail"; rm -f "$log"
  cat gpurun_out/sanitizer_$tool.txt
}
run memcheck ${MEMCHECK_TIMEOUT:-900} $SMALL
# the kernels with shared-memory hand-offs: TMA / mbarrier tiles of the DOF kernels, closure staging, warp-private tiles
# with bulk copy-out (register, tensor-product, fused physical basis), the derivative-operator and interpolation kernels
RACE_K="T_apply or permute or closure or register_path or tensor_product_staged or derivative_operator or physical_basis or interpolat or broadcast or geometry"
run racecheck ${RACECHECK_TIMEOUT:-900} $SMALL -k "$RACE_K"
run synccheck ${SYNCCHECK_TIMEOUT:-600} $SMALL -k "$RACE_K"
