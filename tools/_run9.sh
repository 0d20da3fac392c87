set -u
mkdir -p gpurun_out
PYTEST_TIMEOUT=1500 bash tools/gpu_round.sh
python tools/library_comparator.py > gpurun_out/r02_library_comparator.json 2> gpurun_out/libcomp.err; cat gpurun_out/r02_library_comparator.json; tail -n 2 gpurun_out/libcomp.err
