#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
echo "== tests: postprocess, letterbox, network, conv"
timeout -s KILL 900 $PT tests/test_gpu_postprocess.py tests/test_gpu_letterbox.py tests/test_gpu_network.py tests/test_gpu_conv.py tests/test_gpu_ingest.py -m gpu -x --deselect tests/test_gpu_postprocess.py::test_candidate_arena_grows_instead_of_truncating 2>&1 | tail -5
echo "== stages + predict"
timeout -s KILL 300 python scripts/stage_times.py 2>&1 | tail -3
echo "== A/B weight prefetch"
for rnd in 1 2; do
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_LIB_PATH=$PWD/scratch/lib_nopre.so timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
