#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
echo "== f16x2 epilogue variant: parity at real shapes"
YB_LIB_PATH=$PWD/scratch/lib_f16x2.so timeout -s KILL 900 $PT tests/test_gpu_baseline_shapes.py tests/test_gpu_network.py tests/test_gpu_zoo.py -m gpu -s > gpurun_out/tests_f16x2.log 2>&1; grep -aE "PARITY|stage-wise|passed|failed" gpurun_out/tests_f16x2.log | cut -c1-220
echo "== ncu decode v4"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"decode_" -s 3 -c 1 -o gpurun_out/r2_decode_v4 -f python scripts/one_step.py 4 > gpurun_out/ncu_dec.log 2>&1; tail -1 gpurun_out/ncu_dec.log
