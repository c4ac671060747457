#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
echo "== conv tests (default lib)"
timeout -s KILL 600 $PT tests/test_gpu_conv.py tests/test_gpu_engine.py -m gpu -x 2>&1 | tail -4
echo "== conv tests (single-lane lib)"
YB_LIB_PATH=$PWD/scratch/lib_single.so timeout -s KILL 600 $PT tests/test_gpu_conv.py -m gpu -x 2>&1 | tail -3
echo "== A/B step"
for rnd in 1 2; do
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_LIB_PATH=$PWD/scratch/lib_single.so timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
echo "== layer times"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/r2_layers_run5_uniform.txt 2>&1; tail -1 gpurun_out/r2_layers_run5_uniform.txt
YB_LIB_PATH=$PWD/scratch/lib_single.so timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/r2_layers_run5_single.txt 2>&1; tail -1 gpurun_out/r2_layers_run5_single.txt
