#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider -x"
echo "== conv + network tests (store warp)"
timeout -s KILL 900 $PT tests/test_gpu_conv.py tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py -m gpu 2>&1 | tail -3
echo "== A/B: default (store warp) | in-group store issue"
for rnd in 1 2; do
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_LIB_PATH=$PWD/scratch/lib_groupstore.so timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/layer_times_sw.txt 2>&1; tail -1 gpurun_out/layer_times_sw.txt
