#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider -x"
echo "== conv + pool + network + baseline tests (4 accumulator stages, SPP pool groups)"
timeout -s KILL 1200 $PT tests/test_gpu_conv.py tests/test_gpu_pool_upsample.py tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py tests/test_gpu_zoo.py -m gpu 2>&1 | tail -3
echo "== layer times"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/layer_times_acc4.txt 2>&1; tail -1 gpurun_out/layer_times_acc4.txt
echo "== A/B: default (4 stages) | YB_ACC2=1"
for rnd in 1 2; do
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_ACC2=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
