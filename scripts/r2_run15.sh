#!/bin/bash
# N-split resident weights, bias hoist, exact-SiLU guard, chained tails with own accumulators.
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider -x"
echo "== conv tests (all)"
timeout -s KILL 900 $PT tests/test_gpu_conv.py -m gpu -s > gpurun_out/conv_tests.log 2>&1; grep -aE "violations=[1-9]|viol [1-9]|passed|failed|Error" gpurun_out/conv_tests.log | cut -c1-300 | tail -12
echo "== network + baseline shapes + engine + zoo + p6 + v4"
timeout -s KILL 1200 $PT tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py tests/test_gpu_engine.py tests/test_gpu_zoo.py tests/test_p6.py tests/test_v4.py -m gpu -s 2>&1 | grep -aE "PARITY|stage-wise|passed|failed|Error|assert" | cut -c1-250 | tail -30
echo "== layer times"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/layer_times_v15.txt 2>&1; tail -1 gpurun_out/layer_times_v15.txt
YB_NO_CHAIN=1 timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/layer_times_v15_nochain.txt 2>&1; tail -1 gpurun_out/layer_times_v15_nochain.txt
echo "== A/B plan time: default | no chain | no nsplit | no chain+no nsplit | no guard"
for rnd in 1 2; do
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_NO_CHAIN=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_NO_NSPLIT=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_NO_CHAIN=1 YB_NO_NSPLIT=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_LIB_PATH=$PWD/scratch/lib_noguard.so timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
