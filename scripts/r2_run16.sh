#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider"
echo "== conv tests"
timeout -s KILL 600 $PT tests/test_gpu_conv.py -m gpu -s > gpurun_out/conv_tests.log 2>&1; grep -aE "violations=[1-9]|viol [1-9]|passed|failed|^FAILED|Error" gpurun_out/conv_tests.log | cut -c1-300 | tail -12
echo "== postprocess tests"
timeout -s KILL 600 $PT tests/test_gpu_postprocess.py tests/test_gpu_logits_decoder.py -m gpu > gpurun_out/post_tests.log 2>&1; grep -aE "passed|failed|^FAILED|^E  " gpurun_out/post_tests.log | cut -c1-300 | tail -20
echo "== network + baseline shapes + engine + zoo + p6 + v4"
timeout -s KILL 1200 $PT -x tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py tests/test_gpu_engine.py tests/test_gpu_zoo.py tests/test_p6.py tests/test_v4.py -m gpu -s 2>&1 | grep -aE "PARITY|stage-wise|passed|failed|Error|assert" | cut -c1-250 | tail -30
echo "== layer times"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/layer_times_v16.txt 2>&1; tail -1 gpurun_out/layer_times_v16.txt
echo "== A/B plan time: default | no chain | no nsplit | no chain+no nsplit | no guard | no hoist"
for rnd in 1 2; do
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_NO_CHAIN=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_NO_NSPLIT=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_NO_CHAIN=1 YB_NO_NSPLIT=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_LIB_PATH=$PWD/scratch/lib_noguard.so timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_LIB_PATH=$PWD/scratch/lib_nohoist.so timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
echo "== stage times"
timeout -s KILL 300 python scripts/stage_times.py 2>&1 | tail -3
timeout -s KILL 120 python scripts/nms_phases.py 2>&1 | tail -1
