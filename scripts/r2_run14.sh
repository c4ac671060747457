#!/bin/bash
# Chained tails with their own TMEM accumulators: unit tests, network suites, timing A/B, post-kernel profile.
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider -x"
echo "== chain unit tests"
timeout -s KILL 300 $PT tests/test_gpu_conv.py -m gpu -k "chain" -s > gpurun_out/chain_tests.log 2>&1; grep -aE "^chain|violations|passed|failed|Error|error" gpurun_out/chain_tests.log | cut -c1-300 | tail -40
echo "== conv tests (all) + network + baseline shapes + zoo"
timeout -s KILL 900 $PT tests/test_gpu_conv.py tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py tests/test_gpu_engine.py tests/test_gpu_zoo.py -m gpu -s 2>&1 | grep -aE "PARITY|stage-wise|passed|failed|Error|assert" | cut -c1-250 | tail -30
echo "== layer times (chained)"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/layer_times_chain.txt 2>&1; tail -1 gpurun_out/layer_times_chain.txt
echo "== A/B plan time"
for rnd in 1 2; do
  YB_NO_CHAIN=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
echo "== nms phases"
timeout -s KILL 120 python scripts/nms_phases.py 2>&1 | tail -2
echo "== ncu post kernels"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"letterbox|decode_rows|nms_image" -s 6 -c 3 -o gpurun_out/r2_post_v6 -f python scripts/one_step.py 4 > gpurun_out/ncu_post.log 2>&1; tail -2 gpurun_out/ncu_post.log
