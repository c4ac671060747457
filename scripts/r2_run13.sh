#!/bin/bash
# Chained pointwise tails: unit tests, then the conv / network / baseline-shape suites, then timing with and without.
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider -x"
echo "== chain unit tests"
timeout -s KILL 300 $PT tests/test_gpu_conv.py -m gpu -k "chain" -s 2>&1 | grep -aE "^chain|passed|failed|Error|error|assert" | cut -c1-250 | tail -40
echo "== conv tests (all)"
timeout -s KILL 600 $PT tests/test_gpu_conv.py -m gpu 2>&1 | tail -3
echo "== network + baseline shapes + engine"
timeout -s KILL 900 $PT tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py tests/test_gpu_engine.py tests/test_gpu_zoo.py -m gpu -s 2>&1 | grep -aE "PARITY|stage-wise|passed|failed|violations|Error|assert" | cut -c1-250 | tail -30
echo "== layer times (chained)"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/layer_times_chain.txt 2>&1; tail -1 gpurun_out/layer_times_chain.txt
echo "== A/B plan time"
for rnd in 1 2; do
  YB_NO_CHAIN=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
