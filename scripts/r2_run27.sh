#!/bin/bash
PT="python -m pytest -q -p no:cacheprovider -x"
echo "== conv tests (wide variant)"
timeout -s KILL 600 $PT tests/test_gpu_conv.py -m gpu 2>&1 | tail -2
echo "== network tests with the unfused plan (wide 1x1 layers in the network)"
YB_NO_CHAIN=1 timeout -s KILL 900 $PT tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py -m gpu 2>&1 | tail -2
echo "== A/B: chained (default) | unfused + wide | unfused, two groups"
for rnd in 1 2; do
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_NO_CHAIN=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_NO_CHAIN=1 YB_NO_WIDE=1 timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
YB_NO_CHAIN=1 timeout -s KILL 200 python scripts/layer_times.py 2>&1 | grep -E "body.2|body.4|layer_blocks.0|head.head.0|sum of"
