#!/bin/bash
# Round-2 first experiment: the banded stem variant (YB_STEM_BAND=1), never run on a GPU yet.
#   1. correctness: stage-wise + network parity tests with the variant on (each under a hard timeout: a wrong
#      expect_tx byte count would hang the CTA, not fail)
#   2. speed: A/B of the full step in one session
mkdir -p gpurun_out
YB_STEM_BAND=1 timeout -s KILL 120 python -m pytest tests/test_gpu_network.py -q -m gpu -x 2>&1 | tail -3
for v in 0 1 0 1; do
  YB_STEM_BAND=$v timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1 | sed "s/^/band=$v /"
done
