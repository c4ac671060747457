#!/bin/bash
# Runs the GPU parity tests file by file (each in its own process, each under a timeout) and a short bench.
# Usage (on the GPU box, from the repo root): bash scripts/gpu_tests.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(grep -E 'passed|failed|error' gpurun_out/$name.log | tail -1)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run letterbox 300 python -m pytest tests/test_gpu_letterbox.py -q -m gpu -s
run postprocess 300 python -m pytest tests/test_gpu_postprocess.py -q -m gpu -s
for m in 0 2; do
  run conv_patch_mode$m 240 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "test_patch_conv_view_modes and ${m}]"
done
run conv_patch_misc 240 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "test_patch_conv_ragged or test_patch_conv_matches"
for t in test_conv1x1 test_conv1x1_ragged test_conv3x3 test_conv3x3_crosses test_bottleneck test_head_conv test_wide_output test_bf16 test_large_m test_rejects test_r31_activations; do
  run conv_$t 240 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "$t"
done
run pool 300 python -m pytest tests/test_gpu_pool_upsample.py -q -m gpu -s
run zoo 300 python -m pytest tests/test_gpu_zoo.py -q -m gpu -s
run p6 300 python -m pytest tests/test_p6.py -q -m gpu -s
run v4 300 python -m pytest tests/test_v4.py -q -m gpu -s
run ingest 300 python -m pytest tests/test_gpu_ingest.py -q -m gpu -s
run logits_decoder 300 python -m pytest tests/test_gpu_logits_decoder.py -q -m gpu -s
run network 240 python -m pytest tests/test_gpu_network.py -q -m gpu -s
run smoke 300 python __graft_entry__.py smoke
if [ "$1" != "quick" ]; then
  run bench 300 python bench.py --steps 20 --warmup 5
  tail -1 gpurun_out/bench.log > gpurun_out/bench.json
fi
cat gpurun_out/summary.txt
