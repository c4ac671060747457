#!/bin/bash
# round 2, GPU call 1: banded stem (never run before), baseline per-layer table, ncu captures of post kernels + deep layers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== band stem correctness"
YB_STEM_BAND=1 timeout -s KILL 300 python -m pytest tests/test_gpu_network.py -q -m gpu -x 2>&1 | tail -3
echo "== band A/B"
for v in 0 1 0 1; do
  YB_STEM_BAND=$v timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1 | sed "s/^/band=$v /"
done
echo "== layer times"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/r2_layers_base.txt 2>&1; tail -3 gpurun_out/r2_layers_base.txt
YB_STEM_BAND=1 timeout -s KILL 200 python scripts/layer_times.py 2>&1 | head -3
echo "== ncu post kernels"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"letterbox|decode_candidates|nms_image" -s 9 -c 3 -o gpurun_out/r2_post_base -f python scripts/one_step.py 4 > gpurun_out/ncu_post.log 2>&1; tail -2 gpurun_out/ncu_post.log
echo "== ncu deep convs"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 141 -c 4 -o gpurun_out/r2_deep_base -f python scripts/one_step.py 4 > gpurun_out/ncu_deep.log 2>&1; tail -2 gpurun_out/ncu_deep.log
ls -la gpurun_out | tail
