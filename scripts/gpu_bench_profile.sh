#!/bin/bash
# Bench line + ncu launch list (+ optional full ncu captures).  Run under gpurun (1 GPU).
#   bash scripts/gpu_bench_profile.sh [conv|post|all]
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log > gpurun_out/bench.json
cat gpurun_out/bench.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precondition 0 > gpurun_out/ncu_launches.log 2>&1
if [ "$1" == "conv" ] || [ "$1" == "all" ]; then
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 60 -c 8 \
    -o gpurun_out/conv_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precondition 0 > gpurun_out/ncu_conv.log 2>&1
fi
if [ "$1" == "post" ] || [ "$1" == "all" ]; then
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"nms_image|decode_cand|letterbox" -s 3 -c 3 \
    -o gpurun_out/post_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precondition 0 > gpurun_out/ncu_post.log 2>&1
fi
ls -la gpurun_out | head -40
