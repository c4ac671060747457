#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/tests_full.log 2>&1
echo "== summary"; tail -25 gpurun_out/tests_full.log
echo "== parity lines"; grep -E "^PARITY|^stage-wise|plan creation" gpurun_out/tests_full.log
echo "== layer times"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/r2_layers_run3.txt 2>&1; tail -2 gpurun_out/r2_layers_run3.txt
timeout -s KILL 200 python scripts/layer_times.py yolov5m 16 640 20 bf16 > gpurun_out/r2_layers_m_run3.txt 2>&1; tail -2 gpurun_out/r2_layers_m_run3.txt
echo "== step"
timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
echo "== post kernels launch times"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"letterbox|decode_|nms_image" -s 9 -c 3 --csv python scripts/one_step.py 4 2>&1 | grep -E "letterbox|decode|nms" | cut -c1-200
