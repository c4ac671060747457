#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
echo "== conv tests"
timeout -s KILL 900 $PT tests/test_gpu_conv.py -m gpu -x 2>&1 | tail -15
echo "== layer times"
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/r2_layers_run4.txt 2>&1; tail -2 gpurun_out/r2_layers_run4.txt
echo "== step"
timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
echo "== network / baseline shapes / engine / postprocess / letterbox tests"
timeout -s KILL 1500 $PT tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py tests/test_gpu_engine.py tests/test_gpu_postprocess.py tests/test_gpu_letterbox.py -m gpu -s > gpurun_out/tests_run4.log 2>&1; tail -12 gpurun_out/tests_run4.log; grep -aE "^PARITY|^\.?PARITY|stage-wise" gpurun_out/tests_run4.log | cut -c1-230
echo "== layer times m/x"
timeout -s KILL 200 python scripts/layer_times.py yolov5m 16 640 10 bf16 > gpurun_out/r2_layers_m_run4.txt 2>&1; tail -1 gpurun_out/r2_layers_m_run4.txt
timeout -s KILL 300 python scripts/layer_times.py yolov5x 8 1280 5 f16 > gpurun_out/r2_layers_x_run4.txt 2>&1; tail -1 gpurun_out/r2_layers_x_run4.txt
