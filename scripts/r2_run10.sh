#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
echo "== tests: postprocess, letterbox, network, ingest"
timeout -s KILL 900 $PT tests/test_gpu_postprocess.py tests/test_gpu_letterbox.py tests/test_gpu_network.py tests/test_gpu_ingest.py tests/test_zz_letterbox_cv2.py -m gpu -x 2>&1 | tail -4
echo "== stages + predict"
timeout -s KILL 300 python scripts/stage_times.py 2>&1 | tail -3
echo "== f16x2 epilogue variant: stage-wise parity at the bench shape + conv tests, then A/B"
YB_LIB_PATH=$PWD/scratch/lib_f16x2.so timeout -s KILL 900 $PT tests/test_gpu_conv.py "tests/test_gpu_baseline_shapes.py::test_c2_yolov5s_bs32_640_fp16_every_launch_at_bench_shape" "tests/test_gpu_baseline_shapes.py::test_c2_yolov5s_bs32_640_fp16_detections_vs_oracle" "tests/test_gpu_baseline_shapes.py::test_c5_yolov5x_1280_fp16_every_launch" -m gpu -s 2>&1 | grep -aE "PARITY|stage-wise|passed|failed|violations" | head -20
for rnd in 1 2; do
  timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  YB_LIB_PATH=$PWD/scratch/lib_f16x2.so timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
done
