#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider"
timeout -s KILL 900 $PT tests/test_gpu_letterbox.py tests/test_gpu_ingest.py tests/test_zz_letterbox_cv2.py tests/test_gpu_network.py -m gpu 2>&1 | tail -3
timeout -s KILL 300 python scripts/stage_times.py 2>&1 | tail -2
