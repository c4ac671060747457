#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider"
echo "== letterbox + postprocess + ingest + network tests"
timeout -s KILL 900 $PT tests/test_gpu_letterbox.py tests/test_gpu_postprocess.py tests/test_gpu_logits_decoder.py tests/test_gpu_ingest.py tests/test_zz_letterbox_cv2.py tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py tests/test_p6.py -m gpu 2>&1 | tail -4
echo "== stage times"
timeout -s KILL 300 python scripts/stage_times.py 2>&1 | tail -2
timeout -s KILL 120 python scripts/nms_phases.py 2>&1 | tail -1
