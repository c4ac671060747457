#!/bin/bash
PT="python -m pytest -q -p no:cacheprovider -x"
timeout -s KILL 600 $PT tests/test_gpu_postprocess.py tests/test_gpu_logits_decoder.py tests/test_gpu_network.py tests/test_p6.py -m gpu 2>&1 | tail -3
timeout -s KILL 120 python scripts/nms_phases.py 2>&1 | tail -1
timeout -s KILL 300 python scripts/stage_times.py 2>&1 | tail -2 | head -1
