#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider"
echo "== chain determinism + new multi-level postprocess test"
timeout -s KILL 300 $PT "tests/test_gpu_conv.py" -m gpu -k chain -s 2>&1 | grep -aE "NONDET|store_first=0 differs|passed|failed" | cut -c1-300
timeout -s KILL 300 $PT "tests/test_gpu_postprocess.py" -m gpu -k multi_level -s 2>&1 | grep -aE "candidates|passed|failed|^E  " | cut -c1-300 | tail -30
echo "== p6 debug"
timeout -s KILL 300 python scripts/debug_p6.py 2>&1 | tail -20 | cut -c1-400
