#!/bin/bash
mkdir -p gpurun_out
YB_PATCH_LOADER=1 timeout -s KILL 400 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "patch_conv and not 2]" > gpurun_out/q_conv_cpasync.log 2>&1; echo "cpasync conv rc=$? $(tail -1 gpurun_out/q_conv_cpasync.log)"
python scripts/conv_ablation.py 2>&1 | tee gpurun_out/ablation.txt
