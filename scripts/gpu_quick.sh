#!/bin/bash
# Fast iteration: conv + network parity in one process each, ablation table, bench line + launch list.
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "not view_modes or 0]" > gpurun_out/q_conv.log 2>&1; echo "conv rc=$? $(tail -1 gpurun_out/q_conv.log)"
timeout -s KILL 400 python -m pytest tests/test_gpu_network.py tests/test_gpu_pool_upsample.py -q -m gpu > gpurun_out/q_net.log 2>&1; echo "net rc=$? $(tail -1 gpurun_out/q_net.log)"
if [ "$1" != "noabl" ]; then python scripts/conv_ablation.py 2>&1 | tee gpurun_out/ablation.txt; fi
bash scripts/gpu_bench_profile.sh $2 | head -3 | cut -c1-400
