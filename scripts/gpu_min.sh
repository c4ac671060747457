#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_network.py -q -m gpu -x > gpurun_out/q_net.log 2>&1; echo "net rc=$? $(tail -1 gpurun_out/q_net.log)"
timeout -s KILL 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json
