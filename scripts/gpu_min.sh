#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_network.py -q -m gpu -x > gpurun_out/q_net.log 2>&1; echo "net rc=$? $(tail -1 gpurun_out/q_net.log)"
timeout -s KILL 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
