#!/bin/bash
# Round-end validation: the driver's own test command, smoke, bench line, ncu launch list, full ncu capture of
# the conv kernels.  Everything under a timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout -s KILL 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log | cut -c1-160)"
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench.json; cut -c1-200 gpurun_out/bench.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precondition 0 > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"conv_umma|conv3x3_patch" -s 104 -c 8 \
  -o gpurun_out/conv_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precondition 0 > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out | grep -E "ncu-rep|bench.json|launches"
