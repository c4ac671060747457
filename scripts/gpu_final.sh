#!/bin/bash
# Round-end validation: full GPU test suite, bench line, ncu launch list, full ncu capture of the conv kernels.
bash scripts/gpu_tests.sh quick
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench.json; cut -c1-200 gpurun_out/bench.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"conv_umma|conv3x3_patch" -s 104 -c 12 \
  -o gpurun_out/conv_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out | grep -E "ncu-rep|bench.json|launches"
