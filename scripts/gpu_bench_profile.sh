#!/bin/bash
# Bench line + ncu launch list + one full ncu capture of the conv kernel (1 GPU).  Run under gpurun.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log > gpurun_out/bench.json
cat gpurun_out/bench.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
if [ "$1" == "full" ]; then
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 60 -c 8 \
    -o gpurun_out/conv_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out | head -30
