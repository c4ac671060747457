#!/bin/bash
# DRAM traffic + duration of every conv launch of one step (3 metrics, cheap), and the 2-GPU check is separate.
mkdir -p gpurun_out
timeout -s KILL 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:"conv_umma|conv3x3_patch" -s 156 -c 52 --csv --log-file gpurun_out/conv_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precondition 0 > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/ncu_traffic.log | cut -c1-200
