#!/bin/bash
# Only the tests added after the last full run (cheap validation): bash scripts/gpu_new.sh
mkdir -p gpurun_out; : > gpurun_out/summary_new.txt
for t in "$@"; do
  echo "=== $t" | tee -a gpurun_out/summary_new.txt
  timeout -s KILL 280 python -m pytest $t -q -m gpu -s > gpurun_out/new_$(basename $t .py).log 2>&1
  echo "rc=$? $(grep -E 'passed|failed|error' gpurun_out/new_$(basename $t .py).log | tail -1)" | tee -a gpurun_out/summary_new.txt
  grep -E "^FAILED|^E  |Error" gpurun_out/new_$(basename $t .py).log | head -12 | tee -a gpurun_out/summary_new.txt
done
