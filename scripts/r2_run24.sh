#!/bin/bash
mkdir -p gpurun_out
echo "== epilogue phase clocks (unfused plan)"
YB_LIB_PATH=$PWD/scratch/lib_epitime.so timeout -s KILL 200 python scripts/epi_timing.py 2>&1 | tail -40 | cut -c1-250
echo "== pool G sweep"
for g in 1 2 4 8; do
  echo -n "G=$g: "; YB_POOL_G=$g timeout -s KILL 200 python scripts/layer_times.py 2>&1 | grep -E "pool" | cut -c1-60
done
