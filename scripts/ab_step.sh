#!/bin/bash
# A/B of library builds in one session (same box, interleaved): bash scripts/ab_step.sh scratch/lib_a.so scratch/lib_b.so ...
for rnd in 1 2; do
  for lib in "$@"; do
    YB_LIB_PATH=$PWD/$lib timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
  done
done
