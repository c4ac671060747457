#!/bin/bash
# A/B two builds on this box, install the faster one in-tree, then run the round-end validation with it.
#   bash scripts/gpu_pick_final.sh scratch/lib_a.so scratch/lib_b.so
mkdir -p gpurun_out
bash scripts/ab_step.sh "$@" | tee gpurun_out/ab.txt
pick=$(python - "$@" <<'PY'
import re, sys
best, best_ms = None, 1e9
for lib in sys.argv[1:]:
    name = lib.split("/")[-1]
    ms = [float(v) for line in open("gpurun_out/ab.txt") if line.startswith(name) for v in line.split("plan ms")[1].split()]
    m = sum(ms) / max(len(ms), 1) if ms else 1e9
    if m < best_ms:
        best, best_ms = lib, m
print(best)
PY
)
echo "picked $pick" | tee gpurun_out/picked.txt
cp "$pick" yolort_b200/libyolort_b200.so
bash scripts/gpu_final.sh
