#!/bin/bash
# Round-end validation: the driver's own test command, smoke, bench line, per-layer times, stage times, ncu launch list,
# epilogue phase profile is separate (needs the instrumented library).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
grep -aE "^FAILED|^E  " gpurun_out/pytest_gpu.log | head -8 | cut -c1-300
timeout -s KILL 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log | cut -c1-200)"
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/layer_times.txt 2>&1; tail -1 gpurun_out/layer_times.txt
timeout -s KILL 300 python scripts/stage_times.py 2>&1 | tail -2
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precondition 0 > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"letterbox|decode_rows|nms_image" -s 6 -c 3 -o gpurun_out/r2_post_final -f python scripts/one_step.py 4 > gpurun_out/ncu_post.log 2>&1; tail -1 gpurun_out/ncu_post.log
for c in c4 c5; do timeout -s KILL 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; tail -1 gpurun_out/bench_$c.log > gpurun_out/bench_$c.json; cut -c1-200 gpurun_out/bench_$c.json; done
ls -la gpurun_out | grep -E "bench|launches|layer|ncu-rep"
