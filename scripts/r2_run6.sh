#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
echo "== conv + engine tests"
timeout -s KILL 600 $PT tests/test_gpu_conv.py tests/test_gpu_engine.py -m gpu -x 2>&1 | tail -4
echo "== step"
timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
timeout -s KILL 200 python scripts/layer_times.py > gpurun_out/r2_layers_run6.txt 2>&1; tail -1 gpurun_out/r2_layers_run6.txt
echo "== bench c2"
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_run6.json 2> gpurun_out/bench_c2_run6.err; tail -c 3000 gpurun_out/bench_c2_run6.json; tail -3 gpurun_out/bench_c2_run6.err
echo "== bench c3 c4 c5"
for c in c3 c4 c5; do
  timeout -s KILL 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${c}_run6.json 2> gpurun_out/bench_${c}_run6.err; tail -c 1500 gpurun_out/bench_${c}_run6.json; tail -3 gpurun_out/bench_${c}_run6.err
done
echo "== ncu 1x1 layer with source"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 116 -c 1 -o gpurun_out/r2_1x1_body2cv3 -f python scripts/one_step.py 4 > gpurun_out/ncu_1x1.log 2>&1; tail -2 gpurun_out/ncu_1x1.log
