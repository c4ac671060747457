mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_patch -s 36 -c 4 \
    -o gpurun_out/patch_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_patch.log 2>&1
tail -3 gpurun_out/ncu_patch.log | cut -c1-300
