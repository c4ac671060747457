#!/bin/bash
# Source-level ncu captures (one step of yolov5s batch 32 640x640): post kernels, the first patch-kernel launches
# (stem, body.1, body.2 F3 chain, body.3, body.4 F2/F3 chains, body.6.m.0.cv2), the first im2col/1x1 launches.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout -s KILL 300 $NCU -k regex:"nms_image|decode_rows" -s 4 -c 2 -o gpurun_out/r2_post_v7 -f python scripts/one_step.py 4 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout -s KILL 300 $NCU -k regex:"conv3x3_patch" -s 28 -c 7 -o gpurun_out/r2_patch_v7 -f python scripts/one_step.py 4 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
timeout -s KILL 300 $NCU -k regex:"conv_umma" -s 62 -c 4 -o gpurun_out/r2_umma_v7 -f python scripts/one_step.py 4 > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
timeout -s KILL 120 python scripts/nms_phases.py 2>&1 | tail -1
ls -la gpurun_out/*.ncu-rep
