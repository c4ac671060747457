#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
echo "== tests: postprocess, letterbox, network, ingest, zoo"
timeout -s KILL 900 $PT tests/test_gpu_postprocess.py tests/test_gpu_letterbox.py tests/test_gpu_network.py tests/test_gpu_ingest.py tests/test_gpu_logits_decoder.py tests/test_zz_letterbox_cv2.py -m gpu -x --deselect tests/test_gpu_postprocess.py::test_candidate_arena_grows_instead_of_truncating 2>&1 | tail -5
echo "== stages + predict"
timeout -s KILL 300 python scripts/stage_times.py 2>&1 | tail -3
echo "== ncu launch list (yb kernels of one step)"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_umma|conv3x3_patch|letterbox|decode_|nms_image|spp_pool|upsample2x|init_counters" -s 177 -c 59 --csv --log-file gpurun_out/r2_launches_final.csv python scripts/one_step.py 4 > /dev/null 2>&1; grep -c "gpu__time" gpurun_out/r2_launches_final.csv
echo "== ncu post kernels (full)"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"letterbox|decode_|nms_image" -s 9 -c 3 -o gpurun_out/r2_post_final -f python scripts/one_step.py 4 > gpurun_out/ncu_post.log 2>&1; tail -1 gpurun_out/ncu_post.log
