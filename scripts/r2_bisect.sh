#!/bin/bash
PT="python -m pytest -q -p no:cacheprovider -x"
for lib in default nohoist noguard nohoist_noguard; do
  if [ $lib = default ]; then unset YB_LIB_PATH; else export YB_LIB_PATH=$PWD/scratch/lib_$lib.so; fi
  echo "== $lib: conv1x1 / patch / full conv file"
  timeout -s KILL 120 $PT "tests/test_gpu_conv.py::test_conv1x1" -m gpu 2>&1 | tail -1
  timeout -s KILL 120 $PT "tests/test_gpu_conv.py::test_patch_conv_channel_widths" -m gpu 2>&1 | tail -1
  timeout -s KILL 300 $PT tests/test_gpu_conv.py -m gpu 2>&1 | tail -1
done
unset YB_LIB_PATH
echo "== postprocess tests (new NMS sort/resolve/bitmatrix, decode compaction)"
timeout -s KILL 600 $PT tests/test_gpu_postprocess.py tests/test_gpu_logits_decoder.py -m gpu 2>&1 | tail -2
