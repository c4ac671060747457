#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider"
echo "== chain F2/F3 determinism"
timeout -s KILL 300 $PT "tests/test_gpu_conv.py::test_chain_3x3_into_next_cv1_and_cv3" -m gpu -s > gpurun_out/chain2.log 2>&1; grep -aE "^chain|NONDET|store_first=0 differs|passed|failed" gpurun_out/chain2.log | cut -c1-400
echo "== p6 e2e: default / no chain no nsplit / noguard"
timeout -s KILL 300 $PT "tests/test_p6.py::test_gpu_end_to_end_vs_reference_fixture_p6" -m gpu -s 2>&1 | grep -aE "p6 e2e|passed|failed|differ" | cut -c1-300
YB_NO_CHAIN=1 YB_NO_NSPLIT=1 timeout -s KILL 300 $PT "tests/test_p6.py::test_gpu_end_to_end_vs_reference_fixture_p6" -m gpu -s 2>&1 | grep -aE "p6 e2e|passed|failed|differ" | cut -c1-300
YB_LIB_PATH=$PWD/scratch/lib_noguard.so YB_NO_CHAIN=1 YB_NO_NSPLIT=1 timeout -s KILL 300 $PT "tests/test_p6.py::test_gpu_end_to_end_vs_reference_fixture_p6" -m gpu -s 2>&1 | grep -aE "p6 e2e|passed|failed|differ" | cut -c1-300
echo "== v4"
timeout -s KILL 300 $PT tests/test_v4.py tests/test_p6.py -m gpu 2>&1 | tail -3
