#!/bin/bash
mkdir -p gpurun_out
echo "== new tests (stats)"
timeout -s KILL 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_engine.py -m gpu -q -s 2>&1 | grep -E "PARITY|stage-wise|stage |passed|failed|FAILED|Error|error|plan creation|assert" | head -80
echo "== whole gpu suite"
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== smoke"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== ab"
timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
