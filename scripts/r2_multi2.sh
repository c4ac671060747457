#!/bin/bash
# 2 GPUs: hardware test of predict_sharded + non-current device, then weak / strong / c3 bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
PT="python -m pytest -q -p no:cacheprovider --timeout 600 --timeout-method thread"
timeout -s KILL 900 $PT tests/test_gpu_sharded.py "tests/test_gpu_engine.py::test_model_on_a_non_current_device" -m gpu 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for mode in "--scaling weak" "--scaling strong" "--config c3"; do
  tag=$(echo $mode | tr -d ' -')
  timeout -s KILL 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline $mode > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n2_$tag.json').read().strip().splitlines()[-1])
    print('$tag', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'scaling', d['scaling'], 'img/gpu', d['config']['images_per_gpu_per_step'], 'e2e', round(d['e2e']['value'],1))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/bench_n2_$tag.err').read()[-1500:])
PY
done
