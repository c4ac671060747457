#!/bin/bash
# N GPUs (argument): strong scaling of configs[1] (32 images in total) and configs[2] (yolov5m, 128 images in total, bf16)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
for mode in "--scaling strong" "--config c3"; do
  tag=$(echo $mode | tr -d ' -')
  timeout -s KILL 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $mode > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$tag.json').read().strip().splitlines()[-1])
    print('N=$N $tag', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'scaling', d['scaling'], 'img/gpu', d['config']['images_per_gpu_per_step'], 'e2e', round(d['e2e']['value'],1), 'stages', {k:(round(v['us'],1) if isinstance(v,dict) else v) for k,v in d['roofline_stages'].items() if isinstance(v,dict)})
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/bench_n${N}_$tag.err').read()[-1500:])
PY
done
