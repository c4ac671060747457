#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
echo "== tests: network (predict pipeline), postprocess, letterbox, engine, logits decoder, ingest"
timeout -s KILL 900 $PT tests/test_gpu_network.py tests/test_gpu_postprocess.py tests/test_gpu_letterbox.py tests/test_gpu_engine.py tests/test_gpu_logits_decoder.py tests/test_gpu_ingest.py -m gpu --durations=8 2>&1 | tail -22
echo "== step"
timeout -s KILL 120 python scripts/ab_step.py 40 2>&1 | tail -1
echo "== bench c2"
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_run7.json 2> gpurun_out/bench_c2_run7.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2_run7.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'sync', d['e2e']['sync_call']['value'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'plan_ms', d['roofline']['plan_ms'])
print('stages', {k:(v if not isinstance(v,dict) else {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items()}) for k,v in d['roofline_stages'].items()})
print('heavy', d.get('nms_heavy_load')); print('parity', d.get('parity'))
PY
tail -3 gpurun_out/bench_c2_run7.err
echo "== ncu launch list of one step"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 174 -c 59 --csv --log-file gpurun_out/r2_launches_final.csv python scripts/one_step.py 4 > /dev/null 2>&1; wc -l gpurun_out/r2_launches_final.csv
echo "== ncu post kernels (full)"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"letterbox|decode_|nms_image" -s 9 -c 3 -o gpurun_out/r2_post_final -f python scripts/one_step.py 4 > gpurun_out/ncu_post.log 2>&1; tail -1 gpurun_out/ncu_post.log
echo "== ncu deep convs (full): the 20x20 / 40x40 3x3 layers on the patch kernel"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_patch -s 36 -c 12 -o gpurun_out/r2_patch_final -f python scripts/one_step.py 4 > gpurun_out/ncu_patch.log 2>&1; tail -1 gpurun_out/ncu_patch.log
