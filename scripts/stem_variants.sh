#!/bin/bash
# Compare stem formulations: bench plan_ms for each (short runs).
for cfg in "4 1" "4 0" "2 1" "2 0" "1 1"; do
  set -- $cfg
  YB_STEM_PACK=$1 YB_STEM_IM2COL=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pack=$1 im2col=$2', round(d['value']), d['ms_per_step'], d['roofline']['plan_ms'])"
done
