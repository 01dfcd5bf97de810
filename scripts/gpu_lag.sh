#!/bin/bash
for lag in ${LAGS:-1 2 3 5 7}; do
  DAN_B200_LAG=$lag timeout 300 python bench.py --steps 3 --warmup 3 --batch 2072 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('lag $lag value %.0f stack ms %.2f frac %.3f clocks %s' % (d['value'], r['class_ms_per_step']['conv_stack'], r['frac'], d['clocks']['sm_mhz']))"
done
