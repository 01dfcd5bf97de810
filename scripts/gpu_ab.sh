#!/bin/bash
# development: same-box A/B of library builds (conv_stack class ms per step); LIBS="a.so b.so", alternated REPS times
for r in $(seq 1 ${REPS:-2}); do
for lib in $LIBS; do
DAN_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps ${STEPS:-3} --warmup 3 --batch ${BATCH:-2072} --no-cpu-baseline --no-train --no-fp32 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib', 'conv_stack ms/step', round(r['class_ms_per_step']['conv_stack'],2), 'value', round(d['value']), 'mhz', d['clocks']['sm_mhz'], 'W', d['clocks']['power_w'])"
done; done
