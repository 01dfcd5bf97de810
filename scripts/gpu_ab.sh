#!/bin/bash
# development: same-box A/B of library builds (conv_stack class ms per step); LIBS="a.so b.so", alternated REPS times
for r in $(seq 1 ${REPS:-2}); do
for lib in $LIBS; do
DAN_B200_LIB=$PWD/$lib DAN_B200_STACKDEBUG=${DBG:-0} timeout 300 python bench.py --steps ${STEPS:-3} --warmup 3 --batch ${BATCH:-1024} --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']
print('$lib', 'conv_stack ms/step', round(r['class_ms_per_step']['conv_stack'],2), 'value', round(d['value']), 'mhz', d['clocks']['sm_mhz'])"
done; done
