#!/bin/bash
# development: same-box A/B of library builds on the DEFAULT (long, power-capped) bench: value, clock, power; LIBS="a.so b.so"
for r in $(seq 1 ${REPS:-2}); do
for lib in $LIBS; do
DAN_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']; c=d['clocks']
print('$lib', 'value', round(d['value']), 'stack ms', round(r['class_ms_per_step']['conv_stack'],2), 'mhz', c['sm_mhz'], 'W', c.get('power_w'))"
done; done
