#!/bin/bash
# development: board power / SM clock / conv-stack time of the bench under the stack kernel's debug modes
# (bit 0 = no MMAs, bit 1 = no epilogue work, bit 2 = no T stores, bit 4 = no weight streaming)
for d in ${MODES:-0 1 2 16 18}; do
DAN_B200_STACKDEBUG=$d timeout 300 python bench.py --steps ${STEPS:-6} --warmup 3 --batch 4144 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']; c=d['clocks']
print('debug=$d stack ms/step', round(r['class_ms_per_step']['conv_stack'],2), 'step ms', round(d['ms_per_step'],2), 'power W', c.get('power_w'), 'mhz', c['sm_mhz'], c['reasons'])"
done
