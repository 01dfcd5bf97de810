#!/bin/bash
# development: bench value for a few (batch, pass size) settings; ARGS="--batch 4144|--batch 4144 --pass-candidates 296|..."
IFS='|' read -ra SETS <<< "${ARGS}"
for r in $(seq 1 ${REPS:-2}); do
for a in "${SETS[@]}"; do
timeout 300 python bench.py --steps ${STEPS:-4} --warmup 3 --no-cpu-baseline $a 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']
print('$a', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'stack ms', round(r['class_ms_per_step']['conv_stack'],2), 'frac', round(r['frac'],3), 'mhz', d['clocks']['sm_mhz'])"
done; done
