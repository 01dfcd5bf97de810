#!/bin/bash
# development: stack-kernel component ceilings, timed with CUDA events (conv_stack class of bench.py).
# debug bit 0 = no MMAs, bit 1 = no epilogue work, bit 2 = no T stores
for d in ${MODES:-0 1 2 3 4}; do
DAN_B200_STACKDEBUG=$d timeout 300 python bench.py --steps 2 --warmup 3 --batch 1024 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']
print('debug=$d conv_stack ms/step', round(r['class_ms_per_step']['conv_stack'],2), 'per launch', round(r['avg_launch_ms'],4))"
done
