#!/bin/bash
# development: stack-kernel component ceilings. debug 1 = no MMAs (epilogue + streaming only), 2 = no epilogue work (MMA + weight streaming only), 3 = streaming only
for d in 0 1 2 3; do
echo "== debug=$d"
DAN_B200_STACKDEBUG=$d DAN_B200_STACKPROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --batch 512 --no-cpu-baseline 2>&1 >/dev/null | grep stackprof | tail -2
done
