#!/bin/bash
# one `ncu --set full` capture of the dominant kernel (both segments of one pass), after the same command ran clean without ncu
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 148 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dan_stack -s 2 -c 2 -f -o gpurun_out/stack_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
