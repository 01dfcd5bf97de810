#!/bin/bash
# ncu --set full capture of the conv-stack kernel (both segment launches of one 148-candidate pass); the plain run comes first
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 148 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dan_stack -s ${SKIP:-4} -c 2 -f -o gpurun_out/stack_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
