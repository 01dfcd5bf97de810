#!/bin/bash
# development: real-mode event trace of CTA 0 (both segments' first launches) + per-role cycle counters
mkdir -p gpurun_out
DAN_B200_STACKTRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --batch ${BATCH:-296} --no-cpu-baseline > gpurun_out/trace_bench.json 2> gpurun_out/trace_bench.err; echo "trace rc=$?"
DAN_B200_STACKPROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --batch ${BATCH:-296} --no-cpu-baseline 2>&1 >/dev/null | grep stackprof | tail -4 > gpurun_out/stackprof.txt
cat gpurun_out/stackprof.txt
