#!/bin/bash
# development: per-role cycle counters of the conv-stack kernel (library built with `python -m dl4vc_b200.build --prof`)
DAN_B200_LIB=$PWD/dl4vc_b200/libdan_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --batch ${BATCH:-296} --no-cpu-baseline 2>&1 >/dev/null | grep stkprof | tail -${N:-8}
