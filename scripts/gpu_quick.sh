#!/bin/bash
# quick GPU iteration: bf16 parity tests first (bounded), then an optional short bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "${K:-bf16}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_quick.log
if [ "${BENCH:-1}" = "1" ]; then
  timeout 300 python bench.py --steps ${STEPS:-3} --warmup 3 --batch ${BATCH:-2048} --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
  tail -c 1800 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
fi
