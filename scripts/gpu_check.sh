#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, and the ncu launch list of the same bench command.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps ${STEPS:-5} --warmup 3 --batch ${BATCH:-4096} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
if [ "${NCU:-1}" = "1" ]; then
  python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  echo "ncu rc=$?"
fi
tail -5 gpurun_out/pytest_gpu.log
