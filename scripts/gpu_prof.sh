#!/bin/bash
# development: per-role cycle counters of the fused stack kernel + a short bench
set -u
mkdir -p gpurun_out
DAN_B200_STACKPROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --batch 512 --no-cpu-baseline > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err; echo "prof rc=$?"
grep stackprof gpurun_out/prof_bench.err | tail -4
timeout 300 python bench.py --steps ${STEPS:-3} --warmup 3 --batch ${BATCH:-2048} --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'conv frac',round(r['frac'],3),'classes',{k:round(v,1) for k,v in r['class_ms_per_step'].items()}, 'clk',d['clocks'])
PY
tail -3 gpurun_out/bench_quick.err
