#!/bin/bash
# ncu launch list (per-launch gpu__time_duration) of one small bench step; the bench itself is run first without ncu.
set -u
mkdir -p gpurun_out
B=${BATCH:-296}
python bench.py --steps 1 --warmup 1 --batch $B --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-60} -c ${COUNT:-120} --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --batch $B --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; k = h.index('Kernel Name'); v = h.index('Metric Value'); u = h.index('Metric Unit')
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= v: continue
    name = r[k].split('(')[0][-60:]
    t = float(r[v].replace(',', '')); 
    if r[u] == 'ns': t /= 1e3
    elif r[u] == 'ms': t *= 1e3
    tot[name] += t; cnt[name] += 1
s = sum(tot.values())
for n, t in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{t:10.1f} us {100*t/s:5.1f}%  x{cnt[n]:4d}  avg {t/cnt[n]:8.1f} us  {n}")
PY
