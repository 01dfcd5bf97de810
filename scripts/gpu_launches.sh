#!/bin/bash
# per-launch device times of a short bench (ncu launch list; compare SHARES / per-kernel means, not absolutes)
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --batch ${BATCH:-592} --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c ${COUNT:-300} --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --batch ${BATCH:-592} --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches.csv 2>/dev/null | head -12
python - <<'PY'
import csv
rows = list(csv.DictReader([l for l in open("gpurun_out/launches.csv") if l.startswith('"')]))
v = [float(r["Metric Value"].replace(",", "")) / 1e3 for r in rows if "dan_stack_kernel" in r["Kernel Name"]]
print("stack launches (us):", [round(x) for x in v[:12]])
PY
