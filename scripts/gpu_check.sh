#!/bin/bash
# GPU iteration: parity tests (bounded), then an optional short bench. K=<pytest -k expr> BENCH=0/1 STEPS BATCH
set -u
mkdir -p gpurun_out
timeout ${TMO:-600} python -m pytest tests -m gpu -x -q ${K:+-k "$K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -${TAILN:-30} gpurun_out/pytest_gpu.log
if [ "${BENCH:-1}" = "1" ]; then
  timeout 300 python bench.py --steps ${STEPS:-5} --warmup 3 --batch ${BATCH:-4144} --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("value %.0f e2e %.0f ms/step %.2f | stack frac %.3f whole %.3f | class ms %s | clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["frac"], r["whole_forward_frac"], {k: round(v, 2) for k, v in r["class_ms_per_step"].items()}, d["clocks"]))
except Exception as e:
    print("bench parse failed", e)
PY
  tail -5 gpurun_out/bench_quick.err
fi
