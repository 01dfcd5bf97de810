#!/bin/bash
# end-of-round evidence on one B200: full GPU test suite, smoke, the default bench line, the reference arm, the training line,
# then (only after those plain runs) the ncu launch list and the ncu --set full capture of the conv-stack kernel
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"
timeout 300 python bench.py --mode train --steps 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc=$?"
COUNT=400 scripts/gpu_launches.sh 2>&1 | tail -14
scripts/gpu_ncu_full.sh > /dev/null 2>&1; ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
for f in ("bench", "bench_reference_arm", "bench_train"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step") if k in d}, "e2e", d.get("e2e", {}).get("value"), "clocks", d.get("clocks"), "roofline frac", (d.get("roofline") or {}).get("frac"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "fp32", (d.get("fp32") or {}).get("value"), "train", (d.get("train") or {}).get("value"))
    except Exception as e:
        print(f, "parse failed", e)
PY
