#!/bin/bash
# round 2, call m: the default bench line (e2e, full fit, CPU baseline) after the sweep split, and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02m_bench_default.json 2> gpurun_out/r02m_bench_default.err; echo "default rc=$?"; tail -3 gpurun_out/r02m_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02m_bench_reference.json 2> gpurun_out/r02m_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02m_bench_default.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d.get("e2e"), {k: (v if not isinstance(v,(list,dict)) else "...") for k,v in (d.get("full_fit") or {}).items()})
print(json.dumps(d["roofline"])[:1500])
print(d.get("cpu_baseline"))
print(open("gpurun_out/r02m_bench_reference.json").read()[:600])
PY
