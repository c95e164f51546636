#!/bin/bash
# round 2, call ac: what the driver runs at round end -- GPU tests, smoke(), the bench command and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02ac_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02ac_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02ac_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02ac_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02ac_bench_reference.json 2> gpurun_out/r02ac_bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02ac_bench_cfg2_1gpu.json 2> gpurun_out/r02ac_bench_cfg2_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py > gpurun_out/r02ac_bench_default.json 2> gpurun_out/r02ac_bench_default.err; echo "default rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02ac_bench_cfg2_1gpu.json","gpurun_out/r02ac_bench_default.json","gpurun_out/r02ac_bench_reference.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "%.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), "fit", (d.get("full_fit") or {}).get("wall_seconds"), "launches", d.get("gpu_launches"), "roofline frac", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("clocks") or {}))
PY
