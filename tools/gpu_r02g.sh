#!/bin/bash
# round 2, call g: all GPU tests (1 GPU) + the cfg3 / cfg1 / cfg4 lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02g_pytest.log
python bench.py --workload cfg3_zinb_10kx50k --steps 4 --warmup 3 > gpurun_out/r02g_cfg3.json 2> gpurun_out/r02g_cfg3.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/r02g_cfg3.err
python bench.py --workload cfg1_standin_12kx9k --steps 4 --warmup 3 > gpurun_out/r02g_cfg1.json 2> gpurun_out/r02g_cfg1.err; echo "cfg1 rc=$?"; tail -3 gpurun_out/r02g_cfg1.err
timeout 1500 python bench.py --workload cfg4_fast_30kx1M --steps 1 --warmup 0 > gpurun_out/r02g_cfg4_1gpu.json 2> gpurun_out/r02g_cfg4_1gpu.err; echo "cfg4 rc=$?"; tail -3 gpurun_out/r02g_cfg4_1gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4g"%d["value"], d["unit"], "ms/step %.1f"%d["ms_per_step"], d["roofline"].get("kernels_ms_per_block") or {k[:8]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()}, d.get("stages"), "e2e", d.get("e2e",{}).get("value"), "full_fit", (d.get("full_fit") or {}))
    except Exception as e:
        print(f, "failed", e)
PY
