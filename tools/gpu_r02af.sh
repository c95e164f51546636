#!/bin/bash
# round 2, call af: `best` keeps the completed Huber scales (no repeated exact-MAD rows after rejected steps): tests + whole fit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02af_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02af_pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02af_bench.json").read().strip().splitlines()[-1]); ff=d["full_fit"]
print("%.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "fit %.3f inner %.3f disp %.3f"%(ff["wall_seconds"], ff["inner_loop_seconds"], ff["dispersion_seconds"]), "exact rows", ff["exact_mad_rows"], ff["inner_loop_brackets"].get("exact_mad"), ff["inner_loop_brackets"].get("pass1"), "logl", ff["logl_outer"])
PY
