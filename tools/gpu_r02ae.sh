#!/bin/bash
# round 2, call ae: certificate hints seeded from a sample of every row's deviances (cold contexts): tests, whole fit, e2e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02ae_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02ae_pytest_gpu.log
for s in 0 1; do
RUVNB_NO_SEED=$s timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02ae_bench_noseed$s.json 2> gpurun_out/r02ae_bench_noseed$s.err; echo "noseed=$s rc=$?"
done
python - <<'PY'
import json
for s in (0,1):
    d=json.loads(open("gpurun_out/r02ae_bench_noseed%d.json"%s).read().strip().splitlines()[-1]); ff=d["full_fit"]
    print("noseed",s,"%.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "fit %.3f inner %.3f"%(ff["wall_seconds"], ff["inner_loop_seconds"]), "exact rows", ff["exact_mad_rows"], ff["inner_loop_brackets"].get("exact_mad"), d["exact_mad_rows_per_step"][:6], "logl", ff["logl_outer"][-1])
PY
