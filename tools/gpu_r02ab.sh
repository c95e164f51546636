#!/bin/bash
# round 2, call ab: width of a fresh certificate hint window (RUVNB_HINT_HALF) against the exact-MAD rows / time of a whole fit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for h in 0.4 0.55 0.7 0.8; do
RUVNB_HINT_HALF=$h timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02ab_fullfit_h$h.json 2> gpurun_out/r02ab_fullfit_h$h.err; echo "h=$h rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ab_fullfit_h*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); ff=d["full_fit"]
    print(f.split("_h")[1], "fit %.3f s inner %.3f"%(ff["wall_seconds"], ff["inner_loop_seconds"]), "exact rows", ff["exact_mad_rows"], "exact ms", ff["inner_loop_brackets"].get("exact_mad"), "bench exact rows", d["exact_mad_rows_per_step"], "ms/step %.2f"%d["ms_per_step"], "logl", ff["logl_outer"][-1])
PY
