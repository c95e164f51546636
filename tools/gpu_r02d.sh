#!/bin/bash
# round 2, call d: the cell workloads -- small shape first (both pipelines), then BASELINE configs 5 and 4 at full size on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --workload cells_small_3kx40k --steps 2 --warmup 1 > gpurun_out/r02d_small_getres.json 2> gpurun_out/r02d_small_getres.err; echo "small getres rc=$?"; tail -3 gpurun_out/r02d_small_getres.err
BENCH_FAST=1 python bench.py --workload cells_small_3kx40k --steps 1 --warmup 0 > gpurun_out/r02d_small_fast.json 2> gpurun_out/r02d_small_fast.err; echo "small fast rc=$?"; tail -3 gpurun_out/r02d_small_fast.err
timeout 1500 python bench.py --workload cfg5_getres_30kx1M --steps 1 --warmup 1 > gpurun_out/r02d_cfg5_1gpu.json 2> gpurun_out/r02d_cfg5_1gpu.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/r02d_cfg5_1gpu.err
timeout 1500 python bench.py --workload cfg4_fast_30kx1M --steps 1 --warmup 0 > gpurun_out/r02d_cfg4_1gpu.json 2> gpurun_out/r02d_cfg4_1gpu.err; echo "cfg4 rc=$?"; tail -3 gpurun_out/r02d_cfg4_1gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.3g"%d["value"], d["unit"], "ms/step %.1f"%d["ms_per_step"], d["roofline"]["kernels_ms_per_block"], d.get("stages"), d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value"), "gen %.0fs"%d["datagen_seconds"])
    except Exception as e:
        print(f, "failed", e)
PY
