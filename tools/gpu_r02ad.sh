#!/bin/bash
# round 2, call ad: two-stage column sums; GPU tests + bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02ad_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02ad_pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02ad_bench_cfg2_1gpu.json 2> gpurun_out/r02ad_bench_cfg2_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02ad_bench_cfg2_1gpu.json").read().strip().splitlines()[-1])
print("%.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "e2e", d["e2e"]["value"], "fit", d["full_fit"]["wall_seconds"], {k[:12]:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()}, d["clocks"])
PY
