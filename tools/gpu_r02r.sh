#!/bin/bash
# round 2, call r: W update with the control-gene counts staged through shared memory (cp.async, a gene tile x 64 cells at once)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_irls.py tests/test_golden.py tests/test_gpu_batch.py tests/test_gpu_zinb.py -m gpu -q -x > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02r_pytest.log
B="--steps 20 --warmup 5 --no-full-fit --no-e2e --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02r_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "gene-fits/s", round(d["ms_per_step"],2), "ms/step", {k[:10]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()})
PY
