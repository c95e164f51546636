#!/bin/bash
# round 2, call n: ingest / egress formats (CSC upload, sparse sink, file sink), best slab stage shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_formats.py -m gpu -q -x > gpurun_out/r02n_pytest_formats.log 2>&1; echo "formats rc=$?"; tail -25 gpurun_out/r02n_pytest_formats.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02n_pytest.log
B="--steps 20 --warmup 5 --no-full-fit --no-e2e --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02n_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "gene-fits/s", round(d["ms_per_step"],2), "ms/step", {k[:10]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()})
PY
