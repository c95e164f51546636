#!/bin/bash
# round 2, call w: k_ll_w with four cells per lane (RUVNB_LLW = 1: 6 rows per CTA / 160 registers, 2: 8 rows per CTA / 128 registers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-full-fit --no-e2e --no-cpu-baseline"
for v in 0 1 2; do
  RUVNB_LLW=$v timeout 600 python bench.py $B > gpurun_out/r02w_bench_llw$v.json 2> gpurun_out/r02w_bench_llw$v.err; echo "llw $v rc=$?"
done
for v in 1 2; do
  RUVNB_LLW=$v timeout 600 python -m pytest tests/test_gpu_irls.py tests/test_golden.py tests/test_gpu_batch.py -m gpu -q -x > gpurun_out/r02w_pytest_llw$v.log 2>&1; echo "pytest llw $v rc=$?"; tail -2 gpurun_out/r02w_pytest_llw$v.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02w_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], round(d["value"]), "gene-fits/s", round(d["ms_per_step"],2), "ms/step", {k[:8]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
