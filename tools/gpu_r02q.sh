#!/bin/bash
# round 2, call q: A/B of build variants of the sweeps (constants from constant memory, warp-uniform fast branch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp ruviiinb_b200/libruvnb_b200.so /tmp/lib_base.so
B="--steps 20 --warmup 5 --no-full-fit --no-e2e --no-cpu-baseline"
for v in base v1 v2 v12; do
  if [ $v = base ]; then cp /tmp/lib_base.so ruviiinb_b200/libruvnb_b200.so; else cp ruviiinb_b200/variants/lib_$v.so ruviiinb_b200/libruvnb_b200.so; fi
  timeout 600 python bench.py $B > gpurun_out/r02q_bench_$v.json 2> gpurun_out/r02q_bench_$v.err; echo "$v rc=$?"
done
timeout 600 python -m pytest tests/test_gpu_irls.py tests/test_golden.py tests/test_gpu_batch.py -m gpu -q -x > gpurun_out/r02q_pytest_v12.log 2>&1; echo "pytest v12 rc=$?"; tail -3 gpurun_out/r02q_pytest_v12.log
cp /tmp/lib_base.so ruviiinb_b200/libruvnb_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02q_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], round(d["value"]), "gene-fits/s", round(d["ms_per_step"],2), "ms/step", {k[:8]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
