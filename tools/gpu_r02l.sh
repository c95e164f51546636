#!/bin/bash
# round 2, call l: slab kernels as mbarrier producer/consumer pipelines (no CTA barrier), stage shapes A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02l_pytest.log
RUVNB_SLAB_CFG=3 timeout 600 python -m pytest tests/test_gpu_irls.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02l_pytest_cfg3.log 2>&1; echo "pytest cfg3 rc=$?"; tail -3 gpurun_out/r02l_pytest_cfg3.log
B="--steps 8 --warmup 6 --no-full-fit --no-e2e --no-cpu-baseline"
for c in 0 1 2 3; do
RUVNB_SLAB_CFG=$c timeout 600 python bench.py $B > gpurun_out/r02l_bench_slabcfg$c.json 2> gpurun_out/r02l_bench_slabcfg$c.err; echo "cfg $c rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02l_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], round(d["value"]), "gene-fits/s", round(d["ms_per_step"],2), "ms/step", {k[:8]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()}, d["exact_mad_rows_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
P="--workload prof_2kx100k --steps 2 --warmup 4 --no-full-fit --no-e2e --no-cpu-baseline"
python bench.py $P > gpurun_out/r02l_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_gram|k_pass2w' -s 6 -c 4 -o gpurun_out/r02l_prof python bench.py $P > gpurun_out/r02l_ncu.log 2>&1; echo "ncu rc=$?"
