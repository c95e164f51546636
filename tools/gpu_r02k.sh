#!/bin/bash
# round 2, call j: k_gram and k_pass2w on the bulk-copy (TMA) slab skeleton: tests, headline line, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02k_pytest.log
B="--steps 8 --warmup 6 --no-full-fit --no-e2e --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/r02k_bench_tma.json 2> gpurun_out/r02k_bench_tma.err; echo "tma rc=$?"; tail -2 gpurun_out/r02k_bench_tma.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02k_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], round(d["value"]), "gene-fits/s", round(d["ms_per_step"],2), "ms/step", {k[:8]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()}, d["exact_mad_rows_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
P="--workload prof_2kx100k --steps 2 --warmup 4 --no-full-fit --no-e2e --no-cpu-baseline"
python bench.py $P > gpurun_out/r02k_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ll_w|k_gram|k_pass2w' -s 9 -c 6 -o gpurun_out/r02k_prof python bench.py $P > gpurun_out/r02k_ncu.log 2>&1; echo "ncu rc=$?"
