#!/bin/bash
# round 2, first GPU call: parity tests, the bench line with the fused sweep (8 / 6 rows per CTA) and without it, ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02a_smoke.log
python bench.py --steps 8 --warmup 4 > gpurun_out/r02a_bench_fused8.json 2> gpurun_out/r02a_bench_fused8.err; echo "bench rc=$?"
RUVNB_LLP1_NW=6 python bench.py --steps 8 --warmup 4 --no-full-fit --no-e2e --no-cpu-baseline > gpurun_out/r02a_bench_fused6.json 2> gpurun_out/r02a_bench_fused6.err; echo "bench6 rc=$?"
RUVNB_NO_FUSE=1 python bench.py --steps 8 --warmup 4 --no-full-fit --no-e2e --no-cpu-baseline > gpurun_out/r02a_bench_nofuse.json 2> gpurun_out/r02a_bench_nofuse.err; echo "nofuse rc=$?"
python - <<'PY'
import json
for n in ("fused8","fused6","nofuse"):
    try:
        d=json.loads(open(f"gpurun_out/r02a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), "gene-fits/s", round(d["ms_per_step"],2), "ms/step", {k:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()}, d.get("full_fit",{}).get("wall_seconds"), d.get("e2e",{}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
P="--workload prof_2kx100k --steps 2 --warmup 3 --no-full-fit --no-e2e --no-cpu-baseline"
python bench.py $P > gpurun_out/r02a_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ll_pass1|k_pass2' -s 4 -c 4 -o gpurun_out/r02a_prof python bench.py $P > gpurun_out/r02a_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -15
