#!/bin/bash
# round 2, call t: the driver's bench command with its ncu launch list, the reference arm, and the other BASELINE workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02t_bench_cfg2_1gpu.json 2> gpurun_out/r02t_bench_cfg2_1gpu.err; echo "cfg2 rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02t_bench_reference.json 2> gpurun_out/r02t_bench_reference.err; echo "reference rc=$?"
P="--steps 2 --warmup 8 --no-full-fit --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02t_ncu_launches_cfg2.csv python bench.py $P > gpurun_out/r02t_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 python bench.py --workload cfg3_zinb_10kx50k --steps 4 --warmup 3 > gpurun_out/r02t_cfg3.json 2> gpurun_out/r02t_cfg3.err; echo "cfg3 rc=$?"; tail -2 gpurun_out/r02t_cfg3.err
timeout 900 python bench.py --workload cfg1_standin_12kx9k --steps 4 --warmup 3 > gpurun_out/r02t_cfg1.json 2> gpurun_out/r02t_cfg1.err; echo "cfg1 rc=$?"; tail -2 gpurun_out/r02t_cfg1.err
timeout 1500 python bench.py --workload cfg4_fast_30kx1M --steps 1 --warmup 0 > gpurun_out/r02t_cfg4_1gpu.json 2> gpurun_out/r02t_cfg4_1gpu.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/r02t_cfg4_1gpu.err
timeout 1500 python bench.py --workload cfg5_getres_30kx1M --steps 1 --warmup 1 > gpurun_out/r02t_cfg5_1gpu.json 2> gpurun_out/r02t_cfg5_1gpu.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/r02t_cfg5_1gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02t_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4g"%d["value"], d["unit"], "ms/step %.1f"%d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), "full_fit", (d.get("full_fit") or {}).get("wall_seconds"), d.get("stages"))
    except Exception as e:
        print(f, "failed", e)
PY
