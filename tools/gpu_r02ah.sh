#!/bin/bash
# round 2, call ah (8 GPUs): the driver's bench command at N = 8, 4, 2 with the final library (NCCL control slab, seeded hints, ...)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29570+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02ah_cfg2_${n}gpu.json 2> gpurun_out/r02ah_cfg2_${n}gpu.err; echo "N=$n rc=$?"
done
python - <<'PY'
import json
for n in (2,4,8):
    d=json.loads(open("gpurun_out/r02ah_cfg2_%dgpu.json"%n).read().strip().splitlines()[-1])
    print(n, "%.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["e2e"]["seconds"], "fit", d["full_fit"]["wall_seconds"], "inner %.3f disp %.3f"%(d["full_fit"]["inner_loop_seconds"], d["full_fit"]["dispersion_seconds"]), d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), d.get("cpu_baseline",{}).get("value"))
PY
