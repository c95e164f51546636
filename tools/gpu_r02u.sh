#!/bin/bash
# round 2, call u (2 GPUs): the 2-rank tests (control slab from the host / over NCCL, cell shards), cfg 2 / cfg 4 / cfg 5 on two ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cells.py -m gpu -q > gpurun_out/r02u_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -6 gpurun_out/r02u_pytest_multi.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02u_cfg2_2gpu.json 2> gpurun_out/r02u_cfg2_2gpu.err; echo "cfg2 nccl-slab rc=$?"
RUVNB_BENCH_YCTL_HOST=1 timeout 600 $T --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-full-fit --no-cpu-baseline > gpurun_out/r02u_cfg2_2gpu_hostslab.json 2> gpurun_out/r02u_cfg2_2gpu_hostslab.err; echo "cfg2 host-slab rc=$?"
timeout 900 $T --master-port 29553 bench.py --gpus 2 --workload cfg5_getres_30kx1M --steps 1 --warmup 1 > gpurun_out/r02u_cfg5_2gpu.json 2> gpurun_out/r02u_cfg5_2gpu.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/r02u_cfg5_2gpu.err
timeout 900 $T --master-port 29554 bench.py --gpus 2 --workload cfg4_fast_30kx1M --steps 1 --warmup 0 > gpurun_out/r02u_cfg4_2gpu.json 2> gpurun_out/r02u_cfg4_2gpu.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/r02u_cfg4_2gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02u_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4g"%d["value"], d["unit"], "ms/step %.1f"%d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("h2d_bytes_per_step"), "full_fit", (d.get("full_fit") or {}).get("wall_seconds"), d.get("stages"))
    except Exception as e:
        print(f, "failed", e)
PY
