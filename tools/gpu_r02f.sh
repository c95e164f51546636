#!/bin/bash
# round 2, call f: all GPU tests (2 GPUs) after the robust dispersion prior and the get.res kernel changes; cell workloads again
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02f_pytest_2gpu.log
CUDA_VISIBLE_DEVICES=0 python bench.py --workload cells_small_3kx40k --steps 2 --warmup 1 > gpurun_out/r02f_small_getres.json 2> gpurun_out/r02f_small_getres.err; echo "small rc=$?"; tail -3 gpurun_out/r02f_small_getres.err
CUDA_VISIBLE_DEVICES=0 timeout 1500 python bench.py --workload cfg5_getres_30kx1M --steps 1 --warmup 1 > gpurun_out/r02f_cfg5_1gpu.json 2> gpurun_out/r02f_cfg5_1gpu.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/r02f_cfg5_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload cfg5_getres_30kx1M --steps 1 --warmup 1 > gpurun_out/r02f_cfg5_2gpu.json 2> gpurun_out/r02f_cfg5_2gpu.err; echo "cfg5 2gpu rc=$?"; tail -3 gpurun_out/r02f_cfg5_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload cfg4_fast_30kx1M --steps 1 --warmup 0 > gpurun_out/r02f_cfg4_2gpu.json 2> gpurun_out/r02f_cfg4_2gpu.err; echo "cfg4 2gpu rc=$?"; tail -3 gpurun_out/r02f_cfg4_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 4 > gpurun_out/r02f_cfg2_2gpu.json 2> gpurun_out/r02f_cfg2_2gpu.err; echo "cfg2 2gpu rc=$?"; tail -3 gpurun_out/r02f_cfg2_2gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4g"%d["value"], d["unit"], "ms/step %.1f"%d["ms_per_step"], d["roofline"].get("kernels_ms_per_block") or {k[:8]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()}, d.get("stages"), "e2e", d.get("e2e",{}).get("value"), "full_fit", (d.get("full_fit") or {}).get("wall_seconds"))
    except Exception as e:
        print(f, "failed", e)
PY
