#!/bin/bash
# round 2, call o (8 GPUs): the 2-rank GPU tests, then the driver's bench command at N = 1, 2, 4, 8 (strong scaling of cfg 2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cells.py -m gpu -q > gpurun_out/r02o_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -6 gpurun_out/r02o_pytest_multi.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02o_cfg2_1gpu.json 2> gpurun_out/r02o_cfg2_1gpu.err; echo "N=1 rc=$?"
for n in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02o_cfg2_${n}gpu.json 2> gpurun_out/r02o_cfg2_${n}gpu.err; echo "N=$n rc=$?"; tail -2 gpurun_out/r02o_cfg2_${n}gpu.err
done
python - <<'PY'
import json,glob
base=None
for n in (1,2,4,8):
    f="gpurun_out/r02o_cfg2_%dgpu.json"%n
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if n==1: base=(d["value"], d["e2e"]["value"])
        print(n, "%.4g"%d["value"], "eff %.3f"%(d["value"]/base[0]/n), "ms/step %.2f"%d["ms_per_step"], {k[:8]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()}, "e2e %.4g eff %.3f"%(d["e2e"]["value"], d["e2e"]["value"]/base[1]/n), "full_fit", (d.get("full_fit") or {}).get("wall_seconds"), "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("samples"))
    except Exception as e:
        print(f, "failed", e)
PY
