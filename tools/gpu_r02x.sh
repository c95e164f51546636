#!/bin/bash
# round 2, call x (8 GPUs): stream-segment timings of the iteration at N = 8 and N = 4 (what the non-sweep time is made of)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29560+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-full-fit --no-cpu-baseline > gpurun_out/r02x_cfg2_${n}gpu.json 2> gpurun_out/r02x_cfg2_${n}gpu.err; echo "N=$n rc=$?"
done
python - <<'PY'
import json
for n in (8,4):
    d=json.loads(open("gpurun_out/r02x_cfg2_%dgpu.json"%n).read().strip().splitlines()[-1])
    ks=d["roofline"]["kernels"]
    print(n, "%.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
    tot=0
    for k,v in ks.items():
        t=v["ms"]*v["launches_per_step"]; tot+=t
        print("   %-28s %.3f ms x %.2f"%(k[:28], v["ms"], v["launches_per_step"]))
    print("   sum of brackets %.3f ms"%tot)
PY
