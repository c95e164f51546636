#!/bin/bash
# A/B on one box: Horner vs Estrin evaluation of the exp / log polynomials (fastmath.cuh, -DFM_ESTRIN), then the parity tests on the Estrin build
B="--steps 20 --warmup 5 --no-cpu-baseline --no-full-fit --no-e2e"
python bench.py $B > gpurun_out/r02at_bench_horner.json 2> gpurun_out/r02at_horner.err; echo "horner rc=$?"
cp ruviiinb_b200/variants/estrin/libruvnb_b200.so ruviiinb_b200/libruvnb_b200.so
python bench.py $B > gpurun_out/r02at_bench_estrin.json 2> gpurun_out/r02at_estrin.err; echo "estrin rc=$?"
python - <<'P'
import json
for n in ("horner","estrin"):
    d=json.loads(open(f"gpurun_out/r02at_bench_{n}.json").read().strip().splitlines()[-1])
    print(n, d["ms_per_step"], {k.split(" ")[0]:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items() if v["ms"]>0.5})
P
timeout 100 python -m pytest tests/test_gpu_irls.py tests/test_golden.py tests/test_gpu_disp.py tests/test_gpu_getres.py -x -q -m gpu > gpurun_out/r02at_pytest_estrin.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02at_pytest_estrin.log
