#!/bin/bash
# round 2, call e: ncu of the get.res kernels on the small cell workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="--workload cells_small_3kx40k --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $P > gpurun_out/r02e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_res_elem|k_res_caps' -s 6 -c 4 -o gpurun_out/r02e_res python bench.py $P > gpurun_out/r02e_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | grep r02e
