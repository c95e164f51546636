#!/bin/bash
# round 2, call s: ncu of the per-cell W update at cfg 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="--steps 2 --warmup 6 --no-full-fit --no-e2e --no-cpu-baseline"
python bench.py $P > gpurun_out/r02s_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_w_update_fast' -s 6 -c 1 -o gpurun_out/r02s_prof python bench.py $P > gpurun_out/r02s_ncu.log 2>&1; echo "ncu rc=$?"
