#!/bin/bash
# round 2, call ak: row parts for the cached scoring sweeps of the dispersion step (few rows: stragglers, small gene shards)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02ak_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ak_pytest_gpu.log
python tools/trace_disp.py > gpurun_out/r02ak_trace_disp.txt 2>&1; grep -E "sweep [4-9]:|estimateDisp" gpurun_out/r02ak_trace_disp.txt | tail -12
for p in 0 1; do
RUVNB_NO_DISP_PARTS=$p timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02ak_fit_noparts$p.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/r02ak_fit_noparts$p.json').read().strip().splitlines()[-1]);ff=d['full_fit'];print('noparts=$p fit',round(ff['wall_seconds'],3),'inner',round(ff['inner_loop_seconds'],3),'disp',round(ff['dispersion_seconds'],3),ff['logl_outer'])"
done
