#!/bin/bash
# round 2, call b: parity tests after the faillist fix; fused-sweep variants (rows per CTA x cell order) on the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02b_pytest.log
B="--steps 6 --warmup 4 --no-full-fit --no-e2e --no-cpu-baseline"
for v in "8 0" "6 0" "8 1" "6 1"; do set -- $v
  RUVNB_LLP1_NW=$1 RUVNB_LLP1_SEQ=$2 python bench.py $B > gpurun_out/r02b_bench_nw$1_seq$2.json 2> gpurun_out/r02b_bench_nw$1_seq$2.err; echo "nw=$1 seq=$2 rc=$?"
done
RUVNB_NO_FUSE=1 python bench.py $B > gpurun_out/r02b_bench_nofuse.json 2> gpurun_out/r02b_bench_nofuse.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02b_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], round(d["value"]), "gene-fits/s", round(d["ms_per_step"],2), "ms/step", {k[:8]:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
