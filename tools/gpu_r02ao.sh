#!/bin/bash
# final check of the round: the whole GPU suite, smoke, the default bench line
set -x
python -m pytest tests/ -x -q -m gpu > gpurun_out/r02ao_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ao_pytest.log
tail -3 gpurun_out/r02ao_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02ao_bench.json 2> gpurun_out/r02ao_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02ao_bench.json
