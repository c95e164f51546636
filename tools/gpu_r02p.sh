#!/bin/bash
# round 2, call p: batch-specific dispersion + pseudo-cells, then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -q -x > gpurun_out/r02p_pytest_batch.log 2>&1; echo "batch rc=$?"; tail -40 gpurun_out/r02p_pytest_batch.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02p_pytest.log
