"""Gene-sharded fit on >= 2 GPUs (one process per GPU under torchrun, NCCL): same decisions and 1e-6 parity with the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("yctl", ["host", "nccl", "nccl+batch"])
def test_two_rank_fit_matches_oracle(yctl):
    """yctl: the replicated control-gene slab comes from a host copy on every rank, or from the owning ranks over NCCL; +batch: with
    a batch-specific dispersion (two batches) on the gene shards."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", {"host": "29517", "nccl": "29518"}.get(yctl, "29519"), os.path.join(ROOT, "tests", "mp_fit_check.py")]
    env = dict(os.environ, MP_FIT_YCTL=yctl.split("+")[0], MP_FIT_BATCH="1" if yctl.endswith("batch") else "0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
