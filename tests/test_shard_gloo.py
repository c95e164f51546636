"""N > 1 host logic on CPU: world_size-2 gloo processes shard genes / cell chunks and reproduce, with all-reduces of
per-shard partial sums, the quantities the GPU path exchanges: the global Gram/score sum behind
`alpha.mean` (R/ruvIIInb.R:351-354), the gathered alpha for the column clamp (:371-376), the disjoint W slices and the
scalar log-likelihood that makes every rank take the same accept/halve branch (:486-488)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ruviiinb_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, G, N, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)  # same data on every rank
        A = rng.standard_normal((G, k, k)); A = A @ A.transpose(0, 2, 1)
        b = rng.standard_normal((G, k))
        alpha = rng.standard_normal((G, k))
        ll = rng.standard_normal(G)
        g0, g1 = shard.gene_range(G, world, rank)
        # (1) all-reduce of sum_g [A_g | b_g]
        part = torch.tensor(np.concatenate([A[g0:g1].sum(0).ravel(), b[g0:g1].sum(0)]))
        dist.all_reduce(part)
        full = np.concatenate([A.sum(0).ravel(), b.sum(0)])
        ok1 = np.allclose(part.numpy(), full, rtol=1e-12)
        # (2) alpha slices summed into the zero-padded full matrix (what irls_body does with alpha_full)
        pad = np.zeros((G, k)); pad[g0:g1] = alpha[g0:g1]
        t = torch.tensor(pad); dist.all_reduce(t)
        ok2 = np.array_equal(t.numpy(), alpha)
        # (3) disjoint W slices by whole chunks
        nch = (N + 127) // 128
        c0, c1 = shard.chunk_range(nch, world, rank)
        Wfull = rng.standard_normal((nch * 128, k))
        Wpad = np.zeros_like(Wfull); Wpad[c0 * 128:c1 * 128] = Wfull[c0 * 128:c1 * 128]
        t = torch.tensor(Wpad); dist.all_reduce(t)
        ok3 = np.array_equal(t.numpy(), Wfull)
        # (4) the scalar every rank bases its accept/halve decision on
        s = torch.tensor([ll[g0:g1].sum()]); dist.all_reduce(s)
        gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, s)
        ok4 = all(float(x) == float(s) for x in gathered) and abs(float(s) - ll.sum()) < 1e-12 * max(1.0, abs(ll.sum()))
        # coverage of the shards
        r = torch.tensor([g0, g1, c0, c1]); allr = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, r)
        rows = sorted((int(x[0]), int(x[1])) for x in allr)
        ok5 = rows[0][0] == 0 and rows[-1][1] == G and all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        ret[rank] = bool(ok1 and ok2 and ok3 and ok4 and ok5)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("G,N", [(101, 1000), (4, 130)])
def test_gloo_world2_exchange_steps(G, N):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, G, N, 3, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_ranges_cover_and_handle_more_ranks_than_rows():
    for G, world in [(20000, 8), (7, 8), (1, 2), (16, 4)]:
        r = [shard.gene_range(G, world, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == G
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert sum(b - a for a, b in r) == G
