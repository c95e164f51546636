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


def test_shard_ranges_cover_and_refuse_more_ranks_than_rows():
    for G, world in [(20000, 8), (7, 7), (17, 4), (16, 4)]:
        r = [shard.gene_range(G, world, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == G
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert sum(b - a for a, b in r) == G and all(b > a for a, b in r)
    # a rank without rows cannot hold a context (ruvnb_set_design refuses g0 >= g1): an explicit error, not an empty range
    for G, world, rank in [(7, 8, 7), (1, 2, 1), (9, 8, 5)]:
        with pytest.raises(ValueError, match="without a row"):
            shard.gene_range(G, world, rank)


def _cells_worker(rank, world, port, ret):
    """Host side of a cell-sharded run: block-aligned cell ranges, the gathered annotation, the subsample's
    columns assembled from every rank's piece, the padded all-gather of the alpha slices, and the Winsorize quantile from
    per-shard histograms summed over the ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ruviiinb_b200 import api
        from oracle import rstats
        rng = np.random.default_rng(0)
        G, N, bs, k = 23, 1337, 100, 3
        Y = rng.poisson(3.0, (G, N)).astype(np.int32)
        Y[2] = rng.poisson(6000.0, N)                      # quantile beyond the 4096 direct bins
        ranks = api.Ranks(dist)
        c0, c1 = api.cell_shard_range(N, world, rank, bs)
        ok = c0 % bs == 0 and (c1 == N or c1 % bs == 0)
        spans = ranks.allgather((c0, c1))
        ok = ok and spans[0][0] == 0 and spans[-1][1] == N and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        # the subsample's columns: every rank contributes its own, the union is in subsample order
        sub = np.sort(rng.choice(N, 60, replace=False))
        mine = (sub >= c0) & (sub < c1)
        parts = ranks.allgather((np.where(mine)[0], Y[:, sub[mine]]))
        Ysub = np.empty((G, sub.size), dtype=np.int32)
        for idx, cc in parts:
            Ysub[:, idx] = cc
        ok = ok and np.array_equal(Ysub, Y[:, sub])
        # alpha slices of ceil(G / world) rows, zero-padded, with the non-finite count in the last slot (irls.cu k_pack_alpha)
        alpha = rng.standard_normal((G, k))
        g0, g1 = shard.gene_range(G, world, rank)
        per = -(-G // world)
        sl = np.zeros(per * k + 1); sl[:(g1 - g0) * k] = alpha[g0:g1].ravel(); sl[-1] = float(rank == 1)
        got = [torch.zeros(per * k + 1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(got, torch.tensor(sl))
        full = np.concatenate([g.numpy()[:-1] for g in got])[:G * k].reshape(G, k)
        ok = ok and np.array_equal(full, alpha) and sum(float(g[-1]) for g in got) == 1.0
        # Winsorize cap = ceil(type-7 quantile at 0.995) from summed histograms (winsor.cuh k_wz_hist / k_wz_pick_direct + descent)
        hist = np.stack([np.bincount(np.minimum(Y[g, c0:c1], 4095), minlength=4096) for g in range(G)])
        t = torch.tensor(hist); dist.all_reduce(t); hist = t.numpy()
        index = 1.0 + (N - 1) * 0.995
        klo, khi, frac = int(np.floor(index)) - 1, int(np.ceil(index)) - 1, index - np.floor(index)
        caps = np.empty(G)
        for g in range(G):
            cum = np.cumsum(hist[g])
            a, b = int(np.searchsorted(cum, klo + 1)), int(np.searchsorted(cum, khi + 1))
            if b < 4095:
                caps[g] = np.ceil(a if (frac == 0 or a == b) else (1 - frac) * a + frac * b)
            else:   # radix descent on the full value, one level's histogram summed per step
                vals = []
                for rk in (klo, khi):
                    prefix, pmask = 0, 0
                    for shift, bits in ((20, 12), (8, 12), (0, 8)):
                        v = Y[g, c0:c1].astype(np.int64)
                        sel = (v & pmask) == prefix
                        h = torch.tensor(np.bincount((v[sel] >> shift) & ((1 << bits) - 1), minlength=1 << bits)); dist.all_reduce(h)
                        cum = np.cumsum(h.numpy()); bsel = int(np.searchsorted(cum, rk + 1))
                        rk -= int(cum[bsel - 1]) if bsel > 0 else 0
                        prefix |= bsel << shift; pmask |= ((1 << bits) - 1) << shift
                    vals.append(prefix)
                a, b = vals
                caps[g] = np.ceil(a if (frac == 0 or a == b) else (1 - frac) * a + frac * b)
        ref = np.ceil(np.array([rstats.r_quantile7(Y[g].astype(np.float64), [0.995])[0] for g in range(G)]))
        ok = ok and np.array_equal(caps, ref) and caps[2] > 4095
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_cell_shard_host_logic():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_cells_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_cell_shard_ranges_are_block_aligned():
    from ruviiinb_b200 import api
    for N, world, bs in [(1000000, 8, 5000), (1337, 2, 100), (300, 4, 256), (5000, 3, 5000)]:
        r = [api.cell_shard_range(N, world, k, bs) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == N and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(a % bs == 0 or a == N for a, _ in r)   # (ranks beyond the last block hold no cell)
