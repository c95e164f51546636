"""Run under torchrun (one rank per GPU): gene-sharded whole NB fit against the single-process oracle.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_fit_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ruviiinb as orc  # noqa: E402
from ruviiinb_b200 import _lib, shard, synth  # noqa: E402
from tests.parity import relerr  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G, N, k, m, n_ctl = 203, 1300, 2, 3, 50
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=77, gmean_mu=0.5, gmean_sd=1.0)
    Yf = Y.astype(np.float64)
    W0 = orc.init_W_svd(Yf, ctl, k)
    rng = np.random.default_rng(78)
    sched = np.stack([truth["psi"] * np.exp(0.2 * rng.standard_normal(G) / (i + 1)) for i in range(3)])
    calls = {"i": 0}
    batched = os.environ.get("MP_FIT_BATCH") == "1"    # a batch-specific dispersion (two batches) on the gene shards
    batch = None
    if batched:
        batch = rng.integers(1, 3, size=N); batch[:2] = [1, 2]
        sched = np.stack([sched, sched * np.exp(0.4 * rng.standard_normal(sched.shape))], axis=2)    # (3, G, 2)

    def psi_fn(Yx, offs, wt, psi_old):
        i = calls["i"]
        calls["i"] += 1
        return sched[min(i // 2, 2)][:, i % 2] if batched else sched[min(i, 2)]

    trace = []
    ref = orc.ruvIII_nb(Yf, M, ctl, k=k, outer_maxit=3, W0=W0, psi_fn=psi_fn, winsorize=False, trace=trace, batch=batch, batch_disp=batched)
    uid = shard.broadcast_unique_id(dist, _lib.load(), device=dev)
    Yl, Yctl, (g0, g1) = shard.local_counts(Y, ctl, world, rank)
    opts = _lib.default_opts(k=k, outer_maxit=3)
    with _lib.Context(local) as ctx:
        ctx.comm_init(uid, rank, world)
        if batched:
            ctx.set_batches(batch)
        ctx.set_design(G, N, truth["grp"], ctl, rows=(g0, g1))
        # the control slab: from a host copy on every rank, or (MP_FIT_YCTL=nccl) from the owning ranks over NCCL
        ctx.upload_counts(Yl, Yctl=None if os.environ.get("MP_FIT_YCTL") == "nccl" else Yctl)
        logl, recs, stats = ctx.fit_nb(opts, W0=W0, psi_inject=sched[:, g0:g1])
        got = {n: ctx.get_state(n) for n in ("W", "alpha", "gmean", "Mb", "psi")}
        if batched:     # (the estimators refuse a batch context; get.res with batch psi is covered on one GPU)
            dec = [(r["outer"], r["iter"], r["degener"], r["accepted"]) for r in recs]
            refdec = [(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace]
            err = dict(W=relerr(got["W"], ref["W"]), alpha=relerr(got["alpha"], ref["a"][g0:g1]), gmean=relerr(got["gmean"], ref["gmean"][g0:g1]),
                       Mb=relerr(got["Mb"], ref["Mb"][g0:g1], floor=1e-6), psi=relerr(got["psi"], ref["psi"][g0:g1]), logl=relerr(logl, ref["logl"]))
            ok = dec == refdec and all(v < 1e-6 for v in err.values())
            t = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            print(f"rank {rank}/{world} rows [{g0},{g1}) batched ok={ok} err={ {k_: float(f'{v:.2e}') for k_, v in err.items()} }", flush=True)
            dist.destroy_process_group()
            sys.exit(0 if int(t.item()) == 1 else 1)
        # the dispersion step and get.res on the sharded context: both need quantities of ALL genes (the trend / prior of
        # estimateDisp, the block-wise Pearson clip limits), gathered over NCCL
        from oracle import edger, getres
        grp = truth["grp"]
        disp = ctx.estimate_disp_r(opts, robust=True)
        psi_dev = ctx.get_state("psi")
        ref_disp = edger.estimate_disp(Yf, ref["a"] @ ref["W"].T + ref["Mb"][:, grp])
        ctx.set_state("psi", ref["psi"][g0:g1])
        pear = ctx.get_res(1, block_size=512)
        ref_pear = getres.get_res(dict(counts=Y, a=ref["a"], W=ref["W"], gmean=ref["gmean"], Mb=ref["Mb"][:, grp], psi=ref["psi"]),
                                  getres.PEARSON, block_size=512)
    dec = [(r["outer"], r["iter"], r["degener"], r["accepted"]) for r in recs]
    refdec = [(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace]
    err = dict(W=relerr(got["W"], ref["W"]), alpha=relerr(got["alpha"], ref["a"][g0:g1]), gmean=relerr(got["gmean"], ref["gmean"][g0:g1]),
               Mb=relerr(got["Mb"], ref["Mb"][g0:g1], floor=1e-6), logl=relerr(logl, ref["logl"]),
               disp_common=abs(disp["common"] / ref_disp["common"] - 1), disp_psi=relerr(psi_dev, ref_disp["tagwise"][g0:g1]),
               pearson=relerr(pear, ref_pear[g0:g1]) / 10)   # the two fits differ at 1e-7; Pearson is allowed 1e-5
    ok = dec == refdec and all(v < 1e-6 for v in err.values())
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print(f"rank {rank}/{world} rows [{g0},{g1}) ok={ok} err={ {k_: float(f'{v:.2e}') for k_, v in err.items()} } inner={stats['n_inner']}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
