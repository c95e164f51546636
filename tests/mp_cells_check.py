"""Run under torchrun (one rank per GPU): cell-sharded fastruvIII.nb and get.res against the same calls in ONE process.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/mp_cells_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ruviiinb_b200 import api, synth  # noqa: E402
from tests.parity import relerr  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G, N, k, m, bs = 100, 1800, 2, 3, 256
    Y, Mfull, ctl, truth = synth.make_counts(G, N, k, m, 28, seed=189, gmean_mu=0.9, gmean_sd=0.8)
    rng = np.random.default_rng(3)
    Y[7, rng.choice(N, 40, replace=False)] = 60000          # Winsorize beyond the direct histogram bins
    Y[12] = 0; Y[12, [5, 1700]] = 3                         # emptied by the cap: dropped on every rank
    annotated = rng.random(N) < 0.4
    M = Mfull.astype(bool) & annotated[:, None]
    subsamples = np.sort(rng.choice(np.where(annotated)[0], 300, replace=False))
    psi_idx = np.arange(0, 300, 3)
    # single process (this rank's GPU), the whole problem
    one = api.fastruvIII_nb(Y, M, ctl, k=k, outer_maxit=3, subsamples=subsamples, psi_idx=psi_idx, block_size=bs, device=local, assay="PAC")
    # sharded: this rank's cells only
    ranks = api.Ranks(dist, device=dev)
    c0, c1 = api.cell_shard_range(N, world, rank, bs)
    sh = api.fastruvIII_nb(Y[:, c0:c1], M[c0:c1], ctl, k=k, outer_maxit=3, subsamples=subsamples, psi_idx=psi_idx, block_size=bs, device=local,
                           ranks=ranks, N_total=N, c0=c0, assay="PAC")
    err = dict(W=relerr(sh["W"], one["W"][c0:c1]), a=relerr(sh["a"], one["a"]), psi=relerr(sh["psi"], one["psi"]),
               Mb=relerr(sh["Mb"], one["Mb"][:, c0:c1], floor=1e-6), counts=float((sh["counts"] != one["counts"][:, c0:c1]).sum()),
               pac=float((sh["PAC"] != one["PAC"][:, c0:c1]).mean()), zero=float((sh["zero_genes"] != one["zero_genes"]).sum()))
    ok = err["W"] < 1e-10 and err["a"] < 1e-12 and err["Mb"] < 1e-9 and err["counts"] == 0 and err["pac"] <= 1e-5 and err["zero"] == 0 \
        and sh["W_sweeps"] == one["W_sweeps"] and one["zero_genes"][12]
    # get.res of a ruvIII.nb-style fit (Mb by replicate group), Pearson and PAC, sharded by cell
    fit = dict(counts=one["counts"], W=one["W"], a=one["a"], gmean=one["gmean"], psi=one["psi"], M=Mfull[:, :m], ctl=one["ctl"],
               Mb=0.3 * np.random.default_rng(5).standard_normal((one["a"].shape[0], m)))
    for kind in ("pearson", "quantile"):
        full = api.get_res(fit, type=kind, block_size=bs, device=local)
        part = api.get_res(dict(fit, counts=fit["counts"][:, c0:c1], W=fit["W"][c0:c1], M=fit["M"][c0:c1]), type=kind, block_size=bs, device=local,
                           ranks=ranks, N_total=N, c0=c0)
        if kind == "pearson":
            d = np.abs(part - full[:, c0:c1]).max() / np.abs(full).max()
            ok = ok and d <= 1e-12
        else:   # colMeans(W) is summed in another order over the ranks: only quantile-boundary ties may move
            d = (part != full[:, c0:c1]).mean()
            ok = ok and d <= 1e-5
        err["res_" + kind] = float(d)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print(f"rank {rank}/{world} cells [{c0},{c1}) ok={ok} err={ {k_: float(f'{v:.2e}') for k_, v in err.items()} }", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
