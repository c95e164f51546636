"""Regenerates the golden fixtures in this directory:  python tests/golden/make_golden.py

PARITY UNPINNED: the reference (limfuxing/ruvIIInb, pure R) ships no tests or golden vectors and cannot run
here (no R), so these vectors are outputs of the NumPy restatement in oracle/ (which cites the reference
file:line for every step), frozen so that (i) the oracle cannot drift silently and (ii) the GPU tests have a
fixed target that does not depend on the oracle being importable at the same version.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ruviiinb as orc  # noqa: E402
from ruviiinb_b200 import synth  # noqa: E402
from tests.parity import oracle_init  # noqa: E402


def one_iteration(G, N, k, m, n_ctl, seed, robust):
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed)
    grp = truth["grp"]
    W0, gmean, alpha, adev = oracle_init(Y, ctl, k)
    psi = truth["psi"]
    state = dict(gmean=gmean, Mb=np.zeros((G, m)), alpha=alpha, alpha_dev=adev, W=W0, psi=psi, step=np.ones(G),
                 wt_zinb=np.ones((G, N)), pi0=np.zeros(G))
    Yf = Y.astype(np.float64)
    ll0 = orc.loglik_matrix(Yf, gmean[:, None] + alpha @ W0.T, psi).sum(axis=1)
    rep_ind = [np.where(grp == j)[0] for j in range(m)]
    new, ll1, aux = orc.inner_iteration(Yf, state, grp, rep_ind, ctl, 1.345 if robust else 100.0, 0.01, 16.0)
    return dict(Y=Y.astype(np.int32), grp=grp.astype(np.int32), ctl=ctl, robust=np.array(robust), W0=W0, psi=psi,
                init_gmean=gmean, init_alpha=alpha, init_adev=adev, loglik0=ll0, alpha=new["alpha"],
                alpha_dev=new["alpha_dev"], W=new["W"], gmean=new["gmean"], Mb=new["Mb"], loglik1=ll1, sigma=aux["sigma"])


def full_fit(G, N, k, m, n_ctl, seed):
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=0.5, gmean_sd=1.0)
    assert ((Y > 0).sum(axis=1) > 0).all()  # no all-zero gene: the psi schedule is indexed by input gene
    Yf = Y.astype(np.float64)
    W0 = orc.init_W_svd(Yf, ctl, k)
    rng = np.random.default_rng(seed + 1)
    sched = np.stack([truth["psi"] * np.exp(0.2 * rng.standard_normal(G) / (i + 1)) for i in range(3)])
    calls = {"i": 0}

    def psi_fn(Yx, offs, wt, psi_old):
        i = min(calls["i"], 2)
        calls["i"] += 1
        return sched[i]

    trace = []
    ref = orc.ruvIII_nb(Yf, M, ctl, k=k, outer_maxit=3, W0=W0, psi_fn=psi_fn, winsorize=False, trace=trace)
    dec = np.array([(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace], dtype=np.int32)
    return dict(Y=Y.astype(np.int32), grp=truth["grp"].astype(np.int32), ctl=ctl, W0=W0, psi_sched=sched, decisions=dec,
                logl_new=np.array([r["logl_new"] for r in trace]), logl_outer=ref["logl"], W=ref["W"], alpha=ref["a"],
                gmean=ref["gmean"], Mb=ref["Mb"], psi=ref["psi"])


def dispersion(G, N, k, m, n_ctl, seed):
    """estimateDisp.par on a fitted-looking state: the APL grid, common / trended / tagwise dispersion, prior d.f."""
    from oracle import edger
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=0.5)
    rng = np.random.default_rng(seed)
    W = truth["W"] / np.sqrt((truth["W"] ** 2).sum(axis=0))
    alpha = truth["alpha"] * np.sqrt((truth["W"] ** 2).sum(axis=0)) + 0.05 * rng.standard_normal((G, k))
    Mb = truth["Mb"] + 0.05 * rng.standard_normal((G, m))
    Y[3] = 0                      # an all-zero gene (rowSums < 5: not selected, takes the trend's value)
    offs = alpha @ W.T + Mb[:, truth["grp"]]
    Yf = Y.astype(np.float64)
    sel = Yf.sum(axis=1) >= 5
    l0 = np.full((G, 21), np.nan)
    l0[sel] = edger.apl_grid(Yf[sel], offs[sel])
    ref = edger.estimate_disp(Yf, offs)
    return dict(Y=Y.astype(np.int32), grp=truth["grp"].astype(np.int32), ctl=ctl, W=W, alpha=alpha, Mb=Mb, l0=l0,
                common=np.array(ref["common"]), trended=ref["trended"], tagwise=ref["tagwise"], prior_df=np.asarray(ref["prior_df"]))


def residuals(G, N, k, m, n_ctl, seed, block_size):
    """get.res of a fitted-looking state: Pearson, logcounts, percentile-adjusted counts, in blocks of `block_size` cells."""
    from oracle import getres
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=0.3)
    W = truth["W"] / np.sqrt((truth["W"] ** 2).sum(axis=0))
    a = truth["alpha"] * np.sqrt((truth["W"] ** 2).sum(axis=0))
    out = dict(counts=Y, a=a, W=W, gmean=truth["gmean"], Mb=truth["Mb"][:, truth["grp"]], psi=truth["psi"])
    return dict(Y=Y.astype(np.int32), grp=truth["grp"].astype(np.int32), ctl=ctl, W=W, alpha=a, gmean=truth["gmean"], Mb=truth["Mb"],
                psi=truth["psi"], block_size=np.array(block_size), pearson=getres.get_res(out, 1, block_size),
                logcounts=getres.get_res(out, 2, block_size), pac=getres.get_res(out, 3, block_size))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "disp_k2_m3.npz"), **dispersion(90, 420, 2, 3, 20, 2029))
    np.savez_compressed(os.path.join(HERE, "res_k2_m3.npz"), **residuals(40, 300, 2, 3, 12, 2030, 128))
    np.savez_compressed(os.path.join(HERE, "iter_k2_m3.npz"), **one_iteration(48, 300, 2, 3, 12, 2026, False))
    np.savez_compressed(os.path.join(HERE, "iter_k3_m4_robust.npz"), **one_iteration(40, 260, 3, 4, 12, 2027, True))
    np.savez_compressed(os.path.join(HERE, "fit_k2_m3.npz"), **full_fit(40, 260, 2, 3, 12, 2028))
    print("golden fixtures written to", HERE)
