"""GPU parity of the ZINB path (H4/H5, R/ruvIIInb.R:237-264,435-467,494-504) against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(G=72, N=640, k=2, m=3, n_ctl=20, seed=61, all_cells=True):
    from oracle import nelder_mead, ruviiinb as orc
    from ruviiinb_b200 import synth
    from tests.parity import oracle_init
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, zinb=True, gmean_mu=0.8, gmean_sd=0.8)
    grp = truth["grp"]
    zeroinf = np.ones(N, bool) if all_cells else (np.arange(N) % 3 != 0)
    W0, gmean, alpha, adev = oracle_init(Y, ctl, k)
    psi = truth["psi"].copy()
    Yf = Y.astype(np.float64)
    mu = np.exp(gmean[:, None] + alpha @ W0.T)
    pi0 = nelder_mead.pi0_optim(Yf[:, zeroinf], mu[:, zeroinf], psi)          # :237-251
    pi0 = np.where(np.isnan(pi0), 0.0, pi0)
    wt = orc.zinb_weights(Yf, mu, psi, pi0, zeroinf, np.ones((G, N)))         # :254-264
    state = dict(gmean=gmean, Mb=np.zeros((G, m)), alpha=alpha, alpha_dev=adev, W=W0, psi=psi, step=np.ones(G), wt_zinb=wt, pi0=pi0)
    return Y, Yf, ctl, grp, zeroinf, state, truth


@pytest.mark.parametrize("all_cells", [True, False])
def test_zinb_one_iteration(all_cells):
    from oracle import nelder_mead, ruviiinb as orc
    from ruviiinb_b200 import _lib
    from tests.parity import make_ctx, relerr
    Y, Yf, ctl, grp, zeroinf, state, truth = _setup(all_cells=all_cells)
    G, N = Y.shape
    m = state["Mb"].shape[1]
    rep_ind = [np.where(grp == j)[0] for j in range(m)]
    ll0 = orc.loglik_matrix(Yf, state["gmean"][:, None] + state["alpha"] @ state["W"].T, state["psi"], state["pi0"], zeroinf).sum(axis=1)
    new, ll1, aux = orc.inner_iteration(Yf, state, grp, rep_ind, ctl, 100.0, 0.01, 16.0, zeroinf=zeroinf, pi0_fn=nelder_mead.pi0_optim)
    opts = _lib.default_opts(k=2, zinb=1)
    with make_ctx(Y, grp, ctl, zeroinf=zeroinf) as ctx:
        ctx.set_state("W", state["W"])
        ctx.init_coef(opts)
        ctx.set_state("psi", state["psi"])
        assert ctx.zinb_pi0(opts) == 0
        assert relerr(ctx.get_state("pi0"), state["pi0"]) < 1e-9
        tot0 = ctx.loglik(opts)
        assert relerr(ctx.get_state("loglik"), ll0) < 1e-10
        tot1, bad = ctx.irls_iteration(opts)
        err = {n: relerr(ctx.get_state(n), new[n], floor=1e-6) for n in ("alpha", "W", "gmean", "Mb", "pi0")}
        err["loglik1"] = abs(tot1 - ll1.sum()) / abs(ll1.sum())
    assert all(v < 1e-8 for v in err.values()), err


def test_zinb_full_fit_trajectory():
    from oracle import nelder_mead, ruviiinb as orc
    from ruviiinb_b200 import _lib, synth
    from tests.parity import make_ctx, relerr
    G, N, k, m = 60, 520, 2, 3
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, 16, seed=63, zinb=True, gmean_mu=1.0, gmean_sd=0.7)
    assert ((Y > 0).sum(axis=1) > 0).all()
    Yf = Y.astype(np.float64)
    zeroinf = np.ones(N, bool)
    W0 = orc.init_W_svd(Yf, ctl, k)
    sched = np.stack([truth["psi"] * f for f in (1.1, 1.0, 0.95)])
    calls = {"i": 0}

    def psi_fn(Yx, offs, wt, psi_old):
        i = min(calls["i"], 2)
        calls["i"] += 1
        return sched[i]

    trace = []
    ref = orc.ruvIII_nb(Yf, M, ctl, k=k, outer_maxit=3, zeroinf=zeroinf, W0=W0, psi_fn=psi_fn, pi0_fn=nelder_mead.pi0_optim,
                        winsorize=False, trace=trace)
    opts = _lib.default_opts(k=k, outer_maxit=3, zinb=1)
    with make_ctx(Y, truth["grp"], ctl, zeroinf=zeroinf) as ctx:
        logl, recs, stats = ctx.fit_nb(opts, W0=W0, psi_inject=sched)
        got = {n: ctx.get_state(n) for n in ("W", "alpha", "gmean", "Mb", "pi0")}
    dec = [(r["outer"], r["iter"], r["degener"], r["accepted"]) for r in recs]
    assert dec == [(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace]
    err = dict(W=relerr(got["W"], ref["W"]), alpha=relerr(got["alpha"], ref["a"]), gmean=relerr(got["gmean"], ref["gmean"]),
               Mb=relerr(got["Mb"], ref["Mb"], floor=1e-6), pi0=relerr(got["pi0"], ref["pi0"]), logl=relerr(logl, ref["logl"]))
    assert all(v < 1e-6 for v in err.values()), err
