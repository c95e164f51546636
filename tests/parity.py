"""Shared parity harness: run the CUDA path (through the C ABI) and the NumPy oracle on the same
seeded inputs and report scaled errors.  Test infrastructure -- the only place besides bench.py's
cpu_baseline leg that imports `oracle`."""
import numpy as np

from oracle import ruviiinb as orc
from ruviiinb_b200 import _lib, synth


def relerr(a, b, floor=1e-300):
    """max |a-b| / max(floor, max|b|): scaled absolute error, NaN-aware (NaN pattern must match).
    `floor` guards quantities that are mathematically zero (Mb with a single replicate group)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    if (na != nb).any():
        return np.inf
    if nb.all():
        return 0.0
    d = np.abs(np.where(nb, 0.0, a - b)).max()
    return float(d / max(floor, np.abs(b[~nb]).max()))


def oracle_init(Y, ctl, k, W0=None):
    """R/ruvIIInb.R:160-202 on the oracle side: W0 (SVD stand-in), gmean, alpha, alpha.dev."""
    Yf = Y.astype(np.float64)
    W = orc.init_W_svd(Yf, ctl, k) if W0 is None else W0
    G = Y.shape[0]
    coef = np.empty((G, k + 1))
    Z0 = np.log(Yf + 1.0)
    for g in range(G):
        coef[g] = orc.get_coef_wt(Z0[g], Yf[g] + 1.0, W)
    alpha = coef[:, 1:].copy()
    adev = alpha - alpha.mean(axis=0)[None, :]
    alpha[~np.isfinite(alpha)] = 0.0
    return W, coef[:, 0].copy(), alpha, adev


def make_ctx(Y, grp, ctl, device=0, rows=None, layout="gene_major", zeroinf=None):
    G, N = Y.shape
    ctx = _lib.Context(device)
    ctx.set_design(G, N, grp, ctl, zeroinf=zeroinf, rows=rows)
    if rows is None:
        ctx.upload_counts(Y, layout=layout)
    else:
        ctx.upload_counts(Y[rows[0]:rows[1]], Yctl=Y[np.asarray(ctl, bool)], layout=layout)
    return ctx


def one_iteration_case(G=256, N=1500, k=2, m=3, n_ctl=64, seed=7, robust=False, fastpath=1, layout="gene_major",
                       gmean_mu=-0.5, step=None):
    """init (H2) + log-lik (H6) + one inner iteration (H7-H13) on both sides."""
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=gmean_mu)
    grp = truth["grp"]
    rep_ind = [np.where(grp == j)[0] for j in range(m)]
    Yf = Y.astype(np.float64)
    W0, gmean, alpha, adev = oracle_init(Y, ctl, k)
    psi = truth["psi"].copy()
    stepv = np.ones(G) if step is None else step
    state = dict(gmean=gmean, Mb=np.zeros((G, m)), alpha=alpha, alpha_dev=adev, W=W0, psi=psi, step=stepv,
                 wt_zinb=np.ones((G, N)), pi0=np.zeros(G))
    ll0 = orc.loglik_matrix(Yf, gmean[:, None] + alpha @ W0.T, psi).sum(axis=1)
    k_huber = 1.345 if robust else 100.0
    new, ll1, aux = orc.inner_iteration(Yf, state, grp, rep_ind, ctl, k_huber, 0.01, 16.0)

    err = {}
    opts = _lib.default_opts(k=k, robust=int(robust), huber_fastpath=fastpath)
    with make_ctx(Y, grp, ctl, layout=layout) as ctx:
        ctx.set_state("W", W0)
        ctx.init_coef(opts)
        err["init_gmean"] = relerr(ctx.get_state("gmean"), gmean)
        err["init_alpha"] = relerr(ctx.get_state("alpha"), alpha)
        err["init_adev"] = relerr(ctx.get_state("alpha_dev"), adev)
        ctx.set_state("psi", psi)
        ctx.set_state("step", stepv)
        tot0 = ctx.loglik(opts)
        err["loglik0"] = relerr(ctx.get_state("loglik"), ll0)
        err["loglik0_sum"] = abs(tot0 - ll0.sum()) / abs(ll0.sum())
        tot1, bad = ctx.irls_iteration(opts)
        err["alpha"] = relerr(ctx.get_state("alpha"), new["alpha"])
        err["alpha_dev"] = relerr(ctx.get_state("alpha_dev"), new["alpha_dev"])
        err["W"] = relerr(ctx.get_state("W"), new["W"])
        err["gmean"] = relerr(ctx.get_state("gmean"), new["gmean"])
        err["Mb"] = relerr(ctx.get_state("Mb"), new["Mb"], floor=1e-6)
        err["loglik1_sum"] = abs(tot1 - ll1.sum()) / abs(ll1.sum())
        launches = ctx.kernel_launches
        sigma = ctx.get_state("sigma")
    return dict(err=err, bad=bad, oracle_bad=[aux["n_bad_alpha"], aux["n_bad_W"], aux["n_bad_Mb"]], launches=launches,
                sigma=sigma, oracle_sigma=aux["sigma"])


def fit_case(G=200, N=1200, k=2, m=3, n_ctl=50, seed=11, robust=False, fastpath=1, inner_maxit=50, outer_maxit=3,
             n_psi=3, rows=None):
    """Whole NB fit (R/ruvIIInb.R:266-609) with injected W0 and an injected psi schedule on both sides."""
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed)
    grp = truth["grp"]
    Yf = Y.astype(np.float64)
    W0 = orc.init_W_svd(Yf, ctl, k)
    rng = np.random.default_rng(seed + 1)
    sched = [truth["psi"] * np.exp(0.2 * rng.standard_normal(G) / (i + 1)) for i in range(n_psi)]
    calls = {"i": 0}

    def psi_fn(Yx, offs, wt, psi_old):
        i = min(calls["i"], n_psi - 1)
        calls["i"] += 1
        return sched[i]

    trace = []
    ref = orc.ruvIII_nb(Yf, M, ctl, k=k, robust=robust, inner_maxit=inner_maxit, outer_maxit=outer_maxit, W0=W0,
                        psi_fn=psi_fn, winsorize=False, trace=trace)
    opts = _lib.default_opts(k=k, robust=int(robust), huber_fastpath=fastpath, inner_maxit=inner_maxit,
                             outer_maxit=outer_maxit)
    with make_ctx(Y, grp, ctl, rows=rows) as ctx:
        sl = slice(None) if rows is None else slice(rows[0], rows[1])
        logl, recs, stats = ctx.fit_nb(opts, W0=W0, psi_inject=np.stack(sched)[:, sl])
        got = {n: ctx.get_state(n) for n in ("W", "alpha", "gmean", "Mb", "psi")}
    err = dict(W=relerr(got["W"], ref["W"]), alpha=relerr(got["alpha"], ref["a"][sl]),
               gmean=relerr(got["gmean"], ref["gmean"][sl]), Mb=relerr(got["Mb"], ref["Mb"][sl], floor=1e-6),
               psi=relerr(got["psi"], ref["psi"][sl]))
    if rows is None:
        n = min(len(logl), len(ref["logl"]))
        err["logl_outer"] = relerr(logl[:n], ref["logl"][:n]) if n else np.inf
    return dict(err=err, logl=logl, ref_logl=ref["logl"], recs=recs, ref_trace=trace, stats=stats,
                n_inner_ref=ref["n_inner"])
