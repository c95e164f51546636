"""GPU parity of the dispersion step (H3, R/estimateDisp.par.R) against oracle/edger.py on the same inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(G=150, N=900, k=2, m=3, n_ctl=40, seed=31, gmean_mu=0.5):
    from ruviiinb_b200 import synth
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=gmean_mu)
    rng = np.random.default_rng(seed)
    W = truth["W"] / np.sqrt((truth["W"] ** 2).sum(axis=0))
    alpha = truth["alpha"] * np.sqrt((truth["W"] ** 2).sum(axis=0)) + 0.05 * rng.standard_normal((G, k))
    Mb = truth["Mb"] + 0.05 * rng.standard_normal((G, m))
    offs = alpha @ W.T + Mb[:, truth["grp"]]
    return Y, ctl, truth, W, alpha, Mb, offs


def _ctx(Y, ctl, truth, W, alpha, Mb, k):
    from ruviiinb_b200 import _lib
    from tests.parity import make_ctx
    ctx = make_ctx(Y, truth["grp"], ctl)
    ctx.set_state("W", W)
    ctx.set_state("alpha", alpha)
    ctx.set_state("Mb", Mb)
    ctx.set_state("gmean", truth["gmean"])  # must NOT enter the offset (R/ruvIIInb.R:206)
    return ctx, _lib.default_opts(k=k)


@pytest.mark.parametrize("ecache", [True, False])
def test_apl_grid_matches_oracle(ecache, monkeypatch):
    """l0 of the 21-point grid against the oracle, on the exp(offset) cache and on the recompute path (what a shard too
    large for the cache runs), then once more on the same context: the second call starts from the first call's
    intercepts and must land on the same values."""
    from oracle import edger
    if not ecache:
        monkeypatch.setenv("RUVNB_NO_ECACHE", "1")
    Y, ctl, truth, W, alpha, Mb, offs = _case()
    Y[7] = 0            # all-zero gene: rowSums < 5 -> not selected
    Y[8, :] = 0
    Y[8, 3] = 2         # rowSums = 2 < 5
    Y[9] = np.random.default_rng(1).poisson(700.0, Y.shape[1])   # counts beyond the 128-entry tables
    ctx, opts = _ctx(Y, ctl, truth, W, alpha, Mb, 2)
    with ctx:
        l0, rs = ctx.disp_apl_grid(opts)
        l0b, _ = ctx.disp_apl_grid(opts)
    assert np.array_equal(rs, Y.sum(axis=1))
    ok = np.isfinite(l0)
    assert np.array_equal(ok, np.isfinite(l0b)) and np.abs(l0b[ok] / l0[ok] - 1).max() < 1e-9   # both within the NR tolerance of the same root
    sel = rs >= 5
    assert np.isnan(l0[~sel]).all() and not sel[7] and not sel[8]
    ref = edger.apl_grid(Y[sel].astype(float), offs[sel])
    err = np.abs(l0[sel] - ref) / np.abs(ref)
    assert err.max() < 1e-9, err.max()


@pytest.mark.parametrize("robust", [True, False])
@pytest.mark.parametrize("prior_df", [None, 0.0, 7.5])
def test_estimate_disp_matches_oracle(prior_df, robust):
    """The whole of estimateDisp.par: robust = TRUE (per-gene prior d.f. from limma's robust squeezeVar with the AveLogCPM trend)
    is what the reference always passes; robust = FALSE is fitFDist with the same trend; an injected prior d.f. skips both."""
    from oracle import edger
    Y, ctl, truth, W, alpha, Mb, offs = _case(G=180, N=700, seed=33)
    Y[11] *= 6; Y[12] = np.random.default_rng(1).poisson(3.0, Y.shape[1]) * np.random.default_rng(2).integers(0, 30, Y.shape[1])   # outlier genes
    ref = edger.estimate_disp(Y.astype(float), offs, prior_df=prior_df, robust=robust)
    ctx, opts = _ctx(Y, ctl, truth, W, alpha, Mb, 2)
    with ctx:
        out = ctx.estimate_disp_r(opts, robust=robust, prior_df=prior_df)
        psi = ctx.get_state("psi")
        if not robust:   # the older entry point is the non-robust form
            ctx.set_state("psi", np.ones(Y.shape[0]))
            ex = ctx.estimate_disp_ex(opts, prior_df=prior_df)
            # (a second call starts its grid fits from the first call's intercepts: equal to the Newton tolerance, not to the bit)
            assert np.allclose(ctx.get_state("psi"), psi, rtol=1e-8) and abs(ex["prior_df"] - out["prior_df"][ref["sel"]][0]) <= 1e-8 * max(1.0, ex["prior_df"])
    assert abs(out["common"] / ref["common"] - 1) < 1e-8
    assert np.abs(out["trended"] / ref["trended"] - 1).max() < 1e-7
    fin = np.isfinite(ref["prior_df"])
    assert np.array_equal(np.isfinite(out["prior_df"]), fin)
    assert np.all(np.abs(out["prior_df"][fin] - ref["prior_df"][fin]) <= 1e-6 * np.abs(ref["prior_df"][fin]))
    if robust and prior_df is None:
        assert len(np.unique(out["prior_df"][ref["sel"]])) > 1      # outlier genes carry a smaller prior d.f.
    assert np.abs(psi / ref["tagwise"] - 1).max() < 1e-6   # north_star tolerance on psi


def test_disp_newton_is_the_nb_mle_at_fixed_means():
    """The north_star variant (not the reference's estimator): per-gene Newton on log psi; checked against a SciPy
    root of the same score equation, and reported against the reference-style tagwise value."""
    from scipy import optimize, special
    from ruviiinb_b200 import _lib
    Y, ctl, truth, W, alpha, Mb, offs = _case(G=60, N=1200, seed=35)
    Y[4] = np.random.default_rng(2).poisson(400.0, Y.shape[1])
    gmean = truth["gmean"].copy(); gmean[4] = np.log(400.0)
    mu = np.exp(gmean[:, None] + offs)
    ctx, opts = _ctx(Y, ctl, truth, W, alpha, Mb, 2)
    with ctx:
        ctx.set_state("gmean", gmean)
        ctx.set_state("psi", np.full(Y.shape[0], 0.5))
        psi = ctx.disp_newton(opts, maxit=50)

    def score(th, y, m):
        r = np.exp(-th)
        return np.sum(special.digamma(y + r) - special.digamma(r) + np.log(r) + 1 - np.log(m + r) - (y + r) / (m + r))

    for g in (0, 4, 17, 33):
        th = optimize.brentq(score, -12, 8, args=(Y[g].astype(float), mu[g]), xtol=1e-13)
        assert abs(np.log(psi[g]) - th) < 1e-6, (g, psi[g], np.exp(th))
