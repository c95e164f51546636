"""GPU parity of get.res (R1-R4, R/get.res.R:12-129) against oracle/getres.py: Pearson (block-wise clip), logcounts,
percentile-adjusted counts (exact integers)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fitted(G=96, N=1100, k=2, m=3, seed=51):
    from ruviiinb_b200 import synth
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, 24, seed=seed, gmean_mu=0.3)
    rng = np.random.default_rng(seed)
    Y[5] = rng.poisson(1500.0, N)           # counts far beyond the pmf-recurrence's short range
    W = truth["W"] / np.sqrt((truth["W"] ** 2).sum(axis=0))
    a = truth["alpha"] * np.sqrt((truth["W"] ** 2).sum(axis=0))
    gmean = truth["gmean"].copy()
    gmean[5] = np.log(1500.0)
    return Y, ctl, truth, W, a, gmean


def _run(kind, block_size, per_cell_mb):
    from oracle import getres
    from ruviiinb_b200 import _lib
    from tests.parity import make_ctx
    Y, ctl, truth, W, a, gmean = _fitted()
    grp = truth["grp"]
    Mb = truth["Mb"]
    Mb_cells = Mb[:, grp]
    if per_cell_mb:  # fastruvIII.nb's Mb.all: a genuinely per-cell matrix
        Mb_cells = Mb_cells + 0.1 * np.random.default_rng(9).standard_normal(Mb_cells.shape)
    ref = getres.get_res(dict(counts=Y, a=a, W=W, gmean=gmean, Mb=Mb_cells, psi=truth["psi"]), kind, block_size=block_size)
    with make_ctx(Y, grp, ctl) as ctx:
        ctx.set_state("W", W); ctx.set_state("alpha", a); ctx.set_state("gmean", gmean)
        ctx.set_state("Mb", Mb); ctx.set_state("psi", truth["psi"])
        got = ctx.get_res(kind, block_size=block_size, Mb_cells=Mb_cells if per_cell_mb else None)
    return got, ref


@pytest.mark.parametrize("per_cell_mb", [False, True])
def test_pearson(per_cell_mb):
    got, ref = _run(1, 256, per_cell_mb)
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()


def test_logcounts():
    got, ref = _run(2, 5000, False)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("per_cell_mb", [False, True])
def test_pac_exact(per_cell_mb):
    got, ref = _run(3, 300, per_cell_mb)
    mism = (got != ref)
    assert mism.mean() <= 1e-5, (mism.sum(), got[mism][:5], ref[mism][:5])   # quantile-boundary ties only


@pytest.mark.parametrize("all_cells", [True, False])
def test_zinb_fit_residuals(all_cells):
    """get.res of a zero-inflated fit (R/get.res.R:52-57,69-70,84-85,95-96): Pearson, logcounts and PAC with ZIM's dzinb / pzinb / qzinb;
    the reference reads out$pi0[, batch[c]], so with one batch every cell takes the zero-inflation flag of cell 1."""
    from oracle import getres
    from ruviiinb_b200 import api
    Y, ctl, truth, W, a, gmean = _fitted(G=60, N=700, seed=53)
    N = Y.shape[1]
    rng = np.random.default_rng(7)
    pi0 = rng.beta(2.0, 5.0, Y.shape[0])
    Y = np.where(rng.random(Y.shape) < pi0[:, None], 0, Y)
    zi = np.ones(N, bool) if all_cells else (np.arange(N) % 3 != 1)     # cell 1 (the column every cell reads) is flagged in both
    M = np.zeros((N, 3), bool); M[np.arange(N), truth["grp"]] = True
    out = dict(counts=Y, W=W, a=a, gmean=gmean, Mb=truth["Mb"], psi=truth["psi"], M=M, ctl=ctl, pi0=np.outer(pi0, zi.astype(float)))
    ref_out = dict(out, Mb=truth["Mb"][:, truth["grp"]])
    for kind, code in (("pearson", 1), ("logcounts", 2), ("quantile", 3)):
        got = api.get_res(out, type=kind, block_size=256)
        ref = getres.get_res(ref_out, code, block_size=256)
        if code == 3:
            assert (got != ref).mean() <= 1e-4, kind
        else:
            assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max(), kind
    nb = api.get_res(dict(out, pi0=None), type="quantile", block_size=256)
    assert (nb != got).mean() > 0.01        # the zero-inflated branch really differs from the NB one
