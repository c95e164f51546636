"""GPU parity of fastruvIII.nb's full-data passes (F5 W for all cells, F6 Mb.all, F7 psi cap) against oracle/fastruv.py."""
import numpy as np
import pytest


def _problem(G=120, N=1500, k=3, m=4, n_ctl=40, seed=81):
    from ruviiinb_b200 import synth
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=0.4)
    rng = np.random.default_rng(seed)
    sc = np.sqrt((truth["W"] ** 2).sum(axis=0))
    Wtrue = truth["W"] / sc
    alpha = truth["alpha"] * sc
    sub = np.sort(rng.choice(N, 300, replace=False))             # the fitted subsample (R/fastruvIIInb.R:118-137), injected
    Wsub = Wtrue[sub] + 0.002 * rng.standard_normal((300, k))
    W_init = np.zeros((N, k))
    W_init[sub] = Wsub
    wt_ctl = rng.uniform(0.3, 1.5, n_ctl)
    annot = np.full(N, -1, dtype=np.int32)
    annot[sub] = truth["grp"][sub]
    return Y, ctl, truth, alpha, Wtrue, Wsub, W_init, wt_ctl, annot


@pytest.mark.gpu
def test_fast_W_all_matches_oracle():
    from oracle import fastruv, rstats
    from ruviiinb_b200 import _lib
    from tests.parity import make_ctx
    Y, ctl, truth, alpha, Wtrue, Wsub, W_init, wt_ctl, annot = _problem()
    refW, ref_sweeps, ref_crit = fastruv.fast_W_all(Y, ctl, truth["gmean"], alpha, wt_ctl, W_init, Wsub)
    with make_ctx(Y, truth["grp"], ctl) as ctx:
        ctx.set_state("W", W_init)
        ctx.set_state("alpha", alpha)
        ctx.set_state("gmean", truth["gmean"])
        W, sweeps, crit = ctx.fast_W_all(W_init, wt_ctl, rstats.r_median(Wsub, axis=0), rstats.r_mad(Wsub, axis=0))
    assert sweeps == ref_sweeps
    assert np.abs(W - refW).max() <= 1e-10 * np.abs(refW).max()
    assert abs(crit - ref_crit) <= 1e-6 * ref_crit


@pytest.mark.gpu
@pytest.mark.parametrize("all_annotated", [False, True])
def test_fast_Mb_all_matches_oracle(all_annotated):
    from oracle import fastruv
    from tests.parity import make_ctx
    Y, ctl, truth, alpha, Wtrue, Wsub, W_init, wt_ctl, annot = _problem(seed=83)
    if all_annotated:
        annot = truth["grp"].astype(np.int32)
    ref = fastruv.fast_Mb_all(Y, ctl, truth["gmean"], alpha, Wtrue, truth["psi"], truth["Mb"], annot)
    with make_ctx(Y, truth["grp"], ctl) as ctx:
        ctx.set_state("W", Wtrue); ctx.set_state("alpha", alpha); ctx.set_state("gmean", truth["gmean"])
        ctx.set_state("psi", truth["psi"]); ctx.set_state("Mb", truth["Mb"])
        got = ctx.fast_Mb_all(annot, block_size=400)
    assert np.abs(got - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max())


def test_cap_psi_host():
    from oracle import fastruv
    from ruviiinb_b200 import _lib
    psi = np.exp(np.random.default_rng(0).normal(-1, 0.8, 501))
    psi[3] = 500.0
    assert np.allclose(_lib.fast_cap_psi(psi), fastruv.cap_psi(psi), rtol=1e-14)
