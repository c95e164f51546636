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


@pytest.mark.gpu
def test_fast_subsample_fit_matches_oracle():
    """F3 + F4 (R/fastruvIIInb.R:223-623): the subsample fit with the dispersion estimated on <= 100 cells, against the
    oracle with the same U0 / psi cells: same accept / halve decisions per iteration, 1e-6 on every output."""
    from oracle import edger, fastruv
    from ruviiinb_b200 import api, synth
    from tests.parity import relerr
    G, ns, k, m = 110, 520, 2, 3
    Y, M, ctl, truth = synth.make_counts(G, ns, k, m, 30, seed=87, gmean_mu=0.9, gmean_sd=0.8)
    assert ((Y > 0).sum(axis=1) > 0).all()
    grp = truth["grp"]
    Z = np.log(Y + 1.0)
    zbar = np.stack([Z[:, grp == j].mean(axis=1) for j in range(m)], axis=1)
    _, _, vt = np.linalg.svd(Z - zbar[:, grp], full_matrices=False)
    U0 = vt[:k].T.copy()
    psi_idx = np.arange(3, ns, 6)[:80]

    def psi_fn(Yp, offs_p):
        return edger.estimate_disp(Yp, offs_p)["tagwise"]

    rt, gt = [], []
    ref = fastruv.fast_fit_sub(Y, grp, ctl, k, U0, psi_fn, psi_idx, outer_maxit=3, trace=rt)
    got = api.fastruvIII_nb_subfit(Y, M, ctl, k=k, outer_maxit=3, U0=U0, psi_idx=psi_idx, trace=gt)
    assert [(r["outer"], r["iter"], bool(r["degener"]), bool(r["accepted"])) for r in gt] == \
           [(r["outer"], r["iter"], bool(r["degener"]), bool(r["accepted"])) for r in rt]
    err = {n: relerr(got[n], ref[n], floor=1e-6) for n in ("alpha", "gmean", "Mb", "Wsub", "psi", "wt_ctl", "logl")}
    assert all(v < 1e-6 for v in err.values()), err


@pytest.mark.gpu
def test_fastruvIII_nb_end_to_end():
    """The whole fastruvIII.nb (Winsorize, subsample fit, W for all cells, Mb.all, psi cap) against the oracle pieces
    chained the same way; only 40 % of the cells are annotated."""
    from oracle import edger, fastruv, rstats
    from ruviiinb_b200 import api, synth
    from tests.parity import relerr
    G, N, k, m = 100, 900, 2, 3
    Y, Mfull, ctl, truth = synth.make_counts(G, N, k, m, 28, seed=89, gmean_mu=0.9, gmean_sd=0.8)
    rng = np.random.default_rng(3)
    annotated = rng.random(N) < 0.4
    M = Mfull.astype(bool) & annotated[:, None]
    subsamples = np.sort(rng.choice(np.where(annotated)[0], 240, replace=False))
    psi_idx = np.arange(0, 240, 3)
    Yw, _ = rstats.winsorize_counts(Y.astype(np.float64))
    assert ((Yw > 0).sum(axis=1) > 0).all()
    grp_sub = truth["grp"][subsamples]
    Z = np.log(Yw[:, subsamples] + 1.0)
    zbar = np.stack([Z[:, grp_sub == j].mean(axis=1) for j in range(m)], axis=1)
    _, _, vt = np.linalg.svd(Z - zbar[:, grp_sub], full_matrices=False)
    U0 = vt[:k].T.copy()
    sub = fastruv.fast_fit_sub(Yw[:, subsamples], grp_sub, ctl, k, U0, lambda y, o: edger.estimate_disp(y, o)["tagwise"], psi_idx, outer_maxit=3)
    W_init = np.zeros((N, k)); W_init[subsamples] = sub["Wsub"]
    refW, _, _ = fastruv.fast_W_all(Yw, ctl, sub["gmean"], sub["alpha"], sub["wt_ctl"], W_init, sub["Wsub"])
    annot = np.where(annotated, truth["grp"], -1)
    refMb = fastruv.fast_Mb_all(Yw, ctl, sub["gmean"], sub["alpha"], refW, sub["psi"], sub["Mb"], annot)
    out = api.fastruvIII_nb(Y, M, ctl, k=k, outer_maxit=3, subsamples=subsamples, U0=U0, psi_idx=psi_idx, block_size=256)
    assert relerr(out["W"], refW) < 1e-6 and relerr(out["a"], sub["alpha"]) < 1e-6 and relerr(out["gmean"], sub["gmean"]) < 1e-6
    assert relerr(out["Mb"], refMb, floor=1e-6) < 1e-5
    assert relerr(out["psi"], fastruv.cap_psi(sub["psi"])) < 1e-6
    pac = api.get_res(out, block_size=300)      # makeSCE(assays = 'logPAC') path: per-cell Mb
    assert pac.shape == (G, N) and np.isfinite(pac).all()
