"""The host mirror of the reference's exported functions (ruviiinb_b200/api.py) used the way the reference is used:
ruvIII.nb end to end -- Winsorize, initial W, dispersion estimation, alternating IRLS -- against the oracle, then get.res."""
import numpy as np
import pytest


def _data(G=90, N=700, k=2, m=3, n_ctl=24, seed=91):
    from ruviiinb_b200 import synth
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=0.8, gmean_sd=0.9)
    Y[7] = 0                      # an all-zero gene is dropped (R/ruvIIInb.R:57-62)
    Y[11, 5] = 4000               # an outlier count is winsorised (:53)
    Y[13] = 0; Y[13, [3, N // 2]] = [2, 5]   # expressed in < 0.5 % of the cells: the cap (99.5 % quantile = 0) empties the row, which is
    return Y, M, ctl, truth              # dropped AFTER the cap like an all-zero gene (:53-62)


@pytest.mark.gpu
def test_ruvIII_nb_end_to_end_matches_oracle():
    from oracle import edger, ruviiinb as orc
    from ruviiinb_b200 import api
    from tests.parity import relerr
    Y, M, ctl, truth = _data()
    # the SAME initial W on both sides (irlba's start is random in the reference)
    Yw, _ = __import__("oracle.rstats", fromlist=["x"]).winsorize_counts(Y.astype(np.float64))
    keep = (Yw > 0).sum(axis=1) > 0
    W0 = api.init_W(Yw[keep], ctl[keep], 2)

    def psi_fn(Yx, offs, wt, psi_old):   # estimateDisp.par(Y, 1, offset = offs, weights = wt, tagwise, robust) restated
        return edger.estimate_disp(Yx, offs)["tagwise"]

    trace = []
    ref = orc.ruvIII_nb(Y.astype(np.float64), M, ctl, k=2, outer_maxit=3, W0=W0, psi_fn=psi_fn, trace=trace)
    out = api.ruvIII_nb(Y, M, ctl, k=2, outer_maxit=3, W0=W0, return_trace=True)
    assert out["zero_genes"][7] and out["zero_genes"][13] and out["counts"].shape[0] == Y.shape[0] - 2 == ref["counts"].shape[0]
    assert np.array_equal(out["counts"], ref["counts"].astype(np.int32))
    dec = [(r["outer"], r["iter"], r["degener"], r["accepted"]) for r in out["trace"]]
    assert dec == [(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace]
    err = dict(W=relerr(out["W"], ref["W"]), a=relerr(out["a"], ref["a"]), gmean=relerr(out["gmean"], ref["gmean"]),
               Mb=relerr(out["Mb"], ref["Mb"], floor=1e-6), psi=relerr(out["psi"], ref["psi"]), logl=relerr(out["logl"], ref["logl"]))
    assert all(v < 1e-6 for v in err.values()), err     # north_star tolerance
    # get.res on the fit, the way makeSCE calls it (default type: the percentile-adjusted counts win)
    from oracle import getres
    grp = np.argmax(M, axis=1)
    ref_out = dict(counts=ref["counts"], a=ref["a"], W=ref["W"], gmean=ref["gmean"], Mb=ref["Mb"][:, grp], psi=ref["psi"])
    pac = api.get_res(out, block_size=256)
    ref_pac = getres.get_res(ref_out, getres.QUANTILE, block_size=256)
    assert (pac != ref_pac).mean() < 1e-3      # integer counts; the two fits differ at 1e-7, so a few boundary ties may flip
    pear = api.get_res(out, type="pearson", block_size=256)
    assert np.abs(pear - getres.get_res(ref_out, getres.PEARSON, block_size=256)).max() < 1e-5


def test_api_argument_errors_mirror_stop():
    from ruviiinb_b200 import api
    Y, M, ctl, truth = _data(G=20, N=50, n_ctl=5)
    bad = M.copy(); bad[:, 0] = 0
    with pytest.raises(ValueError, match="exactly one replicate group|non-zero entry"):
        api.ruvIII_nb(Y, bad, ctl)
    with pytest.raises(NotImplementedError):
        api.ruvIII_nb(Y, M, ctl, ortho_W=True)
    with pytest.raises(ValueError, match="one entry per cell"):
        api.ruvIII_nb(Y, M, ctl, batch=np.ones(3), batch_disp=True)
    with pytest.raises(ValueError, match="unknown type"):
        api.get_res(dict(pi0=None, psi=np.ones(3)), type="foo")
    with pytest.raises(ValueError, match="'cData' must be a user-specified data frame"):       # R/makeSCE.R:15-16
        api.makeSCE(dict(batch=None))
    with pytest.raises(ValueError, match="Wrong names of corrected assays"):                   # R/makeSCE.R:17-18
        api.makeSCE(dict(batch=None), cData={}, assays=("logPAC", "tpm"))
    # estimateDisp.par: the reference's own stop() conditions, then what the device path does not implement
    with pytest.raises(ValueError, match="Incorrect length of group"):
        api.estimateDisp_par(Y, group=np.ones(3))
    with pytest.raises(ValueError, match="Incorrect length of lib.size"):
        api.estimateDisp_par(Y, lib_size=np.ones(3))
    with pytest.raises(ValueError, match="offsets must be finite"):
        api.estimateDisp_par(Y, offset=np.full(Y.shape, np.inf))
    with pytest.raises(ValueError, match="device path"):
        api.estimateDisp_par(Y, trend_method="loess")
    with pytest.raises(ValueError, match="weight matrix"):
        api.estimateDisp_par(Y, weights=np.ones(Y.shape))
    with pytest.raises(ValueError, match="not of the form"):
        api.estimateDisp_par(Y, offset=np.random.default_rng(0).standard_normal(Y.shape))


def test_offset_factorisation_is_exact():
    """api._factor_offset: alpha W' + Mb[, grp] is recovered (as some exact factorisation) with and without the groups."""
    from ruviiinb_b200 import api
    rng = np.random.default_rng(4)
    G, N, k, m = 40, 90, 3, 4
    grp = rng.integers(0, m, N).astype(np.int32); grp[:m] = np.arange(m)
    O = rng.standard_normal((G, k)) @ rng.standard_normal((N, k)).T + rng.standard_normal((G, m))[:, grp]
    for g, mm in ((grp, m), (np.zeros(N, dtype=np.int32), 1)):
        Mb, alpha, W = api._factor_offset(O, g, mm)
        assert alpha.shape[1] <= 8 and np.abs(Mb[:, g] + alpha @ W.T - O).max() < 1e-9 * np.abs(O).max()
    Mb, alpha, W = api._factor_offset(np.broadcast_to(rng.standard_normal(G)[:, None], (G, N)), np.zeros(N, dtype=np.int32), 1)
    assert alpha.shape == (G, 1) and not alpha.any() and not W.any()      # a per-gene constant: nothing left after Mb


@pytest.mark.gpu
@pytest.mark.parametrize("with_groups", [True, False])
def test_estimateDisp_par_matches_oracle(with_groups):
    """estimateDisp.par called on its own, as the reference exports it: y and an offset matrix in, the list out; robust = FALSE
    (its default) and robust = TRUE (how ruvIII.nb calls it)."""
    from oracle import edger
    from ruviiinb_b200 import api, synth
    Y, M, ctl, truth = synth.make_counts(160, 650, 2, 3, 40, seed=35, gmean_mu=0.5)
    grp = truth["grp"]
    Wn = truth["W"] / np.sqrt((truth["W"] ** 2).sum(axis=0))
    offs = (truth["alpha"] * np.sqrt((truth["W"] ** 2).sum(axis=0))) @ Wn.T + truth["Mb"][:, grp]
    for robust in (False, True):
        ref = edger.estimate_disp(Y.astype(float), offs, robust=robust)
        out = api.estimateDisp_par(Y, offset=offs, replicate_of_cell=grp if with_groups else None, robust=robust)
        assert abs(out["common.dispersion"] / ref["common"] - 1) < 1e-8
        assert np.abs(out["trended.dispersion"] / ref["trended"] - 1).max() < 1e-7
        fin = np.isfinite(ref["prior_df"])
        assert np.abs(np.broadcast_to(out["prior.df"], fin.shape)[fin] / ref["prior_df"][fin] - 1).max() < 1e-6
        assert np.abs(out["tagwise.dispersion"] / ref["tagwise"] - 1).max() < 1e-6
        assert np.all(np.asarray(out["prior.n"]) == np.asarray(out["prior.df"]) / (Y.shape[1] - 1))


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 3])
def test_init_W_device_matches_dense_svd(k):
    """H1 (R/ruvIIInb.R:160-168): subspace iteration on the resident control rows against a dense LAPACK SVD."""
    from ruviiinb_b200 import _lib, api, synth
    Y, M, ctl, truth = synth.make_counts(300, 1700, 3, 4, 120, seed=97, gmean_mu=1.0)
    ref = api.init_W(Y, ctl, k)
    A = np.log(Y[ctl] / Y[ctl].mean(axis=1, keepdims=True) + 1.0)
    sv_ref = np.linalg.svd(A, compute_uv=False)[:k]
    with _lib.Context(0) as ctx:
        ctx.set_design(Y.shape[0], Y.shape[1], truth["grp"], ctl)
        ctx.upload_counts(Y)
        W, sv, iters = ctx.init_W(k, max_iter=400, tol=1e-13)
        assert np.array_equal(ctx.get_state("W"), W)
    assert np.allclose(sv, sv_ref, rtol=1e-9)
    assert np.abs(W - ref).max() < 1e-6 * np.abs(ref).max(), (np.abs(W - ref).max(), iters)
    assert np.allclose(W.T @ W, np.eye(k), atol=1e-10)
