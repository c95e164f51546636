"""Batch-specific dispersion psi[g, batch(c)] (`batch.disp = TRUE`, R/ruvIIInb.R:213-220,278-283,306-315,542-548; R/get.res.R:59-62,
101-121) and the pseudo-cells of `use.pseudosample = TRUE` (:83-144) -- SURVEY.md section 8(f)2.  The CUDA path (chunks that never
mix batches, one dispersion column and one set of count tables per batch) against the oracle on the same inputs."""
import numpy as np
import pytest


def _case(G=180, N=1300, k=2, m=4, n_ctl=48, seed=23, nb=2):
    from ruviiinb_b200 import synth
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=seed, gmean_mu=0.3)
    rng = np.random.default_rng(seed + 5)
    batch = rng.integers(1, nb + 1, size=N)                 # every replicate group meets every batch
    batch[:nb] = np.arange(1, nb + 1)
    psiB = np.stack([truth["psi"] * np.exp(0.5 * rng.standard_normal(G)) for _ in range(nb)], axis=1)   # G x nb, clearly different
    return Y, M, ctl, truth, batch, psiB


@pytest.mark.gpu
@pytest.mark.parametrize("nb,robust", [(2, False), (3, False), (2, True)])
def test_one_iteration_with_batch_psi_matches_oracle(nb, robust):
    from oracle import ruviiinb as orc
    from ruviiinb_b200 import _lib
    from tests.parity import oracle_init, relerr
    Y, M, ctl, truth, batch, psiB = _case(nb=nb)
    G, N = Y.shape
    k, m = 2, M.shape[1]
    grp = truth["grp"]
    rep_ind = [np.where(grp == j)[0] for j in range(m)]
    Yf = Y.astype(np.float64)
    W0, gmean, alpha, adev = oracle_init(Y, ctl, k)
    psic = psiB[:, batch - 1]                                 # psi[, batch]
    state = dict(gmean=gmean, Mb=np.zeros((G, m)), alpha=alpha, alpha_dev=adev, W=W0, psi=psic, step=np.ones(G), wt_zinb=np.ones((G, N)),
                 pi0=np.zeros(G))
    ll0 = orc.loglik_matrix(Yf, gmean[:, None] + alpha @ W0.T, psic).sum(axis=1)
    new, ll1, aux = orc.inner_iteration(Yf, state, grp, rep_ind, ctl, 1.345 if robust else 100.0, 0.01, 16.0)
    new2, ll2, _ = orc.inner_iteration(Yf, dict(new, step=np.ones(G), psi=psic), grp, rep_ind, ctl, 1.345 if robust else 100.0, 0.01, 16.0)
    opts = _lib.default_opts(k=k, robust=int(robust))
    with _lib.Context(0) as ctx:
        b0 = ctx.set_batches(batch)
        assert np.array_equal(b0, batch - 1)
        ctx.set_design(G, N, grp, ctl)
        ctx.upload_counts(Y)
        assert np.array_equal(ctx.get_counts(), Y)            # the batch-aware cell order is invisible at the boundary
        ctx.set_state("W", W0)
        ctx.init_coef(opts)
        ctx.set_state("psi", psiB)
        assert np.array_equal(ctx.get_state("psi"), psiB)
        tot0 = ctx.loglik(opts)
        assert relerr(ctx.get_state("loglik"), ll0) < 1e-10 and abs(tot0 - ll0.sum()) < 1e-10 * abs(ll0.sum())
        tot1, bad = ctx.irls_iteration(opts)
        err = dict(alpha=relerr(ctx.get_state("alpha"), new["alpha"]), W=relerr(ctx.get_state("W"), new["W"]),
                   gmean=relerr(ctx.get_state("gmean"), new["gmean"]), Mb=relerr(ctx.get_state("Mb"), new["Mb"], floor=1e-6),
                   ll=abs(tot1 - ll1.sum()) / abs(ll1.sum()))
        assert all(v < 1e-9 for v in err.values()), err
        ctx.accept()
        tot2, _ = ctx.irls_iteration(opts)                    # the second iteration starts from the fused sweep's sums
        err2 = dict(alpha=relerr(ctx.get_state("alpha"), new2["alpha"]), W=relerr(ctx.get_state("W"), new2["W"]),
                    gmean=relerr(ctx.get_state("gmean"), new2["gmean"]), ll=abs(tot2 - ll2.sum()) / abs(ll2.sum()))
        assert all(v < 1e-8 for v in err2.values()), err2
        # the estimators refuse a batch context (one estimateDisp per batch belongs to one context per batch)
        with pytest.raises(_lib.RuvnbError, match="batch"):
            ctx.estimate_disp(opts)
    # equal columns reproduce the single-dispersion fit exactly (same arithmetic, only the cell order differs)
    with _lib.Context(0) as c1, _lib.Context(0) as c2:
        c2.set_batches(batch)
        for c in (c1, c2):
            c.set_design(G, N, grp, ctl); c.upload_counts(Y); c.set_state("W", W0); c.init_coef(opts)
        c1.set_state("psi", truth["psi"]); c2.set_state("psi", np.repeat(truth["psi"][:, None], nb, axis=1))
        t1, _ = c1.irls_iteration(opts); t2, _ = c2.irls_iteration(opts)
        assert abs(t1 - t2) < 1e-11 * abs(t1) and relerr(c2.get_state("W"), c1.get_state("W")) < 1e-11


@pytest.mark.gpu
def test_fit_with_batch_psi_schedule_matches_oracle():
    from oracle import ruviiinb as orc, rstats
    from ruviiinb_b200 import api
    from tests.parity import relerr
    Y, M, ctl, truth, batch, psiB = _case(G=150, N=1100, seed=31)
    Yw, _ = rstats.winsorize_counts(Y.astype(np.float64))
    keep = (Yw > 0).sum(axis=1) > 0                            # genes the cap empties leave on both sides (R/ruvIIInb.R:57-62)
    W0 = api.init_W(Yw[keep], ctl[keep], 2)
    rng = np.random.default_rng(2)
    sched = [psiB * np.exp(0.2 * rng.standard_normal(psiB.shape) / (i + 1)) for i in range(3)]
    calls = {"i": 0}

    def psi_fn(Yx, offs, wt, psi_old):                        # called once per batch and refresh: the columns of the schedule in turn
        i = calls["i"]; calls["i"] += 1
        return sched[min(i // 2, 2)][keep, i % 2]

    trace = []
    ref = orc.ruvIII_nb(Y.astype(np.float64), M, ctl, k=2, outer_maxit=3, W0=W0, psi_fn=psi_fn, trace=trace, batch=batch, batch_disp=True)
    out = api.ruvIII_nb(Y, M, ctl, k=2, outer_maxit=3, W0=W0, psi=np.stack(sched), batch=batch, batch_disp=True, return_trace=True)
    assert np.array_equal(out["counts"], ref["counts"].astype(np.int32))
    dec = [(r["outer"], r["iter"], r["degener"], r["accepted"]) for r in out["trace"]]
    assert dec == [(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace]
    err = dict(W=relerr(out["W"], ref["W"]), a=relerr(out["a"], ref["a"]), gmean=relerr(out["gmean"], ref["gmean"]),
               Mb=relerr(out["Mb"], ref["Mb"], floor=1e-6), psi=relerr(out["psi"], ref["psi"]), logl=relerr(out["logl"], ref["logl"]))
    assert all(v < 1e-6 for v in err.values()), err
    assert out["psi"].shape == (int(keep.sum()), 2)


@pytest.mark.gpu
def test_live_batch_dispersion_and_get_res():
    """batch.disp = TRUE end to end with the live estimator (one context per batch behind the fit's dispersion hook) against the
    oracle that runs its estimateDisp restatement per batch; then get.res with the G x nbatch psi."""
    from oracle import edger, getres, ruviiinb as orc, rstats
    from ruviiinb_b200 import api
    from tests.parity import relerr
    Y, M, ctl, truth, batch, _ = _case(G=120, N=900, seed=41)
    Yw, _ = rstats.winsorize_counts(Y.astype(np.float64))
    keep = (Yw > 0).sum(axis=1) > 0
    W0 = api.init_W(Yw[keep], ctl[keep], 2)

    def psi_fn(Yx, offs, wt, psi_old):
        return edger.estimate_disp(Yx, offs)["tagwise"]

    trace = []
    ref = orc.ruvIII_nb(Y.astype(np.float64), M, ctl, k=2, outer_maxit=3, W0=W0, psi_fn=psi_fn, trace=trace, batch=batch, batch_disp=True)
    out = api.ruvIII_nb(Y, M, ctl, k=2, outer_maxit=3, W0=W0, batch=batch, batch_disp=True, return_trace=True)
    assert np.array_equal(out["counts"], ref["counts"].astype(np.int32)) and out["psi"].shape == ref["psi"].shape == (int(keep.sum()), 2)
    dec = [(r["outer"], r["iter"], r["degener"], r["accepted"]) for r in out["trace"]]
    assert dec == [(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace]
    err = dict(W=relerr(out["W"], ref["W"]), a=relerr(out["a"], ref["a"]), gmean=relerr(out["gmean"], ref["gmean"]),
               Mb=relerr(out["Mb"], ref["Mb"], floor=1e-6), psi=relerr(out["psi"], ref["psi"]), logl=relerr(out["logl"], ref["logl"]))
    assert all(v < 1e-6 for v in err.values()), err
    assert np.abs(out["psi"][:, 0] / out["psi"][:, 1] - 1.0).max() > 1e-3      # the two batches really carry different dispersions
    # get.res: psi[, curr.batch] in the Pearson residual and the mid-p step, psi[, ref.batch] in the quantile step
    grp = np.argmax(M, axis=1)
    ref_out = dict(counts=ref["counts"], a=ref["a"], W=ref["W"], gmean=ref["gmean"], Mb=ref["Mb"][:, grp], psi=ref["psi"], batch=batch)
    pac = api.get_res(out, type="quantile", block_size=256)
    ref_pac = getres.get_res(ref_out, getres.QUANTILE, block_size=256)
    assert (pac != ref_pac).mean() < 1e-3
    pear = api.get_res(out, type="pearson", block_size=256)
    assert np.abs(pear - getres.get_res(ref_out, getres.PEARSON, block_size=256)).max() < 1e-5
    # and it is not what a single column would give
    one = dict(out, psi=out["psi"][:, 0])
    assert np.abs(api.get_res(one, type="pearson", block_size=256) - pear).max() > 1e-6


@pytest.mark.gpu
def test_pseudosample_fit_matches_oracle():
    """use.pseudosample = TRUE: the same pseudo-cells on both sides (the draws are an injection seam), the fit on the extended matrix
    with the 2-column dispersion of :138-141, outputs trimmed as :618-624."""
    from oracle import edger, pseudocells, ruviiinb as orc, rstats
    from ruviiinb_b200 import api
    from tests.parity import relerr
    Y, M, ctl, truth, batch, _ = _case(G=110, N=800, seed=53)
    N = Y.shape[1]
    Yw, _ = rstats.winsorize_counts(Y.astype(np.float64))
    keep = (Yw > 0).sum(axis=1) > 0
    Yk = Yw[keep].astype(np.int32)
    # both restatements build the same cells from the same generator state
    Ye, Me, be = pseudocells.add_pseudocells(Yk, M, batch, nc_pool=20, batch_disp=False, rng=np.random.default_rng(9))
    Ya, Ma, ba = api.add_pseudocells(Yk, M, batch, nc_pool=20, batch_disp=False, rng=np.random.default_rng(9))
    assert np.array_equal(Ye, Ya) and np.array_equal(Me, Ma) and np.array_equal(be, ba)
    assert Ye.shape[1] > N and Me.shape == (Ye.shape[1], M.shape[1] + M.shape[1]) and set(be.tolist()) == {1, 2}
    assert Me[N:, :M.shape[1]].sum() == 0 and (Me[N:].sum(axis=1) == 1).all()     # pseudo-cells form replicate groups of their own
    W0 = api.init_W(Ye.astype(np.float64), ctl[keep], 2)

    def psi_fn(Yx, offs, wt, psi_old):
        return edger.estimate_disp(Yx, offs)["tagwise"]

    trace = []
    ref = orc.ruvIII_nb(Ye.astype(np.float64), Me, ctl[keep], k=2, outer_maxit=3, W0=W0, psi_fn=psi_fn, trace=trace, winsorize=False,
                        batch=be, batch_disp=True)      # nbatch = 2 whatever batch.disp says (:145)
    out = api.ruvIII_nb(Y, M, ctl, k=2, outer_maxit=3, W0=W0, batch=batch, use_pseudosample=True, pseudo=(Ye, Me, be), return_trace=True)
    dec = [(r["outer"], r["iter"], r["degener"], r["accepted"]) for r in out["trace"]]
    assert dec == [(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace]
    assert out["W"].shape == (N, 2) and out["M"].shape == (N, Me.shape[1]) and out["psi"].shape == (int(keep.sum()),) and out["counts"].shape[1] == N
    err = dict(W=relerr(out["W"], ref["W"][:N]), a=relerr(out["a"], ref["a"]), gmean=relerr(out["gmean"], ref["gmean"]),
               Mb=relerr(out["Mb"], ref["Mb"], floor=1e-6), psi=relerr(out["psi"], ref["psi"][:, 0]), logl=relerr(out["logl"], ref["logl"]))
    assert all(v < 1e-6 for v in err.values()), err


@pytest.mark.gpu
def test_zinb_iteration_with_batch_psi_matches_oracle():
    """ZINB with a batch-specific dispersion (psivec <- psi[i, batch][zeroinf], R/ruvIIInb.R:243-245,258-262,443-465): pi0 by
    Nelder-Mead, the E-step weights and one IRLS pass, psi and psi_w read per batch."""
    from oracle import nelder_mead, ruviiinb as orc
    from ruviiinb_b200 import _lib, synth
    from tests.parity import oracle_init, relerr
    G, N, k, m, nb = 64, 600, 2, 3, 2
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, 18, seed=71, zinb=True, gmean_mu=0.8, gmean_sd=0.8)
    grp = truth["grp"]
    rng = np.random.default_rng(72)
    batch = rng.integers(1, nb + 1, size=N); batch[:nb] = np.arange(1, nb + 1)
    psiB = np.stack([truth["psi"] * np.exp(0.4 * rng.standard_normal(G)) for _ in range(nb)], axis=1)
    psic = psiB[:, batch - 1]
    zeroinf = np.arange(N) % 4 != 0
    Yf = Y.astype(np.float64)
    W0, gmean, alpha, adev = oracle_init(Y, ctl, k)
    mu = np.exp(gmean[:, None] + alpha @ W0.T)
    pi0 = nelder_mead.pi0_optim(Yf[:, zeroinf], mu[:, zeroinf], psic[:, zeroinf])
    pi0 = np.where(np.isnan(pi0), 0.0, pi0)
    wt = orc.zinb_weights(Yf, mu, psic, pi0, zeroinf, np.ones((G, N)))
    state = dict(gmean=gmean, Mb=np.zeros((G, m)), alpha=alpha, alpha_dev=adev, W=W0, psi=psic, step=np.ones(G), wt_zinb=wt, pi0=pi0)
    rep_ind = [np.where(grp == j)[0] for j in range(m)]
    ll0 = orc.loglik_matrix(Yf, gmean[:, None] + alpha @ W0.T, psic, pi0, zeroinf).sum(axis=1)
    new, ll1, aux = orc.inner_iteration(Yf, state, grp, rep_ind, ctl, 100.0, 0.01, 16.0, zeroinf=zeroinf, pi0_fn=nelder_mead.pi0_optim)
    opts = _lib.default_opts(k=k, zinb=1)
    with _lib.Context(0) as ctx:
        ctx.set_batches(batch)
        ctx.set_design(G, N, grp, ctl, zeroinf=zeroinf)
        ctx.upload_counts(Y)
        ctx.set_state("W", W0)
        ctx.init_coef(opts)
        ctx.set_state("psi", psiB)
        assert ctx.zinb_pi0(opts) == 0
        assert relerr(ctx.get_state("pi0"), pi0) < 1e-9
        assert np.array_equal(ctx.get_state("psi_w"), psiB)
        ctx.loglik(opts)
        assert relerr(ctx.get_state("loglik"), ll0) < 1e-10
        tot1, bad = ctx.irls_iteration(opts)
        err = {n: relerr(ctx.get_state(n), new[n], floor=1e-6) for n in ("alpha", "W", "gmean", "Mb", "pi0")}
        err["loglik1"] = abs(tot1 - ll1.sum()) / abs(ll1.sum())
    assert all(v < 1e-8 for v in err.values()), err
