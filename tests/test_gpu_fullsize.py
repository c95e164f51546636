"""BASELINE config 2 at FULL size (20 000 genes x 100 000 cells, k = 5, 50 replicate groups) on the device, checked through
what does not depend on the size: every operator on the path is gene-local once W is fixed, so a sample of rows of the
full-size run is compared with the oracle run on just those rows (all 100 000 cells), and the totals are checked against
their parts.  Synthetic counts are drawn on the GPU exactly as bench.py draws them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cfg2_full_size_sampled_rows_match_oracle():
    import torch
    import bench
    from oracle import edger
    from oracle import ruviiinb as orc
    from ruviiinb_b200 import _lib

    G, N, k, m, n_ctl = bench.WORKLOADS["cfg2_nb_20kx100k"]
    free, _ = torch.cuda.mem_get_info(0)
    if free < 40 * 2**30:
        pytest.skip("needs ~35 GB of HBM")
    dev = torch.device("cuda", 0)
    tp = bench.truth_params(G, N, k, m, n_ctl)
    Yh = torch.empty((G, N), dtype=torch.int32, pin_memory=True).numpy()
    bench.gen_rows_gpu(tp, np.arange(G), Yh, dev)
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    Yh[17] = 0                      # an all-zero gene and a nearly empty one ride along
    Yh[18] = 0; Yh[18, 5] = 3
    grp = np.asarray(tp["grp"])
    opts = _lib.default_opts(k=k)
    with _lib.Context(0) as ctx:
        ctx.set_design(G, N, grp, tp["ctl"])
        ctx.upload_counts(Yh)
        ctx.set_state("W", tp["W0"]); ctx.init_coef(opts); ctx.set_state("psi", tp["psi"])
        trace = [ctx.loglik(opts)]
        for _ in range(2):
            ll, bad = ctx.irls_iteration(opts)
            ctx.accept()
            trace.append(ll)
        total = ctx.loglik(opts)
        per_gene = ctx.get_state("loglik")
        W, alpha, gmean, Mb, psi = (ctx.get_state(n) for n in ("W", "alpha", "gmean", "Mb", "psi"))
        l0, rs = ctx.disp_apl_grid(opts)
    # totals against their parts
    assert abs(total - trace[-1]) <= 1e-12 * abs(total)          # the iteration's returned log-likelihood is the state's
    assert abs(per_gene.sum() - total) <= 1e-12 * abs(total)
    assert np.array_equal(rs, Yh.sum(axis=1, dtype=np.int64).astype(np.float64))
    assert trace[1] > trace[0]                                    # from the initial coefficients the first pass improves
    # sampled rows against the oracle on those rows alone
    rng = np.random.default_rng(5)
    ctl_idx = np.flatnonzero(tp["ctl"])
    s = np.unique(np.concatenate([rng.choice(G, 10, replace=False), ctl_idx[:2], [17, 18, int(np.argmax(rs))]]))
    Ys = Yh[s].astype(np.float64)
    offs = Mb[s][:, grp] + alpha[s] @ W.T
    ref_ll = orc.loglik_matrix(Ys, gmean[s, None] + offs, psi[s]).sum(axis=1)
    assert np.abs(per_gene[s] - ref_ll).max() <= 1e-10 * np.abs(ref_ll).max()
    sel = rs[s] >= 5
    assert not sel[list(s).index(17)] and not sel[list(s).index(18)] and np.isnan(l0[s][~sel]).all()
    ref_l0 = edger.apl_grid(Ys[sel], offs[sel])
    err = np.abs(l0[s][sel] - ref_l0) / np.abs(ref_l0)
    assert err.max() < 1e-9, err.max()


def test_get_res_full_width_sampled_rows_match_oracle():
    """get.res over all 100 000 cells in the reference's 5 000-cell blocks: logcounts and percentile-adjusted counts are
    gene-local within a block (the caps are per gene), so sampled rows are compared with the oracle run on those rows;
    the Pearson clip limits are quantiles over the whole block slab, so there the comparison is on the unclipped entries
    and on the clipped fraction (1 % by construction)."""
    import torch
    import bench
    from oracle import getres
    from ruviiinb_b200 import _lib

    G, N, k, m, n_ctl = 1500, 100000, 5, 50, 200
    dev = torch.device("cuda", 0)
    tp = bench.truth_params(G, N, k, m, n_ctl)
    Yh = torch.empty((G, N), dtype=torch.int32, pin_memory=True).numpy()
    bench.gen_rows_gpu(tp, np.arange(G), Yh, dev)
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    grp = np.asarray(tp["grp"])
    with _lib.Context(0) as ctx:
        ctx.set_design(G, N, grp, tp["ctl"])
        ctx.upload_counts(Yh)
        for name, key in (("W", "W"), ("alpha", "alpha"), ("gmean", "gmean"), ("Mb", "Mb"), ("psi", "psi")):
            ctx.set_state(name, tp[key])
        got = {kind: ctx.get_res(kind, block_size=5000) for kind in (1, 2, 3)}
    s = np.unique(np.random.default_rng(8).choice(G, 5, replace=False))
    out = dict(counts=Yh[s], a=tp["alpha"][s], W=tp["W"], gmean=tp["gmean"][s], Mb=tp["Mb"][s][:, grp], psi=tp["psi"][s])
    ref2 = getres.get_res(out, 2, block_size=5000)
    assert np.allclose(got[2][s], ref2, rtol=1e-12, atol=1e-14)
    ref3 = getres.get_res(out, 3, block_size=5000)
    mism = got[3][s] != ref3
    assert mism.mean() <= 1e-5, (mism.sum(), got[3][s][mism][:5], ref3[mism][:5])     # quantile-boundary ties only
    assert np.array_equal(got[3], np.round(got[3])) and got[3].min() >= 0             # counts
    # Pearson: per block the two clip limits are attained by 0.5 % of the entries each; everything else is the raw residual
    ref1 = getres.get_res(out, 1, block_size=5000)           # same per-gene caps; its clip limits are the sample's own
    for b0 in range(0, N, 5000):
        blk, rb = got[1][:, b0:b0 + 5000], ref1[:, b0:b0 + 5000]
        lo, hi = blk.min(), blk.max()
        frac = ((blk == lo) | (blk == hi)).mean()
        assert 0.009 < frac < 0.0115, (b0, frac)
        inside = (blk[s] > lo) & (blk[s] < hi) & (rb > rb.min()) & (rb < rb.max())
        assert inside.mean() > 0.95
        assert np.abs(blk[s][inside] - rb[inside]).max() <= 1e-10 * np.abs(rb).max()
