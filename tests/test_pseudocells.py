"""CPU: the pseudo-cell construction (R/ruvIIInb.R:83-144) -- quantcut known answers, structure of the extended design, and the
host mirror (ruviiinb_b200/api.py) against the oracle restatement on the same generator state; the oracle's batch-specific
dispersion against its own single-dispersion fit."""
import numpy as np


def test_quantcut_known_answers():
    from oracle import pseudocells
    from ruviiinb_b200 import api
    x = np.array([5.0, 1.0, 9.0, 3.0, 7.0, 2.0, 8.0, 4.0, 6.0, 10.0])
    # quantile(x, c(0, .5, 1)) = 1, 5.5, 10 -> [1, 5.5], (5.5, 10]
    want2 = np.array([1, 1, 2, 1, 2, 1, 2, 1, 2, 2])
    # quantile(x, (0:3)/3) = 1, 4, 7, 10 -> [1, 4], (4, 7], (7, 10]
    want3 = np.array([2, 1, 3, 1, 2, 1, 3, 1, 2, 3])
    for f in (pseudocells.quantcut, api._quantcut):
        assert np.array_equal(f(x, 1), np.ones(10, dtype=int))
        assert np.array_equal(f(x, 2), want2)
        assert np.array_equal(f(x, 3), want3)
        assert np.array_equal(f(np.full(6, 2.0), 3), np.ones(6, dtype=int))      # all breaks tie: one bin


def test_pseudocells_structure_and_mirror_equals_oracle():
    from oracle import pseudocells
    from ruviiinb_b200 import api, synth
    Y, M, ctl, truth = synth.make_counts(40, 400, 2, 3, 10, seed=3, gmean_mu=0.5)
    rng = np.random.default_rng(1)
    batch = rng.integers(1, 3, size=400)
    for bd in (False, True):
        Ye, Me, be = pseudocells.add_pseudocells(Y, M, batch, nc_pool=20, batch_disp=bd, rng=np.random.default_rng(4))
        Ya, Ma, ba = api.add_pseudocells(Y, M, batch, nc_pool=20, batch_disp=bd, rng=np.random.default_rng(4))
        assert np.array_equal(Ye, Ya) and np.array_equal(Me, Ma) and np.array_equal(be, ba)
        n_new = Ye.shape[1] - 400
        # one pseudo-cell per (stratum, batch, library-size bin): max(round(n / 20), 1) bins
        grp = np.argmax(M, axis=1)
        want = sum(max(int(np.round(((batch == b) & (grp == j)).sum() / 20)), 1) for j in range(3) for b in (1, 2))
        assert n_new == want and np.array_equal(Ye[:, :400], Y)
        assert Me.shape == (400 + n_new, 6) and (Me.sum(axis=1) == 1).all() and Me[:400, 3:].sum() == 0
        assert set(be[400:].tolist()) == ({3, 4} if bd else {2}) and np.array_equal(be[:400], batch if bd else np.ones(400, dtype=int))
        # a pseudo-cell is a binomial thinning of its pool: never above the pool's row sums, about their mean
        assert (Ye[:, 400:].sum(axis=1) <= Y.sum(axis=1)).all()
        assert abs(Ye[:, 400:].mean() / Y.mean() - 1.0) < 0.2


def test_oracle_batch_psi_with_equal_columns_is_the_plain_fit():
    from oracle import ruviiinb as orc
    from ruviiinb_b200 import synth
    Y, M, ctl, truth = synth.make_counts(50, 300, 2, 3, 14, seed=6, gmean_mu=0.5)
    W0 = orc.init_W_svd(Y.astype(np.float64), ctl, 2)
    batch = np.random.default_rng(0).integers(1, 3, size=300)
    one = orc.ruvIII_nb(Y.astype(np.float64), M, ctl, k=2, outer_maxit=2, inner_maxit=4, W0=W0, psi_fn=lambda Yx, o, w, p: truth["psi"])
    two = orc.ruvIII_nb(Y.astype(np.float64), M, ctl, k=2, outer_maxit=2, inner_maxit=4, W0=W0, psi_fn=lambda Yx, o, w, p: truth["psi"],
                        batch=batch, batch_disp=True)
    assert two["psi"].shape == (one["psi"].shape[0], 2)
    for n in ("W", "a", "gmean", "Mb", "logl"):
        assert np.allclose(one[n], two[n], rtol=1e-12, atol=1e-12), n
