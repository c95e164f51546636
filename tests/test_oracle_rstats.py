"""Known-answer tests pinning the R `stats`/MASS/DescTools semantics the oracle restates (oracle/rstats.py).
The expected values are what R prints for the same calls (hand-derivable from the documented definitions:
type-7 quantile, mad constant 1.4826, even-n median = mean of the two middle order statistics)."""
import numpy as np

from oracle import rstats


def test_median_even_odd_nan():
    assert rstats.r_median(np.array([3.0, 1.0, 2.0])) == 2.0
    assert rstats.r_median(np.array([4.0, 1.0, 3.0, 2.0])) == 2.5
    assert np.isnan(rstats.r_median(np.array([1.0, np.nan, 3.0])))


def test_mad_constant_and_axis():
    # mad(c(1,2,3,4,100)) = 1.4826 * median(|x - 3|) = 1.4826 * 1
    assert rstats.r_mad(np.array([1.0, 2.0, 3.0, 4.0, 100.0])) == 1.4826
    # mad(c(1,2,3,4)): median 2.5, |dev| = 1.5,.5,.5,1.5 -> median 1.0
    assert rstats.r_mad(np.array([1.0, 2.0, 3.0, 4.0])) == 1.4826
    X = np.array([[1.0, 2.0, 3.0, 4.0, 100.0], [0.0, 0.0, 0.0, 0.0, 5.0]])
    assert np.array_equal(rstats.r_mad(X, axis=1), np.array([1.4826, 0.0]))


def test_quantile_type7():
    # quantile(c(1,2,3,4,10), .995, type=7): index = 1 + 4*.995 = 4.98 -> .02*4 + .98*10 = 9.88
    q = rstats.r_quantile7(np.array([10.0, 1.0, 3.0, 2.0, 4.0]), [0.0, 0.5, 0.995, 1.0])
    assert np.allclose(q, [1.0, 3.0, 9.88, 10.0], rtol=0, atol=1e-14)
    # ties at the top: no interpolation when x[hi] == x[lo]
    assert rstats.r_quantile7(np.array([0.0, 0.0, 7.0, 7.0]), [0.995])[0] == 7.0


def test_psi_huber_matches_pmin():
    u = np.array([0.5, -2.0, 0.0, np.nan])
    w = rstats.psi_huber(u, 1.0)
    assert w[0] == 1.0 and w[1] == 0.5 and w[2] == 1.0 and np.isnan(w[3])
    # sigma == 0 quirk: pmin(1, 0/|u|) = 0, and 0/0 = NaN
    w0 = rstats.psi_huber(np.array([1.0, 0.0]), 0.0)
    assert w0[0] == 0.0 and np.isnan(w0[1])


def test_winsorize_counts_is_min_with_ceil_cap():
    rng = np.random.default_rng(3)
    Y = rng.poisson(3.0, size=(5, 400)).astype(np.float64)
    Y[0, 7] = 900.0
    Yw, cap = rstats.winsorize_counts(Y)
    assert np.array_equal(Yw, np.minimum(Y, cap[:, None]))
    assert Yw[0, 7] == cap[0] and cap[0] < 900.0
    # 200 values 1..200: index = 1 + 199*.995 = 199.005 -> 199.005 -> ceil 200
    y = np.arange(1.0, 201.0)[None, :]
    _, c = rstats.winsorize_counts(y)
    assert c[0] == 200.0


def test_lmfit_be_is_wls():
    rng = np.random.default_rng(4)
    X = np.column_stack([np.ones(50), rng.standard_normal((50, 2))])
    w = rng.uniform(0.5, 2.0, 50)
    beta = np.array([1.0, -2.0, 0.5])
    y = X @ beta
    assert np.allclose(rstats.lmfit_be(y, X, w), beta, atol=1e-12)
