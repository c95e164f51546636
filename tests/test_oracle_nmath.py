"""The nmath restatements (oracle/nmath.py) against exact fractions and scipy.stats.nbinom."""
import numpy as np
from scipy import stats

from oracle import nmath


def _grid():
    y = np.array([0, 1, 2, 3, 7, 20, 150, 4000], dtype=np.float64)[:, None, None]
    size = np.array([0.05, 0.5, 2.0, 3.3333, 40.0, 1e4])[None, :, None]
    mu = np.array([1e-4, 0.03, 0.7, 1.5, 25.0, 900.0])[None, None, :]
    return np.broadcast_arrays(y, size, mu)


def test_dnbinom_exact_fraction():
    # dnbinom(3, size=2, mu=1.5) = choose(4,3) (4/7)^2 (3/7)^3 = 1728/16807
    v = nmath.dnbinom_mu(3.0, 2.0, 1.5)
    assert abs(float(v) - 1728.0 / 16807.0) < 1e-15
    # dnbinom(0, size=r, mu) = (r/(r+mu))^r
    assert abs(float(nmath.dnbinom_mu(0.0, 0.5, 2.0)) - (0.5 / 2.5) ** 0.5) < 1e-15


def test_dnbinom_log_vs_mpmath_and_scipy():
    """50-digit mpmath is the arbiter (scipy's nbinom.logpmf is itself off by ~1e-9 at size = 1e4, mu = 1e-4)."""
    import mpmath as mp
    mp.mp.dps = 50
    y, size, mu = _grid()
    got = nmath.dnbinom_mu_log(y, size, mu)

    def exact(yy, r, m):
        yy, r, m = mp.mpf(float(yy)), mp.mpf(float(r)), mp.mpf(float(m))
        return float(mp.loggamma(yy + r) - mp.loggamma(r) - mp.loggamma(yy + 1) + r * mp.log(r / (r + m)) + yy * mp.log(m / (r + m)))

    ref = np.vectorize(exact)(y, size, mu)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-13)
    assert np.allclose(got, stats.nbinom.logpmf(y, size, size / (size + mu)), rtol=1e-8, atol=1e-9)


def test_closed_form_equals_saddle_point():
    """The lgamma closed form the CUDA kernels evaluate vs R's saddle-point dnbinom_mu."""
    y, size, mu = _grid()
    a = nmath.nb_loglik_closed(y, size, mu)
    b = nmath.dnbinom_mu_log(y, size, mu)
    assert np.allclose(a, b, rtol=1e-11, atol=1e-9)


def test_pnbinom_qnbinom_vs_scipy():
    y, size, mu = _grid()
    p = size / (size + mu)
    assert np.allclose(nmath.pnbinom_mu(y, size, mu), stats.nbinom.cdf(y, size, p), rtol=1e-11, atol=1e-300)
    assert np.all(nmath.pnbinom_mu(-1.0, size, mu) == 0.0)
    probs = np.array([0.005, 0.2, 0.5, 0.9, 0.995])
    s = np.array([0.5, 2.0, 40.0])[:, None, None]
    m = np.array([0.03, 1.5, 25.0, 300.0])[None, :, None]
    q = nmath.qnbinom_mu(probs[None, None, :], s, m)
    ref = stats.nbinom.ppf(probs[None, None, :], s, s / (s + m))
    assert np.array_equal(q, ref)


def test_log_dbl_min_constant():
    assert nmath.LOG_DBL_MIN == -708.3964185322641
