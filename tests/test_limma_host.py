"""The prior-d.f. step of the dispersion estimator on the HOST (csrc/limma_host.h: limma's fitFDist with a covariate and
fitFDistRobustly, stats::lowess, natural-spline regression, F / chi-square functions, Brent root) against the NumPy / SciPy
restatement in oracle/limma.py -- two code bases, one algorithm; no device needed."""
import numpy as np
import pytest


def _case(n, df1, df2, seed, outliers=True):
    rng = np.random.default_rng(seed)
    cov = rng.normal(5.0, 2.0, n)
    x = np.exp(0.3 * np.sin(cov)) * rng.f(df1, df2, n)
    if outliers:
        x[:max(3, n // 150)] *= 50.0
    return x, cov


@pytest.mark.parametrize("n,df1,df2", [(3000, 99.0, 8.0), (20000, 99999.0, 30.0), (500, 49.0, 3.0), (40, 99.0, 6.0)])
@pytest.mark.parametrize("robust", [False, True])
@pytest.mark.parametrize("with_cov", [True, False])
def test_squeeze_var_df_prior_matches_scipy_restatement(n, df1, df2, robust, with_cov):
    from oracle import limma
    from ruviiinb_b200 import _lib
    x, cov = _case(n, df1, df2, seed=int(n + df1))
    c = cov if with_cov else None
    got = _lib.squeeze_var_df_prior(x, df1, c, robust)
    ref = np.broadcast_to(limma.squeeze_var_df_prior(x, df1, c, robust), got.shape)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.abs(got[fin] / ref[fin] - 1).max() < 1e-6     # the Brent root behind df2 stops at 1e-8 on df2 / (1 + df2)
    if robust and n >= 500:
        assert got[0] < 0.5 * np.median(got)                # the planted outliers are shrunk less (smaller prior d.f.)


def test_robust_prior_resists_outliers():
    from ruviiinb_b200 import _lib
    x, cov = _case(5000, 199.0, 10.0, seed=3)
    plain = _lib.squeeze_var_df_prior(x, 199.0, cov, False)[0]
    rob = np.median(_lib.squeeze_var_df_prior(x, 199.0, cov, True))
    assert abs(rob - 10.0) < 1.5 and plain < 8.0            # the moment estimator is dragged down by 0.7 % outliers


def test_no_outliers_means_one_prior_df():
    from ruviiinb_b200 import _lib
    x, cov = _case(4000, 99.0, 12.0, seed=4, outliers=False)
    got = _lib.squeeze_var_df_prior(x, 99.0, cov, True)
    assert abs(np.median(got) - 12.0) < 2.0 and (got == np.median(got)).mean() > 0.95


def test_lowess_reproduces_a_line_and_resists_outliers():
    from oracle import limma
    rng = np.random.default_rng(1)
    x = np.sort(rng.random(800))
    y = 2.0 * x + 1.0
    assert np.abs(limma.lowess(x, y, 0.4, 3) - y).max() < 1e-12
    y2 = y.copy(); y2[::40] += 50.0
    assert np.abs(limma.lowess(x, y2, 0.4, 3) - y)[5:-5].max() < 0.05
