"""The edgeR-side restatements (oracle/edger.py) against properties that pin them without R:
FMM spline reproduces cubics exactly and interpolates; the interpolant maximiser finds the analytic maximum of a cubic
bump; the one-group NB GLM solves its score equation; the grid APL peaks near the simulating dispersion."""
import numpy as np

from oracle import edger


def test_fmm_spline_reproduces_a_cubic():
    x = np.linspace(-10, 10, 21)
    p = np.poly1d([0.01, -0.2, 0.5, 3.0])
    b, c, d = edger.fmm_spline(x, p(x))
    # on every segment the spline IS the cubic: compare value and derivative at the midpoint
    for i in range(20):
        s = 0.5 * (x[i + 1] - x[i])
        val = ((d[i] * s + c[i]) * s + b[i]) * s + p(x[i])
        assert abs(val - p(x[i] + s)) < 1e-10
        assert abs(b[i] - p.deriv()(x[i])) < 1e-10


def test_fmm_spline_interpolates_and_is_c1():
    rng = np.random.default_rng(0)
    x = np.linspace(-10, 10, 21)
    y = rng.standard_normal(21).cumsum()
    b, c, d = edger.fmm_spline(x, y)
    for i in range(20):
        h = x[i + 1] - x[i]
        assert abs(((d[i] * h + c[i]) * h + b[i]) * h + y[i] - y[i + 1]) < 1e-9
        assert abs((3 * d[i] * h + 2 * c[i]) * h + b[i] - b[i + 1]) < 1e-9


def test_maximize_interpolant_on_cubic_bump():
    x = edger.GRID_PTS
    p = np.poly1d([-0.002, -0.5, 1.3, 0.0])  # local max where p' = 0
    roots = p.deriv().roots
    xm = [r for r in roots if p.deriv(2)(r) < 0][0]
    got = edger.maximize_interpolant(x, p(x)[None, :])[0]
    assert abs(got - xm) < 1e-9
    # boundary maximum stays on the grid
    assert edger.maximize_interpolant(x, (-x)[None, :])[0] == -10.0


def test_glm_one_group_solves_score_equation():
    rng = np.random.default_rng(1)
    G, N = 12, 300
    off = 0.4 * rng.standard_normal((G, N))
    y = rng.poisson(np.exp(1.0 + off) * rng.gamma(2.0, 0.5, (G, N))).astype(float)
    y[0] = 0.0
    beta = edger.glm_one_group(y, off, 0.5)
    assert beta[0] == -np.inf
    mu = np.exp(beta[1:, None] + off[1:])
    score = ((y[1:] - mu) / (1 + 0.5 * mu)).sum(axis=1)
    assert np.abs(score).max() < 1e-6


def test_apl_grid_peaks_near_truth_and_estimate_runs():
    rng = np.random.default_rng(2)
    G, N = 40, 600
    psi = 0.4
    off = 0.3 * rng.standard_normal((G, N))
    mu = np.exp(rng.normal(1.0, 0.7, G)[:, None] + off)
    y = rng.poisson(rng.gamma(1 / psi, mu * psi)).astype(float)
    l0 = edger.apl_grid(y, off)
    assert l0.shape == (G, 21)
    common = 0.1 * 2.0 ** edger.maximize_interpolant(edger.GRID_PTS, l0.sum(axis=0)[None, :])[0]
    assert 0.3 < common < 0.5
    out = edger.estimate_disp(y, off)
    assert out["tagwise"].shape == (G,) and np.all(out["tagwise"] > 0.1) and np.all(out["tagwise"] < 1.5)
    assert abs(out["common"] - common) < 1e-12
    # a huge prior collapses tagwise onto the trend; a zero prior leaves each gene at its own APL maximum
    o0 = edger.estimate_disp(y, off, prior_df=0.0)
    own = 0.1 * 2.0 ** edger.maximize_interpolant(edger.GRID_PTS, l0)
    assert np.allclose(o0["tagwise"], own, rtol=1e-12)


def test_fit_f_dist_recovers_prior_df():
    rng = np.random.default_rng(3)
    d0, d1, s0 = 8.0, 20.0, 0.7
    s2 = s0 * (rng.chisquare(d1, 20000) / d1) / (rng.chisquare(d0, 20000) / d0)
    sc, df2 = edger.fit_f_dist(s2, d1)
    assert abs(df2 - d0) < 1.0 and abs(sc - s0) < 0.05
