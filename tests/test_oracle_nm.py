"""R's Nelder-Mead (oracle/nelder_mead.py) on problems with known answers, and the 1-D zero-inflation fit."""
import numpy as np

from oracle import nelder_mead as nm


def test_nmmin_quadratic_2d_and_1d():
    x, f, cnt, fail = nm.nmmin(lambda p: (p[0] - 1.0) ** 2 + 3 * (p[1] + 2.0) ** 2 + 5.0, [0.0, 0.0])
    assert fail == 0 and abs(x[0] - 1) < 1e-3 and abs(x[1] + 2) < 1e-3 and abs(f - 5.0) < 1e-6
    x, f, cnt, fail = nm.nmmin(lambda p: np.exp(p[0] - 0.7) - (p[0] - 0.7), [-2.0])   # minimum at 0.7, not symmetric
    assert fail == 0 and abs(x[0] - 0.7) < 1e-3 and cnt < 100


def test_nmmin_first_simplex_is_r_like():
    """optim's first vertex is par + 0.1 |par| (step 0.2 for p = -2): the second evaluation must be at -1.8."""
    seen = []

    def fn(p):
        seen.append(float(p[0]))
        return (p[0] + 1.0) ** 2
    nm.nmmin(fn, [-2.0])
    assert seen[0] == -2.0 and seen[1] == -1.8
    assert seen[2] == 2 * -1.8 - (-2.0)      # reflection of the worse vertex through the better one


def test_pi0_optim_recovers_inflation():
    rng = np.random.default_rng(0)
    N, psi, pi0 = 4000, 0.5, 0.3
    mu = np.exp(rng.normal(0.5, 0.4, N))
    y = rng.poisson(rng.gamma(1 / psi, mu * psi)).astype(float)
    y[rng.random(N) < pi0] = 0.0
    est = nm.pi0_optim(y[None, :], mu[None, :], np.array([psi]))
    assert abs(est[0] - pi0) < 0.04
