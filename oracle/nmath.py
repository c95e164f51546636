"""Restatement of the R `nmath` negative-binomial routines the reference calls (oracle, test-only).

R's C sources (src/nmath/{dnbinom,dbinom,bd0,stirlerr,pnbinom,qnbinom}.c) are NOT under
/root/reference; this follows their published algorithm (Loader's saddle-point expansion for
the density, the regularised incomplete beta for the CDF, Cornish-Fisher start plus discrete
search with the 64-epsilon left-continuity fuzz for the quantile).  Cross-checked against
scipy.stats.nbinom in tests/test_oracle_nmath.py.

Reference call sites: dnbinom R/ruvIIInb.R:262,280-288,313-317,465,473-481,502,581-589;
pnbinom/dnbinom/qnbinom R/get.res.R:88-99,108-121.
"""
import numpy as np
from scipy import special

M_LN_SQRT_2PI = 0.918938533204672741780329736406
M_LN_2PI = 1.837877066409345483560659472811
DBL_EPSILON = np.finfo(np.float64).eps
DBL_MIN = np.finfo(np.float64).tiny
LOG_DBL_MIN = float(np.log(DBL_MIN))  # log(.Machine$double.xmin) = -708.3964185322641


def stirlerr(n):
    """stirlerr(n) = log(n!) - log(sqrt(2 pi n) (n/e)^n)."""
    n = np.asarray(n, dtype=np.float64)
    out = np.empty_like(n)
    small = n <= 15.0
    with np.errstate(divide="ignore", invalid="ignore"):
        ns = n[small]
        out[small] = special.gammaln(ns + 1.0) - (ns + 0.5) * np.log(ns) + ns - M_LN_SQRT_2PI
        nb = n[~small]
        nn = nb * nb
        S0, S1, S2, S3, S4 = 1 / 12.0, 1 / 360.0, 1 / 1260.0, 1 / 1680.0, 1 / 1188.0
        out[~small] = (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / nb
    return out


def bd0(x, np_):
    """bd0(x, np) = x log(x/np) + np - x, evaluated stably near x ~ np (Loader 2000)."""
    x = np.asarray(x, dtype=np.float64)
    np_ = np.asarray(np_, dtype=np.float64)
    x, np_ = np.broadcast_arrays(x, np_)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = x * np.log(x / np_) + np_ - x
        near = np.abs(x - np_) < 0.1 * (x + np_)
        if near.any():
            xn, nn = x[near], np_[near]
            v = (xn - nn) / (xn + nn)
            s = (xn - nn) * v
            ej = 2.0 * xn * v
            v2 = v * v
            # |v| < 0.1, so 40 terms take v^(2j) below 1e-80: converged for every element
            for j in range(1, 40):
                ej = ej * v2
                s = s + ej / (2 * j + 1)
            out[near] = s
    return out


def _dbinom_raw_log(x, n, p, q):
    """log dbinom_raw(x, n, p, q) for 0 < p, q < 1 (vectorised; x, n may be non-integer)."""
    x, n, p, q = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in (x, n, p, q)))
    with np.errstate(divide="ignore", invalid="ignore"):
        lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q)
        lf = M_LN_2PI + np.log(x) + np.log1p(-x / n)
        out = lc - 0.5 * lf
        x0 = x == 0
        if x0.any():
            lc0 = np.where(p < 0.1, -bd0(n, n * q) - n * p, n * np.log(q))
            out = np.where(x0, np.where(n == 0, 0.0, lc0), out)
        xn = (x == n) & ~x0
        if xn.any():
            lcn = np.where(q < 0.1, -bd0(n, n * p) - n * q, n * np.log(p))
            out = np.where(xn, lcn, out)
        out = np.where(p == 0, np.where(x == 0, 0.0, -np.inf), out)
        out = np.where(q == 0, np.where(x == n, 0.0, -np.inf), out)
    return out


def dnbinom_mu_log(x, size, mu):
    """log dnbinom(x, size = size, mu = mu)  (R dnbinom_mu, give_log = TRUE)."""
    x, size, mu = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in (x, size, mu)))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        p = size / (size + x)
        ans = _dbinom_raw_log(size, x + size, size / (size + mu), mu / (size + mu))
        out = np.log(p) + ans
        # x == 0 branch
        z = np.where(size < mu, np.log(size / (size + mu)), np.log1p(-mu / (size + mu)))
        out = np.where(x == 0, size * z, out)
        # x < 1e-10 * size branch (MM's formula)
        tiny = (x != 0) & (x < 1e-10 * size)
        if tiny.any():
            pp = np.where(size < mu, np.log(size / (1 + size / mu)), np.log(mu / (1 + mu / size)))
            alt = x * pp - mu - special.gammaln(x + 1) + np.log1p(x * (x - 1) / (2 * size))
            out = np.where(tiny, alt, out)
        # Poisson limit for infinite size
        inf_size = ~np.isfinite(size)
        if inf_size.any():
            pois = -stirlerr(x) - bd0(x, mu) - 0.5 * (M_LN_2PI + np.log(x))
            pois = np.where(x == 0, -mu, pois)
            out = np.where(inf_size, pois, out)
        out = np.where((x == 0) & (size == 0), 0.0, out)
        out = np.where((mu < 0) | (size < 0) | np.isnan(mu) | np.isnan(size), np.nan, out)
        out = np.where((x < 0) | ~np.isfinite(x), -np.inf, out)
    return out


def dnbinom_mu(x, size, mu):
    with np.errstate(over="ignore", invalid="ignore"):
        return np.exp(dnbinom_mu_log(x, size, mu))


def nb_loglik_closed(y, size, mu):
    """Closed-form NB log pmf through lgamma: the formula the CUDA kernels evaluate.

    lgamma(y+r) - lgamma(r) - lgamma(y+1) + r log(r/(r+mu)) + y log(mu/(r+mu))."""
    y, size, mu = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in (y, size, mu)))
    with np.errstate(divide="ignore", invalid="ignore"):
        lr = np.log(mu + size)
        return (special.gammaln(y + size) - special.gammaln(size) - special.gammaln(y + 1.0)
                + size * (np.log(size) - lr) + np.where(y > 0, y * (np.log(mu) - lr), 0.0))


def pnbinom_mu(x, size, mu):
    """pnbinom(x, size = size, mu = mu): lower tail.  R: bratio(size, x+1, size/(size+mu), ...)."""
    x, size, mu = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in (x, size, mu)))
    xf = np.floor(x + 1e-7)
    with np.errstate(invalid="ignore", divide="ignore"):
        val = special.betainc(size, np.maximum(xf, 0.0) + 1.0, size / (size + mu))
    return np.where(x < 0, 0.0, val)


def qnbinom_mu(p, size, mu):
    """qnbinom(p, size = size, mu = mu): smallest integer y with pnbinom(y) >= p (1 - 64 eps).

    R's qnbinom starts from a Cornish-Fisher guess and walks the CDF; the value it returns is
    the definition above whatever the starting point.  Vectorised walk (R/get.res.R:99,121)."""
    from scipy.stats import norm
    p, size, mu = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in (p, size, mu)))
    shape = p.shape
    p, size, mu = p.ravel().copy(), size.ravel(), mu.ravel()
    prob = size / (size + mu)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        Q = 1.0 / prob
        P = (1.0 - prob) * Q
        sigma = np.sqrt(size * P * Q)
        gamma = (Q + P) / sigma
        z = norm.ppf(p)
        y = np.rint(size * P + sigma * (z + gamma * (z * z - 1) / 6))
    y = np.where(np.isfinite(y), np.maximum(y, 0.0), 0.0)
    pf = p * (1 - 64 * DBL_EPSILON)
    zc = pnbinom_mu(y, size, mu)
    # search left while cdf(y-1) >= pf ; search right while cdf(y) < pf
    go_left = zc >= pf
    active = np.ones_like(y, dtype=bool)
    for _ in range(10_000_000):
        left = active & go_left
        right = active & ~go_left
        if left.any():
            idx = np.where(left)[0]
            yl = y[idx]
            done = yl == 0
            zl = pnbinom_mu(yl - 1, size[idx], mu[idx])
            done |= zl < pf[idx]
            y[idx[~done]] = yl[~done] - 1
            active[idx[done]] = False
        if right.any():
            idx = np.where(right)[0]
            y[idx] = y[idx] + 1
            zr = pnbinom_mu(y[idx], size[idx], mu[idx])
            active[idx[zr >= pf[idx]]] = False
        if not active.any():
            break
    y = np.where(p + 1.01 * DBL_EPSILON >= 1.0, np.inf, y)
    return y.reshape(shape)
