"""NumPy / SciPy restatement of the limma pieces `estimateDisp.par` reaches through `squeezeVar(s2, df, covariate = AveLogCPM,
robust = TRUE, winsor.tail.p = c(0.05, 0.1))` (reference R/estimateDisp.par.R:166-168) -- oracle, test-only.

limma is third-party code that is NOT under /root/reference (DESCRIPTION Imports: limma): this file restates the published
algorithm of limma 3.46+ (`fitFDist`, `fitFDistRobustly`, `loessFit` -> `stats::lowess`, `splines::ns`) from its documentation and
source as remembered -- PARITY UNPINNED (no R here to generate fixtures).  The F / chi-square distribution functions come from
SciPy, so the product's own C++ implementations (csrc/limma_host.h) are checked against an independent code base.

    fit_f_dist(x, df1, covariate)            -> (scale[n] or scalar, df2)
    fit_f_dist_robustly(x, df1, covariate)   -> dict(scale, df2, df2_shrunk[n])
    squeeze_var_df_prior(s2, df, covariate, robust)  -> df.prior (scalar, or per-gene when robust)
"""
import numpy as np
from scipy import optimize, special, stats

from . import rstats


def trigamma_inverse(x):
    """limma::trigammaInverse for a scalar x > 0 (Newton on 1/trigamma, as in the package)."""
    if x > 1e7:
        return 1.0 / np.sqrt(x)
    if x < 1e-6:
        return 1.0 / x
    y = 0.5 + 1.0 / x
    for _ in range(50):
        tri = special.polygamma(1, y)
        dif = tri * (1.0 - tri / x) / special.polygamma(2, y)
        y = y + dif
        if -dif / y < 1e-8:
            break
    return float(y)


def ns_basis(x, df=4):
    """A basis of the natural cubic splines `splines::ns(x, df = df, intercept = TRUE)` spans: boundary knots at range(x),
    df - 2 interior knots at the type-7 quantiles j / (df - 1).  Only the column space matters to lm.fit's fitted values and
    residual sum of squares, so the truncated-power form (Hastie et al., ESL eq. 5.4) on x scaled to [0, 1] is used."""
    x = np.asarray(x, dtype=np.float64)
    lo, hi = x.min(), x.max()
    t = (x - lo) / (hi - lo)
    nik = df - 2
    inner = rstats.r_quantile7(x, [(j + 1.0) / (nik + 1.0) for j in range(nik)]) if nik > 0 else np.array([])
    kn = np.concatenate([[0.0], (np.asarray(inner) - lo) / (hi - lo), [1.0]])
    Kn = len(kn)

    def d(j):
        return (np.maximum(t - kn[j], 0.0) ** 3 - np.maximum(t - kn[Kn - 1], 0.0) ** 3) / (kn[Kn - 1] - kn[j])
    cols = [np.ones_like(t), t] + [d(j) - d(Kn - 2) for j in range(Kn - 2)]
    return np.stack(cols, axis=1)


def fit_f_dist(x, df1, covariate=None):
    """limma::fitFDist: moment estimation of scale and df2 of a scaled F from log x; with a covariate, E[log x] is a natural
    cubic spline in it (4 d.f. for n >= 30)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    df1 = np.broadcast_to(np.asarray(df1, dtype=np.float64), x.shape).copy()
    ok = np.isfinite(df1) & (df1 > 1e-15) & np.isfinite(x) & (x > -1e-15)
    if not ok.all():
        raise ValueError("oracle fit_f_dist: filter non-finite / negative variances before the call")
    splinedf = 1
    if covariate is not None:
        covariate = np.asarray(covariate, dtype=np.float64)
        splinedf = 1 + (n >= 3) + (n >= 6) + (n >= 30)
        splinedf = min(splinedf, len(np.unique(covariate)))
        if splinedf < 2:
            covariate = None
    x = np.maximum(x, 0.0)
    m = np.median(x)
    if m == 0:
        m = 1.0
    x = np.maximum(x, 1e-5 * m)
    z = np.log(x)
    e = z - special.digamma(df1 / 2.0) + np.log(df1 / 2.0)
    if covariate is None:
        emean = e.mean()
        evar = ((e - emean) ** 2).sum() / (n - 1.0)
    else:
        X = ns_basis(covariate, splinedf)
        coef, _, rank, _ = np.linalg.lstsq(X, e, rcond=None)
        emean = X @ coef
        evar = ((e - emean) ** 2).sum() / (n - rank)
    evar = evar - special.polygamma(1, df1 / 2.0).mean()
    if evar > 0:
        df2 = 2.0 * trigamma_inverse(evar)
        s20 = np.exp(emean + special.digamma(df2 / 2.0) - np.log(df2 / 2.0))
    else:
        df2 = np.inf
        s20 = np.exp(emean) if covariate is not None else x.mean()
    return s20, df2


def lowess(x, y, f=2.0 / 3.0, iters=3, delta=None):
    """stats::lowess (Cleveland's `clowess`): x ascending.  Local linear fits with tricube weights over the ns nearest points,
    `iters` robustifying passes with bisquare weights on 6 median|residual|, fits skipped (linearly interpolated) for points
    within `delta` of the last fitted one."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    if delta is None:
        delta = 0.01 * (x[-1] - x[0])
    if n < 2:
        return y.copy()
    ns = max(min(n, int(f * n + 1e-7)), 2)
    ys = np.zeros(n)
    rw = np.ones(n)
    res = np.zeros(n)
    for it in range(iters + 1):
        nleft, nright = 0, ns - 1
        last = -1
        i = 0
        while True:
            if nright < n - 1:
                d1 = x[i] - x[nleft]
                d2 = x[nright + 1] - x[i]
                if d1 > d2:          # the window moves right
                    nleft += 1
                    nright += 1
                    continue
            # fitted value at x[i]
            h = max(x[i] - x[nleft], x[nright] - x[i])
            h9, h1 = 0.999 * h, 0.001 * h
            w = np.zeros(n)
            a = 0.0
            j = nleft
            nrt = nleft
            while j < n:
                r = abs(x[j] - x[i])
                if r <= h9:
                    w[j] = 1.0 if r <= h1 else (1.0 - (r / h) ** 3) ** 3
                    if it > 0:
                        w[j] *= rw[j]
                    a += w[j]
                elif x[j] > x[i]:
                    break
                j += 1
            nrt = j - 1
            if a <= 0.0:
                ok = False
            else:
                ok = True
                w[nleft:nrt + 1] /= a
                if h > 0.0:
                    a = (w[nleft:nrt + 1] * x[nleft:nrt + 1]).sum()
                    b = x[i] - a
                    c = (w[nleft:nrt + 1] * (x[nleft:nrt + 1] - a) ** 2).sum()
                    if np.sqrt(c) > 0.001 * (x[n - 1] - x[0]):
                        b /= c
                        w[nleft:nrt + 1] *= (b * (x[nleft:nrt + 1] - a) + 1.0)
                ys[i] = (w[nleft:nrt + 1] * y[nleft:nrt + 1]).sum()
            if not ok:
                ys[i] = y[i]
            if last < i - 1:       # interpolate the skipped points
                denom = x[i] - x[last]
                for j in range(last + 1, i):
                    alpha = (x[j] - x[last]) / denom
                    ys[j] = alpha * ys[i] + (1.0 - alpha) * ys[last]
            last = i
            cut = x[last] + delta
            i = last + 1
            while i < n:
                if x[i] > cut:
                    break
                if x[i] == x[last]:
                    ys[i] = ys[last]
                    last = i
                i += 1
            i = max(last + 1, i - 1)
            if last >= n - 1:
                break
        res = y - ys
        if it >= iters:
            break
        sc = np.abs(res).mean()           # overall scale estimate
        rw = np.abs(res)
        m1 = n // 2
        srt = np.sort(rw)
        if n % 2 == 0:
            cmad = 3.0 * (srt[m1] + srt[m1 - 1])
        else:
            cmad = 6.0 * srt[m1]
        if cmad < 1e-7 * sc:              # effectively zero
            break
        c9, c1 = 0.999 * cmad, 0.001 * cmad
        r = np.abs(res)
        rw = np.where(r <= c1, 1.0, np.where(r <= c9, (1.0 - (r / cmad) ** 2) ** 2, 0.0))
    return ys


def loess_fit_trend(z, covariate):
    """limma::loessFit(z, covariate, span = 0.4, iterations = 4) without weights: limma orders by x, calls
    stats::lowess(x, y, f = span, iter = iterations - 1) and puts the fit back in the original order."""
    z = np.asarray(z, dtype=np.float64)
    x = np.asarray(covariate, dtype=np.float64)
    o = np.argsort(x, kind="stable")
    fitted = np.empty(len(x))
    fitted[o] = lowess(x[o], z[o], f=0.4, iters=3)
    return fitted


_GL = np.polynomial.legendre.leggauss(128)
_GL_NODES, _GL_WEIGHTS = (_GL[0] + 1.0) / 2.0, _GL[1] / 2.0      # statmod::gauss.quad.prob(128, "uniform")


def _winsorized_moments(df1, df2, tail):
    if df2 > 1e6:      # beyond this the F quantiles / density are the chi-square ones to 1e-6 (the product makes the same cut)
        df2 = np.inf
    if np.isinf(df2):
        fq = stats.chi2.ppf([tail[0], 1.0 - tail[1]], df1) / df1
        dens = lambda f: stats.chi2.pdf(f * df1, df1) * df1
    else:
        fq = stats.f.ppf([tail[0], 1.0 - tail[1]], df1, df2)
        dens = lambda f: stats.f.pdf(f, df1, df2)
    zq = np.log(fq)
    q = fq / (1.0 + fq)
    nodes = q[0] + (q[1] - q[0]) * _GL_NODES
    fnodes = nodes / (1.0 - nodes)
    znodes = np.log(fnodes)
    f = dens(fnodes) / (1.0 - nodes) ** 2
    q21 = q[1] - q[0]
    wtmean = q21 * (_GL_WEIGHTS * f * znodes).sum() + (zq * np.asarray(tail)).sum()
    wtvar = q21 * (_GL_WEIGHTS * f * (znodes - wtmean) ** 2).sum() + ((zq - wtmean) ** 2 * np.asarray(tail)).sum()
    return wtmean, wtvar


def fit_f_dist_robustly(x, df1, covariate=None, winsor_tail_p=(0.05, 0.1)):
    """limma::fitFDistRobustly for a constant df1 (what the device path produces: df.residual = N - 1 for every gene)."""
    x = np.asarray(x, dtype=np.float64).copy()
    n = len(x)
    df1 = float(df1)
    m = np.median(x)
    if m <= 0:
        raise ValueError("x has non-positive median")
    x = np.where(x < m * 1e-12, m * 1e-12, x)
    nr_scale, nr_df2 = fit_f_dist(x, df1, covariate)
    tail = np.array([winsor_tail_p[0], winsor_tail_p[1]], dtype=np.float64)
    prob = np.array([tail[0], 1.0 - tail[1]])
    if (tail < 1.0 / n).all():
        return dict(scale=nr_scale, df2=nr_df2, df2_shrunk=np.full(n, nr_df2))
    z = np.log(x)
    if covariate is None:
        ztrend = stats.trim_mean(z, tail[1])
        zresid = z - ztrend
    else:
        ztrend = loess_fit_trend(z, covariate)
        zresid = z - ztrend
    zrq = rstats.r_quantile7(zresid, prob)
    zwins = np.minimum(np.maximum(zresid, zrq[0]), zrq[1])
    zwmean = zwins.mean()
    zwvar = ((zwins - zwmean) ** 2).mean() * n / (n - 1.0)
    mom_mean, mom_var = _winsorized_moments(df1, np.inf, tail)
    funval_inf = np.log(zwvar / mom_var)
    rank = stats.rankdata  # average ranks, as R's rank()
    if funval_inf <= 0:
        ztc = ztrend + zwmean - mom_mean
        fstat = np.exp(z - ztc)
        tailp = stats.chi2.sf(fstat * df1, df1)
        emp = (n - rank(fstat) + 0.5) / n
        pno = np.minimum(tailp / emp, 1.0)
        shr = np.full(n, np.inf)
        O = pno < 1
        shr[O] = pno[O] * n * df1
        o = np.argsort(tailp, kind="stable")
        shr[o] = np.maximum.accumulate(shr[o])
        return dict(scale=np.exp(ztc), df2=np.inf, df2_shrunk=shr)
    if np.isinf(nr_df2):
        return dict(scale=nr_scale, df2=nr_df2, df2_shrunk=np.full(n, nr_df2))

    def fun(u):
        df2 = u / (1.0 - u)
        return np.log(zwvar / _winsorized_moments(df1, df2, tail)[1])
    rbx = nr_df2 / (1.0 + nr_df2)
    f_low = fun(rbx)
    if f_low >= 0:
        df2 = nr_df2
    else:
        # uniroot(fun, interval = c(rbx, 1), f.lower = funvalLow, f.upper = funvalInf): the upper end is never evaluated by fun
        u = optimize.brentq(lambda t: funval_inf if t >= 1.0 else fun(t), rbx, 1.0, xtol=1e-8, rtol=8.9e-16)
        df2 = u / (1.0 - u) if u < 1.0 else np.inf
    mom_mean, mom_var = _winsorized_moments(df1, df2, tail)
    ztc = ztrend + zwmean - mom_mean
    if np.isinf(df2):
        return dict(scale=np.exp(ztc), df2=np.inf, df2_shrunk=np.full(n, np.inf))
    fstat = np.exp(z - ztc)
    log_tailp = stats.f.logsf(fstat, df1, df2)
    tailp = np.exp(log_tailp)
    log_emp = np.log(n - rank(fstat) + 0.5) - np.log(n)
    log_pno = np.minimum(log_tailp - log_emp, 0.0)
    pno = np.exp(log_pno)
    po = -np.expm1(log_pno)
    if (log_pno < 0).any():
        min_ltp = log_tailp.min()
        if np.isneginf(min_ltp):
            shr = pno * df2
        else:
            df2_out = np.log(0.5) / min_ltp * df2
            new_ltp = stats.f.logsf(fstat.max(), df1, df2_out)
            df2_out = np.log(0.5) / new_ltp * df2_out
            shr = pno * df2 + po * df2_out
        o = np.argsort(log_tailp, kind="stable")
        so = shr[o]
        mm = np.cumsum(so) / np.arange(1, n + 1)
        imin = int(np.argmin(mm))
        so[:imin + 1] = mm[imin]
        shr = np.empty(n)
        shr[o] = np.maximum.accumulate(so)
    else:
        shr = np.full(n, df2)
    return dict(scale=np.exp(ztc), df2=df2, df2_shrunk=shr)


def squeeze_var_df_prior(s2, df, covariate=None, robust=False, winsor_tail_p=(0.05, 0.1)):
    """df.prior of limma::squeezeVar(var = s2, df, covariate, robust, winsor.tail.p): a scalar, or one value per gene when robust."""
    if robust:
        return fit_f_dist_robustly(s2, df, covariate, winsor_tail_p)["df2_shrunk"]
    return fit_f_dist(s2, df, covariate)[1]
