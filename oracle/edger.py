"""Restatement of the arithmetic `estimateDisp.par` (reference R/estimateDisp.par.R:3-202) delegates to
edgeR / limma / locfit -- oracle, test-only.  PARITY UNPINNED: none of those packages is under /root/reference and R
is not installed; what follows restates their published algorithms:

  glm_one_group        edgeR C++ `glm_one_group` (one-group NB GLM by Newton-Raphson, tol 1e-10, maxit 50), reached from
                       `glmFit` -> `mglmOneWay` -> `mglmOneGroup` for the intercept-only design the reference passes
                       (R/ruvIIInb.R:210: design = as.matrix(rep(1, nsc)))
  adjusted_profile_lik edgeR `adjustedProfileLik` / C++ `compute_apl`: sum w l_NB(y; mu, phi) - 1/2 log sum w mu/(1+phi mu)
  fmm_spline, maximize_interpolant   R `splines.c:fmm_spline` + edgeR C++ `interpolator::find_max`
  ave_log_cpm          edgeR C++ `ave_log_cpm` (prior.count = 2)
  tricube_trend        stands in for `locfitByCol(..., degree = 0)` (locfit evaluates the same local-constant tricube
                       smoother on kd-tree vertices and interpolates; this is the smoother itself) -- documented deviation
  fit_f_dist           limma `fitFDist` without covariate; stands in for `squeezeVar(robust = TRUE, covariate = ...)`
                       (`fitFDistRobustly` + natural-spline trend are not restated) -- documented deviation
  estimate_disp        the control flow of R/estimateDisp.par.R:27-201 with design = 1

`zerofit` partitioning (R/estimateDisp.par.R:78-80: cells with count < 1e-4 AND fitted < 1e-4 are dropped per gene)
is not applied: such cells contribute O(1e-4) to the profile likelihood.
"""
import numpy as np
from scipy import special

GRID_PTS = np.linspace(-10.0, 10.0, 21)          # spline.pts (R/estimateDisp.par.R:33)
GRID_DISP = 0.1 * 2.0 ** GRID_PTS                # spline.disp (:35)


def glm_one_group(y, offset, disp, weights=None, maxit=50, tol=1e-10):
    """Rows of y (G x N) against offsets (G x N): beta_g with mu = exp(beta + offset).  disp scalar or (G,)."""
    y = np.asarray(y, dtype=np.float64)
    G, N = y.shape
    w = np.ones_like(y) if weights is None else np.asarray(weights, dtype=np.float64)
    disp = np.broadcast_to(np.asarray(disp, dtype=np.float64).reshape(-1, 1), (G, 1))
    eo = np.exp(offset)
    with np.errstate(divide="ignore", invalid="ignore"):
        num = np.where(y > 1e-10, y / eo * w, 0.0).sum(axis=1)
        beta = np.log(num / w.sum(axis=1))
    nonzero = (y > 1e-10).any(axis=1)
    beta = np.where(nonzero, beta, -np.inf)
    active = nonzero.copy()
    for _ in range(maxit):
        if not active.any():
            break
        idx = np.where(active)[0]
        mu = np.exp(beta[idx, None]) * eo[idx]
        den = 1.0 + mu * disp[idx]
        dl = ((y[idx] - mu) / den * w[idx]).sum(axis=1)
        info = (mu / den * w[idx]).sum(axis=1)
        step = dl / info
        beta[idx] += step
        active[idx[np.abs(step) < tol]] = False
    return beta


def nb_unit_loglik(y, mu, phi):
    """edgeR `compute_unit_nb_loglik`: log NB pmf with size 1/phi (Poisson when phi == 0)."""
    r = 1.0 / phi
    with np.errstate(divide="ignore", invalid="ignore"):
        lm = np.log(mu + r)
        out = special.gammaln(y + r) - special.gammaln(y + 1.0) - special.gammaln(r) + r * np.log(r) - (y + r) * lm
        out = out + np.where(y > 0, y * np.log(mu), 0.0)
    return out


def adjusted_profile_lik(phi, y, offset, weights=None):
    """APL_g(phi) for the intercept-only design: fit beta at dispersion phi, Cox-Reid adjust."""
    y = np.asarray(y, dtype=np.float64)
    w = np.ones_like(y) if weights is None else np.asarray(weights, dtype=np.float64)
    beta = glm_one_group(y, offset, phi, w)
    beta = np.maximum(beta, -1e8)  # mglmOneWay
    mu = np.exp(beta[:, None] + offset)
    ll = (w * nb_unit_loglik(y, mu, phi)).sum(axis=1)
    ww = (w * mu / (1.0 + phi * mu)).sum(axis=1)
    low = 1e-10
    adj = 0.5 * np.where(ww < low, np.log(low), np.log(np.maximum(ww, low)))
    return ll - adj


def apl_grid(y, offset, weights=None):
    """l0: G x 21 (R/estimateDisp.par.R:84-127)."""
    return np.stack([adjusted_profile_lik(d, y, offset, weights) for d in GRID_DISP], axis=1)


def fmm_spline(x, y):
    """Forsythe-Malcolm-Moler cubic spline coefficients b, c, d (R splines.c:fmm_spline)."""
    n = len(x)
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    b = np.zeros(n); c = np.zeros(n); d = np.zeros(n)
    if n < 3:
        t = y[1] - y[0]
        b[0] = t / (x[1] - x[0]); b[1] = b[0]
        return b, c, d
    nm1 = n - 1
    d[0] = x[1] - x[0]
    c[1] = (y[1] - y[0]) / d[0]
    for i in range(1, nm1):
        d[i] = x[i + 1] - x[i]
        b[i] = 2.0 * (d[i - 1] + d[i])
        c[i + 1] = (y[i + 1] - y[i]) / d[i]
        c[i] = c[i + 1] - c[i]
    b[0] = -d[0]
    b[nm1] = -d[nm1 - 1]
    c[0] = c[nm1] = 0.0
    if n > 3:
        c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0])
        c[nm1] = c[nm1 - 1] / (x[nm1] - x[nm1 - 2]) - c[nm1 - 2] / (x[nm1 - 1] - x[nm1 - 3])
        c[0] = c[0] * d[0] * d[0] / (x[3] - x[0])
        c[nm1] = -c[nm1] * d[nm1 - 1] * d[nm1 - 1] / (x[nm1] - x[nm1 - 3])
    for i in range(1, n):
        t = d[i - 1] / b[i - 1]
        b[i] = b[i] - t * d[i - 1]
        c[i] = c[i] - t * c[i - 1]
    c[nm1] = c[nm1] / b[nm1]
    for i in range(nm1 - 1, -1, -1):
        c[i] = (c[i] - d[i] * c[i + 1]) / b[i]
    b[nm1] = (y[nm1] - y[nm1 - 1]) / d[nm1 - 1] + d[nm1 - 1] * (c[nm1 - 1] + 2.0 * c[nm1])
    for i in range(nm1):
        b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i])
        d[i] = (c[i + 1] - c[i]) / d[i]
        c[i] = 3.0 * c[i]
    c[nm1] = 3.0 * c[nm1]
    d[nm1] = d[nm1 - 1]
    return b, c, d


def _quad_first_root(a, b, c):
    back = b * b - 4.0 * a * c
    if not back >= 0 or a == 0:
        return None
    front = -b / (2.0 * a)
    return front - np.sqrt(back) / (2.0 * a)


def maximize_interpolant_1(x, y):
    """edgeR `interpolator::find_max` for one row."""
    y = np.asarray(y, dtype=np.float64)
    at = int(np.argmax(y))  # first maximum, like the C++ loop with '>'
    maxed, x_max = y[at], x[at]
    b, c, d = fmm_spline(x, y)
    n = len(x)
    if at > 0:
        s = _quad_first_root(3 * d[at - 1], 2 * c[at - 1], b[at - 1])
        if s is not None and 0 < s < x[at] - x[at - 1]:
            t = ((d[at - 1] * s + c[at - 1]) * s + b[at - 1]) * s + y[at - 1]
            if t > maxed:
                maxed, x_max = t, s + x[at - 1]
    if at < n - 1:
        s = _quad_first_root(3 * d[at], 2 * c[at], b[at])
        if s is not None and 0 < s < x[at + 1] - x[at]:
            t = ((d[at] * s + c[at]) * s + b[at]) * s + y[at]
            if t > maxed:
                maxed, x_max = t, s + x[at]
    return x_max


def maximize_interpolant(x, Y):
    return np.array([maximize_interpolant_1(x, row) for row in np.atleast_2d(Y)])


def ave_log_cpm(y, offset, dispersion, prior_count=2.0, weights=None):
    """edgeR `aveLogCPM` with an offset matrix (C++ `ave_log_cpm` + `add_prior`)."""
    y = np.asarray(y, dtype=np.float64)
    lib = np.exp(offset)
    prior = prior_count * lib / lib.mean(axis=1, keepdims=True)
    off2 = np.log(lib + 2.0 * prior)
    beta = glm_one_group(y + prior, off2, dispersion, weights)
    return (beta + np.log(1e6)) / np.log(2.0)


def default_span(ntags):
    """edgeR::WLEB: span = 1 if ntags <= 50 else 0.25 + 0.75 sqrt(50/ntags)."""
    return 1.0 if ntags <= 50 else 0.25 + 0.75 * (50.0 / ntags) ** 0.5


def tricube_trend(l0, covariate, span):
    """Local-constant (degree 0) tricube smoother of every column of l0 against the covariate, nearest-neighbour
    bandwidth = distance to the ceil(span n)-th nearest point (locfit alpha = span, deg = 0)."""
    x = np.asarray(covariate, dtype=np.float64)
    n = len(x)
    q = min(n, max(1, int(np.ceil(span * n))))
    out = np.empty_like(l0)
    for i in range(n):
        dist = np.abs(x - x[i])
        h = np.partition(dist, q - 1)[q - 1]
        if h <= 0:
            wgt = (dist <= 0).astype(np.float64)
        else:
            u = np.minimum(dist / h, 1.0)
            wgt = (1.0 - u ** 3) ** 3
        out[i] = (wgt[:, None] * l0).sum(axis=0) / wgt.sum()
    return out


def trigamma_inverse(x):
    """limma::trigammaInverse (Newton on 1/trigamma)."""
    x = np.asarray(x, dtype=np.float64)
    y = 0.5 + 1.0 / x
    for _ in range(50):
        tri = special.polygamma(1, y)
        dif = tri * (1.0 - tri / x) / special.polygamma(2, y)
        y = y + dif
        if np.max(-dif / y) < 1e-8:
            break
    return y


def fit_f_dist(s2, df1):
    """limma::fitFDist(x, df1) without covariate: moments of log s2 -> (scale, df2)."""
    s2 = np.asarray(s2, dtype=np.float64)
    df1 = np.broadcast_to(np.asarray(df1, dtype=np.float64), s2.shape)
    ok = np.isfinite(s2) & np.isfinite(df1) & (s2 > -1e-15) & (df1 > 1e-15)
    x, d1 = np.maximum(s2[ok], 0.0), df1[ok]
    n = len(x)
    if n == 0:
        return np.nan, np.nan
    if n == 1:
        return float(x[0]), 0.0
    m = np.median(x)
    if m == 0:
        m = 1.0
    x = np.maximum(x, 1e-5 * m)
    z = np.log(x)
    e = z - special.digamma(d1 / 2.0) + np.log(d1 / 2.0)
    emean = e.mean()
    evar = (n / (n - 1.0) * (e - emean) ** 2).mean() - special.polygamma(1, d1 / 2.0).mean()
    if evar > 0:
        df2 = 2.0 * float(trigamma_inverse(np.array([evar]))[0])
        s20 = float(np.exp(emean + special.digamma(df2 / 2.0) - np.log(df2 / 2.0)))
    else:
        df2 = np.inf
        s20 = float(np.exp(emean))
    return s20, df2


def nb_deviance_one_group(y, mu, phi):
    """Sum of unit NB deviances (edgeR `compute_unit_nb_deviance`, exact form for phi > 0)."""
    r = 1.0 / phi
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(y > 0, y * np.log(y / mu), 0.0) - (y + r) * np.log((y + r) / (mu + r))
    return 2.0 * t.sum(axis=1)


def estimate_disp(y, offset, weights=None, prior_df=None, min_row_sum=5.0, robust=True):
    """Control flow of R/estimateDisp.par.R:27-201 for design = 1, trend.method = "locfit", tagwise = TRUE; robust = TRUE is what
    every call site of the reference passes (R/ruvIIInb.R:210,540).  The prior d.f. come from limma::squeezeVar(s2, df = N - 1,
    covariate = AveLogCPM[sel], robust) restated in oracle/limma.py: one value, or one per selected gene when robust.
    Returns dict(common, trended, tagwise, span, prior_df, prior_n, l0, sel, ave_log_cpm)."""
    from . import limma
    y = np.asarray(y, dtype=np.float64)
    G, N = y.shape
    sel = y.sum(axis=1) >= min_row_sum                                   # :29
    l0 = apl_grid(y[sel], offset[sel], None if weights is None else weights[sel])   # :84-127
    overall = maximize_interpolant(GRID_PTS, l0.sum(axis=0)[None, :])[0]  # :130
    common = 0.1 * 2.0 ** overall                                         # :132
    alc = ave_log_cpm(y, offset, common, weights=weights)                 # :134
    span = default_span(int(sel.sum()))
    m0 = tricube_trend(l0, alc[sel], span)                                # :136-139 (locfit stand-in)
    trend = maximize_interpolant(GRID_PTS, m0)
    disp_trend = 0.1 * 2.0 ** trend                                       # :142
    trended = np.full(G, disp_trend[np.argmin(alc[sel])])                 # :143-145
    trended[sel] = disp_trend
    ns = int(sel.sum())
    if prior_df is None:                                                  # :156-169
        beta = glm_one_group(y[sel], offset[sel], disp_trend, None if weights is None else weights[sel])
        mu = np.exp(beta[:, None] + offset[sel])
        dev = nb_deviance_one_group(y[sel], mu, disp_trend[:, None])
        s2 = np.maximum(dev / (N - 1.0), 0.0)                             # df.residual = N - 1 (no zerofit partition)
        prior_df_sel = np.broadcast_to(limma.squeeze_var_df_prior(s2, N - 1.0, alc[sel], robust), (ns,)).astype(np.float64)
    else:
        prior_df_sel = np.full(ns, float(prior_df))
    prior_n = prior_df_sel / (N - 1.0)                                    # :171
    tagwise = trended.copy()
    too_large = prior_n > 1e6                                             # :176
    if not too_large.all():
        temp_n = np.where(too_large, 1e6, prior_n)
        ind = 0.1 * 2.0 ** maximize_interpolant(GRID_PTS, l0 + temp_n[:, None] * m0)   # WLEB individual, :181-184
        tw = tagwise[sel]
        if robust:
            tw[~too_large] = ind[~too_large]                              # :189
        else:
            tw = ind                                                      # :186
        tagwise[sel] = tw
    prior_full = np.full(G, np.inf if robust else prior_df_sel[0])        # :192-198
    prior_full[sel] = prior_df_sel
    return dict(common=common, trended=trended, tagwise=tagwise, span=span, prior_df=prior_full, prior_n=prior_full / (N - 1.0), l0=l0,
                sel=sel, ave_log_cpm=alc, m0=m0)
