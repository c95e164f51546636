"""R `stats` / MASS / DescTools / Rfast semantics the reference relies on (oracle, test-only).

Each function names the R routine it restates and where the reference calls it.
"""
import numpy as np

MAD_CONST = 1.4826  # stats::mad default `constant`


def r_median(x, axis=None):
    """stats::median: mean of the two middle order statistics for even n; NaN if any NaN.

    Used by `mad` (R/ruvIIInb.R:713,723,734,409) and the clamps (:371,428)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        return np.median(x, axis=axis)  # numpy propagates NaN like R's median(na.rm=FALSE)


def r_mad(x, axis=None):
    """stats::mad(x) = 1.4826 * median(|x - median(x)|)  (center = median, constant = 1.4826)."""
    x = np.asarray(x, dtype=np.float64)
    med = r_median(x, axis=axis)
    if axis is not None:
        med = np.expand_dims(med, axis)
    with np.errstate(invalid="ignore"):
        return MAD_CONST * r_median(np.abs(x - med), axis=axis)


def r_quantile7(x, probs):
    """stats::quantile(x, probs, type = 7) on a 1-D vector.

    index = 1 + (n-1) p ; lo = floor(index) ; hi = ceiling(index);
    q = x[lo] ; where index > lo and x[hi] != x[lo]: q = (1-h) x[lo] + h x[hi], h = index - lo.
    Called through DescTools::Winsorize (R/ruvIIInb.R:53), rowQuantiles (R/fastruvIIInb.R:83)
    and quantile() (R/get.res.R:63)."""
    xs = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = xs.size
    probs = np.atleast_1d(np.asarray(probs, dtype=np.float64))
    index = 1.0 + max(n - 1, 0) * probs
    lo = np.floor(index).astype(np.int64)
    hi = np.ceil(index).astype(np.int64)
    qs = xs[lo - 1].copy()
    xh = xs[hi - 1]
    sel = (index > lo) & (xh != qs)
    h = index - lo
    qs[sel] = (1.0 - h[sel]) * qs[sel] + h[sel] * xh[sel]
    return qs


def psi_huber(u, k):
    """MASS::psi.huber(u, k) = pmin(1, k/abs(u)); 0/0 -> NaN propagates like R's pmin."""
    u = np.asarray(u, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = k / np.abs(u)
    out = np.minimum(1.0, r)  # np.minimum propagates NaN, as pmin(na.rm=FALSE) does
    return out


def winsorize_counts(Y):
    """ceiling(t(apply(Y, 1, DescTools::Winsorize, probs = c(0, 0.995))))  (R/ruvIIInb.R:53).

    Per gene: cap = Q_{0.995} (type 7); values above the cap are set to it; then ceiling.
    For integer counts this equals min(Y, ceil(cap)).  Returns (Yw, cap_ceil)."""
    Y = np.asarray(Y)
    G = Y.shape[0]
    caps = np.empty(G, dtype=np.float64)
    for g in range(G):
        caps[g] = r_quantile7(Y[g], [0.995])[0]
    Yw = np.minimum(Y.astype(np.float64), caps[:, None])
    Yw = np.ceil(Yw)
    return Yw, np.ceil(caps)


def lmfit_be(y, x, w):
    """Rfast::lmfit(y, x, w)$be = solve(crossprod(x, w*x), crossprod(x, w*y))."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    xtwx = x.T @ (w[:, None] * x)
    xtwy = x.T @ (w * y)
    return np.linalg.solve(xtwx, xtwy)
