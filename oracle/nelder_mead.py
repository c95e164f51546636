"""R `optim(par, fn)` with its default method: Nelder-Mead as implemented in R's src/appl/optim.c:nmmin
(alpha = 1, beta = 0.5, gamma = 2, abstol = -Inf, reltol = sqrt(.Machine$double.eps), maxit = 500) -- oracle, test-only.
The reference calls it one-dimensionally for the zero-inflation probability: optim(p = -2, fn = logl.zinb, ...)
(R/ruvIIInb.R:247,447).  PARITY UNPINNED: R's C source is not under /root/reference; this restates its published
algorithm (Nash's Compact Numerical Methods, algorithm 19) step for step."""
import numpy as np

BIG = 1.0e35
RELTOL = float(np.sqrt(np.finfo(np.float64).eps))


def nmmin(fn, par, abstol=-np.inf, intol=RELTOL, alpha=1.0, bet=0.5, gamm=2.0, maxit=500):
    """Returns (x, fmin, fncount, fail).  Raises ValueError when fn(par) is not finite (R: error())."""
    B = np.array(par, dtype=np.float64).ravel().copy()
    n = B.size
    f = fn(B.copy())
    if not np.isfinite(f):
        raise ValueError("function cannot be evaluated at initial parameters")
    funcount = 1
    convtol = intol * (abs(f) + intol)
    n1, C = n + 1, n + 2
    P = np.zeros((n1, n + 2))      # rows 0..n-1: coordinates; row n: function values; columns: vertices + centroid
    P[n1 - 1, 0] = f
    P[:n, 0] = B
    L = 1
    size = 0.0
    step = max(0.1 * np.abs(B).max(), 0.0)
    if step == 0.0:
        step = 0.1
    for j in range(2, n1 + 1):
        P[:n, j - 1] = B
        trystep = step
        while P[j - 2, j - 1] == B[j - 2]:
            P[j - 2, j - 1] = B[j - 2] + trystep
            trystep *= 10
        size += trystep
    oldsize = size
    calcvert = True
    fail = 0
    while True:
        if calcvert:
            for j in range(n1):
                if j + 1 != L:
                    B = P[:n, j].copy()
                    f = fn(B.copy())
                    if not np.isfinite(f):
                        f = BIG
                    funcount += 1
                    P[n1 - 1, j] = f
            calcvert = False
        VL = P[n1 - 1, L - 1]
        VH = VL
        H = L
        for j in range(1, n1 + 1):
            if j != L:
                f = P[n1 - 1, j - 1]
                if f < VL:
                    L = j
                    VL = f
                if f > VH:
                    H = j
                    VH = f
        if VH <= VL + convtol or VL <= abstol:
            break
        for i in range(n):
            temp = -P[i, H - 1]
            for j in range(n1):      # same accumulation order as the C source
                temp += P[i, j]
            P[i, C - 1] = temp / n
        B = (1.0 + alpha) * P[:n, C - 1] - alpha * P[:n, H - 1]
        f = fn(B.copy())
        if not np.isfinite(f):
            f = BIG
        funcount += 1
        VR = f
        if VR < VL:
            P[n1 - 1, C - 1] = f
            for i in range(n):
                f = gamm * B[i] + (1 - gamm) * P[i, C - 1]
                P[i, C - 1] = B[i]
                B[i] = f
            f = fn(B.copy())
            if not np.isfinite(f):
                f = BIG
            funcount += 1
            if f < VR:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = f
            else:
                P[:n, H - 1] = P[:n, C - 1]
                P[n1 - 1, H - 1] = VR
        else:
            if VR < VH:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = VR
            B = (1 - bet) * P[:n, H - 1] + bet * P[:n, C - 1]
            f = fn(B.copy())
            if not np.isfinite(f):
                f = BIG
            funcount += 1
            if f < P[n1 - 1, H - 1]:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = f
            elif VR >= VH:
                calcvert = True
                size = 0.0
                for j in range(n1):
                    if j + 1 != L:
                        for i in range(n):
                            P[i, j] = bet * (P[i, j] - P[i, L - 1]) + P[i, L - 1]
                            size += abs(P[i, j] - P[i, L - 1])
                if size < oldsize:
                    oldsize = size
                else:
                    fail = 10
                    break
        if not funcount <= maxit:
            break
    if funcount > maxit:
        fail = 1
    return P[:n, L - 1].copy(), P[n1 - 1, L - 1], funcount, fail


def pi0_optim(Yz, muz, psi):
    """Per gene: 1/(1 + exp(-optim(p = -2, fn = logl.zinb, y, mu, size = 1/psi)$par))  (R/ruvIIInb.R:239-250, :770-772);
    NaN where optim errors (tryCatch -> NA)."""
    from . import nmath
    G = Yz.shape[0]
    out = np.full(G, np.nan)
    for g in range(G):
        y, mu, size = Yz[g], muz[g], 1.0 / psi[g]
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            nb = nmath.dnbinom_mu(y, size, mu)
        zero = y == 0

        def fn(p):
            with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
                om = 1.0 / (1.0 + np.exp(-p[0]))
                d = om * zero + (1.0 - om) * nb       # ZIM::dzinb
                return float(-np.log(d).sum())
        try:
            x, _, _, _ = nmmin(fn, [-2.0])
            out[g] = 1.0 / (1.0 + np.exp(-x[0]))
        except ValueError:
            pass
    return out
