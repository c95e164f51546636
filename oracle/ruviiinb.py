"""NumPy float64 restatement of `ruvIII.nb` (reference R/ruvIIInb.R:27-629) -- oracle, test-only.

Dense G x N float64 temporaries exactly as the R code builds them, so this is only for sizes
that fit (tests use G <= ~2000, N <= ~5000).  Every block cites the reference lines it follows.

Injection seams: the initial W (`irlba`, random start, :167) and the
dispersion routine (`estimateDisp.par`, third-party edgeR arithmetic, :205-221,:535-565) are
arguments, because neither can be restated bit-faithfully offline.

Batch-specific dispersion (`batch.disp = TRUE`, :213-220,278-283,306-315,542-548) is an argument of `ruvIII_nb`; the pseudo-cells
of `use.pseudosample = TRUE` (:83-152) are prepared by oracle/pseudocells.py.  Out of scope: `ortho.W`.
"""
import numpy as np

from . import nmath, rstats
from .nmath import LOG_DBL_MIN


# --------------------------------------------------------------------------------------
# helpers: R/ruvIIInb.R:639-744
# --------------------------------------------------------------------------------------
def projection(rep_ind, Zmat, Wmat=None, lam=0.0):
    """`projection` with `rowAvg` / `rowWtAvg` (R/ruvIIInb.R:639-675): G x m group (weighted) means."""
    cols = []
    for ind in rep_ind:
        if Wmat is None:
            cols.append(Zmat[:, ind].mean(axis=1))  # rowAvg :657-665 (length-1 group: the column itself)
        else:
            num = (Zmat[:, ind] * Wmat[:, ind]).sum(axis=1)  # rowWtAvg :667-675
            den = Wmat[:, ind].sum(axis=1) + lam
            with np.errstate(divide="ignore", invalid="ignore"):
                cols.append(num / den)
    return np.stack(cols, axis=1)


def huber_weight_rows(signdev, k_huber):
    """psi.huber(signdev_g, k = k.huber * mad(signdev_g)/0.6745), row-wise (R/ruvIIInb.R:409-410,713-714)."""
    sigma = rstats.r_mad(signdev, axis=1) / 0.6745
    return rstats.psi_huber(signdev, (k_huber * sigma)[:, None]), sigma


def get_coef_wt(y, wt, W):
    """`get.coef.wt` (R/ruvIIInb.R:704-706): WLS of y on [1, W] with weights wt."""
    X = np.column_stack([np.ones(W.shape[0]), W])
    return rstats.lmfit_be(y, X, wt)


def gene_gram(RM_W, RM_Z, signdev, wt, k_huber):
    """Per-gene weights, Gram and score of `get.coef2` / `get.adev` (R/ruvIIInb.R:712-730), all genes.

    Returns w (G x N, normalised by its row mean), A (G x k x k) and c = RM_W' diag(w) RM_Z (G x k)."""
    hub, _ = huber_weight_rows(signdev, k_huber)
    w = hub * wt
    with np.errstate(divide="ignore", invalid="ignore"):
        w = w / w.mean(axis=1, keepdims=True)
    k = RM_W.shape[1]
    # A_g = RM_W' diag(w_g) RM_W for all genes at once: w (G x N) times the N x k*k outer products
    P = (RM_W[:, :, None] * RM_W[:, None, :]).reshape(RM_W.shape[0], k * k)
    with np.errstate(invalid="ignore"):
        A = (w @ P).reshape(-1, k, k)
        c = (w * RM_Z) @ RM_W
    return w, A, c


def getW_all(Zc, signdev_c, wt_c, alpha_c, k_huber):
    """`getW` for every cell (R/ruvIIInb.R:732-744): sequential one-regressor WLS over control genes.

    Zc, signdev_c, wt_c: n_c x N (control-gene slab of Z - gmean, signdev, sig.inv*wt.zinb)."""
    nc, N = Zc.shape
    k = alpha_c.shape[1]
    sigma = rstats.r_mad(signdev_c, axis=0) / 0.6745
    w = rstats.psi_huber(signdev_c, (k_huber * sigma)[None, :]) * wt_c
    with np.errstate(divide="ignore", invalid="ignore"):
        w = w / w.mean(axis=0, keepdims=True)
        W = np.zeros((N, k))
        resid = Zc.copy()
        for j in range(k):
            a = alpha_c[:, j][:, None]
            W[:, j] = (w * a * resid).sum(axis=0) / (w * a * a).sum(axis=0)  # lmfit, single regressor
            resid = Zc - alpha_c[:, : j + 1] @ W[:, : j + 1].T
    return W


def clamp_cols(A):
    """median +/- 4 mad clamp of each column (R/ruvIIInb.R:371-376)."""
    A = A.copy()
    med = rstats.r_median(A, axis=0)
    mad = rstats.r_mad(A, axis=0)
    for i in range(A.shape[1]):
        hi, lo = med[i] + 4 * mad[i], med[i] - 4 * mad[i]
        with np.errstate(invalid="ignore"):
            A[A[:, i] > hi, i] = hi
            A[A[:, i] < lo, i] = lo
    return A


def clamp_rows(B):
    """median +/- 4 mad clamp of each row (R/ruvIIInb.R:428-433)."""
    return clamp_cols(B.T).T


def _psi_cols(psi):
    """psi as something that broadcasts against G x N: a vector (one dispersion per gene, nbatch == 1: outer(1/psi, ncells),
    R/ruvIIInb.R:286,309,317) or the G x N matrix psi[, batch] of a batch-specific fit (:280,307,313)."""
    psi = np.asarray(psi)
    return psi[:, None] if psi.ndim == 1 else psi


def loglik_matrix(Y, lmu, psi, pi0=None, zeroinf=None):
    """Per-element log-likelihood with +-Inf/NaN -> log(DBL_MIN) (R/ruvIIInb.R:276-291)."""
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        mu = np.exp(lmu)
        size = 1.0 / _psi_cols(psi)
        if pi0 is None or not np.any(pi0 > 0):
            temp = nmath.dnbinom_mu_log(Y, size, mu)
        else:
            # ZIM::dzinb(x, k, lambda, omega, log): omega*1[x==0] + (1-omega)*dnbinom(x, size=k, mu=lambda)
            omega = np.outer(pi0, zeroinf.astype(np.float64))
            dens = omega * (Y == 0) + (1.0 - omega) * nmath.dnbinom_mu(Y, size, mu)
            temp = np.log(dens)
    temp = np.where(np.isfinite(temp), temp, LOG_DBL_MIN)
    return temp


def working_quantities(Y, gmean, Mb, alpha, W, psi, grp, step):
    """lmu.hat, sig.inv, signdev, Z of R/ruvIIInb.R:303-321 (ncells == 1; psi a vector, or psi[, batch] as a G x N matrix)."""
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        lmu = gmean[:, None] + Mb[:, grp] + alpha @ W.T
        mu = np.exp(lmu)
        sig_inv = 1.0 / (np.exp(-lmu) + _psi_cols(psi))
        size = 1.0 / _psi_cols(psi)
        signdev = np.sign(Y - mu) * np.sqrt(
            2.0 * (nmath.dnbinom_mu_log(Y, size, Y) - nmath.dnbinom_mu_log(Y, size, mu)))
        Z = lmu + step[:, None] * ((Y + 0.01) / (mu + 0.01) - 1.0)
    return lmu, sig_inv, signdev, Z


def zinb_weights(Y, mu, psi, pi0, zeroinf, wt_zinb):
    """E-step weights for zero counts in zero-inflated cells (R/ruvIIInb.R:254-264,457-467)."""
    size = 1.0 / _psi_cols(psi)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        d0 = nmath.dnbinom_mu(0.0, size, mu)
        w = 1.0 - pi0[:, None] / (pi0[:, None] + (1.0 - pi0[:, None]) * d0)
    sel = (Y == 0) & zeroinf[None, :]
    out = wt_zinb.copy()
    out[sel] = w[sel]
    return out


def init_W_svd(Y, ctl, k):
    """Stand-in for irlba (R/ruvIIInb.R:160-168): dense LAPACK SVD of log(Yc/rowMeans + 1).

    Sign convention (irlba's is arbitrary): each right singular vector's largest-|.| entry is positive."""
    Yc = Y[ctl]
    Yn = Yc / Yc.mean(axis=1, keepdims=True)
    _, _, vt = np.linalg.svd(np.log(Yn + 1.0), full_matrices=False)
    V = vt[:k].T.copy()
    for j in range(k):
        if V[np.argmax(np.abs(V[:, j])), j] < 0:
            V[:, j] = -V[:, j]
    return V


# --------------------------------------------------------------------------------------
# one inner IRLS iteration, R/ruvIIInb.R:303-484 -- exposed for operator-level parity tests
# --------------------------------------------------------------------------------------
def inner_iteration(Y, state, grp, rep_ind, ctl, k_huber, lambda_a, lambda_b, zeroinf=None, pi0_fn=None):
    """One pass of the body of `while(!conv.beta)` up to the candidate log-likelihood.

    `state` is a dict with gmean, Mb, alpha, alpha_dev, W, psi, step, wt_zinb, pi0.
    Returns (new_state, loglik_tmp, aux) -- acceptance is decided by the caller (H14)."""
    gmean, Mb, alpha, adev, W = state["gmean"], state["Mb"], state["alpha"], state["alpha_dev"], state["W"]
    psi, step, wt_zinb, pi0 = state["psi"], state["step"], state["wt_zinb"], state["pi0"]
    G, N = Y.shape
    k = W.shape[1]
    # :303-321
    lmu, sig_inv, signdev, Z = working_quantities(Y, gmean, Mb, alpha, W, psi, grp, step)
    # :328-340  projections
    RM_Z = Z - projection(rep_ind, Z)[:, grp]
    RM_W = W - projection(rep_ind, W.T).T[grp]  # projection.K1 (k == 1) computes the same group means
    alpha_old = alpha
    # :346-354  alpha mean
    wt = sig_inv * wt_zinb
    w, A, c = gene_gram(RM_W, RM_Z, signdev, wt, k_huber)
    b = c - np.einsum("gij,gj->gi", A, adev)
    with np.errstate(invalid="ignore"):
        try:
            alpha_mean = np.linalg.solve(A.sum(axis=0), b.sum(axis=0))
        except np.linalg.LinAlgError:
            alpha_mean = np.full(k, np.nan)
    # :356-365  alpha dev
    rhs = c - np.einsum("gij,j->gi", A, alpha_mean)
    Areg = A + lambda_a * np.eye(k)[None]
    adev_new = np.full((G, k), np.nan)
    ok = np.isfinite(Areg).all(axis=(1, 2)) & np.isfinite(rhs).all(axis=1)
    if ok.any():
        try:
            adev_new[ok] = np.linalg.solve(Areg[ok], rhs[ok][..., None])[..., 0]
        except np.linalg.LinAlgError:
            pass
    alpha_new = adev_new + alpha_mean[None, :]
    n_bad_alpha = int((~np.isfinite(alpha_new)).sum())
    if n_bad_alpha:
        alpha_new = alpha_old  # :368
    alpha_new = clamp_cols(alpha_new)  # :371-376
    # :380-402  W
    W_old = W
    Zc = (Z - gmean[:, None])[ctl]
    W_new = getW_all(Zc, signdev[ctl], wt[ctl], alpha_new[ctl], k_huber)
    with np.errstate(divide="ignore", invalid="ignore"):
        W_new = W_new / np.sqrt((W_new ** 2).sum(axis=0))[None, :]
    n_bad_W = int((~np.isfinite(W_new)).sum())
    if n_bad_W:
        W_new = W_old
    # :405-415  gmean
    Wa_new = alpha_new @ W_new.T
    Z_res = Z - Wa_new - Mb[:, grp]
    hub, sigma = huber_weight_rows(signdev, k_huber)
    wtvec = hub * sig_inv * wt_zinb
    with np.errstate(divide="ignore", invalid="ignore"):
        gmean_new = (Z_res * (wtvec / wtvec.sum(axis=1, keepdims=True))).sum(axis=1)
    # :418-433  Mb
    Z_res = Z - Wa_new - gmean_new[:, None]
    Mb_new = projection(rep_ind, Z_res, Wmat=wtvec, lam=lambda_b)
    n_bad_Mb = int((~np.isfinite(Mb_new)).sum())
    if np.isnan(Mb_new).any():
        Mb_new = Mb.copy()
    Mb_new = np.where(np.isfinite(Mb_new), Mb_new, 0.0)
    Mb_new = clamp_rows(Mb_new)
    # :435-467  pi0 / ZINB weights
    pi0_new, wt_zinb_new = pi0, wt_zinb
    if zeroinf is not None and zeroinf.any():
        with np.errstate(over="ignore"):
            mu_new = np.exp(Wa_new + Mb_new[:, grp] + gmean_new[:, None])
        tmp = pi0_fn(Y[:, zeroinf], mu_new[:, zeroinf], psi if np.ndim(psi) == 1 else psi[:, zeroinf])   # :443-445
        pi0_new = np.where(np.isnan(tmp), pi0, tmp)
        wt_zinb_new = zinb_weights(Y, mu_new, psi, pi0_new, zeroinf, wt_zinb)
    # :469-484  candidate log-likelihood
    lmu_new = gmean_new[:, None] + Mb_new[:, grp] + Wa_new
    loglik_tmp = loglik_matrix(Y, lmu_new, psi, pi0_new, zeroinf).sum(axis=1)
    new_state = dict(state, gmean=gmean_new, Mb=Mb_new, alpha=alpha_new, alpha_dev=adev_new,  # :356 keeps the new alpha.dev even when alpha is reverted at :368
                     W=W_new, wt_zinb=wt_zinb_new, pi0=pi0_new)
    aux = dict(alpha_mean=alpha_mean, A=A, c=c, sigma=sigma, n_bad_alpha=n_bad_alpha, n_bad_W=n_bad_W,
               n_bad_Mb=n_bad_Mb, Z=Z, signdev=signdev, sig_inv=sig_inv, lmu=lmu, RM_W=RM_W)
    return new_state, loglik_tmp, aux


# --------------------------------------------------------------------------------------
# the driver, R/ruvIIInb.R:27-629
# --------------------------------------------------------------------------------------
def ruvIII_nb(Y, M, ctl, k=2, robust=False, lambda_a=0.01, lambda_b=16.0, step_fac=0.5, inner_maxit=50,
              outer_maxit=25, zeroinf=None, W0=None, psi_fn=None, pi0_fn=None, winsorize=True, trace=None,
              batch=None, batch_disp=False):
    """Restated `ruvIII.nb` (ortho.W = FALSE; pseudo-cells are prepared by the caller, oracle/pseudocells.py).

    batch (N labels, 1..nbatch as in the reference) with batch_disp = TRUE: psi is a G x nbatch matrix, every use reads
    psi[, batch] (:280,307,313) and every refresh runs psi_fn once per batch on that batch's cells (:213-220,542-548).

    Y: G x N counts; M: N x m 0/1 replicate matrix (each cell in exactly one group, :206);
    ctl: length-G bool.  `W0`: injected initial W (N x k) or None -> dense SVD stand-in.
    `psi_fn(Y, offs, wt_zinb, psi_old)` -> psi (G): injected dispersion routine.
    `trace`: optional list that receives one dict per inner iteration (decisions + log-lik).
    Returns the reference's list as a dict (counts, W, M, ctl, logl, a, Mb, gmean, pi0, psi, L.a, L.b)."""
    Y = np.asarray(Y, dtype=np.float64)
    M = np.asarray(M).astype(bool)
    ctl = np.asarray(ctl).astype(bool)
    N = Y.shape[1]
    zeroinf = np.zeros(N, dtype=bool) if zeroinf is None else np.asarray(zeroinf).astype(bool)  # :39-40
    # :53-62  Winsorize + ceiling, drop all-zero genes
    if winsorize:
        Y, _ = rstats.winsorize_counts(Y)
    zero_genes = (Y > 0).sum(axis=1) == 0
    if zero_genes.any():
        Y = Y[~zero_genes]
        ctl = ctl[~zero_genes]
    G = Y.shape[0]
    k_huber = 1.345 if robust else 100.0  # :155
    grp = np.argmax(M, axis=1)  # apply(M,1,which) :50
    rep_ind = [np.where(M[:, j])[0] for j in range(M.shape[1])]  # :171
    m = M.shape[1]
    # :157-168  sf == 1 -> no size-factor offset; W init
    W = init_W_svd(Y, ctl, k) if W0 is None else np.array(W0, dtype=np.float64)
    # :178-198  init gmean/alpha by WLS of log(Y+1) on [1, W], weights Y+1
    Zres0 = np.log(Y + 1.0)
    coef = np.empty((G, k + 1))
    for g in range(G):
        coef[g] = get_coef_wt(Zres0[g], Y[g] + 1.0, W)
    alpha = coef[:, 1:].copy()
    alpha_mean = alpha.mean(axis=0)
    alpha_dev = alpha - alpha_mean[None, :]
    Mb = np.zeros((G, m))
    alpha[~np.isfinite(alpha)] = 0.0  # :195
    gmean = coef[:, 0].copy()
    wt_zinb = np.ones((G, N))
    pi0 = np.zeros(G)
    # :64-81,88-91: without batch labels, or with batch.disp = FALSE, there is ONE dispersion per gene
    bsel = None
    if batch is not None and batch_disp:
        labels, b0 = np.unique(np.asarray(batch), return_inverse=True)
        if labels.size > 1:
            bsel = [np.where(b0 == B)[0] for B in range(labels.size)]

    def refresh_psi(offs, wt, psi_old):
        """estimateDisp.par on all cells (:210,540), or once per batch on that batch's cells (:213-220,542-548)."""
        if bsel is None:
            return psi_fn(Y, offs, wt, psi_old)
        cols = []
        for B, idx in enumerate(bsel):
            try:
                cols.append(psi_fn(Y[:, idx], offs[:, idx], wt[:, idx], None if psi_old is None else psi_old[:, B]))
            except Exception:           # tryCatch(..., error = function(e) NA), :545
                cols.append(np.full(G, np.nan))
        return np.stack(cols, axis=1)

    def cols_of(psi_):
        """psi as the kernels of this file take it: the vector, or psi[, batch] (G x N)."""
        return psi_ if bsel is None else psi_[:, b0]

    # :205-221  initial psi
    offs = alpha @ W.T + Mb[:, grp]
    psi = refresh_psi(offs, wt_zinb, None)
    # :237-264  initial pi0 / weights
    if zeroinf.any():
        with np.errstate(over="ignore"):
            mu = np.exp(offs + gmean[:, None])
        tmp = pi0_fn(Y[:, zeroinf], mu[:, zeroinf], psi if bsel is None else cols_of(psi)[:, zeroinf])
        pi0 = np.where(np.isnan(tmp), pi0, tmp)
        wt_zinb = zinb_weights(Y, mu, cols_of(psi), pi0, zeroinf, wt_zinb)

    conv = False
    iter_outer = 0
    logl_outer = []
    n_inner_total = 0
    while not conv:  # :271
        iter_outer += 1
        logl_beta = []
        lmu = gmean[:, None] + Mb[:, grp] + alpha @ W.T
        loglik = loglik_matrix(Y, lmu, cols_of(psi), pi0, zeroinf).sum(axis=1)  # :276-291
        step = np.ones(G)  # :293
        best = dict(W=W, psi=psi, alpha=alpha, Mb=Mb, alpha_dev=alpha_dev, gmean=gmean, pi0=pi0)  # :296
        it, halving = 1, 0
        conv_beta = False
        while not conv_beta:  # :300
            state = dict(gmean=gmean, Mb=Mb, alpha=alpha, alpha_dev=alpha_dev, W=W, psi=cols_of(psi), step=step,
                         wt_zinb=wt_zinb, pi0=pi0)
            new, loglik_tmp, aux = inner_iteration(Y, state, grp, rep_ind, ctl, k_huber, lambda_a, lambda_b,
                                                   zeroinf, pi0_fn)
            n_inner_total += 1
            gmean, Mb, alpha, alpha_dev, W = new["gmean"], new["Mb"], new["alpha"], new["alpha_dev"], new["W"]
            wt_zinb, pi0 = new["wt_zinb"], new["pi0"]
            # :486-520  accept / halve / revert
            s_old, s_new = loglik.sum(), loglik_tmp.sum()
            degener = bool(s_old > s_new) if np.isfinite(s_old) and np.isfinite(s_new) else True
            rec = dict(outer=iter_outer, iter=it, halving=halving, logl_old=s_old, logl_new=s_new, degener=degener)
            if degener:
                check = loglik > loglik_tmp
                step = step.copy()
                step[check] *= step_fac
                W, alpha, Mb, psi = best["W"], best["alpha"], best["Mb"], best["psi"]
                alpha_dev, gmean, pi0 = best["alpha_dev"], best["gmean"], best["pi0"]
                if zeroinf.any():  # :494-504
                    with np.errstate(over="ignore"):
                        mu = np.exp(alpha @ W.T + Mb[:, grp] + gmean[:, None])
                    wt_zinb = zinb_weights(Y, mu, cols_of(psi), pi0, zeroinf, wt_zinb)
                halving += 1
                if halving >= 3:
                    loglik_tmp = loglik
                    degener = False
            if not degener:
                loglik = loglik_tmp
                logl_beta.append(loglik.sum())
                best = dict(W=W, psi=psi, alpha=alpha, Mb=Mb, alpha_dev=alpha_dev, gmean=gmean, pi0=pi0)
                it += 1
                halving = 0
            rec.update(accepted=not degener, logl=logl_beta[-1] if logl_beta else np.nan)
            if trace is not None:
                trace.append(rec)
            conv_logl = False
            if it > 2:  # :523-525
                conv_logl = (logl_beta[-1] - logl_beta[-2]) / abs(logl_beta[-2]) < 1e-4
            conv_beta = it >= inner_maxit or conv_logl
        # :529-565  psi update
        logl_outer_tmp = loglik.sum()
        updt_psi = True
        if iter_outer > 1:
            updt_psi = abs(logl_outer[-1] - logl_outer_tmp) / abs(logl_outer[-1]) > 1e-8
        offs = best["alpha"] @ best["W"].T + best["Mb"][:, grp]
        if updt_psi:
            psi_new = refresh_psi(offs, wt_zinb, psi)
            okp = np.isfinite(psi_new)
            psi = np.where(okp, psi_new, psi)
        W, alpha, Mb, alpha_dev, gmean, pi0 = (best["W"], best["alpha"], best["Mb"], best["alpha_dev"],
                                                best["gmean"], best["pi0"])
        best["psi"] = psi
        lmu = gmean[:, None] + Mb[:, grp] + alpha @ W.T
        logl_outer.append(loglik_matrix(Y, lmu, cols_of(psi), pi0, zeroinf).sum())  # :577-593
        if iter_outer > 1:  # :596-598
            conv = ((logl_outer[-1] - logl_outer[-2]) / abs(logl_outer[-2]) < 1e-4) or iter_outer >= outer_maxit
    return {"counts": Y, "W": W, "M": M, "ctl": ctl, "logl": np.array(logl_outer), "a": alpha, "Mb": Mb,
            "gmean": gmean, "pi0": pi0, "psi": psi, "L.a": lambda_a, "L.b": lambda_b,
            "n_inner": n_inner_total, "zero_genes": zero_genes}
