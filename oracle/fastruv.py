"""NumPy restatement of the full-data passes of `fastruvIII.nb` (reference R/fastruvIIInb.R) -- oracle, test-only:
F5 W for all cells (:632-716), F6 Mb.all (:719-748), F7 the psi cap (:765-770).  The subsample fit that precedes them
(:118-623) is NOT restated here (SURVEY.md section 8 rows F2-F4 are not built yet); its outputs are inputs."""
import numpy as np

from . import rstats


def _w_sweep(Yc, gmean_c, alpha_c, wt_ctl, W, lambda_a):
    """One pass of :639-651 / :667-684 over all cells (the reference's blocks are independent column ranges)."""
    k = W.shape[1]
    lmu = gmean_c[:, None] + alpha_c @ W.T
    Zc = lmu - gmean_c[:, None] + (Yc + 0.01) / (np.exp(lmu) + 0.01) - 1.0
    Wn = np.empty_like(W)
    for j in range(k):
        if j > 0:
            Zc = Zc - np.outer(alpha_c[:, j - 1], Wn[:, j - 1])
        aw = alpha_c[:, j] * wt_ctl
        Wn[:, j] = (Zc.T @ aw) / (lambda_a + aw @ alpha_c[:, j])
    return Wn


def _clamp(W, med, mad):
    W = W.copy()
    for j in range(W.shape[1]):
        hi, lo = med[j] + 4.0 * mad[j], med[j] - 4.0 * mad[j]
        W[:, j] = np.where(W[:, j] > hi, hi, np.where(W[:, j] < lo, lo, W[:, j]))
    return W


def fast_W_all(Y, ctl, gmean, alpha, wt_ctl, W_init, Wsub, lambda_a=0.01, max_sweeps=8, tol=1e-5):
    Yc = Y[ctl].astype(np.float64)
    gm, al = gmean[ctl], alpha[ctl]
    med, mad = rstats.r_median(Wsub, axis=0), rstats.r_mad(Wsub, axis=0)       # :653-654
    W = _clamp(_w_sweep(Yc, gm, al, wt_ctl, W_init, lambda_a), med, mad)        # initial pass :637-661
    sweeps, crit = 1, np.nan
    it = 1
    while True:                                                                 # :664-713
        W_old = W
        W = _clamp(_w_sweep(Yc, gm, al, wt_ctl, W_old, lambda_a), med, mad)
        sweeps += 1
        with np.errstate(divide="ignore", invalid="ignore"):
            crit = np.mean((np.abs(W - W_old) / np.abs(W_old)) ** 2)
        it += 1
        if crit < tol or it > max_sweeps:
            break
    return W, sweeps, crit


def fast_Mb_all(Y, ctl, gmean, alpha, W, psi, Mb_sub, annot_grp, lambda_b=16.0):
    Yf = Y.astype(np.float64)
    G, N = Yf.shape
    annot = annot_grp >= 0
    Mb_all = Yf.copy()                                                          # :721
    Mb_all[:, annot] = Mb_sub[:, annot_grp[annot]]                              # :723
    if annot.sum() < N:
        base = gmean[:, None] + alpha @ W.T
        for _ in range(3):                                                      # :727-738
            with np.errstate(over="ignore", invalid="ignore"):
                lmu = base + Mb_all
                Z = Mb_all + ((Yf + 0.01) / (np.exp(lmu) + 0.01) - 1.0)
                wt = 1.0 / (np.exp(-lmu) + psi[:, None])
                Mb_all = Z * (wt / (wt + lambda_b))
    Mb_all[:, annot] = Mb_sub[:, annot_grp[annot]]                              # :740
    Mb_all[ctl] = 0.0                                                           # :741
    med = rstats.r_median(Mb_all[:, annot], axis=1)[:, None]                    # :744-748
    mad = rstats.r_mad(Mb_all[:, annot], axis=1)[:, None]
    hi, lo = med + 4.0 * mad, med - 4.0 * mad
    return np.where(Mb_all > hi, hi, np.where(Mb_all < lo, lo, Mb_all))


def cap_psi(psi):
    lp = np.log(psi)
    cap = np.exp(rstats.r_median(lp) + 3.0 * rstats.r_mad(lp))                  # :766-768
    return np.where(psi > cap, cap, psi)


# --------------------------------------------------------------------------------------------------------------
# F3 + F4: the subsample fit (R/fastruvIIInb.R:223-623), nbatch == 1, no pseudo-cells, ortho.W = FALSE.
# Injection seams: U0 (irlba on RM_Z, :228), psi_idx (sample(), :218) and psi_fn (estimateDisp.par, :273-311,:569-575).
# --------------------------------------------------------------------------------------------------------------
def _row_avg(Z, grp, m):
    return np.stack([Z[:, grp == j].mean(axis=1) for j in range(m)], axis=1)          # rowAvg (:857-865)


def _row_wt_avg(Z, Wm, grp, m, lam):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.stack([(Z[:, grp == j] * Wm[:, grp == j]).sum(axis=1) / (Wm[:, grp == j].sum(axis=1) + lam) for j in range(m)], axis=1)


def _clamp_cols(A):
    A = A.copy()
    med, mad = rstats.r_median(A, axis=0), rstats.r_mad(A, axis=0)
    for i in range(A.shape[1]):
        hi, lo = med[i] + 4 * mad[i], med[i] - 4 * mad[i]
        A[A[:, i] > hi, i] = hi
        A[A[:, i] < lo, i] = lo
    return A


def _seq_ridge_W(Zc, alpha_c, wt, lambda_a):
    """:447-457 (and :247-254 with wt = 1): sequential ridge regressions of Z.c on alpha.c[, j]."""
    k = alpha_c.shape[1]
    W = np.zeros((Zc.shape[1], k))
    Zc = Zc.copy()
    for j in range(k):
        if j > 0:
            Zc = Zc - np.outer(alpha_c[:, j - 1], W[:, j - 1])
        aw = alpha_c[:, j] * wt
        W[:, j] = (Zc.T @ aw) / (lambda_a + aw @ alpha_c[:, j])
    return W


def _cell_loglik(Y, lmu, psi):
    from . import nmath
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        t = nmath.dnbinom_mu_log(Y, (1.0 / psi)[:, None], np.exp(lmu))
    t = np.where(np.isfinite(t), t, nmath.LOG_DBL_MIN)
    return t.sum(axis=0)                                                                # per CELL (:334-346)


def fast_fit_sub(Ysub, grp, ctl, k, U0, psi_fn, psi_idx, lambda_a=0.01, lambda_b=16.0, step_fac=0.5, inner_maxit=50,
                 outer_maxit=25, trace=None):
    """Returns dict(alpha, gmean, Mb, Wsub, psi, wt_ctl, logl)."""
    Y = np.asarray(Ysub, dtype=np.float64)
    G, ns = Y.shape
    m = int(grp.max()) + 1
    Z = np.log(Y + 1.0)                                                                 # :224
    gmean = Z.mean(axis=1)
    alpha = Z @ U0 @ np.linalg.inv(U0.T @ U0 + lambda_a * np.eye(k))                    # :230 (chol2inv)
    alpha = _clamp_cols(alpha)                                                          # :233-238
    Wsub = _clamp_cols(_seq_ridge_W(Z[ctl] - gmean[ctl, None], alpha[ctl], np.ones(ctl.sum()), lambda_a))   # :239-262
    Mb = _row_avg(Z - alpha @ Wsub.T - gmean[:, None], grp, m)                          # :270 (lambda ignored: rowAvg)
    Mb[ctl] = 0.0
    offs = alpha @ Wsub.T + Mb[:, grp]
    psi = psi_fn(Y[:, psi_idx], offs[:, psi_idx])                                       # :273-311
    psi = np.where(np.isnan(psi), 0.0, psi)
    conv, it_outer, logl_outer = False, 0, []
    while not conv:
        it_outer += 1
        logl_beta = []
        lmu = gmean[:, None] + Mb[:, grp] + alpha @ Wsub.T
        with np.errstate(over="ignore", divide="ignore"):
            sig_inv = 1.0 / (np.exp(-lmu) + psi[:, None])
        wt_ctl = sig_inv[ctl].mean(axis=1)                                              # :331
        loglik = _cell_loglik(Y, lmu, psi)
        step = np.ones(ns)                                                              # :350
        best = dict(W=Wsub, psi=psi, a=alpha, Mb=Mb, gmean=gmean, wtctl=wt_ctl)
        it, halving, conv_beta = 1, 0, False
        while not conv_beta:
            lmu = gmean[:, None] + Mb[:, grp] + alpha @ Wsub.T
            with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
                mu = np.exp(lmu)
                sig_inv = 1.0 / (np.exp(-lmu) + psi[:, None])
                Z = lmu + ((Y + 0.01) / (mu + 0.01) - 1.0) * step[None, :]              # :379
            RM_Z = Z - _row_avg(Z, grp, m)[:, grp]                                      # :386-387
            RM_W = Wsub - _row_avg(Wsub.T, grp, m).T[grp]
            alpha_old = alpha
            wt_cell = sig_inv.mean(axis=0)                                              # :405
            p98 = rstats.r_quantile7(wt_cell, [0.98])[0]
            wt_cell = np.where(wt_cell >= p98, p98, wt_cell)
            b = (RM_Z * wt_cell[None, :]) @ RM_W
            alpha = b @ np.linalg.inv((RM_W * wt_cell[:, None]).T @ RM_W + lambda_a * np.eye(k))   # :411
            if not np.isfinite(alpha).all():
                alpha = alpha_old
            alpha = _clamp_cols(alpha)                                                  # :421-426
            wt_ctl = sig_inv[ctl].mean(axis=1)                                          # :445
            Wsub = _seq_ridge_W(Z[ctl] - gmean[ctl, None], alpha[ctl], wt_ctl, lambda_a)   # :443-457
            Zres = Z - alpha @ Wsub.T - Mb[:, grp]
            gmean = (Zres * sig_inv).sum(axis=1) / sig_inv.sum(axis=1)                  # :478-480
            Zres = Z - alpha @ Wsub.T - gmean[:, None]
            Mb_old = Mb
            Mb = _row_wt_avg(Zres, sig_inv, grp, m, lambda_b)                           # :488
            Mb[ctl] = 0.0
            if np.isnan(Mb).any():
                Mb = Mb_old.copy()
            Mb = np.where(np.isfinite(Mb), Mb, 0.0)
            Mb = _clamp_cols(Mb.T).T                                                    # :499-505
            lmu = gmean[:, None] + Mb[:, grp] + alpha @ Wsub.T
            loglik_tmp = _cell_loglik(Y, lmu, psi)
            s_old, s_new = loglik.sum(), loglik_tmp.sum()
            degener = bool(s_old > s_new) if np.isfinite(s_old) and np.isfinite(s_new) else True
            rec = dict(outer=it_outer, iter=it, degener=degener, logl_new=s_new)
            if degener:                                                                 # :535-545
                step = np.where(loglik > loglik_tmp, step * step_fac, step)
                Wsub, alpha, Mb, psi, gmean, wt_ctl = best["W"], best["a"], best["Mb"], best["psi"], best["gmean"], best["wtctl"]
                halving += 1
                if halving >= 3:
                    loglik_tmp = loglik
                    degener = False
            if not degener:
                loglik = loglik_tmp
                logl_beta.append(loglik.sum())
                best = dict(W=Wsub, psi=psi, a=alpha, Mb=Mb, gmean=gmean, wtctl=wt_ctl)
                it += 1
                halving = 0
            rec["accepted"] = not degener
            if trace is not None:
                trace.append(rec)
            conv_logl = it > 2 and (logl_beta[-1] - logl_beta[-2]) / abs(logl_beta[-2]) < 1e-4
            conv_beta = it >= inner_maxit or conv_logl
        updt = True
        if it_outer > 1:
            updt = abs(logl_outer[-1] - loglik.sum()) / abs(logl_outer[-1]) > 1e-8
        offs = best["a"] @ best["W"].T + best["Mb"][:, grp]
        if updt:
            pn = psi_fn(Y[:, psi_idx], offs[:, psi_idx])
            psi = np.where(np.isfinite(pn), pn, psi)
        Wsub, alpha, Mb, gmean = best["W"], best["a"], best["Mb"], best["gmean"]
        best["psi"] = psi
        lmu = gmean[:, None] + Mb[:, grp] + alpha @ Wsub.T
        logl_outer.append(_cell_loglik(Y, lmu, psi).sum())
        if it_outer > 1:
            conv = (logl_outer[-1] - logl_outer[-2]) / abs(logl_outer[-2]) < 1e-4 or it_outer >= outer_maxit
    return dict(alpha=best["a"], gmean=best["gmean"], Mb=best["Mb"], Wsub=best["W"], psi=psi, wt_ctl=best["wtctl"],
                logl=np.array(logl_outer))
