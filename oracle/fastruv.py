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
