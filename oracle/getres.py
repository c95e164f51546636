"""NumPy restatement of `get.res` (reference R/get.res.R:12-129), vector psi, NB and ZINB fits -- oracle, test-only.

`out` is a dict with counts (G x N), a (G x k), W (N x k), gmean (G), Mb (G x N per cell -- fastruvIII.nb's Mb.all; a
ruvIII.nb fit must be expanded with Mb[:, grp] first), psi (G).  Blocks of `block_size` cells
are processed independently, exactly as the reference does: the per-gene mean caps (:38-45,77-80) and the Pearson clip
limits (:63-66) are computed WITHIN each block.
"""
import numpy as np

from . import nmath, rstats

PEARSON, LOGCOUNTS, QUANTILE = 1, 2, 3


def _cap_rows(mu):
    """mu winsorised per gene at exp(rowMedians(log mu) + 4 rowMads(log mu))  (:38-45)."""
    with np.errstate(divide="ignore"):
        lm = np.log(mu)
    cap = np.exp(rstats.r_median(lm, axis=1) + 4.0 * rstats.r_mad(lm, axis=1))
    with np.errstate(invalid="ignore"):
        return np.where(mu > cap[:, None], cap[:, None], mu)


def _omega(out, s, e):
    """out$pi0[, curr.batch] (R/get.res.R:53): column batch[c] (1-based) of the G x N matrix pi0 -- as the reference indexes it."""
    if out.get("pi0") is None:
        return None
    batch = out.get("batch")
    batch = np.ones(out["counts"].shape[1], dtype=np.int64) if batch is None else np.asarray(batch).astype(np.int64)
    return np.asarray(out["pi0"])[:, batch[s:e] - 1]


def get_res(out, kind, block_size=5000):
    """ZINB fits (out["pi0"]: the G x N matrix outer(pi0, zeroinf) of R/ruvIIInb.R:611-626): ZIM 1.1.0's one-line definitions are
    restated -- dzinb = omega 1[x = 0] + (1 - omega) dnbinom, pzinb(q) = omega + (1 - omega) pnbinom(q), qzinb(p) =
    qnbinom(pmax(0, (p - omega)/(1 - omega))) (third-party, parity unpinned)."""
    Y = np.asarray(out["counts"], dtype=np.float64)
    G, N = Y.shape
    a, W, gmean, Mb, psi = out["a"], out["W"], out["gmean"], out["Mb"], np.asarray(out["psi"])
    # a batch-specific fit: psi is G x nbatch and read as out$psi[, curr.batch] (:59-62,101-108); the quantile step uses the
    # column ref.batch = which.min(colMeans(out$psi)) (:113,120)
    bpsi = psi.ndim == 2
    if bpsi:
        batch = np.asarray(out["batch"]).astype(np.int64)
        ref = int(np.argmin(psi.mean(axis=0)))
    block_size = min(block_size, N)
    res = np.empty((G, N))
    wa_mean = a @ W.mean(axis=0)                                       # :36
    for s in range(0, N, block_size):
        e = min(s + block_size, N)
        Ysub = Y[:, s:e]
        Wa = a @ W[s:e].T                                              # :32
        mu = _cap_rows(np.exp(Wa + gmean[:, None]))                    # :33,38-40
        mu_full = _cap_rows(np.exp(Wa + gmean[:, None] + Mb[:, s:e]))  # :34,42-44
        om = _omega(out, s, e)
        psic = psi[:, batch[s:e] - 1] if bpsi else psi[:, None]
        if kind == PEARSON:                                            # :49-67
            if om is None:
                dev = (Ysub - mu) / np.sqrt(mu_full * (1.0 + mu_full * psic))
            else:                                                      # :52-53
                dev = (Ysub - mu * (1.0 - om)) / np.sqrt(mu_full * (1.0 - om) * (1.0 + mu_full * (psic + om)))
            v = dev[~np.isnan(dev)]
            lims = rstats.r_quantile7(v, [0.005, 0.995])
            dev = np.where(dev < lims[0], lims[0], dev)
            dev = np.where(dev > lims[1], lims[1], dev)
        elif kind == LOGCOUNTS:                                        # :68-74
            dev = np.log(Ysub / (np.exp(Wa) * (1.0 if om is None else 1.0 - om)) + 1.0)   # :69-73
        else:                                                          # :75-124
            mu_nouv = _cap_rows(np.exp(Mb[:, s:e] + gmean[:, None] + wa_mean[:, None]))
            size = 1.0 / psic
            size_q = (1.0 / psi[:, ref])[:, None] if bpsi else size
            lo = nmath.pnbinom_mu(Ysub - 1.0, size, mu_full)
            d = nmath.dnbinom_mu(Ysub, size, mu_full)
            if om is not None:                                         # ZIM::pzinb(Ysub - 1), + ZIM::dzinb(Ysub), :84-85
                lo = om + (1.0 - om) * lo
                d = om * (Ysub == 0) + (1.0 - om) * d
            hi = lo + d
            p = np.clip((lo + hi) / 2.0, 0.005, 0.995)
            if om is not None:                                         # ZIM::qzinb(p, omega = rowMeans(out$pi0)), :95-96
                avg = np.asarray(out["pi0"])[:, ref][:, None] if bpsi else np.asarray(out["pi0"]).mean(axis=1)[:, None]   # :112,118 / :95
                p = np.maximum(0.0, (p - avg) / (1.0 - avg))
            dev = np.where(p == 0.0, 0.0, nmath.qnbinom_mu(np.where(p == 0.0, 0.5, p), size_q, mu_nouv))
        res[:, s:e] = dev
    return res
