"""NumPy restatement of `get.res` (reference R/get.res.R:12-129), NB branch with a vector psi -- oracle, test-only.

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


def get_res(out, kind, block_size=5000):
    Y = np.asarray(out["counts"], dtype=np.float64)
    G, N = Y.shape
    a, W, gmean, Mb, psi = out["a"], out["W"], out["gmean"], out["Mb"], out["psi"]
    block_size = min(block_size, N)
    res = np.empty((G, N))
    wa_mean = a @ W.mean(axis=0)                                       # :36
    for s in range(0, N, block_size):
        e = min(s + block_size, N)
        Ysub = Y[:, s:e]
        Wa = a @ W[s:e].T                                              # :32
        mu = _cap_rows(np.exp(Wa + gmean[:, None]))                    # :33,38-40
        mu_full = _cap_rows(np.exp(Wa + gmean[:, None] + Mb[:, s:e]))  # :34,42-44
        if kind == PEARSON:                                            # :49-67
            dev = (Ysub - mu) / np.sqrt(mu_full * (1.0 + mu_full * psi[:, None]))
            v = dev[~np.isnan(dev)]
            lims = rstats.r_quantile7(v, [0.005, 0.995])
            dev = np.where(dev < lims[0], lims[0], dev)
            dev = np.where(dev > lims[1], lims[1], dev)
        elif kind == LOGCOUNTS:                                        # :68-74
            dev = np.log(Ysub / np.exp(Wa) + 1.0)
        else:                                                          # :75-124
            mu_nouv = _cap_rows(np.exp(Mb[:, s:e] + gmean[:, None] + wa_mean[:, None]))
            size = (1.0 / psi)[:, None]
            lo = nmath.pnbinom_mu(Ysub - 1.0, size, mu_full)
            hi = lo + nmath.dnbinom_mu(Ysub, size, mu_full)
            p = np.clip((lo + hi) / 2.0, 0.005, 0.995)
            dev = nmath.qnbinom_mu(p, size, mu_nouv)
        res[:, s:e] = dev
    return res
