"""TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package).

CPU restatement of the pseudo-cell construction of `ruvIII.nb(use.pseudosample = TRUE)`, R/ruvIIInb.R:83-144 -- the branch that
runs: `strata <- apply(M, 1, which)` (:50) is never NULL, so the code at :113-141 applies.  gtools::quantcut is third-party
(not under /root/reference; restated from its documentation: quantile breaks, right-closed intervals, lowest value included);
`rbinom` draws come from R's RNG in the reference and from a numpy Generator here, so the cells themselves are an injection
seam (parity unpinned for the draws, pinned for everything else by the tests in tests/test_pseudocells.py)."""
import numpy as np

from . import rstats


def quantcut(x, q):
    """as.numeric(gtools::quantcut(x, q = q))."""
    x = np.asarray(x, dtype=np.float64)
    probs = np.linspace(0.0, 1.0, int(q) + 1)
    br = np.unique(rstats.r_quantile7(x, probs))
    if br.size < 2:
        return np.ones(x.size, dtype=np.int64)
    out = np.empty(x.size, dtype=np.int64)
    for i, v in enumerate(x):                      # cut(x, br, include.lowest = TRUE): (b[j-1], b[j]], the first one closed
        j = 1
        while j < br.size - 1 and v > br[j]:
            j += 1
        out[i] = j
    return out


def add_pseudocells(Y, M, batch, nc_pool=20, batch_disp=False, rng=None):
    """Returns (Y_ext, M_ext, batch_ext) as R/ruvIIInb.R:113-141 leaves Y, M and batch."""
    rng = np.random.default_rng(0) if rng is None else rng
    Y = np.asarray(Y)
    M = np.asarray(M).astype(bool)
    batch = np.asarray(batch).astype(np.int64)
    nsc = Y.shape[1]
    strata = np.array([int(np.where(M[c])[0][0]) for c in range(nsc)])       # :50
    sf = Y.sum(axis=0) / Y.sum(axis=0).mean()                                # :54-55
    bmax = int(batch.max())
    batch_ext = list(batch)
    Ycols = []
    uniq = lambda v: list(dict.fromkeys(list(v)))                            # R's unique(): order of first appearance
    for st in uniq(strata):                                                  # :115
        temp_rows = 0
        for btc in uniq(batch):                                              # :117
            mask = (batch == btc) & (strata == st)
            n = int(mask.sum())
            nq = max(int(np.round(n / nc_pool)), 1)                          # :118 (round half to even in R and numpy)
            if n > 1:                                                        # :119
                qsf = quantcut(sf[mask], nq)                                 # :120
                cols = np.where(mask)[0]
                for qv in range(1, int(qsf.max()) + 1):                      # :121
                    pool = Y[:, cols[qsf == qv]]                             # :122
                    if pool.shape[1] == 0:
                        continue
                    Ycols.append(rng.binomial(pool.sum(axis=1).astype(np.int64), 1.0 / pool.shape[1]))   # :123
                    batch_ext.append(btc + bmax)                             # :126
                    temp_rows += 1                                           # :127
        if (strata == st).sum() > 1 and temp_rows:                           # :132-135
            rows = M.shape[0]
            M = np.vstack([M, np.zeros((temp_rows, M.shape[1]), dtype=bool)])
            col = np.zeros(M.shape[0], dtype=bool)
            col[rows:] = True
            M = np.hstack([M, col[:, None]])
    if not Ycols:
        return Y, M, batch
    Y_ext = np.concatenate([Y, np.stack(Ycols, axis=1).astype(Y.dtype)], axis=1)
    batch_ext = np.asarray(batch_ext, dtype=np.int64)
    if not batch_disp:                                                       # :138-139
        batch_ext = np.where(batch_ext > bmax, 2, 1)
    return Y_ext, M, batch_ext
