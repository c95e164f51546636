#!/usr/bin/env python
"""How much does the stand-in for locfit move the dispersions?  edgeR's WLEB smooths every column of the APL grid l0 against AveLogCPM
with locfit(deg = 0, kern = tricube, nn = span): locfit fits at the vertices of an adaptive tree over the covariate and interpolates
between them, the product (and the oracle) evaluate the same local-constant tricube fit directly at every gene.  Locfit itself cannot be
restated reliably from memory, so this script brackets the effect of its evaluation structure: the same smoother fitted only at V
vertices (quantiles of the covariate, V = 9, 17, 33 -- locfit's default tree has of that order for span ~ 0.3-0.5) and interpolated
linearly / by a cubic spline to the genes, against direct evaluation.  Reported: max and median |delta psi / psi| of the tagwise
dispersion (robust prior), for few cells (prior.n large, fastruvIII.nb's 100-cell estimate) and many cells.  CPU only (NumPy oracle)."""
import json
import sys
import os

import numpy as np
from scipy import interpolate

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import edger, limma  # noqa: E402


def tagwise_with_trend(d, m0):
    N, sel, l0 = d["N"], d["sel"], d["l0"]
    prior_n = d["prior_df_sel"] / (N - 1.0)
    return 0.1 * 2.0 ** edger.maximize_interpolant(edger.GRID_PTS, l0 + np.minimum(prior_n, 1e6)[:, None] * m0)


def run(G, N, seed):
    rng = np.random.default_rng(seed)
    psi = np.exp(rng.normal(np.log(0.3), 0.5, G))
    gm = np.clip(rng.normal(0.5, 1.3, G), -3, 5)
    off = 0.3 * rng.standard_normal((G, N))
    mu = np.exp(gm[:, None] + off)
    y = rng.poisson(rng.gamma(1.0 / psi[:, None], mu * psi[:, None])).astype(np.float64)
    ref = edger.estimate_disp(y, off, robust=True)
    sel = ref["sel"]
    x = ref["ave_log_cpm"][sel]
    d = dict(N=N, sel=sel, l0=ref["l0"], prior_df_sel=ref["prior_df"][sel])
    base = tagwise_with_trend(d, ref["m0"])
    assert np.allclose(base, ref["tagwise"][sel], rtol=1e-12)
    out = {}
    span = ref["span"]
    for V in (9, 17, 33):
        xv = np.quantile(x, np.linspace(0, 1, V))
        # the smoother of the product, evaluated at the vertices only
        q = min(len(x), max(1, int(np.ceil(span * len(x)))))
        mv = np.empty((V, ref["l0"].shape[1]))
        for i, xx in enumerate(xv):
            dist = np.abs(x - xx)
            h = np.partition(dist, q - 1)[q - 1]
            w = (1.0 - np.minimum(dist / h, 1.0) ** 3) ** 3 if h > 0 else (dist <= 0).astype(float)
            mv[i] = (w[:, None] * ref["l0"]).sum(axis=0) / w.sum()
        xu, iu = np.unique(xv, return_index=True)
        for kind in ("linear", "cubic"):
            f = interpolate.interp1d(xu, mv[iu], axis=0, kind=kind if len(xu) > 3 else "linear")
            tw = tagwise_with_trend(d, f(np.clip(x, xu[0], xu[-1])))
            r = np.abs(tw / base - 1.0)
            out[f"V{V}_{kind}"] = {"max": float(r.max()), "median": float(np.median(r)), "p99": float(np.quantile(r, 0.99))}
    return {"G": G, "N": N, "prior_df_median": float(np.median(d["prior_df_sel"])), "prior_n_median": float(np.median(d["prior_df_sel"]) / (N - 1.0)), "rel_change_tagwise": out}


if __name__ == "__main__":
    res = [run(1500, 100, 1), run(1500, 1000, 2)]
    print(json.dumps(res, indent=1))
