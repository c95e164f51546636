"""Synthetic NB / ZINB count data of the shapes BASELINE.json names.

There is no network and the reference's example blob (`data/cellline_sce.rda`) is missing from
its checkout, so every test and benchmark input is generated here from a fixed seed.
"""
import numpy as np

SEED0 = 20261019

CONFIGS = {
    # name: (G, N, k, m, n_ctl, zinb)
    "cfg1_cellline_standin": (12000, 9027, 2, 2, 1085, False),
    "cfg2_nb_20kx100k": (20000, 100000, 5, 50, 1000, False),
    "cfg3_zinb_10kx50k": (10000, 50000, 3, 20, 1000, True),
    "cfg4_fast_30kx1M": (30000, 1000000, 2, 20, 1000, False),
}


def draw_groups(rng, N, m):
    """Replicate-group label per cell: Categorical(Dirichlet(5 1_m)), redrawn until every group has
    >= 2 cells and the sizes are not all equal (the `apply(M,2,which)` matrix trap, R/ruvIIInb.R:171)."""
    while True:
        p = rng.dirichlet(5.0 * np.ones(m))
        grp = rng.choice(m, size=N, p=p).astype(np.int32)
        sizes = np.bincount(grp, minlength=m)
        if sizes.min() >= 2 and (m == 1 or len(set(sizes.tolist())) > 1):
            return grp


def make_counts(G, N, k, m, n_ctl, seed=SEED0, zinb=False, gmean_mu=-0.5, gmean_sd=1.5, chunk=2048,
                grp=None):
    """Draw (Y int32 G x N, truth dict).  Y ~ Poisson(Gamma(shape = 1/psi, scale = mu psi)),
    log mu = gmean + Mb[:, grp] + alpha W'.  The first n_ctl genes are the negative controls (Mb = 0).
    Generated in gene chunks so the 20k x 100k case never holds a dense float64 G x N matrix."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if grp is None:
        grp = draw_groups(rng, N, m)
    W = rng.standard_normal((N, k))
    alpha = 0.3 * rng.standard_normal((G, k))
    gmean = np.clip(gmean_mu + gmean_sd * rng.standard_normal(G), -4.0, 5.0)
    Mb = 0.5 * rng.standard_normal((G, m))
    Mb[:n_ctl] = 0.0
    psi = np.exp(np.log(0.3) + 0.5 * rng.standard_normal(G))
    pi0 = rng.beta(2.0, 5.0, size=G) if zinb else np.zeros(G)
    Y = np.empty((G, N), dtype=np.int32)
    for g0 in range(0, G, chunk):
        g1 = min(G, g0 + chunk)
        lmu = gmean[g0:g1, None] + Mb[g0:g1][:, grp] + alpha[g0:g1] @ W.T
        mu = np.exp(lmu)
        shape = (1.0 / psi[g0:g1])[:, None]
        lam = rng.gamma(shape, mu * psi[g0:g1, None])
        y = rng.poisson(lam)
        if zinb:
            y = np.where(rng.random(y.shape) < pi0[g0:g1, None], 0, y)
        Y[g0:g1] = y.astype(np.int32)
    ctl = np.zeros(G, dtype=bool)
    ctl[:n_ctl] = True
    M = np.zeros((N, m), dtype=np.uint8)
    M[np.arange(N), grp] = 1
    truth = dict(W=W, alpha=alpha, gmean=gmean, Mb=Mb, psi=psi, pi0=pi0, grp=grp)
    return Y, M, ctl, truth
