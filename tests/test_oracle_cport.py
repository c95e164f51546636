"""The C/OpenMP twin (oracle/irls_oracle.c, the CPU baseline bench.py times) against the NumPy oracle."""
import numpy as np
import pytest

from oracle import cport
from ruviiinb_b200 import synth


@pytest.mark.parametrize("k,m,robust", [(1, 1, False), (2, 3, False), (3, 4, True)])
def test_c_twin_matches_numpy_oracle(k, m, robust):
    if not cport.available():
        pytest.skip("oracle/_ref/libirls_oracle.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    G, N, n_ctl = 60, 420, 16
    Y, M, ctl, truth = synth.make_counts(G, N, k, m, n_ctl, seed=40 + k)
    rng = np.random.default_rng(1)
    W = truth["W"] + 0.2 * rng.standard_normal((N, k))
    W /= np.sqrt((W ** 2).sum(axis=0))
    alpha = truth["alpha"] * 30 + 0.1 * rng.standard_normal((G, k))
    state = dict(gmean=truth["gmean"].copy(), Mb=truth["Mb"] * 0.5, alpha=alpha, alpha_dev=alpha - alpha.mean(axis=0),
                 W=W, psi=truth["psi"].copy(), step=np.where(np.arange(G) % 4 == 0, 0.5, 1.0))
    kh = 1.345 if robust else 100.0
    a, lla = cport.inner_iteration(Y, state, truth["grp"], ctl, k_huber=kh)
    b, llb = cport.inner_iteration(Y, state, truth["grp"], ctl, k_huber=kh, force_numpy=True)
    for n in ("gmean", "Mb", "alpha", "alpha_dev", "W"):
        scale = max(1e-6, np.abs(b[n]).max())
        assert np.abs(a[n] - b[n]).max() / scale < 1e-10, n
    assert np.allclose(lla, llb, rtol=1e-11)
    assert a["bad"] == b["bad"]
