"""ctypes loader for the C/OpenMP oracle (oracle/irls_oracle.c) -- test infrastructure / CPU baseline only.

`inner_iteration` runs ONE inner IRLS iteration of ruvIII.nb (R/ruvIIInb.R:303-484) on the CPU: through
oracle/_ref/libirls_oracle.so when it was built (all host threads, one task per gene / cell like the
reference's bpmapply), else through the NumPy restatement (oracle/ruviiinb.py).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libirls_oracle.so")
_lib = None


def available():
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        vp = C.c_void_p
        _lib.irls_oracle_iteration.argtypes = [vp, C.c_long, C.c_long, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                               C.c_double, C.c_double, C.c_double, vp, vp, C.c_int]
        _lib.irls_oracle_iteration.restype = C.c_int
        _lib.irls_oracle_threads.restype = C.c_int
    return _lib


def threads():
    return _load().irls_oracle_threads() if available() else 1


def inner_iteration(Y, state, grp, ctl, k_huber=100.0, lambda_a=0.01, lambda_b=16.0, nthreads=0, force_numpy=False):
    """state: dict(gmean, Mb [G,m], alpha [G,k], alpha_dev, W [N,k], psi, step).  Returns (new_state, loglik[G])."""
    G, N = Y.shape
    k = state["W"].shape[1]
    m = state["Mb"].shape[1]
    if force_numpy or not available():
        from . import ruviiinb as orc
        rep_ind = [np.where(grp == j)[0] for j in range(m)]
        st = dict(state, wt_zinb=np.ones((G, N)), pi0=np.zeros(G))
        new, ll, aux = orc.inner_iteration(Y.astype(np.float64), st, grp, rep_ind, np.asarray(ctl, bool), k_huber, lambda_a, lambda_b)
        out = {n: new[n] for n in ("gmean", "Mb", "alpha", "alpha_dev", "W", "psi", "step")}
        out["bad"] = [aux["n_bad_alpha"], aux["n_bad_W"], aux["n_bad_Mb"]]
        return out, ll
    lib = _load()
    Yc = np.ascontiguousarray(Y, dtype=np.int32)
    g32 = np.ascontiguousarray(grp, dtype=np.int32)
    c8 = np.ascontiguousarray(np.asarray(ctl).astype(np.uint8))
    new = {n: np.ascontiguousarray(state[n], dtype=np.float64).copy() for n in ("gmean", "Mb", "alpha", "alpha_dev", "W")}
    psi = np.ascontiguousarray(state["psi"], dtype=np.float64)
    step = np.ascontiguousarray(state["step"], dtype=np.float64)
    ll = np.empty(G)
    bad = np.zeros(3, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.irls_oracle_iteration(p(Yc), G, N, k, m, p(g32), p(c8), p(new["gmean"]), p(new["Mb"]), p(new["alpha"]),
                                   p(new["alpha_dev"]), p(new["W"]), p(psi), p(step), k_huber, lambda_a, lambda_b, p(ll), p(bad),
                                   int(nthreads))
    if rc != 0:
        raise MemoryError("irls_oracle_iteration: allocation failed")
    new["psi"], new["step"], new["bad"] = psi, step, bad.tolist()
    return new, ll
