"""Generates tests/golden/midsize_fit_k5_m50.npz: the ORACLE's whole ruvIII.nb fit (oracle/ruviiinb.py + oracle/edger.py, live
dispersion estimation at every psi refresh) on a 2000 x 10000, k = 5, m = 50 problem -- too slow (~40 min of NumPy) to run
inside the GPU test, so its outputs are frozen here; the inputs are regenerated from the seed by the test.
    python tests/golden/make_midsize_fit.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import edger, ruviiinb as orc  # noqa: E402
from ruviiinb_b200 import synth  # noqa: E402

CASE = dict(G=2000, N=10000, k=5, m=50, n_ctl=200, seed=123, gmean_mu=0.0, gmean_sd=1.2)


def main():
    c = CASE
    Y, M, ctl, truth = synth.make_counts(c["G"], c["N"], c["k"], c["m"], c["n_ctl"], seed=c["seed"], gmean_mu=c["gmean_mu"], gmean_sd=c["gmean_sd"])
    Yf = Y.astype(np.float64)
    W0 = orc.init_W_svd(Yf, ctl, c["k"])
    psis = []

    def psi_fn(Yx, offs, wt, psi_old):
        r = edger.estimate_disp(Yx, offs)["tagwise"]
        psis.append(r.copy())
        return r

    trace = []
    t0 = time.time()
    ref = orc.ruvIII_nb(Yf, M, ctl, k=c["k"], W0=W0, psi_fn=psi_fn, winsorize=False, trace=trace)
    print("oracle fit: %.0f s, %d inner iterations, logl %s" % (time.time() - t0, len(trace), ref["logl"]))
    dec = np.array([(r["outer"], r["iter"], int(r["degener"]), int(r["accepted"])) for r in trace], dtype=np.int32)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "midsize_fit_k5_m50.npz"), W0=W0, decisions=dec,
                        logl_new=np.array([r["logl_new"] for r in trace]), logl=ref["logl"], W=ref["W"], a=ref["a"], gmean=ref["gmean"], Mb=ref["Mb"],
                        psi=ref["psi"], psi_refreshes=np.stack(psis), case=np.array(list(c.values()), dtype=np.float64))


if __name__ == "__main__":
    main()
