#!/usr/bin/env python
"""Wall-clock of the operators that are NOT on bench.py's line, at BASELINE.json config-2 scale (20k genes x 100k cells,
k = 5): Winsorize, the dispersion step, get.res (Pearson / PAC, one 5000-cell block streamed at a time) and the
fastruvIII.nb full-data passes.  Prints one JSON object; run on a B200 (`gpurun -- python tools/time_ops.py`)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from ruviiinb_b200 import _lib  # noqa: E402


def main():
    import torch
    G, N, k, m, n_ctl = bench.WORKLOADS["cfg2_nb_20kx100k"] if len(sys.argv) < 2 else bench.WORKLOADS[sys.argv[1]]
    dev = torch.device("cuda", 0)
    tp = bench.truth_params(G, N, k, m, n_ctl)
    Yh = torch.empty((G, N), dtype=torch.int32, pin_memory=True).numpy()
    bench.gen_rows_gpu(tp, np.arange(G), Yh, dev)
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    opts = _lib.default_opts(k=k)
    opts.verbose = int(os.environ.get("RUVNB_VERBOSE", "0"))   # stage timings of the dispersion step on stderr
    out = {"workload": f"{G} x {N}, k={k}, m={m}, n_ctl={n_ctl}"}

    def timed(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        out[name + "_s"] = round(time.perf_counter() - t0, 4)
        return r

    with _lib.Context(0) as ctx:
        ctx.set_design(G, N, tp["grp"], tp["ctl"])
        timed("upload_counts_8GB", lambda: ctx.upload_counts(Yh))
        timed("winsorize", lambda: ctx.winsorize())
        ctx.set_state("W", tp["W0"])
        timed("init_coef", lambda: ctx.init_coef(opts))
        ctx.set_state("psi", tp["psi"])
        for _ in range(6):
            ctx.irls_iteration(opts); ctx.accept()
        psi_keep = ctx.get_state("psi")
        timed("estimate_disp_first_call", lambda: ctx.estimate_disp_ex(opts))   # allocates its work space
        ctx.set_state("psi", psi_keep)
        d = timed("estimate_disp", lambda: ctx.estimate_disp_ex(opts))
        out["estimate_disp_common"] = d["common"]; out["estimate_disp_prior_df"] = d["prior_df"]
        timed("disp_newton", lambda: ctx.disp_newton(opts, maxit=30))
        # get.res: time the first two 5000-cell blocks' worth by running on a 10000-cell problem slice is not possible
        # on a resident context, so time the whole pass and report per-block figures
        res = np.empty((G, N), dtype=np.float64, order="F")
        res.fill(1.0)   # really touch the pages (np.zeros maps them lazily): first-touch faults would be timed otherwise
        import ctypes as C
        for kind, nm in ((1, "get_res_pearson"), (2, "get_res_logcounts"), (3, "get_res_pac")):
            timed(nm, lambda: ctx._ck(ctx.lib.ruvnb_get_res(ctx.h, kind, 5000, None, res.ctypes.data_as(C.c_void_p))))
            out[nm + "_Melem_per_s"] = round(G * N / out[nm + "_s"] / 1e6, 1)
        annot = np.where(np.arange(N) % 10 == 0, tp["grp"], -1).astype(np.int32)
        timed("fast_Mb_all", lambda: ctx._ck(ctx.lib.ruvnb_fast_Mb_all(ctx.h, 16.0, annot.ctypes.data_as(C.c_void_p), 5000, res.ctypes.data_as(C.c_void_p))))
        Wt = ctx.get_state("W")
        sub = np.arange(0, N, 40)
        W_init = np.zeros((N, k)); W_init[sub] = Wt[sub]
        med = np.median(Wt[sub], axis=0); mad = 1.4826 * np.median(np.abs(Wt[sub] - med), axis=0)
        W, sweeps, crit = timed("fast_W_all", lambda: ctx.fast_W_all(W_init, np.ones(n_ctl), med, mad))
        out["fast_W_all_sweeps"] = sweeps
    # BASELINE.json config 3: ZINB, k = 3, 10k features x 50k samples, >= 90 % zeros, CSR upload, every cell zero-inflated
    import scipy.sparse as sp
    G3, N3, k3, m3, nc3 = 10000, 50000, 3, 20, 1000
    rng = np.random.Generator(np.random.PCG64(bench.SEED + 1))
    from ruviiinb_b200 import synth
    tp3 = bench.truth_params(G3, N3, k3, m3, nc3, seed=bench.SEED + 1)
    tp3["gmean"] = np.clip(-2.5 + rng.standard_normal(G3), -4.0, 5.0)
    Y3 = np.empty((G3, N3), dtype=np.int32)
    bench.gen_rows_gpu(tp3, np.arange(G3), Y3, dev)
    pi0 = rng.beta(2.0, 5.0, size=G3)
    Y3[rng.random(Y3.shape, dtype=np.float32) < pi0[:, None].astype(np.float32)] = 0
    out["cfg3_zero_fraction"] = float((Y3 == 0).mean())
    csr = sp.csr_matrix(Y3)
    o3 = _lib.default_opts(k=k3, zinb=1)
    with _lib.Context(0) as ctx:
        ctx.set_design(G3, N3, tp3["grp"], tp3["ctl"], zeroinf=np.ones(N3, bool))
        timed("cfg3_upload_csr", lambda: ctx.upload_counts_csr(csr.indptr, csr.indices, csr.data))
        ctx.set_state("W", tp3["W0"])
        ctx.init_coef(o3)
        ctx.set_state("psi", tp3["psi"])
        timed("cfg3_zinb_pi0_init", lambda: ctx.zinb_pi0(o3))
        ctx.loglik(o3)
        ctx.irls_iteration(o3); ctx.accept()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.irls_iteration(o3); ctx.accept()
        torch.cuda.synchronize()
        out["cfg3_zinb_iteration_s"] = round((time.perf_counter() - t0) / 3, 4)
        out["cfg3_zinb_gene_fits_per_s"] = round(G3 / out["cfg3_zinb_iteration_s"], 1)
        timed("cfg3_estimate_disp_weighted", lambda: ctx.estimate_disp_ex(o3))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
