#!/usr/bin/env python
"""Per-sweep trace of the dispersion step at config-2 scale (RUVNB_TRACE_DISP=1: time and converged rows per NR sweep)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RUVNB_TRACE_DISP"] = "1"
import bench  # noqa: E402
from ruviiinb_b200 import _lib  # noqa: E402
import torch
G, N, k, m, n_ctl = bench.WORKLOADS["cfg2_nb_20kx100k"]
dev = torch.device("cuda", 0)
tp = bench.truth_params(G, N, k, m, n_ctl)
Yh = torch.empty((G, N), dtype=torch.int32, pin_memory=True).numpy()
bench.gen_rows_gpu(tp, np.arange(G), Yh, dev)
torch.cuda.synchronize(); torch.cuda.empty_cache()
opts = _lib.default_opts(k=k); opts.verbose = 1
with _lib.Context(0) as ctx:
    ctx.set_design(G, N, tp["grp"], tp["ctl"]); ctx.upload_counts(Yh); ctx.winsorize()
    ctx.set_state("W", tp["W0"]); ctx.init_coef(opts); ctx.set_state("psi", tp["psi"])
    for _ in range(6):
        ctx.irls_iteration(opts); ctx.accept()
    ctx.estimate_disp_ex(opts)
