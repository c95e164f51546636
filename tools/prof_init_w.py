#!/usr/bin/env python
"""ruvnb_init_W (H1: truncated SVD of the control rows) at cfg-2 width: wall time per iteration count.  Under
`ncu --metrics gpu__time_duration.sum -k regex:k_svd` the launch list shows where an iteration's time goes."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ruviiinb_b200 import _lib  # noqa: E402
import torch
G, N, k, m, n_ctl = 2000, 100000, 5, 50, 1000
dev = torch.device("cuda", 0)
tp = bench.truth_params(G, N, k, m, n_ctl)
Yh = torch.empty((G, N), dtype=torch.int32, pin_memory=True).numpy()
bench.gen_rows_gpu(tp, np.arange(G), Yh, dev)
torch.cuda.synchronize()
out = {}
with _lib.Context(0) as ctx:
    ctx.set_design(G, N, tp["grp"], tp["ctl"]); ctx.upload_counts(Yh); ctx.winsorize()
    for mi in [int(a) for a in sys.argv[1:]] or [5, 100]:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        W0, sv, it = ctx.init_W(k, max_iter=mi)
        torch.cuda.synchronize(); out[f"init_W_max{mi}"] = {"seconds": round(time.perf_counter() - t0, 4), "iterations": it, "sv": [float(x) for x in sv]}
    Q1 = np.linalg.qr(W0 - 0)[0]; Q2 = np.linalg.qr(tp["W"] - tp["W"].mean(0))[0]
    out["canonical_corr_truth"] = [float(x) for x in np.linalg.svd(Q1.T @ Q2, compute_uv=False)]
print(json.dumps(out))
