"""bench_workloads.py -- the BASELINE.json configurations besides the headline one, as `bench.py --workload ...`:

  cfg5_getres_30kx1M   get.res (Pearson / logcounts / PAC) on 30k genes x 1M cells from a fitted model, 1/2/4/8 GPUs (config 5)
  cfg4_fast_30kx1M     fastruvIII.nb: Winsorize, subsample fit, W for all cells, Mb.all + PAC (config 4)
  cfg3_zinb_10kx50k    ruvIII.nb ZINB inner iterations, k = 3, 10k x 50k, >= 90 % zeros, CSR upload (config 3)
  cfg1_standin_12kx9k  ruvIII.nb NB whole fit on the stand-in for the bundled cellline data, k = 2, 2 groups (config 1)

Cells are sharded over the ranks in whole 5000-cell blocks (strong scaling: the problem is fixed); the counts are generated
on the device and handed to the library slab by slab (`ruvnb_upload_counts_rows`), so no rank needs the matrix in host memory.
Every function returns the dict that bench.py prints as its ONE JSON line.
"""
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))

CELL_WORKLOADS = {
    # name: G, N, k, m, n_ctl, annotated fraction
    "cfg5_getres_30kx1M": (30000, 1000000, 3, 20, 1000, 1.0),
    "cfg4_fast_30kx1M": (30000, 1000000, 2, 20, 1000, 0.1),
    "cells_small_3kx40k": (3000, 40000, 2, 5, 200, 0.1),          # same code path, seconds: smoke / tests
}
BLOCK = 5000


def cells_truth(G, N, k, m, n_ctl, frac, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.dirichlet(5.0 * np.ones(m))
    grp = rng.choice(m, size=N, p=p).astype(np.int32)
    W = rng.standard_normal((N, k)) / np.sqrt(N)                    # unit-norm columns, as a fit leaves them
    alpha = 0.3 * np.sqrt(N) * rng.standard_normal((G, k))
    gmean = np.clip(-0.5 + 1.5 * rng.standard_normal(G), -4.0, 5.0)
    Mb = 0.5 * rng.standard_normal((G, m))
    Mb[:n_ctl] = 0.0
    psi = np.exp(np.log(0.3) + 0.5 * rng.standard_normal(G))
    ctl = np.zeros(G, dtype=bool)
    ctl[:n_ctl] = True
    annotated = rng.random(N) < frac if frac < 1.0 else np.ones(N, dtype=bool)
    return dict(grp=grp, W=W, alpha=alpha, gmean=gmean, Mb=Mb, psi=psi, ctl=ctl, annotated=annotated)


def gen_block_gpu(tp, g0, g1, c0, c1, device, seed):
    """int32 [g1 - g0, c1 - c0] on the device: Y ~ Poisson(Gamma(1/psi, mu psi)), log mu = gmean + Mb[, grp] + alpha W'."""
    import torch
    Wt = torch.as_tensor(tp["W"][c0:c1], device=device)
    grp = torch.as_tensor(tp["grp"][c0:c1].astype(np.int64), device=device)
    lmu = torch.as_tensor(tp["gmean"][g0:g1], device=device)[:, None] + torch.as_tensor(tp["Mb"][g0:g1], device=device)[:, grp] \
        + torch.as_tensor(tp["alpha"][g0:g1], device=device) @ Wt.T
    psi = torch.as_tensor(tp["psi"][g0:g1], device=device)[:, None]
    torch.manual_seed(seed)
    lam = torch._standard_gamma((1.0 / psi).expand_as(lmu).contiguous()) * torch.exp(lmu) * psi
    return torch.poisson(lam).to(torch.int32).contiguous()


def upload_generated(ctx, tp, G, c0, c1, device, seed, chunk=250):
    """The rank's cells of every gene, generated on the device chunk by chunk and permuted into the resident layout."""
    import torch
    for g0 in range(0, G, chunk):
        g1 = min(G, g0 + chunk)
        y = gen_block_gpu(tp, g0, g1, c0, c1, device, seed * 100003 + g0 * 64 + (c0 // BLOCK) % 64)
        ctx.upload_counts_rows(None, g0, device_ptr=y.data_ptr(), nrows=g1 - g0)
        del y
    ctx.upload_counts_done()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def _peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        d = json.load(open(pk))
        if "hbm_gbs" in d:
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def _maxr(x, dist, dev, world):
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sumr(x, dist, dev, world):
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def cpu_getres_sample(workload, kind=3, genes=96, seed=5):
    """The oracle's get.res (NumPy restatement of R/get.res.R:32-124) on ONE 5000-cell block and a bounded gene sample."""
    from oracle import getres
    G, N, k, m, n_ctl, frac = CELL_WORKLOADS[workload]
    tp = cells_truth(G, N, k, m, n_ctl, frac, seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    rows = np.sort(rng.choice(G, genes, replace=False))
    B = min(BLOCK, N)
    lmu = tp["gmean"][rows, None] + tp["Mb"][rows][:, tp["grp"][:B]] + tp["alpha"][rows] @ tp["W"][:B].T
    Y = rng.poisson(rng.gamma(1.0 / tp["psi"][rows, None], np.exp(lmu) * tp["psi"][rows, None])).astype(np.float64)
    out = dict(counts=Y, a=tp["alpha"][rows], W=tp["W"][:B], gmean=tp["gmean"][rows], Mb=tp["Mb"][rows][:, tp["grp"][:B]], psi=tp["psi"][rows])
    t0 = time.perf_counter()
    getres.get_res(out, kind, block_size=B)
    dt = time.perf_counter() - t0
    return genes * B / dt, dict(kind="port", cores=1, sample=f"{genes} genes x one {B}-cell block of {workload}, type {kind} "
                                f"(3 = percentile-adjusted counts), NumPy restatement of R/get.res.R:32-124 ({dt:.1f} s)")


# ------------------------------------------------------------------------------------------------------------------------
def run_getres(args, rank, world, local, dist, json_out, ClockSampler):
    import torch
    from ruviiinb_b200 import _lib, api, shard
    G, N, k, m, n_ctl, frac = CELL_WORKLOADS[args.workload]
    dev = torch.device("cuda", local)
    seed = 20261019 + 5
    tp = cells_truth(G, N, k, m, n_ctl, frac, seed)
    c0, c1 = api.cell_shard_range(N, world, rank, BLOCK)
    Nl = c1 - c0
    config = {"workload": f"get.res (R/get.res.R) on a fitted NB model, {G} genes x {N} cells, k={k}, Mb by replicate group (m={m}), "
                          f"blocks of {BLOCK} cells ({args.workload})",
              "l2": "inputs larger than L2 (int32 counts %.1f GB per pass)" % (4.0 * G * N / 1e9), "sharding": "cells (whole blocks)" if world > 1 else "none"}
    ctx = _lib.Context(local)
    if world > 1:
        ctx.comm_init(shard.broadcast_unique_id(dist, _lib.load(), device=dev), rank, world)
    ctx.set_design(G, Nl, tp["grp"][c0:c1], tp["ctl"], m=m)
    ctx.set_cell_shard(N, c0)
    t_gen = time.perf_counter()
    upload_generated(ctx, tp, G, c0, c1, dev, seed)
    t_gen = time.perf_counter() - t_gen
    for name, key in (("W", "W"), ("alpha", "alpha"), ("gmean", "gmean"), ("Mb", "Mb"), ("psi", "psi")):
        ctx.set_state(name, tp[key][c0:c1] if name == "W" else tp[key])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ctx.get_res_ex(3, block_size=BLOCK, out_kind=_lib.OUT_NONE)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    clocks.begin()
    dsec, stats = 0.0, None
    for _ in range(args.steps):
        _, stats = ctx.get_res_ex(3, block_size=BLOCK, out_kind=_lib.OUT_NONE)
        dsec += stats["device_seconds"]
    barrier()
    clocks.end()
    clk = clocks.stop() if rank == 0 else None
    dmax = _maxr(dsec, dist, dev, world)
    value = G * N * args.steps / dmax
    checksum = _sumr(stats["checksum"], dist, dev, world)
    other = {}
    for kind, nm in ((1, "pearson"), (2, "logcounts")):
        ctx.get_res_ex(kind, block_size=BLOCK, out_kind=_lib.OUT_NONE, cells=(0, min(Nl, 4 * BLOCK)))
        _, st = ctx.get_res_ex(kind, block_size=BLOCK, out_kind=_lib.OUT_NONE)
        other[nm] = {"elements_per_s": G * N / _maxr(st["device_seconds"], dist, dev, world), "caps_ms": st["caps_ms"], "elem_ms": st["elem_ms"],
                     "select_ms": st["select_ms"]}
    # ---- end to end: host buffers in, host buffers out, range by range through fresh contexts -----------------------------------
    e2e = None
    if not args.no_e2e:
        R = min(Nl, 5 * BLOCK)
        Yh = torch.empty((G, R), dtype=torch.int32, pin_memory=True)
        ycols = ctx.get_cells(np.arange(R))                                  # the first R cells of this rank, R layout
        Yh.copy_(torch.from_numpy(np.ascontiguousarray(ycols)))
        out_t = torch.empty((R, G), dtype=torch.int32, pin_memory=True)      # column-major G x R
        out_np = out_t.numpy().T
        nrng = max(1, Nl // R)
        barrier()
        t0 = time.perf_counter()
        for r in range(nrng):
            with _lib.Context(local) as c2:
                c2.set_design(G, R, tp["grp"][c0:c0 + R], tp["ctl"], m=m)
                c2.upload_counts(Yh.numpy())
                for name, key in (("W", "W"), ("alpha", "alpha"), ("gmean", "gmean"), ("Mb", "Mb"), ("psi", "psi")):
                    c2.set_state(name, tp[key][c0:c0 + R] if name == "W" else tp[key])
                c2.get_res_ex(3, block_size=BLOCK, out_kind=_lib.OUT_I32, out=out_np)
        barrier()
        dt = _maxr(time.perf_counter() - t0, dist, dev, world)
        e2e = {"value": G * R * nrng * world / dt, "unit": "elements/s", "h2d_bytes_per_step": int(Yh.numel() * 4 + tp["W"][:R].nbytes + 8 * G * (k + 2 + m)),
               "d2h_bytes_per_step": int(out_t.numel() * 4),
               "what": f"per step one range of {R} cells: fresh context, upload_counts (pinned host int32 -> HBM), set_state x5, get_res_ex(PAC, int32 "
                       f"sink into pinned host memory); {nrng} ranges per rank, wall clock, max over ranks"}
    ctx.close()
    if rank != 0:
        return None
    peak, peak_src = _peaks()
    nblk = -(-Nl // BLOCK)
    kern = {"k_res_caps": stats["caps_ms"] / nblk, "k_res_elem<PAC>": stats["elem_ms"] / nblk}
    bytes_elem = (4.0 + 4.0) * G * BLOCK          # int32 count in, int32 PAC out per element; parameters are L2-resident
    traffic, tsrc = None, None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        tr = json.load(open(tr_path))
        traffic = tr.get(args.workload, {}).get("k_res_elem")
        tsrc = tr.get("source_cells")
    ach = bytes_elem / (kern["k_res_elem<PAC>"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_res_elem<QUANTILE> (one 5000-cell block)", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_elem,
                "note": "FP64-bound (pmf recurrences of the mid-p and of qnbinom, 2 exp + 2 log per element); the other kernel of the pass, k_res_caps, "
                        "is a shared-memory selection (3 x median + MAD of 5000 values per gene and block) with no HBM stream of its own",
                "kernels_ms_per_block": kern}
    line = {"metric": "get.res percentile-adjusted counts, elements/sec", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dmax / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clk, "gpu_launches": int(stats["launches"] * args.steps),
            "roofline": roofline, "checksum_pac": checksum, "other_types": other, "datagen_seconds": t_gen}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        v, desc = cpu_getres_sample(args.workload)
        line["cpu_baseline"] = dict(value=v, unit="elements/s", **desc)
    return line


# ------------------------------------------------------------------------------------------------------------------------
def run_fast(args, rank, world, local, dist, json_out, ClockSampler):
    """fastruvIII.nb on resident counts: Winsorize over all cells, the subsample fit (<= 3000 cells, every rank), W for all cells,
    Mb.all + PAC in one pass per block, psi cap.  A step = the whole pipeline after the upload."""
    import torch
    from ruviiinb_b200 import _lib, api, shard
    G, N, k, m, n_ctl, frac = CELL_WORKLOADS[args.workload]
    dev = torch.device("cuda", local)
    seed = 20261019 + 4
    tp = cells_truth(G, N, k, m, n_ctl, frac, seed)
    c0, c1 = api.cell_shard_range(N, world, rank, BLOCK)
    Nl = c1 - c0
    annot_all = np.where(tp["annotated"], tp["grp"], -1).astype(np.int32)
    config = {"workload": f"fastruvIII.nb (R/fastruvIIInb.R), {G} genes x {N} cells, k={k}, {m} replicate groups, {int(100 * frac)} % of the cells annotated, "
                          f"subsample <= 3000 cells ({args.workload})",
              "l2": "inputs larger than L2 (int32 counts %.1f GB per pass)" % (4.0 * G * N / 1e9), "sharding": "cells (whole blocks)" if world > 1 else "none"}
    ctx = _lib.Context(local)
    if world > 1:
        ctx.comm_init(shard.broadcast_unique_id(dist, _lib.load(), device=dev), rank, world)
    ctx.set_design(G, Nl, np.where(annot_all[c0:c1] >= 0, annot_all[c0:c1], 0).astype(np.int32), tp["ctl"], m=m)
    ctx.set_cell_shard(N, c0)
    t_gen = time.perf_counter()
    upload_generated(ctx, tp, G, c0, c1, dev, seed)
    t_gen = time.perf_counter() - t_gen

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rng = np.random.default_rng(seed)
    p = min(0.2, 3000.0 / (annot_all >= 0).sum())
    subsamples = np.sort(np.concatenate([rng.choice(np.where(annot_all == j)[0], size=int(np.ceil(p * (annot_all == j).sum())), replace=False) for j in range(m)]))
    mine = (subsamples >= c0) & (subsamples < c1)
    ranks = api.Ranks(dist if world > 1 else None, device=dev)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    clocks.begin()
    stage = {}
    t0 = time.perf_counter()
    _, nz = ctx.winsorize()
    torch.cuda.synchronize(); stage["winsorize_s"] = time.perf_counter() - t0
    keep = nz > 0
    if not keep.all():
        ctx.set_row_mask(keep)
    t1 = time.perf_counter()
    cols = ctx.get_cells(subsamples[mine] - c0)
    parts = ranks.allgather((np.where(mine)[0], cols))
    Ysub = np.empty((G, subsamples.size), dtype=np.int32)
    for idx, cc in parts:
        Ysub[:, idx] = cc
    Msub = np.zeros((subsamples.size, m), dtype=bool)
    Msub[np.arange(subsamples.size), annot_all[subsamples]] = True
    sub = api.fastruvIII_nb_subfit(Ysub[keep], Msub, tp["ctl"][keep], k=k, device=local, seed=seed)
    stage["subsample_fit_s"] = time.perf_counter() - t1
    t2 = time.perf_counter()

    def full(x, fill=0.0):
        o = np.full((G,) + x.shape[1:], fill)
        o[keep] = x
        return o
    W_init = np.zeros((Nl, k))
    W_init[subsamples[mine] - c0] = sub["Wsub"][mine]
    med = np.median(sub["Wsub"], axis=0)
    mad = 1.4826 * np.median(np.abs(sub["Wsub"] - med[None, :]), axis=0)
    wt_full = np.zeros(int(tp["ctl"].sum()))
    wt_full[keep[tp["ctl"]]] = sub["wt_ctl"]
    ctx.set_state("W", W_init); ctx.set_state("alpha", full(sub["alpha"])); ctx.set_state("gmean", full(sub["gmean"]))
    ctx.set_state("psi", full(sub["psi"], 1.0)); ctx.set_state("Mb", full(sub["Mb"]))
    W, sweeps, crit = ctx.fast_W_all(W_init, wt_full, med, mad)
    torch.cuda.synchronize(); stage["W_all_s"] = time.perf_counter() - t2
    t3 = time.perf_counter()
    _, _, st = ctx.fast_res(3, annot_all[c0:c1], lambda_b=16.0, block_size=BLOCK, out_kind=_lib.OUT_NONE)
    psi_cap = _lib.fast_cap_psi(sub["psi"])
    torch.cuda.synchronize(); stage["Mb_all_PAC_s"] = time.perf_counter() - t3
    barrier()
    clocks.end()
    wall = _maxr(time.perf_counter() - t0, dist, dev, world)
    clk = clocks.stop() if rank == 0 else None
    stage = {k_: _maxr(v, dist, dev, world) for k_, v in stage.items()}
    dres = _maxr(st["device_seconds"], dist, dev, world)
    checksum = _sumr(st["checksum"], dist, dev, world)
    corr = None
    ctx.close()
    if rank != 0:
        return None
    # how well W tracks the truth on this rank's cells (canonical correlations) -- a sanity figure, not a parity claim
    q1, q2 = np.linalg.qr(W - W.mean(0))[0], np.linalg.qr(tp["W"][c0:c1] - tp["W"][c0:c1].mean(0))[0]
    corr = [float(x) for x in np.linalg.svd(q1.T @ q2, compute_uv=False)]
    peak, peak_src = _peaks()
    nblk = -(-Nl // BLOCK)
    bytes_blk = (4.0 + 8.0 + 8.0 + 4.0) * G * BLOCK     # count in; Mb.all block written and read back on the device; int32 PAC out
    ach = bytes_blk / ((st["mb_ms"] + st["elem_ms"]) / nblk * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_fast_mb + k_res_elem<QUANTILE> (one 5000-cell block)", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_blk,
                "note": "FP64-bound elementwise passes (6 exp per element in the three Mb.all sweeps, the PAC recurrences); k_res_caps is a shared-memory selection",
                "kernels_ms_per_block": {"k_fast_mb": st["mb_ms"] / nblk, "k_res_caps": st["caps_ms"] / nblk, "k_res_elem<PAC>": st["elem_ms"] / nblk}}
    full_pass = stage["winsorize_s"] + stage["W_all_s"] + stage["Mb_all_PAC_s"]
    line = {"metric": "fastruvIII.nb full-data passes (Winsorize + W for all cells + Mb.all + PAC), elements/sec", "value": G * N / full_pass, "unit": "elements/s",
            "n_gpus": world, "steps": 1, "warmup": 0, "ms_per_step": 1e3 * wall, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clk, "gpu_launches": int(st["launches"]), "roofline": roofline,
            "wall_seconds_after_upload": wall, "stages": stage, "W_sweeps": sweeps, "W_crit": crit, "fast_res_device_seconds": dres, "checksum_pac": checksum,
            "subsample_cells": int(subsamples.size), "dropped_genes": int((~keep).sum()), "canonical_corr_W_truth_rank0": corr, "datagen_seconds": t_gen,
            "e2e": {"value": G * N / wall, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(Ysub.nbytes + W.nbytes),
                    "what": "the whole pipeline after the (device-generated) counts are resident: Winsorize, subsample gather + fit, W for all cells, "
                            "Mb.all + PAC with the checksum sink, psi cap; wall clock, max over ranks.  The count matrix itself (120 GB) is generated on the "
                            "device: no host of this pool holds it"}}
    if not args.no_cpu_baseline:
        v, desc = cpu_getres_sample(args.workload)
        line["cpu_baseline"] = dict(value=v, unit="elements/s", **desc)
    return line


# ------------------------------------------------------------------------------------------------------------------------
FIT_WORKLOADS = {
    # name: G, N, k, m, n_ctl, zinb
    "cfg3_zinb_10kx50k": (10000, 50000, 3, 20, 1000, True),
    "cfg1_standin_12kx9k": (12000, 9027, 2, 2, 1085, False),
}


def run_fit(args, rank, world, local, dist, json_out, ClockSampler):
    """BASELINE configs 3 (ZINB, CSR upload) and 1 (the stand-in for the bundled cellline data): K inner IRLS iterations timed on the
    device as on the headline line, plus the whole ruvIII.nb fit through the public mirror (wall clock).  One GPU."""
    import torch
    import bench
    from ruviiinb_b200 import _lib, api
    if world > 1:
        raise SystemExit("these workloads are single-GPU lines (the headline workload carries the gene-sharded scaling)")
    G, N, k, m, n_ctl, zinb = FIT_WORKLOADS[args.workload]
    dev = torch.device("cuda", local)
    seed = bench.SEED + (1 if zinb else -1)
    tp = bench.truth_params(G, N, k, m, n_ctl, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    if zinb:
        tp["gmean"] = np.clip(-2.5 + rng.standard_normal(G), -4.0, 5.0)
    if m == 2:     # the cellline stand-in: group sizes 4330 / 4697
        grp = np.zeros(N, dtype=np.int32); grp[rng.permutation(N)[:4697]] = 1
        tp["grp"] = grp
    Y = np.empty((G, N), dtype=np.int32)
    bench.gen_rows_gpu(tp, np.arange(G), Y, dev)
    zi = None
    if zinb:
        pi0 = rng.beta(2.0, 5.0, size=G)
        Y[rng.random(Y.shape, dtype=np.float32) < pi0[:, None].astype(np.float32)] = 0
        zi = np.ones(N, bool)
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    opts = _lib.default_opts(k=k, zinb=int(zinb))
    config = {"workload": f"ruvIII.nb {'ZINB' if zinb else 'NB'} inner IRLS iteration, k={k}, {G} genes x {N} cells, {m} replicate groups, {n_ctl} control genes "
                          f"({args.workload}{', %.0f %% zeros, CSR upload' % (100 * (Y == 0).mean()) if zinb else ''})",
              "l2": "inputs larger than L2 (int32 counts %.1f GB per pass)" % (4.0 * G * N / 1e9), "sharding": "none"}
    ctx = _lib.Context(local)
    ctx.set_design(G, N, tp["grp"], tp["ctl"], zeroinf=zi)
    if zinb:
        import scipy.sparse as sp
        csr = sp.csr_matrix(Y)
        t0 = time.perf_counter(); ctx.upload_counts_csr(csr.indptr, csr.indices, csr.data); t_up = time.perf_counter() - t0
    else:
        t0 = time.perf_counter(); ctx.upload_counts(Y); t_up = time.perf_counter() - t0
    ctx.set_state("W", tp["W0"]); ctx.init_coef(opts); ctx.set_state("psi", tp["psi"])
    if zinb:
        ctx.zinb_pi0(opts)
    ctx.loglik(opts)
    for _ in range(args.warmup):
        ctx.irls_iteration(opts); ctx.accept()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.kernel_time_reset()
    l0 = ctx.kernel_launches
    clocks = ClockSampler(local); clocks.start()
    torch.cuda.synchronize(); clocks.begin()
    e0.record(stream)
    logls = []
    for _ in range(args.steps):
        ll, bad = ctx.irls_iteration(opts); ctx.accept(); logls.append(ll)
    e1.record(stream)
    torch.cuda.synchronize(); clocks.end()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - l0
    kt = ctx.kernel_times()
    ctx.close()
    # the whole fit through the public mirror: Winsorize, device SVD start, robust estimateDisp at every psi refresh, alternating IRLS
    M = np.zeros((N, m), bool); M[np.arange(N), tp["grp"]] = True
    t0 = time.perf_counter()
    fit = api.ruvIII_nb(Y, M, tp["ctl"], k=k, zeroinf=zi, device=local, return_trace=True)
    wall = time.perf_counter() - t0
    peak, peak_src = _peaks()
    kern = {n: {"ms": tot / cnt, "launches_per_step": cnt / args.steps} for n, (tot, cnt) in kt.items() if cnt}
    dom = max(kern, key=lambda n: kern[n]["ms"] * kern[n]["launches_per_step"])
    ach = 4.0 * G * N / (kern[dom]["ms"] * 1e-3) / 1e9
    line = {"metric": "NB IRLS gene-fits/sec", "value": G * args.steps / (ms * 1e-3), "unit": "gene-fits/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clk, "gpu_launches": int(launches), "logl_trace": logls,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 4.0 * G * N, "note": "FP64-pipe-bound sweep (DESIGN.md section 5)", "kernels": kern},
            "e2e": {"value": G * fit["stats"]["n_inner"] / wall, "unit": "gene-fits/s", "h2d_bytes_per_step": int(Y.nbytes / max(1, fit["stats"]["n_inner"])),
                    "d2h_bytes_per_step": int((fit["W"].nbytes + fit["a"].nbytes + fit["Mb"].nbytes + Y.nbytes) / max(1, fit["stats"]["n_inner"])),
                    "what": "the whole ruvIII.nb through api.ruvIII_nb with host buffers: upload, Winsorize, zero-gene filter, device SVD start, estimateDisp "
                            "(robust) at every psi refresh, all inner iterations, read-back of counts and parameters; gene-fits = G x inner iterations / wall"},
            "full_fit": {"wall_seconds": wall, "outer_iterations": fit["stats"]["n_outer"], "inner_iterations": fit["stats"]["n_inner"],
                         "inner_loop_seconds": fit["stats"]["inner_seconds"], "dispersion_seconds": fit["stats"]["disp_seconds"], "upload_seconds_bench_ctx": t_up,
                         "logl_outer": [float(x) for x in fit["logl"]], "dropped_genes": int(fit["zero_genes"].sum())}}
    if not args.no_cpu_baseline:
        v, desc, _ = bench.cpu_sample_run("cfg2_nb_20kx100k", 2, 1)
        desc["sample"] += " (the NB inner iteration of the headline workload; the C restatement has no ZINB branch)" if zinb else ""
        line["cpu_baseline"] = dict(value=v, unit="gene-fits/s", **desc)
    return line


def run_reference(args):
    v, desc = cpu_getres_sample(args.workload)
    G, N, k, m, n_ctl, frac = CELL_WORKLOADS[args.workload]
    return {"impl": "reference", "metric": "get.res percentile-adjusted counts, elements/sec", "value": v, "unit": "elements/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": args.workload}, "cpu_baseline": dict(value=v, unit="elements/s", **desc),
            "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
