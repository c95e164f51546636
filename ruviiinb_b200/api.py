"""Host-side mirror of the reference's exported R functions (NAMESPACE:3-6) over the C ABI of libruvnb_b200.

    ruvIII_nb(Y, M, ctl, k=2, ...)      R/ruvIIInb.R:27-629     -> dict with the reference's list elements
    get_res(out, type=..., ...)         R/get.res.R:12-129      -> G x N array
    makeSCE(obj, cData, batch, assays)  R/makeSCE.R:12-43       -> dict standing in for the SingleCellExperiment
    estimateDisp_par(y, offset=...)     R/estimateDisp.par.R:3-202 -> dict with the reference's list elements
    fastruvIII_nb_finish(...)           R/fastruvIIInb.R:625-792 (the full-data passes after the subsample fit)

Same argument names (dots become underscores), defaults and error behaviour: invalid input raises ValueError where the R
code calls stop(); numerical "problematic entries" are reverted inside the library, never raised.  `ncores` is accepted and
ignored (the fork pool is replaced by the GPU).  Everything numerical happens in the CUDA library: there is no CPU
fallback, and this module never imports the test oracle.
"""
import numpy as np

from . import _lib


def _is_sparse(Y):
    """A scipy.sparse matrix: the stand-in for R's dgCMatrix (counts of ruvIII.nb's output list, R/ruvIIInb.R:611-612)."""
    return hasattr(Y, "tocsc") and hasattr(Y, "nnz")


def _upload_any(ctx, Y, rows=None):
    """Counts into the context: a dense array goes up as int32 rows, a sparse matrix as the slots of its CSC form (the device
    transposes it into the gene-major resident layout, include/ruvnb.h ruvnb_upload_counts_csc)."""
    if _is_sparse(Y):
        Yc = Y.tocsr()[rows].tocsc() if rows is not None else Y.tocsc()
        Yc.sort_indices()
        ctx.upload_counts_csc(Yc.indptr, Yc.indices, Yc.data, block_cols=max(1, (1 << 27) // max(1, Yc.shape[0])))
    else:
        Yd = np.asarray(Y)
        ctx.upload_counts(np.ascontiguousarray(np.ceil(Yd[rows] if rows is not None else Yd), dtype=np.int32))

_UNSUPPORTED = "is not part of the B200 path (SURVEY.md section 2: out of scope, off by default in the reference)"


def _as_logical_ctl(ctl, G, rownames=None):
    """R/ruvIIInb.R:42-45: a logical vector, or indices / names of the control genes."""
    ctl = np.asarray(ctl)
    if ctl.dtype == bool:
        if ctl.size != G:
            raise ValueError("ctl must have one entry per gene")
        return ctl.copy()
    out = np.zeros(G, dtype=bool)
    if ctl.dtype.kind in "iu":
        out[ctl] = True      # 0-based positions (R: 1-based)
        return out
    if rownames is None:
        raise ValueError("character ctl needs rownames")
    out[np.isin(np.asarray(rownames), ctl)] = True
    return out


def init_W(Y, ctl, k):
    """Host check of `ruvnb_init_W` (R/ruvIIInb.R:160-168): the k leading right singular vectors of
    log(Y[ctl,]/rowMeans(Y[ctl,]) + 1) by a dense eigen-decomposition of the n_ctl x n_ctl Gram, each vector signed so that
    its largest entry is positive.  The fit itself uses the device routine (`Context.init_W`); this is kept for tests and
    for callers that want to inject the same W0 on both sides of a comparison."""
    Yc = np.asarray(Y)[np.asarray(ctl, bool)].astype(np.float64)
    A = np.log(Yc / Yc.mean(axis=1, keepdims=True) + 1.0)
    w, U = np.linalg.eigh(A @ A.T)
    order = np.argsort(w)[::-1][:k]
    V = (A.T @ U[:, order]) / np.sqrt(w[order])[None, :]
    for j in range(k):
        if V[np.argmax(np.abs(V[:, j])), j] < 0:
            V[:, j] = -V[:, j]
    return V


def _quantcut(x, q):
    """as.numeric(gtools::quantcut(x, q = q)) (R/ruvIIInb.R:97,120): bin 1..q of every value, breaks at the type-7 quantiles
    seq(0, 1, length = q + 1), intervals closed on the right, the lowest value included; tied breaks are merged."""
    x = np.asarray(x, dtype=np.float64)
    br = np.unique(np.quantile(x, np.linspace(0.0, 1.0, int(q) + 1)))
    if br.size < 2:
        return np.ones(x.size, dtype=np.int64)
    return np.clip(np.searchsorted(br, x, side="left"), 1, br.size - 1).astype(np.int64)


def add_pseudocells(Y, M, batch, nc_pool=20, batch_disp=False, rng=None):
    """The pseudo-cells of `use.pseudosample = TRUE` (R/ruvIIInb.R:83-144) for a (winsorized) G x N count matrix: inside every
    replicate group (stratum) and batch the cells are binned by library size (gtools::quantcut of sf = colSums(Y)/mean, :54-55,
    into max(round(n / nc.pool), 1) bins) and every pool adds one cell drawn as rbinom(size = rowSums(pool), prob = 1/ncol(pool))
    (:122-123); the pseudo-cells of a stratum form a replicate group of their own (a new column of M, :130-135) and get batch
    btc + max(batch) -- or 2 against 1 for the real cells when batch.disp = FALSE (:138-141).  R's RNG stream cannot be
    reproduced: `rng` is a numpy Generator (default: seed 0).  Returns (Y_ext, M_ext, batch_ext)."""
    rng = np.random.default_rng(0) if rng is None else rng
    Y = np.asarray(Y)
    M = np.asarray(M).astype(bool)
    batch = np.asarray(batch).astype(np.int64)
    G, nsc = Y.shape
    strata = np.argmax(M, axis=1)                        # apply(M, 1, which), :50
    sf = Y.sum(axis=0).astype(np.float64)
    sf = sf / sf.mean()                                  # :54-55
    bmax = int(batch.max())
    new_cols, new_batch = [], []
    Mx = M
    for st in list(dict.fromkeys(strata.tolist())):      # unique(na.omit(strata)): order of first appearance
        added = 0
        for btc in list(dict.fromkeys(batch.tolist())):  # unique(batch)
            sel = np.where((batch == btc) & (strata == st))[0]
            if sel.size > 1:                             # :119
                nq = max(int(np.round(sel.size / nc_pool)), 1)   # R's round(): half to even, like numpy
                qsf = _quantcut(sf[sel], nq)
                for qv in range(1, int(qsf.max()) + 1):
                    pool = Y[:, sel[qsf == qv]]
                    if pool.shape[1] == 0:
                        continue
                    new_cols.append(rng.binomial(pool.sum(axis=1).astype(np.int64), 1.0 / pool.shape[1]))
                    new_batch.append(btc + bmax)
                    added += 1
        if (strata == st).sum() > 1 and added:           # :130-135: a new replicate-group column for this stratum's pseudo-cells
            rows = Mx.shape[0]
            Mx = np.vstack([Mx, np.zeros((added, Mx.shape[1]), dtype=bool)])
            col = np.zeros(Mx.shape[0], dtype=bool)
            col[rows:] = True
            Mx = np.hstack([Mx, col[:, None]])
    if not new_cols:
        return Y, M, batch
    Yx = np.concatenate([Y, np.stack(new_cols, axis=1).astype(Y.dtype)], axis=1)
    bx = np.concatenate([batch, np.asarray(new_batch, dtype=np.int64)])
    if not batch_disp:
        bx = np.where(bx > bmax, 2, 1)                   # :138-139
    return Yx, Mx, bx


def ruvIII_nb(Y, M, ctl, k=2, robust=False, ortho_W=False, lambda_a=0.01, lambda_b=16, batch=None, step_fac=0.5,
              inner_maxit=50, outer_maxit=25, ncores=2, use_pseudosample=False, nc_pool=20, batch_disp=False,
              zeroinf=None, W0=None, psi=None, device=0, rownames=None, return_trace=False, rng=None, pseudo=None):
    """`ruvIII.nb` (R/ruvIIInb.R:27-629).  Y: G x N counts; M: N x m replicate matrix with exactly one TRUE per row
    (:206); ctl: logical / index vector of control genes.  Extra arguments of the mirror: W0 (injects the initial W
    instead of the SVD; with pseudo-cells it has one row per cell INCLUDING them), psi (injects a dispersion vector or schedule
    instead of estimateDisp; G x nbatch per entry for a batch-specific fit), rng (numpy Generator for the pseudo-cell draws) or
    pseudo = (Y_ext, M_ext, batch_ext) (injects the pseudo-cells themselves: they depend on R's RNG in the reference), device.
    Returns the reference's list as a dict: counts, W, M, ctl, logl, a, Mb, gmean, pi0, psi, L.a, L.b, batch."""
    if batch is None and batch_disp:            # :65-68
        batch_disp = False
    if batch is None and use_pseudosample:      # :71-74
        use_pseudosample = False
    if ortho_W:
        raise NotImplementedError("ortho.W " + _UNSUPPORTED)
    sparse_in = _is_sparse(Y)
    if not sparse_in:
        Y = np.asarray(Y)
    if Y.ndim != 2:
        raise ValueError("Y must be a genes x cells matrix")
    G, N = Y.shape
    M = np.asarray(M).astype(bool)                                   # :46-49
    if M.shape[0] != N:
        raise ValueError("M must have one row per cell")
    if (M.sum(axis=1) != 1).any():
        raise ValueError("every cell must belong to exactly one replicate group (R/ruvIIInb.R:206 indexes Mb[, apply(M, 1, which)])")
    if (M.sum(axis=0) == 0).any():
        raise ValueError("Each column of M matrix must have at least one non-zero entry")   # R/fastruvIIInb.R:60
    ctl = _as_logical_ctl(ctl, G, rownames)
    batch = np.ones(N, dtype=np.int64) if batch is None else np.asarray(batch)              # :77-80
    if batch.shape[0] != N:
        raise ValueError("batch must have one entry per cell")
    zi = np.zeros(N, dtype=bool) if zeroinf is None else np.asarray(zeroinf).astype(bool)   # :39-40
    if use_pseudosample:                        # :147-152: only NB models with pseudo-samples
        zi = np.zeros(N, dtype=bool)
    # :53-62: Winsorize + ceiling on the device, THEN drop the genes that are all zero (a gene expressed in fewer than 0.5 % of
    # the cells has a 99.5 % quantile of 0: the cap zeroes the whole row, and the reference removes it only after the cap).
    # Genes that are all zero in the raw data are certainly all zero after it: they never travel.
    grp = np.argmax(M, axis=1).astype(np.int32)                      # apply(M, 1, which), :50
    opts = _lib.default_opts(k=int(k), robust=int(bool(robust)), lambda_a=float(lambda_a), lambda_b=float(lambda_b),
                             step_fac=float(step_fac), inner_maxit=int(inner_maxit), outer_maxit=int(outer_maxit),
                             zinb=int(zi.any()))
    keep = np.asarray((Y > 0).sum(axis=1)).ravel() > 0
    nb_fit = len(np.unique(batch)) if (batch_disp or use_pseudosample) else 1    # :88-91,145: batch labels only matter for psi
    ctx = _lib.Context(device)
    bctx = []
    try:
        if nb_fit > 1 and not use_pseudosample:
            ctx.set_batches(batch)
        ctx.set_design(int(keep.sum()), N, grp, ctl[keep], zeroinf=zi if zi.any() else None)
        _upload_any(ctx, Y, rows=np.where(keep)[0])
        _, nz = ctx.winsorize()
        counts = ctx.get_counts()
        M_fit, batch_fit, N_fit, counts_fit = M, batch, N, counts
        if (nz == 0).any() or use_pseudosample:
            # rows emptied by the cap leave (:57-62), pseudo-cells join (:83-144): the fit context holds that matrix
            ctx.close()
            alive = nz > 0
            keep[np.where(keep)[0][~alive]] = False
            counts = counts_fit = np.ascontiguousarray(counts[alive])
            if use_pseudosample:
                counts_fit, M_fit, batch_fit = pseudo if pseudo is not None else add_pseudocells(counts, M, batch, nc_pool, batch_disp, rng)
                counts_fit = np.ascontiguousarray(counts_fit, dtype=np.int32)
                N_fit = counts_fit.shape[1]
                nb_fit = len(np.unique(batch_fit))
                grp = np.argmax(M_fit, axis=1).astype(np.int32)
                zi = np.zeros(N_fit, dtype=bool)
            ctx = _lib.Context(device)
            if nb_fit > 1:
                ctx.set_batches(batch_fit)
            ctx.set_design(counts_fit.shape[0], N_fit, grp, ctl[keep], zeroinf=zi if zi.any() else None)
            ctx.upload_counts(counts_fit)   # already capped (the cap of a capped row is the same cap)
        ctlk = ctl[keep]
        if ctlk.sum() == 0:
            raise ValueError("no control gene is left after removing all-zero genes")
        if W0 is None:
            W0, _, _ = ctx.init_W(int(k))        # H1 on the device (subspace iteration on the resident control rows)
        psi_inject = None
        if psi is not None:
            psi_inject = np.asarray(psi, dtype=np.float64)
            psi_inject = (psi_inject.reshape(-1, G, nb_fit) if nb_fit > 1 else np.atleast_2d(psi_inject))[:, keep]
        elif nb_fit > 1:
            # one estimateDisp.par per batch on that batch's cells (:213-220,542-548): a context per batch, refreshed through the
            # fit's dispersion hook
            b0 = np.unique(batch_fit, return_inverse=True)[1]
            for B in range(nb_fit):
                idx = np.where(b0 == B)[0]
                present, gc = np.unique(grp[idx], return_inverse=True)     # replicate groups that occur in this batch
                c2 = _lib.Context(device)
                bctx.append((c2, idx, present))
                c2.set_design(counts_fit.shape[0], idx.size, gc.astype(np.int32), ctlk, zeroinf=zi[idx] if zi.any() else None)
                c2.upload_counts(np.ascontiguousarray(counts_fit[:, idx]))

            def refresh(c):
                W_, a_, Mb_, old = c.get_state("W"), c.get_state("alpha"), c.get_state("Mb"), c.get_state("psi")
                new = old.copy()
                for B, (c2, idx, present) in enumerate(bctx):
                    c2.set_state("W", W_[idx])
                    c2.set_state("alpha", a_)
                    c2.set_state("Mb", Mb_[:, present])
                    if zi.any():
                        c2.set_state("pi0", c.get_state("pi0"))
                        c2.set_state("psi_w", c.get_state("psi_w")[:, B])
                    try:
                        c2.estimate_disp(opts)
                        col = c2.get_state("psi")
                    except _lib.RuvnbError:                                # tryCatch(..., error = function(e) NA), :545
                        col = np.full(old.shape[0], np.nan)
                    ok = np.isfinite(col)                                  # :564
                    new[ok, B] = col[ok]
                c.set_state("psi", new)

            ctx.set_psi_callback(refresh)
        logl, recs, stats = ctx.fit_nb(opts, W0=np.asarray(W0, dtype=np.float64), psi_inject=psi_inject)
        psi_out, W_out = ctx.get_state("psi"), ctx.get_state("W")
        if use_pseudosample:                                               # :618-624: drop what belongs to the pseudo-cells
            nb_org = len(np.unique(batch)) if batch_disp else 1
            W_out = W_out[:N]
            M_fit = M_fit[:N]
            psi_out = psi_out[:, :nb_org] if psi_out.ndim == 2 else psi_out
            if psi_out.ndim == 2 and psi_out.shape[1] == 1:
                psi_out = psi_out[:, 0]
        out = {"counts": counts, "W": W_out, "M": M_fit, "ctl": ctlk, "logl": logl, "a": ctx.get_state("alpha"),
               "Mb": ctx.get_state("Mb"), "gmean": ctx.get_state("gmean"), "psi": psi_out,
               "L.a": lambda_a, "L.b": lambda_b, "batch": batch}
        pi0 = ctx.get_state("pi0")
    finally:
        ctx.close()
        for c2, _, _ in bctx:
            c2.close()
    out["pi0"] = np.outer(pi0, zi[:N].astype(np.float64)) if zi.any() else None   # :611-626
    if sparse_in:                                                              # :611-612: counts leave as a sparse matrix
        import scipy.sparse as sp
        out["counts"] = sp.csc_matrix(out["counts"])
    out["zero_genes"] = ~keep
    if return_trace:
        out["trace"], out["stats"] = recs, stats
    return out


def get_res(out, type=("logcounts", "pearson", "quantile"), batch=None, block_size=5000, device=0, ranks=None, N_total=None, c0=0,
            sink="f64", return_stats=False, sink_file=None):
    """`get.res` (R/get.res.R:12-129).  `type`: like the reference, every requested kind is evaluated in the order
    pearson, logcounts, quantile and the LAST one is returned (:49,68,75), so the default returns the percentile-adjusted
    counts.  out["Mb"] may be G x N (fastruvIII.nb's Mb.all) or G x m (a ruvIII.nb fit: expanded by replicate group --
    the reference indexes Mb by cell, :35, and would need the caller to expand it).

    Cell shards (`ranks`, `N_total`, `c0` as in `fastruvIII_nb`): blocks are independent (:27-126) except for colMeans(W)
    (:36), which the library sums over the ranks; `out` then holds this rank's cells.  sink: "f64" (the reference's doubles),
    "i32" (PAC as int32), "log1p" (log(PAC + 1)), "none" (checksum only; with return_stats), "csc" / "csc_log1p" (a
    scipy.sparse.csc_matrix of the values / of log(PAC + 1): every block is compacted on the device and only its non-zero entries
    cross PCIe -- the `as(..., "sparseMatrix")` of R/makeSCE.R:24-38).  sink_file: the dense result is realised block by block into a
    file-backed array (.npy, column-major) instead of host memory -- the role of the reference's HDF5 AutoRealizationSink
    (R/get.res.R:21-23,125; HDF5 itself is not available in this environment) -- and returned as a numpy.memmap.
    out["counts"] may be dense or a scipy.sparse matrix (the dgCMatrix ruvIII.nb returns).
    Zero-inflated fits (out["pi0"]: the G x N matrix of R/ruvIIInb.R:611-626) take the ZIM branches (:52-57,69-70,84-85,95-96)."""
    kinds = [type] if isinstance(type, str) else list(type)
    for t in kinds:
        if t not in ("pearson", "logcounts", "quantile"):
            raise ValueError(f"unknown type '{t}'")
    last = [t for t in ("pearson", "logcounts", "quantile") if t in kinds][-1]
    psi_m = np.asarray(out["psi"], dtype=np.float64)
    bpsi = psi_m.ndim == 2 and psi_m.shape[1] > 1          # a batch-specific fit: out$psi[, curr.batch] (:59-62,101-121)
    if psi_m.ndim == 2 and not bpsi:
        psi_m = psi_m[:, 0]
    if bpsi:
        if batch is None:
            batch = out.get("batch")
        if batch is None:
            raise ValueError("a batch-specific dispersion needs the batch of every cell")
        bu = np.unique(batch)
        if bu.min() < 1 or bu.max() > psi_m.shape[1] or (bu != bu.astype(np.int64)).any():
            raise ValueError("batch must take values in 1..ncol(psi): it indexes the columns of psi (R/get.res.R:60)")
    ranks = ranks or Ranks()
    Y = out["counts"] if _is_sparse(out["counts"]) else np.ascontiguousarray(out["counts"], dtype=np.int32)
    G, N = Y.shape
    N_total = N if N_total is None else int(N_total)
    M = np.asarray(out["M"]).astype(bool)
    Mb = np.asarray(out["Mb"], dtype=np.float64)
    per_cell = Mb.shape[1] == N and Mb.shape[1] != M.shape[1]
    grp = np.argmax(M, axis=1).astype(np.int32) if (M.sum(axis=1) == 1).all() else np.zeros(N, dtype=np.int32)
    m = max(int(g.max()) for g in ranks.allgather(grp)) + 1
    kind = {"f64": _lib.OUT_F64, "i32": _lib.OUT_I32, "log1p": _lib.OUT_LOG1P, "none": _lib.OUT_NONE, "csc": _lib.OUT_F64,
            "csc_log1p": _lib.OUT_LOG1P}[sink]
    with _lib.Context(device) as ctx:
        if ranks.world > 1:
            ctx.comm_init(ranks.unique_id(), ranks.rank, ranks.world)
        if bpsi:   # the labels index the columns of psi as they are (this rank's cells may not meet every batch)
            bl = np.asarray(batch).astype(np.int64)
            ctx.set_batches((bl[c0:c0 + N] if len(bl) != N else bl) - 1, nbatch=psi_m.shape[1])
        ctx.set_design(G, N, grp, np.asarray(out["ctl"], bool), m=m)
        if ranks.world > 1 or N_total != N:
            ctx.set_cell_shard(N_total, c0)
        _upload_any(ctx, Y)
        ctx.set_state("W", out["W"])
        ctx.set_state("alpha", out["a"])
        ctx.set_state("gmean", out["gmean"])
        ctx.set_state("psi", psi_m)
        Mbg = np.zeros((G, m))
        if not per_cell:
            Mbg[:, :Mb.shape[1]] = Mb
        ctx.set_state("Mb", Mbg)
        if out.get("pi0") is not None:       # a ZINB fit: out$pi0 = outer(pi0, zeroinf), read as out$pi0[, batch[c]] (R/get.res.R:53,84-85)
            P = np.asarray(out["pi0"], dtype=np.float64)
            if P.shape[1] != N_total:
                raise ValueError("get.res of a zero-inflated fit needs the whole G x N pi0 matrix (its columns are indexed by batch)")
            zi = (P != 0).any(axis=0)
            b = np.ones(N_total, dtype=np.int64) if batch is None else np.asarray(batch).astype(np.int64)
            ctx.set_state("pi0", P.max(axis=1))
            # the quantile call's omega: rowMeans(out$pi0) (:95), or column ref.batch of out$pi0 with a batch-specific psi (:112-114)
            ctx.set_res_zinb(zi[b[c0:c0 + N] - 1], float(zi[int(np.argmin(psi_m.mean(axis=0)))]) if bpsi else float(zi.mean()))
        if sink.startswith("csc"):
            import scipy.sparse as sp
            cp, ri, vv, st = ctx.get_res_csc(_lib.RES_KINDS[last], block_size=block_size, Mb_cells=Mb if per_cell else None,
                                             log1p=sink == "csc_log1p")
            res = sp.csc_matrix((vv, ri, cp), shape=(getattr(ctx, "Gout", None) or G, N))
        else:
            dst = None
            if sink_file is not None and kind != _lib.OUT_NONE:
                dst = np.lib.format.open_memmap(sink_file, mode="w+", dtype=np.int32 if kind == _lib.OUT_I32 else np.float64,
                                                shape=(G, N), fortran_order=True)
            res, st = ctx.get_res_ex(_lib.RES_KINDS[last], block_size=block_size, Mb_cells=Mb if per_cell else None, out_kind=kind, out=dst)
            if dst is not None:
                dst.flush()
    return (res, st) if return_stats else res


def makeSCE(obj, cData=None, batch=None, assays=("pearson", "logPAC"), device=0):
    """`makeSCE` (R/makeSCE.R:12-43).  There is no SingleCellExperiment class outside R: the return value is a dict with the same
    parts -- assays (counts, and per request logcounts / pearson / logPAC, each a scipy.sparse.csc_matrix as the reference stores
    `as(..., "sparseMatrix")`), colData (= cData), rowData (alpha1..alphak and psi, :20-21).  Same argument checks and messages."""
    import scipy.sparse as sp
    batch = obj.get("batch")                                                        # :14
    if cData is None:
        raise ValueError("'cData' must be a user-specified data frame.")           # :15-16
    want = [assays] if isinstance(assays, str) else list(assays)
    if any(a not in ("logcounts", "logPAC", "pearson") for a in want):
        raise ValueError("'Wrong names of corrected assays data was specified. Available corrected data assays are logcounts,logPAC and pearson")
    a = np.asarray(obj["a"])
    row_data = {f"alpha{j + 1}": a[:, j] for j in range(a.shape[1])}
    row_data["obj.psi"] = np.asarray(obj["psi"])
    counts = obj["counts"] if _is_sparse(obj["counts"]) else sp.csc_matrix(np.asarray(obj["counts"]))
    sce = {"assays": {"counts": counts}, "colData": cData, "rowData": row_data}
    if "logcounts" in want:
        sce["assays"]["logcounts"] = get_res(obj, type="logcounts", batch=batch, sink="csc", device=device)
    if "pearson" in want:
        sce["assays"]["pearson"] = get_res(obj, type="pearson", batch=batch, sink="csc", device=device)
    if "logPAC" in want:
        sce["assays"]["logPAC"] = get_res(obj, type="quantile", batch=batch, sink="csc_log1p", device=device)   # :36-38
    return sce


def _factor_offset(offset, grp, m, tol=1e-9):
    """offset (G x N) = Mb[:, grp] + alpha W' exactly, with k <= RUVNB_MAX_K: the form every offset of the reference has
    (alpha W' + Mb[, grp], R/ruvIIInb.R:205-221,535-565) and the only one the device holds.  Group means go to Mb, the
    rest must have numerical rank <= 8."""
    O = np.asarray(offset, dtype=np.float64)
    G, N = O.shape
    cnt = np.bincount(grp, minlength=m).astype(np.float64)
    Mb = np.zeros((G, m))
    for j in range(m):
        Mb[:, j] = O[:, grp == j].sum(axis=1) / cnt[j]
    R = O - Mb[:, grp]
    scale = max(1.0, float(np.abs(O).max()))
    if np.abs(R).max() <= tol * scale:
        return Mb, np.zeros((G, 1)), np.zeros((N, 1))
    U, sv, Vt = np.linalg.svd(R, full_matrices=False)
    for k in range(1, min(_lib.MAX_K, len(sv)) + 1):
        alpha, W = U[:, :k] * sv[:k], Vt[:k].T
        if np.abs(R - alpha @ W.T).max() <= tol * scale:
            return Mb, alpha, W
    raise ValueError(f"offset is not of the form alpha W' + Mb[, group] with at most {_lib.MAX_K} factors "
                     "(the only offsets the reference passes, R/ruvIIInb.R:205-221); pass replicate_of_cell if the "
                     "offset has a replicate-group part")


def estimateDisp_par(y, design=None, group=None, lib_size=None, offset=None, prior_df=None, trend_method="locfit",
                     tagwise=True, span=None, min_row_sum=5, grid_length=21, grid_range=(-10, 10), robust=False,
                     winsor_tail_p=(0.05, 0.1), tol=1e-06, weights=None, BPPARAM=None, replicate_of_cell=None, device=0):
    """estimateDisp.par (R/estimateDisp.par.R:3-202) as ruvIII.nb calls it: design = 1 (one coefficient), an offset matrix
    of the form alpha W' + Mb[, grp], the 21-point grid 0.1 * 2^seq(-10, 10), trend.method = "locfit" (stood in for by the
    local-constant tricube smoother), tagwise = TRUE, robust = FALSE (the function's own default) or TRUE
    (what ruvIII.nb passes: per-gene prior.df from limma's robust squeezeVar).  Returns the reference's list elements
    (common.dispersion, trended.dispersion, tagwise.dispersion, span, prior.df, prior.n).  Arguments the device path does
    not implement raise ValueError instead of being ignored.  `replicate_of_cell` (0-based group of every cell) lets an
    offset with a replicate-group part of any size be taken apart; `BPPARAM` is accepted and ignored."""
    Y = np.asarray(y)
    if Y.ndim != 2:
        raise ValueError("y must be a matrix")
    G, N = Y.shape
    if G == 0:
        return dict(span=span, **{"prior.df": prior_df, "prior.n": None})                       # :13-14
    if design is not None:
        d = np.asarray(design, dtype=np.float64)
        if d.shape not in ((N,), (N, 1)) or not np.all(d == 1.0):
            raise ValueError("only design = matrix(1, ncol(y)) (one intercept) is on the device path")
    if group is not None and len(np.unique(np.asarray(group))) != 1:
        raise ValueError("a grouping with more than one level needs a design with more than one coefficient: not on the device path")
    if group is not None and len(np.asarray(group)) != N:
        raise ValueError("Incorrect length of group.")                                              # :19-20
    if weights is not None:
        raise ValueError("observation weights are derived on the device from the ZINB state (opts.zinb); a weight matrix is not accepted")
    if trend_method != "locfit" or not tagwise or span is not None:
        raise ValueError("only trend.method = 'locfit', tagwise = TRUE, span = NULL are on the device path")
    if tuple(winsor_tail_p) != (0.05, 0.1):
        raise ValueError("winsor.tail.p = c(0.05, 0.1) (the default every call site of the reference uses) is the one on the device path")
    if grid_length != 21 or tuple(grid_range) != (-10, 10) or min_row_sum != 5:
        raise ValueError("the device grid is 0.1 * 2^seq(-10, 10, length = 21) with min.row.sum = 5")
    if lib_size is not None and len(np.asarray(lib_size)) != N:
        raise ValueError("Incorrect length of lib.size.")                                           # :24-25
    if offset is None:                                                                              # .compressOffsets: log(lib.size)
        ls = Y.sum(axis=0).astype(np.float64) if lib_size is None else np.asarray(lib_size, dtype=np.float64)
        offset = np.broadcast_to(np.log(ls)[None, :], (G, N))
    offset = np.asarray(offset, dtype=np.float64)
    if offset.shape != (G, N):
        raise ValueError("offset must be a matrix of the dimensions of y")
    if not np.all(np.isfinite(offset)):
        raise ValueError("offsets must be finite values")
    grp = np.zeros(N, dtype=np.int32) if replicate_of_cell is None else np.asarray(replicate_of_cell, dtype=np.int32)
    if grp.shape != (N,) or grp.min() < 0:
        raise ValueError("replicate_of_cell must give a 0-based group for every column of y")
    m = int(grp.max()) + 1
    Mb, alpha, W = _factor_offset(offset, grp, m)
    k = alpha.shape[1]
    opts = _lib.default_opts(k=k)
    ctl = np.zeros(G, dtype=bool)
    ctl[0] = True                                   # the dispersion step does not read the control set
    with _lib.Context(device) as ctx:
        ctx.set_design(G, N, grp, ctl)
        ctx.upload_counts(np.ascontiguousarray(Y, dtype=np.int32))
        ctx.set_state("W", W); ctx.set_state("alpha", alpha); ctx.set_state("Mb", Mb)
        out = ctx.estimate_disp_r(opts, robust=bool(robust), prior_df=prior_df)
        tag = ctx.get_state("psi")
    ns = int((Y.sum(axis=1) >= 5).sum())
    pdf = out["prior_df"] if robust else float(out["prior_df"][Y.sum(axis=1) >= 5][0])        # :192-198: per gene when robust
    return {"common.dispersion": out["common"], "trended.dispersion": out["trended"], "tagwise.dispersion": tag,
            "span": 1.0 if ns <= 50 else 0.25 + 0.75 * np.sqrt(50.0 / ns),                          # edgeR::WLEB default
            "prior.df": pdf, "prior.n": pdf / (N - 1.0)}                                           # :171, ncoefs = 1


def _top_right_singular_vectors(Ysub, grp, k, device):
    """U0 = irlba(RM_Z, nv = k)$v (R/fastruvIIInb.R:223-228), RM_Z = log(Ysub + 1) minus its replicate-group means: the k leading
    eigenvectors of the nsub x nsub Gram RM_Z' RM_Z.  irlba starts from a random vector, so the reference's own U0 is defined up to
    its tolerance and sign only (parity runs inject U0); this is set-up plumbing, not the hot path: the Gram and its symmetric
    eigen-decomposition run through torch on the context's GPU (a 30 000 x 3 000 dense SVD takes half a minute on one host thread)."""
    import torch
    dev = torch.device("cuda", int(device))
    Z = torch.log(torch.as_tensor(np.ascontiguousarray(Ysub), device=dev).to(torch.float64) + 1.0)
    g = torch.as_tensor(np.asarray(grp, dtype=np.int64), device=dev)
    m = int(g.max().item()) + 1
    for j in range(m):
        sel = g == j
        Z[:, sel] -= Z[:, sel].mean(dim=1, keepdim=True)
    w, V = torch.linalg.eigh(Z.T @ Z)
    U0 = V[:, torch.argsort(w, descending=True)[:k]].cpu().numpy()
    for j in range(k):                                   # a fixed sign: largest entry positive
        if U0[np.argmax(np.abs(U0[:, j])), j] < 0:
            U0[:, j] = -U0[:, j]
    return np.ascontiguousarray(U0)


def _estimate_disp_subset(Ysub, grp, ctl, alpha, W, Mb, cols, k, device):
    """estimateDisp.par(Ysub[, cols], design = 1, offset = (alpha W' + Mb[, grp])[, cols]) (R/fastruvIIInb.R:273-311,569-575)
    on a second context that holds only those columns."""
    cols = np.asarray(cols)
    g = grp[cols]
    present, gc = np.unique(g, return_inverse=True)       # replicate groups that occur among the psi cells
    opts = _lib.default_opts(k=int(k))
    with _lib.Context(device) as c2:
        c2.set_design(Ysub.shape[0], len(cols), gc.astype(np.int32), ctl)
        c2.upload_counts(np.ascontiguousarray(Ysub[:, cols], dtype=np.int32))
        c2.set_state("W", W[cols])
        c2.set_state("alpha", alpha)
        c2.set_state("Mb", Mb[:, present])
        try:
            c2.estimate_disp_r(opts, robust=True)           # robust = TRUE, R/fastruvIIInb.R:282,569
        except _lib.RuvnbError:
            return np.full(Ysub.shape[0], np.nan)         # tryCatch(..., error = function(e) rep(NA, ngene))
        return c2.get_state("psi")


def fastruvIII_nb_subfit(Ysub, Msub, ctl, k=2, lambda_a=0.01, lambda_b=16, step_fac=0.5, inner_maxit=50, outer_maxit=25,
                         U0=None, psi_idx=None, psi=None, device=0, seed=0, trace=None):
    """The subsample fit of `fastruvIII.nb` (R/fastruvIIInb.R:223-623) on the already-subsampled cells (every cell of
    Msub annotated).  U0: the k right singular vectors of RM_Z (irlba, :228) -- computed by a dense SVD when None;
    psi_idx: the <= 100 cells used for the dispersion (:215-221) -- drawn with NumPy when None; psi: injected dispersion
    schedule (list of G-vectors) instead of estimateDisp.  Returns dict(alpha, gmean, Mb, Wsub, psi, wt_ctl, logl)."""
    Ysub = np.ascontiguousarray(Ysub, dtype=np.int32)
    G, ns = Ysub.shape
    Msub = np.asarray(Msub).astype(bool)
    if (Msub.sum(axis=1) != 1).any():
        raise ValueError("every subsampled cell must belong to exactly one replicate group")
    grp = np.argmax(Msub, axis=1).astype(np.int32)
    ctl = np.asarray(ctl, bool)
    rng = np.random.default_rng(seed)
    if psi_idx is None:
        psi_idx = np.sort(rng.choice(ns, size=min(100, int(round(0.5 * ns))), replace=False))    # :218
    if U0 is None:
        U0 = _top_right_singular_vectors(Ysub, grp, k, device)
    opts = _lib.default_opts(k=int(k), lambda_a=float(lambda_a), lambda_b=float(lambda_b), step_fac=float(step_fac),
                             inner_maxit=int(inner_maxit), outer_maxit=int(outer_maxit))
    calls = {"i": 0}

    def new_psi(ctx):
        if psi is not None:
            i = min(calls["i"], len(psi) - 1)
            calls["i"] += 1
            return np.asarray(psi[i], dtype=np.float64)
        return _estimate_disp_subset(Ysub, grp, ctl, ctx.get_state("alpha"), ctx.get_state("W"), ctx.get_state("Mb"), psi_idx, k, device)

    with _lib.Context(device) as ctx:
        ctx.set_design(G, ns, grp, ctl)
        ctx.upload_counts(Ysub)
        ctx.fast_init(opts, U0)                                                              # F3
        p0 = new_psi(ctx)
        ctx.set_state("psi", np.where(np.isnan(p0), 0.0, p0))                               # :311
        conv, it_outer, logl_outer = False, 0, []
        while not conv:                                                                      # :319
            it_outer += 1
            logl_beta = []
            s_old = ctx.fast_begin_outer(opts)
            ctx.snapshot_best()                                                              # :352
            it, halving, conv_beta = 1, 0, False
            while not conv_beta:
                s_new, bad = ctx.fast_iteration(opts)
                degener = (not (np.isfinite(s_old) and np.isfinite(s_new))) or s_old > s_new   # :533-534
                rec = dict(outer=it_outer, iter=it, degener=degener, logl_new=s_new)
                if degener:
                    ctx.fast_halve_steps(opts)                                               # :538-539
                    ctx.restore_best()
                    halving += 1
                    if halving >= 3:
                        degener = False                                                      # loglik.tmp <- loglik
                else:
                    ctx.fast_accept()
                    s_old = s_new
                if not degener:
                    logl_beta.append(s_old)
                    ctx.snapshot_best()
                    it += 1
                    halving = 0
                rec["accepted"] = not degener
                if trace is not None:
                    trace.append(rec)
                conv_logl = it > 2 and (logl_beta[-1] - logl_beta[-2]) / abs(logl_beta[-2]) < 1e-4
                conv_beta = it >= inner_maxit or conv_logl
            updt = True
            if it_outer > 1:
                updt = abs(logl_outer[-1] - s_old) / abs(logl_outer[-1]) > 1e-8
            ctx.restore_best()
            if updt:
                pn = new_psi(ctx)
                cur = ctx.get_state("psi")
                ctx.set_state("psi", np.where(np.isfinite(pn), pn, cur))                     # :577
            logl_outer.append(ctx.fast_begin_outer(opts))                                    # :588-604
            if it_outer > 1:
                conv = (logl_outer[-1] - logl_outer[-2]) / abs(logl_outer[-2]) < 1e-4 or it_outer >= outer_maxit
        ctx.snapshot_best()
        return {"alpha": ctx.get_state("alpha"), "gmean": ctx.get_state("gmean"), "Mb": ctx.get_state("Mb"),
                "Wsub": ctx.get_state("W"), "psi": ctx.get_state("psi"), "wt_ctl": ctx.get_state("wt_ctl"),
                "logl": np.array(logl_outer), "psi_idx": psi_idx}


class Ranks:
    """How the ranks of a cell-sharded run talk on the host side (small objects only: labels, the <= 3000 subsampled columns).
    `dist` is torch.distributed (any backend) or None for a single process; the bulk collectives (histograms, W column sums,
    the W-sweep criterion) run inside the library over NCCL (`ruvnb_set_cell_shard`)."""

    def __init__(self, dist=None, device=None):
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.device = device

    def allgather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def unique_id(self):
        from . import shard
        return shard.broadcast_unique_id(self.dist, _lib.load(), device=self.device)


def cell_shard_range(N_total, world, rank, block_size):
    """Cells [c0, c1) of rank `rank`: whole blocks of `block_size` cells (blocks are the unit of independence of get.res and of
    the fastruvIII.nb passes, R/get.res.R:27-126, R/fastruvIIInb.R:639-689,727-738), as evenly as the block count allows."""
    nb = -(-N_total // block_size)
    per = -(-nb // world)
    return min(N_total, rank * per * block_size), min(N_total, (rank + 1) * per * block_size)


def fastruvIII_nb(Y, M, ctl, k=2, lambda_a=0.01, lambda_b=16, step_fac=0.5, inner_maxit=50, outer_maxit=25, pCells_touse=0.2,
                  block_size=5000, subsamples=None, U0=None, psi_idx=None, device=0, seed=0, ranks=None, N_total=None, c0=0,
                  assay="logPAC", return_Mb_all=True, pinned_out=None):
    """`fastruvIII.nb` (R/fastruvIIInb.R:29-792), nbatch == 1, no pseudo-cells: Winsorize (:81-95), stratified subsample of the
    annotated cells (<= 3000, :118-137; NumPy RNG unless `subsamples` is given), the subsample fit (:223-623), W for all cells
    (:632-716), Mb.all + the percentile-adjusted counts of `makeSCE(assays = 'logPAC')` in one pass per block (:719-748,790),
    the psi cap (:765-770).  The count matrix is uploaded ONCE and never comes back: the subsample's columns are gathered on
    the device.

    Cell shards: with `ranks` (a `Ranks` over torch.distributed), Y / M are THIS rank's cells [c0, c0 + N) of N_total (c0 on
    a block boundary: `cell_shard_range`); `subsamples` are global cell indices.  Per-cell outputs (W, Mb, the assay) are the
    rank's own cells; per-gene outputs are identical on every rank.
    Returns the reference's `out` list as a dict plus the assay ("logPAC": log(PAC + 1) doubles, "PAC": int32, None)."""
    ranks = ranks or Ranks()
    Y = np.asarray(Y)
    M = np.asarray(M).astype(bool)
    G, N = Y.shape
    N_total = N if N_total is None else int(N_total)
    ctl = np.asarray(ctl, bool)
    annotated = M.sum(axis=1) > 0
    grp_local = np.where(annotated, np.argmax(M, axis=1), -1).astype(np.int32)
    annot_all = np.concatenate(ranks.allgather(grp_local))             # replicate group of every cell of the problem, -1 = none
    if annot_all.size != N_total:
        raise ValueError("the ranks' cells do not add up to N_total")
    if (annot_all >= 0).mean() < 0.01:
        raise ValueError("Less than 1% of cells have replicate annotation")                 # :53
    m = M.shape[1]
    if (np.bincount(annot_all[annot_all >= 0], minlength=m) == 0).any():
        raise ValueError("Each column of M matrix must have at least one non-zero entry")   # :60
    rng = np.random.default_rng(seed)
    if subsamples is None:                                                                   # :118-137 (same draw on every rank)
        p = min(pCells_touse, 3000.0 / (annot_all >= 0).sum())
        subsamples = np.sort(np.concatenate([rng.choice(np.where(annot_all == j)[0], size=int(np.ceil(p * (annot_all == j).sum())), replace=False)
                                             for j in range(m)]))
    subsamples = np.asarray(subsamples, dtype=np.int64)
    with _lib.Context(device) as ctx:
        if ranks.world > 1:
            ctx.comm_init(ranks.unique_id(), ranks.rank, ranks.world)
        ctx.set_design(G, N, np.where(annotated, grp_local, 0).astype(np.int32), ctl, m=m)
        if ranks.world > 1 or N_total != N:
            ctx.set_cell_shard(N_total, c0)
        ctx.upload_counts(np.ascontiguousarray(Y, dtype=np.int32))
        _, nz = ctx.winsorize()                                                              # :81-95, quantile over ALL cells
        keep = nz > 0                                                                        # :97-101
        if not keep.all():
            ctx.set_row_mask(keep)
        ctlk = ctl[keep]
        # the subsample's columns leave the device (G x <= 3000 int32), every rank its own, and are put in subsample order
        mine = (subsamples >= c0) & (subsamples < c0 + N)
        cols = ctx.get_cells(subsamples[mine] - c0)
        parts = ranks.allgather((np.where(mine)[0], cols))
        Ysub = np.empty((G, subsamples.size), dtype=np.int32)
        for idx, cc in parts:
            Ysub[:, idx] = cc
        Msub = np.zeros((subsamples.size, m), dtype=bool)
        Msub[np.arange(subsamples.size), annot_all[subsamples]] = True
        # the subsample fit is small (<= 3000 cells): every rank runs it on a second context and gets the same result
        sub = fastruvIII_nb_subfit(Ysub[keep], Msub, ctlk, k=k, lambda_a=lambda_a, lambda_b=lambda_b, step_fac=step_fac,
                                   inner_maxit=inner_maxit, outer_maxit=outer_maxit, U0=U0, psi_idx=psi_idx, device=device, seed=seed)
        # full-data passes on the resident counts: rows the cap emptied stay resident with neutral parameters and are not reported
        def full(x, fill=0.0):
            out = np.full((G,) + x.shape[1:], fill)
            out[keep] = x
            return out
        W_init = np.zeros((N, k))
        W_init[subsamples[mine] - c0] = sub["Wsub"][mine]                                     # :635-636
        med = np.median(sub["Wsub"], axis=0)
        mad = 1.4826 * np.median(np.abs(sub["Wsub"] - med[None, :]), axis=0)                 # :653-654
        wt_full = np.zeros(int(ctl.sum()))
        wt_full[keep[ctl]] = sub["wt_ctl"]
        ctx.set_state("W", W_init)
        ctx.set_state("alpha", full(sub["alpha"]))
        ctx.set_state("gmean", full(sub["gmean"]))
        ctx.set_state("psi", full(sub["psi"], 1.0))
        ctx.set_state("Mb", full(sub["Mb"]))
        W, sweeps, crit = ctx.fast_W_all(W_init, wt_full, med, mad, lambda_a=lambda_a)       # :632-716
        res = None
        Mb_all = None
        if assay is not None or return_Mb_all:
            kind = {"logPAC": _lib.OUT_LOG1P, "PAC": _lib.OUT_I32, None: _lib.OUT_NONE}[assay]
            res, Mb_all, st = ctx.fast_res(_lib.RES_KINDS["quantile"], grp_local, lambda_b=lambda_b, block_size=block_size,
                                           out_kind=kind, out=pinned_out, want_Mb_all=return_Mb_all)   # :719-748 + makeSCE
            if Mb_all is not None:
                Mb_all = Mb_all[keep]
        counts = ctx.get_counts()[keep] if return_Mb_all else None   # the `counts` element (small problems / tests only)
    out = {"counts": counts, "W": W, "M": M, "ctl": ctlk, "logl": sub["logl"], "a": sub["alpha"], "Mb": Mb_all, "gmean": sub["gmean"],
           "psi": _lib.fast_cap_psi(sub["psi"]), "L.a": lambda_a, "L.b": lambda_b, "subsamples": subsamples, "Wsub": sub["Wsub"],
           "Mb.sub": sub["Mb"], "W_sweeps": sweeps, "zero_genes": ~keep, assay or "assay": res}
    return out


def fastruvIII_nb_finish(Y, M, ctl, alpha, gmean, psi, Mb_sub, Wsub, subsamples, wt_ctl, lambda_a=0.01, lambda_b=16,
                         block_size=5000, device=0):
    """The full-data tail of `fastruvIII.nb` (R/fastruvIIInb.R:625-792) given the subsample fit (alpha, gmean, psi,
    Mb_sub G x m, Wsub rows for the cells `subsamples`, wt_ctl = best.wtctl): W for all cells, Mb.all, the psi cap.
    The subsample fit (:118-623) is `fastruvIII_nb_subfit`."""
    from . import _lib as L
    Y = np.ascontiguousarray(Y, dtype=np.int32)
    G, N = Y.shape
    M = np.asarray(M).astype(bool)
    annotated = M.sum(axis=1) > 0
    if annotated.mean() < 0.01:
        raise ValueError("Less than 1% of cells have replicate annotation")                 # R/fastruvIIInb.R:53
    if (M.sum(axis=0) == 0).any():
        raise ValueError("Each column of M matrix must have at least one non-zero entry")   # :60
    grp = np.where(annotated, np.argmax(M, axis=1), 0).astype(np.int32)
    annot = np.where(annotated, np.argmax(M, axis=1), -1).astype(np.int32)
    k = Wsub.shape[1]
    W_init = np.zeros((N, k))
    W_init[np.asarray(subsamples)] = Wsub                                                    # :635-636
    med = np.median(Wsub, axis=0)
    mad = 1.4826 * np.median(np.abs(Wsub - med[None, :]), axis=0)                            # :653-654
    with L.Context(device) as ctx:
        ctx.set_design(G, N, grp, np.asarray(ctl, bool))
        ctx.upload_counts(Y)
        ctx.set_state("W", W_init)
        ctx.set_state("alpha", alpha)
        ctx.set_state("gmean", gmean)
        ctx.set_state("psi", psi)
        ctx.set_state("Mb", Mb_sub)
        W, sweeps, crit = ctx.fast_W_all(W_init, wt_ctl, med, mad, lambda_a=lambda_a)
        Mb_all = ctx.fast_Mb_all(annot, lambda_b=lambda_b, block_size=block_size)
    return {"counts": Y, "W": W, "M": M, "ctl": np.asarray(ctl, bool), "a": alpha, "Mb": Mb_all, "gmean": gmean,
            "psi": L.fast_cap_psi(psi), "L.a": lambda_a, "L.b": lambda_b, "subsamples": subsamples, "Wsub": Wsub,
            "Mb.sub": Mb_sub, "W_sweeps": sweeps}
