# ruvIIInb_gpu.R -- the exported functions of limfuxing/ruvIIInb with their bodies routed through libruvnb_b200 (.Call shim in
# src/ruvnb_shim.c).  Signatures, defaults, stop() texts and return objects are the reference's (R/ruvIIInb.R:27-28,627-628;
# R/fastruvIIInb.R:29-30,53,60,111,786-791; R/estimateDisp.par.R:3-8,199-201; R/get.res.R:12; R/makeSCE.R:12): a user switches by
# loading this file after the package.  `ncores` / `BPPARAM` stay in the signatures and are ignored (the fork pool is the GPU).
# batch.disp = TRUE runs one estimateDisp.par per batch through the fit's dispersion hook; use.pseudosample = TRUE builds the
# pseudo-cells on the host (gtools::quantcut, rbinom) before the fit.  Not on the device path (stops with a message): ortho.W.
# R is not installed in the build image of the library: these wrappers are exercised only where R and the package's imports exist.

.ruvnb_opts <- function(robust, lambda.a, lambda.b, step.fac, inner.maxit, outer.maxit, zinb = FALSE)
  list(robust = robust, lambda.a = lambda.a, lambda.b = lambda.b, step.fac = step.fac, inner.maxit = as.integer(inner.maxit),
       outer.maxit = as.integer(outer.maxit), zinb = zinb)

.ruvnb_unsupported <- function(ortho.W) {
  if (ortho.W) stop("ortho.W = TRUE is not on the GPU path")
}

# a context over counts given as a dense matrix or as a dgCMatrix (its slots go to the device as they are)
.ruvnb_open <- function(Y, grp0, m, ctl, zeroinf = NULL, batch0 = NULL, capped = FALSE, device = 0L) {
  if (methods::is(Y, "dgCMatrix"))
    return(.Call("ruvnb_R_open_csc", Y@p, Y@i, Y@x, dim(Y), as.integer(grp0), as.integer(m), as.logical(ctl), zeroinf,
                 if (is.null(batch0)) NULL else as.integer(batch0), capped, as.integer(device)))
  Y <- as.matrix(Y); storage.mode(Y) <- "integer"
  .Call("ruvnb_R_open", Y, as.integer(grp0), as.integer(m), as.logical(ctl), zeroinf, if (is.null(batch0)) NULL else as.integer(batch0),
        capped, as.integer(device))
}

# as.numeric(gtools::quantcut(x, q)) and the pseudo-cells of R/ruvIIInb.R:113-141 (the branch that runs: strata is never NULL)
.ruvnb_pseudocells <- function(Y, M, batch, nc.pool, batch.disp) {
  strata <- apply(M, 1, which); sf <- colSums(Y); sf <- sf / mean(sf); nsc <- ncol(Y)
  batch.ext <- batch
  for (st in unique(na.omit(strata))) {
    temp.M <- M
    for (btc in unique(batch)) {
      sel <- batch == btc & strata == st
      nq <- max(round(sum(sel) / nc.pool), 1)
      if (sum(sel) > 1) {
        qsf <- as.numeric(gtools::quantcut(sf[sel], q = nq))
        for (q in 1:max(qsf)) {
          pool <- Y[, 1:nsc, drop = FALSE][, sel, drop = FALSE][, qsf == q, drop = FALSE]
          Y <- cbind(Y, rbinom(nrow(pool), size = rowSums(pool), prob = 1 / ncol(pool)))
          batch.ext <- c(batch.ext, btc + max(batch)); temp.M <- rbind(temp.M, FALSE)
        }
      }
    }
    if (sum(strata == st) > 1 && nrow(temp.M) > nrow(M)) {
      ps.mcol <- rep(FALSE, nrow(temp.M)); ps.mcol[seq(nrow(M) + 1, nrow(temp.M), 1)] <- TRUE
      M <- cbind(temp.M, ps.mcol)
    }
  }
  if (!batch.disp) batch.ext <- ifelse(batch.ext > max(batch), 2, 1)
  list(Y = Y, M = M, batch = batch.ext)
}

ruvIII.nb <- function(Y, M, ctl, k = 2, robust = FALSE, ortho.W = FALSE, lambda.a = 0.01, lambda.b = 16, batch = NULL, step.fac = 0.5,
                      inner.maxit = 50, outer.maxit = 25, ncores = 2, use.pseudosample = FALSE, nc.pool = 20, batch.disp = FALSE,
                      zeroinf = NULL, device = 0L) {
  .ruvnb_unsupported(ortho.W)
  if (is.null(zeroinf)) zeroinf <- rep(FALSE, ncol(Y))                                   # R/ruvIIInb.R:39-40
  if (!is.logical(ctl)) ctl <- rownames(Y) %in% ctl                                      # :42-45
  mode(M) <- "logical"                                                                   # :47-48
  strata <- apply(M, 1, which)                                                           # :50
  if (is.list(strata) || length(strata) != ncol(Y))
    stop("every cell must belong to exactly one replicate group (ruvIII.nb indexes Mb[, apply(M, 1, which)], R/ruvIIInb.R:206)")
  if (is.null(batch) & batch.disp) {                                                     # :65-68
    print(paste0("Batch variable not supplied...cannot fit batch-specific dispersion parameter.")); batch.disp <- FALSE
  }
  if (is.null(batch) & use.pseudosample) {                                               # :71-74
    print(paste0("Batch variable not supplied...cannot use pseudosample to define rep matrix")); use.pseudosample <- FALSE
  }
  if (is.null(batch)) batch <- rep(1, ncol(Y))                                           # :77-80
  nsc.org <- ncol(Y)
  nbatch.org <- if (batch.disp) length(unique(batch)) else 1                             # :88-91
  if (use.pseudosample) zeroinf <- rep(FALSE, ncol(Y))                                   # :147-152
  sparse.in <- methods::is(Y, "dgCMatrix")
  raw.zero <- Matrix::rowSums(Y > 0) == 0              # certainly empty after the cap too: never uploaded
  zi <- if (any(zeroinf)) as.logical(zeroinf) else NULL
  b0 <- if (batch.disp && length(unique(batch)) > 1) match(batch, sort(unique(batch))) - 1L else NULL
  ctx <- .ruvnb_open(Y[!raw.zero, , drop = FALSE], strata - 1L, ncol(M), ctl[!raw.zero], zi, b0, FALSE, device)
  on.exit(.Call("ruvnb_R_close", ctx), add = TRUE)
  zero.genes <- raw.zero
  zero.genes[!raw.zero] <- attr(ctx, "nonzero") == 0                                     # :57 -- AFTER the Winsorization
  Mfit <- M; bfit <- batch
  if (any(attr(ctx, "nonzero") == 0) || use.pseudosample) {
    # rows emptied by the cap leave (:58-62), pseudo-cells join (:83-144): a context over that matrix, already capped
    Yw <- .Call("ruvnb_R_counts", ctx, sum(!raw.zero), ncol(Y))[attr(ctx, "nonzero") > 0, , drop = FALSE]
    .Call("ruvnb_R_close", ctx)
    if (use.pseudosample) {
      ps <- .ruvnb_pseudocells(Yw, M, batch, nc.pool, batch.disp)
      Yw <- ps$Y; Mfit <- ps$M; bfit <- ps$batch; strata <- apply(Mfit, 1, which); zi <- NULL
      b0 <- match(bfit, sort(unique(bfit))) - 1L
    }
    storage.mode(Yw) <- "integer"
    ctx <- .ruvnb_open(Yw, strata - 1L, ncol(Mfit), ctl[!zero.genes], zi, b0, TRUE, device)
  }
  if (any(zero.genes)) print(paste0(sum(zero.genes), "genes with all zero counts are removed following a Winsorization step."))
  ctl <- ctl[!zero.genes]
  G <- sum(!zero.genes); N <- length(strata); m <- ncol(Mfit)
  opts <- .ruvnb_opts(robust, lambda.a, lambda.b, step.fac, inner.maxit, outer.maxit, zinb = !is.null(zi))
  hook <- NULL
  if (!is.null(b0)) {
    # estimateDisp.par once per batch on that batch's cells (:213-220,542-548): one context per batch, refreshed from the fit's hook
    Yall <- .Call("ruvnb_R_counts", ctx, G, N)
    bctx <- lapply(sort(unique(b0)), function(B) {
      idx <- which(b0 == B); present <- sort(unique(strata[idx]))
      list(idx = idx, present = present,
           ctx = .ruvnb_open(Yall[, idx, drop = FALSE], match(strata[idx], present) - 1L, length(present), ctl, if (is.null(zi)) NULL else zi[idx],
                             NULL, TRUE, device))
    })
    on.exit(for (b in bctx) .Call("ruvnb_R_close", b$ctx), add = TRUE)
    hook <- function(cx) {
      cur <- .Call("ruvnb_R_get_fit", cx, G, N, m, as.integer(k))
      psi.new <- cur$psi
      for (B in seq_along(bctx)) {
        b <- bctx[[B]]
        .Call("ruvnb_R_set_fit", b$ctx, cur$W[b$idx, , drop = FALSE], cur$a, NULL, cur$Mb[, b$present, drop = FALSE], NULL)
        if (!is.null(zi)) .Call("ruvnb_R_set_zinb", b$ctx, cur$pi0, cur$psi.w[, B], NULL, 0)
        est <- tryCatch(.Call("ruvnb_R_estimate_disp", b$ctx, G, as.integer(k), TRUE, NULL, opts)$tagwise.dispersion, error = function(e) NA)
        ok <- !is.na(est) & !is.infinite(est)                                            # :564
        if (length(est) == G) psi.new[ok, B] <- est[ok]
      }
      .Call("ruvnb_R_set_fit", cx, NULL, NULL, NULL, NULL, psi.new)
      invisible(NULL)
    }
  }
  fit <- .Call("ruvnb_R_fit", ctx, G, N, m, as.integer(k), NULL, NULL, hook, opts)
  counts <- .Call("ruvnb_R_counts", ctx, G, N)[, seq_len(nsc.org), drop = FALSE]
  dimnames(counts) <- list(rownames(Y)[!zero.genes], colnames(Y))
  pi0 <- if (!is.null(zi)) Matrix::Matrix(outer(fit$pi0, as.numeric(zeroinf)), sparse = TRUE) else NULL   # :611-626
  W <- fit$W; psi <- fit$psi
  if (use.pseudosample) {                                                                # :618-624
    W <- as.matrix(W[1:nsc.org, ]); Mfit <- Mfit[1:nsc.org, ]; psi <- if (is.matrix(psi)) psi[, 1:nbatch.org] else psi
  }
  list(counts = methods::as(counts, "sparseMatrix"), W = W, M = Mfit, ctl = ctl, logl = fit$logl, a = fit$a, Mb = fit$Mb,
       gmean = fit$gmean, pi0 = pi0, psi = psi, L.a = lambda.a, L.b = lambda.b, batch = batch)   # :627-628
}

estimateDisp.par <- function(y, design = NULL, group = NULL, lib.size = NULL, offset = NULL, prior.df = NULL, trend.method = "locfit",
                             tagwise = TRUE, span = NULL, min.row.sum = 5, grid.length = 21, grid.range = c(-10, 10), robust = FALSE,
                             winsor.tail.p = c(0.05, 0.1), tol = 1e-06, weights = NULL, BPPARAM = NULL, replicate.of.cell = NULL,
                             device = 0L, ...) {
  y <- as.matrix(y)
  ntags <- nrow(y); nlibs <- ncol(y)
  if (ntags == 0) return(list(span = span, prior.df = prior.df, prior.n = NULL))          # R/estimateDisp.par.R:13-14
  if (!is.null(group) && length(group) != nlibs) stop("Incorrect length of group.")       # :19-20
  if (!is.null(lib.size) && length(lib.size) != nlibs) stop("Incorrect length of lib.size.")   # :24-25
  if (!is.null(design) && (ncol(as.matrix(design)) != 1 || any(as.matrix(design) != 1))) stop("only design = matrix(1, ncol(y)) is on the GPU path")
  if (trend.method != "locfit" || !tagwise || !is.null(span) || grid.length != 21 || any(grid.range != c(-10, 10)) || min.row.sum != 5 ||
      any(winsor.tail.p != c(0.05, 0.1)) || !is.null(weights))
    stop("this combination of arguments is not on the GPU path (trend.method = 'locfit', tagwise = TRUE, the default grid, no weight matrix)")
  if (is.null(offset)) offset <- matrix(log(if (is.null(lib.size)) colSums(y) else lib.size), ntags, nlibs, byrow = TRUE)
  if (any(!is.finite(offset))) stop("offsets must be finite values")
  grp <- if (is.null(replicate.of.cell)) rep(0L, nlibs) else as.integer(replicate.of.cell)
  m <- max(grp) + 1L
  # offset = Mb[, grp] + alpha W' (the only offsets ruvIII.nb passes, R/ruvIIInb.R:205-221): group means, then a truncated SVD of the rest
  Mb <- sapply(seq_len(m) - 1L, function(j) rowMeans(offset[, grp == j, drop = FALSE]))
  Mb <- matrix(Mb, ntags, m)
  R <- offset - Mb[, grp + 1L, drop = FALSE]
  sv <- svd(R); kk <- max(1L, sum(sv$d > 1e-9 * max(1, max(abs(offset)))))
  if (kk > 8L) stop("offset is not of the form alpha W' + Mb[, group] with at most 8 factors")
  alpha <- sv$u[, seq_len(kk), drop = FALSE] %*% diag(sv$d[seq_len(kk)], kk); W <- sv$v[, seq_len(kk), drop = FALSE]
  storage.mode(y) <- "integer"
  ctl <- c(TRUE, rep(FALSE, ntags - 1L))
  ctx <- .ruvnb_open(y, grp, m, ctl, NULL, NULL, TRUE, device)     # (estimateDisp.par applies no count cap)
  on.exit(.Call("ruvnb_R_close", ctx), add = TRUE)
  .Call("ruvnb_R_set_fit", ctx, W, alpha, NULL, Mb, NULL)
  out <- .Call("ruvnb_R_estimate_disp", ctx, ntags, kk, robust, prior.df, list())
  ns <- sum(rowSums(y) >= 5)
  pdf <- if (robust) out$prior.df else out$prior.df[rowSums(y) >= 5][1]
  list(common.dispersion = out$common.dispersion, trended.dispersion = out$trended.dispersion, tagwise.dispersion = out$tagwise.dispersion,
       span = if (ns <= 50) 1 else 0.25 + 0.75 * sqrt(50 / ns), prior.df = pdf, prior.n = pdf / (nlibs - 1))   # :199-201
}

# get.res (R/get.res.R:12-129).  sparse = TRUE returns the dgCMatrix makeSCE would make of the result (every block is compacted on
# the device); log1p = TRUE (type 'quantile' only) returns log(PAC + 1).
get.res <- function(out, type = c("logcounts", "pearson", "quantile"), batch = NULL, block.size = 5000, device = 0L, sparse = FALSE,
                    log1p = FALSE) {
  last <- tail(intersect(c("pearson", "logcounts", "quantile"), type), 1)                # the reference keeps the last one (R/get.res.R:49,68,75)
  Y <- out$counts
  M <- out$M; mode(M) <- "logical"
  G <- nrow(Y); N <- ncol(Y)
  per.cell <- ncol(out$Mb) == N && ncol(out$Mb) != ncol(M)                               # fastruvIII.nb's Mb.all vs a ruvIII.nb fit
  grp <- if (all(rowSums(M) == 1)) apply(M, 1, which) - 1L else rep(0L, N)
  m <- max(grp) + 1L
  bpsi <- !is.null(dim(out$psi)) && ncol(out$psi) > 1                                    # out$psi[, curr.batch], :59-62,101-121
  if (is.null(batch)) batch <- if (is.null(out$batch)) rep(1, N) else out$batch
  if (bpsi && !all(batch %in% seq_len(ncol(out$psi)))) stop("batch must take values in 1..ncol(out$psi)")
  ctx <- .ruvnb_open(Y, grp, m, out$ctl, NULL, if (bpsi) as.integer(batch) - 1L else NULL, TRUE, device)   # (the counts of `out` are already capped)
  on.exit(.Call("ruvnb_R_close", ctx), add = TRUE)
  .Call("ruvnb_R_set_fit", ctx, out$W, out$a, out$gmean, if (per.cell) matrix(0, G, m) else as.matrix(out$Mb)[, seq_len(m), drop = FALSE],
        if (bpsi) as.matrix(out$psi) else as.numeric(out$psi))
  if (!is.null(out$pi0)) {                             # a zero-inflated fit: out$pi0 = outer(pi0, zeroinf), read as out$pi0[, batch[c]] (:53,84-85)
    P <- as.matrix(out$pi0); zi <- colSums(P != 0) > 0
    .Call("ruvnb_R_set_zinb", ctx, apply(P, 1, max), NULL, zi[batch],
          if (bpsi) as.numeric(zi[which.min(colMeans(out$psi))]) else mean(zi))           # :95 / :112-114
  }
  ty <- match(last, c("pearson", "logcounts", "quantile"))
  Mb.all <- if (per.cell) as.matrix(out$Mb) else NULL
  if (sparse) {
    s <- .Call("ruvnb_R_get_res_csc", ctx, G, N, ty, as.integer(block.size), Mb.all, log1p)
    return(Matrix::sparseMatrix(i = s$i, p = s$p, x = s$x, dims = c(G, N), dimnames = dimnames(out$counts), index1 = FALSE))
  }
  res <- .Call("ruvnb_R_get_res", ctx, G, N, ty, as.integer(block.size), Mb.all)
  if (log1p) res <- log(res + 1)
  dimnames(res) <- dimnames(out$counts)
  DelayedArray::DelayedArray(res)                                                         # R/get.res.R:127-128
}

# makeSCE (R/makeSCE.R:12-43): the assays come back from the device already sparse
makeSCE <- function(obj, cData = NULL, batch = NULL, assays = c("pearson", "logPAC"), device = 0L) {
  batch <- obj$batch
  if (is.null(cData)) stop("'cData' must be a user-specified data frame.")
  if (any(!assays %in% c("logcounts", "logPAC", "pearson")))
    stop("'Wrong names of corrected assays data was specified. Available corrected data assays are logcounts,logPAC and pearson")
  colnames(obj$a) <- paste0("alpha", 1:ncol(obj$a))
  features.data <- data.frame(obj$a, obj$psi)
  sce.obj <- SingleCellExperiment::SingleCellExperiment(assays = list(counts = methods::as(obj$counts, "sparseMatrix")), colData = cData,
                                                        rowData = features.data)
  one <- function(name, type, log1p = FALSE) {
    v <- tryCatch(get.res(obj, type = type, batch = batch, device = device, sparse = TRUE, log1p = log1p), error = function(e) NULL)
    if (is.null(v)) warning(paste0("failed to return ", name)) else SummarizedExperiment::assays(sce.obj, withDimnames = FALSE)[[name]] <<- v
  }
  if (any(assays == "logcounts")) one("logcounts", "logcounts")
  if (any(assays == "pearson")) one("pearson", "pearson")
  if (any(assays == "logPAC")) one("logPAC", "quantile", log1p = TRUE)
  sce.obj
}

fastruvIII.nb <- function(Y, M, cData = NULL, ctl, k = 2, robust = FALSE, ortho.W = FALSE, lambda.a = 0.01, lambda.b = 16, batch = NULL,
                          step.fac = 0.5, inner.maxit = 50, outer.maxit = 25, ncores = 2, use.pseudosample = FALSE, nc.pool = 20,
                          batch.disp = FALSE, pCells.touse = 0.2, block.size = 5000, device = 0L) {
  .ruvnb_unsupported(ortho.W)
  if (!is.logical(ctl)) ctl <- rownames(Y) %in% ctl
  if (any(rowSums(M) == 0)) {                                                             # R/fastruvIIInb.R:47-54
    n.unannot <- sum(rowSums(M) == 0)
    print(paste0(n.unannot, " cells have all zero entries in the M matrix. They will be considered as un-annotated cells"))
    if ((ncol(Y) - n.unannot) < (0.01 * ncol(Y)))
      stop("Less than 1% of cells have known annotation. Please increase the pct of cells with known annotation to 1%")
  }
  if (any(colSums(M) == 0))                                                               # :57-61
    stop(paste0("Columns ", which(colSums(M) == 0), " of the M matrix has zero sum. Pls remove these columns and re-run"))
  if (!is.null(batch) && !is.numeric(batch))                                              # :110-112
    stop(paste0("Batch variable is not a numeric vector. Please supply batch variable as numeric vector with 1,2...B values"))
  stop("fastruvIII.nb on the GPU: the subsample fit loop lives in the Python mirror (ruviiinb_b200/api.py fastruvIII_nb_subfit, a line-for-line ",
       "restatement of R/fastruvIIInb.R:319-623 over the ruvnb_fast_* entry points); the full-data passes are bound here as ",
       ".Call('ruvnb_R_fast_W_all', ...) and .Call('ruvnb_R_fast_res', ...)")
}
