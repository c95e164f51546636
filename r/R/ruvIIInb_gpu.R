# ruvIIInb_gpu.R -- the exported functions of limfuxing/ruvIIInb with their bodies routed through libruvnb_b200 (.Call shim in
# src/ruvnb_shim.c).  Signatures, defaults, stop() texts and return objects are the reference's (R/ruvIIInb.R:27-28,627-628;
# R/fastruvIIInb.R:29-30,53,60,111,786-791; R/estimateDisp.par.R:3-8,199-201; R/get.res.R:12; R/makeSCE.R:12): a user switches by
# loading this file after the package.  `ncores` / `BPPARAM` stay in the signatures and are ignored (the fork pool is the GPU).
# Not on the device path (they stop with a message): use.pseudosample, batch.disp with more than one batch, ortho.W.
# R is not installed in the build image of the library: these wrappers are exercised only where R and the package's imports exist.

.ruvnb_opts <- function(robust, lambda.a, lambda.b, step.fac, inner.maxit, outer.maxit, zinb = FALSE)
  list(robust = robust, lambda.a = lambda.a, lambda.b = lambda.b, step.fac = step.fac, inner.maxit = as.integer(inner.maxit),
       outer.maxit = as.integer(outer.maxit), zinb = zinb)

.ruvnb_unsupported <- function(use.pseudosample, batch.disp, batch, ortho.W) {
  if (use.pseudosample) stop("use.pseudosample = TRUE is not on the GPU path")
  if (ortho.W) stop("ortho.W = TRUE is not on the GPU path")
  if (batch.disp && !is.null(batch) && length(unique(batch)) > 1) stop("batch-specific dispersion (batch.disp = TRUE) is not on the GPU path")
}

ruvIII.nb <- function(Y, M, ctl, k = 2, robust = FALSE, ortho.W = FALSE, lambda.a = 0.01, lambda.b = 16, batch = NULL, step.fac = 0.5,
                      inner.maxit = 50, outer.maxit = 25, ncores = 2, use.pseudosample = FALSE, nc.pool = 20, batch.disp = FALSE,
                      zeroinf = NULL, device = 0L) {
  .ruvnb_unsupported(use.pseudosample, batch.disp, batch, ortho.W)
  if (is.null(zeroinf)) zeroinf <- rep(FALSE, ncol(Y))                                   # R/ruvIIInb.R:39-40
  if (!is.logical(ctl)) ctl <- rownames(Y) %in% ctl                                      # :42-45
  mode(M) <- "logical"                                                                   # :47-48
  strata <- apply(M, 1, which)                                                           # :50
  if (is.list(strata) || length(strata) != ncol(Y))
    stop("every cell must belong to exactly one replicate group (ruvIII.nb indexes Mb[, apply(M, 1, which)], R/ruvIIInb.R:206)")
  if (is.null(batch)) batch <- rep(1, ncol(Y))                                           # :64-66
  storage.mode(Y) <- "integer"
  raw.zero <- rowSums(Y > 0) == 0                      # certainly empty after the cap too: never uploaded
  ctx <- .Call("ruvnb_R_open", Y[!raw.zero, , drop = FALSE], as.integer(strata - 1L), ncol(M), ctl[!raw.zero],
               if (any(zeroinf)) as.logical(zeroinf) else NULL, as.integer(device))
  on.exit(.Call("ruvnb_R_close", ctx), add = TRUE)
  zero.genes <- raw.zero
  zero.genes[!raw.zero] <- attr(ctx, "nonzero") == 0                                     # :57 -- AFTER the Winsorization
  if (any(attr(ctx, "nonzero") == 0)) {                # rows emptied by the cap: a context over the surviving rows (:58-62)
    Yw <- .Call("ruvnb_R_counts", ctx, sum(!raw.zero), ncol(Y))
    .Call("ruvnb_R_close", ctx)
    ctx <- .Call("ruvnb_R_open", Yw[attr(ctx, "nonzero") > 0, , drop = FALSE], as.integer(strata - 1L), ncol(M), ctl[!zero.genes],
                 if (any(zeroinf)) as.logical(zeroinf) else NULL, as.integer(device))
  }
  if (any(zero.genes)) print(paste0(sum(zero.genes), "genes with all zero counts are removed following a Winsorization step."))
  ctl <- ctl[!zero.genes]
  G <- sum(!zero.genes); N <- ncol(Y)
  fit <- .Call("ruvnb_R_fit", ctx, G, N, ncol(M), as.integer(k), NULL, NULL,
               .ruvnb_opts(robust, lambda.a, lambda.b, step.fac, inner.maxit, outer.maxit, zinb = any(zeroinf)))
  counts <- .Call("ruvnb_R_counts", ctx, G, N)
  dimnames(counts) <- list(rownames(Y)[!zero.genes], colnames(Y))
  pi0 <- if (any(zeroinf)) Matrix::Matrix(outer(fit$pi0, as.numeric(zeroinf)), sparse = TRUE) else NULL   # :611-626
  list(counts = methods::as(counts, "sparseMatrix"), W = fit$W, M = M, ctl = ctl, logl = fit$logl, a = fit$a, Mb = fit$Mb,
       gmean = fit$gmean, pi0 = pi0, psi = fit$psi, L.a = lambda.a, L.b = lambda.b, batch = batch)   # :627-628
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
  ctx <- .Call("ruvnb_R_open", y, grp, m, ctl, NULL, as.integer(device))
  on.exit(.Call("ruvnb_R_close", ctx), add = TRUE)
  .Call("ruvnb_R_set_fit", ctx, W, alpha, NULL, Mb, NULL)
  out <- .Call("ruvnb_R_estimate_disp", ctx, ntags, kk, robust, prior.df, list())
  ns <- sum(rowSums(y) >= 5)
  pdf <- if (robust) out$prior.df else out$prior.df[rowSums(y) >= 5][1]
  list(common.dispersion = out$common.dispersion, trended.dispersion = out$trended.dispersion, tagwise.dispersion = out$tagwise.dispersion,
       span = if (ns <= 50) 1 else 0.25 + 0.75 * sqrt(50 / ns), prior.df = pdf, prior.n = pdf / (nlibs - 1))   # :199-201
}

get.res <- function(out, type = c("logcounts", "pearson", "quantile"), batch = NULL, block.size = 5000, device = 0L) {
  if (!is.null(out$pi0)) stop("get.res for zero-inflated fits is not on the GPU path")
  if (!is.null(dim(out$psi))) stop("batch-specific dispersion is not on the GPU path")
  last <- tail(intersect(c("pearson", "logcounts", "quantile"), type), 1)                # the reference keeps the last one (R/get.res.R:49,68,75)
  Y <- as.matrix(out$counts); storage.mode(Y) <- "integer"
  M <- out$M; mode(M) <- "logical"
  per.cell <- ncol(out$Mb) == ncol(Y) && ncol(out$Mb) != ncol(M)                         # fastruvIII.nb's Mb.all vs a ruvIII.nb fit
  grp <- if (all(rowSums(M) == 1)) apply(M, 1, which) - 1L else rep(0L, ncol(Y))
  m <- max(grp) + 1L
  ctx <- .Call("ruvnb_R_open", Y, as.integer(grp), m, out$ctl, NULL, as.integer(device))   # (the counts of `out` are already capped)
  on.exit(.Call("ruvnb_R_close", ctx), add = TRUE)
  .Call("ruvnb_R_set_fit", ctx, out$W, out$a, out$gmean, if (per.cell) matrix(0, nrow(Y), m) else out$Mb, out$psi)
  res <- .Call("ruvnb_R_get_res", ctx, nrow(Y), ncol(Y), match(last, c("pearson", "logcounts", "quantile")), as.integer(block.size),
               if (per.cell) as.matrix(out$Mb) else NULL)
  dimnames(res) <- dimnames(out$counts)
  DelayedArray::DelayedArray(res)                                                         # R/get.res.R:127-128
}

fastruvIII.nb <- function(Y, M, cData = NULL, ctl, k = 2, robust = FALSE, ortho.W = FALSE, lambda.a = 0.01, lambda.b = 16, batch = NULL,
                          step.fac = 0.5, inner.maxit = 50, outer.maxit = 25, ncores = 2, use.pseudosample = FALSE, nc.pool = 20,
                          batch.disp = FALSE, pCells.touse = 0.2, block.size = 5000, device = 0L) {
  .ruvnb_unsupported(use.pseudosample, batch.disp, batch, ortho.W)
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
