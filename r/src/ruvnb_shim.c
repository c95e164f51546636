/* ruvnb_shim.c -- the `.Call` entry points that bind libruvnb_b200 (include/ruvnb.h) under the R API of limfuxing/ruvIIInb.
 *
 * The reference package is pure R (DESCRIPTION:35 NeedsCompilation: no, NAMESPACE:1-6); a maintainer adds this file as
 * src/ruvnb_shim.c, `useDynLib(ruvIIInb, .registration = TRUE)` to NAMESPACE and links with -lruvnb_b200 (r/src/Makevars).
 * No arithmetic happens here: every function converts R objects to the plain pointers of the C ABI and back.  R is not
 * installed in the build image, so the test-suite compiles this file against a stub of the R API (tests/stubs/) -- it is built
 * for real only where `R CMD config` exists (r/README.md).
 *
 * Conventions: matrices are R's column-major (the library's boundary layout); replicate groups arrive 0-based
 * (apply(M, 1, which) - 1L); argument errors become R errors through Rf_error with the library's message; numerical
 * "problematic entries" are not errors (the library reverts them as R/ruvIIInb.R:368,402,424-425 do).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "ruvnb.h"

static void ctx_finalizer(SEXP p) {
  ruvnb_ctx* c = (ruvnb_ctx*)R_ExternalPtrAddr(p);
  if (c) { ruvnb_destroy(c); R_ClearExternalPtr(p); }
}
static ruvnb_ctx* ctx_of(SEXP p) {
  ruvnb_ctx* c = (ruvnb_ctx*)R_ExternalPtrAddr(p);
  if (!c) Rf_error("ruvnb: the GPU context has been released");
  return c;
}
#define CHECK(ctx, call) do { int rc_ = (call); if (rc_ != RUVNB_OK) Rf_error("ruvnb: %s", ruvnb_last_error(ctx)); } while (0)

static SEXP list_get(SEXP lst, const char* name) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < Rf_xlength(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
  return R_NilValue;
}
static void opts_from_list(ruvnb_opts* o, SEXP lst, int k) {
  ruvnb_opts_default(o);
  o->k = k;
  SEXP v;
  if ((v = list_get(lst, "robust")) != R_NilValue) o->robust = Rf_asLogical(v) == 1;           /* R/ruvIIInb.R:155 */
  if ((v = list_get(lst, "lambda.a")) != R_NilValue) o->lambda_a = Rf_asReal(v);
  if ((v = list_get(lst, "lambda.b")) != R_NilValue) o->lambda_b = Rf_asReal(v);
  if ((v = list_get(lst, "step.fac")) != R_NilValue) o->step_fac = Rf_asReal(v);
  if ((v = list_get(lst, "inner.maxit")) != R_NilValue) o->inner_maxit = Rf_asInteger(v);
  if ((v = list_get(lst, "outer.maxit")) != R_NilValue) o->outer_maxit = Rf_asInteger(v);
  if ((v = list_get(lst, "zinb")) != R_NilValue) o->zinb = Rf_asLogical(v) == 1;
  if ((v = list_get(lst, "verbose")) != R_NilValue) o->verbose = Rf_asLogical(v) == 1;
}

/* context + design shared by the two `open` entry points.  batch: integer N, 0-based batch of every cell, or NULL (one dispersion
 * per gene); with batches the dispersion is psi[, batch] (R/ruvIIInb.R:278-283,306-315) */
static SEXP open_design(int G, int N, SEXP grp, SEXP m_, SEXP ctl, SEXP zeroinf, SEXP batch, SEXP device, ruvnb_ctx** out) {
  if (Rf_length(grp) != N || Rf_length(ctl) != G) Rf_error("ruvnb: M / ctl do not match the dimensions of Y");
  ruvnb_ctx* ctx = NULL;
  if (ruvnb_create(&ctx, Rf_asInteger(device)) != RUVNB_OK) Rf_error("ruvnb: no usable CUDA device (there is no CPU fallback)");
  SEXP ext = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ext, ctx_finalizer, TRUE);
  unsigned char* c8 = (unsigned char*)R_alloc(G, 1);
  for (int g = 0; g < G; ++g) c8[g] = LOGICAL(ctl)[g] == 1;
  unsigned char* z8 = NULL;
  if (zeroinf != R_NilValue) {
    z8 = (unsigned char*)R_alloc(N, 1);
    for (int c = 0; c < N; ++c) z8[c] = LOGICAL(zeroinf)[c] == 1;
  }
  if (batch != R_NilValue) {
    if (Rf_length(batch) != N) Rf_error("ruvnb: batch must have one entry per cell");
    int nb = 0;
    for (int c = 0; c < N; ++c) if (INTEGER(batch)[c] + 1 > nb) nb = INTEGER(batch)[c] + 1;
    CHECK(ctx, ruvnb_set_batches(ctx, INTEGER(batch), N, nb));
  }
  CHECK(ctx, ruvnb_set_design(ctx, G, N, Rf_asInteger(m_), INTEGER(grp), c8, z8, 0, G));
  *out = ctx;
  UNPROTECT(1);
  return ext;
}
/* Winsorize the resident counts (R/ruvIIInb.R:53) unless `capped`; attribute "nonzero": rowSums(Y > 0) afterwards (:57-62) */
static SEXP finish_open(ruvnb_ctx* ctx, SEXP ext, int G, int capped) {
  PROTECT(ext);
  int64_t* nz = (int64_t*)R_alloc(G, sizeof(int64_t));
  if (!capped) CHECK(ctx, ruvnb_winsorize(ctx, NULL, nz));
  SEXP nzv = PROTECT(Rf_allocVector(REALSXP, G));
  for (int g = 0; g < G; ++g) REAL(nzv)[g] = capped ? NA_REAL : (double)nz[g];
  Rf_setAttrib(ext, Rf_install("nonzero"), nzv);
  UNPROTECT(2);
  return ext;
}

/* ruvnb_R_open(Y integer G x N, grp integer N (0-based), m, ctl logical G, zeroinf logical N or NULL, batch integer N (0-based) or
 * NULL, capped logical, device) -> external pointer owning the context with the counts resident and (unless capped) Winsorized
 * (R/ruvIIInb.R:34-62) */
SEXP ruvnb_R_open(SEXP Y, SEXP grp, SEXP m_, SEXP ctl, SEXP zeroinf, SEXP batch, SEXP capped, SEXP device) {
  if (!Rf_isInteger(Y) || !Rf_isMatrix(Y)) Rf_error("ruvnb: Y must be an integer matrix (genes x cells)");
  const int G = Rf_nrows(Y), N = Rf_ncols(Y);
  ruvnb_ctx* ctx = NULL;
  SEXP ext = PROTECT(open_design(G, N, grp, m_, ctl, zeroinf, batch, device, &ctx));
  CHECK(ctx, ruvnb_upload_counts_dense(ctx, INTEGER(Y), RUVNB_LAYOUT_CELL_MAJOR, NULL));
  ext = finish_open(ctx, ext, G, Rf_asLogical(capped) == 1);
  UNPROTECT(1);
  return ext;
}
/* The same from the slots of a dgCMatrix (`Y@p`, `Y@i`, `Y@x`, dim(Y)): the transpose into the resident gene-major layout happens
 * on the device, nothing dense is built on the host (the `counts` of a fit are a dgCMatrix, R/ruvIIInb.R:611-612) */
SEXP ruvnb_R_open_csc(SEXP p, SEXP i, SEXP x, SEXP dims, SEXP grp, SEXP m_, SEXP ctl, SEXP zeroinf, SEXP batch, SEXP capped, SEXP device) {
  const int G = INTEGER(dims)[0], N = INTEGER(dims)[1];
  if (!Rf_isInteger(p) || !Rf_isInteger(i) || !Rf_isReal(x) || Rf_length(p) != N + 1) Rf_error("ruvnb: p / i / x are not the slots of a dgCMatrix of these dimensions");
  ruvnb_ctx* ctx = NULL;
  SEXP ext = PROTECT(open_design(G, N, grp, m_, ctl, zeroinf, batch, device, &ctx));
  CHECK(ctx, ruvnb_upload_counts_csc(ctx, INTEGER(p), INTEGER(i), REAL(x), 0, N));
  CHECK(ctx, ruvnb_upload_counts_done(ctx));
  ext = finish_open(ctx, ext, G, Rf_asLogical(capped) == 1);
  UNPROTECT(1);
  return ext;
}

/* ruvnb_R_counts(ctx) -> the resident (Winsorized) counts, integer G x N: the `counts` element (R/ruvIIInb.R:611,627) */
SEXP ruvnb_R_counts(SEXP ext, SEXP G_, SEXP N_) {
  ruvnb_ctx* ctx = ctx_of(ext);
  const int G = Rf_asInteger(G_), N = Rf_asInteger(N_);
  int32_t* buf = (int32_t*)R_alloc((size_t)G * N, 4);
  CHECK(ctx, ruvnb_get_counts(ctx, buf));                                       /* gene-major */
  SEXP out = PROTECT(Rf_allocMatrix(INTSXP, G, N));
  int* o = INTEGER(out);
  for (int g = 0; g < G; ++g) for (int c = 0; c < N; ++c) o[g + (size_t)G * c] = buf[(size_t)g * N + c];
  UNPROTECT(1);
  return out;
}

/* ruvnb_R_get_fit(ctx, G, N, m, k) -> list(W, a, Mb, gmean, psi (G, or G x nbatch), pi0, psi.w): the current parameter set */
SEXP ruvnb_R_get_fit(SEXP ext, SEXP G_, SEXP N_, SEXP m_, SEXP k_) {
  ruvnb_ctx* ctx = ctx_of(ext);
  const int G = Rf_asInteger(G_), N = Rf_asInteger(N_), m = Rf_asInteger(m_), k = Rf_asInteger(k_), nb = ruvnb_nbatch(ctx);
  SEXP W = PROTECT(Rf_allocMatrix(REALSXP, N, k)), a = PROTECT(Rf_allocMatrix(REALSXP, G, k));
  SEXP Mb = PROTECT(Rf_allocMatrix(REALSXP, G, m)), gm = PROTECT(Rf_allocVector(REALSXP, G));
  SEXP ps = PROTECT(nb > 1 ? Rf_allocMatrix(REALSXP, G, nb) : Rf_allocVector(REALSXP, G)), p0 = PROTECT(Rf_allocVector(REALSXP, G));
  SEXP pw = PROTECT(nb > 1 ? Rf_allocMatrix(REALSXP, G, nb) : Rf_allocVector(REALSXP, G));
  CHECK(ctx, ruvnb_get_state(ctx, "W", REAL(W)));   CHECK(ctx, ruvnb_get_state(ctx, "alpha", REAL(a)));
  CHECK(ctx, ruvnb_get_state(ctx, "Mb", REAL(Mb))); CHECK(ctx, ruvnb_get_state(ctx, "gmean", REAL(gm)));
  CHECK(ctx, ruvnb_get_state(ctx, "psi", REAL(ps))); CHECK(ctx, ruvnb_get_state(ctx, "pi0", REAL(p0)));
  CHECK(ctx, ruvnb_get_state(ctx, "psi_w", REAL(pw)));
  const char* nm[] = {"W", "a", "Mb", "gmean", "psi", "pi0", "psi.w", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, W); SET_VECTOR_ELT(out, 1, a); SET_VECTOR_ELT(out, 2, Mb); SET_VECTOR_ELT(out, 3, gm);
  SET_VECTOR_ELT(out, 4, ps); SET_VECTOR_ELT(out, 5, p0); SET_VECTOR_ELT(out, 6, pw);
  UNPROTECT(8);
  return out;
}

/* the dispersion hook of ruvnb_fit_nb (ruvnb_set_psi_callback) as an R closure: hook(ctx) reads the state with ruvnb_R_get_fit and
 * writes the new dispersion with ruvnb_R_set_fit(psi = ) -- one estimateDisp.par per batch lives there (R/ruvIIInb.R:213-220,542-548) */
struct hook_call { SEXP fn, ext; };
static int psi_hook_trampoline(void* user, ruvnb_ctx* ctx) {
  struct hook_call* h = (struct hook_call*)user;
  int err = 0;
  SEXP call = PROTECT(Rf_lang2(h->fn, h->ext));
  R_tryEval(call, R_GlobalEnv, &err);
  UNPROTECT(1);
  (void)ctx;
  return err ? RUVNB_EARG : RUVNB_OK;
}

/* ruvnb_R_fit(ctx, G, N, m, k, W0 (N x k or NULL: device SVD, R/ruvIIInb.R:160-168), psi (G or G x nbatch; NULL: estimateDisp.par
 * at every refresh), psi.hook (function(ctx) or NULL), opts list) -> list(W, a, Mb, gmean, psi, pi0, psi.w, logl): the inner / outer
 * loops of R/ruvIIInb.R:266-609 */
SEXP ruvnb_R_fit(SEXP ext, SEXP G_, SEXP N_, SEXP m_, SEXP k_, SEXP W0, SEXP psi, SEXP hook, SEXP opts_) {
  ruvnb_ctx* ctx = ctx_of(ext);
  const int k = Rf_asInteger(k_);
  ruvnb_opts o;
  opts_from_list(&o, opts_, k);
  if (W0 == R_NilValue) CHECK(ctx, ruvnb_init_W(ctx, k, 0, 0.0, NULL, NULL, NULL));
  ruvnb_fit_stats st;
  double* logl = (double*)R_alloc(o.outer_maxit > 0 ? o.outer_maxit : 1, sizeof(double));
  struct hook_call hc; hc.fn = hook; hc.ext = ext;
  CHECK(ctx, ruvnb_set_psi_callback(ctx, hook == R_NilValue ? NULL : psi_hook_trampoline, &hc));
  int rc = ruvnb_fit_nb(ctx, &o, W0 == R_NilValue ? NULL : REAL(W0), psi == R_NilValue ? NULL : REAL(psi), psi == R_NilValue ? 0 : 1,
                        NULL, 0, logl, &st);
  ruvnb_set_psi_callback(ctx, NULL, NULL);
  if (rc != RUVNB_OK) Rf_error("ruvnb: %s", ruvnb_last_error(ctx));
  SEXP fit = PROTECT(ruvnb_R_get_fit(ext, G_, N_, m_, k_));
  SEXP ll = PROTECT(Rf_allocVector(REALSXP, st.n_outer));
  for (int i = 0; i < st.n_outer; ++i) REAL(ll)[i] = logl[i];
  const char* nm[] = {"W", "a", "Mb", "gmean", "psi", "pi0", "psi.w", "logl", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  for (int i = 0; i < 7; ++i) SET_VECTOR_ELT(out, i, VECTOR_ELT(fit, i));
  SET_VECTOR_ELT(out, 7, ll);
  UNPROTECT(3);
  return out;
}

/* ruvnb_R_set_fit(ctx, W, a, gmean, Mb (G x m), psi): the fitted parameters of an `out` list put back on a context (get.res,
 * estimateDisp.par called on their own) */
SEXP ruvnb_R_set_fit(SEXP ext, SEXP W, SEXP a, SEXP gmean, SEXP Mb, SEXP psi) {
  ruvnb_ctx* ctx = ctx_of(ext);
  if (W != R_NilValue) CHECK(ctx, ruvnb_set_state(ctx, "W", REAL(W), Rf_ncols(W)));
  if (a != R_NilValue) CHECK(ctx, ruvnb_set_state(ctx, "alpha", REAL(a), 0));
  if (gmean != R_NilValue) CHECK(ctx, ruvnb_set_state(ctx, "gmean", REAL(gmean), 0));
  if (Mb != R_NilValue) CHECK(ctx, ruvnb_set_state(ctx, "Mb", REAL(Mb), 0));
  if (psi != R_NilValue) CHECK(ctx, ruvnb_set_state(ctx, "psi", REAL(psi), 0));
  return R_NilValue;
}

/* ruvnb_R_get_res(ctx, G, N, type (1 pearson, 2 logcounts, 3 quantile), block.size, Mb.all (G x N or NULL)) -> G x N matrix:
 * what get.res writes block by block into its sink (R/get.res.R:27-128) */
SEXP ruvnb_R_get_res(SEXP ext, SEXP G_, SEXP N_, SEXP type, SEXP block_size, SEXP Mb_all) {
  ruvnb_ctx* ctx = ctx_of(ext);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(G_), Rf_asInteger(N_)));
  CHECK(ctx, ruvnb_get_res(ctx, Rf_asInteger(type), (int64_t)Rf_asInteger(block_size), Mb_all == R_NilValue ? NULL : REAL(Mb_all), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* ruvnb_R_set_zinb(ctx, pi0 (G), psi.w (G or G x nbatch, or NULL), zflag logical N or NULL, mean.zeroinf): the zero-inflation part
 * of a fit; zflag / mean.zeroinf switch get.res to the ZIM branches (R/get.res.R:52-57,69-70,84-85,95-96; include/ruvnb.h) */
SEXP ruvnb_R_set_zinb(SEXP ext, SEXP pi0, SEXP psi_w, SEXP zflag, SEXP mean_zeroinf) {
  ruvnb_ctx* ctx = ctx_of(ext);
  CHECK(ctx, ruvnb_set_state(ctx, "pi0", REAL(pi0), 0));
  if (psi_w != R_NilValue) CHECK(ctx, ruvnb_set_state(ctx, "psi_w", REAL(psi_w), 0));
  if (zflag != R_NilValue) {
    const int N = Rf_length(zflag);
    unsigned char* z8 = (unsigned char*)R_alloc(N, 1);
    for (int c = 0; c < N; ++c) z8[c] = LOGICAL(zflag)[c] == 1;
    CHECK(ctx, ruvnb_set_res_zinb(ctx, z8, Rf_asReal(mean_zeroinf)));
  }
  return R_NilValue;
}

/* ruvnb_R_get_res_csc(ctx, G, N, type, block.size, Mb.all or NULL, log1p) -> list(p, i, x): the slots of the dgCMatrix makeSCE stores
 * (`as(..., "sparseMatrix")`, R/makeSCE.R:24-38); every block is compacted on the device.  The sink grows by doubling. */
SEXP ruvnb_R_get_res_csc(SEXP ext, SEXP G_, SEXP N_, SEXP type, SEXP block_size, SEXP Mb_all, SEXP log1p_) {
  ruvnb_ctx* ctx = ctx_of(ext);
  const int G = Rf_asInteger(G_), N = Rf_asInteger(N_);
  int64_t* cp = (int64_t*)R_alloc((size_t)N + 1, sizeof(int64_t));
  int64_t cap = (int64_t)G * (N < 64 ? N : 64), nnz = 0;
  for (;;) {
    int32_t* ri = (int32_t*)R_alloc((size_t)cap, sizeof(int32_t));
    double* vv = (double*)R_alloc((size_t)cap, sizeof(double));
    int rc = ruvnb_get_res_csc(ctx, Rf_asInteger(type), (int64_t)Rf_asInteger(block_size), Mb_all == R_NilValue ? NULL : REAL(Mb_all),
                               Rf_asLogical(log1p_) == 1, 0, N, cp, ri, vv, cap, &nnz, NULL);
    if (rc == RUVNB_ENOMEM && cap < (int64_t)G * N) { cap = cap * 4 < (int64_t)G * N ? cap * 4 : (int64_t)G * N; continue; }
    if (rc != RUVNB_OK) Rf_error("ruvnb: %s", ruvnb_last_error(ctx));
    if (nnz > 2147483647) Rf_error("ruvnb: more than 2^31 - 1 non-zero entries do not fit a dgCMatrix");
    SEXP p = PROTECT(Rf_allocVector(INTSXP, N + 1)), i = PROTECT(Rf_allocVector(INTSXP, nnz)), x = PROTECT(Rf_allocVector(REALSXP, nnz));
    for (int c = 0; c <= N; ++c) INTEGER(p)[c] = (int)cp[c];
    if (nnz) { memcpy(INTEGER(i), ri, (size_t)nnz * sizeof(int32_t)); memcpy(REAL(x), vv, (size_t)nnz * sizeof(double)); }
    const char* nm[] = {"p", "i", "x", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(out, 0, p); SET_VECTOR_ELT(out, 1, i); SET_VECTOR_ELT(out, 2, x);
    UNPROTECT(4);
    return out;
  }
}

/* ruvnb_R_estimate_disp(ctx, G, robust, prior.df or NULL, opts) -> list(common.dispersion, trended.dispersion, tagwise.dispersion,
 * prior.df): estimateDisp.par for design = 1 and the offset the context's parameters define (R/estimateDisp.par.R:27-201) */
SEXP ruvnb_R_estimate_disp(SEXP ext, SEXP G_, SEXP k_, SEXP robust, SEXP prior_df, SEXP opts_) {
  ruvnb_ctx* ctx = ctx_of(ext);
  const int G = Rf_asInteger(G_);
  ruvnb_opts o;
  opts_from_list(&o, opts_, Rf_asInteger(k_));
  double common = 0.0;
  SEXP tr = PROTECT(Rf_allocVector(REALSXP, G)), pd = PROTECT(Rf_allocVector(REALSXP, G)), tag = PROTECT(Rf_allocVector(REALSXP, G));
  CHECK(ctx, ruvnb_estimate_disp_r(ctx, &o, Rf_asLogical(robust) == 1, prior_df == R_NilValue ? -1.0 : Rf_asReal(prior_df), &common,
                                   REAL(tr), REAL(pd)));
  CHECK(ctx, ruvnb_get_state(ctx, "psi", REAL(tag)));
  const char* nm[] = {"common.dispersion", "trended.dispersion", "tagwise.dispersion", "prior.df", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, Rf_ScalarReal(common)); SET_VECTOR_ELT(out, 1, tr); SET_VECTOR_ELT(out, 2, tag); SET_VECTOR_ELT(out, 3, pd);
  UNPROTECT(4);
  return out;
}

/* the full-data passes of fastruvIII.nb on a context that holds ALL cells (R/fastruvIIInb.R:632-748,790) */
SEXP ruvnb_R_fast_W_all(SEXP ext, SEXP k_, SEXP lambda_a, SEXP wt_ctl, SEXP W_init, SEXP W_med, SEXP W_mad) {
  ruvnb_ctx* ctx = ctx_of(ext);
  SEXP W = PROTECT(Rf_allocMatrix(REALSXP, Rf_nrows(W_init), Rf_ncols(W_init)));
  int sweeps = 0; double crit = 0.0;
  CHECK(ctx, ruvnb_fast_W_all(ctx, Rf_asInteger(k_), Rf_asReal(lambda_a), REAL(wt_ctl), REAL(W_init), REAL(W_med), REAL(W_mad), 8, 1e-5,
                              REAL(W), &sweeps, &crit));
  UNPROTECT(1);
  return W;
}
SEXP ruvnb_R_fast_res(SEXP ext, SEXP G_, SEXP N_, SEXP type, SEXP lambda_b, SEXP annot, SEXP block_size, SEXP log1p_) {
  ruvnb_ctx* ctx = ctx_of(ext);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(G_), Rf_asInteger(N_)));
  CHECK(ctx, ruvnb_fast_res(ctx, Rf_asInteger(type), Rf_asReal(lambda_b), INTEGER(annot), (int64_t)Rf_asInteger(block_size),
                            Rf_asLogical(log1p_) == 1 ? RUVNB_OUT_LOG1P : RUVNB_OUT_F64, REAL(out), NULL, 0, -1, NULL));
  UNPROTECT(1);
  return out;
}

SEXP ruvnb_R_close(SEXP ext) { ctx_finalizer(ext); return R_NilValue; }

static const R_CallMethodDef call_methods[] = {
  {"ruvnb_R_open", (DL_FUNC)&ruvnb_R_open, 8},           {"ruvnb_R_counts", (DL_FUNC)&ruvnb_R_counts, 3},
  {"ruvnb_R_open_csc", (DL_FUNC)&ruvnb_R_open_csc, 11},  {"ruvnb_R_get_fit", (DL_FUNC)&ruvnb_R_get_fit, 5},
  {"ruvnb_R_set_zinb", (DL_FUNC)&ruvnb_R_set_zinb, 5},   {"ruvnb_R_get_res_csc", (DL_FUNC)&ruvnb_R_get_res_csc, 7},
  {"ruvnb_R_fit", (DL_FUNC)&ruvnb_R_fit, 9},             {"ruvnb_R_set_fit", (DL_FUNC)&ruvnb_R_set_fit, 6},
  {"ruvnb_R_get_res", (DL_FUNC)&ruvnb_R_get_res, 6},     {"ruvnb_R_estimate_disp", (DL_FUNC)&ruvnb_R_estimate_disp, 6},
  {"ruvnb_R_fast_W_all", (DL_FUNC)&ruvnb_R_fast_W_all, 7}, {"ruvnb_R_fast_res", (DL_FUNC)&ruvnb_R_fast_res, 8},
  {"ruvnb_R_close", (DL_FUNC)&ruvnb_R_close, 1},         {NULL, NULL, 0}};

void R_init_ruvIIInb(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
