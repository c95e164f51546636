/*
 * ruvnb.h -- C ABI of libruvnb_b200: the B200-native RUV-III-NB alternating-IRLS path.
 *
 * The reference (limfuxing/ruvIIInb 0.9.2.1) is a pure-R package with NO native interface
 * (DESCRIPTION:35 `NeedsCompilation: no`, NAMESPACE:1-6).  Its data-parallel seams are the
 * per-gene / per-cell helper closures it dispatches through BiocParallel::bpmapply and foreach
 * (R/ruvIIInb.R:183,239,350,356,385,439; R/estimateDisp.par.R:84-113).  Every entry point below
 * replaces one of those seams and cites it; INTEGRATION.md shows the `.Call` shim that binds
 * them from R and the ctypes binding used by the Python host mirror.
 *
 * Conventions
 *   - Every function returns 0 on success, a negative RUVNB_E* code on failure;
 *     ruvnb_last_error(ctx) returns the message.  Numerical "problematic entries" (NaN/Inf in
 *     an update) are NOT errors: the reference swallows and reverts them (R/ruvIIInb.R:368,402,
 *     424-425) and so do we, reporting the counts.
 *   - All matrices crossing the boundary are HOST arrays, double precision, column-major,
 *     shaped exactly like the R objects (W: N x k, a: G x k, Mb: G x m, gmean/psi: G).
 *     Counts are int32.  The caller owns host buffers; the library owns all device memory.
 *   - One ctx == one GPU == one host thread.  Multi-GPU = one process (ctx) per GPU with genes
 *     sharded across ranks; ruvnb_comm_init wires the ranks with NCCL.
 *   - There is no CPU fallback: ruvnb_create fails if no CUDA device is usable.
 */
#ifndef RUVNB_H
#define RUVNB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RUVNB_ABI_VERSION 3

#define RUVNB_OK 0
#define RUVNB_EARG (-1)    /* invalid argument */
#define RUVNB_ECUDA (-2)   /* CUDA runtime error */
#define RUVNB_ENCCL (-3)   /* NCCL error */
#define RUVNB_ESTATE (-4)  /* call out of order (e.g. fit before upload) */
#define RUVNB_ENOMEM (-5)  /* problem does not fit the device */

#define RUVNB_MAX_K 8      /* unwanted factors supported by the register-resident Gram kernels */

typedef struct ruvnb_ctx ruvnb_ctx;

/* Layout of the count matrix handed to ruvnb_upload_counts_dense. */
#define RUVNB_LAYOUT_GENE_MAJOR 0 /* Y[g*N + c]  (NumPy C order G x N)        */
#define RUVNB_LAYOUT_CELL_MAJOR 1 /* Y[g + G*c]  (R integer matrix, col-major) */

/* Constants the reference hard-codes or takes as arguments (R/ruvIIInb.R:27-28,155,321,374,431,
 * 506,524,533,597).  ruvnb_opts_default fills the reference defaults. */
typedef struct ruvnb_opts {
  int k;               /* number of unwanted factors (ruvIII.nb `k`)                 */
  int robust;          /* `robust`: Huber k = 1.345 if set, else 100 (:155)            */
  double lambda_a;     /* `lambda.a` ridge on alpha (:729)                            */
  double lambda_b;     /* `lambda.b` ridge on Mb (:420,669)                           */
  double step_fac;     /* `step.fac` (:491)                                           */
  int inner_maxit;     /* `inner.maxit` (:526)                                        */
  int outer_maxit;     /* `outer.maxit` (:597)                                        */
  double inner_tol;    /* 1e-4 (:524)                                                 */
  double outer_tol;    /* 1e-4 (:597)                                                 */
  double psi_tol;      /* 1e-8 (:533)                                                 */
  int zinb;            /* fit ZINB for cells flagged zeroinf (:237-264,435-467)        */
  int huber_fastpath;  /* 1: skip the exact per-gene/per-cell MAD when a counting      */
                       /*    certificate proves every Huber weight is exactly 1;       */
                       /* 0: always compute the exact MAD (results are identical)      */
  int verbose;         /* print the reference's trace lines (:301,514,594) to stderr   */
} ruvnb_opts;

/* Per-inner-iteration record: the decisions of R/ruvIIInb.R:486-527. */
typedef struct ruvnb_trace_rec {
  int outer, iter, halving;
  int degener, accepted;
  double logl_old, logl_new, logl;
  int bad_alpha, bad_W, bad_Mb; /* "Problematic ... entries" counts (:367,401,423) */
} ruvnb_trace_rec;

typedef struct ruvnb_fit_stats {
  int n_outer;              /* outer iterations run                                       */
  int n_inner;              /* inner iterations run (accepted + rejected)                 */
  int n_trace;              /* records written to `trace` (<= trace_cap)                  */
  double inner_seconds;     /* device time in the inner IRLS loops (CUDA events)          */
  double disp_seconds;      /* device+host time in the dispersion step                    */
  double total_seconds;     /* whole ruvnb_fit_nb call                                    */
  long long kernel_launches;/* kernels launched by this library during the call           */
  int n_exact_mad_rows;     /* gene rows that needed the exact MAD (certificate failed)   */
  int n_exact_mad_cells;    /* cells that needed the exact MAD                            */
} ruvnb_fit_stats;

/* ---- life cycle ------------------------------------------------------------------------- */
int ruvnb_abi_version(void);
void ruvnb_opts_default(ruvnb_opts* o);
/* device: CUDA ordinal.  Replaces doParallel::registerDoParallel(ncores) (R/ruvIIInb.R:34-37):
 * the worker pool becomes one GPU. */
int ruvnb_create(ruvnb_ctx** out, int device);
void ruvnb_destroy(ruvnb_ctx* ctx);
const char* ruvnb_last_error(const ruvnb_ctx* ctx);
long long ruvnb_kernel_launches(const ruvnb_ctx* ctx);
/* Device time of the Y-streaming sweep kernels, measured with CUDA events on the library's own stream
 * (bench.py's roofline line).  which: 0 unused (the Huber certificate now rides in the log-likelihood sweep),
 * 1 pass1 (Gram/score), 2 w_update (per-cell), 3 pass2 (group sums), 4 loglik + certificate (k_ll_w when fused), 5 exact-MAD (signdev + select),
 * 6 gram (Gram / score / group sums from the slabs of the fused sweep); stream segments between the sweeps (small kernels +
 * collectives, what limits a many-GPU iteration): 7 alpha phase (normalise, all-reduce, solve, all-gather, clamp), 8 W exchange
 * (all-reduce of the slices, column norms), 9 gmean / Mb finish, 10 tail (certificate finish, log-likelihood total, all-reduce).
 * Accumulated since ruvnb_kernel_time_reset. */
#define RUVNB_NKERN 11
int ruvnb_kernel_time(const ruvnb_ctx* ctx, int which, double* ms_total, long long* launches);
void ruvnb_kernel_time_reset(ruvnb_ctx* ctx);
/* The CUDA stream every kernel of this context is launched on (a cudaStream_t), so a host can
 * bracket calls with its own events. */
void* ruvnb_stream(const ruvnb_ctx* ctx);
/* Gene rows / cells of the last ruvnb_irls_iteration whose Huber certificate failed (exact MAD computed). */
int ruvnb_last_exact_rows(const ruvnb_ctx* ctx);
int ruvnb_last_exact_cells(const ruvnb_ctx* ctx);
/* Why gene rows failed the Huber certificate since the last call (tuning aid; resets the counters): [0] no usable
 * hint, [1] NaN deviance, [2] max|d| = 0, [3] max|d| outgrew the hint, [4]/[5] the median left the hint window
 * below/above, [6] more than half of the row inside the window, [7] fast path off. */
int ruvnb_cert_diag(ruvnb_ctx* ctx, int out[8]);
/* Measured FP64 FMA throughput of the context's device in TFLOP/s (2 flops per DFMA; a register-only DFMA-chain
 * kernel, best of 5): the denominator of the FP64 roofline bench.py reports beside the HBM one. */
int ruvnb_measure_fp64(ruvnb_ctx* ctx, double* tflops);
/* Host evaluation of the table-driven exp (which = 0) / log (which = 1) / unchecked log on the interleaved table
 * (which = 2, normal positive arguments) the kernels use (csrc/fastmath.cuh is __host__ __device__): lets the CPU test-suite check the very source the GPU runs.  No device needed. */
int ruvnb_selftest_math(int which, const double* x, double* out, int64_t n);
/* Host evaluation of the prior-d.f. step of the dispersion estimator (csrc/limma_host.h: limma::squeezeVar(s2, df = df1,
 * covariate, robust, winsor.tail.p = c(0.05, 0.1))$df.prior as R/estimateDisp.par.R:166-168 calls it): df_prior_out receives n
 * values (all equal when robust = 0).  covariate may be NULL.  No device needed: the CPU test-suite checks it against the
 * SciPy-based restatement. */
int ruvnb_selftest_squeeze_var(const double* s2, int64_t n, double df1, const double* covariate, int robust, double* df_prior_out);

/* ---- multi-GPU wiring (gene sharding) -------------------------------- */
/* id: 128 bytes from ruvnb_comm_unique_id on rank 0, distributed by the host (torch.distributed
 * broadcast in the Python mirror; Rmpi or a file in an R host).  After this call the ctx owns genes
 * [g0, g1) of the G passed to ruvnb_set_design and all-reduces/all-gathers the small per-iteration
 * quantities over NCCL. */
int ruvnb_comm_unique_id(unsigned char id[128]);
int ruvnb_comm_init(ruvnb_ctx* ctx, const unsigned char id[128], int rank, int nranks);

/* ---- problem set-up ------------------------------------------------------------------------- */
/* grp_of_cell: N labels in 0..m-1 (apply(M,1,which), R/ruvIIInb.R:50); ctl: G flags (:42-45);
 * zeroinf: N flags or NULL (:39-40).  G is the GLOBAL gene count; [g0,g1) the rows this ctx holds
 * (g0 = 0, g1 = G on a single GPU). */
/* Batch of every cell (0-based, N entries, nbatch <= 127) for a BATCH-SPECIFIC dispersion psi[g, batch(c)]
 * (`batch.disp = TRUE`, R/ruvIIInb.R:213-220,278-283,306-315,542-548; R/get.res.R:59-62,101-121).  Must precede ruvnb_set_design: the
 * resident cell order keeps the batches of a replicate group apart.  State "psi" / "psi_w" then have G x nbatch entries (column-major,
 * R's psi matrix); ruvnb_irls_iteration, ruvnb_loglik, ruvnb_fit_nb, ruvnb_zinb_pi0 and ruvnb_get_res index the dispersion by the
 * cell's batch.  The dispersion estimators (ruvnb_estimate_disp*, ruvnb_disp_apl_grid, ruvnb_disp_newton) and the fastruvIII.nb
 * operators refuse such a context: the reference calls estimateDisp once per batch on that batch's cells (:213-220), which a host
 * does with one context per batch (ruvnb_set_psi_callback). */
int ruvnb_set_batches(ruvnb_ctx* ctx, const int32_t* batch_of_cell, int64_t N, int nbatch);
int ruvnb_nbatch(const ruvnb_ctx* ctx);
/* Host hook for every dispersion refresh of ruvnb_fit_nb (R/ruvIIInb.R:205-221,535-565) in place of the built-in estimator: called
 * with the state to estimate from as the current state; it reads with ruvnb_get_state and writes the new dispersion with
 * ruvnb_set_state(ctx, "psi", ...) -- keeping the old value where the new one is not finite (:564) is the hook's job.  Returns
 * RUVNB_OK or a negative error.  fn = NULL removes the hook.  A psi schedule passed to ruvnb_fit_nb takes precedence. */
typedef int (*ruvnb_psi_fn)(void* user, ruvnb_ctx* ctx);
int ruvnb_set_psi_callback(ruvnb_ctx* ctx, ruvnb_psi_fn fn, void* user);
int ruvnb_set_design(ruvnb_ctx* ctx, int64_t G, int64_t N, int m, const int32_t* grp_of_cell,
                     const uint8_t* ctl, const uint8_t* zeroinf, int64_t g0, int64_t g1);
/* Y: the (g1-g0) x N rows this ctx owns, plus -- when sharded -- Yctl: the n_ctl x N control-gene
 * rows (replicated on every rank for the per-cell W update; NULL on a single GPU). */
/* (gene shards: Yctl = NULL after ruvnb_comm_init makes every rank copy the control rows it owns into the replicated slab and
 * broadcast them over NCCL -- nothing but the rank's own rows crosses PCIe; the ranks' row ranges must partition 0..G) */
int ruvnb_upload_counts_dense(ruvnb_ctx* ctx, const int32_t* Y, int layout, const int32_t* Yctl);
/* CSR rows (the sparse / ZINB input, BASELINE.json config 3): indptr (rows+1), indices, values. */
int ruvnb_upload_counts_csr(ruvnb_ctx* ctx, const int64_t* indptr, const int32_t* indices,
                            const int32_t* values, const int64_t* ctl_indptr,
                            const int32_t* ctl_indices, const int32_t* ctl_values);

/* H0  Winsorize + ceiling (R/ruvIIInb.R:53, DescTools::Winsorize probs=c(0,.995), type-7):
 * caps the resident counts in place; cap_out (local rows, may be NULL) receives ceil(Q_.995);
 * nonzero_out (local rows, may be NULL) receives rowSums(Y>0) for the zero-gene filter (:57). */
int ruvnb_winsorize(ruvnb_ctx* ctx, double* cap_out, int64_t* nonzero_out);
/* The resident counts of the local rows (after ruvnb_winsorize: the winsorized ones), gene-major Gl x N int32:
 * the `counts` element of the reference's return list (R/ruvIIInb.R:611,627). */
int ruvnb_get_counts(ruvnb_ctx* ctx, int32_t* Y_out);

/* ---- state (all local gene rows; column-major like the R objects) ---------------------------- */
/* names: "W"(N x k) "alpha"(G x k) "alpha_dev"(G x k) "gmean"(G) "Mb"(G x m) "psi"(G) "step"(G)
 *        "pi0"(G) "loglik"(G); G here = local rows. */
int ruvnb_set_state(ruvnb_ctx* ctx, const char* name, const double* host, int k);
int ruvnb_get_state(ruvnb_ctx* ctx, const char* name, double* host);

/* H1  initial W (R/ruvIIInb.R:160-168): the k leading right singular vectors of log(Y[ctl,]/rowMeans(Y[ctl,]) + 1), by
 *     subspace iteration on the resident control rows (the matrix is never formed) + one Rayleigh-Ritz rotation; each
 *     vector signed so that its largest entry is positive.  The reference uses irlba (random start, tol 1e-5), so its own
 *     W0 is defined only up to that: parity runs inject W0 instead.  Sets "W"; W_out (N x k), sv_out (k singular values)
 *     and iters_out may be NULL.  max_iter <= 0 -> 100, tol <= 0 -> 1e-9 (relative change of the Ritz values, ~3e-5 on the vectors). */
int ruvnb_init_W(ruvnb_ctx* ctx, int k, int max_iter, double tol, double* W_out, double* sv_out, int* iters_out);

/* ---- step operators: one per reference helper seam ------------------------------------------ */
/* H2  get.coef.wt (R/ruvIIInb.R:181-198,704-706): WLS of log(Y+1) on [1,W] with weights Y+1;
 *     sets gmean, alpha, alpha_dev, Mb = 0. */
int ruvnb_init_coef(ruvnb_ctx* ctx, const ruvnb_opts* o);
/* H6  per-gene log-likelihood at the current state (R/ruvIIInb.R:276-291): sets "loglik",
 *     returns the sum over ALL genes (all ranks) in *total. */
int ruvnb_loglik(ruvnb_ctx* ctx, const ruvnb_opts* o, double* total);
/* H7-H13  one pass of the inner-loop body (R/ruvIIInb.R:303-484): alpha (get.coef2 + get.adev +
 *     clamp), W (getW + normalise), gmean, Mb (+ clamp), (ZINB: pi0 + E-step), then the candidate
 *     log-likelihood.  The candidate becomes the current state; `bad` receives the three
 *     problematic-entry counts; *logl_new the candidate total. */
int ruvnb_irls_iteration(ruvnb_ctx* ctx, const ruvnb_opts* o, double* logl_new, int bad[3]);
/* H4/H5  ZINB (opts.zinb and cells flagged zeroinf): pi0 of the current parameter set by optim()'s one-dimensional
 *     Nelder-Mead from p = -2 (R/ruvIIInb.R:237-251,770-772), then the E-step -- the weights wt.zinb (:254-264) are
 *     never stored: every sweep recomputes them from (pi0, the psi of this call).  bad: non-finite pi0 entries (:453). */
int ruvnb_zinb_pi0(ruvnb_ctx* ctx, const ruvnb_opts* o, int* bad);
/* H14 snapshots: best.* <- current (R/ruvIIInb.R:296,516) / current <- best.* (:492). */
int ruvnb_snapshot_best(ruvnb_ctx* ctx);
int ruvnb_restore_best(ruvnb_ctx* ctx);
/* H14 per-gene step halving: step[g] *= step_fac where loglik_old[g] > loglik_new[g] (:490-491). */
int ruvnb_halve_steps(ruvnb_ctx* ctx, const ruvnb_opts* o);
/* H14 accept: loglik <- loglik.tmp (:512); the candidate of the last ruvnb_irls_iteration is already
 * the current state. */
int ruvnb_accept(ruvnb_ctx* ctx);

/* H3  estimateDisp.par (R/estimateDisp.par.R:3-202) with design = 1, offset = alpha W' + Mb[,grp],
 *     weights = wt.zinb (call sites R/ruvIIInb.R:205-221,535-565).
 *     ruvnb_disp_apl_grid: the Cox-Reid adjusted profile likelihood on the 21-point grid
 *     0.1*2^seq(-10,10) (:33-37,84-127) -> l0 (local rows x 21, column-major; rows with
 *     rowSums(y) < 5 are NaN), and AveLogCPM-ready intercepts.
 *     ruvnb_estimate_disp: grid + common + trend + prior + WLEB -> sets "psi" (tagwise).
 *     The one-group fits stop at edgeR's tolerance (|step| < 1e-10, at most 50 sweeps) or one sweep earlier when the
 *     next step is predicted below it; start values differ from edgeR's (neighbouring grid point, earlier call) but the
 *     root of the concave score is the same.  Device memory: while free memory allows (need + max(8 GB, 1/8 of the
 *     device)), the context keeps exp(offset) of every (local gene, cell) for the duration of the step (8 B per
 *     element, allocated at the first call); the environment variable RUVNB_NO_ECACHE=1 turns that off.
 *     RUVNB_TRACE_DISP=1 prints every sweep's time and converged rows on stderr. */
int ruvnb_disp_apl_grid(ruvnb_ctx* ctx, const ruvnb_opts* o, double* l0, double* rowsum);
int ruvnb_estimate_disp(ruvnb_ctx* ctx, const ruvnb_opts* o, double* common_out);
/* Same with the outputs of estimateDisp.par exposed: prior_df_in >= 0 injects prior.df (the reference derives it with
 * limma::squeezeVar; this entry point uses its NON-robust form, fitFDist with the AveLogCPM spline trend); < 0 estimates it.  trended_out: local rows; prior_df_out: the prior d.f. used. */
int ruvnb_estimate_disp_ex(ruvnb_ctx* ctx, const ruvnb_opts* o, double prior_df_in, double* common_out, double* trended_out,
                           double* prior_df_out);
/* estimateDisp.par's `robust` argument (R/estimateDisp.par.R:7,166-168,186-198): robust = 1 is what every call site of the reference
 * passes (R/ruvIIInb.R:210,540; R/fastruvIIInb.R:282,569) and what ruvnb_estimate_disp / ruvnb_fit_nb use -- limma's
 * squeezeVar(robust = TRUE, covariate = AveLogCPM, winsor.tail.p = c(0.05, 0.1)): lowess trend of log s2, Winsorized moments,
 * ONE PRIOR D.F. PER GENE (outlier genes are shrunk less), tagwise values only where prior.n <= 1e6.  robust = 0: limma's fitFDist
 * with the natural-spline trend in AveLogCPM (one prior d.f.).  prior_df_vec_out: local rows (Inf outside the selected genes when
 * robust).  limma's arithmetic is restated in csrc/limma_host.h (third-party, parity unpinned). */
int ruvnb_estimate_disp_r(ruvnb_ctx* ctx, const ruvnb_opts* o, int robust, double prior_df_in, double* common_out, double* trended_out,
                          double* prior_df_vec_out);
/* The north_star variant: per-gene Newton on log(psi) of the NB log-likelihood at the CURRENT fitted means, with the
 * digamma / trigamma differences psi0(y + r) - psi0(r), psi1(y + r) - psi1(r) tabulated per gene and iteration (one
 * sweep per Newton iteration, FP64 pipe).  NOT what the reference computes (it interpolates the 21-point APL grid and
 * shrinks towards a trend); reported separately.  Starts from the current "psi", which it leaves untouched. */
int ruvnb_disp_newton(ruvnb_ctx* ctx, const ruvnb_opts* o, double* psi_out, int maxit);

/* ---- the whole fit on device (R/ruvIIInb.R:266-609) ------------------------------------------ */
/* W0: injected initial W (N x k, column-major) -- the reference draws it from irlba with a random
 * start (:167), so parity runs inject it.  psi_inject: NULL -> ruvnb_estimate_disp at every psi
 * refresh; else `n_psi` consecutive G-vectors used for the initial and subsequent refreshes (the last
 * one repeats).  trace/trace_cap: optional per-iteration records.  logl_outer: outer_maxit doubles. */
int ruvnb_fit_nb(ruvnb_ctx* ctx, const ruvnb_opts* o, const double* W0, const double* psi_inject,
                 int n_psi, ruvnb_trace_rec* trace, int trace_cap, double* logl_outer,
                 ruvnb_fit_stats* stats);

/* ---- fastruvIII.nb streaming passes (R/fastruvIIInb.R) ------------------------------------------------------------ */
/* F3/F4 (:223-623) the subsample fit, on a context that holds the SUBSAMPLED cells (every cell annotated; the caller draws
 * `subsamples`, :118-137, and passes U0 = irlba(RM_Z)$v, :228 -- both are R-RNG dependent and therefore injected).
 *   ruvnb_fast_init         F3: gmean, alpha, Wsub, Mb from Z = log(Y + 1); "psi" is then set by the caller (estimateDisp
 *                           on <= 100 cells, :273-311: a second context over those columns + ruvnb_estimate_disp)
 *   ruvnb_fast_begin_outer  per-cell log-likelihood and wt.ctl of the current parameters, steps <- 1 (:322-352)
 *   ruvnb_fast_iteration    one inner iteration (:357-531): candidate becomes current; *total_new = sum of its per-cell
 *                           log-likelihoods; bad = problematic alpha / W / Mb entries
 *   ruvnb_fast_halve_steps  per-CELL step halving (:538-539);  ruvnb_fast_accept  loglik <- loglik.tmp (:548)
 * ruvnb_snapshot_best / ruvnb_restore_best carry best.wtctl as well.  State names: "cell_loglik", "cell_step" (N), "wt_ctl". */
int ruvnb_fast_init(ruvnb_ctx* ctx, const ruvnb_opts* o, const double* U0);
int ruvnb_fast_begin_outer(ruvnb_ctx* ctx, const ruvnb_opts* o, double* total);
int ruvnb_fast_iteration(ruvnb_ctx* ctx, const ruvnb_opts* o, double* total_new, int bad[3]);
int ruvnb_fast_halve_steps(ruvnb_ctx* ctx, const ruvnb_opts* o);
int ruvnb_fast_accept(ruvnb_ctx* ctx);
/* F5 (:632-716) W for ALL cells from the control genes.  The current state supplies alpha and gmean; wt_ctl: n_ctl
 * weights (best.wtctl); W_init: N x k (zeros with the Wsub rows filled in, :635-636); W_med/W_mad: colMedians/colMads of
 * Wsub (:653-654); max_sweeps = 8, tol = 1e-5 in the reference (:709-712).  W_out (N x k, may be NULL) and "W". */
int ruvnb_fast_W_all(ruvnb_ctx* ctx, int k, double lambda_a, const double* wt_ctl, const double* W_init, const double* W_med,
                     const double* W_mad, int max_sweeps, double tol, double* W_out, int* n_sweeps, double* crit_out);
/* F6 (:719-748) Mb.all, G x N column-major doubles streamed to the host in blocks of block_size cells.  annot_grp[c]:
 * replicate group of cell c (0-based) or -1 for unannotated cells; "Mb" holds the subsample fit's G x m matrix. */
int ruvnb_fast_Mb_all(ruvnb_ctx* ctx, double lambda_b, const int32_t* annot_grp, int64_t block_size, double* Mb_all_out);
/* F7 (:765-770) psi capped at exp(median(log psi) + 3 mad(log psi)), in place (host vector; no device needed). */
int ruvnb_fast_cap_psi(double* psi, int64_t G);

/* ---- cell shards: the full-data passes of fastruvIII.nb and get.res at 10^6 cells ----------------
 * Blocks of `block.size` cells are independent in get.res (R/get.res.R:27-126) and in the W / Mb passes of fastruvIII.nb
 * (R/fastruvIIInb.R:639-689,727-738), so a problem too large for one GPU is cut by CELL: every rank (one context per GPU)
 * holds ALL genes of the cells [c0, c0 + N) of N_total, N being the N of ruvnb_set_design and c0 a multiple of the block size
 * used later.  The three quantities that cross blocks are summed over the communicator of ruvnb_comm_init (call it BEFORE
 * ruvnb_set_design): the per-gene count histograms of the Winsorize quantile (R/fastruvIIInb.R:81-95), colSums(W) for
 * colMeans(W) (R/get.res.R:36) and the convergence criterion of the W sweeps (R/fastruvIIInb.R:709).  A cell-sharded context
 * refuses the gene-sharded operators (ruvnb_irls_iteration, ruvnb_fit_nb, ...). */
int ruvnb_set_cell_shard(ruvnb_ctx* ctx, int64_t N_total, int64_t c0);
/* Rows (genes) that get.res / the fused Mb + PAC pass report: alive[g] = 0 drops local row g from every output (the reference
 * removes the genes the Winsorize cap emptied, R/fastruvIIInb.R:97-101, before anything else sees them; here they stay resident
 * and are skipped).  Outputs then have sum(alive) rows.  NULL restores all rows. */
int ruvnb_set_row_mask(ruvnb_ctx* ctx, const uint8_t* alive);
/* Counts already in device memory (a slab of nrows gene rows x N cells, gene-major, row0 = first local row): the upload
 * permutes them into the resident layout without touching the host.  on_device = 0 takes a host slab the same way (a matrix
 * larger than host memory arrives in pieces).  Call ruvnb_upload_counts_done after the last slab. */
int ruvnb_upload_counts_rows(ruvnb_ctx* ctx, const int32_t* Y_rows, int on_device, int64_t row0, int64_t nrows);
int ruvnb_upload_counts_done(ruvnb_ctx* ctx);
/* Counts as the slots of a dgCMatrix (compressed sparse COLUMN, column = cell): what `Matrix::Matrix(Y, sparse = TRUE)` holds, what
 * ruvIII.nb returns as `counts` (R/ruvIIInb.R:611-612) and what get.res / makeSCE read back (R/get.res.R:14, R/makeSCE.R:22).  A block
 * of columns [c0, c0 + ncols): p = ncols + 1 offsets into i / x (p[0] need not be 0, so a window of the caller's own slots can be
 * passed), i = 0-based gene rows, x = counts as doubles (rounded up like :53).  The transpose into the gene-major resident layout
 * happens on the device; the first block must start at cell 0 (it zero-fills the matrix); finish with ruvnb_upload_counts_done. */
int ruvnb_upload_counts_csc(ruvnb_ctx* ctx, const int32_t* p, const int32_t* i, const double* x, int64_t c0, int64_t ncols);
/* The resident (winsorized) counts of `n` local cells (0-based, any order), all local rows: out is G x n column-major int32 --
 * how the subsample of fastruvIII.nb (R/fastruvIIInb.R:118-137) leaves the device without the whole matrix travelling. */
int ruvnb_get_cells(ruvnb_ctx* ctx, const int64_t* cells, int64_t n, int32_t* out);
/* Page-locked host memory for the large result matrices (a pageable destination runs the device->host stream at a fifth of
 * the PCIe rate). */
void* ruvnb_host_alloc(size_t bytes);
void ruvnb_host_free(void* p);

/* ---- get.res (R/get.res.R:12-129) -------------------------------------------------------------- */
#define RUVNB_RES_PEARSON 1
#define RUVNB_RES_LOGCOUNTS 2
#define RUVNB_RES_QUANTILE 3 /* mid-p percentile-adjusted counts (PAC) */
/* Mb_cells: G x N per-cell Mb (fastruvIII.nb's Mb.all) or NULL to expand the resident G x m Mb by
 * group.  out: G x N column-major doubles (what write_block receives, :125). */
int ruvnb_get_res(ruvnb_ctx* ctx, int type, int64_t block_size, const double* Mb_cells, double* out);
/* get.res of a ZERO-INFLATED fit (R/get.res.R:52-57,69-70,84-85,95-96,104-105; ZIM::dzinb / pzinb / qzinb restated): the state's "pi0"
 * enters as omega_gc = pi0_g where zflag_of_cell[c] != 0, else 0.  The reference reads out$pi0[, curr.batch] -- column batch[c] of
 * the G x N matrix outer(pi0, zeroinf) -- so the flag of cell c is zeroinf[batch[c]] (with one batch: zeroinf[1] for every cell); the
 * host passes exactly that vector.  mean_zeroinf: rowMeans(out$pi0) = pi0_g * mean(zeroinf), the omega of the qzinb call.  NULL
 * switches back to the NB formulas. */
int ruvnb_set_res_zinb(ruvnb_ctx* ctx, const uint8_t* zflag_of_cell, double mean_zeroinf);
/* The same with the sink chosen by the caller.  out_kind: */
#define RUVNB_OUT_F64 0    /* doubles, G x N column-major (what ruvnb_get_res writes)                                     */
#define RUVNB_OUT_I32 1    /* int32 (type RUVNB_RES_QUANTILE only: PAC are integers; half the bytes of the stream)        */
#define RUVNB_OUT_LOG1P 2  /* log(PAC + 1) as doubles: makeSCE's logPAC assay (R/makeSCE.R:36-38)                         */
#define RUVNB_OUT_NONE 3   /* nothing leaves the device: only `checksum` (sum of all values written) -- full-size runs    */
typedef struct ruvnb_res_stats {
  double checksum;          /* sum over all reported elements of the values written (clipped Pearson, PAC, ...)            */
  double device_seconds;    /* first kernel to last kernel of the call, CUDA events on the library's stream                */
  double caps_ms, elem_ms, select_ms, mb_ms;  /* device time in k_res_caps / k_res_elem / the Pearson quantile select +    */
                            /* clip / the Mb.all block (k_fast_mb or transpose), summed over the blocks                    */
  long long elements;       /* reported rows x cells                                                                       */
  long long launches;       /* kernels launched by the call                                                                */
  long long clipped_lo, clipped_hi;  /* reserved                                                                           */
} ruvnb_res_stats;
/* cells [cell_begin, cell_end) only (cell_begin a multiple of block_size, cell_end = -1: all): `out` then starts at cell_begin.
 * stats may be NULL. */
int ruvnb_get_res_ex(ruvnb_ctx* ctx, int type, int64_t block_size, const double* Mb_cells, int out_kind, void* out,
                     int64_t cell_begin, int64_t cell_end, ruvnb_res_stats* stats);
/* The same into the slots of a dgCMatrix (compressed sparse column, column = cell): what makeSCE stores -- `as(get.res(...),
 * "sparseMatrix")`, `as(log(get.res(type = 'quantile') + 1), "sparseMatrix")` (R/makeSCE.R:24-38).  Every block is compacted on the
 * device, so only its non-zero entries cross PCIe.  log1p != 0: values are log(PAC + 1) (type RUVNB_RES_QUANTILE only).  colptr has
 * (cell_end - cell_begin) + 1 slots, rowidx / values `capacity` slots (0-based rows in ascending order within a column); *nnz returns
 * the entries written.  RUVNB_ENOMEM when capacity is too small (*nnz = entries needed so far; nothing else is valid then). */
int ruvnb_get_res_csc(ruvnb_ctx* ctx, int type, int64_t block_size, const double* Mb_cells, int log1p, int64_t cell_begin, int64_t cell_end,
                      int64_t* colptr, int32_t* rowidx, double* values, int64_t capacity, int64_t* nnz, ruvnb_res_stats* stats);
/* fastruvIII.nb's tail as ONE pass per block (R/fastruvIIInb.R:719-748 + makeSCE(assays = 'logPAC') -> get.res(type = 'quantile'),
 * R/fastruvIIInb.R:790, R/makeSCE.R:36-38): the block of Mb.all is computed on the device (as ruvnb_fast_Mb_all does), feeds the
 * caps and the PAC of the same block and is dropped -- the 8 bytes per element of Mb.all never cross PCIe (240 GB at 30k x 1M).
 * Mb_all_out (G x N doubles, column-major) may be NULL; type as ruvnb_get_res (normally RUVNB_RES_QUANTILE). */
int ruvnb_fast_res(ruvnb_ctx* ctx, int type, double lambda_b, const int32_t* annot_grp, int64_t block_size, int out_kind, void* out,
                   double* Mb_all_out, int64_t cell_begin, int64_t cell_end, ruvnb_res_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* RUVNB_H */
