// ctx.cuh -- host-side context of libruvnb_b200: device buffers, error macros, NCCL wiring, and the host functions the
// translation units share (core.cu, irls.cu, sweep_*.cu, disp.cu, res.cu, fast.cu, svd.cu).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/ruvnb.h"
#include "fastmath.cuh"

struct NMState;   // zinb.cuh

// ---- NCCL through dlopen: single-GPU use never needs libnccl, and a host that already loaded
// torch's bundled libnccl.so.2 shares that copy (same soname) --------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string* err) {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { *err = std::string("dlopen libnccl failed: ") + dlerror(); return false; }
    GetUniqueId = (int (*)(ncclUniqueId_t*))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId_t, int))dlsym(h, "ncclCommInitRank");
    CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    Broadcast = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclBroadcast");
    GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather || !Broadcast) { *err = "libnccl lacks a required symbol"; return false; }
    return true;
  }
};
inline NcclApi g_nccl;   // one copy for the whole library (C++17 inline variable)
enum { NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_F64 = 8 };  // ncclDataType_t values (nccl.h)
enum { NCCL_SUM = 0 };

struct StateSet {  // one parameter set on device
  double *gmean = nullptr, *Mb = nullptr, *alpha = nullptr, *adev = nullptr, *W = nullptr, *psi = nullptr, *pi0 = nullptr;
  // Huber scale of THIS parameter set, written by the log-likelihood sweep (certificate) and completed by the
  // exact path: +Inf = every weight provably 1, 0 = not evaluated yet, else mad(signdev)/0.6745
  double* sigma = nullptr;
  double* wtctl = nullptr;  // fastruvIII.nb subsample fit: wt.ctl = rowMeans(sig.inv[ctl,]) of this parameter set (best.wtctl)
  double* psi_w = nullptr;  // ZINB: the dispersion the last E-step (wt.zinb update) used -- R/ruvIIInb.R:262,465,502
  bool sig_valid = false;   // sigma was produced for the parameters currently held
  bool p1_valid = false;    // ctx->A, u, sw, ZS, VS hold the (unit-Huber-weight) Gram / score sums of THIS parameter set with the
                            // current steps: left by the fused log-likelihood sweep that scored it (kernels.cuh k_ll_pass1)
  int nfail = -1;           // rows of `sigma` still marked 0 (exact MAD to be computed), as read back by the host; -1 = not read
  bool maybe_active = false; // some rows went through the exact path: their sigma may be finite (live Huber weights)
};

struct ruvnb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;          // side stream: the few rows with live Huber weights run beside the certified bulk
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  long long launches = 0;
  std::vector<void*> allocs;

  // design -----------------------------------------------------------------------------------
  long long G = 0, N = 0, g0 = 0, g1 = 0;
  int Gl = 0, m = 0, K = 0;
  bool have_design = false, have_counts = false;
  int empty_group = -1;            // first replicate group without a cell (legal on a cell shard only), or -1
  std::vector<int> grp;
  std::vector<uint8_t> ctl, zeroinf;
  bool any_zeroinf = false;
  // cell permutation: groups contiguous, each padded to whole 128-cell chunks
  long long Np = 0;
  int nchunks = 0;
  std::vector<int> pos2cell, cell2pos, chunk_grp, chunk_valid, group_chunk0, group_n;
  int *d_pos2cell = nullptr, *d_cell2pos = nullptr, *d_chunk_grp = nullptr, *d_chunk_valid = nullptr,
      *d_group_chunk0 = nullptr, *d_group_n = nullptr;
  int* d_chunk_meta = nullptr;     // [nchunks] (batch << 24) | (group << 8) | valid cells: one load per chunk in the row walk
  // batch-specific dispersion (R/ruvIIInb.R:278-283,306-315: psi[, batch]): cells carry a batch label, psi / psi_w / the count tables
  // have one column per batch ([nbatch][Gl], R's G x nbatch column-major), and inside a replicate group the cells are ordered by
  // batch so that every 128-cell chunk belongs to ONE batch (ruvnb_set_batches, before ruvnb_set_design)
  ruvnb_psi_fn psi_cb = nullptr; void* psi_cb_user = nullptr;   // ruvnb_set_psi_callback
  int nbatch = 1;
  std::vector<int> batch;          // [N] batch of each cell, 0-based (empty: one batch)
  std::vector<int> chunk_batch, batch_n, group_nch;
  int* d_chunk_batch = nullptr;    // [nchunks]
  int* d_group_nch = nullptr;      // [m] chunks of each group's run (its per-batch sub-runs are padded separately)
  // control genes
  int n_ctl = 0;
  std::vector<int> ctl_gene;       // global gene index of each control
  int* d_ctl_local = nullptr;      // local row of each control gene, or -1
  int* d_ctl_global = nullptr;     // global row of each control gene
  const int32_t** d_ctl_rows = nullptr;
  // counts
  int* d_roworder = nullptr; // [Gl] launch order of the gene rows: heavy rows (many counts >= YT) first
  bool roworder_ready = false;
  int32_t* dY = nullptr;     // [Gl][Np]
  int32_t* dYctl = nullptr;  // [n_ctl][Np] (sharded runs only)
  // state
  StateSet cur, nw, best;
  double *step = nullptr, *loglik = nullptr, *loglik_tmp = nullptr;
  bool have_W = false;
  // scratch
  double *A = nullptr, *u = nullptr, *sw = nullptr, *ZS = nullptr, *VS = nullptr, *cvec = nullptr, *Ab = nullptr,
         *red = nullptr, *amean = nullptr, *S0 = nullptr, *S1 = nullptr, *Wbar = nullptr,
         *RMW = nullptr, *sumsq = nullptr, *med = nullptr, *mad = nullptr, *ctlpack = nullptr, *alpha_full = nullptr,
         *scal = nullptr;
  double* colsum_part = nullptr;   // [COLSUM_B][64] partial column sums (k_colsum_part -> k_colsum_fin)
  double* Wc = nullptr;      // [Gl][Np] weight cache: huber x sig.inv x wt.zinb of every element, written by pass 1 / the fused sweep and
  bool Wc_tried = false;     //   read back by pass 2 (kernels.cuh k_pass2w); taken when HBM has the room, RUVNB_NO_WCACHE=1 turns it off
  double* SWZ = nullptr;     // [nparts][Gl][m] per-group sum w Z of the same sweep
  double* S0c = nullptr;     // [nparts][Gl][m] per-group sum w of the same sweep (own buffer: ctx->S0 is scratch of other operators)
  double* WZc = nullptr;     // [Gl][Np] w Z of every element: the fused sweep's second slab (k_ll_w -> k_gram), allocated with Wc
  double* D = nullptr; long long D_rows = 0;  // exact-MAD scratch [D_rows][Np]
  double* rowmax = nullptr; int* rownan = nullptr;  // per scratch slot: max|d|, NaN flag
  double *dp_ps1 = nullptr, *dp_ps2 = nullptr;   // [DISP_MAXP][Gl][DISP_ND] partial sums of a scoring sweep launched with row parts
  bool opt_no_disp_parts = false;               // RUVNB_NO_DISP_PARTS=1
  bool best_eq_cur = false;      // `best` and `cur` hold the same parameter set (last copy_set between them, nothing changed since)
  bool hints_seeded = false;     // the rows got hints from a sample of their deviances (irls.cu seed_hints); RUVNB_NO_SEED=1 turns it off
  bool opt_no_seed = false;
  double* hint = nullptr;        // [Gl][3] certificate hints lo, hi, tg (k_row_sigma -> k_loglik)
  double* sigma_exact = nullptr; // [Gl] last exact sigma of each row (diagnostics / tests)
  // table-driven math (fastmath.cuh, kernels.cuh k_build_ytabs)
  MathTabs* d_mtabs = nullptr;
  double *tabL = nullptr, *tabR = nullptr, *lgr = nullptr;   // [Gl][YT], [Gl][YT], [Gl]
  double* ctl_tabR = nullptr;                                 // [n_ctl][YT]
  const double* ytab_psi = nullptr; long long ytab_ver = -1, psi_ver = 0;  // which psi the tables were built from
  int *w_fail = nullptr;         // [1 + Np] fast-W failures: count, then positions
  int *cert_diag = nullptr;      // [8] certificate failure reasons (ruvnb_cert_diag)
  // fastruvIII.nb subsample fit (fastfit.cuh), allocated by ruvnb_fast_init
  double *ff_Z = nullptr, *ff_SI = nullptr, *ff_ZS = nullptr, *ff_sirow = nullptr, *ff_wtcell = nullptr, *ff_step = nullptr,
         *ff_cll = nullptr, *ff_cll_tmp = nullptr, *ff_b = nullptr, *ff_Ginv = nullptr, *ff_alpha_c = nullptr, *ff_A = nullptr, *ff_ones = nullptr;
  uint8_t* ff_ctlmask = nullptr;
  int* ff_ctl_rows_idx = nullptr;
  bool ff_ready = false;
  // ZINB (zinb.cuh)
  bool zinb_on = false;          // opts.zinb && some cell is flagged zero-inflated
  uint8_t* d_zinf = nullptr;     // [Np] zero-inflated cells, permuted like the counts
  double *zq = nullptr, *z_spos = nullptr, *z_xeval = nullptr, *z_feval = nullptr;  // q rows [Gl][Np], S_pos, NM points / values
  int *z_npos = nullptr, *z_active = nullptr;
  NMState* z_nm = nullptr;
  int last_nm_rounds = 0;
  // dispersion scratch (disp.cuh), allocated on first use
  double *dp_beta = nullptr, *dp_beta0 = nullptr, *dp_rowsum = nullptr, *dp_emean = nullptr, *dp_l0 = nullptr, *dp_phi_grid = nullptr,
         *dp_tabG = nullptr, *dp_tabRg = nullptr, *dp_lgr_grid = nullptr, *dp_zeros = nullptr, *dp_beta1 = nullptr, *dp_phi_row = nullptr,
         *dp_dev = nullptr, *dp_tabR2 = nullptr;
  double *dp_step = nullptr;                 // [Gl][21] previous NR step per column (stopping rule, disp.cuh DispArgs)
  double *dp_info = nullptr;                 // [Gl][21] Fisher information of the converged grid fits (Cox-Reid term)
  double *dp_xs = nullptr, *dp_m0 = nullptr; // trend smoother scratch: [G] AveLogCPM sorted, [G][21] smoothed l0
  int* dp_perm = nullptr;                    // [G]
  int* dp_flag = nullptr;                    // [2]: some |step| >= tol; rows still iterating
  int* dp_done = nullptr;
  int* dp_live = nullptr;                    // [Gl] list of the rows still iterating
  int* dp_hist = nullptr;                    // [Gl][YT] count histogram of each row (valid while hist_ready)
  bool hist_ready = false;
  double* dp_E = nullptr;                    // [Gl][Np] exp(offset) cache of the dispersion step (disp.cuh DispArgs::Ecache), if it fits
  bool dp_E_tried = false;
  double* dp_syo = nullptr;                  // [Gl] sum y offset per row
  bool dp_beta_warm = false;                 // dp_beta holds the converged grid intercepts of an earlier call on these counts
  bool dp_ready = false;
  int disp_nr_sweeps = 0;        // Newton-Raphson sweeps of the last dispersion call (diagnostics)
  int *flags = nullptr;      // [8] device ints: 0 bad_alpha 1 bad_W 2 bad_Mb 3 nan_Mb 4 n_fail 5..7 spare
  int *faillist = nullptr;
  const double* faillist_for = nullptr;  // the sigma buffer whose unevaluated rows faillist currently lists
  int nparts = 1;            // row parts of the IRLS sweeps (kernels.cuh RowGeom::nparts), chosen from the shard size
  int launch_parts = 1;      // parts of the launch being issued (geom() and LAUNCH_ROWS read it)
  double* llparts = nullptr; // [nparts][Gl] partial log-likelihoods
  int* certc = nullptr;      // [nparts][Gl][4] certificate counters
  int *list_inf = nullptr, *list_fin = nullptr, *part_counts = nullptr;  // rows by Huber status (k_partition_sigma)
  int *h_flags = nullptr;    // pinned mirror
  double* h_scal = nullptr;  // pinned [16]
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // cell shard (fastruvIII.nb full-data passes, get.res at 10^6 cells): this context holds cells [cell0, cell0 + N) of
  // N_total and ALL genes; what crosses ranks is summed over the communicator (ruvnb_set_cell_shard)
  bool cell_shard = false;
  long long N_total = 0, cell0 = 0;
  uint8_t* res_zflag = nullptr; double res_zavg = 0.0;   // get.res of a ZINB fit (ruvnb_set_res_zinb)
  uint8_t* d_row_alive = nullptr; int* d_orow = nullptr; int Gout = 0;   // rows reported by get.res / F6 (ruvnb_set_row_mask)
  bool even_shards = false;  // rank r holds rows [r per, min(G, (r+1) per)), per = ceil(G / nranks): slices can be all-gathered
  double* agather = nullptr; // [nranks][per K + 1] alpha slices + the rank's non-finite count (all-gather buffer)
  cudaEvent_t ev_ctl = nullptr;  // the control-gene parameter exchange on the side stream has finished
  // switches read once from the environment (tuning / A-B measurements)
  bool opt_no_fuse = false;  // RUVNB_NO_FUSE=1: separate log-likelihood and Gram sweeps (three Y sweeps per iteration)
  double opt_hint_half = 0.7;   // RUVNB_HINT_HALF: half-width of a fresh certificate hint window in units of (MAD - t), < 0.85.  Measured on the
                                // whole fit at cfg 2 (exact-MAD rows / ms): 0.4 67 111 / 127, 0.55 60 472 / 114, 0.7 54 818 / 104, 0.8 62 192 / 118
  int opt_slab_cfg = 0;      // RUVNB_SLAB_CFG bit 0: k_gram stage shape B, bit 1: k_pass2w stage shape B (A/B runs)
  // stats of the last iteration
  int last_fail_rows = 0, last_active_rows = 0, last_fail_cells = 0;
  // per-kernel device timing of the sweep kernels (CUDA events on ctx->stream, read at the
  // end-of-iteration sync): 0 devstats 1 pass1 2 w_update 3 pass2 4 loglik 5 exact-MAD
  cudaEvent_t kev[RUVNB_NKERN][2] = {};
  bool kev_armed[RUVNB_NKERN] = {};
  double kms[RUVNB_NKERN] = {};
  long long kcnt[RUVNB_NKERN] = {};

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    err = buf;
    return code;
  }
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return ctx->fail(RUVNB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define CKL()                                                                                      \
  do {                                                                                             \
    ctx->launches++;                                                                               \
    cudaError_t e_ = cudaGetLastError();                                                           \
    if (e_ != cudaSuccess)                                                                         \
      return ctx->fail(RUVNB_ECUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)
#define RET(x) do { int r_ = (x); if (r_ != RUVNB_OK) return r_; } while (0)

// call-scoped device buffers: released on every way out of the call (error returns included)
struct ScopedDev {
  std::vector<void*> ptrs;
  ~ScopedDev() { for (void* q : ptrs) cudaFree(q); }
  template <class T>
  cudaError_t alloc(T** out, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes ? bytes : 1);
    if (e == cudaSuccess) { ptrs.push_back(q); *out = (T*)q; }
    return e;
  }
};

template <class T>
static int dev_alloc(ruvnb_ctx* ctx, T** p, long long n) {
  void* q = nullptr;
  if (n <= 0) n = 1;
  cudaError_t e = cudaMalloc(&q, (size_t)n * sizeof(T));
  if (e != cudaSuccess) return ctx->fail(RUVNB_ENOMEM, "cudaMalloc of %lld bytes failed: %s", n * (long long)sizeof(T), cudaGetErrorString(e));
  ctx->allocs.push_back(q);
  *p = (T*)q;
  return RUVNB_OK;
}
template <class T>
static int dev_upload(ruvnb_ctx* ctx, T** p, const std::vector<T>& v) {
  RET(dev_alloc(ctx, p, (long long)v.size()));
  if (!v.empty()) CK(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return RUVNB_OK;
}
static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// all-reduce (sum) in place over the ranks; no-op on a single GPU
static int allreduce(ruvnb_ctx* ctx, void* buf, size_t count, int dtype) {
  if (ctx->nranks <= 1) return RUVNB_OK;
  int r = g_nccl.AllReduce(buf, buf, count, dtype, NCCL_SUM, ctx->comm, ctx->stream);
  if (r != 0) return ctx->fail(RUVNB_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return RUVNB_OK;
}

// bracket a sweep kernel with events; collected by kev_collect() after the next stream sync
#define KEV_BEGIN(id) do { if (ctx->kev[id][0]) { CK(cudaEventRecord(ctx->kev[id][0], ctx->stream)); } } while (0)
#define KEV_END(id) do { if (ctx->kev[id][0]) { CK(cudaEventRecord(ctx->kev[id][1], ctx->stream)); ctx->kev_armed[id] = true; } } while (0)
static void kev_collect(ruvnb_ctx* ctx) {
  for (int i = 0; i < RUVNB_NKERN; ++i) {
    if (!ctx->kev_armed[i]) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->kev[i][0], ctx->kev[i][1]) == cudaSuccess) { ctx->kms[i] += ms; ctx->kcnt[i]++; }
    ctx->kev_armed[i] = false;
  }
}

#define DISPATCH_K(KVAL, ...)                                        \
  switch (KVAL) {                                                    \
    case 1: { constexpr int KT = 1; __VA_ARGS__; } break;            \
    case 2: { constexpr int KT = 2; __VA_ARGS__; } break;            \
    case 3: { constexpr int KT = 3; __VA_ARGS__; } break;            \
    case 4: { constexpr int KT = 4; __VA_ARGS__; } break;            \
    case 5: { constexpr int KT = 5; __VA_ARGS__; } break;            \
    case 6: { constexpr int KT = 6; __VA_ARGS__; } break;            \
    case 7: { constexpr int KT = 7; __VA_ARGS__; } break;            \
    case 8: { constexpr int KT = 8; __VA_ARGS__; } break;            \
    default: return ctx->fail(RUVNB_EARG, "k = %d outside 1..%d", KVAL, RUVNB_MAX_K); \
  }

// ---- host functions shared between the translation units ------------------------------------------------------------
int ensure_state(ruvnb_ctx* ctx, int K);                              // core.cu
int check_ready(ruvnb_ctx* ctx, const ruvnb_opts* o);                 // core.cu
int ensure_roworder(ruvnb_ctx* ctx);                                  // core.cu
int ensure_ytabs(ruvnb_ctx* ctx, const double* psi);                  // core.cu
int make_rmw(ruvnb_ctx* ctx, const double* W);                        // irls.cu
struct FastMbPrep { double *d_rmed = nullptr, *d_rmad = nullptr; int* d_annot = nullptr; uint8_t* d_ctl = nullptr; };
int fast_mb_prepare(ruvnb_ctx* ctx, const int32_t* annot_grp, ScopedDev* own, FastMbPrep* p);   // fast.cu
