// core.cu -- C ABI of libruvnb_b200 (include/ruvnb.h): context life cycle, NCCL wiring, design, count upload, Winsorize, state access.
// sm_100a only; no CPU fallback anywhere in this library.
#include <stdlib.h>
#include "ctx.cuh"
#include "launch.cuh"
#include "winsor.cuh"

extern "C" {

int ruvnb_abi_version(void) { return RUVNB_ABI_VERSION; }

void ruvnb_opts_default(ruvnb_opts* o) {
  // defaults of ruvIII.nb (R/ruvIIInb.R:27-28) and its hard-coded constants (:524,533,597)
  o->k = 2; o->robust = 0; o->lambda_a = 0.01; o->lambda_b = 16.0; o->step_fac = 0.5;
  o->inner_maxit = 50; o->outer_maxit = 25; o->inner_tol = 1e-4; o->outer_tol = 1e-4; o->psi_tol = 1e-8;
  o->zinb = 0; o->huber_fastpath = 1; o->verbose = 0;
}

int ruvnb_create(ruvnb_ctx** out, int device) {
  if (!out) return RUVNB_EARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return RUVNB_ECUDA;  // no CPU fallback
  ruvnb_ctx* ctx = new ruvnb_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&ctx->stream) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_flags, 8 * sizeof(int)) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_scal, 16 * sizeof(double)) != cudaSuccess) {
    delete ctx;
    return RUVNB_ECUDA;
  }
  for (int i = 0; i < RUVNB_NKERN; ++i) { cudaEventCreate(&ctx->kev[i][0]); cudaEventCreate(&ctx->kev[i][1]); }
  cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->ev_ctl, cudaEventDisableTiming);
  { const char* e = getenv("RUVNB_NO_FUSE"); ctx->opt_no_fuse = e && *e && *e != '0'; }
  { const char* e = getenv("RUVNB_SLAB_CFG"); ctx->opt_slab_cfg = e ? atoi(e) : 0; }
  { const char* e = getenv("RUVNB_NO_SEED"); ctx->opt_no_seed = e && *e && *e != '0'; }
  { const char* e = getenv("RUVNB_NO_DISP_PARTS"); ctx->opt_no_disp_parts = e && *e && *e != '0'; }
  { const char* e = getenv("RUVNB_HINT_HALF"); double v = e ? atof(e) : 0.0; if (v > 0.05 && v < 0.85) ctx->opt_hint_half = v; }
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  *out = ctx;
  return RUVNB_OK;
}

void ruvnb_destroy(ruvnb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) g_nccl.CommDestroy(ctx->comm);
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
  for (int i = 0; i < RUVNB_NKERN; ++i) { if (ctx->kev[i][0]) cudaEventDestroy(ctx->kev[i][0]); if (ctx->kev[i][1]) cudaEventDestroy(ctx->kev[i][1]); }
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_ctl) cudaEventDestroy(ctx->ev_ctl);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* ruvnb_last_error(const ruvnb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int ruvnb_kernel_time(const ruvnb_ctx* ctx, int which, double* ms_total, long long* launches) {
  if (!ctx || which < 0 || which >= RUVNB_NKERN) return RUVNB_EARG;
  if (ms_total) *ms_total = ctx->kms[which];
  if (launches) *launches = ctx->kcnt[which];
  return RUVNB_OK;
}
int ruvnb_last_exact_rows(const ruvnb_ctx* ctx) { return ctx ? ctx->last_fail_rows : -1; }
// measured FP64 FMA throughput of the device (TFLOP/s, 2 flops per DFMA), best of 5
int ruvnb_measure_fp64(ruvnb_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return RUVNB_EARG;
  CK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device));
  double* d = nullptr;
  CK(cudaMalloc((void**)&d, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 2048, grid = prop.multiProcessorCount * 8;
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0, ctx->stream));
    k_fp64_peak<<<grid, 256, 0, ctx->stream>>>(d, 1e-9, 0.999999, iters);
    CKL();
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1)); CK(cudaFree(d));
  *tflops = best;
  return RUVNB_OK;
}
// why rows failed the Huber certificate since the last call (tuning aid): [0] no usable hint, [1] NaN deviance,
// [2] max|d| = 0, [3] max|d| grew past the hint's margin, [4]/[5] the median left [lo, hi] below/above,
// [6] more than half of the row inside the window, [7] fast path off.  Resets the counters.
int ruvnb_cert_diag(ruvnb_ctx* ctx, int out[8]) {
  if (!ctx || !out) return RUVNB_EARG;
  for (int i = 0; i < 8; ++i) out[i] = 0;
  if (!ctx->cert_diag) return RUVNB_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->cert_diag, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemsetAsync(ctx->cert_diag, 0, 8 * sizeof(int), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RUVNB_OK;
}
int ruvnb_last_exact_cells(const ruvnb_ctx* ctx) { return ctx ? ctx->last_fail_cells : -1; }
int ruvnb_selftest_math(int which, const double* x, double* out, int64_t n) {
  if (!x || !out || n < 0 || which < 0 || which > 2) return RUVNB_EARG;
  static MathTabs mt;
  static FmRL rl[FM_NT];
  static bool built = false;
  if (!built) { build_math_tabs(&mt); for (int j = 0; j < FM_NT; ++j) { rl[j].rc = mt.rc[j]; rl[j].lc = mt.lc[j]; } built = true; }
  // 2: the unchecked logarithm on the interleaved table (disp.cuh, cached APL sweep); normal positive arguments only
  for (int64_t i = 0; i < n; ++i) out[i] = which == 0 ? fexp(x[i], mt.ex) : which == 1 ? flog(x[i], mt.rc, mt.lc) : flog_nc2(x[i], rl);
  return RUVNB_OK;
}
void* ruvnb_stream(const ruvnb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
void ruvnb_kernel_time_reset(ruvnb_ctx* ctx) {
  if (!ctx) return;
  for (int i = 0; i < RUVNB_NKERN; ++i) { ctx->kms[i] = 0.0; ctx->kcnt[i] = 0; }
}
long long ruvnb_kernel_launches(const ruvnb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- multi-GPU ------------------------------------------------------------------------------------
int ruvnb_comm_unique_id(unsigned char id[128]) {
  std::string err;
  if (!g_nccl.load(&err)) return RUVNB_ENCCL;
  ncclUniqueId_t u;
  if (g_nccl.GetUniqueId(&u) != 0) return RUVNB_ENCCL;
  memcpy(id, u.internal, 128);
  return RUVNB_OK;
}
int ruvnb_comm_init(ruvnb_ctx* ctx, const unsigned char id[128], int rank, int nranks) {
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return RUVNB_EARG;
  if (nranks == 1) { ctx->rank = 0; ctx->nranks = 1; return RUVNB_OK; }
  if (!g_nccl.load(&ctx->err)) return RUVNB_ENCCL;
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId_t u;
  memcpy(u.internal, id, 128);
  int r = g_nccl.CommInitRank(&ctx->comm, nranks, u, rank);
  if (r != 0) return ctx->fail(RUVNB_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  ctx->rank = rank; ctx->nranks = nranks;
  // NCCL connects its rings / trees lazily at the first collective (hundreds of ms at 8 ranks): do that here, as part
  // of the one-time set-up, not inside the first IRLS iteration
  int* warm = nullptr;
  CK(cudaMalloc((void**)&warm, 4));
  CK(cudaMemsetAsync(warm, 0, 4, ctx->stream));
  RET(allreduce(ctx, warm, 1, NCCL_INT32));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(warm));
  return RUVNB_OK;
}

// ---- problem set-up -------------------------------------------------------------------------------
// batch of every cell (0-based), for a batch-specific dispersion psi[g, batch(c)] (R/ruvIIInb.R:213-220,278-283,306-315,542-548).
// Must precede ruvnb_set_design: the cell permutation keeps the batches of a replicate group apart.
int ruvnb_set_batches(ruvnb_ctx* ctx, const int32_t* batch_of_cell, int64_t N, int nbatch) {
  if (!ctx) return RUVNB_EARG;
  if (ctx->have_design) return ctx->fail(RUVNB_ESTATE, "set_batches: call it before ruvnb_set_design (the cell order depends on it)");
  if (nbatch < 1 || nbatch > 127 || (nbatch > 1 && (!batch_of_cell || N <= 0))) return ctx->fail(RUVNB_EARG, "set_batches: nbatch = %d outside 1..127, or no labels", nbatch);
  ctx->nbatch = nbatch; ctx->batch.clear();
  if (nbatch == 1) return RUVNB_OK;
  ctx->batch.assign(batch_of_cell, batch_of_cell + N);
  // (a batch without a cell is legal: a cell shard of get.res may not meet every batch)
  for (int64_t c = 0; c < N; ++c) {
    const int b = ctx->batch[c];
    if (b < 0 || b >= nbatch) { ctx->nbatch = 1; ctx->batch.clear(); return ctx->fail(RUVNB_EARG, "set_batches: batch label %d of cell %lld outside 0..%d", b, (long long)c, nbatch - 1); }
  }
  return RUVNB_OK;
}
int ruvnb_nbatch(const ruvnb_ctx* ctx) { return ctx ? ctx->nbatch : 0; }

int ruvnb_set_design(ruvnb_ctx* ctx, int64_t G, int64_t N, int m, const int32_t* grp_of_cell, const uint8_t* ctl,
                     const uint8_t* zeroinf, int64_t g0, int64_t g1) {
  if (!ctx) return RUVNB_EARG;
  if (ctx->have_design) return ctx->fail(RUVNB_ESTATE, "design already set: one context holds one problem");
  if (G <= 0 || N <= 0 || m <= 0 || !grp_of_cell || !ctl || g0 < 0 || g1 > G || g0 >= g1)
    return ctx->fail(RUVNB_EARG, "set_design: bad sizes G=%lld N=%lld m=%d rows [%lld,%lld)", (long long)G, (long long)N, m, (long long)g0, (long long)g1);
  if (N > (1ll << 30) || (g1 - g0) > (1ll << 30)) return ctx->fail(RUVNB_EARG, "set_design: N or local G above 2^30");
  CK(cudaSetDevice(ctx->device));
  ctx->G = G; ctx->N = N; ctx->m = m; ctx->g0 = g0; ctx->g1 = g1; ctx->Gl = (int)(g1 - g0);
  ctx->grp.assign(grp_of_cell, grp_of_cell + N);
  ctx->ctl.assign(ctl, ctl + G);
  ctx->zeroinf.assign(N, 0);
  if (zeroinf) { ctx->zeroinf.assign(zeroinf, zeroinf + N); for (auto z : ctx->zeroinf) ctx->any_zeroinf |= (z != 0); }
  ctx->group_n.assign(m, 0);
  for (long long c = 0; c < N; ++c) {
    int j = ctx->grp[c];
    if (j < 0 || j >= m) return ctx->fail(RUVNB_EARG, "set_design: group label %d of cell %lld outside 0..%d", j, c, m - 1);
    ctx->group_n[j]++;
  }
  // every column of M must be non-empty (R/fastruvIIInb.R:60 stop(); ruvIII.nb's projection would
  // divide by zero, R/ruvIIInb.R:657-665)
  // (a cell shard may hold no cell of some group: only the IRLS operators need every group, and check_ready refuses them)
  ctx->empty_group = -1;
  for (int j = 0; j < m; ++j) if (ctx->group_n[j] == 0 && ctx->empty_group < 0) ctx->empty_group = j;
  // runs of cells: one per (group, batch), group-major, each padded to whole chunks -- a group stays one contiguous run of chunks
  // (its per-batch sub-runs follow each other), and a chunk never mixes batches
  const int nb = ctx->nbatch;
  if (nb > 1 && (long long)ctx->batch.size() != N) return ctx->fail(RUVNB_EARG, "set_design: ruvnb_set_batches was given %lld cells, the design has %lld", (long long)ctx->batch.size(), (long long)N);
  std::vector<int> run_n((size_t)m * nb, 0);
  ctx->batch_n.assign(nb, 0);
  for (long long c = 0; c < N; ++c) { const int b = nb > 1 ? ctx->batch[c] : 0; run_n[(size_t)ctx->grp[c] * nb + b]++; ctx->batch_n[b]++; }
  ctx->group_chunk0.assign(m, 0);
  std::vector<int> run_chunk0((size_t)m * nb, 0);
  int nch = 0;
  for (int j = 0; j < m; ++j) {
    ctx->group_chunk0[j] = nch;
    for (int b = 0; b < nb; ++b) { run_chunk0[(size_t)j * nb + b] = nch; nch += (run_n[(size_t)j * nb + b] + CH - 1) / CH; }
  }
  ctx->group_nch.assign(m, 0);
  for (int j = 0; j < m; ++j) ctx->group_nch[j] = (j + 1 < m ? ctx->group_chunk0[j + 1] : nch) - ctx->group_chunk0[j];
  nch = (nch + CHPAD - 1) / CHPAD * CHPAD;  // whole tiles: trailing chunks are empty (valid = 0)
  ctx->nchunks = nch; ctx->Np = (long long)nch * CH;
  ctx->chunk_grp.assign(nch, ctx->m - 1); ctx->chunk_valid.assign(nch, 0); ctx->chunk_batch.assign(nch, 0);
  for (int j = 0; j < m; ++j)
    for (int b = 0; b < nb; ++b) {
      const int n = run_n[(size_t)j * nb + b], nc = (n + CH - 1) / CH, q0 = run_chunk0[(size_t)j * nb + b];
      for (int q = 0; q < nc; ++q) {
        ctx->chunk_grp[q0 + q] = j; ctx->chunk_batch[q0 + q] = b;
        ctx->chunk_valid[q0 + q] = std::min(CH, n - q * CH);
      }
    }
  ctx->pos2cell.assign(ctx->Np, -1); ctx->cell2pos.assign(N, 0);
  std::vector<long long> fill((size_t)m * nb);
  for (size_t r = 0; r < fill.size(); ++r) fill[r] = (long long)run_chunk0[r] * CH;
  for (long long c = 0; c < N; ++c) {  // stable: cells keep their order inside a run
    long long p = fill[(size_t)ctx->grp[c] * nb + (nb > 1 ? ctx->batch[c] : 0)]++;
    ctx->pos2cell[p] = (int)c; ctx->cell2pos[c] = (int)p;
  }
  ctx->ctl_gene.clear();
  for (long long g = 0; g < G; ++g) if (ctx->ctl[g]) ctx->ctl_gene.push_back((int)g);
  ctx->n_ctl = (int)ctx->ctl_gene.size();
  if (ctx->n_ctl == 0) return ctx->fail(RUVNB_EARG, "set_design: no control gene");
  std::vector<int> ctl_local(ctx->n_ctl);
  for (int i = 0; i < ctx->n_ctl; ++i) {
    long long g = ctx->ctl_gene[i];
    ctl_local[i] = (g >= g0 && g < g1) ? (int)(g - g0) : -1;
  }
  RET(dev_upload(ctx, &ctx->d_pos2cell, ctx->pos2cell));
  RET(dev_upload(ctx, &ctx->d_cell2pos, ctx->cell2pos));
  RET(dev_upload(ctx, &ctx->d_chunk_grp, ctx->chunk_grp));
  RET(dev_upload(ctx, &ctx->d_chunk_valid, ctx->chunk_valid));
  {
    if (m > 65535) return ctx->fail(RUVNB_EARG, "set_design: more than 65535 replicate groups");
    std::vector<int> meta(nch);
    for (int q = 0; q < nch; ++q) meta[q] = (ctx->chunk_batch[q] << 24) | (ctx->chunk_grp[q] << 8) | ctx->chunk_valid[q];
    RET(dev_upload(ctx, &ctx->d_chunk_meta, meta));
    RET(dev_upload(ctx, &ctx->d_chunk_batch, ctx->chunk_batch));
    RET(dev_upload(ctx, &ctx->d_group_nch, ctx->group_nch));
  }
  RET(dev_upload(ctx, &ctx->d_group_chunk0, ctx->group_chunk0));
  RET(dev_upload(ctx, &ctx->d_group_n, ctx->group_n));
  RET(dev_upload(ctx, &ctx->d_ctl_local, ctl_local));
  RET(dev_upload(ctx, &ctx->d_ctl_global, ctx->ctl_gene));
  if (ctx->any_zeroinf) {
    std::vector<uint8_t> zp(ctx->Np, 0);
    for (long long c = 0; c < N; ++c) zp[ctx->cell2pos[c]] = ctx->zeroinf[c] ? 1 : 0;
    RET(dev_upload(ctx, &ctx->d_zinf, zp));
  }
  RET(dev_alloc(ctx, &ctx->flags, 8));
  RET(dev_alloc(ctx, &ctx->faillist, ctx->Gl));
  RET(dev_alloc(ctx, &ctx->list_inf, ctx->Gl)); RET(dev_alloc(ctx, &ctx->list_fin, ctx->Gl)); RET(dev_alloc(ctx, &ctx->part_counts, 2));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_design = true;
  return RUVNB_OK;
}

// upload `rows` gene rows (host index list or contiguous range) into dst[rows][Np], permuted + padded
static int upload_rows(ruvnb_ctx* ctx, const int32_t* Y, int layout, long long nrows_total_host, const int* rowidx,
                       long long row0, long long rows, int32_t* dst) {
  const long long N = ctx->N, Np = ctx->Np;
  if (layout == RUVNB_LAYOUT_GENE_MAJOR) {
    // slabs of whole rows through a staging buffer (<= 256 MB)
    long long slab = std::max(1ll, (256ll << 20) / (N * 4));
    slab = std::min(slab, rows);
    int32_t* stage = nullptr;
    CK(cudaMalloc((void**)&stage, (size_t)slab * N * 4));
    for (long long r0 = 0; r0 < rows; r0 += slab) {
      long long nr = std::min(slab, rows - r0);
      if (rowidx) {
        for (long long r = 0; r < nr; ++r)
          CK(cudaMemcpyAsync(stage + r * N, Y + (long long)rowidx[r0 + r] * N, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
      } else {
        CK(cudaMemcpyAsync(stage, Y + (row0 + r0) * N, (size_t)nr * N * 4, cudaMemcpyHostToDevice, ctx->stream));
      }
      k_permute_rows<<<nblk(nr * Np, 256), 256, 0, ctx->stream>>>(stage, nr, N, N, 1, ctx->d_pos2cell, Np, dst + r0 * Np);
      CKL();
      CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaFree(stage));
  } else {
    // R layout: Y[g + Gh*c], Gh = rows of the host matrix.  Slabs of whole cells.
    const long long Gh = nrows_total_host;
    long long slab = std::max(1ll, (256ll << 20) / (Gh * 4));
    slab = std::min(slab, N);
    int32_t* stage = nullptr;
    int* d_rowidx = nullptr;
    CK(cudaMalloc((void**)&stage, (size_t)slab * Gh * 4));
    if (rowidx) {
      CK(cudaMalloc((void**)&d_rowidx, (size_t)rows * 4));
      CK(cudaMemcpyAsync(d_rowidx, rowidx, (size_t)rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    // pads first
    k_fill_i32<<<nblk(rows * Np, 256), 256, 0, ctx->stream>>>(dst, rows * Np, -1);
    CKL();
    for (long long c0 = 0; c0 < N; c0 += slab) {
      long long nc = std::min(slab, N - c0);
      CK(cudaMemcpyAsync(stage, Y + c0 * Gh, (size_t)nc * Gh * 4, cudaMemcpyHostToDevice, ctx->stream));
      dim3 grid(nblk(rows, 32), nblk(nc, 32));
      k_scatter_cellmajor<<<grid, dim3(32, 8), 0, ctx->stream>>>(stage, Gh, d_rowidx, row0, rows, c0, nc, ctx->d_cell2pos, Np, dst);
      CKL();
      CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaFree(stage));
    if (d_rowidx) CK(cudaFree(d_rowidx));
  }
  return RUVNB_OK;
}

// The control-gene rows of a gene-sharded fit without a host copy: every control gene is resident on the rank that owns its row, so
// each rank copies its own control rows into the replicated slab and broadcasts them over NVLink (one ncclBroadcast per owning
// rank) -- instead of every rank uploading the whole n_ctl x N slab over PCIe (0.4 GB per rank at cfg 2).
static int exchange_ctl_rows(ruvnb_ctx* ctx) {
  const int nr = ctx->nranks;
  long long* d_rg = nullptr;
  ScopedDev own;
  CK(own.alloc(&d_rg, (size_t)2 * nr * 8));
  const long long mine[2] = {ctx->g0, ctx->g1};
  CK(cudaMemcpyAsync(d_rg + 2 * ctx->rank, mine, 16, cudaMemcpyHostToDevice, ctx->stream));
  int r = g_nccl.AllGather(d_rg + 2 * ctx->rank, d_rg, 2, NCCL_INT64, ctx->comm, ctx->stream);
  if (r != 0) return ctx->fail(RUVNB_ENCCL, "ncclAllGather (shard ranges): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  std::vector<long long> rg((size_t)2 * nr);
  CK(cudaMemcpyAsync(rg.data(), d_rg, rg.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  long long covered = 0;
  for (int q = 0; q < nr; ++q) {
    // ctl_gene is ascending: the controls of rank q are one run [i0, i1) of the slab
    const int i0 = (int)(std::lower_bound(ctx->ctl_gene.begin(), ctx->ctl_gene.end(), (int)rg[2 * q]) - ctx->ctl_gene.begin());
    const int i1 = (int)(std::lower_bound(ctx->ctl_gene.begin(), ctx->ctl_gene.end(), (int)rg[2 * q + 1]) - ctx->ctl_gene.begin());
    if (i1 <= i0) continue;
    covered += i1 - i0;
    if (q == ctx->rank)
      for (int i = i0; i < i1; ++i)
        CK(cudaMemcpyAsync(ctx->dYctl + (long long)i * ctx->Np, ctx->dY + (long long)(ctx->ctl_gene[i] - ctx->g0) * ctx->Np, (size_t)ctx->Np * 4,
                           cudaMemcpyDeviceToDevice, ctx->stream));
    int32_t* seg = ctx->dYctl + (long long)i0 * ctx->Np;
    r = g_nccl.Broadcast(seg, seg, (size_t)(i1 - i0) * ctx->Np, NCCL_INT32, q, ctx->comm, ctx->stream);
    if (r != 0) return ctx->fail(RUVNB_ENCCL, "ncclBroadcast (control rows): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if (covered != ctx->n_ctl) return ctx->fail(RUVNB_EARG, "upload_counts: the ranks' row ranges own %lld of the %d control genes (overlapping or incomplete shards)", covered, ctx->n_ctl);
  return RUVNB_OK;
}

int ruvnb_upload_counts_dense(ruvnb_ctx* ctx, const int32_t* Y, int layout, const int32_t* Yctl) {
  if (!ctx || !Y) return RUVNB_EARG;
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "upload_counts: call ruvnb_set_design first");
  if (layout != RUVNB_LAYOUT_GENE_MAJOR && layout != RUVNB_LAYOUT_CELL_MAJOR) return ctx->fail(RUVNB_EARG, "upload_counts: bad layout %d", layout);
  CK(cudaSetDevice(ctx->device));
  const bool sharded = (ctx->Gl != ctx->G);
  if (sharded && !Yctl && ctx->nranks < 2) return ctx->fail(RUVNB_EARG, "upload_counts: a gene shard needs the control-gene rows (Yctl), or a communicator to fetch them from the owning ranks");
  if (!ctx->dY) RET(dev_alloc(ctx, &ctx->dY, (long long)ctx->Gl * ctx->Np));
  // Y holds the LOCAL rows (Gl x N)
  RET(upload_rows(ctx, Y, layout, ctx->Gl, nullptr, 0, ctx->Gl, ctx->dY));
  std::vector<const int32_t*> rows(ctx->n_ctl);
  if (sharded) {
    if (!ctx->dYctl) RET(dev_alloc(ctx, &ctx->dYctl, (long long)ctx->n_ctl * ctx->Np));
    if (Yctl) RET(upload_rows(ctx, Yctl, layout, ctx->n_ctl, nullptr, 0, ctx->n_ctl, ctx->dYctl));
    else RET(exchange_ctl_rows(ctx));
    for (int i = 0; i < ctx->n_ctl; ++i) rows[i] = ctx->dYctl + (long long)i * ctx->Np;
  } else {
    for (int i = 0; i < ctx->n_ctl; ++i) rows[i] = ctx->dY + (long long)ctx->ctl_gene[i] * ctx->Np;
  }
  if (!ctx->d_ctl_rows) RET(dev_alloc(ctx, &ctx->d_ctl_rows, ctx->n_ctl));
  CK(cudaMemcpyAsync((void*)ctx->d_ctl_rows, rows.data(), rows.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_counts = true; ctx->hints_seeded = false; ctx->roworder_ready = false; ctx->hist_ready = false; ctx->dp_beta_warm = false;
  return RUVNB_OK;
}

int ruvnb_upload_counts_csr(ruvnb_ctx* ctx, const int64_t* indptr, const int32_t* indices, const int32_t* values,
                            const int64_t* ctl_indptr, const int32_t* ctl_indices, const int32_t* ctl_values) {
  // CSR rows are expanded on the device into the same padded gene-major int32 tiles the dense path
  // uses: the IRLS sweeps touch every (gene, cell) pair whatever the sparsity (mu > 0 everywhere), so
  // the streaming layout stays dense; CSR only saves the host->device transfer.
  if (!ctx || !indptr || !indices || !values) return RUVNB_EARG;
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "upload_counts: call ruvnb_set_design first");
  CK(cudaSetDevice(ctx->device));
  const bool sharded = (ctx->Gl != ctx->G);
  if (sharded && (!ctl_indptr || !ctl_indices || !ctl_values)) return ctx->fail(RUVNB_EARG, "upload_counts_csr: a gene shard needs the control-gene rows");
  auto expand = [&](const int64_t* ip, const int32_t* ix, const int32_t* vl, long long rows, int32_t* dst) -> int {
    long long nnz = ip[rows] - ip[0];
    long long* d_ip = nullptr; int32_t *d_ix = nullptr, *d_vl = nullptr;
    CK(cudaMalloc((void**)&d_ip, (size_t)(rows + 1) * 8));
    CK(cudaMalloc((void**)&d_ix, (size_t)std::max(1ll, nnz) * 4));
    CK(cudaMalloc((void**)&d_vl, (size_t)std::max(1ll, nnz) * 4));
    CK(cudaMemcpyAsync(d_ip, ip, (size_t)(rows + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
      CK(cudaMemcpyAsync(d_ix, ix + ip[0], (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(d_vl, vl + ip[0], (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    k_fill_pad<<<nblk(rows * ctx->Np, 256), 256, 0, ctx->stream>>>(dst, rows, ctx->Np, ctx->d_pos2cell);
    CKL();
    k_csr_scatter<<<nblk(rows * 32, 256), 256, 0, ctx->stream>>>(d_ip, d_ix, d_vl, rows, ip[0], ctx->d_cell2pos, ctx->Np, dst);
    CKL();
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(d_ip)); CK(cudaFree(d_ix)); CK(cudaFree(d_vl));
    return RUVNB_OK;
  };
  if (!ctx->dY) RET(dev_alloc(ctx, &ctx->dY, (long long)ctx->Gl * ctx->Np));
  RET(expand(indptr, indices, values, ctx->Gl, ctx->dY));
  std::vector<const int32_t*> rows(ctx->n_ctl);
  if (sharded) {
    if (!ctx->dYctl) RET(dev_alloc(ctx, &ctx->dYctl, (long long)ctx->n_ctl * ctx->Np));
    RET(expand(ctl_indptr, ctl_indices, ctl_values, ctx->n_ctl, ctx->dYctl));
    for (int i = 0; i < ctx->n_ctl; ++i) rows[i] = ctx->dYctl + (long long)i * ctx->Np;
  } else {
    for (int i = 0; i < ctx->n_ctl; ++i) rows[i] = ctx->dY + (long long)ctx->ctl_gene[i] * ctx->Np;
  }
  if (!ctx->d_ctl_rows) RET(dev_alloc(ctx, &ctx->d_ctl_rows, ctx->n_ctl));
  CK(cudaMemcpyAsync((void*)ctx->d_ctl_rows, rows.data(), rows.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_counts = true; ctx->hints_seeded = false; ctx->roworder_ready = false; ctx->hist_ready = false; ctx->dp_beta_warm = false;
  return RUVNB_OK;
}

// H0: Winsorize + ceiling on the resident counts (R/ruvIIInb.R:53), rowSums(Y > 0) (:57)
int ruvnb_winsorize(ruvnb_ctx* ctx, double* cap_out, int64_t* nonzero_out) {
  if (!ctx) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "winsorize: no counts uploaded");
  CK(cudaSetDevice(ctx->device));
  const int Gl = ctx->Gl;
  ScopedDev own;
  double* d_cap = nullptr; long long* d_nz = nullptr;
  CK(own.alloc(&d_cap, (size_t)Gl * 8));
  CK(own.alloc(&d_nz, (size_t)Gl * 8));
  if (ctx->cell_shard && (ctx->nranks > 1 || getenv("RUVNB_WZ_HIST"))) {   // (the variable: the same path on one rank, for tests)
    // cells over the ranks: the quantile of a gene needs the counts of ALL its cells.  Per-gene histograms of the values
    // 0..4094 (+ one overflow bin) are summed over the ranks; genes whose quantile falls into the overflow bin (counts in
    // the thousands for > 0.5 % of the cells) repeat the step on the value's upper bits, then on the lower 12 of that bucket
    unsigned* hist = nullptr; int* d_need = nullptr; unsigned *d_v[2] = {nullptr, nullptr}; long long* d_rank = nullptr;
    CK(own.alloc(&hist, (size_t)Gl * WZ_BINS * 4)); CK(own.alloc(&d_need, 4)); CK(own.alloc(&d_v[0], (size_t)Gl * 4)); CK(own.alloc(&d_v[1], (size_t)Gl * 4));
    CK(own.alloc(&d_rank, (size_t)Gl * 8));
    const long long n = ctx->N_total;
    CK(cudaMemsetAsync(hist, 0, (size_t)Gl * WZ_BINS * 4, ctx->stream));
    k_wz_hist<<<Gl, WZ_NT, 0, ctx->stream>>>(ctx->dY, ctx->Np, 0, WZ_BINS, 0u, nullptr, d_cap, 1, hist); CKL();
    RET(allreduce(ctx, hist, (size_t)Gl * WZ_BINS, NCCL_INT32));
    CK(cudaMemsetAsync(d_need, 0, 4, ctx->stream));
    k_wz_pick_direct<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, n, 0.995, hist, d_cap, d_need); CKL();
    CK(cudaMemcpyAsync(ctx->h_flags + 7, d_need, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_flags[7] > 0) {   // the same on every rank: the histograms are global
      const int shifts[3] = {20, 8, 0}, bits[3] = {12, 12, 8};
      for (int which = 0; which < 2; ++which) {
        k_wz_desc_init<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, n, 0.995, which, d_cap, d_v[which], d_rank); CKL();
        unsigned pmask = 0;
        for (int lv = 0; lv < 3; ++lv) {
          const int nb = 1 << bits[lv];
          CK(cudaMemsetAsync(hist, 0, (size_t)Gl * WZ_BINS * 4, ctx->stream));
          k_wz_hist<<<Gl, WZ_NT, 0, ctx->stream>>>(ctx->dY, ctx->Np, shifts[lv], nb, pmask, d_v[which], d_cap, 0, hist); CKL();
          RET(allreduce(ctx, hist, (size_t)Gl * WZ_BINS, NCCL_INT32));
          k_wz_desc_pick<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, hist, shifts[lv], nb, d_cap, d_v[which], d_rank); CKL();
          pmask |= (unsigned)(nb - 1) << shifts[lv];
        }
      }
      k_wz_desc_finish<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, n, 0.995, d_v[0], d_v[1], d_cap); CKL();
    }
    k_wz_apply<<<Gl, WZ_NT, 0, ctx->stream>>>(ctx->dY, ctx->Np, d_cap, d_nz); CKL();
    RET(allreduce(ctx, d_nz, Gl, NCCL_INT64));
  } else {
    k_winsorize<<<Gl, WZ_NT, 0, ctx->stream>>>(ctx->dY, ctx->Np, ctx->N, 0.995, d_cap, d_nz);
    CKL();
    if (ctx->dYctl) {  // the replicated control-gene rows get the same treatment (same rows, same caps)
      k_winsorize<<<ctx->n_ctl, WZ_NT, 0, ctx->stream>>>(ctx->dYctl, ctx->Np, ctx->N, 0.995, nullptr, nullptr);
      CKL();
    }
  }
  if (cap_out) CK(cudaMemcpyAsync(cap_out, d_cap, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (nonzero_out) CK(cudaMemcpyAsync(nonzero_out, d_nz, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->cur.sig_valid = false; ctx->roworder_ready = false; ctx->hist_ready = false; ctx->dp_beta_warm = false;
  return RUVNB_OK;
}

// the resident (winsorized) counts of the local rows, gene-major Gl x N: the `counts` element of the
// reference's return list (R/ruvIIInb.R:611,627)
int ruvnb_get_counts(ruvnb_ctx* ctx, int32_t* Y_out) {
  if (!ctx || !Y_out) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "get_counts: no counts uploaded");
  CK(cudaSetDevice(ctx->device));
  const long long N = ctx->N;
  long long slab = std::max(1ll, (256ll << 20) / (N * 4));
  slab = std::min<long long>(slab, ctx->Gl);
  int32_t* stage = nullptr;
  CK(cudaMalloc((void**)&stage, (size_t)slab * N * 4));
  for (long long r0 = 0; r0 < ctx->Gl; r0 += slab) {
    long long nr = std::min<long long>(slab, ctx->Gl - r0);
    k_unpermute_rows<<<nblk(nr * N, 256), 256, 0, ctx->stream>>>(ctx->dY + r0 * ctx->Np, nr, N, ctx->Np, ctx->d_cell2pos, stage);
    CKL();
    CK(cudaMemcpyAsync(Y_out + r0 * N, stage, (size_t)nr * N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaFree(stage));
  return RUVNB_OK;
}

// ---- cell shards, row mask, slab upload, cell gather (include/ruvnb.h) ---------------------------------------------------
int ruvnb_set_cell_shard(ruvnb_ctx* ctx, int64_t N_total, int64_t c0) {
  if (!ctx) return RUVNB_EARG;
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "set_cell_shard: call ruvnb_set_design first");
  if (ctx->Gl != ctx->G) return ctx->fail(RUVNB_EARG, "set_cell_shard: a cell shard holds ALL genes (set_design rows [0, G))");
  if (N_total < ctx->N || c0 < 0 || c0 + ctx->N > N_total) return ctx->fail(RUVNB_EARG, "set_cell_shard: cells [%lld, %lld) outside 0..%lld", (long long)c0, (long long)(c0 + ctx->N), (long long)N_total);
  ctx->cell_shard = true; ctx->N_total = N_total; ctx->cell0 = c0;
  return RUVNB_OK;
}

int ruvnb_set_row_mask(ruvnb_ctx* ctx, const uint8_t* alive) {
  if (!ctx) return RUVNB_EARG;
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "set_row_mask: call ruvnb_set_design first");
  CK(cudaSetDevice(ctx->device));
  if (!alive) { ctx->d_orow = nullptr; ctx->Gout = ctx->Gl; return RUVNB_OK; }   // the buffer stays in ctx->allocs
  std::vector<int> orow(ctx->Gl);
  int n = 0;
  for (int g = 0; g < ctx->Gl; ++g) orow[g] = alive[g] ? n++ : -1;
  if (n == 0) return ctx->fail(RUVNB_EARG, "set_row_mask: no row left");
  int* d = nullptr;
  RET(dev_upload(ctx, &d, orow));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->d_orow = d; ctx->Gout = n;
  return RUVNB_OK;
}

int ruvnb_upload_counts_rows(ruvnb_ctx* ctx, const int32_t* Y_rows, int on_device, int64_t row0, int64_t nrows) {
  if (!ctx || !Y_rows) return RUVNB_EARG;
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "upload_counts_rows: call ruvnb_set_design first");
  if (row0 < 0 || nrows <= 0 || row0 + nrows > ctx->Gl) return ctx->fail(RUVNB_EARG, "upload_counts_rows: rows [%lld, %lld) outside the %d local rows", (long long)row0, (long long)(row0 + nrows), ctx->Gl);
  if (ctx->Gl != ctx->G) return ctx->fail(RUVNB_EARG, "upload_counts_rows: gene shards take their control rows through ruvnb_upload_counts_dense");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->dY) RET(dev_alloc(ctx, &ctx->dY, (long long)ctx->Gl * ctx->Np));
  if (on_device) {
    k_permute_rows<<<nblk(nrows * ctx->Np, 256), 256, 0, ctx->stream>>>(Y_rows, nrows, ctx->N, ctx->N, 1, ctx->d_pos2cell, ctx->Np, ctx->dY + row0 * ctx->Np);
    CKL();
    CK(cudaStreamSynchronize(ctx->stream));   // the caller may reuse its slab
  } else {
    RET(upload_rows(ctx, Y_rows, RUVNB_LAYOUT_GENE_MAJOR, nrows, nullptr, 0, nrows, ctx->dY + row0 * ctx->Np));
  }
  return RUVNB_OK;
}
// A block of columns [c0, c0 + ncols) of a dgCMatrix -- what `Matrix::Matrix(Y, sparse = TRUE)` holds and what ruvIII.nb returns as
// `counts` (R/ruvIIInb.R:611-612): p (ncols + 1 offsets, p[0] need not be 0: a window of the caller's slot), i (0-based gene rows),
// x (doubles).  The first block (c0 == 0) zero-fills the resident matrix; blocks may arrive in any order after it.
int ruvnb_upload_counts_csc(ruvnb_ctx* ctx, const int32_t* p, const int32_t* i, const double* x, int64_t c0, int64_t ncols) {
  if (!ctx || !p || ncols < 0) return RUVNB_EARG;
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "upload_counts_csc: call ruvnb_set_design first");
  if (ctx->Gl != ctx->G) return ctx->fail(RUVNB_EARG, "upload_counts_csc: gene shards take their rows through ruvnb_upload_counts_csr");
  if (c0 < 0 || c0 + ncols > ctx->N) return ctx->fail(RUVNB_EARG, "upload_counts_csc: cells [%lld, %lld) outside 0..%lld", (long long)c0, (long long)(c0 + ncols), (long long)ctx->N);
  const long long nnz = ncols ? (long long)p[ncols] - p[0] : 0;
  if (nnz < 0 || (nnz && (!i || !x))) return ctx->fail(RUVNB_EARG, "upload_counts_csc: inconsistent p / i / x");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->dY) {
    if (c0 != 0) return ctx->fail(RUVNB_ESTATE, "upload_counts_csc: the first block must start at cell 0 (it zero-fills the matrix)");
    RET(dev_alloc(ctx, &ctx->dY, (long long)ctx->Gl * ctx->Np));
  }
  if (c0 == 0) { k_fill_pad<<<nblk((long long)ctx->Gl * ctx->Np, 256), 256, 0, ctx->stream>>>(ctx->dY, ctx->Gl, ctx->Np, ctx->d_pos2cell); CKL(); }
  if (ncols == 0) return RUVNB_OK;
  ScopedDev own;
  int *d_p = nullptr, *d_i = nullptr, *d_bad = nullptr; double* d_x = nullptr;
  CK(own.alloc(&d_p, (size_t)(ncols + 1) * 4)); CK(own.alloc(&d_i, (size_t)std::max(1ll, nnz) * 4)); CK(own.alloc(&d_x, (size_t)std::max(1ll, nnz) * 8));
  CK(own.alloc(&d_bad, 4));
  CK(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
  CK(cudaMemcpyAsync(d_p, p, (size_t)(ncols + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (nnz) {
    CK(cudaMemcpyAsync(d_i, i + p[0], (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_x, x + p[0], (size_t)nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  k_csc_scatter<<<nblk(ncols * 32, 256), 256, 0, ctx->stream>>>(d_p, d_i, d_x, ncols, c0, p[0], ctx->Gl, ctx->d_cell2pos, ctx->Np, ctx->dY, d_bad); CKL();
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (bad) return ctx->fail(RUVNB_EARG, "upload_counts_csc: %d entries with a row outside 0..%d or a negative / non-finite count", bad, ctx->Gl - 1);
  return RUVNB_OK;
}
int ruvnb_upload_counts_done(ruvnb_ctx* ctx) {
  if (!ctx) return RUVNB_EARG;
  if (!ctx->dY) return ctx->fail(RUVNB_ESTATE, "upload_counts_done: no slab uploaded");
  CK(cudaSetDevice(ctx->device));
  std::vector<const int32_t*> rows(ctx->n_ctl);
  for (int i = 0; i < ctx->n_ctl; ++i) rows[i] = ctx->dY + (long long)ctx->ctl_gene[i] * ctx->Np;
  if (!ctx->d_ctl_rows) RET(dev_alloc(ctx, &ctx->d_ctl_rows, ctx->n_ctl));
  CK(cudaMemcpyAsync((void*)ctx->d_ctl_rows, rows.data(), rows.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_counts = true; ctx->hints_seeded = false; ctx->roworder_ready = false; ctx->hist_ready = false; ctx->dp_beta_warm = false;
  return RUVNB_OK;
}

int ruvnb_get_cells(ruvnb_ctx* ctx, const int64_t* cells, int64_t n, int32_t* out) {
  if (!ctx || !cells || !out || n < 0) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "get_cells: no counts uploaded");
  if (n == 0) return RUVNB_OK;
  CK(cudaSetDevice(ctx->device));
  std::vector<int> pos(n);
  for (int64_t i = 0; i < n; ++i) {
    if (cells[i] < 0 || cells[i] >= ctx->N) return ctx->fail(RUVNB_EARG, "get_cells: cell %lld outside 0..%lld", (long long)cells[i], (long long)ctx->N - 1);
    pos[i] = ctx->cell2pos[cells[i]];
  }
  ScopedDev own;
  int* d_pos = nullptr; int32_t* d_out = nullptr;
  CK(own.alloc(&d_pos, (size_t)n * 4)); CK(own.alloc(&d_out, (size_t)n * ctx->Gl * 4));
  CK(cudaMemcpyAsync(d_pos, pos.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  k_gather_cells<<<dim3(nblk(ctx->Gl, 32), nblk(n, 32)), dim3(32, 8), 0, ctx->stream>>>(ctx->dY, ctx->Gl, ctx->Np, d_pos, n, d_out); CKL();
  CK(cudaMemcpyAsync(out, d_out, (size_t)n * ctx->Gl * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RUVNB_OK;
}

void* ruvnb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void ruvnb_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"

// ---- state ----------------------------------------------------------------------------------------
static int alloc_set(ruvnb_ctx* ctx, StateSet* s) {
  const long long Gl = ctx->Gl;
  RET(dev_alloc(ctx, &s->gmean, Gl));
  RET(dev_alloc(ctx, &s->Mb, Gl * ctx->m));
  RET(dev_alloc(ctx, &s->alpha, Gl * ctx->K));
  RET(dev_alloc(ctx, &s->adev, Gl * ctx->K));
  RET(dev_alloc(ctx, &s->W, ctx->Np * ctx->K));
  RET(dev_alloc(ctx, &s->psi, Gl * ctx->nbatch));     // [nbatch][Gl]
  RET(dev_alloc(ctx, &s->pi0, Gl));
  RET(dev_alloc(ctx, &s->sigma, Gl));
  RET(dev_alloc(ctx, &s->psi_w, Gl * ctx->nbatch));
  RET(dev_alloc(ctx, &s->wtctl, ctx->n_ctl));
  CK(cudaMemsetAsync(s->wtctl, 0, (size_t)ctx->n_ctl * 8, ctx->stream));
  k_fill<<<nblk(Gl * ctx->nbatch, 256), 256, 0, ctx->stream>>>(s->psi_w, Gl * ctx->nbatch, 1.0);
  CKL();
  CK(cudaMemsetAsync(s->sigma, 0, Gl * 8, ctx->stream));
  s->sig_valid = false;
  CK(cudaMemsetAsync(s->gmean, 0, Gl * 8, ctx->stream));
  CK(cudaMemsetAsync(s->Mb, 0, Gl * ctx->m * 8, ctx->stream));
  CK(cudaMemsetAsync(s->alpha, 0, Gl * ctx->K * 8, ctx->stream));
  CK(cudaMemsetAsync(s->adev, 0, Gl * ctx->K * 8, ctx->stream));
  CK(cudaMemsetAsync(s->W, 0, ctx->Np * ctx->K * 8, ctx->stream));
  CK(cudaMemsetAsync(s->pi0, 0, Gl * 8, ctx->stream));
  k_fill<<<nblk(Gl * ctx->nbatch, 256), 256, 0, ctx->stream>>>(s->psi, Gl * ctx->nbatch, 1.0);
  CKL();
  return RUVNB_OK;
}

int ensure_state(ruvnb_ctx* ctx, int K) {
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "call ruvnb_set_design first");
  if (K < 1 || K > RUVNB_MAX_K) return ctx->fail(RUVNB_EARG, "k = %d outside 1..%d", K, RUVNB_MAX_K);
  if (ctx->K == K) return RUVNB_OK;
  if (ctx->K != 0) return ctx->fail(RUVNB_ESTATE, "k changed from %d to %d: one context holds one problem", ctx->K, K);
  CK(cudaSetDevice(ctx->device));
  ctx->K = K;
  const long long Gl = ctx->Gl, m = ctx->m;
  const int KK = K * (K + 1) / 2;
  RET(alloc_set(ctx, &ctx->cur));
  RET(alloc_set(ctx, &ctx->nw));
  RET(alloc_set(ctx, &ctx->best));
  RET(dev_alloc(ctx, &ctx->step, Gl));
  RET(dev_alloc(ctx, &ctx->loglik, Gl));
  RET(dev_alloc(ctx, &ctx->loglik_tmp, Gl));
  k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->step, Gl, 1.0);
  CKL();
  CK(cudaMemsetAsync(ctx->loglik, 0, Gl * 8, ctx->stream));
  CK(cudaMemsetAsync(ctx->loglik_tmp, 0, Gl * 8, ctx->stream));
  // row parts: enough CTAs for >= 6 waves of 296 (2 CTAs x 148 SMs) so that a small gene shard does not leave most
  // of a wave empty (20000 rows on one GPU: 1 part; 2500 rows per GPU on eight: 8 parts)
  {
    // cost model of a sweep: waves of 296 resident CTAs x tiles per part; the smallest product wins (at most 24 parts:
    // the part-major partial sums grow with P).  20000 rows: 1 part (8.4 waves); 2500 rows: 17 parts (18.0 waves of 12 tiles)
    const long long ctas = (Gl + NWARP - 1) / NWARP;
    const long long ntiles = ctx->nchunks / TileCfg_SC(K);
    int P = 1;
    if (ctas < 6 * 296) {
      long long best = -1;
      for (int p = 1; p <= 24 && p <= ntiles; ++p) {
        const long long waves = (ctas * p + 295) / 296, tpp = (ntiles + p - 1) / p;
        const long long cost = waves * tpp * 16 + p;   // ties -> fewer parts
        if (best < 0 || cost < best) { best = cost; P = p; }
      }
    }
    if (const char* e = getenv("RUVNB_NPARTS")) { int v = atoi(e); if (v >= 1 && v <= 24) P = v; }
    ctx->nparts = P;
  }
  const long long P = ctx->nparts;
  RET(dev_alloc(ctx, &ctx->A, P * Gl * KK));
  RET(dev_alloc(ctx, &ctx->u, P * Gl * K));
  RET(dev_alloc(ctx, &ctx->sw, P * Gl));
  RET(dev_alloc(ctx, &ctx->ZS, P * Gl * m));
  RET(dev_alloc(ctx, &ctx->VS, P * Gl * m * K));
  RET(dev_alloc(ctx, &ctx->llparts, P * Gl));
  RET(dev_alloc(ctx, &ctx->certc, P * Gl * 4));
  RET(dev_alloc(ctx, &ctx->cvec, Gl * K));
  RET(dev_alloc(ctx, &ctx->colsum_part, 64 * 64));
  RET(dev_alloc(ctx, &ctx->Ab, Gl * (KK + K)));
  RET(dev_alloc(ctx, &ctx->red, 64));
  RET(dev_alloc(ctx, &ctx->amean, K));
  RET(dev_alloc(ctx, &ctx->S0, P * Gl * m));
  RET(dev_alloc(ctx, &ctx->S1, P * Gl * m));
  RET(dev_alloc(ctx, &ctx->sigma_exact, Gl));
  RET(dev_alloc(ctx, &ctx->hint, Gl * 3));
  k_fill<<<nblk(Gl * 3, 256), 256, 0, ctx->stream>>>(ctx->hint, Gl * 3, -1.0);  // tg < 0: no hint yet
  CKL();
  RET(dev_alloc(ctx, &ctx->tabL, Gl * ctx->nbatch * YT));   // [nbatch][Gl][YT]
  RET(dev_alloc(ctx, &ctx->tabR, Gl * ctx->nbatch * YT));
  RET(dev_alloc(ctx, &ctx->lgr, Gl * ctx->nbatch));
  RET(dev_alloc(ctx, &ctx->ctl_tabR, (long long)ctx->n_ctl * ctx->nbatch * YT));
  RET(dev_alloc(ctx, &ctx->w_fail, ctx->Np + 1));
  RET(dev_alloc(ctx, &ctx->cert_diag, 8));
  CK(cudaMemsetAsync(ctx->cert_diag, 0, 8 * sizeof(int), ctx->stream));
  CK(cudaMemsetAsync(ctx->w_fail, 0, sizeof(int), ctx->stream));
  {
    MathTabs mt;
    build_math_tabs(&mt);
    RET(dev_alloc(ctx, &ctx->d_mtabs, 1));
    CK(cudaMemcpyAsync(ctx->d_mtabs, &mt, sizeof(MathTabs), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // mt is a stack object
  }
  RET(dev_alloc(ctx, &ctx->Wbar, m * K));
  RET(dev_alloc(ctx, &ctx->RMW, ctx->Np * K));
  RET(dev_alloc(ctx, &ctx->sumsq, K));
  RET(dev_alloc(ctx, &ctx->med, K));
  RET(dev_alloc(ctx, &ctx->mad, K));
  RET(dev_alloc(ctx, &ctx->ctlpack, (long long)ctx->n_ctl * (3 + 2 * ctx->nbatch + 2 * K + m)));
  RET(dev_alloc(ctx, &ctx->scal, 16));
  if (ctx->nranks > 1) {
    RET(dev_alloc(ctx, &ctx->alpha_full, ctx->G * K + 1));
    const long long per = (ctx->G + ctx->nranks - 1) / ctx->nranks;
    RET(dev_alloc(ctx, &ctx->agather, (long long)ctx->nranks * (2 * per * K + 1)));
  }
  return RUVNB_OK;
}

enum StateKind { SK_VEC, SK_GK, SK_GM, SK_W, SK_CELL, SK_CTL, SK_GB };   // SK_GB: Gl x nbatch, column-major on both sides
static int state_lookup(ruvnb_ctx* ctx, const char* name, double** p, StateKind* kind) {
  StateSet& s = ctx->cur;
  if (!strcmp(name, "W")) { *p = s.W; *kind = SK_W; }
  else if (!strcmp(name, "alpha")) { *p = s.alpha; *kind = SK_GK; }
  else if (!strcmp(name, "alpha_dev")) { *p = s.adev; *kind = SK_GK; }
  else if (!strcmp(name, "gmean")) { *p = s.gmean; *kind = SK_VEC; }
  else if (!strcmp(name, "Mb")) { *p = s.Mb; *kind = SK_GM; }
  else if (!strcmp(name, "psi")) { *p = s.psi; *kind = SK_GB; }
  else if (!strcmp(name, "pi0")) { *p = s.pi0; *kind = SK_VEC; }
  else if (!strcmp(name, "psi_w")) { *p = s.psi_w; *kind = SK_GB; }
  else if (!strcmp(name, "step")) { *p = ctx->step; *kind = SK_VEC; }
  else if (!strcmp(name, "loglik")) { *p = ctx->loglik; *kind = SK_VEC; }
  else if (!strcmp(name, "sigma")) { *p = ctx->sigma_exact; *kind = SK_VEC; }
  else if (!strcmp(name, "wt_ctl")) { *p = s.wtctl; *kind = SK_CTL; }
  else if (!strcmp(name, "cell_loglik") && ctx->ff_cll) { *p = ctx->ff_cll; *kind = SK_CELL; }
  else if (!strcmp(name, "cell_step") && ctx->ff_step) { *p = ctx->ff_step; *kind = SK_CELL; }
  else return ctx->fail(RUVNB_EARG, "unknown state name '%s'", name);
  return RUVNB_OK;
}

extern "C" {

int ruvnb_set_state(ruvnb_ctx* ctx, const char* name, const double* host, int k) {
  if (!ctx || !name || !host) return RUVNB_EARG;
  if (k > 0) RET(ensure_state(ctx, k));
  if (ctx->K == 0) return ctx->fail(RUVNB_ESTATE, "set_state: k unknown -- set \"W\" first (with k)");
  CK(cudaSetDevice(ctx->device));
  double* p; StateKind kind;
  RET(state_lookup(ctx, name, &p, &kind));
  const long long Gl = ctx->Gl, K = ctx->K, m = ctx->m, N = ctx->N, Np = ctx->Np;
  std::vector<double> tmp;
  switch (kind) {
    case SK_VEC: tmp.assign(host, host + Gl); break;
    case SK_GB: tmp.assign(host, host + Gl * ctx->nbatch); break;
    case SK_GK: tmp.resize(Gl * K); for (long long g = 0; g < Gl; ++g) for (long long l = 0; l < K; ++l) tmp[g * K + l] = host[g + Gl * l]; break;
    case SK_GM: tmp.resize(Gl * m); for (long long g = 0; g < Gl; ++g) for (long long j = 0; j < m; ++j) tmp[g * m + j] = host[g + Gl * j]; break;
    case SK_W:
      tmp.assign(Np * K, 0.0);
      for (long long l = 0; l < K; ++l) for (long long c = 0; c < N; ++c) tmp[l * Np + ctx->cell2pos[c]] = host[c + N * l];
      ctx->have_W = true;
      break;
    case SK_CELL:
      tmp.assign(Np, 0.0);
      for (long long c = 0; c < N; ++c) tmp[ctx->cell2pos[c]] = host[c];
      break;
    case SK_CTL: tmp.assign(host, host + ctx->n_ctl); break;
  }
  CK(cudaMemcpyAsync(p, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->cur.sig_valid = false;                  // the Huber scale belongs to the parameters it was computed for
  ctx->best_eq_cur = false;
  if (!strcmp(name, "psi")) ctx->psi_ver++;    // count tables follow psi
  return RUVNB_OK;
}

int ruvnb_get_state(ruvnb_ctx* ctx, const char* name, double* host) {
  if (!ctx || !name || !host) return RUVNB_EARG;
  if (ctx->K == 0) return ctx->fail(RUVNB_ESTATE, "get_state: no state yet");
  CK(cudaSetDevice(ctx->device));
  double* p; StateKind kind;
  RET(state_lookup(ctx, name, &p, &kind));
  const long long Gl = ctx->Gl, K = ctx->K, m = ctx->m, N = ctx->N, Np = ctx->Np;
  long long n = kind == SK_VEC ? Gl : kind == SK_GB ? Gl * ctx->nbatch : kind == SK_GK ? Gl * K : kind == SK_GM ? Gl * m : kind == SK_CELL ? Np : kind == SK_CTL ? ctx->n_ctl : Np * K;
  std::vector<double> tmp(n);
  CK(cudaMemcpyAsync(tmp.data(), p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  switch (kind) {
    case SK_VEC: memcpy(host, tmp.data(), Gl * 8); break;
    case SK_GB: memcpy(host, tmp.data(), (size_t)Gl * ctx->nbatch * 8); break;
    case SK_GK: for (long long g = 0; g < Gl; ++g) for (long long l = 0; l < K; ++l) host[g + Gl * l] = tmp[g * K + l]; break;
    case SK_GM: for (long long g = 0; g < Gl; ++g) for (long long j = 0; j < m; ++j) host[g + Gl * j] = tmp[g * m + j]; break;
    case SK_W: for (long long l = 0; l < K; ++l) for (long long c = 0; c < N; ++c) host[c + N * l] = tmp[l * Np + ctx->cell2pos[c]]; break;
    case SK_CELL: for (long long c = 0; c < N; ++c) host[c] = tmp[ctx->cell2pos[c]]; break;
    case SK_CTL: memcpy(host, tmp.data(), (size_t)ctx->n_ctl * 8); break;
  }
  return RUVNB_OK;
}

}  // extern "C"
int check_ready(ruvnb_ctx* ctx, const ruvnb_opts* o) {
  if (!ctx || !o) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "no counts uploaded");
  RET(ensure_state(ctx, o->k));
  if (!ctx->have_W) return ctx->fail(RUVNB_ESTATE, "W not set: the initial W (irlba in the reference, R/ruvIIInb.R:167) is injected with ruvnb_set_state(\"W\")");
  if (ctx->empty_group >= 0) return ctx->fail(RUVNB_EARG, "set_design: replicate group %d has no cell", ctx->empty_group);
  if (ctx->cell_shard && ctx->nranks > 1) return ctx->fail(RUVNB_ESTATE, "this context holds a CELL shard: the IRLS / dispersion operators shard by gene (ruvnb_set_design rows)");
  ctx->zinb_on = o->zinb && ctx->any_zeroinf;   // ZINB only matters for cells flagged zeroinf (:237,254)
  ctx->launch_parts = 1;                         // a failed call must not leave a sweep's part count behind
  CK(cudaSetDevice(ctx->device));
  return RUVNB_OK;
}

// launch order of the rows (load balance of the sweeps): rows with many counts beyond the tables go first
int ensure_roworder(ruvnb_ctx* ctx) {
  if (ctx->roworder_ready) return RUVNB_OK;
  const int Gl = ctx->Gl;
  int* d_nbig = nullptr;
  CK(cudaMalloc((void**)&d_nbig, (size_t)Gl * 4));
  k_count_big<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(ctx->dY, Gl, ctx->Np, d_nbig); CKL();
  std::vector<int> nbig(Gl), order(Gl);
  CK(cudaMemcpyAsync(nbig.data(), d_nbig, (size_t)Gl * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(d_nbig));
  for (int i = 0; i < Gl; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nbig[a] > nbig[b]; });
  if (!ctx->d_roworder) RET(dev_alloc(ctx, &ctx->d_roworder, Gl));
  CK(cudaMemcpyAsync(ctx->d_roworder, order.data(), (size_t)Gl * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->roworder_ready = true;
  return RUVNB_OK;
}

// per-gene count tables for the dispersions in `psi` (rebuilt only when psi changed)
int ensure_ytabs(ruvnb_ctx* ctx, const double* psi) {
  RET(ensure_roworder(ctx));
  if (ctx->ytab_ver == ctx->psi_ver && ctx->ytab_psi != nullptr) return RUVNB_OK;
  k_build_ytabs<<<ctx->Gl * ctx->nbatch, YT, 0, ctx->stream>>>(psi, ctx->Gl * ctx->nbatch, ctx->tabL, ctx->tabR, ctx->lgr);
  CKL();
  ctx->ytab_ver = ctx->psi_ver; ctx->ytab_psi = psi;
  return RUVNB_OK;
}
