// ruvnb.cu -- C ABI of libruvnb_b200 (include/ruvnb.h): context, data upload, the NB IRLS operators
// and the whole-fit driver.  sm_100a only; no CPU fallback anywhere in this file.
#include <stdlib.h>
#include <chrono>
#include "ctx.cuh"
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
  for (int j = 0; j < m; ++j)
    if (ctx->group_n[j] == 0) return ctx->fail(RUVNB_EARG, "set_design: replicate group %d has no cell", j);
  ctx->group_chunk0.assign(m, 0);
  int nch = 0;
  for (int j = 0; j < m; ++j) { ctx->group_chunk0[j] = nch; nch += (ctx->group_n[j] + CH - 1) / CH; }
  const int nch_real = nch;
  nch = (nch + CHPAD - 1) / CHPAD * CHPAD;  // whole tiles: trailing chunks are empty (valid = 0)
  ctx->nchunks = nch; ctx->Np = (long long)nch * CH;
  ctx->chunk_grp.assign(nch, ctx->m - 1); ctx->chunk_valid.assign(nch, 0);
  (void)nch_real;
  for (int j = 0; j < m; ++j) {
    int nc = (ctx->group_n[j] + CH - 1) / CH;
    for (int q = 0; q < nc; ++q) {
      ctx->chunk_grp[ctx->group_chunk0[j] + q] = j;
      ctx->chunk_valid[ctx->group_chunk0[j] + q] = std::min(CH, ctx->group_n[j] - q * CH);
    }
  }
  ctx->pos2cell.assign(ctx->Np, -1); ctx->cell2pos.assign(N, 0);
  std::vector<long long> fill(m);
  for (int j = 0; j < m; ++j) fill[j] = (long long)ctx->group_chunk0[j] * CH;
  for (long long c = 0; c < N; ++c) {  // stable: cells keep their order inside a group
    long long p = fill[ctx->grp[c]]++;
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
  RET(dev_upload(ctx, &ctx->d_group_chunk0, ctx->group_chunk0));
  RET(dev_upload(ctx, &ctx->d_group_n, ctx->group_n));
  RET(dev_upload(ctx, &ctx->d_ctl_local, ctl_local));
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

int ruvnb_upload_counts_dense(ruvnb_ctx* ctx, const int32_t* Y, int layout, const int32_t* Yctl) {
  if (!ctx || !Y) return RUVNB_EARG;
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "upload_counts: call ruvnb_set_design first");
  if (layout != RUVNB_LAYOUT_GENE_MAJOR && layout != RUVNB_LAYOUT_CELL_MAJOR) return ctx->fail(RUVNB_EARG, "upload_counts: bad layout %d", layout);
  CK(cudaSetDevice(ctx->device));
  const bool sharded = (ctx->Gl != ctx->G);
  if (sharded && !Yctl) return ctx->fail(RUVNB_EARG, "upload_counts: a gene shard needs the control-gene rows (Yctl)");
  if (!ctx->dY) RET(dev_alloc(ctx, &ctx->dY, (long long)ctx->Gl * ctx->Np));
  // Y holds the LOCAL rows (Gl x N)
  RET(upload_rows(ctx, Y, layout, ctx->Gl, nullptr, 0, ctx->Gl, ctx->dY));
  std::vector<const int32_t*> rows(ctx->n_ctl);
  if (sharded) {
    if (!ctx->dYctl) RET(dev_alloc(ctx, &ctx->dYctl, (long long)ctx->n_ctl * ctx->Np));
    RET(upload_rows(ctx, Yctl, layout, ctx->n_ctl, nullptr, 0, ctx->n_ctl, ctx->dYctl));
    for (int i = 0; i < ctx->n_ctl; ++i) rows[i] = ctx->dYctl + (long long)i * ctx->Np;
  } else {
    for (int i = 0; i < ctx->n_ctl; ++i) rows[i] = ctx->dY + (long long)ctx->ctl_gene[i] * ctx->Np;
  }
  if (!ctx->d_ctl_rows) RET(dev_alloc(ctx, &ctx->d_ctl_rows, ctx->n_ctl));
  CK(cudaMemcpyAsync((void*)ctx->d_ctl_rows, rows.data(), rows.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_counts = true; ctx->roworder_ready = false; ctx->hist_ready = false; ctx->dp_beta_warm = false;
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
  ctx->have_counts = true; ctx->roworder_ready = false; ctx->hist_ready = false; ctx->dp_beta_warm = false;
  return RUVNB_OK;
}

// H0: Winsorize + ceiling on the resident counts (R/ruvIIInb.R:53), rowSums(Y > 0) (:57)
int ruvnb_winsorize(ruvnb_ctx* ctx, double* cap_out, int64_t* nonzero_out) {
  if (!ctx) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "winsorize: no counts uploaded");
  CK(cudaSetDevice(ctx->device));
  const int Gl = ctx->Gl;
  double* d_cap = nullptr; long long* d_nz = nullptr;
  CK(cudaMalloc((void**)&d_cap, (size_t)Gl * 8));
  CK(cudaMalloc((void**)&d_nz, (size_t)Gl * 8));
  k_winsorize<<<Gl, WZ_NT, 0, ctx->stream>>>(ctx->dY, ctx->Np, ctx->N, 0.995, d_cap, d_nz);
  CKL();
  if (ctx->dYctl) {  // the replicated control-gene rows get the same treatment (same rows, same caps)
    k_winsorize<<<ctx->n_ctl, WZ_NT, 0, ctx->stream>>>(ctx->dYctl, ctx->Np, ctx->N, 0.995, nullptr, nullptr);
    CKL();
  }
  if (cap_out) CK(cudaMemcpyAsync(cap_out, d_cap, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (nonzero_out) CK(cudaMemcpyAsync(nonzero_out, d_nz, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(d_cap)); CK(cudaFree(d_nz));
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

}  // extern "C"

// ---- state ----------------------------------------------------------------------------------------
static int alloc_set(ruvnb_ctx* ctx, StateSet* s) {
  const long long Gl = ctx->Gl;
  RET(dev_alloc(ctx, &s->gmean, Gl));
  RET(dev_alloc(ctx, &s->Mb, Gl * ctx->m));
  RET(dev_alloc(ctx, &s->alpha, Gl * ctx->K));
  RET(dev_alloc(ctx, &s->adev, Gl * ctx->K));
  RET(dev_alloc(ctx, &s->W, ctx->Np * ctx->K));
  RET(dev_alloc(ctx, &s->psi, Gl));
  RET(dev_alloc(ctx, &s->pi0, Gl));
  RET(dev_alloc(ctx, &s->sigma, Gl));
  RET(dev_alloc(ctx, &s->psi_w, Gl));
  RET(dev_alloc(ctx, &s->wtctl, ctx->n_ctl));
  CK(cudaMemsetAsync(s->wtctl, 0, (size_t)ctx->n_ctl * 8, ctx->stream));
  k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(s->psi_w, Gl, 1.0);
  CKL();
  CK(cudaMemsetAsync(s->sigma, 0, Gl * 8, ctx->stream));
  s->sig_valid = false;
  CK(cudaMemsetAsync(s->gmean, 0, Gl * 8, ctx->stream));
  CK(cudaMemsetAsync(s->Mb, 0, Gl * ctx->m * 8, ctx->stream));
  CK(cudaMemsetAsync(s->alpha, 0, Gl * ctx->K * 8, ctx->stream));
  CK(cudaMemsetAsync(s->adev, 0, Gl * ctx->K * 8, ctx->stream));
  CK(cudaMemsetAsync(s->W, 0, ctx->Np * ctx->K * 8, ctx->stream));
  CK(cudaMemsetAsync(s->pi0, 0, Gl * 8, ctx->stream));
  k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(s->psi, Gl, 1.0);
  CKL();
  return RUVNB_OK;
}

static int ensure_state(ruvnb_ctx* ctx, int K) {
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
    const long long ctas = (Gl + NWARP - 1) / NWARP;
    int P = 1;
    while (P < 8 && ctas * P < 6 * 296) P *= 2;
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
  RET(dev_alloc(ctx, &ctx->Ab, Gl * (KK + K)));
  RET(dev_alloc(ctx, &ctx->red, 64));
  RET(dev_alloc(ctx, &ctx->amean, K));
  RET(dev_alloc(ctx, &ctx->S0, P * Gl * m));
  RET(dev_alloc(ctx, &ctx->S1, P * Gl * m));
  RET(dev_alloc(ctx, &ctx->sigma_exact, Gl));
  RET(dev_alloc(ctx, &ctx->hint, Gl * 3));
  k_fill<<<nblk(Gl * 3, 256), 256, 0, ctx->stream>>>(ctx->hint, Gl * 3, -1.0);  // tg < 0: no hint yet
  CKL();
  RET(dev_alloc(ctx, &ctx->tabL, Gl * YT));
  RET(dev_alloc(ctx, &ctx->tabR, Gl * YT));
  RET(dev_alloc(ctx, &ctx->lgr, Gl));
  RET(dev_alloc(ctx, &ctx->ctl_tabR, (long long)ctx->n_ctl * YT));
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
  RET(dev_alloc(ctx, &ctx->ctlpack, (long long)ctx->n_ctl * (5 + 2 * K + m)));
  RET(dev_alloc(ctx, &ctx->scal, 16));
  if (ctx->nranks > 1) RET(dev_alloc(ctx, &ctx->alpha_full, ctx->G * K));
  return RUVNB_OK;
}

enum StateKind { SK_VEC, SK_GK, SK_GM, SK_W, SK_CELL, SK_CTL };
static int state_lookup(ruvnb_ctx* ctx, const char* name, double** p, StateKind* kind) {
  StateSet& s = ctx->cur;
  if (!strcmp(name, "W")) { *p = s.W; *kind = SK_W; }
  else if (!strcmp(name, "alpha")) { *p = s.alpha; *kind = SK_GK; }
  else if (!strcmp(name, "alpha_dev")) { *p = s.adev; *kind = SK_GK; }
  else if (!strcmp(name, "gmean")) { *p = s.gmean; *kind = SK_VEC; }
  else if (!strcmp(name, "Mb")) { *p = s.Mb; *kind = SK_GM; }
  else if (!strcmp(name, "psi")) { *p = s.psi; *kind = SK_VEC; }
  else if (!strcmp(name, "pi0")) { *p = s.pi0; *kind = SK_VEC; }
  else if (!strcmp(name, "psi_w")) { *p = s.psi_w; *kind = SK_VEC; }
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
  long long n = kind == SK_VEC ? Gl : kind == SK_GK ? Gl * K : kind == SK_GM ? Gl * m : kind == SK_CELL ? Np : kind == SK_CTL ? ctx->n_ctl : Np * K;
  std::vector<double> tmp(n);
  CK(cudaMemcpyAsync(tmp.data(), p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  switch (kind) {
    case SK_VEC: memcpy(host, tmp.data(), Gl * 8); break;
    case SK_GK: for (long long g = 0; g < Gl; ++g) for (long long l = 0; l < K; ++l) host[g + Gl * l] = tmp[g * K + l]; break;
    case SK_GM: for (long long g = 0; g < Gl; ++g) for (long long j = 0; j < m; ++j) host[g + Gl * j] = tmp[g * m + j]; break;
    case SK_W: for (long long l = 0; l < K; ++l) for (long long c = 0; c < N; ++c) host[c + N * l] = tmp[l * Np + ctx->cell2pos[c]]; break;
    case SK_CELL: for (long long c = 0; c < N; ++c) host[c] = tmp[ctx->cell2pos[c]]; break;
    case SK_CTL: memcpy(host, tmp.data(), (size_t)ctx->n_ctl * 8); break;
  }
  return RUVNB_OK;
}

}  // extern "C"

// ---- launch helpers ---------------------------------------------------------------------------------
static RowGeom geom(ruvnb_ctx* ctx, int nrows, const int* rowlist) {
  RowGeom g;
  g.Yp = ctx->dY; g.Np = ctx->Np; g.nchunks = ctx->nchunks; g.chunk_grp = ctx->d_chunk_grp; g.chunk_valid = ctx->d_chunk_valid;
  g.nrows = nrows; g.rowlist = rowlist; g.nrows_dev = nullptr; g.nparts = ctx->launch_parts; g.Gl = ctx->Gl;
  if (!rowlist && nrows == ctx->Gl && ctx->roworder_ready) g.rowlist = ctx->d_roworder;   // full sweeps: heavy rows first
  g.mtabs = ctx->d_mtabs; g.tabL = ctx->tabL; g.tabR = ctx->tabR; g.lgr = ctx->lgr;
  g.zinf = ctx->zinb_on ? ctx->d_zinf : nullptr;
  return g;
}
template <int K>
static GeneState<K> gstate(ruvnb_ctx* ctx, const StateSet& s, bool with_step) {
  GeneState<K> st;
  st.gmean = s.gmean; st.Mb = s.Mb; st.alpha = s.alpha; st.psi = s.psi; st.step = with_step ? ctx->step : nullptr; st.W = s.W; st.m = ctx->m;
  st.pi0 = s.pi0; st.psi_w = s.psi_w; st.zinb = ctx->zinb_on ? 1 : 0;
  return st;
}
#define P1_NW 8   // rows per CTA of k_pass1 (register budget, see kernels.cuh)
template <int K>
static inline size_t tile_smem(int nsets) { return (size_t)2 * nsets * K * TileCfg<K>::CELLS * sizeof(double); }
static inline unsigned row_grid(int nrows) { return (unsigned)((nrows + NWARP - 1) / NWARP); }
// row-streaming launch: the double-buffered W tiles exceed the 48 KB default dynamic shared memory
#define LAUNCH_ROWS_NW(KERNEL, NW_, NSETS, NROWS, ...)                                                           \
  do {                                                                                                           \
    const size_t smem_ = tile_smem<KT>(NSETS);                                                                   \
    CK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));                   \
    KERNEL<<<(unsigned)(((NROWS) + (NW_) - 1) / (NW_)) * ctx->launch_parts, (NW_) * 32, smem_, ctx->stream>>>(__VA_ARGS__); \
    CKL();                                                                                                       \
  } while (0)
#define LAUNCH_ROWS(KERNEL, NSETS, NROWS, ...)                                                                   \
  do {                                                                                                           \
    const size_t smem_ = tile_smem<KT>(NSETS);                                                                   \
    CK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));                   \
    KERNEL<<<row_grid(NROWS) * ctx->launch_parts, ROW_THREADS, smem_, ctx->stream>>>(__VA_ARGS__);               \
    CKL();                                                                                                       \
  } while (0)

static int check_ready(ruvnb_ctx* ctx, const ruvnb_opts* o) {
  if (!ctx || !o) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "no counts uploaded");
  RET(ensure_state(ctx, o->k));
  if (!ctx->have_W) return ctx->fail(RUVNB_ESTATE, "W not set: the initial W (irlba in the reference, R/ruvIIInb.R:167) is injected with ruvnb_set_state(\"W\")");
  ctx->zinb_on = o->zinb && ctx->any_zeroinf;   // ZINB only matters for cells flagged zeroinf (:237,254)
  ctx->launch_parts = 1;                         // a failed call must not leave a sweep's part count behind
  CK(cudaSetDevice(ctx->device));
  return RUVNB_OK;
}

// launch order of the rows (load balance of the sweeps): rows with many counts beyond the tables go first
static int ensure_roworder(ruvnb_ctx* ctx) {
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
static int ensure_ytabs(ruvnb_ctx* ctx, const double* psi) {
  RET(ensure_roworder(ctx));
  if (ctx->ytab_ver == ctx->psi_ver && ctx->ytab_psi != nullptr) return RUVNB_OK;
  k_build_ytabs<<<ctx->Gl, YT, 0, ctx->stream>>>(psi, ctx->Gl, ctx->tabL, ctx->tabR, ctx->lgr);
  CKL();
  ctx->ytab_ver = ctx->psi_ver; ctx->ytab_psi = psi;
  return RUVNB_OK;
}

// per-gene log-likelihood of parameter set `s` into `out` (R/ruvIIInb.R:276-291) and, in the same sweep, the
// Huber certificate of `s` (s.sigma: +Inf certified / 0 to be evaluated exactly); total (all ranks) into
// *total if non-null
static int loglik_of(ruvnb_ctx* ctx, const ruvnb_opts* o, StateSet& s, double* out, double* total) {
  const double kh = o->robust ? 1.345 : 100.0;  // :155
  RET(ensure_ytabs(ctx, s.psi));
  const int P = ctx->nparts;
  KEV_BEGIN(4);
  ctx->launch_parts = P;
  DISPATCH_K(ctx->K, {
    LAUNCH_ROWS(k_loglik<KT>, 1, ctx->Gl, geom(ctx, ctx->Gl, nullptr), gstate<KT>(ctx, s, false), P > 1 ? ctx->llparts : out, ctx->hint, ctx->certc,
                o->huber_fastpath);
  });
  ctx->launch_parts = 1;
  if (P > 1) { k_sum_parts<<<nblk(ctx->Gl, 256), 256, 0, ctx->stream>>>(out, ctx->llparts, ctx->Gl, P); CKL(); }
  k_cert_finish<<<nblk(ctx->Gl, 128), 128, 0, ctx->stream>>>(ctx->Gl, P, ctx->certc, ctx->hint, s.sigma, kh, o->huber_fastpath, (long long)ctx->N, ctx->cert_diag);
  CKL();
  KEV_END(4);
  s.sig_valid = true; s.maybe_active = false;
  if (total) {
    k_colsum<<<1, 1024, 0, ctx->stream>>>(out, ctx->Gl, 1, ctx->scal);
    CKL();
    RET(allreduce(ctx, ctx->scal, 1, NCCL_F64));
    CK(cudaMemcpyAsync(ctx->h_scal, ctx->scal, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    kev_collect(ctx);
    *total = ctx->h_scal[0];
  }
  return RUVNB_OK;
}

// complete s.sigma: rows the certificate sweep left at 0 get the exact mad/0.6745 (R/ruvIIInb.R:713) -- or +Inf
// when that exact value proves every weight is 1.  *nfail = rows evaluated exactly, *nactive = rows whose
// Huber weights are live (finite sigma) afterwards.
static int complete_sigma(ruvnb_ctx* ctx, const ruvnb_opts* o, StateSet& s, double kh, int* nfail, int* nactive) {
  const int Gl = ctx->Gl;
  k_list_failed<<<1, 1024, 0, ctx->stream>>>(Gl, s.sigma, ctx->faillist, ctx->flags + 4, nullptr);
  CKL();
  CK(cudaMemcpyAsync(ctx->h_flags + 4, ctx->flags + 4, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  kev_collect(ctx);
  const int nf = ctx->h_flags[4];
  *nfail = nf;
  if (nf > 0) {
    // exact path in batches of rows through the deviance scratch (<= 2 GB)
    long long batch = std::max(8ll, (2ll << 30) / (ctx->Np * 8));
    batch = std::min<long long>(batch, Gl);
    if (ctx->D_rows < std::min<long long>(batch, nf)) {
      // first call sizes the scratch for a full batch so that it never grows again
      RET(dev_alloc(ctx, &ctx->D, batch * ctx->Np));
      RET(dev_alloc(ctx, &ctx->rowmax, batch));
      RET(dev_alloc(ctx, &ctx->rownan, batch));
      ctx->D_rows = batch;
    }
    KEV_BEGIN(5);
    for (long long r0 = 0; r0 < nf; r0 += ctx->D_rows) {
      int nr = (int)std::min<long long>(ctx->D_rows, nf - r0);
      CK(cudaMemsetAsync(ctx->rowmax, 0, (size_t)nr * 8, ctx->stream));   // combined across the parts with atomicMax / atomicOr
      CK(cudaMemsetAsync(ctx->rownan, 0, (size_t)nr * 4, ctx->stream));
      {
        int Ps = 1;
        while (Ps < 8 && (long long)((nr + NWARP - 1) / NWARP) * Ps < 2 * 296) Ps *= 2;   // few failing rows: cut them into parts
        ctx->launch_parts = Ps;
      }
      DISPATCH_K(ctx->K, {
        LAUNCH_ROWS(k_signdev<KT>, 1, nr, geom(ctx, nr, ctx->faillist + r0), gstate<KT>(ctx, s, false), ctx->D, ctx->rowmax, ctx->rownan);
      });
      ctx->launch_parts = 1;
      k_row_sigma<<<nr, RS_NT, 0, ctx->stream>>>(ctx->D, ctx->Np, (long long)ctx->N, ctx->faillist + r0, ctx->rowmax, ctx->rownan, kh,
                                                 o->huber_fastpath, s.sigma, ctx->sigma_exact, ctx->hint);
      CKL();
    }
    KEV_END(5);
  }
  if (nf > 0) s.maybe_active = true;  // finite sigmas can only come from the exact path
  *nactive = s.maybe_active ? 1 : 0;
  return RUVNB_OK;
}

// RM_W = W - groupmean(W) (R/ruvIIInb.R:331-340)
static int make_rmw(ruvnb_ctx* ctx, const double* W) {
  const int K = ctx->K, m = ctx->m;
  k_wbar<<<nblk((long long)K * m * 32, 256), 256, 0, ctx->stream>>>(W, ctx->Np, K, m, ctx->d_group_chunk0, ctx->d_group_n, ctx->Wbar);
  CKL();
  k_rmw<<<nblk(ctx->Np * K, 256), 256, 0, ctx->stream>>>(W, ctx->Np, K, ctx->d_chunk_grp, ctx->d_chunk_valid, ctx->Wbar, ctx->RMW);
  CKL();
  return RUVNB_OK;
}

// H4: pi0 of parameter set `s` by R's one-dimensional Nelder-Mead (zinb.cuh); pi0_new may alias s.pi0.  *nbad = non-finite entries.
static int zinb_update_pi0(ruvnb_ctx* ctx, StateSet& s, const double* pi0_old, double* pi0_new, int* nbad) {
  const int Gl = ctx->Gl;
  if (!ctx->zq) {
    RET(dev_alloc(ctx, &ctx->zq, (long long)Gl * ctx->Np)); RET(dev_alloc(ctx, &ctx->z_spos, Gl)); RET(dev_alloc(ctx, &ctx->z_npos, Gl));
    RET(dev_alloc(ctx, &ctx->z_xeval, Gl)); RET(dev_alloc(ctx, &ctx->z_feval, Gl)); RET(dev_alloc(ctx, &ctx->z_nm, Gl)); RET(dev_alloc(ctx, &ctx->z_active, 2));
  }
  RET(ensure_ytabs(ctx, s.psi));
  DISPATCH_K(ctx->K, {
    LAUNCH_ROWS(k_zinb_q<KT>, 1, Gl, geom(ctx, Gl, nullptr), gstate<KT>(ctx, s, false), ctx->zq, ctx->z_spos, ctx->z_npos);
  });
  k_nm_init<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, ctx->z_nm, ctx->z_xeval); CKL();
  int rounds = 0;
  for (;; ++rounds) {
    if (rounds > 2 * NM_MAXIT + 8) return ctx->fail(RUVNB_ESTATE, "zinb: Nelder-Mead state machines did not stop");
    k_nm_eval<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(Gl, ctx->Np, ctx->zq, ctx->z_spos, ctx->z_npos, ctx->z_nm, ctx->z_xeval, ctx->d_mtabs, ctx->z_feval); CKL();
    CK(cudaMemsetAsync(ctx->z_active, 0, sizeof(int), ctx->stream));
    k_nm_step<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, ctx->z_nm, ctx->z_feval, ctx->z_xeval, ctx->z_active); CKL();
    if ((rounds & 3) == 3) {  // look at the number of running genes every 4 rounds
      CK(cudaMemcpyAsync(ctx->h_flags + 7, ctx->z_active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      if (ctx->h_flags[7] == 0) break;
    }
  }
  ctx->last_nm_rounds = rounds + 1;
  CK(cudaMemsetAsync(ctx->z_active + 1, 0, sizeof(int), ctx->stream));
  k_nm_finish<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, ctx->z_nm, pi0_old, pi0_new, ctx->z_active + 1); CKL();
  RET(allreduce(ctx, ctx->z_active + 1, 1, NCCL_INT32));
  CK(cudaMemcpyAsync(ctx->h_flags + 7, ctx->z_active + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *nbad = ctx->h_flags[7];
  return RUVNB_OK;
}

// one pass of the inner-loop body, R/ruvIIInb.R:303-484.  cur -> nw, then the two sets swap.
static int irls_body(ruvnb_ctx* ctx, const ruvnb_opts* o, double* logl_new, int bad[3]) {
  const int Gl = ctx->Gl, K = ctx->K, m = ctx->m, KK = K * (K + 1) / 2;
  const double kh = o->robust ? 1.345 : 100.0;  // :155
  StateSet& c = ctx->cur; StateSet& n = ctx->nw;
  CK(cudaMemsetAsync(ctx->flags, 0, 4 * sizeof(int), ctx->stream));
  RET(ensure_ytabs(ctx, c.psi));
  // Huber scale of the current state (:713,723,409): normally left by the log-likelihood sweep that scored it
  if (!c.sig_valid) RET(loglik_of(ctx, o, c, ctx->loglik, nullptr));
  int nfail = 0, nactive = 0;
  RET(complete_sigma(ctx, o, c, kh, &nfail, &nactive));
  ctx->last_fail_rows = nfail; ctx->last_active_rows = nactive;
  const bool hub = nactive > 0 || ctx->zinb_on;   // the general instantiation also carries the ZINB weights
  k_partition_sigma<<<1, 1024, 0, ctx->stream>>>(Gl, c.sigma, ctx->roworder_ready ? ctx->d_roworder : nullptr, ctx->list_inf, ctx->list_fin, ctx->part_counts);
  CKL();
  // alpha ------------------------------------------------------------------------------------------
  RET(make_rmw(ctx, c.W));
  DISPATCH_K(K, {
    auto st = gstate<KT>(ctx, c, true);
    KEV_BEGIN(1);
    const int P = ctx->nparts;
    ctx->launch_parts = P;
    if (P > 1) {  // groups that a part never meets must read as zero
      CK(cudaMemsetAsync(ctx->ZS, 0, (size_t)P * Gl * m * 8, ctx->stream));
      CK(cudaMemsetAsync(ctx->VS, 0, (size_t)P * Gl * m * KT * 8, ctx->stream));
    }
    // certified rows (sigma = +Inf: every Huber weight is 1) take the lean instantiation, the others the general one;
    // the two lists and their lengths stay on the device (no host round trip), idle CTAs exit at once
    if (ctx->zinb_on) {
      LAUNCH_ROWS_NW((k_pass1<KT, true, P1_NW>), P1_NW, 2, Gl, geom(ctx, Gl, nullptr), st, ctx->RMW, c.sigma, kh, ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS);
    } else {
      // the (few) rows with live Huber weights go FIRST, on the side stream: a row is one warp walking 100k cells, so a
      // handful of them after the bulk would add a whole row-walk (2.5 ms at cfg 2) to every iteration
      if (hub) {
        RowGeom gf = geom(ctx, Gl, ctx->list_fin); gf.nrows_dev = ctx->part_counts + 1;
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream)); CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        cudaStream_t main_ = ctx->stream; ctx->stream = ctx->stream2;
        LAUNCH_ROWS_NW((k_pass1<KT, true, P1_NW>), P1_NW, 2, Gl, gf, st, ctx->RMW, c.sigma, kh, ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS);
        ctx->stream = main_;
        CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
      }
      RowGeom gi = geom(ctx, Gl, ctx->list_inf); gi.nrows_dev = ctx->part_counts;
      LAUNCH_ROWS_NW((k_pass1<KT, false, P1_NW>), P1_NW, 2, Gl, gi, st, ctx->RMW, c.sigma, kh, ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS);
      if (hub) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    ctx->launch_parts = 1;
    if (P > 1) {
      constexpr int KKT = KT * (KT + 1) / 2;
      k_sum_parts<<<nblk((long long)Gl * KKT, 256), 256, 0, ctx->stream>>>(ctx->A, ctx->A, (long long)Gl * KKT, P); CKL();
      k_sum_parts<<<nblk((long long)Gl * KT, 256), 256, 0, ctx->stream>>>(ctx->u, ctx->u, (long long)Gl * KT, P); CKL();
      k_sum_parts<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->sw, ctx->sw, Gl, P); CKL();
      k_sum_parts<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(ctx->ZS, ctx->ZS, (long long)Gl * m, P); CKL();
      k_sum_parts<<<nblk((long long)Gl * m * KT, 256), 256, 0, ctx->stream>>>(ctx->VS, ctx->VS, (long long)Gl * m * KT, P); CKL();
    }
    KEV_END(1);
    k_alpha_finish<KT><<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, m, (double)ctx->N, ctx->d_group_n, ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS, c.adev, ctx->cvec, ctx->Ab);
    CKL();
    k_colsum<<<1, 1024, 0, ctx->stream>>>(ctx->Ab, Gl, KK + K, ctx->red);
    CKL();
    RET(allreduce(ctx, ctx->red, KK + K, NCCL_F64));
    k_alpha_mean<KT><<<1, 32, 0, ctx->stream>>>(ctx->red, ctx->amean);
    CKL();
    k_adev<KT><<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, ctx->A, ctx->cvec, ctx->amean, o->lambda_a, n.alpha, n.adev, ctx->flags + 0);
    CKL();
  });
  RET(allreduce(ctx, ctx->flags, 1, NCCL_INT32));
  k_revert_if<<<nblk((long long)Gl * K, 256), 256, 0, ctx->stream>>>(ctx->flags + 0, n.alpha, c.alpha, (long long)Gl * K);  // :368
  CKL();
  // clamp each column to median +- 4 mad over ALL genes (:371-376)
  if (ctx->nranks > 1) {
    CK(cudaMemsetAsync(ctx->alpha_full, 0, (size_t)ctx->G * K * 8, ctx->stream));
    CK(cudaMemcpyAsync(ctx->alpha_full + ctx->g0 * K, n.alpha, (size_t)Gl * K * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    RET(allreduce(ctx, ctx->alpha_full, (size_t)ctx->G * K, NCCL_F64));
    k_strided_medmad<<<K, RS_NT, 0, ctx->stream>>>(ctx->alpha_full, 1, K, (int)ctx->G, ctx->med, ctx->mad);
  } else {
    k_strided_medmad<<<K, RS_NT, 0, ctx->stream>>>(n.alpha, 1, K, Gl, ctx->med, ctx->mad);
  }
  CKL();
  k_clamp_cols<<<nblk((long long)Gl * K, 256), 256, 0, ctx->stream>>>(n.alpha, Gl, K, ctx->med, ctx->mad);
  CKL();
  // W ------------------------------------------------------------------------------------------------
  {
    const int nc = ctx->n_ctl;
    double* pk = ctx->ctlpack;
    CtlPack cp;
    double* p_gmean = pk; double* p_psi = pk + nc; double* p_step = pk + 2 * nc;
    double* p_aold = pk + 3 * nc; double* p_anew = p_aold + (long long)nc * K; double* p_mb = p_anew + (long long)nc * K;
    double* p_pi0 = p_mb + (long long)nc * m; double* p_psiw = p_pi0 + nc;
    k_gather_rows<<<nblk(nc, 256), 256, 0, ctx->stream>>>(c.gmean, ctx->d_ctl_local, nc, 1, p_gmean); CKL();
    k_gather_rows<<<nblk(nc, 256), 256, 0, ctx->stream>>>(c.psi, ctx->d_ctl_local, nc, 1, p_psi); CKL();
    k_gather_rows<<<nblk(nc, 256), 256, 0, ctx->stream>>>(ctx->step, ctx->d_ctl_local, nc, 1, p_step); CKL();
    k_gather_rows<<<nblk((long long)nc * K, 256), 256, 0, ctx->stream>>>(c.alpha, ctx->d_ctl_local, nc, K, p_aold); CKL();
    k_gather_rows<<<nblk((long long)nc * K, 256), 256, 0, ctx->stream>>>(n.alpha, ctx->d_ctl_local, nc, K, p_anew); CKL();
    k_gather_rows<<<nblk((long long)nc * m, 256), 256, 0, ctx->stream>>>(c.Mb, ctx->d_ctl_local, nc, m, p_mb); CKL();
    k_gather_rows<<<nblk(nc, 256), 256, 0, ctx->stream>>>(c.pi0, ctx->d_ctl_local, nc, 1, p_pi0); CKL();
    k_gather_rows<<<nblk(nc, 256), 256, 0, ctx->stream>>>(c.psi_w, ctx->d_ctl_local, nc, 1, p_psiw); CKL();
    RET(allreduce(ctx, pk, (size_t)nc * (5 + 2 * K + m), NCCL_F64));
    cp.gmean = p_gmean; cp.psi = p_psi; cp.step = p_step; cp.alpha_old = p_aold; cp.alpha_new = p_anew; cp.Mb = p_mb;
    cp.pi0 = p_pi0; cp.psi_w = p_psiw; cp.zinf = ctx->d_zinf; cp.zinb = ctx->zinb_on ? 1 : 0;
    // cells are sharded over the ranks in whole chunks; the slices are summed (disjoint) afterwards
    long long ch_per = (ctx->nchunks + ctx->nranks - 1) / ctx->nranks;
    long long pb = std::min<long long>(ctx->Np, ch_per * ctx->rank * CH), pe = std::min<long long>(ctx->Np, ch_per * (ctx->rank + 1) * CH);
    if (ctx->nranks > 1) CK(cudaMemsetAsync(n.W, 0, (size_t)ctx->Np * K * 8, ctx->stream));
    // exact kernel geometry: the per-cell deviance rows of all control genes must fit shared memory
    const int NACC = KK + K;
    int cpc = 32;
    const long long nstr = nc | 1;
    while (cpc > 1 && (size_t)cpc * nstr * 8 > (size_t)200 * 1024) cpc >>= 1;
    if ((size_t)cpc * nstr * 8 > (size_t)200 * 1024) return ctx->fail(RUVNB_ENOMEM, "n_ctl = %d control genes exceed the per-cell shared-memory tile (max 25600)", nc);
    size_t smem = std::max<size_t>((size_t)cpc * (size_t)nstr * 8, (size_t)8 * cpc * NACC * 8);
    if (pe > pb) {
      KEV_BEGIN(2);
      const bool fast = o->huber_fastpath && !o->robust && !ctx->zinb_on;
      if (fast) {
        k_build_ytabs<<<nc, YT, 0, ctx->stream>>>(p_psi, nc, nullptr, ctx->ctl_tabR, nullptr); CKL();
        CK(cudaMemsetAsync(ctx->w_fail, 0, sizeof(int), ctx->stream));
      }
      DISPATCH_K(K, {
        CK(cudaFuncSetAttribute(k_w_update<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (fast) {
          k_w_update_fast<KT><<<nblk(pe - pb, 64), 256, 0, ctx->stream>>>(ctx->d_ctl_rows, nc, ctx->Np, ctx->d_chunk_grp, ctx->d_chunk_valid, cp, ctx->ctl_tabR,
                                                                        ctx->d_mtabs, m, c.W, n.W, kh, pb, pe, ctx->w_fail, ctx->w_fail + 1);
          CKL();
          // cells whose certificate failed: exact per-cell MAD (CTAs beyond the failure count exit at once)
          k_w_update<KT, true><<<nblk(pe - pb, cpc), 256, smem, ctx->stream>>>(ctx->d_ctl_rows, nc, ctx->Np, ctx->d_chunk_grp, ctx->d_chunk_valid, cp, m, c.W, n.W,
                                                                              cpc, kh, pb, pe, ctx->w_fail + 1, ctx->w_fail);
          CKL();
        } else {
          k_w_update<KT, true><<<nblk(pe - pb, cpc), 256, smem, ctx->stream>>>(ctx->d_ctl_rows, nc, ctx->Np, ctx->d_chunk_grp, ctx->d_chunk_valid, cp, m, c.W, n.W,
                                                                              cpc, kh, pb, pe, nullptr, nullptr);
          CKL();
        }
      });
      KEV_END(2);
    }
    RET(allreduce(ctx, n.W, (size_t)ctx->Np * K, NCCL_F64));
    k_w_sumsq<<<K, 1024, 0, ctx->stream>>>(n.W, ctx->Np, ctx->sumsq); CKL();
    k_w_scale<<<nblk(ctx->Np * K, 256), 256, 0, ctx->stream>>>(n.W, ctx->Np, K, ctx->sumsq, ctx->d_chunk_valid, ctx->flags + 1); CKL();
    k_revert_if<<<nblk(ctx->Np * K, 256), 256, 0, ctx->stream>>>(ctx->flags + 1, n.W, c.W, ctx->Np * K); CKL();  // :402
  }
  // gmean, Mb ------------------------------------------------------------------------------------------
  DISPATCH_K(K, {
    auto st = gstate<KT>(ctx, c, true);
    KEV_BEGIN(3);
    const int P = ctx->nparts;
    ctx->launch_parts = P;
    if (P > 1) {
      CK(cudaMemsetAsync(ctx->S0, 0, (size_t)P * Gl * m * 8, ctx->stream));
      CK(cudaMemsetAsync(ctx->S1, 0, (size_t)P * Gl * m * 8, ctx->stream));
    }
    if (ctx->zinb_on) {
      LAUNCH_ROWS((k_pass2<KT, true>), 2, Gl, geom(ctx, Gl, nullptr), st, n.W, n.alpha, c.sigma, kh, ctx->S0, ctx->S1);
    } else {
      if (hub) {
        RowGeom gf = geom(ctx, Gl, ctx->list_fin); gf.nrows_dev = ctx->part_counts + 1;
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream)); CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        cudaStream_t main_ = ctx->stream; ctx->stream = ctx->stream2;
        LAUNCH_ROWS((k_pass2<KT, true>), 2, Gl, gf, st, n.W, n.alpha, c.sigma, kh, ctx->S0, ctx->S1);
        ctx->stream = main_;
        CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
      }
      RowGeom gi = geom(ctx, Gl, ctx->list_inf); gi.nrows_dev = ctx->part_counts;
      LAUNCH_ROWS((k_pass2<KT, false>), 2, Gl, gi, st, n.W, n.alpha, c.sigma, kh, ctx->S0, ctx->S1);
      if (hub) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    ctx->launch_parts = 1;
    if (P > 1) {
      k_sum_parts<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(ctx->S0, ctx->S0, (long long)Gl * m, P); CKL();
      k_sum_parts<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(ctx->S1, ctx->S1, (long long)Gl * m, P); CKL();
    }
    KEV_END(3);
  });
  k_gmean_mb<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, m, ctx->S0, ctx->S1, c.Mb, o->lambda_b, n.gmean, n.Mb, ctx->flags + 3, ctx->flags + 2); CKL();
  RET(allreduce(ctx, ctx->flags + 2, 2, NCCL_INT32));
  k_mb_finalize<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(Gl, m, ctx->flags + 3, c.Mb, n.Mb); CKL();
  // psi is fixed inside the inner loop
  CK(cudaMemcpyAsync(n.psi, c.psi, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  if (ctx->zinb_on) {
    // step 4a (:435-467): pi0 at the NEW parameters, then the E-step weights -- recomputed on the fly from (pi0, psi_w)
    CK(cudaMemcpyAsync(n.pi0, c.pi0, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(n.psi_w, c.psi_w, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    int nbad_pi0 = 0;
    RET(zinb_update_pi0(ctx, n, c.pi0, n.pi0, &nbad_pi0));
    if (nbad_pi0) CK(cudaMemcpyAsync(n.pi0, c.pi0, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));  // :454
    CK(cudaMemcpyAsync(n.psi_w, n.psi, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));             // :457-467
  } else {
    CK(cudaMemcpyAsync(n.pi0, c.pi0, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(n.psi_w, c.psi_w, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  // candidate log-likelihood (:469-484) + its Huber certificate for the next iteration
  RET(loglik_of(ctx, o, n, ctx->loglik_tmp, logl_new));
  CK(cudaMemcpyAsync(ctx->h_flags, ctx->flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_flags + 6, ctx->w_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (bad) { bad[0] = ctx->h_flags[0]; bad[1] = ctx->h_flags[1]; bad[2] = ctx->h_flags[2]; }
  ctx->last_fail_cells = ctx->h_flags[6];
  std::swap(ctx->cur, ctx->nw);
  return RUVNB_OK;
}

static int copy_set(ruvnb_ctx* ctx, StateSet& dst, const StateSet& src) {
  const size_t Gl = ctx->Gl, K = ctx->K, m = ctx->m;
  CK(cudaMemcpyAsync(dst.gmean, src.gmean, Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.Mb, src.Mb, Gl * m * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.alpha, src.alpha, Gl * K * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.adev, src.adev, Gl * K * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.W, src.W, (size_t)ctx->Np * K * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.psi, src.psi, Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.pi0, src.pi0, Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.sigma, src.sigma, Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.psi_w, src.psi_w, Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.wtctl, src.wtctl, (size_t)ctx->n_ctl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  dst.sig_valid = src.sig_valid; dst.maybe_active = src.maybe_active;
  return RUVNB_OK;
}

extern "C" {

int ruvnb_init_coef(ruvnb_ctx* ctx, const ruvnb_opts* o) {
  RET(check_ready(ctx, o));
  const int Gl = ctx->Gl, K = ctx->K;
  StateSet& c = ctx->cur;
  DISPATCH_K(K, {
    LAUNCH_ROWS(k_init_coef<KT>, 1, Gl, geom(ctx, Gl, nullptr), c.W, c.gmean, c.alpha);
  });
  // alpha.dev = alpha - colMeans(alpha) over ALL genes (:187-188), then NaN/Inf alpha -> 0 (:195)
  k_colsum<<<1, 1024, 0, ctx->stream>>>(c.alpha, Gl, K, ctx->red); CKL();
  RET(allreduce(ctx, ctx->red, K, NCCL_F64));
  k_init_adev<<<nblk((long long)Gl * K, 256), 256, 0, ctx->stream>>>(Gl, K, ctx->red, (double)ctx->G, c.alpha, c.adev); CKL();
  CK(cudaMemsetAsync(c.Mb, 0, (size_t)Gl * ctx->m * 8, ctx->stream));  // :191
  CK(cudaMemsetAsync(c.pi0, 0, (size_t)Gl * 8, ctx->stream));         // :202
  CK(cudaStreamSynchronize(ctx->stream));
  c.sig_valid = false;
  return RUVNB_OK;
}

int ruvnb_loglik(ruvnb_ctx* ctx, const ruvnb_opts* o, double* total) {
  RET(check_ready(ctx, o));
  return loglik_of(ctx, o, ctx->cur, ctx->loglik, total);
}

int ruvnb_irls_iteration(ruvnb_ctx* ctx, const ruvnb_opts* o, double* logl_new, int bad[3]) {
  RET(check_ready(ctx, o));
  double t = 0.0;
  RET(irls_body(ctx, o, &t, bad));
  if (logl_new) *logl_new = t;
  return RUVNB_OK;
}

int ruvnb_zinb_pi0(ruvnb_ctx* ctx, const ruvnb_opts* o, int* bad) {
  RET(check_ready(ctx, o));
  if (!ctx->zinb_on) return ctx->fail(RUVNB_ESTATE, "zinb_pi0: needs opts.zinb and zeroinf flags in ruvnb_set_design");
  int nb = 0;
  RET(zinb_update_pi0(ctx, ctx->cur, ctx->cur.pi0, ctx->cur.pi0, &nb));
  CK(cudaMemcpyAsync(ctx->cur.psi_w, ctx->cur.psi, (size_t)ctx->Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->cur.sig_valid = false;
  if (bad) *bad = nb;
  return RUVNB_OK;
}

int ruvnb_snapshot_best(ruvnb_ctx* ctx) {
  if (!ctx || ctx->K == 0) return RUVNB_EARG;
  CK(cudaSetDevice(ctx->device));
  RET(copy_set(ctx, ctx->best, ctx->cur));
  CK(cudaStreamSynchronize(ctx->stream));
  return RUVNB_OK;
}
int ruvnb_restore_best(ruvnb_ctx* ctx) {
  if (!ctx || ctx->K == 0) return RUVNB_EARG;
  CK(cudaSetDevice(ctx->device));
  RET(copy_set(ctx, ctx->cur, ctx->best));
  // "revert zinb weight" (:494-504): the weights are recomputed from the restored parameters with the CURRENT psi
  if (ctx->zinb_on) CK(cudaMemcpyAsync(ctx->cur.psi_w, ctx->cur.psi, (size_t)ctx->Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RUVNB_OK;
}
int ruvnb_halve_steps(ruvnb_ctx* ctx, const ruvnb_opts* o) {
  if (!ctx || !o || ctx->K == 0) return RUVNB_EARG;
  CK(cudaSetDevice(ctx->device));
  k_halve<<<nblk(ctx->Gl, 256), 256, 0, ctx->stream>>>(ctx->Gl, ctx->loglik, ctx->loglik_tmp, o->step_fac, ctx->step); CKL();
  return RUVNB_OK;
}
// accept the candidate: loglik <- loglik.tmp (:512)
int ruvnb_accept(ruvnb_ctx* ctx) {
  if (!ctx || ctx->K == 0) return RUVNB_EARG;
  std::swap(ctx->loglik, ctx->loglik_tmp);
  return RUVNB_OK;
}

// ---- the whole fit: R/ruvIIInb.R:266-609 ------------------------------------------------------------
int ruvnb_fit_nb(ruvnb_ctx* ctx, const ruvnb_opts* o, const double* W0, const double* psi_inject, int n_psi,
                 ruvnb_trace_rec* trace, int trace_cap, double* logl_outer, ruvnb_fit_stats* stats) {
  if (!ctx || !o) return RUVNB_EARG;
  if (W0) RET(ruvnb_set_state(ctx, "W", W0, o->k));
  RET(check_ready(ctx, o));
  cudaEvent_t e0, e1, t0, t1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
  const long long launches0 = ctx->launches;
  CK(cudaEventRecord(t0, ctx->stream));
  const int Gl = ctx->Gl;
  double inner_ms = 0.0, disp_ms = 0.0;
  int n_trace = 0, n_inner = 0, psi_used = 0, exact_rows = 0;
  auto refresh_psi = [&](StateSet& target) -> int {
    // the dispersion routine is an injection seam (edgeR/limma/locfit arithmetic)
    if (psi_inject && n_psi > 0) {
      const double* src = psi_inject + (long long)std::min(psi_used, n_psi - 1) * Gl;
      psi_used++;
      std::vector<double> tmp(Gl), old(Gl);
      CK(cudaMemcpyAsync(old.data(), target.psi, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (int g = 0; g < Gl; ++g) tmp[g] = std::isfinite(src[g]) ? src[g] : old[g];  // :564
      CK(cudaMemcpyAsync(target.psi, tmp.data(), (size_t)Gl * 8, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->psi_ver++; target.sig_valid = false;
      return RUVNB_OK;
    }
    cudaEvent_t d0, d1;
    CK(cudaEventCreate(&d0)); CK(cudaEventCreate(&d1));
    CK(cudaEventRecord(d0, ctx->stream));
    // estimate on the parameters held in `target` (offset = alpha W' + Mb[,grp], :535-536)
    if (&target != &ctx->cur) std::swap(ctx->cur, target);
    int r = ruvnb_estimate_disp(ctx, o, nullptr);
    if (&target != &ctx->cur) std::swap(ctx->cur, target);
    RET(r);
    ctx->psi_ver++; target.sig_valid = false;
    CK(cudaEventRecord(d1, ctx->stream));
    CK(cudaEventSynchronize(d1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, d0, d1)); disp_ms += ms;
    CK(cudaEventDestroy(d0)); CK(cudaEventDestroy(d1));
    return RUVNB_OK;
  };
  // initial estimates (:173-221)
  RET(ruvnb_init_coef(ctx, o));
  RET(refresh_psi(ctx->cur));
  if (ctx->zinb_on) {  // :237-264: initial pi0 (from pi0 = 0) and E-step
    int nbad_pi0 = 0;
    RET(zinb_update_pi0(ctx, ctx->cur, ctx->cur.pi0, ctx->cur.pi0, &nbad_pi0));
    CK(cudaMemcpyAsync(ctx->cur.psi_w, ctx->cur.psi, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->cur.sig_valid = false;
  }
  k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->step, Gl, 1.0); CKL();  // :270
  bool conv = false;
  int iter_outer = 0;
  std::vector<double> lo;
  while (!conv) {
    iter_outer++;
    std::vector<double> logl_beta;
    double s_old = 0.0;
    RET(loglik_of(ctx, o, ctx->cur, ctx->loglik, &s_old));  // :276-291
    k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->step, Gl, 1.0); CKL();  // :293
    RET(copy_set(ctx, ctx->best, ctx->cur));                                    // :296
    int iter = 1, halving = 0;
    bool conv_beta = false;
    CK(cudaEventRecord(e0, ctx->stream));
    while (!conv_beta) {
      if (o->verbose) fprintf(stderr, "Inner iter:%d\n", iter);
      double s_new = 0.0; int bad[3] = {0, 0, 0};
      RET(irls_body(ctx, o, &s_new, bad));
      n_inner++;
      exact_rows += ctx->last_fail_rows;
      bool degener = !(std::isfinite(s_old) && std::isfinite(s_new)) || (s_old > s_new);  // :487-488
      ruvnb_trace_rec rec;
      rec.outer = iter_outer; rec.iter = iter; rec.halving = halving; rec.degener = degener ? 1 : 0;
      rec.logl_old = s_old; rec.logl_new = s_new; rec.bad_alpha = bad[0]; rec.bad_W = bad[1]; rec.bad_Mb = bad[2];
      if (degener) {
        k_halve<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(Gl, ctx->loglik, ctx->loglik_tmp, o->step_fac, ctx->step); CKL();  // :490-491
        RET(copy_set(ctx, ctx->cur, ctx->best));  // :492
        if (ctx->zinb_on) CK(cudaMemcpyAsync(ctx->cur.psi_w, ctx->cur.psi, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));  // :494-504
        halving++;
        if (halving >= 3) { degener = false; }    // :506-509: loglik.tmp <- loglik, accept the restored state
      } else {
        std::swap(ctx->loglik, ctx->loglik_tmp);  // :512
        s_old = s_new;
      }
      if (!degener) {
        logl_beta.push_back(s_old);               // :513
        RET(copy_set(ctx, ctx->best, ctx->cur));  // :516
        iter++; halving = 0;
        if (o->verbose) fprintf(stderr, "Outer Iter %d, Inner iter %d logl-likelihood:%.10g\n", iter_outer, iter - 1, s_old);
      }
      rec.accepted = degener ? 0 : 1;
      rec.logl = logl_beta.empty() ? NAN : logl_beta.back();
      if (trace && n_trace < trace_cap) trace[n_trace++] = rec;
      bool conv_logl = false;
      if (iter > 2) {  // :523-525
        size_t L = logl_beta.size();
        conv_logl = (logl_beta[L - 1] - logl_beta[L - 2]) / fabs(logl_beta[L - 2]) < o->inner_tol;
      }
      conv_beta = iter >= o->inner_maxit || conv_logl;
    }
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    { float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); inner_ms += ms; }
    // psi update (:529-565)
    double logl_outer_tmp = s_old;
    bool updt = true;
    if (iter_outer > 1) updt = fabs(lo.back() - logl_outer_tmp) / fabs(lo.back()) > o->psi_tol;
    RET(copy_set(ctx, ctx->cur, ctx->best));  // :568-573
    if (updt) RET(refresh_psi(ctx->cur));
    double s = 0.0;
    RET(loglik_of(ctx, o, ctx->cur, ctx->loglik, &s));  // :577-593
    lo.push_back(s);
    if (logl_outer && iter_outer <= o->outer_maxit) logl_outer[iter_outer - 1] = s;
    if (o->verbose) fprintf(stderr, "Outer Iter %d logl-likelihood:%.10g\n", iter_outer, s);
    if (iter_outer > 1) {
      size_t L = lo.size();
      conv = ((lo[L - 1] - lo[L - 2]) / fabs(lo[L - 2]) < o->outer_tol) || iter_outer >= o->outer_maxit;  // :596-598
    }
  }
  CK(cudaEventRecord(t1, ctx->stream));
  CK(cudaEventSynchronize(t1));
  if (stats) {
    float ms = 0; CK(cudaEventElapsedTime(&ms, t0, t1));
    stats->n_outer = iter_outer; stats->n_inner = n_inner; stats->n_trace = n_trace;
    stats->inner_seconds = inner_ms * 1e-3; stats->disp_seconds = disp_ms * 1e-3; stats->total_seconds = ms * 1e-3;
    stats->kernel_launches = ctx->launches - launches0;
    stats->n_exact_mad_rows = exact_rows; stats->n_exact_mad_cells = (int)std::min<long long>(ctx->N * (long long)n_inner, 2147483647ll);
  }
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1)); CK(cudaEventDestroy(t0)); CK(cudaEventDestroy(t1));
  return RUVNB_OK;
}

}  // extern "C"

// ---- H3: dispersion (R/estimateDisp.par.R) ------------------------------------------------------------------------------
// dst[r][d0 + i ds] = src[r][src_col + i ds] (a converged neighbour or earlier column) where that is finite and within
// maxdiff of the one-group start value beta0[r], else beta0[r] (src == nullptr: always); i < nd.  From far above the root
// Fisher scoring walks down by 1 per sweep, so an intercept left over from very different parameters (a gene whose
// coefficients were reverted in between) must not be used.
__global__ void k_start_cols(const double* src, int src_ld, int src_col, const double* beta0, double maxdiff, long long n, int ld, int d0, int ds, int nd, double* dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nd) return;
  long long r = i / nd; int d = (int)(i % nd);
  double v = beta0[r];
  if (src) { double p = src[r * src_ld + src_col + d * ds]; if (fabs(p) <= 1.79769313486231570e308 && fabs(p - v) <= maxdiff) v = p; }
  dst[r * ld + d0 + d * ds] = v;
}
// start value for a one-dispersion fit from the converged grid: the intercept is a smooth function of log(phi), so linear
// interpolation between the two bracketing grid columns lands within ~1e-3 of the root (NR then needs 2-3 sweeps, not 8).
// prior_count > 0: the aveLogCPM fit of y + p against log(e^o + 2p) -- shift the start as edgeR's own start is shifted.
__global__ void k_beta_from_grid(const double* beta, const double* phi_grid, int ng, const double* phi_row, const double* emean,
                                 double prior_count, int n, double* out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double lp = log(phi_row[r]);
  int d = 0;
  while (d + 2 < ng && log(phi_grid[d + 1]) < lp) ++d;
  const double l0 = log(phi_grid[d]), l1 = log(phi_grid[d + 1]);
  double t = fmin(fmax((lp - l0) / (l1 - l0), 0.0), 1.0);
  const double b0 = beta[(long long)r * ng + d], b1 = beta[(long long)r * ng + d + 1];
  double b = (b0 == b1) ? b0 : fma(t, b1 - b0, b0);   // -Inf rows stay -Inf (no Inf - Inf)
  if (!(b == b)) b = b0;
  if (prior_count > 0.0) { double pcm = prior_count / emean[r]; b = log((exp(b) + pcm) / fma(2.0, pcm, 1.0)); }
  out[r] = b;
}
__global__ void k_bcast_cols(const double* src, long long n, int ld, int d0, int nd, double* dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nd) return;
  long long r = i / nd; int d = (int)(i % nd);
  dst[r * ld + d0 + d] = src[r];
}
static int disp_ensure(ruvnb_ctx* ctx) {
  if (ctx->dp_ready) return RUVNB_OK;
  const long long Gl = ctx->Gl;
  RET(dev_alloc(ctx, &ctx->dp_beta, Gl * DISP_NG)); RET(dev_alloc(ctx, &ctx->dp_beta0, Gl)); RET(dev_alloc(ctx, &ctx->dp_rowsum, Gl));
  RET(dev_alloc(ctx, &ctx->dp_emean, Gl)); RET(dev_alloc(ctx, &ctx->dp_l0, Gl * DISP_NG)); RET(dev_alloc(ctx, &ctx->dp_phi_grid, DISP_NG));
  RET(dev_alloc(ctx, &ctx->dp_tabG, DISP_NG * YT)); RET(dev_alloc(ctx, &ctx->dp_tabRg, DISP_NG * YT)); RET(dev_alloc(ctx, &ctx->dp_lgr_grid, DISP_NG));
  RET(dev_alloc(ctx, &ctx->dp_zeros, Gl)); RET(dev_alloc(ctx, &ctx->dp_beta1, Gl)); RET(dev_alloc(ctx, &ctx->dp_phi_row, Gl));
  RET(dev_alloc(ctx, &ctx->dp_dev, Gl)); RET(dev_alloc(ctx, &ctx->dp_tabR2, Gl * YT)); RET(dev_alloc(ctx, &ctx->dp_flag, 2));
  RET(dev_alloc(ctx, &ctx->dp_done, Gl)); RET(dev_alloc(ctx, &ctx->dp_live, Gl)); RET(dev_alloc(ctx, &ctx->dp_info, Gl * DISP_NG)); RET(dev_alloc(ctx, &ctx->dp_step, Gl * DISP_NG));
  RET(dev_alloc(ctx, &ctx->dp_hist, Gl * YT)); RET(dev_alloc(ctx, &ctx->dp_syo, Gl));
  RET(dev_alloc(ctx, &ctx->dp_xs, ctx->G)); RET(dev_alloc(ctx, &ctx->dp_m0, ctx->G * DISP_NG)); RET(dev_alloc(ctx, &ctx->dp_perm, ctx->G));
  CK(cudaMemsetAsync(ctx->dp_zeros, 0, (size_t)Gl * 8, ctx->stream));
  double grid[DISP_NG];
  for (int i = 0; i < DISP_NG; ++i) grid[i] = 0.1 * exp2(-10.0 + (double)i);  // spline.disp, R/estimateDisp.par.R:33-35
  CK(cudaMemcpyAsync(ctx->dp_phi_grid, grid, sizeof grid, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  k_build_ytabs<<<DISP_NG, YT, 0, ctx->stream>>>(ctx->dp_phi_grid, DISP_NG, ctx->dp_tabG, ctx->dp_tabRg, ctx->dp_lgr_grid);
  CKL();
  ctx->dp_ready = true;
  return RUVNB_OK;
}
// offset = alpha W' + Mb[, grp] (no gmean): the parameter set `s` with a zero gmean
template <int K>
static GeneState<K> disp_state(ruvnb_ctx* ctx, const StateSet& s) {
  GeneState<K> st = gstate<K>(ctx, s, false);
  st.gmean = ctx->dp_zeros;
  return st;
}
// Newton-Raphson to convergence (edgeR glm_one_group: |step| < 1e-10, at most 50 iterations) for columns d0..d0+nd-1
static int disp_nr(ruvnb_ctx* ctx, const StateSet& s, DispArgs a, int nd) {
  a.done = ctx->dp_done; a.step_prev = ctx->dp_step;
  CK(cudaMemsetAsync(ctx->dp_done, 0, (size_t)ctx->Gl * sizeof(int), ctx->stream));
  CK(cudaMemsetAsync(ctx->dp_step, 0, (size_t)ctx->Gl * DISP_NG * 8, ctx->stream));
  const bool trace = getenv("RUVNB_TRACE_DISP") != nullptr;
  // the bulk of the rows converges within 4-5 sweeps; from the third sweep on only the rows still iterating are launched
  int nlive = ctx->Gl; const int* live = nullptr;
  for (int it = 0; it < 50; ++it) {
    const auto t0 = std::chrono::steady_clock::now();
    CK(cudaMemsetAsync(ctx->dp_flag, 0, 2 * sizeof(int), ctx->stream));
    DISPATCH_K(ctx->K, {
      if (a.Ecache) {
        if (nd == DISP_ND) k_disp_nr_cached<KT, DISP_ND><<<row_grid(nlive), ROW_THREADS, 0, ctx->stream>>>(geom(ctx, nlive, live), disp_state<KT>(ctx, s), a);
        else k_disp_nr_cached<KT, 1><<<row_grid(nlive), ROW_THREADS, 0, ctx->stream>>>(geom(ctx, nlive, live), disp_state<KT>(ctx, s), a);
        CKL();
      } else if (nd == DISP_ND) LAUNCH_ROWS((k_disp_nr<KT, DISP_ND>), 1, nlive, geom(ctx, nlive, live), disp_state<KT>(ctx, s), a);
      else LAUNCH_ROWS((k_disp_nr<KT, 1>), 1, nlive, geom(ctx, nlive, live), disp_state<KT>(ctx, s), a);
    });
    ctx->disp_nr_sweeps++;
    if (it >= 2) { k_list_zero<<<1, 1024, 0, ctx->stream>>>(ctx->Gl, ctx->dp_done, ctx->dp_live, ctx->dp_flag + 1); CKL(); }
    CK(cudaMemcpyAsync(ctx->h_flags + 6, ctx->dp_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (trace) {
      std::vector<int> dn(ctx->Gl); long long nd_rows = 0;
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      CK(cudaMemcpy(dn.data(), ctx->dp_done, (size_t)ctx->Gl * sizeof(int), cudaMemcpyDeviceToHost));
      for (int v : dn) nd_rows += v;
      fprintf(stderr, "  disp NR (nd=%d) sweep %d: %.2f ms over %d rows, %lld of %d rows converged after it\n", nd, it, ms, nlive, nd_rows, ctx->Gl);
    }
    if (ctx->h_flags[6] == 0) break;
    if (it >= 2) { nlive = ctx->h_flags[7]; live = ctx->dp_live; if (nlive <= 0) break; }
    if (trace && it == 49) {
      std::vector<int> lv(nlive); std::vector<double> bb((size_t)ctx->Gl * a.ldb), sp((size_t)ctx->Gl * DISP_NG);
      CK(cudaMemcpy(lv.data(), ctx->dp_live, (size_t)nlive * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(bb.data(), a.beta, bb.size() * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(sp.data(), ctx->dp_step, sp.size() * 8, cudaMemcpyDeviceToHost));
      for (int i = 0; i < std::min(nlive, 4); ++i) {
        fprintf(stderr, "  not converged: row %d:", lv[i]);
        for (int d = 0; d < nd; ++d) fprintf(stderr, " [beta %.17g step %.3g]", bb[(size_t)lv[i] * a.ldb + a.d0 + d * a.ds], sp[(size_t)lv[i] * a.ldb + a.d0 + d * a.ds]);
        fprintf(stderr, "\n");
      }
    }
  }
  return RUVNB_OK;
}
// count histogram of every row (the lgamma part of the grid log-likelihoods, disp.cuh OpDispAPL)
static int ensure_hist(ruvnb_ctx* ctx) {
  if (ctx->hist_ready) return RUVNB_OK;
  k_row_hist<<<ctx->Gl, 256, 0, ctx->stream>>>(ctx->dY, ctx->Np, ctx->dp_hist); CKL();
  ctx->hist_ready = true;
  return RUVNB_OK;
}
// l0 (device, [Gl][21] row-major), rowsum, emean, beta0 for the parameter set `s`
static int disp_grid(ruvnb_ctx* ctx, const StateSet& s) {
  RET(disp_ensure(ctx));
  RET(ensure_hist(ctx));
  const int Gl = ctx->Gl;
  ctx->disp_nr_sweeps = 0;
  // exp(offset) of every cell, kept for the ~40 sweeps of the step when HBM has the room (16 GB at 20k x 100k) next to a
  // reserve for everything allocated later (exact-MAD scratch, get.res slabs); otherwise every sweep recomputes it
  if (!ctx->dp_E_tried) {
    ctx->dp_E_tried = true;
    size_t fr = 0, tot = 0;
    CK(cudaMemGetInfo(&fr, &tot));
    const size_t need = (size_t)Gl * (size_t)ctx->Np * 8, reserve = std::max<size_t>((size_t)8 << 30, tot / 8);
    if (!getenv("RUVNB_NO_ECACHE") && fr > need + reserve) {
      void* q = nullptr;
      if (cudaMalloc(&q, need) == cudaSuccess) { ctx->allocs.push_back(q); ctx->dp_E = (double*)q; } else cudaGetLastError();
    }
  }
  CK(cudaMemsetAsync(ctx->dp_info, 0, (size_t)Gl * DISP_NG * 8, ctx->stream));   // all-zero genes are never iterated
  DISPATCH_K(ctx->K, {
    LAUNCH_ROWS(k_disp_init<KT>, 1, Gl, geom(ctx, Gl, nullptr), disp_state<KT>(ctx, s), ctx->dp_beta0, ctx->dp_rowsum, ctx->dp_emean, (double)ctx->N,
                s.gmean, ctx->dp_syo, ctx->dp_E);
  });
  // three sweeps of seven interleaved grid points: {0,3,..,18}, {1,4,..,19}, {2,5,..,20}.  edgeR starts every fit from
  // the same value (start = NULL, R/estimateDisp.par.R:107-110); the intercept is the unique root of a concave score, so
  // starting the second and third sets from the converged neighbour one grid step below gives the same root (to the NR
  // tolerance 1e-10) in about half the sweeps.
  const int DS = DISP_NG / DISP_ND;   // 3
  for (int c = 0; c < DS; ++c) {
    // start values: a later call on the same counts (the next psi refresh of the fit) starts every column from its own
    // converged intercept of the call before -- the parameters have moved less than one grid step does
    if (ctx->dp_beta_warm) k_start_cols<<<nblk((long long)Gl * DISP_ND, 256), 256, 0, ctx->stream>>>(ctx->dp_beta, DISP_NG, c, ctx->dp_beta0, 2.0, Gl, DISP_NG, c, DS, DISP_ND, ctx->dp_beta);
    else if (c == 0) k_start_cols<<<nblk((long long)Gl * DISP_ND, 256), 256, 0, ctx->stream>>>(nullptr, 1, 0, ctx->dp_beta0, 0.0, Gl, DISP_NG, c, DS, DISP_ND, ctx->dp_beta);
    else k_start_cols<<<nblk((long long)Gl * DISP_ND, 256), 256, 0, ctx->stream>>>(ctx->dp_beta, DISP_NG, c - 1, ctx->dp_beta0, 1e300, Gl, DISP_NG, c, DS, DISP_ND, ctx->dp_beta);
    CKL();
    DispArgs a;
    a.phi_grid = ctx->dp_phi_grid + c; a.phi_row = nullptr; a.beta = ctx->dp_beta; a.ldb = DISP_NG; a.d0 = c; a.ds = DS;
    a.emean = nullptr; a.prior_count = 0.0; a.tol = 1e-10; a.notconv = ctx->dp_flag; a.n = (double)ctx->N; a.gm_real = s.gmean; a.Ecache = ctx->dp_E; a.Np = ctx->Np; a.nchunks = ctx->nchunks; a.syo_row = ctx->dp_syo; a.sy_row = ctx->dp_rowsum;
    a.info_out = ctx->dp_info;
    RET(disp_nr(ctx, s, a, DISP_ND));
    const auto t_apl = std::chrono::steady_clock::now();
    DISPATCH_K(ctx->K, {
      if (a.Ecache) {
        k_disp_apl_cached<KT, DISP_ND><<<row_grid(Gl), ROW_THREADS, 0, ctx->stream>>>(geom(ctx, Gl, nullptr), disp_state<KT>(ctx, s), a, ctx->dp_tabG + (long long)c * YT,
                                                                                    ctx->dp_lgr_grid + c, ctx->dp_hist, ctx->dp_l0, DISP_NG);
        CKL();
      } else {
        LAUNCH_ROWS((k_disp_apl<KT, DISP_ND>), 1, Gl, geom(ctx, Gl, nullptr), disp_state<KT>(ctx, s), a, ctx->dp_tabG + (long long)c * YT,
                    ctx->dp_lgr_grid + c, ctx->dp_hist, ctx->dp_l0, DISP_NG);
      }
    });
    if (getenv("RUVNB_TRACE_DISP")) {
      CK(cudaStreamSynchronize(ctx->stream));
      fprintf(stderr, "  disp APL sweep %d: %.2f ms\n", c, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_apl).count());
    }
  }
  ctx->dp_beta_warm = true;
  return RUVNB_OK;
}

extern "C" {

int ruvnb_disp_apl_grid(ruvnb_ctx* ctx, const ruvnb_opts* o, double* l0, double* rowsum) {
  RET(check_ready(ctx, o));
  RET(disp_grid(ctx, ctx->cur));
  const int Gl = ctx->Gl;
  std::vector<double> h((size_t)Gl * DISP_NG), rs(Gl);
  CK(cudaMemcpyAsync(h.data(), ctx->dp_l0, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(rs.data(), ctx->dp_rowsum, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (l0) for (int g = 0; g < Gl; ++g) for (int d = 0; d < DISP_NG; ++d) l0[g + (size_t)Gl * d] = rs[g] >= 5.0 ? h[(size_t)g * DISP_NG + d] : NAN;  // sel, :29
  if (rowsum) memcpy(rowsum, rs.data(), (size_t)Gl * 8);
  return RUVNB_OK;
}

// The whole of estimateDisp.par for design = 1 (R/estimateDisp.par.R:27-201): sets "psi" of the current parameter set to
// the tagwise dispersion.  prior_df < 0: estimate it (limma fitFDist moments; the reference's robust squeezeVar with a
// covariate trend is not restated); prior_df >= 0: use it (injection seam).
int ruvnb_estimate_disp_ex(ruvnb_ctx* ctx, const ruvnb_opts* o, double prior_df_in, double* common_out, double* trended_out,
                           double* prior_df_out) {
  RET(check_ready(ctx, o));
  StateSet& s = ctx->cur;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_start = now();
  RET(disp_grid(ctx, s));
  CK(cudaStreamSynchronize(ctx->stream));
  const double t_grid = now(); const int sweeps_grid = ctx->disp_nr_sweeps;
  const int Gl = ctx->Gl; const long long G = ctx->G, g0 = ctx->g0;
  const double N = (double)ctx->N;
  // gather l0 / rowsum of all genes on every rank (zero-padded sum over the ranks)
  std::vector<double> l0((size_t)G * DISP_NG, 0.0), rowsum(G, 0.0);
  double* d_l0_full = ctx->dp_l0; double* d_rs_full = ctx->dp_rowsum;
  double *tmpA = nullptr, *tmpB = nullptr;
  ScopedDev scratch;
  if (ctx->nranks > 1) {
    CK(scratch.alloc(&tmpA, (size_t)G * DISP_NG * 8)); CK(scratch.alloc(&tmpB, (size_t)G * 8));
    CK(cudaMemsetAsync(tmpA, 0, (size_t)G * DISP_NG * 8, ctx->stream)); CK(cudaMemsetAsync(tmpB, 0, (size_t)G * 8, ctx->stream));
    CK(cudaMemcpyAsync(tmpA + g0 * DISP_NG, ctx->dp_l0, (size_t)Gl * DISP_NG * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(tmpB + g0, ctx->dp_rowsum, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    RET(allreduce(ctx, tmpA, (size_t)G * DISP_NG, NCCL_F64)); RET(allreduce(ctx, tmpB, (size_t)G, NCCL_F64));
    d_l0_full = tmpA; d_rs_full = tmpB;
  }
  CK(cudaMemcpyAsync(l0.data(), d_l0_full, l0.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(rowsum.data(), d_rs_full, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  double pts[DISP_NG];
  for (int i = 0; i < DISP_NG; ++i) pts[i] = -10.0 + (double)i;
  std::vector<int> sel;  // :29 rowSums(y) >= 5
  for (long long g = 0; g < G; ++g) if (rowsum[g] >= 5.0) sel.push_back((int)g);
  const int ns = (int)sel.size();
  if (ns == 0) return ctx->fail(RUVNB_EARG, "estimate_disp: no gene has rowSums(y) >= 5");
  // common dispersion (:130-132)
  double colsum[DISP_NG] = {0};
  for (int i : sel) for (int d = 0; d < DISP_NG; ++d) colsum[d] += l0[(size_t)i * DISP_NG + d];
  const double common = 0.1 * exp2(maximize_interpolant_h(DISP_NG, pts, colsum));
  if (common_out) *common_out = common;
  const double t_common = now();
  // AveLogCPM (:134): one-group fit of y + prior at the common dispersion
  {
    k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->dp_phi_row, Gl, common); CKL();
    // edgeR's start is log(sum (y + p)/(e^o + 2p) / n) = log((exp(beta0) + pcm) / (1 + 2 pcm)), pcm = prior.count / mean e^o;
    // the same shift of the grid's converged intercept at the common dispersion is closer to the (unique) root
    k_beta_from_grid<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(ctx->dp_beta, ctx->dp_phi_grid, DISP_NG, ctx->dp_phi_row, ctx->dp_emean, 2.0, Gl, ctx->dp_beta1); CKL();
    DispArgs a;
    a.phi_grid = nullptr; a.phi_row = ctx->dp_phi_row; a.beta = ctx->dp_beta1; a.ldb = 1; a.d0 = 0; a.ds = 1;
    a.emean = ctx->dp_emean; a.prior_count = 2.0; a.tol = 1e-10; a.notconv = ctx->dp_flag; a.n = N; a.gm_real = s.gmean; a.Ecache = ctx->dp_E; a.Np = ctx->Np; a.nchunks = ctx->nchunks; a.syo_row = ctx->dp_syo; a.sy_row = ctx->dp_rowsum; a.info_out = nullptr;
    RET(disp_nr(ctx, s, a, 1));
  }
  std::vector<double> alc(G, 0.0);
  {
    std::vector<double> b1(Gl);
    CK(cudaMemcpyAsync(b1.data(), ctx->dp_beta1, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int g = 0; g < Gl; ++g) alc[g0 + g] = (b1[g] + log(1e6)) / log(2.0);
    if (ctx->nranks > 1) {
      CK(cudaMemcpyAsync(tmpB, alc.data(), (size_t)G * 8, cudaMemcpyHostToDevice, ctx->stream));
      RET(allreduce(ctx, tmpB, (size_t)G, NCCL_F64));
      CK(cudaMemcpyAsync(alc.data(), tmpB, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
  }
  const double t_alc = now(); const int sweeps_alc = ctx->disp_nr_sweeps;
  // trend: local-constant tricube smoother of l0 against AveLogCPM over the selected genes (:136-139, locfit stand-in)
  const double span = ns <= 50 ? 1.0 : 0.25 + 0.75 * sqrt(50.0 / (double)ns);  // edgeR::WLEB default
  const int q = std::min(ns, std::max(1, (int)ceil(span * (double)ns)));
  std::vector<int> order(ns);
  for (int i = 0; i < ns; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return alc[sel[a]] < alc[sel[b]]; });
  std::vector<double> xs(ns); std::vector<int> perm(ns);
  for (int i = 0; i < ns; ++i) { xs[i] = alc[sel[order[i]]]; perm[i] = sel[order[i]]; }
  std::vector<double> m0s((size_t)ns * DISP_NG);
  {
    double *d_xs = ctx->dp_xs, *d_m0 = ctx->dp_m0, *d_l0g = nullptr; int* d_perm = ctx->dp_perm;   // ns <= G
    if (ctx->nranks > 1) d_l0g = tmpA; else d_l0g = ctx->dp_l0;
    CK(cudaMemcpyAsync(d_xs, xs.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_perm, perm.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_tricube<<<ns, 128, 0, ctx->stream>>>(d_xs, d_perm, ns, q, d_l0g, DISP_NG, d_m0); CKL();
    CK(cudaMemcpyAsync(m0s.data(), d_m0, m0s.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  // m0 back to gene order; trended dispersion (:140-145)
  std::vector<double> m0((size_t)G * DISP_NG, 0.0), trended(G, 0.0);
  int i_min = 0;
  for (int i = 0; i < ns; ++i) {
    memcpy(&m0[(size_t)perm[i] * DISP_NG], &m0s[(size_t)i * DISP_NG], DISP_NG * 8);
    trended[perm[i]] = 0.1 * exp2(maximize_interpolant_h(DISP_NG, pts, &m0s[(size_t)i * DISP_NG]));
  }
  for (int i = 1; i < ns; ++i) if (alc[sel[i]] < alc[sel[i_min]]) i_min = i;   // which.min(AveLogCPM[sel])
  { std::vector<char> issel(G, 0); for (int i : sel) issel[i] = 1; for (long long g = 0; g < G; ++g) if (!issel[g]) trended[g] = trended[sel[i_min]]; }
  if (trended_out) memcpy(trended_out, trended.data() + g0, (size_t)Gl * 8);
  const double t_trend = now();
  // prior d.f. (:156-169): deviance at the trended dispersion -> limma fitFDist moments
  double prior_df = prior_df_in;
  if (!(prior_df >= 0.0)) {
    CK(cudaMemcpyAsync(ctx->dp_phi_row, trended.data() + g0, (size_t)Gl * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_beta_from_grid<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(ctx->dp_beta, ctx->dp_phi_grid, DISP_NG, ctx->dp_phi_row, nullptr, 0.0, Gl, ctx->dp_beta1); CKL();
    DispArgs a;
    a.phi_grid = nullptr; a.phi_row = ctx->dp_phi_row; a.beta = ctx->dp_beta1; a.ldb = 1; a.d0 = 0; a.ds = 1;
    a.emean = nullptr; a.prior_count = 0.0; a.tol = 1e-10; a.notconv = ctx->dp_flag; a.n = N; a.gm_real = s.gmean; a.Ecache = ctx->dp_E; a.Np = ctx->Np; a.nchunks = ctx->nchunks; a.syo_row = ctx->dp_syo; a.sy_row = ctx->dp_rowsum; a.info_out = nullptr;
    RET(disp_nr(ctx, s, a, 1));
    k_build_ytabs<<<Gl, YT, 0, ctx->stream>>>(ctx->dp_phi_row, Gl, nullptr, ctx->dp_tabR2, nullptr); CKL();
    DISPATCH_K(ctx->K, {
      RowGeom gg = geom(ctx, Gl, nullptr);
      gg.tabR = ctx->dp_tabR2;  // log(y + r) for r = 1/trended
      LAUNCH_ROWS(k_disp_dev<KT>, 1, Gl, gg, disp_state<KT>(ctx, s), ctx->dp_beta1, ctx->dp_phi_row, s.gmean, ctx->dp_dev);
    });
    std::vector<double> dev(G, 0.0);
    CK(cudaMemcpyAsync(dev.data() + g0, ctx->dp_dev, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->nranks > 1) {
      CK(cudaMemcpyAsync(tmpB, dev.data(), (size_t)G * 8, cudaMemcpyHostToDevice, ctx->stream));
      RET(allreduce(ctx, tmpB, (size_t)G, NCCL_F64));
      CK(cudaMemcpyAsync(dev.data(), tmpB, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<double> s2(ns);
    for (int i = 0; i < ns; ++i) s2[i] = std::max(dev[sel[i]] / (N - 1.0), 0.0);
    prior_df = fit_f_dist_df2(s2, N - 1.0);
  }
  if (prior_df_out) *prior_df_out = prior_df;
  const double t_prior = now();
  const double prior_n = prior_df / (N - 1.0);  // :171 (ncoefs = 1)
  // tagwise (:172-190): per gene, maximise the spline through l0 + prior.n m0
  std::vector<double> tag(trended);
  if (!(prior_n > 1e6)) {
    double l0a[DISP_NG];
    for (int i : sel) {
      for (int d = 0; d < DISP_NG; ++d) l0a[d] = l0[(size_t)i * DISP_NG + d] + prior_n * m0[(size_t)i * DISP_NG + d];
      tag[i] = 0.1 * exp2(maximize_interpolant_h(DISP_NG, pts, l0a));
    }
  }
  // psi <- tagwise where finite (R/ruvIIInb.R:564)
  std::vector<double> old(Gl);
  CK(cudaMemcpyAsync(old.data(), s.psi, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int g = 0; g < Gl; ++g) if (std::isfinite(tag[g0 + g])) old[g] = tag[g0 + g];
  CK(cudaMemcpyAsync(s.psi, old.data(), (size_t)Gl * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->psi_ver++; s.sig_valid = false;
  if (o->verbose)
    fprintf(stderr, "estimateDisp: grid APL %.3f s (%d NR sweeps + 3 APL sweeps), gather + common %.3f s, AveLogCPM %.3f s (%d sweeps), "
                    "trend %.3f s, prior d.f. %.3f s (%d sweeps + deviance), tagwise %.3f s\n",
            t_grid - t_start, sweeps_grid, t_common - t_grid, t_alc - t_common, sweeps_alc - sweeps_grid, t_trend - t_alc,
            t_prior - t_trend, ctx->disp_nr_sweeps - sweeps_alc, now() - t_prior);
  return RUVNB_OK;
}

int ruvnb_estimate_disp(ruvnb_ctx* ctx, const ruvnb_opts* o, double* common_out) {
  return ruvnb_estimate_disp_ex(ctx, o, -1.0, common_out, nullptr, nullptr);
}

}  // extern "C"

// ---- get.res (R/get.res.R:12-129) ----------------------------------------------------------------------------------------
extern "C" int ruvnb_get_res(ruvnb_ctx* ctx, int type, int64_t block_size, const double* Mb_cells, double* out) {
  if (!ctx || !out) return RUVNB_EARG;
  if (type < RES_PEARSON || type > RES_QUANTILE) return ctx->fail(RUVNB_EARG, "get_res: type %d is not 1 (pearson), 2 (logcounts) or 3 (quantile)", type);
  if (!ctx->have_counts || ctx->K == 0 || !ctx->have_W) return ctx->fail(RUVNB_ESTATE, "get_res: needs counts and a fitted state (W, alpha, gmean, Mb, psi)");
  CK(cudaSetDevice(ctx->device));
  const int Gl = ctx->Gl, K = ctx->K;
  const long long N = ctx->N;
  long long bs = std::min<long long>(block_size > 0 ? block_size : 5000, N);   // :18
  if (bs > 12000) return ctx->fail(RUVNB_EARG, "get_res: block.size %lld exceeds the 12000 cells a per-gene shared-memory tile holds", bs);
  StateSet& s = ctx->cur;
  double *wm = nullptr, *wam = nullptr, *logcap = nullptr, *slab[2] = {nullptr, nullptr}, *mbblk = nullptr;
  unsigned* hist = nullptr; SelTargets* T = nullptr; int* nanc = nullptr;
  cudaStream_t copy_stream = nullptr; cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  CK(cudaMalloc((void**)&wm, K * 8)); CK(cudaMalloc((void**)&wam, (size_t)Gl * 8)); CK(cudaMalloc((void**)&logcap, (size_t)Gl * 3 * 8));
  for (int i = 0; i < 2; ++i) { CK(cudaMalloc((void**)&slab[i], (size_t)Gl * bs * 8)); CK(cudaEventCreate(&done[i])); CK(cudaEventCreate(&copied[i])); }
  if (Mb_cells) CK(cudaMalloc((void**)&mbblk, (size_t)Gl * bs * 8));
  CK(cudaMalloc((void**)&hist, SEL_NT * SEL_BINS * 4)); CK(cudaMalloc((void**)&T, sizeof(SelTargets))); CK(cudaMalloc((void**)&nanc, 4));
  CK(cudaMemsetAsync(hist, 0, SEL_NT * SEL_BINS * 4, ctx->stream));
  CK(cudaStreamCreate(&copy_stream));
  k_res_wmean<<<K, 1024, 0, ctx->stream>>>(s.W, ctx->Np, (double)N, wm); CKL();
  k_res_wamean<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(Gl, K, s.alpha, wm, wam); CKL();
  CK(cudaFuncSetAttribute(k_res_caps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * bs * 8)));
  CK(cudaFuncSetAttribute(k_sel_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_NT * SEL_BINS * 4));
  int blk = 0;
  for (long long c0 = 0; c0 < N; c0 += bs, ++blk) {
    const int B = (int)std::min<long long>(bs, N - c0);
    const int sl = blk & 1;
    if (blk >= 2) CK(cudaStreamWaitEvent(ctx->stream, copied[sl], 0));   // the slab's previous contents reached the host
    ResArgs a;
    a.Gl = Gl; a.K = K; a.m = ctx->m; a.B = B; a.N = N; a.Np = ctx->Np; a.c0 = c0; a.Yp = ctx->dY;
    a.alpha = s.alpha; a.gmean = s.gmean; a.psi = s.psi; a.W = s.W; a.Mb = s.Mb; a.mb_blk = nullptr;
    a.cell2pos = ctx->d_cell2pos; a.chunk_grp = ctx->d_chunk_grp; a.wa_mean = wam; a.logcap = logcap;
    if (Mb_cells) {  // host G x N column-major: the block is one contiguous Gl x B piece == device [B][Gl]
      CK(cudaMemcpyAsync(mbblk, Mb_cells + (size_t)c0 * Gl, (size_t)Gl * B * 8, cudaMemcpyHostToDevice, ctx->stream));
      a.mb_blk = mbblk;
    }
    if (type != RES_LOGCOUNTS) { k_res_caps<<<Gl, 256, (size_t)2 * B * 8, ctx->stream>>>(a); CKL(); }
    dim3 grid(nblk(Gl, 32), nblk(B, 32));
    if (type == RES_PEARSON) {
      CK(cudaMemsetAsync(nanc, 0, 4, ctx->stream));
      k_res_elem<RES_PEARSON><<<grid, 1024, 0, ctx->stream>>>(a, slab[sl], nanc); CKL();
      // clip at the type-7 0.5 % / 99.5 % quantiles of ALL genes x this block's cells (:63-66)
      long long len = (long long)Gl * B, ntot = len;
      if (ctx->nranks > 1) { ntot = (long long)ctx->G * B; RET(allreduce(ctx, nanc, 1, NCCL_INT32)); }
      k_sel_init<<<1, 1, 0, ctx->stream>>>(T, ntot, nanc); CKL();
      const int shifts[6] = {52, 40, 28, 16, 4, 0}, bits[6] = {12, 12, 12, 12, 12, 4};
      for (int lv = 0; lv < 6; ++lv) {
        k_sel_hist<<<296, 256, SEL_NT * SEL_BINS * 4, ctx->stream>>>(slab[sl], len, shifts[lv], bits[lv], T, hist); CKL();
        RET(allreduce(ctx, hist, SEL_NT * SEL_BINS, NCCL_INT32));
        k_sel_pick<<<1, SEL_NT * 32, 0, ctx->stream>>>(T, hist, shifts[lv], bits[lv]); CKL();
      }
      k_sel_finish<<<1, 1, 0, ctx->stream>>>(T); CKL();
      k_res_clip<<<nblk(len, 256), 256, 0, ctx->stream>>>(slab[sl], len, T); CKL();
    } else if (type == RES_LOGCOUNTS) {
      k_res_elem<RES_LOGCOUNTS><<<grid, 1024, 0, ctx->stream>>>(a, slab[sl], nanc); CKL();
    } else {
      k_res_elem<RES_QUANTILE><<<grid, 1024, 0, ctx->stream>>>(a, slab[sl], nanc); CKL();
    }
    CK(cudaEventRecord(done[sl], ctx->stream));
    CK(cudaStreamWaitEvent(copy_stream, done[sl], 0));
    CK(cudaMemcpyAsync(out + (size_t)c0 * Gl, slab[sl], (size_t)Gl * B * 8, cudaMemcpyDeviceToHost, copy_stream));  // write_block, :125
    CK(cudaEventRecord(copied[sl], copy_stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(copy_stream));
  CK(cudaStreamDestroy(copy_stream));
  for (int i = 0; i < 2; ++i) { CK(cudaFree(slab[i])); CK(cudaEventDestroy(done[i])); CK(cudaEventDestroy(copied[i])); }
  CK(cudaFree(wm)); CK(cudaFree(wam)); CK(cudaFree(logcap)); CK(cudaFree(hist)); CK(cudaFree(T)); CK(cudaFree(nanc));
  if (mbblk) CK(cudaFree(mbblk));
  return RUVNB_OK;
}

// ---- fastruvIII.nb streaming passes (R/fastruvIIInb.R:632-748) -----------------------------------------------------------
// R median of values v[i] each repeated cnt[i] times (host; m is small)
static double weighted_median_h(std::vector<std::pair<double, long long>> vc) {
  std::sort(vc.begin(), vc.end());
  long long n = 0;
  for (auto& p : vc) n += p.second;
  if (n == 0) return NAN;
  auto kth = [&](long long k) { long long acc = 0; for (auto& p : vc) { acc += p.second; if (k < acc) return p.first; } return vc.back().first; };
  return (n & 1) ? kth(n / 2) : 0.5 * (kth(n / 2 - 1) + kth(n / 2));
}

extern "C" {

// F5: W for all cells.  The current state supplies alpha and gmean (all genes); wt_ctl: n_ctl gene weights
// (best.wtctl = rowMeans(sig.inv[ctl,]) of the subsample fit); W_init: N x k (zeros, Wsub rows filled in, :635-636);
// W_med / W_mad: colMedians / colMads of Wsub (:653-654).  W_out: N x k; the resident "W" is updated too.
int ruvnb_fast_W_all(ruvnb_ctx* ctx, int k, double lambda_a, const double* wt_ctl, const double* W_init, const double* W_med,
                     const double* W_mad, int max_sweeps, double tol, double* W_out, int* n_sweeps, double* crit_out) {
  if (!ctx || !wt_ctl || !W_init || !W_med || !W_mad) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "fast_W_all: no counts uploaded");
  if (ctx->nranks > 1) return ctx->fail(RUVNB_ESTATE, "fast_W_all: single-GPU only in this build (cells are independent: shard N across contexts)");
  RET(ensure_state(ctx, k));
  CK(cudaSetDevice(ctx->device));
  RET(ruvnb_set_state(ctx, "W", W_init, k));
  const int K = ctx->K, nc = ctx->n_ctl, KK = K * (K + 1) / 2;
  StateSet& s = ctx->cur;
  double *d_gm = nullptr, *d_al = nullptr, *d_wt = nullptr, *d_A = nullptr, *d_med = nullptr, *d_mad = nullptr, *d_crit = nullptr;
  const unsigned grid = nblk(ctx->Np, 64);
  CK(cudaMalloc((void**)&d_gm, (size_t)nc * 8)); CK(cudaMalloc((void**)&d_al, (size_t)nc * K * 8)); CK(cudaMalloc((void**)&d_wt, (size_t)nc * 8));
  CK(cudaMalloc((void**)&d_A, KK * 8)); CK(cudaMalloc((void**)&d_med, K * 8)); CK(cudaMalloc((void**)&d_mad, K * 8)); CK(cudaMalloc((void**)&d_crit, (size_t)grid * 8));
  k_gather_rows<<<nblk(nc, 256), 256, 0, ctx->stream>>>(s.gmean, ctx->d_ctl_local, nc, 1, d_gm); CKL();
  k_gather_rows<<<nblk((long long)nc * K, 256), 256, 0, ctx->stream>>>(s.alpha, ctx->d_ctl_local, nc, K, d_al); CKL();
  CK(cudaMemcpyAsync(d_wt, wt_ctl, (size_t)nc * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_med, W_med, K * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_mad, W_mad, K * 8, cudaMemcpyHostToDevice, ctx->stream));
  DISPATCH_K(K, { k_fast_gram<KT><<<1, 32, 0, ctx->stream>>>(nc, d_al, d_wt, d_A); CKL(); });
  std::vector<double> part(grid);
  int sweeps = 0; double crit = NAN;
  // the initial pass (:637-651) plus up to max_sweeps refinement sweeps (:664-713)
  for (int it = 0; it <= max_sweeps; ++it) {
    DISPATCH_K(K, {
      k_fast_w_sweep<KT><<<grid, 256, 0, ctx->stream>>>(ctx->d_ctl_rows, nc, ctx->Np, ctx->d_chunk_valid, d_gm, d_al, d_wt, d_A, lambda_a, d_med, d_mad,
                                                      ctx->d_mtabs, s.W, ctx->nw.W, d_crit);
      CKL();
    });
    std::swap(s.W, ctx->nw.W);
    sweeps++;
    CK(cudaMemcpyAsync(part.data(), d_crit, (size_t)grid * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double t = 0.0;
    for (double v : part) t += v;
    crit = t / ((double)ctx->N * K);                       // mean over all N x k entries (:709)
    if (it >= 1 && crit < tol) break;                      // the initial pass has no convergence test
  }
  s.sig_valid = false;
  if (n_sweeps) *n_sweeps = sweeps;
  if (crit_out) *crit_out = crit;
  CK(cudaFree(d_gm)); CK(cudaFree(d_al)); CK(cudaFree(d_wt)); CK(cudaFree(d_A)); CK(cudaFree(d_med)); CK(cudaFree(d_mad)); CK(cudaFree(d_crit));
  if (W_out) RET(ruvnb_get_state(ctx, "W", W_out));
  return RUVNB_OK;
}

// F6: Mb.all (G x N, column-major, streamed to the host in blocks of block_size cells).  annot_grp[c] = replicate group
// of cell c or -1; the current state supplies alpha, gmean, psi, W and the subsample fit's Mb (G x m).
int ruvnb_fast_Mb_all(ruvnb_ctx* ctx, double lambda_b, const int32_t* annot_grp, int64_t block_size, double* Mb_all_out) {
  if (!ctx || !annot_grp || !Mb_all_out) return RUVNB_EARG;
  if (!ctx->have_counts || ctx->K == 0 || !ctx->have_W) return ctx->fail(RUVNB_ESTATE, "fast_Mb_all: needs counts and a state (W, alpha, gmean, Mb, psi)");
  CK(cudaSetDevice(ctx->device));
  const int Gl = ctx->Gl, m = ctx->m; const long long N = ctx->N;
  StateSet& s = ctx->cur;
  // row median / mad over the annotated columns (:744-745): values Mb[g, j] with multiplicity = annotated cells of group j
  std::vector<long long> cnt(m, 0);
  for (long long c = 0; c < N; ++c) { int j = annot_grp[c]; if (j >= m) return ctx->fail(RUVNB_EARG, "fast_Mb_all: group %d outside 0..%d", j, m - 1); if (j >= 0) cnt[j]++; }
  std::vector<double> Mbh((size_t)Gl * m), rmed(Gl), rmad(Gl);
  CK(cudaMemcpyAsync(Mbh.data(), s.Mb, Mbh.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::vector<uint8_t> ctl_local(Gl);
  for (int g = 0; g < Gl; ++g) ctl_local[g] = ctx->ctl[ctx->g0 + g];
  for (int g = 0; g < Gl; ++g) {
    std::vector<std::pair<double, long long>> vc(m);
    for (int j = 0; j < m; ++j) vc[j] = {ctl_local[g] ? 0.0 : Mbh[(size_t)g * m + j], cnt[j]};
    double med = weighted_median_h(vc);
    for (int j = 0; j < m; ++j) vc[j].first = fabs(vc[j].first - med);
    rmed[g] = med; rmad[g] = 1.4826 * weighted_median_h(vc);
  }
  long long bs = std::min<long long>(block_size > 0 ? block_size : 5000, N);
  double *d_rmed = nullptr, *d_rmad = nullptr, *slab[2] = {nullptr, nullptr}; int* d_annot = nullptr; uint8_t* d_ctl = nullptr;
  cudaStream_t copy_stream = nullptr; cudaEvent_t done[2], copied[2];
  CK(cudaMalloc((void**)&d_rmed, (size_t)Gl * 8)); CK(cudaMalloc((void**)&d_rmad, (size_t)Gl * 8)); CK(cudaMalloc((void**)&d_annot, (size_t)N * 4));
  CK(cudaMalloc((void**)&d_ctl, Gl));
  for (int i = 0; i < 2; ++i) { CK(cudaMalloc((void**)&slab[i], (size_t)Gl * bs * 8)); CK(cudaEventCreate(&done[i])); CK(cudaEventCreate(&copied[i])); }
  CK(cudaStreamCreate(&copy_stream));
  CK(cudaMemcpyAsync(d_rmed, rmed.data(), (size_t)Gl * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_rmad, rmad.data(), (size_t)Gl * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_annot, annot_grp, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_ctl, ctl_local.data(), Gl, cudaMemcpyHostToDevice, ctx->stream));
  int blk = 0;
  for (long long c0 = 0; c0 < N; c0 += bs, ++blk) {
    const int B = (int)std::min<long long>(bs, N - c0), sl = blk & 1;
    if (blk >= 2) CK(cudaStreamWaitEvent(ctx->stream, copied[sl], 0));
    FastMbArgs a;
    a.Gl = Gl; a.K = ctx->K; a.m = m; a.B = B; a.Np = ctx->Np; a.c0 = c0; a.Yp = ctx->dY; a.alpha = s.alpha; a.gmean = s.gmean; a.psi = s.psi;
    a.W = s.W; a.Mb = s.Mb; a.cell2pos = ctx->d_cell2pos; a.annot = d_annot; a.ctl_local = d_ctl; a.rmed = d_rmed; a.rmad = d_rmad; a.lambda_b = lambda_b;
    k_fast_mb<<<dim3(nblk(Gl, 32), nblk(B, 32)), 1024, 0, ctx->stream>>>(a, slab[sl]); CKL();
    CK(cudaEventRecord(done[sl], ctx->stream));
    CK(cudaStreamWaitEvent(copy_stream, done[sl], 0));
    CK(cudaMemcpyAsync(Mb_all_out + (size_t)c0 * Gl, slab[sl], (size_t)Gl * B * 8, cudaMemcpyDeviceToHost, copy_stream));
    CK(cudaEventRecord(copied[sl], copy_stream));
  }
  CK(cudaStreamSynchronize(ctx->stream)); CK(cudaStreamSynchronize(copy_stream));
  CK(cudaStreamDestroy(copy_stream));
  for (int i = 0; i < 2; ++i) { CK(cudaFree(slab[i])); CK(cudaEventDestroy(done[i])); CK(cudaEventDestroy(copied[i])); }
  CK(cudaFree(d_rmed)); CK(cudaFree(d_rmad)); CK(cudaFree(d_annot)); CK(cudaFree(d_ctl));
  return RUVNB_OK;
}

// F7: psi capped at exp(median(log psi) + 3 mad(log psi)) (R/fastruvIIInb.R:765-770); host-side, in place
int ruvnb_fast_cap_psi(double* psi, int64_t G) {
  if (!psi || G <= 0) return RUVNB_EARG;
  std::vector<double> l(G);
  for (int64_t i = 0; i < G; ++i) l[i] = log(psi[i]);
  auto median = [](std::vector<double> v) { size_t n = v.size(); std::sort(v.begin(), v.end()); return (n & 1) ? v[n / 2] : 0.5 * (v[n / 2 - 1] + v[n / 2]); };
  double med = median(l);
  for (auto& v : l) v = fabs(v - med);
  double cap = exp(med + 3.0 * 1.4826 * median(l));
  for (int64_t i = 0; i < G; ++i) if (psi[i] > cap) psi[i] = cap;
  return RUVNB_OK;
}

}  // extern "C"

// the north_star variant of the dispersion step: per-gene Newton on log psi at the current fitted means (disp.cuh)
extern "C" int ruvnb_disp_newton(ruvnb_ctx* ctx, const ruvnb_opts* o, double* psi_out, int maxit) {
  RET(check_ready(ctx, o));
  if (!psi_out) return RUVNB_EARG;
  RET(disp_ensure(ctx));
  const int Gl = ctx->Gl;
  StateSet& s = ctx->cur;
  double* theta = ctx->dp_beta1;
  k_log_vec<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(s.psi, Gl, theta); CKL();   // start from the current psi
  if (maxit <= 0) maxit = 30;
  int it = 0;
  for (; it < maxit; ++it) {
    CK(cudaMemsetAsync(ctx->dp_flag, 0, sizeof(int), ctx->stream));
    DISPATCH_K(ctx->K, {
      RowGeom gg = geom(ctx, Gl, nullptr);
      gg.tabL = nullptr;   // the op fills its own per-iteration tables in the warp's shared-memory slots
      LAUNCH_ROWS(k_psi_newton<KT>, 1, Gl, gg, gstate<KT>(ctx, s, false), theta, 1e-8, ctx->dp_flag);
    });
    CK(cudaMemcpyAsync(ctx->h_flags + 7, ctx->dp_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_flags[7] == 0) { ++it; break; }
  }
  ctx->disp_nr_sweeps = it;
  k_exp_vec<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(theta, Gl, ctx->dp_phi_row); CKL();
  CK(cudaMemcpyAsync(psi_out, ctx->dp_phi_row, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RUVNB_OK;
}

// ---- fastruvIII.nb subsample fit (R/fastruvIIInb.R:223-623): F3 initial estimates, F4 one inner iteration ----------------
static FFGeom ffgeom(ruvnb_ctx* ctx) {
  FFGeom g; g.Yp = ctx->dY; g.Np = ctx->Np; g.Gl = ctx->Gl; g.K = ctx->K; g.m = ctx->m; g.chunk_grp = ctx->d_chunk_grp; g.chunk_valid = ctx->d_chunk_valid;
  return g;
}
static int ff_ensure(ruvnb_ctx* ctx) {
  if (ctx->ff_ready) return RUVNB_OK;
  if (ctx->nranks > 1) return ctx->fail(RUVNB_ESTATE, "the subsample fit runs on one GPU (it holds <= 3000 cells)");
  const long long Gl = ctx->Gl, Np = ctx->Np, m = ctx->m, K = ctx->K, nc = ctx->n_ctl;
  RET(dev_alloc(ctx, &ctx->ff_Z, Gl * Np)); RET(dev_alloc(ctx, &ctx->ff_SI, Gl * Np)); RET(dev_alloc(ctx, &ctx->ff_ZS, Gl * m));
  RET(dev_alloc(ctx, &ctx->ff_sirow, Gl)); RET(dev_alloc(ctx, &ctx->ff_wtcell, Np)); RET(dev_alloc(ctx, &ctx->ff_step, Np));
  RET(dev_alloc(ctx, &ctx->ff_cll, Np)); RET(dev_alloc(ctx, &ctx->ff_cll_tmp, Np)); RET(dev_alloc(ctx, &ctx->ff_b, Gl * K));
  RET(dev_alloc(ctx, &ctx->ff_Ginv, K * K)); RET(dev_alloc(ctx, &ctx->ff_alpha_c, nc * K)); RET(dev_alloc(ctx, &ctx->ff_A, K * (K + 1) / 2));
  RET(dev_alloc(ctx, &ctx->ff_ones, std::max(Np, nc))); RET(dev_alloc(ctx, &ctx->ff_ctlmask, Gl)); RET(dev_alloc(ctx, &ctx->ff_ctl_rows_idx, nc));
  k_fill<<<nblk(std::max(Np, nc), 256), 256, 0, ctx->stream>>>(ctx->ff_ones, std::max(Np, nc), 1.0); CKL();
  k_fill<<<nblk(Np, 256), 256, 0, ctx->stream>>>(ctx->ff_step, Np, 1.0); CKL();
  std::vector<uint8_t> mask(Gl); std::vector<int> idx(nc);
  for (long long g = 0; g < Gl; ++g) mask[g] = ctx->ctl[ctx->g0 + g];
  for (long long i = 0; i < nc; ++i) idx[i] = ctx->ctl_gene[i];
  CK(cudaMemcpyAsync(ctx->ff_ctlmask, mask.data(), Gl, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->ff_ctl_rows_idx, idx.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->ff_ready = true;
  return RUVNB_OK;
}
// (X'DX + lambda I)^-1 for the K x K matrix X'DX given row-major on the host (Gauss-Jordan; chol2inv in the reference)
static bool ff_inverse(int K, std::vector<double>& M) {
  std::vector<double> inv(K * K, 0.0);
  for (int i = 0; i < K; ++i) inv[i * K + i] = 1.0;
  for (int c = 0; c < K; ++c) {
    int piv = c;
    for (int r = c + 1; r < K; ++r) if (fabs(M[r * K + c]) > fabs(M[piv * K + c])) piv = r;
    if (!(fabs(M[piv * K + c]) > 0.0)) return false;
    for (int q = 0; q < K; ++q) { std::swap(M[c * K + q], M[piv * K + q]); std::swap(inv[c * K + q], inv[piv * K + q]); }
    double d = M[c * K + c];
    for (int q = 0; q < K; ++q) { M[c * K + q] /= d; inv[c * K + q] /= d; }
    for (int r = 0; r < K; ++r) if (r != c) {
      double f = M[r * K + c];
      for (int q = 0; q < K; ++q) { M[r * K + q] -= f * M[c * K + q]; inv[r * K + q] -= f * inv[c * K + q]; }
    }
  }
  M = inv;
  return true;
}
// alpha <- (sum_c v_c x_c x_c' + lambda I)^-1-weighted projection of b: Gram of the [K][Np] matrix X with cell weights v on the host
static int ff_alpha_from_b(ruvnb_ctx* ctx, const double* X /* dev [K][Np] */, const double* v /* dev [Np] */, double lambda_a, double* alpha_out, int* nbad) {
  const int K = ctx->K; const long long Np = ctx->Np;
  std::vector<double> Xh((size_t)K * Np), vh(Np), Gm(K * K, 0.0);
  CK(cudaMemcpyAsync(Xh.data(), X, Xh.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(vh.data(), v, (size_t)Np * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < K; ++i) for (int j = 0; j <= i; ++j) {
    double s = 0.0;
    for (long long p = 0; p < Np; ++p) s += vh[p] * Xh[(size_t)i * Np + p] * Xh[(size_t)j * Np + p];   // pads hold X = 0
    Gm[i * K + j] = Gm[j * K + i] = s;
  }
  for (int i = 0; i < K; ++i) Gm[i * K + i] += lambda_a;
  if (!ff_inverse(K, Gm)) for (auto& x : Gm) x = NAN;
  CK(cudaMemcpyAsync(ctx->ff_Ginv, Gm.data(), (size_t)K * K * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->flags + 0, 0, sizeof(int), ctx->stream));
  k_ff_alpha<<<nblk(ctx->Gl, 128), 128, 0, ctx->stream>>>(ctx->Gl, K, ctx->ff_b, ctx->ff_Ginv, alpha_out, ctx->flags + 0); CKL();
  CK(cudaMemcpyAsync(ctx->h_flags, ctx->flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *nbad = ctx->h_flags[0];
  return RUVNB_OK;
}
static int ff_clamp_alpha(ruvnb_ctx* ctx, double* alpha) {   // :233-238,421-426
  k_strided_medmad<<<ctx->K, RS_NT, 0, ctx->stream>>>(alpha, 1, ctx->K, ctx->Gl, ctx->med, ctx->mad); CKL();
  k_clamp_cols<<<nblk((long long)ctx->Gl * ctx->K, 256), 256, 0, ctx->stream>>>(alpha, ctx->Gl, ctx->K, ctx->med, ctx->mad); CKL();
  return RUVNB_OK;
}
// Wsub from the stored Z (control genes), gene weights wt (device [n_ctl]) and the NEW alpha; returns non-finite count
static int ff_update_W(ruvnb_ctx* ctx, const double* gmean_old, const double* alpha_new, const double* wt, double lambda_a, double* W_out, int* nbad) {
  const int nc = ctx->n_ctl, K = ctx->K;
  k_gather_rows<<<nblk((long long)nc * K, 256), 256, 0, ctx->stream>>>(alpha_new, ctx->d_ctl_local, nc, K, ctx->ff_alpha_c); CKL();
  CK(cudaMemsetAsync(ctx->flags + 1, 0, sizeof(int), ctx->stream));
  DISPATCH_K(K, {
    k_fast_gram<KT><<<1, 32, 0, ctx->stream>>>(nc, ctx->ff_alpha_c, wt, ctx->ff_A); CKL();
    k_ff_w<KT><<<nblk(ctx->Np, 64), 256, 0, ctx->stream>>>(ffgeom(ctx), ctx->ff_ctl_rows_idx, nc, ctx->ff_Z, gmean_old, ctx->ff_alpha_c, wt, ctx->ff_A, lambda_a,
                                                          W_out, ctx->flags + 1);
    CKL();
  });
  CK(cudaMemcpyAsync(ctx->h_flags + 1, ctx->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *nbad = ctx->h_flags[1];
  return RUVNB_OK;
}
static int ff_cell_loglik(ruvnb_ctx* ctx, StateSet& s, double* out, double* total) {
  RET(ensure_ytabs(ctx, s.psi));
  k_ff_cell_loglik<<<nblk(ctx->Np, 128), 128, 0, ctx->stream>>>(ffgeom(ctx), s.gmean, s.Mb, s.alpha, s.W, s.psi, ctx->tabL, ctx->lgr, out); CKL();
  k_colsum<<<1, 1024, 0, ctx->stream>>>(out, ctx->Np, 1, ctx->scal); CKL();
  CK(cudaMemcpyAsync(ctx->h_scal, ctx->scal, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (total) *total = ctx->h_scal[0];
  return RUVNB_OK;
}

extern "C" {

// F3 (:223-271): Z = log(Y+1), gmean, alpha = Z U0 (U0'U0 + lambda.a I)^-1 (U0: nsub x k, the irlba result -- injected),
// clamps, Wsub by sequential ridge regressions (unit weights) clamped to median +- 4 mad, Mb = unweighted group means, Mb[ctl,] = 0
int ruvnb_fast_init(ruvnb_ctx* ctx, const ruvnb_opts* o, const double* U0) {
  if (!ctx || !o || !U0) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "fast_init: no counts uploaded");
  RET(ruvnb_set_state(ctx, "W", U0, o->k));   // U0 in the W slot: [K][Np], pads 0
  RET(ff_ensure(ctx));
  CK(cudaSetDevice(ctx->device));
  StateSet& c = ctx->cur;
  const int Gl = ctx->Gl, K = ctx->K, m = ctx->m; const long long Np = ctx->Np, N = ctx->N;
  FFGeom ge = ffgeom(ctx);
  k_ff_init_z<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(ge, (double)N, ctx->ff_Z, c.gmean, ctx->ff_ZS); CKL();
  // b = Z U0: k_ff_b with zero group means and unit cell weights
  CK(cudaMemsetAsync(ctx->ff_ZS, 0, (size_t)Gl * m * 8, ctx->stream));
  k_ff_b<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(ge, ctx->ff_Z, ctx->ff_ZS, ctx->d_group_n, ctx->ff_ones, c.W, ctx->ff_b); CKL();
  int nbad = 0;
  RET(ff_alpha_from_b(ctx, c.W, ctx->ff_ones, o->lambda_a, c.alpha, &nbad));
  RET(ff_clamp_alpha(ctx, c.alpha));
  RET(ff_update_W(ctx, c.gmean, c.alpha, ctx->ff_ones, o->lambda_a, ctx->nw.W, &nbad));
  std::swap(c.W, ctx->nw.W);
  // clamp the columns of Wsub to median +- 4 mad over the REAL cells (:256-262): nsub is small, done on the host
  {
    std::vector<double> Wh((size_t)N * K);
    RET(ruvnb_get_state(ctx, "W", Wh.data()));
    for (int l = 0; l < K; ++l) {
      std::vector<double> col(Wh.begin() + (size_t)l * N, Wh.begin() + (size_t)(l + 1) * N), s(col);
      std::sort(s.begin(), s.end());
      auto med_of = [](std::vector<double>& v) { size_t n = v.size(); return (n & 1) ? v[n / 2] : 0.5 * (v[n / 2 - 1] + v[n / 2]); };
      double med = med_of(s);
      for (auto& x : s) x = fabs(x - med);
      std::sort(s.begin(), s.end());
      double mad = 1.4826 * med_of(s), hi = med + 4 * mad, lo = med - 4 * mad;
      for (long long i = 0; i < N; ++i) { double& x = Wh[(size_t)l * N + i]; if (x > hi) x = hi; if (x < lo) x = lo; }
    }
    RET(ruvnb_set_state(ctx, "W", Wh.data(), K));
  }
  k_ff_init_mb<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(ge, ctx->ff_Z, c.gmean, c.alpha, c.W, ctx->d_group_n, ctx->ff_ctlmask, c.Mb); CKL();
  k_fill<<<nblk(Np, 256), 256, 0, ctx->stream>>>(ctx->ff_step, Np, 1.0); CKL();
  CK(cudaStreamSynchronize(ctx->stream));
  c.sig_valid = false;
  return RUVNB_OK;
}

// per-cell log-likelihood of the current parameters (:334-348) -> "cell_loglik"; wt.ctl of the current parameters (:331);
// steps reset to 1 (:350)
int ruvnb_fast_begin_outer(ruvnb_ctx* ctx, const ruvnb_opts* o, double* total) {
  RET(check_ready(ctx, o));
  RET(ff_ensure(ctx));
  StateSet& c = ctx->cur;
  k_fill<<<nblk(ctx->Np, 256), 256, 0, ctx->stream>>>(ctx->ff_step, ctx->Np, 1.0); CKL();
  k_ff_working<<<nblk(ctx->Gl, 8), 256, 0, ctx->stream>>>(ffgeom(ctx), c.gmean, c.Mb, c.alpha, c.W, c.psi, ctx->ff_step, ctx->ff_Z, ctx->ff_SI, ctx->ff_ZS, ctx->ff_sirow); CKL();
  k_gather_rows<<<nblk(ctx->n_ctl, 256), 256, 0, ctx->stream>>>(ctx->ff_sirow, ctx->d_ctl_local, ctx->n_ctl, 1, c.wtctl); CKL();
  k_scale_vec<<<nblk(ctx->n_ctl, 256), 256, 0, ctx->stream>>>(c.wtctl, ctx->n_ctl, 1.0 / (double)ctx->N); CKL();
  return ff_cell_loglik(ctx, c, ctx->ff_cll, total);
}

// one pass of the inner-loop body (:357-531): cur -> candidate (swapped in), candidate per-cell log-likelihood in the tmp slot
int ruvnb_fast_iteration(ruvnb_ctx* ctx, const ruvnb_opts* o, double* total_new, int bad[3]) {
  RET(check_ready(ctx, o));
  RET(ff_ensure(ctx));
  StateSet& c = ctx->cur; StateSet& n = ctx->nw;
  const int Gl = ctx->Gl, K = ctx->K, m = ctx->m; const long long Np = ctx->Np, N = ctx->N;
  FFGeom ge = ffgeom(ctx);
  int nb_a = 0, nb_w = 0;
  k_ff_working<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(ge, c.gmean, c.Mb, c.alpha, c.W, c.psi, ctx->ff_step, ctx->ff_Z, ctx->ff_SI, ctx->ff_ZS, ctx->ff_sirow); CKL();
  RET(make_rmw(ctx, c.W));
  // wt.cell = colMeans(sig.inv) capped at its type-7 98 % quantile (:405-408)
  k_ff_colmean<<<nblk(Np, 128), 128, 0, ctx->stream>>>(ge, ctx->ff_SI, ctx->ff_wtcell); CKL();
  {
    std::vector<double> wc(Np), real; real.reserve(N);
    CK(cudaMemcpyAsync(wc.data(), ctx->ff_wtcell, (size_t)Np * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (long long p = 0; p < Np; ++p) if (ctx->pos2cell[p] >= 0) real.push_back(wc[p]);
    std::sort(real.begin(), real.end());
    double index = 1.0 + (double)(N - 1) * 0.98, lo = floor(index), hi = ceil(index);
    double q = real[(size_t)lo - 1];
    if (index > lo && real[(size_t)hi - 1] != q) { double h = index - lo; q = (1.0 - h) * q + h * real[(size_t)hi - 1]; }
    k_ff_cap<<<nblk(Np, 256), 256, 0, ctx->stream>>>(ctx->ff_wtcell, Np, q); CKL();
  }
  // alpha = RM_Z D RM_W (RM_W' D RM_W + lambda.a I)^-1 (:410-411); NaN -> revert (:418); clamp (:421-426)
  k_ff_b<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(ge, ctx->ff_Z, ctx->ff_ZS, ctx->d_group_n, ctx->ff_wtcell, ctx->RMW, ctx->ff_b); CKL();
  RET(ff_alpha_from_b(ctx, ctx->RMW, ctx->ff_wtcell, o->lambda_a, n.alpha, &nb_a));
  if (nb_a) CK(cudaMemcpyAsync(n.alpha, c.alpha, (size_t)Gl * K * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  RET(ff_clamp_alpha(ctx, n.alpha));
  // wt.ctl = rowMeans(sig.inv[ctl,]) (:445), Wsub (:447-457); non-finite -> keep the previous W (:475; the reference
  // names an undefined W.old there)
  k_gather_rows<<<nblk(ctx->n_ctl, 256), 256, 0, ctx->stream>>>(ctx->ff_sirow, ctx->d_ctl_local, ctx->n_ctl, 1, n.wtctl); CKL();
  k_scale_vec<<<nblk(ctx->n_ctl, 256), 256, 0, ctx->stream>>>(n.wtctl, ctx->n_ctl, 1.0 / (double)N); CKL();
  RET(ff_update_W(ctx, c.gmean, n.alpha, n.wtctl, o->lambda_a, n.W, &nb_w));
  if (nb_w) CK(cudaMemcpyAsync(n.W, c.W, (size_t)Np * K * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  // gmean (:478-480) and Mb (:485-505) from the per-group sums
  k_ff_gsums<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(ge, ctx->ff_Z, ctx->ff_SI, n.alpha, n.W, ctx->S0, ctx->S1); CKL();
  CK(cudaMemsetAsync(ctx->flags + 2, 0, 2 * sizeof(int), ctx->stream));
  k_gmean_mb<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, m, ctx->S0, ctx->S1, c.Mb, o->lambda_b, n.gmean, n.Mb, ctx->flags + 3, ctx->flags + 2); CKL();
  k_ff_zero_ctl_rows<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(Gl, m, ctx->ff_ctlmask, n.Mb); CKL();   // :490
  // NaN among the non-control rows -> whole Mb reverts; NA/Inf -> 0; row clamp (:496-505).  The NaN flag was raised before
  // the control rows were zeroed: recount on the zeroed matrix
  CK(cudaMemsetAsync(ctx->flags + 3, 0, sizeof(int), ctx->stream));
  k_count_nan<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(n.Mb, (long long)Gl * m, ctx->flags + 3); CKL();
  k_mb_finalize<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(Gl, m, ctx->flags + 3, c.Mb, n.Mb); CKL();
  CK(cudaMemcpyAsync(n.psi, c.psi, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(n.pi0, c.pi0, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(n.psi_w, c.psi_w, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  double tot = 0.0;
  RET(ff_cell_loglik(ctx, n, ctx->ff_cll_tmp, &tot));
  CK(cudaMemcpyAsync(ctx->h_flags + 2, ctx->flags + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (bad) { bad[0] = nb_a; bad[1] = nb_w; bad[2] = ctx->h_flags[2]; }
  if (total_new) *total_new = tot;
  n.sig_valid = false;
  std::swap(ctx->cur, ctx->nw);
  return RUVNB_OK;
}
// step[c] *= step.fac for the cells whose log-likelihood dropped (:538-539)
int ruvnb_fast_halve_steps(ruvnb_ctx* ctx, const ruvnb_opts* o) {
  if (!ctx || !o || !ctx->ff_ready) return RUVNB_EARG;
  CK(cudaSetDevice(ctx->device));
  k_ff_halve<<<nblk(ctx->Np, 256), 256, 0, ctx->stream>>>(ctx->Np, ctx->ff_cll, ctx->ff_cll_tmp, o->step_fac, ctx->ff_step); CKL();
  return RUVNB_OK;
}
// loglik <- loglik.tmp (:548)
int ruvnb_fast_accept(ruvnb_ctx* ctx) {
  if (!ctx || !ctx->ff_ready) return RUVNB_EARG;
  std::swap(ctx->ff_cll, ctx->ff_cll_tmp);
  return RUVNB_OK;
}

}  // extern "C"

// ---- H1: initial W (R/ruvIIInb.R:160-168) by subspace iteration on the resident control rows (svd.cuh) -----------------
extern "C" int ruvnb_init_W(ruvnb_ctx* ctx, int k, int max_iter, double tol, double* W_out, double* sv_out, int* iters_out) {
  if (!ctx) return RUVNB_EARG;
  if (!ctx->have_counts) return ctx->fail(RUVNB_ESTATE, "init_W: no counts uploaded");
  RET(ensure_state(ctx, k));
  CK(cudaSetDevice(ctx->device));
  const int K = ctx->K, nc = ctx->n_ctl, KK = K * (K + 1) / 2; const long long Np = ctx->Np;
  if (nc < K) return ctx->fail(RUVNB_EARG, "init_W: %d control genes cannot carry %d factors", nc, K);
  if (max_iter <= 0) max_iter = 100;
  if (!(tol > 0)) tol = 1e-9;   // on the Ritz values, i.e. ~3e-5 on the vectors: irlba's own tolerance is 1e-5
  double *rm = nullptr, *U = nullptr, *G = nullptr, *T = nullptr, *A = nullptr, *Upart = nullptr;
  // cell parts of the A V product: enough CTAs for ~4 per SM, whole 128-cell chunks per part
  const int row_blocks = (nc + 8 * SVD_RPW - 1) / (8 * SVD_RPW);
  int n_sm = 148;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
  int parts = std::max(1, std::min((int)(Np / CH), (4 * n_sm + row_blocks - 1) / row_blocks));
  const long long seg = ((Np / CH + parts - 1) / parts) * CH;
  parts = (int)((Np + seg - 1) / seg);
  ScopedDev scratch;
  CK(scratch.alloc(&rm, (size_t)nc * 8)); CK(scratch.alloc(&U, (size_t)nc * K * 8)); CK(scratch.alloc(&G, KK * 8)); CK(scratch.alloc(&T, K * K * 8));
  CK(scratch.alloc(&Upart, (size_t)parts * nc * K * 8));
  if (scratch.alloc(&A, (size_t)nc * Np * 8) != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(RUVNB_ENOMEM, "init_W: no room for the %d x %lld FP64 slab of log-ratios (%.1f GB)", nc, Np, (double)nc * Np * 8 / 1e9);
  }
  double* V = ctx->cur.W; double* V2 = ctx->nw.W;
  k_svd_rowmean<<<nblk(nc, 8), 256, 0, ctx->stream>>>(ctx->d_ctl_rows, nc, Np, (double)ctx->N, rm); CKL();
  k_svd_build<<<nblk((long long)nc * Np, 256), 256, 0, ctx->stream>>>(ctx->d_ctl_rows, nc, Np, rm, A); CKL();
  auto av = [&](const double* Vin) -> int {
    DISPATCH_K(K, { k_svd_av<KT><<<dim3(row_blocks, parts), 256, 0, ctx->stream>>>(A, nc, Np, seg, Vin, Upart); CKL(); });
    k_svd_sum_parts<<<nblk((long long)nc * K, 256), 256, 0, ctx->stream>>>(Upart, (long long)nc * K, parts, U); CKL();
    return RUVNB_OK;
  };
  k_svd_start<<<nblk(Np * K, 256), 256, 0, ctx->stream>>>(V, Np, K, ctx->d_pos2cell); CKL();
  std::vector<double> Gh(KK), Th(K * K), prev(K, 0.0), diag(K, 0.0);
  // orthonormalise the K vectors in `X` in place: Gram -> Cholesky G = R'R on the host -> X <- X R^-1
  auto orth = [&](double* X) -> int {
    k_svd_gram<<<KK, 1024, 0, ctx->stream>>>(X, Np, K, G); CKL();
    CK(cudaMemcpyAsync(Gh.data(), G, KK * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<double> R(K * K, 0.0);   // upper triangular, row-major
    for (int j = 0; j < K; ++j) {
      for (int i = 0; i <= j; ++i) {
        double s = Gh[j * (j + 1) / 2 + i];
        for (int q = 0; q < i; ++q) s -= R[q * K + i] * R[q * K + j];
        if (i == j) { if (!(s > 0.0)) return ctx->fail(RUVNB_EARG, "init_W: the control-gene matrix has rank < k"); R[j * K + j] = sqrt(s); diag[j] = s; }
        else R[i * K + j] = s / R[i * K + i];
      }
    }
    // Rinv (upper triangular)
    for (auto& x : Th) x = 0.0;
    for (int j = 0; j < K; ++j) {
      Th[j * K + j] = 1.0 / R[j * K + j];
      for (int i = j - 1; i >= 0; --i) {
        double s = 0.0;
        for (int q = i + 1; q <= j; ++q) s -= R[i * K + q] * Th[q * K + j];
        Th[i * K + j] = s / R[i * K + i];
      }
    }
    CK(cudaMemcpyAsync(T, Th.data(), K * K * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_svd_apply<<<nblk(Np, 256), 256, 0, ctx->stream>>>(X, Np, K, T); CKL();
    return RUVNB_OK;
  };
  RET(orth(V));
  int it = 0;
  for (; it < max_iter; ++it) {
    RET(av(V));
    DISPATCH_K(K, { k_svd_atu<KT><<<(unsigned)(Np / 32), 256, 0, ctx->stream>>>(A, nc, Np, U, V2); CKL(); });
    std::swap(V, V2);
    RET(orth(V));   // diag = squared norms of A'A v before normalisation ~ sigma^4 along converged directions
    double ch = 0.0;
    for (int l = 0; l < K; ++l) { ch = std::max(ch, fabs(diag[l] - prev[l]) / std::max(diag[l], 1e-300)); prev[l] = diag[l]; }
    if (it >= 2 && ch < tol) { ++it; break; }
  }
  // Rayleigh-Ritz: B = A V, eigen-decomposition of B'B -> singular values and the rotation of V
  RET(av(V));
  std::vector<double> Uh((size_t)nc * K), BtB(K * K, 0.0), Q;
  CK(cudaMemcpyAsync(Uh.data(), U, Uh.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int a = 0; a < K; ++a) for (int b = 0; b <= a; ++b) {
    double s = 0.0;
    for (int g = 0; g < nc; ++g) s += Uh[(size_t)g * K + a] * Uh[(size_t)g * K + b];
    BtB[a * K + b] = BtB[b * K + a] = s;
  }
  jacobi_eig_h(K, BtB, Q);
  CK(cudaMemcpyAsync(T, Q.data(), K * K * 8, cudaMemcpyHostToDevice, ctx->stream));
  k_svd_apply<<<nblk(Np, 256), 256, 0, ctx->stream>>>(V, Np, K, T); CKL();
  if (V != ctx->cur.W) { std::swap(ctx->cur.W, ctx->nw.W); }   // the result lives in cur.W
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_W = true; ctx->cur.sig_valid = false;
  // sign convention on the host copy, then written back so that "W" and W_out agree
  const long long N = ctx->N;
  std::vector<double> Wh((size_t)N * K);
  RET(ruvnb_get_state(ctx, "W", Wh.data()));
  for (int l = 0; l < K; ++l) {
    long long am = 0;
    for (long long c = 1; c < N; ++c) if (fabs(Wh[(size_t)l * N + c]) > fabs(Wh[(size_t)l * N + am])) am = c;
    if (Wh[(size_t)l * N + am] < 0) for (long long c = 0; c < N; ++c) Wh[(size_t)l * N + c] = -Wh[(size_t)l * N + c];
  }
  RET(ruvnb_set_state(ctx, "W", Wh.data(), K));
  if (W_out) memcpy(W_out, Wh.data(), Wh.size() * 8);
  if (sv_out) for (int l = 0; l < K; ++l) sv_out[l] = sqrt(std::max(BtB[l * K + l], 0.0));
  if (iters_out) *iters_out = it;
  return RUVNB_OK;   // scratch released by its scope
}
