// fast.cu -- fastruvIII.nb: full-data streaming passes (R/fastruvIIInb.R:632-748) and the subsample fit (:223-623)
#include "ctx.cuh"
#include "launch.cuh"
#include "fast.cuh"
#include "fastfit.cuh"

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
  if (ctx->nbatch > 1) return ctx->fail(RUVNB_ESTATE, "fastruvIII.nb has no batch-specific dispersion");
  if (ctx->nranks > 1 && !ctx->cell_shard) return ctx->fail(RUVNB_ESTATE, "fast_W_all: cells are independent -- shard them (ruvnb_set_cell_shard), not the genes");
  RET(ensure_state(ctx, k));
  CK(cudaSetDevice(ctx->device));
  RET(ruvnb_set_state(ctx, "W", W_init, k));
  const int K = ctx->K, nc = ctx->n_ctl, KK = K * (K + 1) / 2;
  StateSet& s = ctx->cur;
  double *d_gm = nullptr, *d_al = nullptr, *d_wt = nullptr, *d_A = nullptr, *d_med = nullptr, *d_mad = nullptr, *d_crit = nullptr;
  const unsigned grid = nblk(ctx->Np, 64);
  ScopedDev own;   // released on every way out
  CK(own.alloc(&d_gm, (size_t)nc * 8)); CK(own.alloc(&d_al, (size_t)nc * K * 8)); CK(own.alloc(&d_wt, (size_t)nc * 8));
  CK(own.alloc(&d_A, KK * 8)); CK(own.alloc(&d_med, K * 8)); CK(own.alloc(&d_mad, K * 8)); CK(own.alloc(&d_crit, (size_t)grid * 8));
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
    double n_all = (double)ctx->N;
    if (ctx->cell_shard && ctx->nranks > 1) {              // the criterion is a mean over the cells of ALL ranks
      CK(cudaMemcpyAsync(ctx->scal, &t, 8, cudaMemcpyHostToDevice, ctx->stream));
      RET(allreduce(ctx, ctx->scal, 1, NCCL_F64));
      CK(cudaMemcpyAsync(&t, ctx->scal, 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      n_all = (double)ctx->N_total;
    }
    crit = t / (n_all * K);                                // mean over all N x k entries (:709)
    if (it >= 1 && crit < tol) break;                      // the initial pass has no convergence test
  }
  s.sig_valid = false;
  if (n_sweeps) *n_sweeps = sweeps;
  if (crit_out) *crit_out = crit;
  if (W_out) RET(ruvnb_get_state(ctx, "W", W_out));
  return RUVNB_OK;
}

}  // extern "C"

// what F6 needs besides the state: annotation of every local cell, control flags, and the per-gene clamp limits
// median +- 4 mad of Mb.all over the ANNOTATED columns (R/fastruvIIInb.R:744-745) = of the values Mb[g, j] with multiplicity
// #annotated cells of group j (summed over the ranks of a cell-sharded problem).  Buffers land in `own` (freed by its scope).
int fast_mb_prepare(ruvnb_ctx* ctx, const int32_t* annot_grp, ScopedDev* own, FastMbPrep* p) {
  const int Gl = ctx->Gl, m = ctx->m; const long long N = ctx->N;
  StateSet& s = ctx->cur;
  std::vector<long long> cnt(m, 0);
  for (long long c = 0; c < N; ++c) { int j = annot_grp[c]; if (j >= m) return ctx->fail(RUVNB_EARG, "fast Mb: group %d outside 0..%d", j, m - 1); if (j >= 0) cnt[j]++; }
  if (ctx->cell_shard && ctx->nranks > 1) {
    long long* d_cnt = nullptr;
    CK(own->alloc(&d_cnt, (size_t)m * 8));
    CK(cudaMemcpyAsync(d_cnt, cnt.data(), (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    RET(allreduce(ctx, d_cnt, m, NCCL_INT64));
    CK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
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
  CK(own->alloc(&p->d_rmed, (size_t)Gl * 8)); CK(own->alloc(&p->d_rmad, (size_t)Gl * 8)); CK(own->alloc(&p->d_annot, (size_t)N * 4)); CK(own->alloc(&p->d_ctl, (size_t)Gl));
  CK(cudaMemcpyAsync(p->d_rmed, rmed.data(), (size_t)Gl * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(p->d_rmad, rmad.data(), (size_t)Gl * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(p->d_annot, annot_grp, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(p->d_ctl, ctl_local.data(), Gl, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope
  return RUVNB_OK;
}

extern "C" {

// F6: Mb.all (G x N, column-major, streamed to the host in blocks of block_size cells).  annot_grp[c] = replicate group
// of cell c or -1; the current state supplies alpha, gmean, psi, W and the subsample fit's Mb (G x m).
int ruvnb_fast_Mb_all(ruvnb_ctx* ctx, double lambda_b, const int32_t* annot_grp, int64_t block_size, double* Mb_all_out) {
  if (!ctx || !annot_grp || !Mb_all_out) return RUVNB_EARG;
  if (!ctx->have_counts || ctx->K == 0 || !ctx->have_W) return ctx->fail(RUVNB_ESTATE, "fast_Mb_all: needs counts and a state (W, alpha, gmean, Mb, psi)");
  if (ctx->nbatch > 1) return ctx->fail(RUVNB_ESTATE, "fastruvIII.nb has no batch-specific dispersion");
  CK(cudaSetDevice(ctx->device));
  const int Gl = ctx->Gl, m = ctx->m; const long long N = ctx->N;
  StateSet& s = ctx->cur;
  ScopedDev own; FastMbPrep pr;
  RET(fast_mb_prepare(ctx, annot_grp, &own, &pr));
  long long bs = std::min<long long>(block_size > 0 ? block_size : 5000, N);
  double* slab[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr; cudaEvent_t done[2], copied[2];
  for (int i = 0; i < 2; ++i) { CK(own.alloc(&slab[i], (size_t)Gl * bs * 8)); CK(cudaEventCreate(&done[i])); CK(cudaEventCreate(&copied[i])); }
  CK(cudaStreamCreate(&copy_stream));
  int blk = 0;
  for (long long c0 = 0; c0 < N; c0 += bs, ++blk) {
    const int B = (int)std::min<long long>(bs, N - c0), sl = blk & 1;
    if (blk >= 2) CK(cudaStreamWaitEvent(ctx->stream, copied[sl], 0));
    FastMbArgs a;
    a.Gl = Gl; a.K = ctx->K; a.m = m; a.B = B; a.Np = ctx->Np; a.c0 = c0; a.Yp = ctx->dY; a.alpha = s.alpha; a.gmean = s.gmean; a.psi = s.psi;
    a.W = s.W; a.Mb = s.Mb; a.cell2pos = ctx->d_cell2pos; a.annot = pr.d_annot; a.ctl_local = pr.d_ctl; a.rmed = pr.d_rmed; a.rmad = pr.d_rmad; a.lambda_b = lambda_b; a.gene_major = 0; a.ldB = 0;
    k_fast_mb<<<dim3(nblk(Gl, 32), nblk(B, 32)), 1024, 0, ctx->stream>>>(a, slab[sl]); CKL();
    CK(cudaEventRecord(done[sl], ctx->stream));
    CK(cudaStreamWaitEvent(copy_stream, done[sl], 0));
    CK(cudaMemcpyAsync(Mb_all_out + (size_t)c0 * Gl, slab[sl], (size_t)Gl * B * 8, cudaMemcpyDeviceToHost, copy_stream));
    CK(cudaEventRecord(copied[sl], copy_stream));
  }
  CK(cudaStreamSynchronize(ctx->stream)); CK(cudaStreamSynchronize(copy_stream));
  CK(cudaStreamDestroy(copy_stream));
  for (int i = 0; i < 2; ++i) { CK(cudaEventDestroy(done[i])); CK(cudaEventDestroy(copied[i])); }
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
  if (ctx->nbatch > 1) return ctx->fail(RUVNB_ESTATE, "this operator has no batch-specific dispersion: run it in one context per batch (see ruvnb_set_psi_callback)");
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
  if (ctx->nbatch > 1) return ctx->fail(RUVNB_ESTATE, "this operator has no batch-specific dispersion: run it in one context per batch (see ruvnb_set_psi_callback)");
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
