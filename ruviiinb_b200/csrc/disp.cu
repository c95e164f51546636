// disp.cu -- H3: the dispersion step (R/estimateDisp.par.R) and the per-gene Newton variant.
#include <stdlib.h>
#include <chrono>
#include "disp_launch.cuh"
#include "limma_host.h"

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
  RET(dev_alloc(ctx, &ctx->dp_ps1, (long long)DISP_MAXP * Gl * DISP_ND)); RET(dev_alloc(ctx, &ctx->dp_ps2, (long long)DISP_MAXP * Gl * DISP_ND));
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
    {
      // few rows (a small gene shard, the stragglers): cut the rows into parts so that the launch fills the device -- a sweep over
      // ONE row otherwise costs a whole row walk (1 ms at 100 000 cells)
      int P = 1;
      if (a.Ecache && !ctx->opt_no_disp_parts) {
        const long long ctas = (nlive + NWARP - 1) / NWARP;
        while (P < DISP_MAXP && ctas * P < 4 * 296 && ctx->nchunks / (2 * P) >= 16) P *= 2;
      }
      a.nparts = P; a.Gl = ctx->Gl; a.part_s1 = ctx->dp_ps1; a.part_s2 = ctx->dp_ps2;
      ctx->launch_parts = P;
      int r = launch_disp_nr(ctx, geom(ctx, nlive, live), s, a, nd);
      ctx->launch_parts = 1;
      RET(r);
    }
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
    RET(launch_disp_apl(ctx, geom(ctx, Gl, nullptr), s, a, ctx->dp_tabG + (long long)c * YT, ctx->dp_lgr_grid + c));
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
  if (ctx->nbatch > 1) return ctx->fail(RUVNB_ESTATE, "this operator has no batch-specific dispersion: run it in one context per batch (see ruvnb_set_psi_callback)");
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
// robust: limma's squeezeVar(robust = TRUE) prior (one prior d.f. per gene, outlier genes shrink less) as every call site of the
// reference asks (R/ruvIIInb.R:210,540, R/fastruvIIInb.R:282,569); 0: the plain fitFDist prior with the AveLogCPM trend.
// prior_df_in >= 0 injects one prior d.f. for all genes.  prior_df_vec_out: local rows (Inf outside the selected genes when robust).
static int estimate_disp_impl(ruvnb_ctx* ctx, const ruvnb_opts* o, int robust, double prior_df_in, double* common_out, double* trended_out,
                              double* prior_df_out, double* prior_df_vec_out) {
  RET(check_ready(ctx, o));
  if (ctx->nbatch > 1) return ctx->fail(RUVNB_ESTATE, "this operator has no batch-specific dispersion: run it in one context per batch (see ruvnb_set_psi_callback)");
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
  int i_min = 0;
  for (int i = 1; i < ns; ++i) if (alc[sel[i]] < alc[sel[i_min]]) i_min = i;   // which.min(AveLogCPM[sel])
  // a gene shard needs the trend (and everything after it) for its OWN genes only, plus the gene of lowest abundance whose trended
  // value the unselected genes take (:143-145): `mine` = their sorted positions (all of them on one GPU)
  const bool sharded = ctx->nranks > 1;
  auto is_mine = [&](int gene) { return !sharded || (gene >= g0 && gene < g0 + Gl) || gene == sel[i_min]; };
  std::vector<int> mine;
  for (int i = 0; i < ns; ++i) if (is_mine(perm[i])) mine.push_back(i);
  {
    double *d_xs = ctx->dp_xs, *d_m0 = ctx->dp_m0, *d_l0g = nullptr; int* d_perm = ctx->dp_perm;   // ns <= G
    if (sharded) d_l0g = tmpA; else d_l0g = ctx->dp_l0;
    CK(cudaMemcpyAsync(d_xs, xs.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_perm, perm.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, ctx->stream));
    int* d_mine = nullptr;
    if (sharded) {
      CK(scratch.alloc(&d_mine, std::max<size_t>(1, mine.size()) * 4));
      CK(cudaMemcpyAsync(d_mine, mine.data(), mine.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!mine.empty()) { k_tricube<<<(unsigned)mine.size(), 128, 0, ctx->stream>>>(d_xs, d_perm, ns, q, d_l0g, DISP_NG, d_m0, d_mine); CKL(); }
    CK(cudaMemcpyAsync(m0s.data(), d_m0, m0s.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));   // (rows of other ranks' genes: not evaluated, not read)
    CK(cudaStreamSynchronize(ctx->stream));
  }
  // m0 back to gene order; trended dispersion (:140-145)
  std::vector<double> m0((size_t)G * DISP_NG, 0.0), trended(G, 0.0);
  for (int i : mine) {
    memcpy(&m0[(size_t)perm[i] * DISP_NG], &m0s[(size_t)i * DISP_NG], DISP_NG * 8);
    trended[perm[i]] = 0.1 * exp2(maximize_interpolant_h(DISP_NG, pts, &m0s[(size_t)i * DISP_NG]));
  }
  { std::vector<char> issel(G, 0); for (int i : sel) issel[i] = 1; for (long long g = 0; g < G; ++g) if (!issel[g]) trended[g] = trended[sel[i_min]]; }
  if (trended_out) memcpy(trended_out, trended.data() + g0, (size_t)Gl * 8);
  const double t_trend = now();
  // prior d.f. (:156-169): deviance at the trended dispersion -> limma::squeezeVar(s2, df = N - 1, covariate = AveLogCPM[sel], robust)
  std::vector<double> prior_df_sel(ns, prior_df_in);   // one value per selected gene
  if (!(prior_df_in >= 0.0)) {
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
    std::vector<double> s2(ns), cov(ns);
    for (int i = 0; i < ns; ++i) { s2[i] = std::max(dev[sel[i]] / (N - 1.0), 0.0); cov[i] = alc[sel[i]]; }   // :163-165 (df.residual = N - 1)
    prior_df_sel = limma_host::squeeze_var_df_prior(s2, N - 1.0, &cov, robust != 0);                          // :166-168
  }
  if (prior_df_out) *prior_df_out = prior_df_sel.empty() ? NAN : prior_df_sel[0];
  const double t_prior = now();
  // tagwise (:170-190): per gene, maximise the spline through l0 + prior.n m0 with prior.n = prior.df / (N - 1) (ncoefs = 1), prior.n
  // capped at 1e6; robust: genes whose prior.n exceeds 1e6 keep the trended value (:186-189)
  std::vector<double> tag(trended);
  {
    bool all_large = true;
    for (int i = 0; i < ns; ++i) all_large &= (prior_df_sel[i] / (N - 1.0) > 1e6);
    if (!all_large) {
      double l0a[DISP_NG];
      for (int q = 0; q < ns; ++q) {
        const int i = sel[q];
        if (sharded && (i < g0 || i >= g0 + Gl)) continue;   // another rank's gene
        double pn = prior_df_sel[q] / (N - 1.0);
        const bool large = pn > 1e6;
        if (large) { if (robust) continue; pn = 1e6; }
        for (int dd = 0; dd < DISP_NG; ++dd) l0a[dd] = l0[(size_t)i * DISP_NG + dd] + pn * m0[(size_t)i * DISP_NG + dd];
        tag[i] = 0.1 * exp2(maximize_interpolant_h(DISP_NG, pts, l0a));
      }
    }
  }
  if (prior_df_vec_out) {   // :192-198: Inf outside the selected genes when robust, else the one value
    std::vector<double> full(G, robust ? INFINITY : (prior_df_sel.empty() ? NAN : prior_df_sel[0]));
    for (int q = 0; q < ns; ++q) full[sel[q]] = prior_df_sel[q];
    memcpy(prior_df_vec_out, full.data() + g0, (size_t)Gl * 8);
  }
  // psi <- tagwise where finite (R/ruvIIInb.R:564)
  std::vector<double> old(Gl);
  CK(cudaMemcpyAsync(old.data(), s.psi, (size_t)Gl * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int g = 0; g < Gl; ++g) if (std::isfinite(tag[g0 + g])) old[g] = tag[g0 + g];
  CK(cudaMemcpyAsync(s.psi, old.data(), (size_t)Gl * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->psi_ver++; s.sig_valid = false; ctx->best_eq_cur = false;
  if (o->verbose)
    fprintf(stderr, "estimateDisp: grid APL %.3f s (%d NR sweeps + 3 APL sweeps), gather + common %.3f s, AveLogCPM %.3f s (%d sweeps), "
                    "trend %.3f s, prior d.f. %.3f s (%d sweeps + deviance), tagwise %.3f s\n",
            t_grid - t_start, sweeps_grid, t_common - t_grid, t_alc - t_common, sweeps_alc - sweeps_grid, t_trend - t_alc,
            t_prior - t_trend, ctx->disp_nr_sweeps - sweeps_alc, now() - t_prior);
  return RUVNB_OK;
}

int ruvnb_estimate_disp_ex(ruvnb_ctx* ctx, const ruvnb_opts* o, double prior_df_in, double* common_out, double* trended_out,
                           double* prior_df_out) {
  return estimate_disp_impl(ctx, o, 0, prior_df_in, common_out, trended_out, prior_df_out, nullptr);
}
int ruvnb_estimate_disp_r(ruvnb_ctx* ctx, const ruvnb_opts* o, int robust, double prior_df_in, double* common_out, double* trended_out,
                          double* prior_df_vec_out) {
  return estimate_disp_impl(ctx, o, robust, prior_df_in, common_out, trended_out, nullptr, prior_df_vec_out);
}
// what the fit calls at every psi refresh: robust = TRUE, as R/ruvIIInb.R:210,540 pass it
int ruvnb_estimate_disp(ruvnb_ctx* ctx, const ruvnb_opts* o, double* common_out) {
  return estimate_disp_impl(ctx, o, 1, -1.0, common_out, nullptr, nullptr, nullptr);
}

}  // extern "C"
// the north_star variant of the dispersion step: per-gene Newton on log psi at the current fitted means (disp.cuh)
extern "C" int ruvnb_disp_newton(ruvnb_ctx* ctx, const ruvnb_opts* o, double* psi_out, int maxit) {
  RET(check_ready(ctx, o));
  if (ctx->nbatch > 1) return ctx->fail(RUVNB_ESTATE, "this operator has no batch-specific dispersion: run it in one context per batch (see ruvnb_set_psi_callback)");
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

// host-only: the prior-d.f. step (limma_host.h) for the CPU test-suite
extern "C" int ruvnb_selftest_squeeze_var(const double* s2, int64_t n, double df1, const double* covariate, int robust, double* df_prior_out) {
  if (!s2 || !df_prior_out || n < 0 || !(df1 > 0.0)) return RUVNB_EARG;
  std::vector<double> x(s2, s2 + n), cov;
  if (covariate) cov.assign(covariate, covariate + n);
  const std::vector<double> r = limma_host::squeeze_var_df_prior(x, df1, covariate ? &cov : nullptr, robust != 0);
  for (int64_t i = 0; i < n; ++i) df_prior_out[i] = r[i];
  return RUVNB_OK;
}
