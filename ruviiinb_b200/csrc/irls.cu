// irls.cu -- the inner-loop body of ruvIII.nb (R/ruvIIInb.R:303-527) and the whole-fit driver (:266-609).
#include <stdlib.h>
#include "sweeps.cuh"
#include "zinb.cuh"

// the fused log-likelihood + Gram sweep applies to plain NB fits with the default Huber constant (robust = TRUE makes
// every Huber weight live, ZINB needs the general instantiation) when the two slabs fit in HBM: see kernels.cuh k_ll_w / k_gram
static int ensure_wcache(ruvnb_ctx* ctx);
static bool fuse_ok(ruvnb_ctx* ctx, const ruvnb_opts* o) {
  if (!(o->huber_fastpath && !o->robust && !ctx->zinb_on && !ctx->opt_no_fuse)) return false;
  if (ensure_wcache(ctx) != RUVNB_OK) return false;
  return ctx->Wc && ctx->WZc;
}
// [Sum l | bad_Mb | nan_Mb] of this rank into scal[0..2] (one all-reduce carries the three)
static __global__ void k_pack_fin(double* scal, const int* flags) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { scal[1] = (double)flags[2]; scal[2] = (double)flags[3]; }
}

// hints for a context whose rows have none (k_hint_sample): 16 whole chunks spread over the non-empty chunks of the row
static int seed_hints(ruvnb_ctx* ctx, const ruvnb_opts* o, StateSet& s) {
  ctx->hints_seeded = true;
  std::vector<int> full;
  for (int c = 0; c < ctx->nchunks; ++c) if (ctx->chunk_valid[c] > 0) full.push_back(c);
  const int ns = 16;   // 2048 cells: the sample median is off by ~0.03 MAD, the window is +- 0.7 of the slack
  if ((int)full.size() < 4 * ns) return RUVNB_OK;   // small problems: the exact path is cheap
  std::vector<int> pick(ns);
  long long n_s = 0;
  for (int i = 0; i < ns; ++i) { pick[i] = full[(size_t)((2 * i + 1) * full.size() / (2 * ns))]; n_s += ctx->chunk_valid[pick[i]]; }
  const int Gl = ctx->Gl;
  ScopedDev own;
  int *d_pick = nullptr, *rn = nullptr; double *Ds = nullptr, *rmx = nullptr, *sg = nullptr;
  if (own.alloc(&Ds, (size_t)Gl * ns * CH * 8) != cudaSuccess) { cudaGetLastError(); return RUVNB_OK; }   // no room: no hints, the exact path runs
  CK(own.alloc(&d_pick, ns * 4)); CK(own.alloc(&rn, (size_t)Gl * 4)); CK(own.alloc(&rmx, (size_t)Gl * 8)); CK(own.alloc(&sg, (size_t)Gl * 8));
  CK(cudaMemcpyAsync(d_pick, pick.data(), ns * 4, cudaMemcpyHostToDevice, ctx->stream));
  k_hint_sample<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(Gl, ctx->K, ctx->m, ctx->Np, ctx->dY, s.gmean, s.Mb, s.alpha, s.psi, s.W, d_pick, ns, ctx->d_chunk_grp,
                                                      ctx->d_chunk_valid, ctx->d_chunk_batch, Ds, rmx, rn); CKL();
  const double kh = o->robust ? 1.345 : 100.0;
  k_row_sigma<<<Gl, RS_NT, 0, ctx->stream>>>(Ds, (long long)ns * CH, n_s, nullptr, rmx, rn, kh, 1, sg, nullptr, ctx->hint, ctx->opt_hint_half); CKL();
  CK(cudaStreamSynchronize(ctx->stream));   // (scratch leaves with this scope)
  return RUVNB_OK;
}

// per-gene log-likelihood of parameter set `s` into `out` (R/ruvIIInb.R:276-291) and, in the same sweep, the
// Huber certificate of `s` (s.sigma: +Inf certified / 0 to be evaluated exactly; the rows left at 0 are listed in
// ctx->faillist).  with_p1: the sweep also leaves the working quantities of `s` in the slabs and k_gram takes the Gram / score /
// group sums of the next inner iteration from them (k_ll_w + k_gram).
// total != nullptr: the sum over all genes of all ranks is brought to the host (ONE stream synchronisation, which also
// brings the number of listed rows and, in fin[0..1], the global counts of problematic / NaN Mb entries).
static int loglik_of(ruvnb_ctx* ctx, const ruvnb_opts* o, StateSet& s, double* out, double* total, bool with_p1 = false, double* fin = nullptr) {
  const double kh = o->robust ? 1.345 : 100.0;  // :155
  const int Gl = ctx->Gl, P = ctx->nparts;
  RET(ensure_ytabs(ctx, s.psi));
  if (o->huber_fastpath && !ctx->hints_seeded && !ctx->opt_no_seed) RET(seed_hints(ctx, o, s));
  if (with_p1) RET(make_rmw(ctx, s.W));   // (fuse_ok has allocated the slabs)
  KEV_BEGIN(4);
  ctx->launch_parts = P;
  int r;
  if (with_p1) {
    if (P > 1) {  // groups that a part never meets must read as zero
      CK(cudaMemsetAsync(ctx->ZS, 0, (size_t)P * Gl * ctx->m * 8, ctx->stream));
      CK(cudaMemsetAsync(ctx->VS, 0, (size_t)P * Gl * ctx->m * ctx->K * 8, ctx->stream));
      CK(cudaMemsetAsync(ctx->SWZ, 0, (size_t)P * Gl * ctx->m * 8, ctx->stream));
      CK(cudaMemsetAsync(ctx->S0c, 0, (size_t)P * Gl * ctx->m * 8, ctx->stream));
    }
    r = launch_ll_w(ctx, geom(ctx, Gl, nullptr), s, P > 1 ? ctx->llparts : out, o->huber_fastpath);
  } else {
    r = launch_loglik(ctx, geom(ctx, Gl, nullptr), s, P > 1 ? ctx->llparts : out, o->huber_fastpath);
  }
  ctx->launch_parts = 1;
  RET(r);
  if (P > 1) { k_sum_parts<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(out, ctx->llparts, Gl, P); CKL(); }
  k_cert_finish<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, P, ctx->certc, ctx->hint, s.sigma, kh, o->huber_fastpath, (long long)ctx->N, ctx->cert_diag);
  CKL();
  KEV_END(4);
  if (with_p1) {
    KEV_BEGIN(6);
    ctx->launch_parts = P;
    r = launch_gram(ctx, geom(ctx, Gl, nullptr));
    ctx->launch_parts = 1;
    RET(r);
    KEV_END(6);
  }
  KEV_BEGIN(10);
  s.sig_valid = true; s.maybe_active = false; s.p1_valid = with_p1; s.nfail = -1;
  if (with_p1) { ctx->cur.p1_valid = ctx->nw.p1_valid = ctx->best.p1_valid = false; s.p1_valid = true; }   // one set of sums
  k_list_failed<<<1, 1024, 0, ctx->stream>>>(Gl, s.sigma, ctx->faillist, ctx->flags + 4, nullptr);
  CKL();
  ctx->faillist_for = s.sigma;
  if (total) {
    k_colsum_part<<<COLSUM_B, 256, 0, ctx->stream>>>(out, Gl, 1, ctx->colsum_part); CKL();
    k_colsum_fin<<<1, 64, 0, ctx->stream>>>(ctx->colsum_part, 1, ctx->scal);
    CKL();
    k_pack_fin<<<1, 32, 0, ctx->stream>>>(ctx->scal, ctx->flags); CKL();
    RET(allreduce(ctx, ctx->scal, 3, NCCL_F64));
    CK(cudaMemcpyAsync(ctx->h_scal, ctx->scal, 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_flags, ctx->flags, 5 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_flags + 6, ctx->w_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KEV_END(10);
    CK(cudaStreamSynchronize(ctx->stream));
    kev_collect(ctx);
    *total = ctx->h_scal[0];
    s.nfail = ctx->h_flags[4];
    if (fin) { fin[0] = ctx->h_scal[1]; fin[1] = ctx->h_scal[2]; }
  }
  return RUVNB_OK;
}

// complete s.sigma: rows the certificate sweep left at 0 get the exact mad/0.6745 (R/ruvIIInb.R:713) -- or +Inf
// when that exact value proves every weight is 1.  *nfail = rows evaluated exactly, *nactive = rows whose
// Huber weights are live (finite sigma) afterwards.  No host synchronisation when the row count came back with the
// log-likelihood total of the sweep that wrote s.sigma.
static int complete_sigma(ruvnb_ctx* ctx, const ruvnb_opts* o, StateSet& s, double kh, int* nfail, int* nactive) {
  const int Gl = ctx->Gl;
  if (s.nfail < 0 || (s.nfail > 0 && ctx->faillist_for != s.sigma)) {
    k_list_failed<<<1, 1024, 0, ctx->stream>>>(Gl, s.sigma, ctx->faillist, ctx->flags + 4, nullptr);
    CKL();
    ctx->faillist_for = s.sigma;
    CK(cudaMemcpyAsync(ctx->h_flags + 4, ctx->flags + 4, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    kev_collect(ctx);
    s.nfail = ctx->h_flags[4];
  }
  const int nf = s.nfail;
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
      int r = launch_signdev(ctx, geom(ctx, nr, ctx->faillist + r0), s);
      ctx->launch_parts = 1;
      RET(r);
      k_row_sigma<<<nr, RS_NT, 0, ctx->stream>>>(ctx->D, ctx->Np, (long long)ctx->N, ctx->faillist + r0, ctx->rowmax, ctx->rownan, kh,
                                                 o->huber_fastpath, s.sigma, ctx->sigma_exact, ctx->hint, ctx->opt_hint_half);
      CKL();
    }
    KEV_END(5);
    s.maybe_active = true;  // finite sigmas can only come from the exact path
    s.nfail = 0;            // sigma is complete now
  }
  *nactive = s.maybe_active ? 1 : 0;
  return RUVNB_OK;
}

// RM_W = W - groupmean(W) (R/ruvIIInb.R:331-340)
int make_rmw(ruvnb_ctx* ctx, const double* W) {
  const int K = ctx->K, m = ctx->m;
  k_wbar<<<nblk((long long)K * m * 32, 256), 256, 0, ctx->stream>>>(W, ctx->Np, K, m, ctx->d_group_chunk0, ctx->d_group_n, ctx->d_group_nch, ctx->d_chunk_valid, ctx->Wbar);
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
// Runs `fn(geometry, huber)` for the certified rows (lean instantiation, main stream) and, when some rows carry live Huber
// weights, for those rows FIRST on the side stream: a row is one warp walking all cells, so a handful of them after the
// bulk would add a whole row-walk (2.5 ms at cfg 2) to every iteration.  With ZINB the general instantiation runs on all rows.
template <class Fn>
static int split_rows(ruvnb_ctx* ctx, bool hub, Fn fn) {
  const int Gl = ctx->Gl;
  if (ctx->zinb_on) return fn(geom(ctx, Gl, nullptr), true);
  if (hub) {
    RowGeom gf = geom(ctx, Gl, ctx->list_fin); gf.nrows_dev = ctx->part_counts + 1;
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream)); CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    cudaStream_t main_ = ctx->stream; ctx->stream = ctx->stream2;
    int r = fn(gf, true);
    ctx->stream = main_;
    RET(r);
    CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
  }
  RowGeom gi = geom(ctx, Gl, ctx->list_inf); gi.nrows_dev = ctx->part_counts;
  RET(fn(gi, false));
  if (hub) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  return RUVNB_OK;
}

// control-gene parameters of the CURRENT state gathered into ctx->ctlpack: [gmean | psi (nb columns) | step | alpha_old (K) |
// alpha_new (K, filled later) | Mb (m) | pi0 | psi_w (nb columns)] x n_ctl; genes of other ranks read 0 (the pack is then summed over
// the ranks).  nb = batches (1 unless the dispersion is batch-specific); Gl = stride between the columns of psi / psi_w.
static __global__ void k_gather_ctlpack(int nc, int K, int m, int nb, long long Gl, const int* ctl_local, const double* gmean, const double* psi, const double* step,
                                        const double* alpha, const double* Mb, const double* pi0, const double* psi_w, double* pk) {
  const int per = 3 + 2 * nb + 2 * K + m;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nc * per) return;
  // field-major layout (as CtlPack expects): locate the field of element i
  double* p_psi = pk + nc; double* p_step = p_psi + (long long)nb * nc;
  double* p_aold = p_step + nc; double* p_anew = p_aold + (long long)nc * K; double* p_mb = p_anew + (long long)nc * K;
  double* p_pi0 = p_mb + (long long)nc * m; double* p_psiw = p_pi0 + nc;
  double* dst = pk + i;
  double v = 0.0;
  if (dst < p_psi) { int g = ctl_local[i]; if (g >= 0) v = gmean[g]; }
  else if (dst < p_step) { long long q = dst - p_psi; int g = ctl_local[q % nc]; if (g >= 0) v = psi[(q / nc) * Gl + g]; }
  else if (dst < p_aold) { int g = ctl_local[dst - p_step]; if (g >= 0) v = step[g]; }
  else if (dst < p_anew) { long long q = dst - p_aold; int g = ctl_local[q / K]; if (g >= 0) v = alpha[(long long)g * K + q % K]; }
  else if (dst < p_mb) { v = 0.0; }
  else if (dst < p_pi0) { long long q = dst - p_mb; int g = ctl_local[q / m]; if (g >= 0) v = Mb[(long long)g * m + q % m]; }
  else if (dst < p_psiw) { int g = ctl_local[dst - p_pi0]; if (g >= 0) v = pi0[g]; }
  else { long long q = dst - p_psiw; int g = ctl_local[q % nc]; if (g >= 0) v = psi_w[(q / nc) * Gl + g]; }
  *dst = v;
}
// this rank's slice of the all-gather buffer: [new alpha rows | current alpha rows] (each zero-padded to `per` rows), then
// its non-finite count.  The current rows ride along because a NaN on ANY rank reverts the whole matrix (:368).
static __global__ void k_pack_alpha(const double* a_new, const double* a_cur, long long nloc, long long per_k, const int* bad, double* slice) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > 2 * per_k) return;
  if (i == 2 * per_k) { slice[i] = (double)*bad; return; }
  const long long q = i < per_k ? i : i - per_k;
  slice[i] = q < nloc ? (i < per_k ? a_new[q] : a_cur[q]) : 0.0;
}
// gathered slices -> alpha_full [G][K] (the new matrix, or the current one when any rank counted a non-finite entry);
// *bad = sum of the ranks' counts
static __global__ void k_unpack_alpha(const double* gath, int nranks, long long per_k, long long GK, double* alpha_full, int* bad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long sl = 2 * per_k + 1;
  double t = 0.0;
  for (int r = 0; r < nranks; ++r) t += gath[(long long)r * sl + 2 * per_k];
  if (i < GK) alpha_full[i] = gath[(i / per_k) * sl + (t > 0.0 ? per_k : 0) + i % per_k];
  if (i == 0) *bad = (int)t;
}
// the zero-padded sum path (uneven shards): buf = [new G K | current G K | count]
static __global__ void k_sumpath_pack(double* buf, long long GK, long long off, long long nloc, const double* a_new, const double* a_cur, const int* bad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nloc) { buf[off + i] = a_new[i]; buf[GK + off + i] = a_cur[i]; }
  if (i == 0) buf[2 * GK] = (double)*bad;
}
static __global__ void k_sumpath_unpack(const double* buf, long long GK, double* alpha_full, int* bad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const double t = buf[2 * GK];
  if (i < GK) alpha_full[i] = buf[(t > 0.0 ? GK : 0) + i];
  if (i == 0) *bad = (int)t;
}
// all parts of the five pass-1 buffers summed in one launch (k_sum_parts five times)
static __global__ void k_sum_parts5(double* A, long long nA, double* u, long long nu, double* sw, long long nsw, double* ZS, long long nZS, double* VS, long long nVS, int P) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double* p; long long n;
  if (i < nA) { p = A; n = nA; }
  else if ((i -= nA) < nu) { p = u; n = nu; }
  else if ((i -= nu) < nsw) { p = sw; n = nsw; }
  else if ((i -= nsw) < nZS) { p = ZS; n = nZS; }
  else if ((i -= nZS) < nVS) { p = VS; n = nVS; }
  else return;
  double s = p[i];
  for (int q = 1; q < P; ++q) s += p[(long long)q * n + i];
  p[i] = s;
}

// the weight cache (ctx->Wc): 8 bytes per element, taken once when free HBM covers it plus a reserve for what is allocated later
// (exact-MAD scratch, the dispersion step's exp(offset) slab, get.res slabs)
static int ensure_wcache(ruvnb_ctx* ctx) {
  if (ctx->Wc_tried) return RUVNB_OK;
  ctx->Wc_tried = true;
  const char* e = getenv("RUVNB_NO_WCACHE");
  if (e && *e && *e != '0') return RUVNB_OK;
  size_t fr = 0, tot = 0;
  CK(cudaMemGetInfo(&fr, &tot));
  const size_t need = (size_t)ctx->Gl * (size_t)ctx->Np * 8, reserve = std::max<size_t>((size_t)24 << 30, tot / 4) + need;   // + the exp(offset) slab
  if (fr > need + reserve) {
    void* q = nullptr;
    if (cudaMalloc(&q, need) == cudaSuccess) {
      ctx->allocs.push_back(q); ctx->Wc = (double*)q;
      RET(dev_alloc(ctx, &ctx->SWZ, (long long)ctx->nparts * ctx->Gl * ctx->m));
      RET(dev_alloc(ctx, &ctx->S0c, (long long)ctx->nparts * ctx->Gl * ctx->m));
      // the second slab (w Z) of the fused sweep pair
      if (fr > 2 * need + reserve && cudaMalloc(&q, need) == cudaSuccess) { ctx->allocs.push_back(q); ctx->WZc = (double*)q; }
      else cudaGetLastError();
    } else cudaGetLastError();
  }
  return RUVNB_OK;
}

static int irls_body(ruvnb_ctx* ctx, const ruvnb_opts* o, double* logl_new, int bad[3]) {
  const int Gl = ctx->Gl, K = ctx->K, m = ctx->m, KK = K * (K + 1) / 2;
  const double kh = o->robust ? 1.345 : 100.0;  // :155
  StateSet& c = ctx->cur; StateSet& n = ctx->nw;
  const bool fuse = fuse_ok(ctx, o);
  CK(cudaMemsetAsync(ctx->flags, 0, 4 * sizeof(int), ctx->stream));
  RET(ensure_ytabs(ctx, c.psi));
  RET(ensure_wcache(ctx));
  // Huber scale of the current state (:713,723,409): normally left by the log-likelihood sweep that scored it
  if (!c.sig_valid) RET(loglik_of(ctx, o, c, ctx->loglik, nullptr, fuse));
  int nfail = 0, nactive = 0;
  RET(complete_sigma(ctx, o, c, kh, &nfail, &nactive));
  ctx->last_fail_rows = nfail; ctx->last_active_rows = nactive;
  if (nfail > 0 && ctx->best_eq_cur) {
    // `best` holds this very parameter set (snapshot or restore just before this pass): it keeps the completed Huber scales, so a
    // rejected step that restores it does not send the same rows through the exact path again
    CK(cudaMemcpyAsync(ctx->best.sigma, c.sigma, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->best.sig_valid = c.sig_valid; ctx->best.maybe_active = c.maybe_active; ctx->best.nfail = c.nfail;
  }
  ctx->best_eq_cur = false;   // the pass ends with a new current state
  const bool hub = nactive > 0 || ctx->zinb_on;   // the general instantiation also carries the ZINB weights
  k_partition_sigma<<<1, 1024, 0, ctx->stream>>>(Gl, c.sigma, ctx->roworder_ready ? ctx->d_roworder : nullptr, ctx->list_inf, ctx->list_fin, ctx->part_counts);
  CKL();
  const int P = ctx->nparts;
  // control-gene parameters of the current state for the W step: gathered (and, when sharded, summed over the ranks) on the
  // side stream while pass 1 runs -- nothing of it depends on this iteration's work except alpha_new, filled in later
  const int nc = ctx->n_ctl;
  double* pk = ctx->ctlpack;
  const int nb = ctx->nbatch;
  const long long per = 3 + 2 * nb + 2 * K + m;
  double* p_gmean = pk; double* p_psi = pk + nc; double* p_step = p_psi + (long long)nb * nc;
  double* p_aold = p_step + nc; double* p_anew = p_aold + (long long)nc * K; double* p_mb = p_anew + (long long)nc * K;
  double* p_pi0 = p_mb + (long long)nc * m; double* p_psiw = p_pi0 + nc;
  {
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream)); CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    k_gather_ctlpack<<<nblk((long long)nc * per, 256), 256, 0, ctx->stream2>>>(nc, K, m, nb, Gl, ctx->d_ctl_local, c.gmean, c.psi, ctx->step, c.alpha, c.Mb, c.pi0, c.psi_w, pk);
    CKL();
    if (ctx->nranks > 1) {
      int r = g_nccl.AllReduce(pk, pk, (size_t)nc * per, NCCL_F64, NCCL_SUM, ctx->comm, ctx->stream2);
      if (r != 0) return ctx->fail(RUVNB_ENCCL, "ncclAllReduce (control-gene pack): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    }
    CK(cudaEventRecord(ctx->ev_ctl, ctx->stream2));
  }
  // alpha ------------------------------------------------------------------------------------------
  {
    KEV_BEGIN(1);
    ctx->launch_parts = P;
    int r = RUVNB_OK;
    if (!c.p1_valid) {
      RET(make_rmw(ctx, c.W));
      if (P > 1) {  // groups that a part never meets must read as zero
        CK(cudaMemsetAsync(ctx->ZS, 0, (size_t)P * Gl * m * 8, ctx->stream));
        CK(cudaMemsetAsync(ctx->VS, 0, (size_t)P * Gl * m * K * 8, ctx->stream));
        if (ctx->Wc) {
          CK(cudaMemsetAsync(ctx->SWZ, 0, (size_t)P * Gl * m * 8, ctx->stream));
          CK(cudaMemsetAsync(ctx->S0c, 0, (size_t)P * Gl * m * 8, ctx->stream));
        }
      }
      // certified rows (sigma = +Inf: every Huber weight is 1) take the lean instantiation, the others the general one;
      // the two lists and their lengths stay on the device (no host round trip), idle CTAs exit at once
      r = split_rows(ctx, hub, [&](const RowGeom& g, bool huber) { return launch_pass1(ctx, huber, g, c, kh); });
    } else if (hub) {
      // the sums came with the sweep pair that scored this state (k_ll_w + k_gram, unit Huber weights; ctx->RMW is that of c.W);
      // only the rows whose weights turned out to be live are redone
      RowGeom gf = geom(ctx, Gl, ctx->list_fin); gf.nrows_dev = ctx->part_counts + 1;
      r = launch_pass1(ctx, true, gf, c, kh);
    }
    ctx->launch_parts = 1;
    RET(r);
    c.p1_valid = false;   // consumed: the sums are normalised in place below
    if (P > 1) {
      const long long nA = (long long)Gl * KK, nu = (long long)Gl * K, nZ = (long long)Gl * m, nV = (long long)Gl * m * K;
      k_sum_parts5<<<nblk(nA + nu + Gl + nZ + nV, 256), 256, 0, ctx->stream>>>(ctx->A, nA, ctx->u, nu, ctx->sw, Gl, ctx->ZS, nZ, ctx->VS, nV, P); CKL();
      if (ctx->Wc) { k_sum_parts5<<<nblk(2 * nZ, 256), 256, 0, ctx->stream>>>(ctx->SWZ, nZ, ctx->S0c, nZ, nullptr, 0, nullptr, 0, nullptr, 0, P); CKL(); }
    }
    KEV_END(1);
  }
  KEV_BEGIN(7);
  DISPATCH_K(K, {
    k_alpha_finish<KT><<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, m, (double)ctx->N, ctx->d_group_n, ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS, c.adev, ctx->cvec, ctx->Ab);
    CKL();
    k_colsum_part<<<COLSUM_B, 256, 0, ctx->stream>>>(ctx->Ab, Gl, KK + K, ctx->colsum_part); CKL();
    k_colsum_fin<<<1, 64, 0, ctx->stream>>>(ctx->colsum_part, KK + K, ctx->red);
    CKL();
    RET(allreduce(ctx, ctx->red, KK + K, NCCL_F64));
    k_alpha_mean<KT><<<1, 32, 0, ctx->stream>>>(ctx->red, ctx->amean);
    CKL();
    k_adev<KT><<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, ctx->A, ctx->cvec, ctx->amean, o->lambda_a, n.alpha, n.adev, ctx->flags + 0);
    CKL();
  });
  // NaN anywhere (any rank) -> the whole alpha reverts (:368); then each column is clamped to median +- 4 mad over ALL genes
  // (:371-376).  Sharded: the slices (and each rank's non-finite count) are all-gathered, so every rank holds the full
  // matrix -- the control-gene rows of it feed the W step
  if (ctx->nranks > 1) {
    const long long per = (ctx->G + ctx->nranks - 1) / ctx->nranks, per_k = per * K, GK = ctx->G * K;
    const bool even = ctx->g0 == std::min<long long>(ctx->G, per * ctx->rank) && ctx->g1 == std::min<long long>(ctx->G, per * (ctx->rank + 1));
    // every rank must take the same branch: uneven shards are a property of the caller's set_design ranges on ALL ranks
    if (even) {
      const long long sl = 2 * per_k + 1;
      double* mine = ctx->agather + (long long)ctx->rank * sl;
      k_pack_alpha<<<nblk(sl, 256), 256, 0, ctx->stream>>>(n.alpha, c.alpha, (long long)Gl * K, per_k, ctx->flags + 0, mine); CKL();
      int r = g_nccl.AllGather(mine, ctx->agather, (size_t)sl, NCCL_F64, ctx->comm, ctx->stream);
      if (r != 0) return ctx->fail(RUVNB_ENCCL, "ncclAllGather (alpha): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
      k_unpack_alpha<<<nblk(GK, 256), 256, 0, ctx->stream>>>(ctx->agather, ctx->nranks, per_k, GK, ctx->alpha_full, ctx->flags + 0); CKL();
    } else {
      CK(cudaMemsetAsync(ctx->agather, 0, (size_t)(2 * GK + 1) * 8, ctx->stream));
      k_sumpath_pack<<<nblk(std::max<long long>(1, (long long)Gl * K), 256), 256, 0, ctx->stream>>>(ctx->agather, GK, ctx->g0 * K, (long long)Gl * K, n.alpha, c.alpha, ctx->flags + 0); CKL();
      RET(allreduce(ctx, ctx->agather, (size_t)(2 * GK + 1), NCCL_F64));
      k_sumpath_unpack<<<nblk(GK, 256), 256, 0, ctx->stream>>>(ctx->agather, GK, ctx->alpha_full, ctx->flags + 0); CKL();
    }
    k_revert_if<<<nblk((long long)Gl * K, 256), 256, 0, ctx->stream>>>(ctx->flags + 0, n.alpha, c.alpha, (long long)Gl * K); CKL();  // :368
    k_strided_medmad<<<K, RS_NT, 0, ctx->stream>>>(ctx->alpha_full, 1, K, (int)ctx->G, ctx->med, ctx->mad); CKL();
    k_clamp_cols<<<nblk((long long)ctx->G * K, 256), 256, 0, ctx->stream>>>(ctx->alpha_full, ctx->G, K, ctx->med, ctx->mad); CKL();
    k_clamp_cols<<<nblk((long long)Gl * K, 256), 256, 0, ctx->stream>>>(n.alpha, Gl, K, ctx->med, ctx->mad); CKL();
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_ctl, 0));   // the control-gene pack (side stream) is complete
    k_gather_rows<<<nblk((long long)nc * K, 256), 256, 0, ctx->stream>>>(ctx->alpha_full, ctx->d_ctl_global, nc, K, p_anew); CKL();
  } else {
    k_revert_if<<<nblk((long long)Gl * K, 256), 256, 0, ctx->stream>>>(ctx->flags + 0, n.alpha, c.alpha, (long long)Gl * K); CKL();  // :368
    k_strided_medmad<<<K, RS_NT, 0, ctx->stream>>>(n.alpha, 1, K, Gl, ctx->med, ctx->mad); CKL();
    k_clamp_cols<<<nblk((long long)Gl * K, 256), 256, 0, ctx->stream>>>(n.alpha, Gl, K, ctx->med, ctx->mad); CKL();
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_ctl, 0));   // the control-gene pack (side stream) is complete
    k_gather_rows<<<nblk((long long)nc * K, 256), 256, 0, ctx->stream>>>(n.alpha, ctx->d_ctl_local, nc, K, p_anew); CKL();
  }
  KEV_END(7);
  // W ------------------------------------------------------------------------------------------------
  {
    CtlPack cp;
    cp.gmean = p_gmean; cp.psi = p_psi; cp.step = p_step; cp.alpha_old = p_aold; cp.alpha_new = p_anew; cp.Mb = p_mb;
    cp.pi0 = p_pi0; cp.psi_w = p_psiw; cp.zinf = ctx->d_zinf; cp.zinb = ctx->zinb_on ? 1 : 0; cp.chunk_batch = ctx->d_chunk_batch;
    // cells are sharded over the ranks in whole chunks; the slices are summed (disjoint) afterwards
    long long ch_per = (ctx->nchunks + ctx->nranks - 1) / ctx->nranks;
    long long pb = std::min<long long>(ctx->Np, ch_per * ctx->rank * CH), pe = std::min<long long>(ctx->Np, ch_per * (ctx->rank + 1) * CH);
    if (ctx->nranks > 1) CK(cudaMemsetAsync(n.W, 0, (size_t)ctx->Np * K * 8, ctx->stream));
    if (pe > pb) {
      KEV_BEGIN(2);
      RET(launch_w_update(ctx, o, cp, p_psi, c, n, kh, pb, pe));
      KEV_END(2);
    }
    KEV_BEGIN(8);
    RET(allreduce(ctx, n.W, (size_t)ctx->Np * K, NCCL_F64));
    k_w_sumsq<<<K, 1024, 0, ctx->stream>>>(n.W, ctx->Np, ctx->sumsq); CKL();
    k_w_scale<<<nblk(ctx->Np * K, 256), 256, 0, ctx->stream>>>(n.W, ctx->Np, K, ctx->sumsq, ctx->d_chunk_valid, ctx->flags + 1); CKL();
    k_revert_if<<<nblk(ctx->Np * K, 256), 256, 0, ctx->stream>>>(ctx->flags + 1, n.W, c.W, ctx->Np * K); CKL();  // :402
    KEV_END(8);
  }
  // gmean, Mb ------------------------------------------------------------------------------------------
  {
    KEV_BEGIN(3);
    ctx->launch_parts = P;
    if (ctx->Wc) {
      // pass 1 cached every element's weight and left the per-group sums of w and w Z: what is left is T = sum w W_new per group
      if (P > 1) CK(cudaMemsetAsync(ctx->VS, 0, (size_t)P * Gl * m * K * 8, ctx->stream));
      int r = launch_pass2w(ctx, geom(ctx, Gl, nullptr), n);
      ctx->launch_parts = 1;
      RET(r);
      if (P > 1) {
        const long long nZ = (long long)Gl * m;
        k_sum_parts<<<nblk(nZ * K, 256), 256, 0, ctx->stream>>>(ctx->VS, ctx->VS, nZ * K, P); CKL();
      }
      k_s1_combine<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(Gl, m, K, ctx->SWZ, ctx->VS, n.alpha, ctx->S1); CKL();
    } else {
      if (P > 1) {
        CK(cudaMemsetAsync(ctx->S0, 0, (size_t)P * Gl * m * 8, ctx->stream));
        CK(cudaMemsetAsync(ctx->S1, 0, (size_t)P * Gl * m * 8, ctx->stream));
      }
      int r = split_rows(ctx, hub, [&](const RowGeom& g, bool huber) { return launch_pass2(ctx, huber, g, c, n, kh); });
      ctx->launch_parts = 1;
      RET(r);
      if (P > 1) {
        const long long nZ = (long long)Gl * m;
        k_sum_parts5<<<nblk(2 * nZ, 256), 256, 0, ctx->stream>>>(ctx->S0, nZ, ctx->S1, nZ, nullptr, 0, nullptr, 0, nullptr, 0, P); CKL();
      }
    }
    KEV_END(3);
  }
  KEV_BEGIN(9);
  k_gmean_mb<<<nblk(Gl, 128), 128, 0, ctx->stream>>>(Gl, m, ctx->Wc ? ctx->S0c : ctx->S0, ctx->S1, c.Mb, o->lambda_b, n.gmean, n.Mb, ctx->flags + 3, ctx->flags + 2); CKL();
  // "any NA -> the whole Mb reverts" (:424) needs the NaN count of ALL ranks.  NB: it is taken speculatively as this rank's
  // own count; the global count travels with the log-likelihood total, and the tail is redone in the (rare) case they
  // disagree.  ZINB (whose tail holds the Nelder-Mead rounds) settles the count first.
  if (ctx->zinb_on) RET(allreduce(ctx, ctx->flags + 2, 2, NCCL_INT32));
  k_mb_finalize<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(Gl, m, ctx->flags + 3, c.Mb, n.Mb); CKL();
  // psi is fixed inside the inner loop
  const size_t psi_bytes = (size_t)Gl * ctx->nbatch * 8;   // [nbatch][Gl]
  CK(cudaMemcpyAsync(n.psi, c.psi, psi_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  KEV_END(9);
  double fin[2] = {0.0, 0.0};
  if (ctx->zinb_on) {
    // step 4a (:435-467): pi0 at the NEW parameters, then the E-step weights -- recomputed on the fly from (pi0, psi_w)
    CK(cudaMemcpyAsync(n.pi0, c.pi0, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(n.psi_w, c.psi_w, psi_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    int nbad_pi0 = 0;
    RET(zinb_update_pi0(ctx, n, c.pi0, n.pi0, &nbad_pi0));
    if (nbad_pi0) CK(cudaMemcpyAsync(n.pi0, c.pi0, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));  // :454
    CK(cudaMemcpyAsync(n.psi_w, n.psi, psi_bytes, cudaMemcpyDeviceToDevice, ctx->stream));             // :457-467
    RET(loglik_of(ctx, o, n, ctx->loglik_tmp, logl_new, false, nullptr));
    fin[0] = ctx->h_flags[2]; fin[1] = ctx->h_flags[3];   // already global
  } else {
    CK(cudaMemcpyAsync(n.pi0, c.pi0, (size_t)Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(n.psi_w, c.psi_w, psi_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    // candidate log-likelihood (:469-484) + its Huber certificate (+ its Gram / score sums) for the next iteration
    RET(loglik_of(ctx, o, n, ctx->loglik_tmp, logl_new, fuse, fin));
    if (ctx->nranks > 1 && fin[1] > 0.0 && ctx->h_flags[3] == 0) {
      // another rank produced NaN in Mb and this one did not revert: revert now and score the candidate again
      const int one = 1;
      CK(cudaMemcpyAsync(ctx->flags + 3, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      k_mb_finalize<<<nblk(Gl, 8), 256, 0, ctx->stream>>>(Gl, m, ctx->flags + 3, c.Mb, n.Mb); CKL();
    }
    if (ctx->nranks > 1 && fin[1] > 0.0) {   // every rank scores again (same branch everywhere: fin is a global count)
      double f2[2];
      RET(loglik_of(ctx, o, n, ctx->loglik_tmp, logl_new, fuse, f2));
    }
  }
  // h_flags[0..4] and [6] came back with the log-likelihood total (loglik_of)
  if (bad) { bad[0] = ctx->h_flags[0]; bad[1] = ctx->h_flags[1]; bad[2] = (ctx->nranks > 1 && !ctx->zinb_on) ? (int)fin[0] : ctx->h_flags[2]; }
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
  CK(cudaMemcpyAsync(dst.psi, src.psi, Gl * ctx->nbatch * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.pi0, src.pi0, Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.sigma, src.sigma, Gl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.psi_w, src.psi_w, Gl * ctx->nbatch * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dst.wtctl, src.wtctl, (size_t)ctx->n_ctl * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  dst.sig_valid = src.sig_valid; dst.maybe_active = src.maybe_active; dst.nfail = src.nfail;
  ctx->best_eq_cur = (&dst == &ctx->best && &src == &ctx->cur) || (&dst == &ctx->cur && &src == &ctx->best);
  if (ctx->faillist_for == dst.sigma) ctx->faillist_for = nullptr;   // the list was made from what this buffer held before
  dst.p1_valid = false;   // the Gram / score sums in ctx->A.. belong to one parameter set (and one step vector) only
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
  return loglik_of(ctx, o, ctx->cur, ctx->loglik, total, fuse_ok(ctx, o));   // the next ruvnb_irls_iteration starts from its sums
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
  CK(cudaMemcpyAsync(ctx->cur.psi_w, ctx->cur.psi, (size_t)ctx->Gl * ctx->nbatch * 8, cudaMemcpyDeviceToDevice, ctx->stream));
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
  ctx->psi_ver++;   // the restored psi may differ from the one the count tables were built from (set_state / estimate_disp in between)
  // "revert zinb weight" (:494-504): the weights are recomputed from the restored parameters with the CURRENT psi
  if (ctx->zinb_on) CK(cudaMemcpyAsync(ctx->cur.psi_w, ctx->cur.psi, (size_t)ctx->Gl * ctx->nbatch * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RUVNB_OK;
}
int ruvnb_halve_steps(ruvnb_ctx* ctx, const ruvnb_opts* o) {
  if (!ctx || !o || ctx->K == 0) return RUVNB_EARG;
  CK(cudaSetDevice(ctx->device));
  k_halve<<<nblk(ctx->Gl, 256), 256, 0, ctx->stream>>>(ctx->Gl, ctx->loglik, ctx->loglik_tmp, o->step_fac, ctx->step); CKL();
  ctx->cur.p1_valid = ctx->nw.p1_valid = ctx->best.p1_valid = false;   // the Gram / score sums were taken with the old steps
  return RUVNB_OK;
}
// host hook for the dispersion refresh of ruvnb_fit_nb (nullptr removes it)
int ruvnb_set_psi_callback(ruvnb_ctx* ctx, ruvnb_psi_fn fn, void* user) {
  if (!ctx) return RUVNB_EARG;
  ctx->psi_cb = fn; ctx->psi_cb_user = user;
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
  const long long GB = (long long)Gl * ctx->nbatch;   // entries of psi: one column per batch
  double inner_ms = 0.0, disp_ms = 0.0;
  int n_trace = 0, n_inner = 0, psi_used = 0, exact_rows = 0;
  auto refresh_psi = [&](StateSet& target) -> int {
    // the dispersion routine is an injection seam (edgeR/limma/locfit arithmetic)
    if (psi_inject && n_psi > 0) {
      const double* src = psi_inject + (long long)std::min(psi_used, n_psi - 1) * GB;
      psi_used++;
      std::vector<double> tmp(GB), old(GB);
      CK(cudaMemcpyAsync(old.data(), target.psi, (size_t)GB * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (long long g = 0; g < GB; ++g) tmp[g] = std::isfinite(src[g]) ? src[g] : old[g];  // :564
      CK(cudaMemcpyAsync(target.psi, tmp.data(), (size_t)GB * 8, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->psi_ver++; target.sig_valid = false; ctx->best_eq_cur = false;
      return RUVNB_OK;
    }
    if (ctx->psi_cb) {
      // host hook (ruvnb_set_psi_callback): the dispersion of `target` is refreshed by the caller -- batch-specific dispersion runs
      // estimateDisp once per batch on that batch's cells (R/ruvIIInb.R:213-220,542-548), which the host mirror does in one
      // context per batch.  The hook reads the state with ruvnb_get_state and writes psi with ruvnb_set_state.
      if (&target != &ctx->cur) std::swap(ctx->cur, target);
      int r = ctx->psi_cb(ctx->psi_cb_user, ctx);
      if (&target != &ctx->cur) std::swap(ctx->cur, target);
      if (r != RUVNB_OK) return ctx->fail(r < 0 ? r : RUVNB_EARG, "fit_nb: the dispersion callback returned %d", r);
      ctx->psi_ver++; target.sig_valid = false; ctx->best_eq_cur = false;
      return RUVNB_OK;
    }
    if (ctx->nbatch > 1) return ctx->fail(RUVNB_ESTATE, "fit_nb: a batch-specific dispersion needs a psi schedule or ruvnb_set_psi_callback (one estimateDisp per batch)");
    cudaEvent_t d0, d1;
    CK(cudaEventCreate(&d0)); CK(cudaEventCreate(&d1));
    CK(cudaEventRecord(d0, ctx->stream));
    // estimate on the parameters held in `target` (offset = alpha W' + Mb[,grp], :535-536)
    if (&target != &ctx->cur) std::swap(ctx->cur, target);
    int r = ruvnb_estimate_disp(ctx, o, nullptr);
    if (&target != &ctx->cur) std::swap(ctx->cur, target);
    RET(r);
    ctx->psi_ver++; target.sig_valid = false; ctx->best_eq_cur = false;
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
    CK(cudaMemcpyAsync(ctx->cur.psi_w, ctx->cur.psi, (size_t)GB * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->cur.sig_valid = false;
  }
  k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->step, Gl, 1.0); CKL();  // :270
  bool conv = false;
  int iter_outer = 0;
  std::vector<double> lo;
  double s_carry = 0.0; bool have_carry = false;
  long long exact_cells = 0;
  while (!conv) {
    iter_outer++;
    std::vector<double> logl_beta;
    double s_old = s_carry;
    if (!have_carry) {
      k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->step, Gl, 1.0); CKL();  // :293 (before the sweep: its Gram sums use the steps)
      RET(loglik_of(ctx, o, ctx->cur, ctx->loglik, &s_old, fuse_ok(ctx, o)));  // :276-291
    }   // else: the sweep that closed the previous outer iteration (:577-593) scored this very state with these psi and steps
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
      exact_cells += ctx->last_fail_cells;
      bool degener = !(std::isfinite(s_old) && std::isfinite(s_new)) || (s_old > s_new);  // :487-488
      ruvnb_trace_rec rec;
      rec.outer = iter_outer; rec.iter = iter; rec.halving = halving; rec.degener = degener ? 1 : 0;
      rec.logl_old = s_old; rec.logl_new = s_new; rec.bad_alpha = bad[0]; rec.bad_W = bad[1]; rec.bad_Mb = bad[2];
      if (degener) {
        k_halve<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(Gl, ctx->loglik, ctx->loglik_tmp, o->step_fac, ctx->step); CKL();  // :490-491
        ctx->cur.p1_valid = ctx->nw.p1_valid = ctx->best.p1_valid = false;
        RET(copy_set(ctx, ctx->cur, ctx->best));  // :492
        if (ctx->zinb_on) CK(cudaMemcpyAsync(ctx->cur.psi_w, ctx->cur.psi, (size_t)GB * 8, cudaMemcpyDeviceToDevice, ctx->stream));  // :494-504
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
    k_fill<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->step, Gl, 1.0); CKL();  // the next outer iteration's :293, before the sweep
    RET(loglik_of(ctx, o, ctx->cur, ctx->loglik, &s, fuse_ok(ctx, o)));  // :577-593; also the next outer iteration's :276-291
    s_carry = s; have_carry = true;
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
    stats->n_exact_mad_rows = exact_rows; stats->n_exact_mad_cells = (int)std::min<long long>(exact_cells, 2147483647ll);
  }
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1)); CK(cudaEventDestroy(t0)); CK(cudaEventDestroy(t1));
  return RUVNB_OK;
}

}  // extern "C"
