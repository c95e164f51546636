// irls.cu -- the inner-loop body of ruvIII.nb (R/ruvIIInb.R:303-527) and the whole-fit driver (:266-609).
#include <stdlib.h>
#include "sweeps.cuh"
#include "zinb.cuh"

// per-gene log-likelihood of parameter set `s` into `out` (R/ruvIIInb.R:276-291) and, in the same sweep, the
// Huber certificate of `s` (s.sigma: +Inf certified / 0 to be evaluated exactly); total (all ranks) into
// *total if non-null
static int loglik_of(ruvnb_ctx* ctx, const ruvnb_opts* o, StateSet& s, double* out, double* total) {
  const double kh = o->robust ? 1.345 : 100.0;  // :155
  RET(ensure_ytabs(ctx, s.psi));
  const int P = ctx->nparts;
  KEV_BEGIN(4);
  ctx->launch_parts = P;
  RET(launch_loglik(ctx, geom(ctx, ctx->Gl, nullptr), s, P > 1 ? ctx->llparts : out, o->huber_fastpath));
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
      RET(launch_signdev(ctx, geom(ctx, nr, ctx->faillist + r0), s));
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
int make_rmw(ruvnb_ctx* ctx, const double* W) {
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
  const int P = ctx->nparts;
  // alpha ------------------------------------------------------------------------------------------
  RET(make_rmw(ctx, c.W));
  {
    KEV_BEGIN(1);
    ctx->launch_parts = P;
    if (P > 1) {  // groups that a part never meets must read as zero
      CK(cudaMemsetAsync(ctx->ZS, 0, (size_t)P * Gl * m * 8, ctx->stream));
      CK(cudaMemsetAsync(ctx->VS, 0, (size_t)P * Gl * m * K * 8, ctx->stream));
    }
    // certified rows (sigma = +Inf: every Huber weight is 1) take the lean instantiation, the others the general one;
    // the two lists and their lengths stay on the device (no host round trip), idle CTAs exit at once
    int r = split_rows(ctx, hub, [&](const RowGeom& g, bool huber) { return launch_pass1(ctx, huber, g, c, kh); });
    ctx->launch_parts = 1;
    RET(r);
    if (P > 1) {
      k_sum_parts<<<nblk((long long)Gl * KK, 256), 256, 0, ctx->stream>>>(ctx->A, ctx->A, (long long)Gl * KK, P); CKL();
      k_sum_parts<<<nblk((long long)Gl * K, 256), 256, 0, ctx->stream>>>(ctx->u, ctx->u, (long long)Gl * K, P); CKL();
      k_sum_parts<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(ctx->sw, ctx->sw, Gl, P); CKL();
      k_sum_parts<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(ctx->ZS, ctx->ZS, (long long)Gl * m, P); CKL();
      k_sum_parts<<<nblk((long long)Gl * m * K, 256), 256, 0, ctx->stream>>>(ctx->VS, ctx->VS, (long long)Gl * m * K, P); CKL();
    }
    KEV_END(1);
  }
  DISPATCH_K(K, {
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
    if (pe > pb) {
      KEV_BEGIN(2);
      RET(launch_w_update(ctx, o, cp, p_psi, c, n, kh, pb, pe));
      KEV_END(2);
    }
    RET(allreduce(ctx, n.W, (size_t)ctx->Np * K, NCCL_F64));
    k_w_sumsq<<<K, 1024, 0, ctx->stream>>>(n.W, ctx->Np, ctx->sumsq); CKL();
    k_w_scale<<<nblk(ctx->Np * K, 256), 256, 0, ctx->stream>>>(n.W, ctx->Np, K, ctx->sumsq, ctx->d_chunk_valid, ctx->flags + 1); CKL();
    k_revert_if<<<nblk(ctx->Np * K, 256), 256, 0, ctx->stream>>>(ctx->flags + 1, n.W, c.W, ctx->Np * K); CKL();  // :402
  }
  // gmean, Mb ------------------------------------------------------------------------------------------
  {
    KEV_BEGIN(3);
    ctx->launch_parts = P;
    if (P > 1) {
      CK(cudaMemsetAsync(ctx->S0, 0, (size_t)P * Gl * m * 8, ctx->stream));
      CK(cudaMemsetAsync(ctx->S1, 0, (size_t)P * Gl * m * 8, ctx->stream));
    }
    int r = split_rows(ctx, hub, [&](const RowGeom& g, bool huber) { return launch_pass2(ctx, huber, g, c, n, kh); });
    ctx->launch_parts = 1;
    RET(r);
    if (P > 1) {
      k_sum_parts<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(ctx->S0, ctx->S0, (long long)Gl * m, P); CKL();
      k_sum_parts<<<nblk((long long)Gl * m, 256), 256, 0, ctx->stream>>>(ctx->S1, ctx->S1, (long long)Gl * m, P); CKL();
    }
    KEV_END(3);
  }
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
