// sweep_w.cu -- launcher of the per-cell W update (H11: k_w_update_fast with its counting certificate, k_w_update exact); see sweeps.cuh
#include "sweeps.cuh"

int launch_w_update(ruvnb_ctx* ctx, const ruvnb_opts* o, const CtlPack& cp, const double* p_psi, const StateSet& c, StateSet& n,
                    double kh, long long pb, long long pe) {
  const int K = ctx->K, m = ctx->m, nc = ctx->n_ctl, KK = K * (K + 1) / 2;
  // exact kernel geometry: the per-cell deviance rows of all control genes must fit shared memory
  const int NACC = KK + K;
  int cpc = 32;
  const long long nstr = nc | 1;
  while (cpc > 1 && (size_t)cpc * nstr * 8 > (size_t)200 * 1024) cpc >>= 1;
  if ((size_t)cpc * nstr * 8 > (size_t)200 * 1024) return ctx->fail(RUVNB_ENOMEM, "n_ctl = %d control genes exceed the per-cell shared-memory tile (max 25600)", nc);
  size_t smem = std::max<size_t>((size_t)cpc * (size_t)nstr * 8, (size_t)8 * cpc * NACC * 8);
  if (pe <= pb) return RUVNB_OK;
  const bool fast = o->huber_fastpath && !o->robust && !ctx->zinb_on;
  if (fast) {
    k_build_ytabs<<<nc * ctx->nbatch, YT, 0, ctx->stream>>>(p_psi, nc * ctx->nbatch, nullptr, ctx->ctl_tabR, nullptr); CKL();   // [nbatch][n_ctl][YT]
    CK(cudaMemsetAsync(ctx->w_fail, 0, sizeof(int), ctx->stream));
  }
  DISPATCH_K(K, {
    CK(cudaFuncSetAttribute(k_w_update<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (fast) {
      CK(cudaFuncSetAttribute(k_w_update_fast<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_FAST_DYN_SMEM));
      k_w_update_fast<KT><<<nblk(pe - pb, 64), 256, W_FAST_DYN_SMEM, ctx->stream>>>(ctx->d_ctl_rows, nc, ctx->Np, ctx->d_chunk_grp, ctx->d_chunk_valid, cp, ctx->ctl_tabR,
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
  return RUVNB_OK;
}
