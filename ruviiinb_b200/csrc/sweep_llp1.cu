// sweep_llp1.cu -- launchers of the fused sweep pair (kernels.cuh): k_ll_w, the candidate's log-likelihood + certificate sweep that
// also leaves the next iteration's element-wise working quantities in two slabs, and k_gram, the Gram / score / group sums taken
// from those slabs; see sweeps.cuh
#include "sweeps.cuh"

int launch_ll_w(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, double* out, int fastpath) {
  DISPATCH_K(ctx->K, {
    LAUNCH_ROWS((k_ll_w<KT>), 1, g.nrows, g, gstate<KT>(ctx, s, true), out, ctx->hint, ctx->certc, fastpath, ctx->Wc, ctx->WZc, ctx->ZS);
  });
  return RUVNB_OK;
}

int launch_gram(ruvnb_ctx* ctx, const RowGeom& g_) {
  RowGeom g = g_;
  g.tabL = nullptr; g.tabR = nullptr; g.lgr = nullptr;   // no count tables in this sweep
#define GRAM_LAUNCH(CFG)                                                                                                           \
  do {                                                                                                                             \
    const size_t smem_ = slab_smem<KT, 2, CFG>();                                                                                  \
    CK(cudaFuncSetAttribute((k_gram<KT, CFG>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));                          \
    k_gram<KT, CFG><<<row_grid(g.nrows) * ctx->launch_parts, ROW_THREADS, smem_, ctx->stream>>>(g, ctx->m, ctx->RMW, ctx->Wc, ctx->WZc, ctx->A, \
                                                                                               ctx->u, ctx->sw, ctx->VS, ctx->S0c, ctx->SWZ);   \
    CKL();                                                                                                                         \
  } while (0)
  DISPATCH_K(ctx->K, {
    if (ctx->opt_slab_cfg & 1) GRAM_LAUNCH(GramCfgB); else GRAM_LAUNCH(GramCfgA);
  });
#undef GRAM_LAUNCH
  return RUVNB_OK;
}
