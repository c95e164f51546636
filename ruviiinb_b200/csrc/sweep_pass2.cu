// sweep_pass2.cu -- launcher of k_pass2 (H12-H13: per-group sums for gmean and Mb); see sweeps.cuh
#include "sweeps.cuh"

int launch_pass2(ruvnb_ctx* ctx, bool huber, const RowGeom& g, const StateSet& c, const StateSet& n, double kh) {
  DISPATCH_K(ctx->K, {
    auto st = gstate<KT>(ctx, c, true);
    if (huber) LAUNCH_ROWS((k_pass2<KT, true>), 2, g.nrows, g, st, n.W, n.alpha, c.sigma, kh, ctx->S0, ctx->S1);
    else LAUNCH_ROWS((k_pass2<KT, false>), 2, g.nrows, g, st, n.W, n.alpha, c.sigma, kh, ctx->S0, ctx->S1);
  });
  return RUVNB_OK;
}

// T = sum w W_new per group from the weight cache of pass 1 (kernels.cuh k_pass2w) -> ctx->VS (free after k_alpha_finish)
int launch_pass2w(ruvnb_ctx* ctx, const RowGeom& g_, const StateSet& n) {
  RowGeom g = g_;
  g.tabL = nullptr; g.tabR = nullptr; g.lgr = nullptr;   // no count tables in this sweep
#define P2W_LAUNCH(CFG)                                                                                                   \
  do {                                                                                                                    \
    const size_t smem_ = slab_smem<KT, 1, CFG>();                                                                         \
    CK(cudaFuncSetAttribute((k_pass2w<KT, CFG>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));               \
    k_pass2w<KT, CFG><<<row_grid(g.nrows) * ctx->launch_parts, ROW_THREADS, smem_, ctx->stream>>>(g, ctx->m, n.W, ctx->Wc, ctx->VS); \
    CKL();                                                                                                                \
  } while (0)
  DISPATCH_K(ctx->K, {
    if (ctx->opt_slab_cfg & 2) P2W_LAUNCH(P2wCfgB); else P2W_LAUNCH(P2wCfgA);
  });
#undef P2W_LAUNCH
  return RUVNB_OK;
}
