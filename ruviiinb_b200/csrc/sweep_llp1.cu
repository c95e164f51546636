// sweep_llp1.cu -- launcher of k_ll_pass1: the candidate's log-likelihood + certificate sweep fused with the next iteration's
// Gram / score sweep (kernels.cuh); see sweeps.cuh
#include "sweeps.cuh"

// rows per CTA: 6 (168 registers per thread, 12 warps per SM: 22.3 ms per sweep at cfg 2) or 8 (128 registers, 16 warps per SM,
// spills in the pair loop: 30.3 ms); RUVNB_LLP1_NW=8 selects the latter for A/B runs
int launch_ll_pass1(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, double* out, int fastpath) {
  const int nw = ctx->opt_llp1_nw == 8 ? 8 : 6;
#define LLP1_LAUNCH(NW_)                                                                                                    \
  LAUNCH_ROWS_NW((k_ll_pass1<KT, NW_>), NW_, 2, g.nrows, g, gstate<KT>(ctx, s, true), ctx->RMW, out, ctx->hint, ctx->certc, \
                 fastpath, ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS, ctx->Wc, ctx->Wc ? ctx->SWZ : nullptr)
  DISPATCH_K(ctx->K, {
    if (nw == 6) LLP1_LAUNCH(6); else LLP1_LAUNCH(8);
  });
#undef LLP1_LAUNCH
  return RUVNB_OK;
}
