// sweep_llp1.cu -- launcher of k_ll_pass1: the candidate's log-likelihood + certificate sweep fused with the next iteration's
// Gram / score sweep (kernels.cuh); see sweeps.cuh
#include "sweeps.cuh"

// rows per CTA: 8 (128 registers per thread, 16 warps per SM) or 6 (168 registers, 12 warps per SM); RUVNB_LLP1_NW selects
#ifndef LLP1_NW_DEFAULT
#define LLP1_NW_DEFAULT 8
#endif
int launch_ll_pass1(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, double* out, int fastpath) {
  const int nw = ctx->opt_llp1_nw == 6 || ctx->opt_llp1_nw == 8 ? ctx->opt_llp1_nw : LLP1_NW_DEFAULT;
  DISPATCH_K(ctx->K, {
    if (nw == 6) {
      LAUNCH_ROWS_NW((k_ll_pass1<KT, 6>), 6, 2, g.nrows, g, gstate<KT>(ctx, s, true), ctx->RMW, out, ctx->hint, ctx->certc, fastpath,
                     ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS);
    } else {
      LAUNCH_ROWS_NW((k_ll_pass1<KT, 8>), 8, 2, g.nrows, g, gstate<KT>(ctx, s, true), ctx->RMW, out, ctx->hint, ctx->certc, fastpath,
                     ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS);
    }
  });
  return RUVNB_OK;
}
