// sweep_pass1.cu -- launcher of k_pass1 (H7-H10: per-gene Gram / score and group sums of Z); see sweeps.cuh
#include "sweeps.cuh"

int launch_pass1(ruvnb_ctx* ctx, bool huber, const RowGeom& g, const StateSet& c, double kh) {
  DISPATCH_K(ctx->K, {
    auto st = gstate<KT>(ctx, c, true);
    if (huber) LAUNCH_ROWS_NW((k_pass1<KT, true, P1_NW>), P1_NW, 2, g.nrows, g, st, ctx->RMW, c.sigma, kh, ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS, ctx->Wc, ctx->S0c, ctx->SWZ);
    else LAUNCH_ROWS_NW((k_pass1<KT, false, P1_NW>), P1_NW, 2, g.nrows, g, st, ctx->RMW, c.sigma, kh, ctx->A, ctx->u, ctx->sw, ctx->ZS, ctx->VS, ctx->Wc, ctx->S0c, ctx->SWZ);
  });
  return RUVNB_OK;
}
