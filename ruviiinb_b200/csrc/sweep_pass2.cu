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
