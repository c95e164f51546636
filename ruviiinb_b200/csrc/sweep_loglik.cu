// sweep_loglik.cu -- launchers of k_loglik (H6 + Huber certificate counters) and k_signdev (exact-MAD path); see sweeps.cuh
#include "sweeps.cuh"

int launch_loglik(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, double* out, int fastpath) {
  DISPATCH_K(ctx->K, {
    LAUNCH_ROWS(k_loglik<KT>, 1, g.nrows, g, gstate<KT>(ctx, s, false), out, ctx->hint, ctx->certc, fastpath);
  });
  return RUVNB_OK;
}

int launch_signdev(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s) {
  DISPATCH_K(ctx->K, {
    LAUNCH_ROWS(k_signdev<KT>, 1, g.nrows, g, gstate<KT>(ctx, s, false), ctx->D, ctx->rowmax, ctx->rownan);
  });
  return RUVNB_OK;
}
