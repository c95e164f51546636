// disp_apl.cu -- launcher of the Cox-Reid adjusted-profile-likelihood sweeps (disp.cuh: k_disp_apl, k_disp_apl_cached)
#include "disp_launch.cuh"

int launch_disp_apl(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, const DispArgs& a, const double* tabG, const double* lgr_grid) {
  DISPATCH_K(ctx->K, {
    if (a.Ecache) {
      k_disp_apl_cached<KT, DISP_ND><<<row_grid(g.nrows), ROW_THREADS, 0, ctx->stream>>>(g, disp_state<KT>(ctx, s), a, tabG, lgr_grid, ctx->dp_hist, ctx->dp_l0, DISP_NG);
      CKL();
    } else {
      LAUNCH_ROWS((k_disp_apl<KT, DISP_ND>), 1, g.nrows, g, disp_state<KT>(ctx, s), a, tabG, lgr_grid, ctx->dp_hist, ctx->dp_l0, DISP_NG);
    }
  });
  return RUVNB_OK;
}
