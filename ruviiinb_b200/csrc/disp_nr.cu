// disp_nr.cu -- launcher of the one-group Fisher-scoring sweeps of the dispersion step (disp.cuh: k_disp_nr, k_disp_nr_cached)
#include "disp_launch.cuh"

int launch_disp_nr(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, const DispArgs& a, int nd) {
  DISPATCH_K(ctx->K, {
    if (a.Ecache) {
      if (nd == DISP_ND) k_disp_nr_cached<KT, DISP_ND><<<row_grid(g.nrows) * g.nparts, ROW_THREADS, 0, ctx->stream>>>(g, disp_state<KT>(ctx, s), a);
      else k_disp_nr_cached<KT, 1><<<row_grid(g.nrows) * g.nparts, ROW_THREADS, 0, ctx->stream>>>(g, disp_state<KT>(ctx, s), a);
      CKL();
      if (g.nparts > 1) {   // (a.nparts == g.nparts: the Newton steps from the parts' partial sums)
        if (nd == DISP_ND) k_disp_nr_finish<DISP_ND><<<nblk(g.nrows, 128), 128, 0, ctx->stream>>>(g.nrows, g.rowlist, a);
        else k_disp_nr_finish<1><<<nblk(g.nrows, 128), 128, 0, ctx->stream>>>(g.nrows, g.rowlist, a);
        CKL();
      }
    } else if (nd == DISP_ND) LAUNCH_ROWS((k_disp_nr<KT, DISP_ND>), 1, g.nrows, g, disp_state<KT>(ctx, s), a);
    else LAUNCH_ROWS((k_disp_nr<KT, 1>), 1, g.nrows, g, disp_state<KT>(ctx, s), a);
  });
  return RUVNB_OK;
}
