// disp_launch.cuh -- launchers of the dispersion sweeps (disp.cuh), split over disp_nr.cu / disp_apl.cu / disp.cu for build time
#pragma once
#include "launch.cuh"
#include "disp.cuh"

// offset = alpha W' + Mb[, grp] (no gmean): the parameter set `s` with a zero gmean
template <int K>
static GeneState<K> disp_state(ruvnb_ctx* ctx, const StateSet& s) {
  GeneState<K> st = gstate<K>(ctx, s, false);
  st.gmean = ctx->dp_zeros;
  return st;
}
// one Fisher-scoring sweep for nd = DISP_ND or 1 dispersions per row, on the exp(offset) cache when a.Ecache is set   (disp_nr.cu)
int launch_disp_nr(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, const DispArgs& a, int nd);
// the adjusted-profile-likelihood sweep for DISP_ND grid dispersions -> ctx->dp_l0                                  (disp_apl.cu)
int launch_disp_apl(ruvnb_ctx* ctx, const RowGeom& g, const StateSet& s, const DispArgs& a, const double* tabG, const double* lgr_grid);
