// launch.cuh -- launch helpers of the row-streaming kernels (kernels.cuh)
#pragma once
#include "ctx.cuh"
#include "kernels.cuh"
// ---- launch helpers ---------------------------------------------------------------------------------
static RowGeom geom(ruvnb_ctx* ctx, int nrows, const int* rowlist) {
  RowGeom g;
  g.Yp = ctx->dY; g.Np = ctx->Np; g.nchunks = ctx->nchunks; g.chunk_grp = ctx->d_chunk_grp; g.chunk_valid = ctx->d_chunk_valid; g.chunk_meta = ctx->d_chunk_meta;
  g.nrows = nrows; g.rowlist = rowlist; g.nrows_dev = nullptr; g.nparts = ctx->launch_parts; g.Gl = ctx->Gl;
  if (!rowlist && nrows == ctx->Gl && ctx->roworder_ready) g.rowlist = ctx->d_roworder;   // full sweeps: heavy rows first
  g.mtabs = ctx->d_mtabs; g.tabL = ctx->tabL; g.tabR = ctx->tabR; g.lgr = ctx->lgr;
  g.zinf = ctx->zinb_on ? ctx->d_zinf : nullptr;
  return g;
}
template <int K>
static GeneState<K> gstate(ruvnb_ctx* ctx, const StateSet& s, bool with_step) {
  GeneState<K> st;
  st.gmean = s.gmean; st.Mb = s.Mb; st.alpha = s.alpha; st.psi = s.psi; st.step = with_step ? ctx->step : nullptr; st.W = s.W; st.m = ctx->m;
  st.pi0 = s.pi0; st.psi_w = s.psi_w; st.zinb = ctx->zinb_on ? 1 : 0;
  return st;
}
#define P1_NW 8   // rows per CTA of k_pass1 (register budget, see kernels.cuh)
template <int K>
static inline size_t tile_smem(int nsets) { return (size_t)2 * nsets * K * TileCfg<K>::CELLS * sizeof(double); }
static inline unsigned row_grid(int nrows) { return (unsigned)((nrows + NWARP - 1) / NWARP); }
// row-streaming launch: the double-buffered W tiles exceed the 48 KB default dynamic shared memory
#define LAUNCH_ROWS_NW(KERNEL, NW_, NSETS, NROWS, ...)                                                           \
  do {                                                                                                           \
    const size_t smem_ = tile_smem<KT>(NSETS);                                                                   \
    CK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));                   \
    KERNEL<<<(unsigned)(((NROWS) + (NW_) - 1) / (NW_)) * ctx->launch_parts, (NW_) * 32, smem_, ctx->stream>>>(__VA_ARGS__); \
    CKL();                                                                                                       \
  } while (0)
#define LAUNCH_ROWS(KERNEL, NSETS, NROWS, ...)                                                                   \
  do {                                                                                                           \
    const size_t smem_ = tile_smem<KT>(NSETS);                                                                   \
    CK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));                   \
    KERNEL<<<row_grid(NROWS) * ctx->launch_parts, ROW_THREADS, smem_, ctx->stream>>>(__VA_ARGS__);               \
    CKL();                                                                                                       \
  } while (0)
