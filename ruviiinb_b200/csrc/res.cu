// res.cu -- get.res (R/get.res.R:12-129)
#include "ctx.cuh"
#include "launch.cuh"
#include "res.cuh"

// ---- get.res (R/get.res.R:12-129) ----------------------------------------------------------------------------------------
extern "C" int ruvnb_get_res(ruvnb_ctx* ctx, int type, int64_t block_size, const double* Mb_cells, double* out) {
  if (!ctx || !out) return RUVNB_EARG;
  if (type < RES_PEARSON || type > RES_QUANTILE) return ctx->fail(RUVNB_EARG, "get_res: type %d is not 1 (pearson), 2 (logcounts) or 3 (quantile)", type);
  if (!ctx->have_counts || ctx->K == 0 || !ctx->have_W) return ctx->fail(RUVNB_ESTATE, "get_res: needs counts and a fitted state (W, alpha, gmean, Mb, psi)");
  CK(cudaSetDevice(ctx->device));
  const int Gl = ctx->Gl, K = ctx->K;
  const long long N = ctx->N;
  long long bs = std::min<long long>(block_size > 0 ? block_size : 5000, N);   // :18
  if (bs > 12000) return ctx->fail(RUVNB_EARG, "get_res: block.size %lld exceeds the 12000 cells a per-gene shared-memory tile holds", bs);
  StateSet& s = ctx->cur;
  double *wm = nullptr, *wam = nullptr, *logcap = nullptr, *slab[2] = {nullptr, nullptr}, *mbblk = nullptr;
  unsigned* hist = nullptr; SelTargets* T = nullptr; int* nanc = nullptr;
  cudaStream_t copy_stream = nullptr; cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  CK(cudaMalloc((void**)&wm, K * 8)); CK(cudaMalloc((void**)&wam, (size_t)Gl * 8)); CK(cudaMalloc((void**)&logcap, (size_t)Gl * 3 * 8));
  for (int i = 0; i < 2; ++i) { CK(cudaMalloc((void**)&slab[i], (size_t)Gl * bs * 8)); CK(cudaEventCreate(&done[i])); CK(cudaEventCreate(&copied[i])); }
  if (Mb_cells) CK(cudaMalloc((void**)&mbblk, (size_t)Gl * bs * 8));
  CK(cudaMalloc((void**)&hist, SEL_NT * SEL_BINS * 4)); CK(cudaMalloc((void**)&T, sizeof(SelTargets))); CK(cudaMalloc((void**)&nanc, 4));
  CK(cudaMemsetAsync(hist, 0, SEL_NT * SEL_BINS * 4, ctx->stream));
  CK(cudaStreamCreate(&copy_stream));
  k_res_wmean<<<K, 1024, 0, ctx->stream>>>(s.W, ctx->Np, (double)N, wm); CKL();
  k_res_wamean<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(Gl, K, s.alpha, wm, wam); CKL();
  CK(cudaFuncSetAttribute(k_res_caps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * bs * 8)));
  CK(cudaFuncSetAttribute(k_sel_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_NT * SEL_BINS * 4));
  int blk = 0;
  for (long long c0 = 0; c0 < N; c0 += bs, ++blk) {
    const int B = (int)std::min<long long>(bs, N - c0);
    const int sl = blk & 1;
    if (blk >= 2) CK(cudaStreamWaitEvent(ctx->stream, copied[sl], 0));   // the slab's previous contents reached the host
    ResArgs a;
    a.Gl = Gl; a.K = K; a.m = ctx->m; a.B = B; a.N = N; a.Np = ctx->Np; a.c0 = c0; a.Yp = ctx->dY;
    a.alpha = s.alpha; a.gmean = s.gmean; a.psi = s.psi; a.W = s.W; a.Mb = s.Mb; a.mb_blk = nullptr;
    a.cell2pos = ctx->d_cell2pos; a.chunk_grp = ctx->d_chunk_grp; a.wa_mean = wam; a.logcap = logcap;
    if (Mb_cells) {  // host G x N column-major: the block is one contiguous Gl x B piece == device [B][Gl]
      CK(cudaMemcpyAsync(mbblk, Mb_cells + (size_t)c0 * Gl, (size_t)Gl * B * 8, cudaMemcpyHostToDevice, ctx->stream));
      a.mb_blk = mbblk;
    }
    if (type != RES_LOGCOUNTS) { k_res_caps<<<Gl, 256, (size_t)2 * B * 8, ctx->stream>>>(a); CKL(); }
    dim3 grid(nblk(Gl, 32), nblk(B, 32));
    if (type == RES_PEARSON) {
      CK(cudaMemsetAsync(nanc, 0, 4, ctx->stream));
      k_res_elem<RES_PEARSON><<<grid, 1024, 0, ctx->stream>>>(a, slab[sl], nanc); CKL();
      // clip at the type-7 0.5 % / 99.5 % quantiles of ALL genes x this block's cells (:63-66)
      long long len = (long long)Gl * B, ntot = len;
      if (ctx->nranks > 1) { ntot = (long long)ctx->G * B; RET(allreduce(ctx, nanc, 1, NCCL_INT32)); }
      k_sel_init<<<1, 1, 0, ctx->stream>>>(T, ntot, nanc); CKL();
      const int shifts[6] = {52, 40, 28, 16, 4, 0}, bits[6] = {12, 12, 12, 12, 12, 4};
      for (int lv = 0; lv < 6; ++lv) {
        k_sel_hist<<<296, 256, SEL_NT * SEL_BINS * 4, ctx->stream>>>(slab[sl], len, shifts[lv], bits[lv], T, hist); CKL();
        RET(allreduce(ctx, hist, SEL_NT * SEL_BINS, NCCL_INT32));
        k_sel_pick<<<1, SEL_NT * 32, 0, ctx->stream>>>(T, hist, shifts[lv], bits[lv]); CKL();
      }
      k_sel_finish<<<1, 1, 0, ctx->stream>>>(T); CKL();
      k_res_clip<<<nblk(len, 256), 256, 0, ctx->stream>>>(slab[sl], len, T); CKL();
    } else if (type == RES_LOGCOUNTS) {
      k_res_elem<RES_LOGCOUNTS><<<grid, 1024, 0, ctx->stream>>>(a, slab[sl], nanc); CKL();
    } else {
      k_res_elem<RES_QUANTILE><<<grid, 1024, 0, ctx->stream>>>(a, slab[sl], nanc); CKL();
    }
    CK(cudaEventRecord(done[sl], ctx->stream));
    CK(cudaStreamWaitEvent(copy_stream, done[sl], 0));
    CK(cudaMemcpyAsync(out + (size_t)c0 * Gl, slab[sl], (size_t)Gl * B * 8, cudaMemcpyDeviceToHost, copy_stream));  // write_block, :125
    CK(cudaEventRecord(copied[sl], copy_stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(copy_stream));
  CK(cudaStreamDestroy(copy_stream));
  for (int i = 0; i < 2; ++i) { CK(cudaFree(slab[i])); CK(cudaEventDestroy(done[i])); CK(cudaEventDestroy(copied[i])); }
  CK(cudaFree(wm)); CK(cudaFree(wam)); CK(cudaFree(logcap)); CK(cudaFree(hist)); CK(cudaFree(T)); CK(cudaFree(nanc));
  if (mbblk) CK(cudaFree(mbblk));
  return RUVNB_OK;
}
