// res.cu -- get.res (R/get.res.R:12-129) and the fused tail of fastruvIII.nb (Mb.all block -> caps -> PAC, R/fastruvIIInb.R:719-748
// + R/makeSCE.R:36-38), block by block of `block.size` cells as the reference streams them.  Sinks: doubles / int32 / log1p to
// host memory (double-buffered device slabs, copies on a second stream) or nothing but a checksum.
#include "ctx.cuh"
#include "launch.cuh"
#include "res.cuh"
#include "fast.cuh"

namespace {

struct EvPair { cudaEvent_t a = nullptr, b = nullptr; int which = 0; };

// where the per-cell Mb of a block comes from
enum MbSource { MB_GROUP, MB_HOST_CELLS, MB_FAST };

struct ResRun {
  int type, out_kind;
  long long bs, cell_begin, cell_end;
  MbSource src;
  const double* Mb_cells;       // MB_HOST_CELLS: host G x N doubles, column-major
  double lambda_b;              // MB_FAST
  const int32_t* annot_grp;     // MB_FAST
  double* Mb_all_out;           // MB_FAST: optional host copy of Mb.all (G x N doubles, column-major)
  void* out;
  ruvnb_res_stats* stats;
  // sparse sink (out_kind F64 / LOG1P values compacted per block): host slots of a dgCMatrix, or colptr == nullptr
  int64_t* csc_colptr = nullptr; int32_t* csc_rowidx = nullptr; double* csc_values = nullptr; int64_t csc_capacity = 0; int64_t* csc_nnz = nullptr;
};

template <int TYPE>
int launch_elem(ruvnb_ctx* ctx, const ResArgs& a, int out_kind, double* tmp, void* slab, int* nanc, double* sum) {
  k_res_elem<TYPE><<<dim3(nblk(a.Gl, 8), nblk(a.B, 32)), 256, 0, ctx->stream>>>(a, tmp, nanc); CKL();
  dim3 grid(nblk(a.Gl, 32), nblk(a.B, 32));
  if (out_kind == RUVNB_OUT_I32) k_res_out<1><<<grid, 256, 0, ctx->stream>>>(a, tmp, slab, sum);
  else if (out_kind == RUVNB_OUT_LOG1P) k_res_out<2><<<grid, 256, 0, ctx->stream>>>(a, tmp, slab, sum);
  else k_res_out<0><<<grid, 256, 0, ctx->stream>>>(a, tmp, slab, sum);
  CKL();
  return RUVNB_OK;
}

int res_run(ruvnb_ctx* ctx, const ResRun& rr) {
  const int type = rr.type;
  if (type < RES_PEARSON || type > RES_QUANTILE) return ctx->fail(RUVNB_EARG, "get_res: type %d is not 1 (pearson), 2 (logcounts) or 3 (quantile)", type);
  if (!ctx->have_counts || ctx->K == 0 || !ctx->have_W) return ctx->fail(RUVNB_ESTATE, "get_res: needs counts and a fitted state (W, alpha, gmean, Mb, psi)");
  if (rr.out_kind < RUVNB_OUT_F64 || rr.out_kind > RUVNB_OUT_NONE) return ctx->fail(RUVNB_EARG, "get_res: unknown sink %d", rr.out_kind);
  if ((rr.out_kind == RUVNB_OUT_I32 || rr.out_kind == RUVNB_OUT_LOG1P) && type != RES_QUANTILE)
    return ctx->fail(RUVNB_EARG, "get_res: the int32 / log1p sinks hold percentile-adjusted counts (type 3) only");
  const bool csc = rr.csc_colptr != nullptr;
  if (csc && (rr.out_kind == RUVNB_OUT_I32 || rr.out_kind == RUVNB_OUT_NONE)) return ctx->fail(RUVNB_EARG, "get_res: the sparse sink stores doubles (sink f64 or log1p)");
  if (csc && (!rr.csc_rowidx || !rr.csc_values || !rr.csc_nnz || rr.csc_capacity < 0)) return ctx->fail(RUVNB_EARG, "get_res: sparse sink without i / x / nnz slots");
  if (rr.out_kind != RUVNB_OUT_NONE && !rr.out && !csc) return ctx->fail(RUVNB_EARG, "get_res: no output buffer");
  CK(cudaSetDevice(ctx->device));
  const int Gl = ctx->Gl, K = ctx->K;
  const long long N = ctx->N;
  const bool gene_sharded = ctx->Gl != ctx->G && ctx->nranks > 1;   // rows over the ranks: the Pearson clip needs all of them
  const long long bs = std::min<long long>(rr.bs > 0 ? rr.bs : 5000, std::max<long long>(N, 1));   // :18
  if (bs > 12288) return ctx->fail(RUVNB_EARG, "get_res: block.size %lld exceeds the 12288 cells a per-gene shared-memory tile holds", bs);
  const long long cb = rr.cell_begin, ce = rr.cell_end < 0 ? N : rr.cell_end;
  if (cb < 0 || ce > N || cb > ce || cb % bs != 0) return ctx->fail(RUVNB_EARG, "get_res: cells [%lld, %lld) must start on a block boundary inside 0..%lld", cb, ce, N);
  if (ctx->cell_shard && ctx->cell0 % bs != 0) return ctx->fail(RUVNB_EARG, "get_res: the cell shard starts at %lld, not on a block of %lld cells", ctx->cell0, bs);
  const int Gout = ctx->d_orow ? ctx->Gout : Gl;
  StateSet& s = ctx->cur;
  int ref_batch = 0;
  if (ctx->nbatch > 1) {   // which.min(colMeans(out$psi)), R/get.res.R:113,120
    if (rr.src == MB_FAST) return ctx->fail(RUVNB_EARG, "fast_res: fastruvIII.nb has no batch-specific dispersion");
    std::vector<double> hp((size_t)Gl * ctx->nbatch);
    CK(cudaMemcpyAsync(hp.data(), s.psi, hp.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double best = 0.0;
    for (int b = 0; b < ctx->nbatch; ++b) {
      double t = 0.0;
      for (int g = 0; g < Gl; ++g) t += hp[(size_t)b * Gl + g];
      if (b == 0 || t < best) { best = t; ref_batch = b; }
    }
  }
  const size_t esz = rr.out_kind == RUVNB_OUT_I32 ? 4 : 8;
  ScopedDev own;
  double *wm = nullptr, *wam = nullptr, *caps = nullptr, *mbstage = nullptr, *mbG = nullptr, *mbslab[2] = {nullptr, nullptr}, *sum = nullptr, *tmp = nullptr;
  void* slab[2] = {nullptr, nullptr};
  unsigned* hist = nullptr; SelTargets* T = nullptr; int* nanc = nullptr;
  const int ldB = (int)((bs + 31) / 32 * 32);
  CK(own.alloc(&wm, K * 8)); CK(own.alloc(&wam, (size_t)Gl * 8)); CK(own.alloc(&caps, (size_t)Gl * 3 * 8)); CK(own.alloc(&sum, 8));
  for (int i = 0; i < 2; ++i) CK(own.alloc((char**)&slab[i], (size_t)Gout * bs * 8));
  CK(own.alloc(&tmp, (size_t)Gl * ldB * 8));   // gene-major results of the block before the transpose into the slab
  if (rr.src != MB_GROUP) CK(own.alloc(&mbG, (size_t)Gl * ldB * 8));
  if (rr.src == MB_HOST_CELLS) CK(own.alloc(&mbstage, (size_t)Gl * bs * 8));
  if (rr.src == MB_FAST && rr.Mb_all_out) for (int i = 0; i < 2; ++i) CK(own.alloc(&mbslab[i], (size_t)Gl * bs * 8));
  CK(own.alloc(&hist, SEL_NT * SEL_BINS * 4)); CK(own.alloc(&T, sizeof(SelTargets))); CK(own.alloc(&nanc, 4));
  CK(cudaMemsetAsync(hist, 0, SEL_NT * SEL_BINS * 4, ctx->stream));
  CK(cudaMemsetAsync(sum, 0, 8, ctx->stream));
  // sparse sink: per slab a compacted copy (row indices + values, worst case dense), the block's column offsets, the running total
  int *c_cnt = nullptr, *c_ridx[2] = {nullptr, nullptr}; double* c_val[2] = {nullptr, nullptr}; long long *c_ptr[2] = {nullptr, nullptr}, *c_tot = nullptr, *c_blk = nullptr;
  long long* h_blk = nullptr;   // pinned: nnz of the block just compacted
  long long csc_done = 0;
  if (csc) {
    CK(own.alloc(&c_cnt, (size_t)bs * 4)); CK(own.alloc(&c_tot, 8)); CK(own.alloc(&c_blk, 8));
    for (int i = 0; i < 2; ++i) { CK(own.alloc(&c_ridx[i], (size_t)Gout * bs * 4)); CK(own.alloc(&c_val[i], (size_t)Gout * bs * 8)); CK(own.alloc(&c_ptr[i], (size_t)bs * 8)); }
    CK(cudaMemsetAsync(c_tot, 0, 8, ctx->stream));
    CK(cudaMallocHost((void**)&h_blk, 8));
  }
  FastMbPrep pr;
  if (rr.src == MB_FAST) RET(fast_mb_prepare(ctx, rr.annot_grp, &own, &pr));
  cudaStream_t copy_stream = nullptr; cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr}, t_begin = nullptr, t_end = nullptr;
  CK(cudaStreamCreate(&copy_stream));
  for (int i = 0; i < 2; ++i) { CK(cudaEventCreate(&done[i])); CK(cudaEventCreate(&copied[i])); }
  CK(cudaEventCreate(&t_begin)); CK(cudaEventCreate(&t_end));
  std::vector<EvPair> evs;
  const bool timing = rr.stats != nullptr;
  auto ev_begin = [&](int which) -> int {
    if (!timing) return RUVNB_OK;
    EvPair p; p.which = which;
    CK(cudaEventCreate(&p.a)); CK(cudaEventCreate(&p.b));
    CK(cudaEventRecord(p.a, ctx->stream));
    evs.push_back(p);
    return RUVNB_OK;
  };
  auto ev_end = [&]() -> int { if (timing) CK(cudaEventRecord(evs.back().b, ctx->stream)); return RUVNB_OK; };
  const long long launches0 = ctx->launches;
  // colMeans(W) over ALL cells (:36): column sums of the local cells (pads hold 0), summed over the ranks of a cell shard
  k_res_wmean<<<K, 1024, 0, ctx->stream>>>(s.W, ctx->Np, 1.0, wm); CKL();
  const double n_all = ctx->cell_shard ? (double)ctx->N_total : (double)N;
  if (ctx->cell_shard) RET(allreduce(ctx, wm, K, NCCL_F64));
  k_res_wamean<<<nblk(Gl, 256), 256, 0, ctx->stream>>>(Gl, K, s.alpha, wm, 1.0 / n_all, wam); CKL();
  const int Bp_max = (int)((bs + RS_NT - 1) / RS_NT * RS_NT);
  CK(cudaFuncSetAttribute(k_res_caps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * Bp_max * 8)));
  CK(cudaFuncSetAttribute(k_sel_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_NT * SEL_BINS * 4));
  CK(cudaEventRecord(t_begin, ctx->stream));
  int blk = 0;
  for (long long c0 = cb; c0 < ce; c0 += bs, ++blk) {
    const int B = (int)std::min<long long>(bs, ce - c0);
    const int sl = blk & 1;
    if (blk >= 2) CK(cudaStreamWaitEvent(ctx->stream, copied[sl], 0));   // the slab's previous contents reached the host
    ResArgs a;
    a.Gl = Gl; a.K = K; a.m = ctx->m; a.B = B; a.ldB = ldB; a.Gout = Gout; a.N = N; a.Np = ctx->Np; a.c0 = c0; a.Yp = ctx->dY;
    a.alpha = s.alpha; a.gmean = s.gmean; a.psi = s.psi; a.W = s.W; a.Mb = s.Mb; a.mbG = nullptr;
    a.cell2pos = ctx->d_cell2pos; a.chunk_grp = ctx->d_chunk_grp; a.chunk_batch = ctx->d_chunk_batch; a.ref_batch = ref_batch; a.orow = ctx->d_orow; a.wa_mean = wam; a.caps = caps; a.mtabs = ctx->d_mtabs;
    a.pi0 = ctx->res_zflag ? s.pi0 : nullptr; a.zflag = ctx->res_zflag; a.zavg = ctx->res_zavg;
    if (rr.src == MB_HOST_CELLS) {  // host G x N column-major: the block is one contiguous Gl x B piece == device [B][Gl]
      RET(ev_begin(3));
      CK(cudaMemcpyAsync(mbstage, rr.Mb_cells + (size_t)c0 * Gl, (size_t)Gl * B * 8, cudaMemcpyHostToDevice, ctx->stream));
      k_res_transpose<<<dim3(nblk(Gl, 32), nblk(B, 32)), dim3(32, 8), 0, ctx->stream>>>(mbstage, Gl, B, ldB, mbG); CKL();
      RET(ev_end());
      a.mbG = mbG;
    } else if (rr.src == MB_FAST) {  // F6 for this block, kept on the device in gene-major order
      RET(ev_begin(3));
      FastMbArgs f;
      f.Gl = Gl; f.K = K; f.m = ctx->m; f.B = B; f.Np = ctx->Np; f.c0 = c0; f.Yp = ctx->dY; f.alpha = s.alpha; f.gmean = s.gmean; f.psi = s.psi;
      f.W = s.W; f.Mb = s.Mb; f.cell2pos = ctx->d_cell2pos; f.annot = pr.d_annot; f.ctl_local = pr.d_ctl; f.rmed = pr.d_rmed; f.rmad = pr.d_rmad;
      f.lambda_b = rr.lambda_b; f.gene_major = 1; f.ldB = ldB;
      k_fast_mb<<<dim3(nblk(Gl, 32), nblk(B, 32)), 1024, 0, ctx->stream>>>(f, mbG); CKL();
      if (rr.Mb_all_out) {   // the caller also wants Mb.all itself: a second launch writes the R-layout slab
        f.gene_major = 0;
        k_fast_mb<<<dim3(nblk(Gl, 32), nblk(B, 32)), 1024, 0, ctx->stream>>>(f, mbslab[sl]); CKL();
      }
      RET(ev_end());
      a.mbG = mbG;
    }
    if (type != RES_LOGCOUNTS) {
      RET(ev_begin(0));
      k_res_caps<<<Gl, RS_NT, (size_t)2 * ((B + RS_NT - 1) / RS_NT * RS_NT) * 8, ctx->stream>>>(a, type == RES_PEARSON ? 3 : 6); CKL();
      RET(ev_end());
    }
    double* sumptr = rr.out_kind == RUVNB_OUT_NONE || timing ? sum : nullptr;
    RET(ev_begin(1));
    if (type == RES_PEARSON) {
      CK(cudaMemsetAsync(nanc, 0, 4, ctx->stream));
      RET(launch_elem<RES_PEARSON>(ctx, a, rr.out_kind, tmp, slab[sl], nanc, nullptr));
      RET(ev_end());
      RET(ev_begin(2));
      // clip at the type-7 0.5 % / 99.5 % quantiles of ALL genes x this block's cells (:63-66)
      long long len = (long long)Gout * B, ntot = len;
      if (gene_sharded) { ntot = (long long)ctx->G * B; RET(allreduce(ctx, nanc, 1, NCCL_INT32)); }
      k_sel_init<<<1, 1, 0, ctx->stream>>>(T, ntot, nanc); CKL();
      const int shifts[6] = {52, 40, 28, 16, 4, 0}, bits[6] = {12, 12, 12, 12, 12, 4};
      for (int lv = 0; lv < 6; ++lv) {
        k_sel_hist<<<296, 256, SEL_NT * SEL_BINS * 4, ctx->stream>>>((const double*)slab[sl], len, shifts[lv], bits[lv], T, hist); CKL();
        if (gene_sharded) RET(allreduce(ctx, hist, SEL_NT * SEL_BINS, NCCL_INT32));
        k_sel_pick<<<1, SEL_NT * 32, 0, ctx->stream>>>(T, hist, shifts[lv], bits[lv]); CKL();
      }
      k_sel_finish<<<1, 1, 0, ctx->stream>>>(T); CKL();
      k_res_clip<<<nblk(len, 256), 256, 0, ctx->stream>>>((double*)slab[sl], len, T, sumptr); CKL();
      RET(ev_end());
    } else if (type == RES_LOGCOUNTS) {
      RET(launch_elem<RES_LOGCOUNTS>(ctx, a, rr.out_kind, tmp, slab[sl], nanc, sumptr));
      RET(ev_end());
    } else {
      RET(launch_elem<RES_QUANTILE>(ctx, a, rr.out_kind, tmp, slab[sl], nanc, sumptr));
      RET(ev_end());
    }
    if (csc) {
      // compact the block on the device; only its non-zero entries cross PCIe.  The host needs the block's count before it can
      // place the copy: one 8-byte read-back per block (the next block's kernels are queued right after it)
      k_csc_count<<<nblk((long long)B * 32, 256), 256, 0, ctx->stream>>>((const double*)slab[sl], Gout, B, c_cnt); CKL();
      k_csc_scan<<<1, 1024, 0, ctx->stream>>>(B, c_cnt, c_tot, c_ptr[sl], c_blk); CKL();
      k_csc_fill<<<nblk((long long)B * 32, 256), 256, 0, ctx->stream>>>((const double*)slab[sl], Gout, B, c_ptr[sl], csc_done, c_ridx[sl], c_val[sl]); CKL();
      CK(cudaMemcpyAsync(h_blk, c_blk, 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaEventRecord(done[sl], ctx->stream));
      CK(cudaEventSynchronize(done[sl]));
      const long long nb = *h_blk;
      if (csc_done + nb > rr.csc_capacity) {
        cudaStreamSynchronize(copy_stream); cudaFreeHost(h_blk);
        *rr.csc_nnz = csc_done + nb;
        return ctx->fail(RUVNB_ENOMEM, "get_res: the sparse sink holds %lld entries, %lld needed after cell %lld", (long long)rr.csc_capacity, csc_done + nb, (long long)(c0 + B));
      }
      CK(cudaMemcpyAsync(rr.csc_colptr + (c0 - cb), c_ptr[sl], (size_t)B * 8, cudaMemcpyDeviceToHost, copy_stream));
      if (nb) {
        CK(cudaMemcpyAsync(rr.csc_rowidx + csc_done, c_ridx[sl], (size_t)nb * 4, cudaMemcpyDeviceToHost, copy_stream));
        CK(cudaMemcpyAsync(rr.csc_values + csc_done, c_val[sl], (size_t)nb * 8, cudaMemcpyDeviceToHost, copy_stream));
      }
      CK(cudaEventRecord(copied[sl], copy_stream));
      csc_done += nb;
      continue;
    }
    CK(cudaEventRecord(done[sl], ctx->stream));
    if (rr.out_kind != RUVNB_OUT_NONE || (rr.src == MB_FAST && rr.Mb_all_out)) {
      CK(cudaStreamWaitEvent(copy_stream, done[sl], 0));
      if (rr.out_kind != RUVNB_OUT_NONE)
        CK(cudaMemcpyAsync((char*)rr.out + (size_t)(c0 - cb) * Gout * esz, slab[sl], (size_t)Gout * B * esz, cudaMemcpyDeviceToHost, copy_stream));  // write_block, :125
      if (rr.src == MB_FAST && rr.Mb_all_out)
        CK(cudaMemcpyAsync(rr.Mb_all_out + (size_t)(c0 - cb) * Gl, mbslab[sl], (size_t)Gl * B * 8, cudaMemcpyDeviceToHost, copy_stream));
      CK(cudaEventRecord(copied[sl], copy_stream));
    } else {
      CK(cudaEventRecord(copied[sl], ctx->stream));
    }
  }
  CK(cudaEventRecord(t_end, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(copy_stream));
  if (csc) { rr.csc_colptr[ce - cb] = csc_done; *rr.csc_nnz = csc_done; CK(cudaFreeHost(h_blk)); }
  if (rr.stats) {
    ruvnb_res_stats* st = rr.stats;
    memset(st, 0, sizeof *st);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, t_begin, t_end));
    st->device_seconds = ms * 1e-3;
    for (auto& p : evs) {
      float t = 0; CK(cudaEventElapsedTime(&t, p.a, p.b));
      (p.which == 0 ? st->caps_ms : p.which == 1 ? st->elem_ms : p.which == 2 ? st->select_ms : st->mb_ms) += t;
    }
    CK(cudaMemcpy(&st->checksum, sum, 8, cudaMemcpyDeviceToHost));
    st->elements = (long long)Gout * (ce - cb);
    st->launches = ctx->launches - launches0;
  }
  for (auto& p : evs) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  CK(cudaStreamDestroy(copy_stream));
  for (int i = 0; i < 2; ++i) { CK(cudaEventDestroy(done[i])); CK(cudaEventDestroy(copied[i])); }
  CK(cudaEventDestroy(t_begin)); CK(cudaEventDestroy(t_end));
  return RUVNB_OK;
}

}  // namespace

extern "C" {

int ruvnb_get_res_ex(ruvnb_ctx* ctx, int type, int64_t block_size, const double* Mb_cells, int out_kind, void* out,
                     int64_t cell_begin, int64_t cell_end, ruvnb_res_stats* stats) {
  if (!ctx) return RUVNB_EARG;
  ResRun rr;
  rr.type = type; rr.out_kind = out_kind; rr.bs = block_size; rr.cell_begin = cell_begin; rr.cell_end = cell_end;
  rr.src = Mb_cells ? MB_HOST_CELLS : MB_GROUP; rr.Mb_cells = Mb_cells; rr.lambda_b = 0.0; rr.annot_grp = nullptr; rr.Mb_all_out = nullptr;
  rr.out = out; rr.stats = stats;
  return res_run(ctx, rr);
}

// get.res of a ZINB fit (R/get.res.R:52-57,69-70,84-85,95-96): the state's "pi0" is used where zflag_of_cell[c] != 0
int ruvnb_set_res_zinb(ruvnb_ctx* ctx, const uint8_t* zflag_of_cell, double mean_zeroinf) {
  if (!ctx) return RUVNB_EARG;
  if (!ctx->have_design) return ctx->fail(RUVNB_ESTATE, "set_res_zinb: call ruvnb_set_design first");
  CK(cudaSetDevice(ctx->device));
  if (!zflag_of_cell) { ctx->res_zflag = nullptr; ctx->res_zavg = 0.0; return RUVNB_OK; }
  uint8_t* d = nullptr;
  RET(dev_alloc(ctx, &d, ctx->N));
  CK(cudaMemcpyAsync(d, zflag_of_cell, (size_t)ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->res_zflag = d; ctx->res_zavg = mean_zeroinf;
  return RUVNB_OK;
}

int ruvnb_get_res(ruvnb_ctx* ctx, int type, int64_t block_size, const double* Mb_cells, double* out) {
  if (!ctx || !out) return RUVNB_EARG;
  return ruvnb_get_res_ex(ctx, type, block_size, Mb_cells, RUVNB_OUT_F64, out, 0, -1, nullptr);
}

int ruvnb_get_res_csc(ruvnb_ctx* ctx, int type, int64_t block_size, const double* Mb_cells, int log1p, int64_t cell_begin, int64_t cell_end,
                      int64_t* colptr, int32_t* rowidx, double* values, int64_t capacity, int64_t* nnz, ruvnb_res_stats* stats) {
  if (!ctx || !colptr || !nnz) return RUVNB_EARG;
  ResRun rr;
  rr.type = type; rr.out_kind = log1p ? RUVNB_OUT_LOG1P : RUVNB_OUT_F64; rr.bs = block_size; rr.cell_begin = cell_begin; rr.cell_end = cell_end;
  rr.src = Mb_cells ? MB_HOST_CELLS : MB_GROUP; rr.Mb_cells = Mb_cells; rr.lambda_b = 0.0; rr.annot_grp = nullptr; rr.Mb_all_out = nullptr;
  rr.out = nullptr; rr.stats = stats;
  rr.csc_colptr = colptr; rr.csc_rowidx = rowidx; rr.csc_values = values; rr.csc_capacity = capacity; rr.csc_nnz = nnz;
  return res_run(ctx, rr);
}

int ruvnb_fast_res(ruvnb_ctx* ctx, int type, double lambda_b, const int32_t* annot_grp, int64_t block_size, int out_kind, void* out,
                   double* Mb_all_out, int64_t cell_begin, int64_t cell_end, ruvnb_res_stats* stats) {
  if (!ctx || !annot_grp) return RUVNB_EARG;
  ResRun rr;
  rr.type = type; rr.out_kind = out_kind; rr.bs = block_size; rr.cell_begin = cell_begin; rr.cell_end = cell_end;
  rr.src = MB_FAST; rr.Mb_cells = nullptr; rr.lambda_b = lambda_b; rr.annot_grp = annot_grp; rr.Mb_all_out = Mb_all_out;
  rr.out = out; rr.stats = stats;
  return res_run(ctx, rr);
}

}  // extern "C"
