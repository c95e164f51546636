// kernels.cuh -- the IRLS kernels of libruvnb_b200 (sm_100a).  Included (through launch.cuh) by every translation unit of the library: core.cu, irls.cu, the sweep_*.cu launchers, disp*.cu, res.cu, fast.cu, svd.cu.
//
// Data layout in HBM:
//   Yp   int32 [Gl][Np]   counts, gene-major; cells permuted so that every replicate group is one
//                         contiguous run of 128-cell chunks (tail of a group's last chunk = -1 pads)
//   W    f64   [K][Np]    unwanted factors, factor-major (== R's column-major N x k, permuted/padded)
//   alpha f64  [Gl][K], Mb f64 [Gl][m], gmean/psi/step f64 [Gl]
// Row-streaming kernels: one warp per gene row, NWARP rows per CTA, all warps walk the chunks in
// lock-step so the K x 128 tile of W (and RM_W / W_new) is staged once per CTA in shared memory with
// cp.async double buffering and reused by the NWARP rows.
#pragma once
#include "common.cuh"
#include "select.cuh"

struct RowGeom {
  const int32_t* Yp;
  long long Np;
  int nchunks;          // multiple of CHPAD; trailing chunks may be empty (chunk_valid == 0)
  const int* chunk_grp;
  const int* chunk_valid;
  const int* chunk_meta; // [nchunks] (group << 8) | valid: what the row walk reads (one load per chunk, fetched a chunk ahead)
  int nrows;            // rows this launch walks (upper bound when nrows_dev is set)
  const int* rowlist;   // nullptr -> slot == row; else row = rowlist[slot] (exact-MAD batches, partitions)
  const int* nrows_dev; // optional device-side row count: slots >= *nrows_dev idle (lists built on the device)
  int nparts;           // each row is cut into nparts runs of whole tiles, one CTA per (8-row block, part): small gene shards
                        // (several GPUs) would otherwise fill a fraction of a wave; partial sums land in part-major buffers
  int Gl;               // rows of the local shard = stride between the parts of an output buffer
  const MathTabs* mtabs;  // exp / log / log(y) tables (global copy, staged into shared memory per CTA)
  const double* tabL;   // [Gl][YT] lgamma(y+r) - lgamma(r) - lgamma(y+1)
  const double* tabR;   // [Gl][YT] log(y + r)
  const double* lgr;    // [Gl] lgamma(r)
  const uint8_t* zinf;  // [Np] zero-inflated cells (permuted like the counts) or nullptr (plain NB)
};

template <int K>
struct GeneState {  // pointers to one parameter set (current, candidate or best)
  const double* gmean;
  const double* Mb;     // [Gl][m]
  const double* alpha;  // [Gl][K]
  const double* psi;
  const double* step;
  const double* W;      // [K][Np]
  int m;
  // ZINB (R/ruvIIInb.R:237-264,435-467): pi0 and the dispersion the last E-step used; zinb = 0 -> plain NB
  const double* pi0;
  const double* psi_w;
  int zinb;
};

// chunks per shared-memory stage: 512 cells for k <= 5, 256 above (tile bytes = 2 x NSETS x K x 8 x cells)
template <int K> struct TileCfg { static constexpr int SC = (K <= 5) ? 4 : 2; static constexpr int CELLS = SC * CH; };
static inline int TileCfg_SC(int K) { return K <= 5 ? 4 : 2; }   // host mirror

// per-CTA shared tables of the row-streaming kernels
struct RowShared {
  MathTabs mt;              // 4 KB
  double tl[NWARP][YT];     // this warp's row: lgamma table
  double tr[NWARP][YT];     // this warp's row: log(y + r)
  double al[NWARP][8];      // this warp's row: alpha_g (kept out of the register file)
  int ybuf[NWARP][2][CH];   // this warp's row: the counts of the current and the next chunk (cp.async ring, 512 B per slot)
  int meta[2][4];           // (group << 8) | valid cells of the chunks of the current / next tile
  const double* wsets[2];
};

// ------------------------------------------------------------------------------------------------
// row-streaming skeleton
// ------------------------------------------------------------------------------------------------
template <int K, int NSETS, int NW = NWARP>
__device__ __forceinline__ void stage_tile(double* dst, const double* const* wsets, long long Np, int tile, int tid, int* meta_dst,
                                           const int* chunk_meta) {
  constexpr int CELLS = TileCfg<K>::CELLS;
  if (tid == 0) __pipeline_memcpy_async(meta_dst, chunk_meta + tile * TileCfg<K>::SC, TileCfg<K>::SC * 4);
  constexpr int UNITS_PER_ROW = CELLS / 2;  // 16-byte units
  constexpr int UNITS = NSETS * K * UNITS_PER_ROW;
  for (int u = tid; u < UNITS; u += NW * 32) {
    int s = u / (K * UNITS_PER_ROW);
    int rem = u - s * (K * UNITS_PER_ROW);
    int l = rem / UNITS_PER_ROW;
    int e2 = rem - l * UNITS_PER_ROW;
    const double* src = wsets[s] + (long long)l * Np + (long long)tile * CELLS + 2 * e2;
    __pipeline_memcpy_async(dst + (s * K + l) * CELLS + 2 * e2, src, 16);
  }
  __pipeline_commit();
}

// NaN-propagating pmin(1, k/|u|)  (MASS::psi.huber; R's pmin keeps NaN, CUDA's fmin does not)
__device__ __forceinline__ double huber_w(double khsig, double d) {
  double q = khsig / fabs(d);
  return (q < 1.0) ? q : ((q != q) ? q : 1.0);
}

template <int K>
struct GeneRegs {
  static constexpr int CELLS = TileCfg<K>::CELLS;
  double gmean, psi, r, step, lgr;
  bool rok;                   // psi in [1e-100, 1e100]
  bool zi;                    // zero-inflation on for this launch
  double pi0, rw, lrw;        // ZINB: pi0_g, size 1/psi_w and its log
  const uint8_t* zinf;
  double* as;                 // alpha_g of the row in shared memory (RowShared::al)
  const double* mb;
  double mbj;
  const double *tl, *tr;      // this warp's count tables (shared memory)
  const MathTabs* mt;         // shared memory
  __device__ __forceinline__ void load(const GeneState<K>& st, int row) {
    gmean = st.gmean[row];
    psi = st.psi[row];
    r = 1.0 / psi;
    rok = (psi > 1e-100) & (psi < 1e100);
    zi = st.zinb != 0;
    pi0 = 0.0; rw = r; lrw = 0.0;
    if (zi) { pi0 = st.pi0[row]; rw = 1.0 / st.psi_w[row]; lrw = log(rw); }
    step = st.step ? st.step[row] : 1.0;
    {
      const int lane = threadIdx.x & 31;
      if (lane < K) as[lane] = st.alpha[(long long)row * K + lane];
      __syncwarp();
    }
    mb = st.Mb + (long long)row * st.m;
    mbj = 0.0;
  }
  // column b of psi / psi_w ([nbatch][Gl]) for this row: what load() took from column 0
  __device__ __forceinline__ void rebatch(const GeneState<K>& st, int row, int b, int Gl) {
    psi = st.psi[(long long)b * Gl + row];
    r = 1.0 / psi;
    rok = (psi > 1e-100) & (psi < 1e100);
    if (zi) { rw = 1.0 / st.psi_w[(long long)b * Gl + row]; lrw = log(rw); } else rw = r;
  }
  __device__ __forceinline__ double lmu(const double* w, int ci) const {
    double s = gmean + mbj;
#pragma unroll
    for (int l = 0; l < K; ++l) s = fma(as[l], w[l * CELLS + ci], s);
    return s;
  }
  // two adjacent cells (ti even) with one 16-byte shared-memory load per factor
  __device__ __forceinline__ void lmu_pair(const double* w, int ti, double& l0, double& l1) const {
    double s0 = gmean + mbj, s1 = s0;
#pragma unroll
    for (int l = 0; l < K; ++l) {
      double2 v = *reinterpret_cast<const double2*>(w + l * CELLS + ti);
      const double al = as[l];
      s0 = fma(al, v.x, s0);
      s1 = fma(al, v.y, s1);
    }
    l0 = s0; l1 = s1;
  }
  // E-step weight of a zero count in a zero-inflated cell: 1 - pi0 / (pi0 + (1 - pi0) dnbinom(0; mu, size = rw))
  // (R/ruvIIInb.R:262,465,502); 1 elsewhere.  pos = position of the cell in the permuted row.
  __device__ __forceinline__ double wz(int y, double mu, long long pos) const {
    if (y != 0 || !zinf[pos]) return 1.0;
    double q = fexp(-rw * (flog(mu + rw, mt->rc, mt->lc) - lrw), mt->ex);
    return 1.0 - pi0 / (pi0 + (1.0 - pi0) * q);
  }
  // both linear predictors inside the range where fexp_nc / flog_nc / frcp need no checks
  __device__ __forceinline__ bool fast_range(double l0, double l1) const { return rok & (fabs(l0) < 230.0) & (fabs(l1) < 230.0); }
  __device__ __forceinline__ double expm(double x) const { return fexp(x, mt->ex); }
  __device__ __forceinline__ double logL(double mu) const { return flog(mu + r, mt->rc, mt->lc); }
};

// ops whose arithmetic depends on the dispersion declare `static constexpr bool USES_PSI = true`, keep their GeneState in `st` and
// provide psi_changed(): the skeleton switches them to the batch of the chunk (batch-specific dispersion, psi[g, batch(c)])
template <class T, class = void> struct OpUsesPsi { static constexpr bool value = false; };
template <class T> struct OpUsesPsi<T, decltype((void)T::USES_PSI)> { static constexpr bool value = T::USES_PSI; };
// ops that take nothing from the counts (they stream another per-element row) declare `static constexpr bool NO_Y = true`
template <class T, class = void> struct OpNoY { static constexpr bool value = false; };
template <class T> struct OpNoY<T, decltype((void)T::NO_Y)> { static constexpr bool value = T::NO_Y; };

// a row of per-element doubles ([Np], same positions as the counts) fetched one 128-cell chunk ahead: a lane's two pairs of
// chunk c sit at c CH + 2 lane and c CH + 64 + 2 lane (the exp(offset) cache of the dispersion sweeps is read the same way)
struct RowF64 {
  const double* base; double2 na, nb, hold; int pre_c, c_end;
  __device__ __forceinline__ void begin(const double* rowbase, int cend) { base = rowbase; pre_c = -1; c_end = cend; }
  __device__ __forceinline__ double2 get(int c, int ti) {
    if (ti & 64) return hold;
    const int lane2 = 2 * (threadIdx.x & 31);
    double2 cur;
    if (pre_c == c) { cur = na; hold = nb; }
    else {
      cur = __ldcs(reinterpret_cast<const double2*>(base + (long long)c * CH + lane2));
      hold = __ldcs(reinterpret_cast<const double2*>(base + (long long)c * CH + 64 + lane2));
    }
    if (c + 1 < c_end) {
      na = __ldcs(reinterpret_cast<const double2*>(base + (long long)(c + 1) * CH + lane2));
      nb = __ldcs(reinterpret_cast<const double2*>(base + (long long)(c + 1) * CH + 64 + lane2));
      pre_c = c + 1;
    }
    return cur;
  }
};

// Op interface: row_begin(row, slot), group_begin(j), pair<FULLCHUNK>(y0, y1, chunk, ti, tile, m0, m1) -- two
// adjacent cells at tile index ti (even), masks m0/m1 false for padding cells (always true when FULLCHUNK) --
// group_end(j), row_end(row).  Op::PADS: pair() must also see chunks that hold no real cell.
template <int K, int NSETS, class Op, int NW = NWARP>
__device__ __forceinline__ void stream_rows(const RowGeom& geo, RowShared* sh, Op& op, double* smem) {
  constexpr int SC = TileCfg<K>::SC, CELLS = TileCfg<K>::CELLS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = blockIdx.x % geo.nparts, rblk = blockIdx.x / geo.nparts;
  const int slot = rblk * NW + warp;
  const int nrows = geo.nrows_dev ? min(geo.nrows, *geo.nrows_dev) : geo.nrows;
  if (rblk * NW >= nrows) return;   // whole CTA beyond the (device-side) list: nothing to do
  const bool active = slot < nrows;
  const int row = active ? (geo.rowlist ? geo.rowlist[slot] : slot) : 0;
  constexpr int TILE = NSETS * K * CELLS;
  // shared tables: math tables for the CTA, count tables for this warp's row
  {
    const double* src = reinterpret_cast<const double*>(geo.mtabs);
    double* dst = reinterpret_cast<double*>(&sh->mt);
    for (int i = tid; i < (int)(sizeof(MathTabs) / 8); i += NW * 32) dst[i] = src[i];
    if (active && geo.tabL) {
      for (int i = lane; i < YT; i += 32) {
        sh->tl[warp][i] = geo.tabL[(long long)row * YT + i];
        sh->tr[warp][i] = geo.tabR[(long long)row * YT + i];
      }
    }
  }
  op.g.tl = sh->tl[warp]; op.g.tr = sh->tr[warp]; op.g.mt = &sh->mt; op.g.as = sh->al[warp];
  op.g.lgr = (active && geo.lgr) ? geo.lgr[row] : 0.0;
  op.g.zinf = geo.zinf;
  __syncthreads();
  if (active) op.row_begin(row, slot);
  const int ntiles_all = geo.nchunks / SC;
  const int tpp = (ntiles_all + geo.nparts - 1) / geo.nparts;           // tiles per part
  const int t0 = min(ntiles_all, part * tpp), ntiles = min(ntiles_all, t0 + tpp);
  int curj = -1, curb = 0;
  // The counts travel HBM -> shared memory with cp.async, one 128-cell chunk (16 B per lane) ahead of the arithmetic: no
  // register holds a load in flight (with ~128 live registers ptxas spilled the prefetched values, i.e. waited for them at
  // once).  A lane reads its four cells {2l, 2l+1, 64+2l, 65+2l} back with two conflict-free 8-byte loads.
  const int* yrow = geo.Yp + (long long)row * geo.Np;
  int* yring = &sh->ybuf[warp][0][0];
  const int c_end = ntiles * SC;
  constexpr bool NOY = OpNoY<Op>::value;
  if (t0 < ntiles) {
    if (active && !NOY) __pipeline_memcpy_async(yring + ((t0 * SC) & 1) * CH + 4 * lane, yrow + (long long)t0 * SC * CH + 4 * lane, 16);
    stage_tile<K, NSETS, NW>(smem, sh->wsets, geo.Np, t0, tid, sh->meta[0], geo.chunk_meta);
  }
  for (int t = t0; t < ntiles; ++t) {
    __pipeline_wait_prior(0);
    __syncthreads();           // tile t, its chunk meta and the counts of its first chunk are in shared memory
    if (t + 1 < ntiles) stage_tile<K, NSETS, NW>(smem + ((t + 1 - t0) & 1) * TILE, sh->wsets, geo.Np, t + 1, tid, sh->meta[(t + 1 - t0) & 1], geo.chunk_meta);
    if (active) {
      const double* tile = smem + ((t - t0) & 1) * TILE;
      const int* mt_ = sh->meta[(t - t0) & 1];
#pragma unroll 1
      for (int sc = 0; sc < SC; ++sc) {
        const int c = t * SC + sc;
        int2 y01 = make_int2(0, 0), y23 = y01;
        if constexpr (!NOY) {
          if (sc > 0) { __pipeline_wait_prior(0); __syncwarp(); }   // the counts of chunk c have landed (all lanes' pieces)
          const int* yb = yring + (c & 1) * CH;
          y01 = *reinterpret_cast<const int2*>(yb + 2 * lane); y23 = *reinterpret_cast<const int2*>(yb + 64 + 2 * lane);
          if (c + 1 < c_end) { __pipeline_memcpy_async(yring + ((c + 1) & 1) * CH + 4 * lane, yrow + (long long)(c + 1) * CH + 4 * lane, 16); __pipeline_commit(); }
        }
        const int mcur = mt_[sc];
        const int nv = mcur & 0xff;
        if (nv == 0 && !Op::PADS) continue;
        if constexpr (OpUsesPsi<Op>::value) {
          // batch-specific dispersion: the chunk's batch picks the column of psi (and of the count tables) -- rare, warp-uniform
          const int b = mcur >> 24;
          if (b != curb && nv > 0) {
            curb = b;
            op.g.rebatch(op.st, row, b, geo.Gl);
            if (geo.tabL) {
              __syncwarp();
              const long long tr_ = ((long long)b * geo.Gl + row) * YT;
              for (int i = lane; i < YT; i += 32) { sh->tl[warp][i] = geo.tabL[tr_ + i]; sh->tr[warp][i] = geo.tabR[tr_ + i]; }
              __syncwarp();
            }
            if (geo.lgr) op.g.lgr = geo.lgr[(long long)b * geo.Gl + row];
            op.psi_changed();
          }
        }
        const int j = (mcur >> 8) & 0xffff;
        if (j != curj && nv > 0) {
          if (curj >= 0) op.group_end(curj);
          curj = j;
          op.group_begin(j);
        }
        const int ti = sc * CH + 2 * lane;
        if (nv == CH) {
          if constexpr (Op::QUAD) {
            op.quad(y01.x, y01.y, y23.x, y23.y, c, ti, tile);   // the lane's four cells in one branch-free block
          } else {
            op.template pair<true>(y01.x, y01.y, c, ti, tile, true, true);
            op.template pair<true>(y23.x, y23.y, c, ti + 64, tile, true, true);
          }
        } else {
          const int ci = 2 * lane;
          op.template pair<false>(y01.x, y01.y, c, ti, tile, ci < nv, ci + 1 < nv);
          op.template pair<false>(y23.x, y23.y, c, ti + 64, tile, ci + 64 < nv, ci + 65 < nv);
        }
      }
    }
    __syncthreads();
  }
  if (active) {
    if (curj >= 0) op.group_end(curj);
    op.row_end(row);
  }
}

// The same walk for ops that take nothing from the W tiles (the dispersion sweeps on the exp(offset) cache): no tile
// staging, no CTA barriers inside the row -- a warp only streams its row's counts one chunk ahead.  pair() gets
// tile == nullptr.
// geo.nparts > 1: every row is cut into nparts runs of chunks, one CTA per (row block, part) -- for launches over few rows (a small
// gene shard, the stragglers of an iteration), whose ops leave partial sums per part.
template <class Op, int NW = NWARP>
__device__ __forceinline__ void stream_rows_lean(const RowGeom& geo, RowShared* sh, Op& op) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = blockIdx.x % geo.nparts, rblk = blockIdx.x / geo.nparts;
  const int slot = rblk * NW + warp;
  const int nrows = geo.nrows_dev ? min(geo.nrows, *geo.nrows_dev) : geo.nrows;
  if (rblk * NW >= nrows) return;
  const bool active = slot < nrows;
  const int row = active ? (geo.rowlist ? geo.rowlist[slot] : slot) : 0;
  {
    const double* src = reinterpret_cast<const double*>(geo.mtabs);
    double* dst = reinterpret_cast<double*>(&sh->mt);
    for (int i = tid; i < (int)(sizeof(MathTabs) / 8); i += NW * 32) dst[i] = src[i];
  }
  op.g.tl = sh->tl[warp]; op.g.tr = sh->tr[warp]; op.g.mt = &sh->mt; op.g.as = sh->al[warp];
  op.g.lgr = 0.0;
  op.g.zinf = geo.zinf;
  __syncthreads();
  if (!active) return;
  op.row_begin(row, slot);
  const int2* yrow2 = reinterpret_cast<const int2*>(geo.Yp + (long long)row * geo.Np);
  int curj = -1;
  const int per = (geo.nchunks + geo.nparts - 1) / geo.nparts;
  const int c_lo = min(geo.nchunks, part * per), c_hi = min(geo.nchunks, c_lo + per);
  int2 ya = make_int2(0, 0), yb = ya;
  int meta = 0;
  if (c_lo < c_hi) { ya = yrow2[c_lo * (CH / 2) + lane]; yb = yrow2[c_lo * (CH / 2) + 32 + lane]; meta = geo.chunk_meta[c_lo]; }
#pragma unroll 1
  for (int c = c_lo; c < c_hi; ++c) {
    const int2 y01 = ya, y23 = yb;
    const int mcur = meta;
    if (c + 1 < c_hi) { ya = yrow2[(c + 1) * (CH / 2) + lane]; yb = yrow2[(c + 1) * (CH / 2) + 32 + lane]; meta = geo.chunk_meta[c + 1]; }
    const int nv = mcur & 0xff;
    if (nv == 0) continue;
    const int j = (mcur >> 8) & 0xffff;
    if (j != curj) {
      if (curj >= 0) op.group_end(curj);
      curj = j;
      op.group_begin(j);
    }
    const int ti = 2 * lane;
    if (nv == CH) {
      op.template pair<true>(y01.x, y01.y, c, ti, nullptr, true, true);
      op.template pair<true>(y23.x, y23.y, c, ti + 64, nullptr, true, true);
    } else {
      op.template pair<false>(y01.x, y01.y, c, ti, nullptr, ti < nv, ti + 1 < nv);
      op.template pair<false>(y23.x, y23.y, c, ti + 64, nullptr, ti + 64 < nv, ti + 65 < nv);
    }
  }
  if (curj >= 0) op.group_end(curj);
  op.row_end(row);
}

// ------------------------------------------------------------------------------------------------
// slab-streaming skeleton: the same walk (one warp per gene row, NWARP rows per CTA, group runs) for ops whose per-element input
// is not the counts but NSLAB rows of doubles ([Gl][Np] slabs written by an earlier sweep).  Everything comes in through the
// bulk-copy engine (cp.async.bulk, 1-D TMA) with mbarrier completion, as a producer / consumer pipeline of D stages of SC
// chunks: thread 0 fetches the K rows of a W tile (shared by the CTA's rows), lane 0 of every warp the matching segment of its
// own row's slabs; D - 1 stages are in flight while one is summed, and no register holds a load (the register prefetch of RowF64
// kept one chunk in flight and cost 12 registers per slab).  A warp releases a W-tile stage by arriving on its `empty` barrier
// -- there is no CTA-wide barrier inside the row, so the warps of a CTA drift up to D - 1 stages apart.
// Op interface: row_begin(row, slot), group_begin(j), pair(ti, tile, slab) with slab[q * CELLS + ti] the element of slab q at
// stage index ti (pads hold the 0 the producing sweep wrote), group_end(j), row_end(row).
// Measured at cfg 2 (k = 5): k_gram 8.5 ms with RowF64 -> 7.4 ms (512-cell stages, D = 2, CTA barriers, one CTA per SM) -> 6.5 ms
// (256-cell stages, two CTAs per SM); k_pass2w 4.0 -> 3.55 ms.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
  } while (!ok);
}
template <int SC_, int D_> struct SlabCfg { static constexpr int SC = SC_, D = D_, CELLS = SC_ * CH; };
template <int K, int NSLAB, class Cfg>
static inline size_t slab_smem() {   // host: dynamic shared memory of a slab-streaming kernel
  return (size_t)Cfg::D * (K + NWARP * NSLAB) * Cfg::CELLS * sizeof(double) + (size_t)Cfg::D * (2 + NWARP) * sizeof(uint64_t);
}
template <int K, int NSLAB, class Cfg, class Op>
__device__ __forceinline__ void stream_slabs(const RowGeom& geo, const double* wset, const double* const (&slabs)[NSLAB], Op& op, double* smem) {
  constexpr int SC = Cfg::SC, D = Cfg::D, CELLS = Cfg::CELLS, TILE = K * CELLS, SLOT = NSLAB * CELLS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = blockIdx.x % geo.nparts, rblk = blockIdx.x / geo.nparts;
  const int slot = rblk * NWARP + warp;
  const int nrows = geo.nrows_dev ? min(geo.nrows, *geo.nrows_dev) : geo.nrows;
  if (rblk * NWARP >= nrows) return;
  const bool active = slot < nrows;
  const int row = active ? (geo.rowlist ? geo.rowlist[slot] : slot) : 0;
  double* tiles = smem;                                        // [D][K][CELLS]
  double* ring = smem + D * TILE + (size_t)warp * D * SLOT;    // this warp: [D][NSLAB][CELLS]
  uint64_t* tile_full = reinterpret_cast<uint64_t*>(smem + D * TILE + (size_t)NWARP * D * SLOT);
  uint64_t* tile_empty = tile_full + D;
  uint64_t* slab_full = tile_empty + D + warp * D;
  if (tid == 0) {
    const int nact = min(NWARP, nrows - rblk * NWARP);   // warps that walk a row (and release the W-tile stages)
    for (int s = 0; s < D; ++s) { mbar_init(tile_full + s, 1); mbar_init(tile_empty + s, nact); }
    for (int w = 0; w < NWARP * D; ++w) mbar_init(tile_empty + D + w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (!active) return;
  op.row_begin(row, slot);
  const int ntiles_all = geo.nchunks / SC;
  const int tpp = (ntiles_all + geo.nparts - 1) / geo.nparts;
  const int t0 = min(ntiles_all, part * tpp), ntiles = min(ntiles_all, t0 + tpp);
  const double* srow[NSLAB];
#pragma unroll
  for (int q = 0; q < NSLAB; ++q) srow[q] = slabs[q] + (long long)row * geo.Np;
  // stage i (tile t0 + i) -> ring slot i % D; its first use needs no `empty` wait
  auto issue = [&](int i) {
    const int s = i % D, t = t0 + i;
    if (tid == 0) {
      if (i >= D) mbar_wait(tile_empty + s, (uint32_t)(i / D - 1) & 1u);   // every warp is done with stage i - D
      mbar_expect_tx(tile_full + s, (uint32_t)(TILE * sizeof(double)));
#pragma unroll
      for (int l = 0; l < K; ++l) bulk_g2s(tiles + s * TILE + l * CELLS, wset + (long long)l * geo.Np + (long long)t * CELLS, CELLS * sizeof(double), tile_full + s);
    }
    if (lane == 0) {   // (the warp's own lanes left stage i - D at the __syncwarp that closed it)
      mbar_expect_tx(slab_full + s, (uint32_t)(SLOT * sizeof(double)));
#pragma unroll
      for (int q = 0; q < NSLAB; ++q) bulk_g2s(ring + s * SLOT + q * CELLS, srow[q] + (long long)t * CELLS, CELLS * sizeof(double), slab_full + s);
    }
  };
  const int nst = ntiles - t0;
  for (int i = 0; i < D - 1 && i < nst; ++i) issue(i);
  int curj = -1;
  for (int i = 0; i < nst; ++i) {
    const int s = i % D;
    const uint32_t ph = (uint32_t)(i / D) & 1u;
    if (i + D - 1 < nst) issue(i + D - 1);
    mbar_wait(tile_full + s, ph);
    mbar_wait(slab_full + s, ph);
    const double* tile = tiles + s * TILE;
    const double* slab = ring + s * SLOT;
#pragma unroll
    for (int sc = 0; sc < SC; ++sc) {
      const int mcur = __ldg(geo.chunk_meta + (t0 + i) * SC + sc);
      const int nv = mcur & 0xff;
      if (nv == 0) continue;
      const int j = (mcur >> 8) & 0xffff;
      if (j != curj) {
        if (curj >= 0) op.group_end(curj);
        curj = j;
        op.group_begin(j);
      }
      const int ti = sc * CH + 2 * lane;
      op.pair(ti, tile, slab);
      op.pair(ti + 64, tile, slab);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tile_empty + s);
  }
  if (curj >= 0) op.group_end(curj);
  op.row_end(row);
}

// ------------------------------------------------------------------------------------------------
// K_build_ytabs: per-gene count tables (rebuilt whenever psi changes -- once per outer iteration):
//   tabL[y] = lgamma(y+r) - lgamma(r) - lgamma(y+1) = sum_{j<y} [log(r+j) - log(j+1)]   (no cancellation)
//   tabR[y] = log(y + r),  lgr = lgamma(r).   One CTA of YT threads per gene.
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(YT) k_build_ytabs(const double* psi, int rows, double* tabL, double* tabR, double* lgr) {
  __shared__ double term[YT];
  const int row = blockIdx.x, y = threadIdx.x;
  if (row >= rows) return;
  const double r = 1.0 / psi[row];
  const double lr = log((double)y + r);
  term[y] = lr - log((double)y + 1.0);
  __syncthreads();
  if (tabL) {
    double s = 0.0;
    for (int j = 0; j < y; ++j) s += term[j];
    tabL[(long long)row * YT + y] = s;
  }
  tabR[(long long)row * YT + y] = lr;
  if (y == 0 && lgr) lgr[row] = lgamma(r);
}

// ------------------------------------------------------------------------------------------------
// K_signdev: signed deviance residuals of every element -> scratch D (exact-MAD path); pads = +Inf
// (they sort last and never reach the middle ranks).  Also per-slot max|d| and a NaN flag.
// ------------------------------------------------------------------------------------------------
template <int K>
struct OpSigndev {
  static constexpr bool PADS = true;
  static constexpr bool QUAD = false;
  static constexpr bool USES_PSI = true;
  __device__ __forceinline__ void psi_changed() {}
  GeneState<K> st;
  GeneRegs<K> g;
  double* Drow;
  double* D;
  long long Np;
  double mx; int nan_; int slot_;
  double* rowmax; int* rownan;
  __device__ __forceinline__ void row_begin(int row, int slot) { g.load(st, row); Drow = D + (long long)slot * Np; mx = 0.0; nan_ = 0; slot_ = slot; }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; }
  __device__ __forceinline__ double one(int y, double lmu) {
    double mu = g.expm(lmu);
    double yd = (double)y;
    double d = nb_signdev_from_half(nb_devhalf_t(y, yd, lmu, g.logL(mu), g.r, g.tr, g.mt), yd, mu);
    if (d != d) nan_ = 1;
    mx = fmax(mx, fabs(d));
    return d;
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double l0, l1;
    g.lmu_pair(tile, ti, l0, l1);
    double2 d;
    d.x = m0 ? one(y0, l0) : inf;
    d.y = m1 ? one(y1, l1) : inf;
    *reinterpret_cast<double2*>(Drow + (long long)c * CH + (ti & (CH - 1))) = d;
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int) {
    double v = warp_max(mx);
    int nn = __any_sync(FULL, nan_);
    // max and or are order-independent: the parts of a row combine with atomics (buffers zeroed by the host)
    if ((threadIdx.x & 31) == 0) {
      atomicMax(reinterpret_cast<unsigned long long*>(rowmax + slot_), (unsigned long long)__double_as_longlong(v));  // v >= 0
      if (nn) atomicOr(rownan + slot_, 1);
    }
  }
};
template <int K>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_signdev(RowGeom geo, GeneState<K> st, double* D, double* rowmax, int* rownan) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = st.W;
  OpSigndev<K> op;
  op.st = st; op.D = D; op.Np = geo.Np; op.rowmax = rowmax; op.rownan = rownan;
  stream_rows<K, 1>(geo, &sh, op, smem);
}

// Exact sigma = mad(signdev)/0.6745 of one row from scratch D (one CTA of RS_NT threads per row), the
// certificate hints for later sweeps, and -- with the fast path on -- promotion to "certified" (+Inf)
// when k sigma >= max|d|, i.e. every Huber weight pmin(1, k sigma/|d|) is exactly 1.
static __global__ void __launch_bounds__(RS_NT) k_row_sigma(const double* D, long long Np, long long n, const int* rowlist,
                                                    const double* rowmax, const int* rownan, double kh, int fastpath,
                                                    double* sigma, double* sigma_exact, double* hint /* [Gl][3] lo, hi, tg */, double hint_half) {
  __shared__ RowSelScratch sc;
  const int slot = blockIdx.x;
  const int row = rowlist ? rowlist[slot] : slot;
  const double* d = D + (long long)slot * Np;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  double med = qnan, madraw = qnan;
  if (!rownan[slot] && n > 0) {
    unsigned long long klo, khi;
    auto key1 = [&](int i) { return f64_key(d[i]); };
    cta_two_middles((int)Np, n, key1, &sc, &klo, &khi);
    med = (n & 1) ? key_f64(khi) : (key_f64(klo) + key_f64(khi)) / 2.0;
    auto key2 = [&](int i) { return f64_key(fabs(d[i] - med)); };
    cta_two_middles((int)Np, n, key2, &sc, &klo, &khi);
    madraw = (n & 1) ? key_f64(khi) : (key_f64(klo) + key_f64(khi)) / 2.0;
  }
  if (threadIdx.x == 0) {
    const double C = MAD_CONST / 0.6745;
    double sg = MAD_CONST * madraw / 0.6745;
    double mx = rowmax[slot];
    if (sigma_exact) sigma_exact[row] = sg;
    double out = sg;
    double lo = qnan, hi = qnan, tg = -1.0;
    if (fastpath && sg == sg) {
      double t_now = mx / (kh * C) * (1.0 + 1e-9);
      if (sg > 0.0 && kh * sg >= mx * (1.0 + 1e-9)) out = __longlong_as_double(0x7ff0000000000000ll);
      double slack = madraw - t_now;
      // window: the median may move by HINT_HALF slack and max|d| may grow by (1 + 0.15 slack / t) before the row needs
      // another exact evaluation; (hi - lo)/2 + tg stays below madraw so that c3 <= (n-1)/2 is attainable
      if (slack > 0.0 && mx < 1.79769313486231570e308) { tg = t_now + 0.15 * slack; lo = med - hint_half * slack; hi = med + hint_half * slack; }
    }
    sigma[row] = out;
    hint[(long long)row * 3 + 0] = lo; hint[(long long)row * 3 + 1] = hi; hint[(long long)row * 3 + 2] = tg;
  }
}

// Certificate hints for a context that has none yet: the signed deviances of a SAMPLE of every row's cells (ns whole chunks spread over
// the row, 128 ns values per row in Ds, pads = +Inf) -- k_row_sigma then takes the median / MAD of the sample and writes the hint
// window.  A hint never changes a result (the certificate sweep re-verifies it on all cells): without this, the first sweep of a
// fit sends every row through the exact path (20 000 row walks + selections at cfg 2, ~45 ms).  One warp per row; plain libm
// arithmetic (the sample is 2 % of a sweep).
static __global__ void __launch_bounds__(256) k_hint_sample(int Gl, int K, int m, long long Np, const int32_t* Yp, const double* gmean, const double* Mb,
                                                    const double* alpha, const double* psi, const double* W, const int* sample_chunks, int ns,
                                                    const int* chunk_grp, const int* chunk_valid, const int* chunk_batch, double* Ds,
                                                    double* rowmax, int* rownan) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= Gl) return;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double gm = gmean[row];
  double al[RUVNB_MAX_K];
  for (int l = 0; l < K; ++l) al[l] = alpha[(long long)row * K + l];
  double mx = 0.0; int nan_ = 0;
  for (int si = 0; si < ns; ++si) {
    const int c = sample_chunks[si], nv = chunk_valid[c];
    const double r = 1.0 / psi[(long long)chunk_batch[c] * Gl + row];
    const double base = gm + Mb[(long long)row * m + chunk_grp[c]];
    for (int e = 0; e < 4; ++e) {
      const int cell = e * 32 + lane;
      const long long pos = (long long)c * CH + cell;
      double d = inf;
      if (cell < nv) {
        double lmu = base;
        for (int l = 0; l < K; ++l) lmu = fma(al[l], W[(long long)l * Np + pos], lmu);
        d = nb_signdev(Yp[(long long)row * Np + pos], lmu, exp(lmu), r);
        if (d != d) nan_ = 1;
        mx = fmax(mx, fabs(d));
      }
      Ds[((long long)row * ns + si) * CH + cell] = d;
    }
  }
  mx = warp_max(mx);
  nan_ = __any_sync(FULL, nan_);
  if (lane == 0) { rowmax[row] = mx; rownan[row] = nan_; }
}

// rows split by their Huber status: sigma = +Inf (every weight provably 1) -> list_inf, anything else (finite or NaN
// sigma: live weights) -> list_fin; both in launch order `order` (or natural order); one CTA, ordered scan
static __global__ void __launch_bounds__(1024) k_partition_sigma(int Gl, const double* sigma, const int* order, int* list_inf, int* list_fin,
                                                         int* counts /* [2] */) {
  __shared__ int wsum[32];
  __shared__ int base_inf, base_fin;
  if (threadIdx.x == 0) { base_inf = 0; base_fin = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int g0 = 0; g0 < Gl; g0 += 1024) {
    const int i = g0 + threadIdx.x;
    const int g = i < Gl ? (order ? order[i] : i) : 0;
    const bool inf = i < Gl && sigma[g] > 1.79769313486231570e308;
    const bool fin = i < Gl && !inf;
    unsigned bi = __ballot_sync(FULL, inf);
    if (lane == 0) wsum[w] = __popc(bi);
    __syncthreads();
    int off = 0, tot = 0;
    for (int q = 0; q < 32; ++q) { if (q < w) off += wsum[q]; tot += wsum[q]; }
    const int nvalid = min(1024, Gl - g0);
    const int pre_inf = off + __popc(bi & ((1u << lane) - 1));
    if (inf) list_inf[base_inf + pre_inf] = g;
    if (fin) list_fin[base_fin + (int)threadIdx.x - pre_inf] = g;   // rank among the non-inf = position - #inf before
    __syncthreads();
    if (threadIdx.x == 0) { base_inf += tot; base_fin += nvalid - tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { counts[0] = base_inf; counts[1] = base_fin; }
}

// rows whose certificate failed (sigma == 0 marker) -> compact list, ascending (one CTA, ordered scan)
static __global__ void __launch_bounds__(1024) k_list_failed(int Gl, const double* sigma, int* list, int* count, int* n_active) {
  __shared__ int wsum[32];
  __shared__ int base, act;
  if (threadIdx.x == 0) { base = 0; act = 0; }
  __syncthreads();
  int myact = 0;
  for (int g0 = 0; g0 < Gl; g0 += 1024) {
    int g = g0 + threadIdx.x;
    double sg = g < Gl ? sigma[g] : 0.0;
    bool f = g < Gl && sg == 0.0;
    if (g < Gl && !f && !(sg > 1.79769313486231570e308)) myact++;  // finite or NaN sigma: Huber weights are live
    unsigned bal = __ballot_sync(FULL, f);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int i = 0; i < w; ++i) off += wsum[i];
    if (f) list[off + __popc(bal & ((1u << lane) - 1))] = g;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 32; ++i) t += wsum[i]; base += t; }
    __syncthreads();
  }
  if (myact) atomicAdd(&act, myact);
  __syncthreads();
  if (threadIdx.x == 0) { *count = base; if (n_active) *n_active = act; }
}

// ------------------------------------------------------------------------------------------------
// K_loglik: per-gene NB log-likelihood (R/ruvIIInb.R:276-291) fused with the "Huber provably inactive"
// certificate for the same parameter set.  The reference computes
// sigma_g = mad(signdev_g)/0.6745 for every gene and weights pmin(1, k sigma/|d|) (:713-714); with
// k.huber = 100 (robust = FALSE, :155) those weights are all exactly 1 unless the row is degenerate.
// Given hints lo <= hi and tg (from the last exact evaluation of the row), one sweep counts
//   c1 = #{d < lo}, c2 = #{d > hi}, c3 = #{lo - tg < d < hi + tg}  and max|d|:
//   c1 <= (n-1)/2 and c2 <= n-1-n/2   =>  both middle order statistics, hence med, lie in [lo, hi];
//   t := max|d| / (k 1.4826/0.6745) <= tg  =>  every element with |d - med| < t lies in (lo - tg, hi + tg);
//   c3 <= (n-1)/2                      =>  median|d - med| >= t  =>  k sigma >= max|d|: all weights are 1.
// Comparisons are done on u = sign(d) d^2/2 (monotone in d): no square root per element.  Rows that
// fail (or have no hint yet) get sigma = 0 and go to the exact path (k_signdev + k_row_sigma).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double u_of_d(double d) { return copysign(0.5 * d * d, d); }
__device__ __forceinline__ double d_of_u(double u) { return copysign(sqrt(2.0 * fabs(u)), u); }
template <int K>
struct OpLoglik {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = true;
  static constexpr bool USES_PSI = true;
  __device__ __forceinline__ void psi_changed() { rlogr = g.r * log(g.r); }
  GeneState<K> st;
  GeneRegs<K> g;
  double acc, rlogr;
  double* out;
  // certificate
  const double* hint; int* certc; int fastpath;   // certc: this part's [Gl][4] counters c1, c2, c3, max key (or nullptr)
  double u_lo, u_hi, u_wl, u_wh; int k_lo, k_hi, k_wl, k_wh, hmaxk; int c1, c2, c3; bool cert;
  // The counting runs on 32-bit integer keys: the high word of u mapped so that integer order = floating order
  // (monotone non-decreasing under the truncation of the low word).  u < u_lo implies key(u) <= key(u_lo), etc., so the
  // integer counts can only OVER-count c1, c2, c3 -- the certificate stays rigorous -- and the FP64 pipe is spared five
  // compares per element.
  __device__ __forceinline__ static int okey(double u) { int h = __double2hiint(u); return h ^ ((h >> 31) & 0x7fffffff); }
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row); acc = 0.0; rlogr = g.r * log(g.r);
    cert = false; c1 = c2 = c3 = 0; hmaxk = 0; u_lo = u_hi = u_wl = u_wh = 0.0; k_lo = k_hi = k_wl = k_wh = 0;
    if (fastpath && certc) {
      double lo = hint[(long long)row * 3 + 0], hi = hint[(long long)row * 3 + 1], tg = hint[(long long)row * 3 + 2];
      cert = tg > 0.0 && lo <= hi;
      if (cert) {
        u_lo = u_of_d(lo); u_hi = u_of_d(hi); u_wl = u_of_d(lo - tg); u_wh = u_of_d(hi + tg);
        k_lo = okey(u_lo); k_hi = okey(u_hi); k_wl = okey(u_wl); k_wh = okey(u_wh);
      }
    }
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; }
  // generic element: any count, any argument range
  __device__ __forceinline__ void one(int y, double lmu, bool m, long long pos) {
    double mu = g.expm(lmu);
    double L = g.logL(mu);
    double yd = (double)y;
    double v = nb_logpmf_t(y, yd, lmu, L, g.r, g.lgr, rlogr, g.tl, g.mt);
    if (g.zi && m && g.zinf[pos]) {
      // ZIM::dzinb(log = TRUE): log(omega 1[y = 0] + (1 - omega) dnbinom(y)), non-finite -> log(DBL_MIN) (:282,290)
      double d = (1.0 - g.pi0) * exp(v);
      if (y == 0) d += g.pi0;
      v = log(d);
      if (!(fabs(v) <= 1.79769313486231570e308)) v = LOG_DBL_MIN;
    }
    acc += m ? v : 0.0;
    if (cert) {  // warp-uniform
      double h = nb_devhalf_t(y, yd, lmu, L, g.r, g.tr, g.mt);
      count(h, yd, mu, m);
    }
  }
  __device__ __forceinline__ void count(double h, double yd, double mu, bool m) {
    const int hi = __double2hiint(h);
    const int kraw = m ? (hi & 0x7fffffff) : 0;   // |h| key; NaN / Inf land at >= 0x7ff00000 and are caught at the row end
    hmaxk = max(hmaxk, kraw);
    const int kpos = max(hi, 0);                  // h < 0 only by rounding when y ~ mu: deviance 0
    const int ks = (yd > mu) ? kpos : ~kpos;      // ordered key of u = sign(y - mu) h
    c1 += m & (ks <= k_lo); c2 += m & (ks >= k_hi); c3 += m & (ks >= k_wl) & (ks <= k_wh);
  }
  // in-table counts and in-range arguments: branch-free (tl[0] = 0 and yd = 0 make the y = 0 case fall out)
  __device__ __forceinline__ void one_fast(int y, double lmu) {
    double mu = fexp_nc(lmu, g.mt->ex);
    double L = flog_nc(mu + g.r, g.mt->rc, g.mt->lc);
    double yd = (double)y;
    acc += fma(-g.r, L, rlogr) + fma(yd, lmu - L, g.tl[y]);
    if (cert) {  // warp-uniform
      double h = fma(yd, g.mt->ly[y] - lmu, -(yd + g.r) * (g.tr[y] - L));
      count(h, yd, mu, true);
    }
  }
  __device__ __forceinline__ void quad(int y0, int y1, int y2, int y3, int c, int ti, const double* tile) {
    double l0, l1, l2, l3;
    g.lmu_pair(tile, ti, l0, l1);
    g.lmu_pair(tile, ti + 64, l2, l3);
    if (!g.zi && (g.fast_range(l0, l1) & g.fast_range(l2, l3)) && (unsigned)(y0 | y1 | y2 | y3) < (unsigned)YT) {
      one_fast(y0, l0); one_fast(y1, l1); one_fast(y2, l2); one_fast(y3, l3);
    } else {
      const long long pos = (long long)c * CH + (ti & (CH - 1));
      one(y0, l0, true, pos); one(y1, l1, true, pos + 1); one(y2, l2, true, pos + 64); one(y3, l3, true, pos + 65);
    }
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    double l0, l1;
    g.lmu_pair(tile, ti, l0, l1);
    if (FULLCHUNK && !g.zi && g.fast_range(l0, l1) && (unsigned)(y0 | y1) < (unsigned)YT) {
      one_fast(y0, l0);
      one_fast(y1, l1);
    } else {
      const long long pos = (long long)c * CH + (ti & (CH - 1));
      if (m0) one(y0, l0, true, pos);
      if (m1) one(y1, l1, true, pos + 1);
    }
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int row) {
    double v = warp_sum(acc);
    const int lane = threadIdx.x & 31;
    if (lane == 0) out[row] = v;
    if (!certc) return;
    int a1 = __reduce_add_sync(FULL, c1), a2 = __reduce_add_sync(FULL, c2), a3 = __reduce_add_sync(FULL, c3);
    const int hk = __reduce_max_sync(FULL, hmaxk);
    if (lane == 0) { int4 o = make_int4(a1, a2, a3, hk); reinterpret_cast<int4*>(certc)[row] = o; }
  }
};
// The certificate decision of a row from the counters of its parts (see the comment above OpLoglik): sigma = +Inf
// (certified) or 0 (exact path), hint recentring, failure diagnostics.  One thread per row.
static __global__ void k_cert_finish(int Gl, int nparts, const int* certc /* [nparts][Gl][4] */, double* hint, double* sigma_out, double kh,
                              int fastpath, long long n, int* diag) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= Gl) return;
  bool ok = false, cert = false;
  if (fastpath) {
    const double lo = hint[(long long)row * 3 + 0], hi = hint[(long long)row * 3 + 1], tg = hint[(long long)row * 3 + 2];
    cert = tg > 0.0 && lo <= hi;
    if (cert) {
      long long a1 = 0, a2 = 0, a3 = 0; int hk = 0;
      for (int p = 0; p < nparts; ++p) {
        const int4 c = reinterpret_cast<const int4*>(certc)[(long long)p * Gl + row];
        a1 += c.x; a2 += c.y; a3 += c.z; hk = max(hk, c.w);
      }
      const bool nb = hk >= 0x7ff00000;                     // some deviance was NaN or infinite
      if (!nb) {
        const double hm = __hiloint2double(hk + 1, 0);      // upper bound of max |h| (one unit of the high word)
        const double C = MAD_CONST / 0.6745;
        const double mx = sqrt(2.0 * hm);
        const double t = mx / (kh * C) * (1.0 + 1e-9);
        const double u_lo = u_of_d(lo), u_hi = u_of_d(hi), u_wl = u_of_d(lo - tg), u_wh = u_of_d(hi + tg);
        // margins actually realised by the u-space thresholds
        const double room = fmin(d_of_u(u_lo) - d_of_u(u_wl), d_of_u(u_wh) - d_of_u(u_hi));
        const long long kl = (n - 1) / 2, ku = n / 2;
        ok = (hk > 0) && (mx > 0.0) && (t <= room) && (a1 <= kl) && (a2 <= n - 1 - ku) && (a3 <= kl);
        if (!ok && diag) {  // why the certificate failed (tuning aid; ruvnb_cert_diag)
          int why = !(hk > 0 && mx > 0.0) ? 2 : !(t <= room) ? 3 : (a1 > kl) ? 4 : (a2 > n - 1 - ku) ? 5 : 6;
          atomicAdd(diag + why, 1);
        }
        if (ok) {
          // recentre the hint on the median estimated from the counts (density inside [lo, hi] taken as uniform).
          // Hints never affect results -- the certificate re-verifies them -- they only decide how often the exact
          // path is needed while the parameters are still moving.
          const double pin = 1.0 - (double)(a1 + a2) / (double)n;
          if (pin > 0.05) {
            const double w = hi - lo;
            const double ctr = lo + (0.5 - (double)a1 / (double)n) * w / pin;
            hint[(long long)row * 3 + 0] = ctr - 0.5 * w;
            hint[(long long)row * 3 + 1] = ctr + 0.5 * w;
          }
        }
      } else if (diag) {
        atomicAdd(diag + 1, 1);
      }
    }
  }
  sigma_out[row] = ok ? __longlong_as_double(0x7ff0000000000000ll) : 0.0;
  if (diag && !cert) atomicAdd(diag + (fastpath ? 0 : 7), 1);   // 0: no usable hint (first visit or live Huber weights)
}
// dst[i] = sum over the parts of src[p * n + i], fixed order (deterministic); in place on part 0 when dst == src
static __global__ void k_sum_parts(double* dst, const double* src, long long n, int nparts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = src[i];
  for (int p = 1; p < nparts; ++p) s += src[(long long)p * n + i];
  dst[i] = s;
}
template <int K>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_loglik(RowGeom geo, GeneState<K> st, double* out /* [nparts][Gl] */, const double* hint,
                                                           int* certc /* [nparts][Gl][4] or nullptr */, int fastpath) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = st.W;
  const long long poff = (long long)(blockIdx.x % geo.nparts) * geo.Gl;
  OpLoglik<K> op;
  op.st = st; op.out = out + poff; op.hint = hint; op.certc = certc ? certc + poff * 4 : nullptr; op.fastpath = fastpath;
  stream_rows<K, 1>(geo, &sh, op, smem);
}

// ------------------------------------------------------------------------------------------------
// K_pass1: per-gene weighted Gram / score of get.coef2 + get.adev (R/ruvIIInb.R:712-730) and the
// unweighted group sums of Z for `projection` (:328-329), in one sweep of the row.
// Outputs per gene: A (K(K+1)/2, lower triangle row-wise), u = sum w RM_W Z (K), sw = sum w,
//                   ZS[j] = sum_{c in j} Z, VS[j][K] = sum_{c in j} w RM_W.
// ------------------------------------------------------------------------------------------------
template <int K, bool HUBER>
struct OpPass1 {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = false;
  static constexpr bool USES_PSI = true;
  __device__ __forceinline__ void psi_changed() {}
  static constexpr int KK = K * (K + 1) / 2;
  static constexpr int CELLS = TileCfg<K>::CELLS;
  GeneState<K> st;
  GeneRegs<K> g;
  const double* sigma; double kh, khsig; bool hub;
  double A[KK], u[K], sw, gz, V[K];
  double *A_out, *u_out, *sw_out, *ZS, *VS;
  // weight slab [Gl][Np] for k_pass2w + the per-group sums S0 = sum w, SWZ = sum w Z (what k_gram leaves after the fused sweep), or nullptr
  double *Wc, *S0, *SWZ; double* wrow; double gs0, gwz; long long Np;
  int row_; int lane;
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row); row_ = row; lane = threadIdx.x & 31;
    wrow = Wc ? Wc + (long long)row * Np : nullptr;
#pragma unroll
    for (int i = 0; i < KK; ++i) A[i] = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) u[i] = 0.0;
    sw = 0.0;
    hub = false;
    // sigma[row] == +Inf marks a row whose Huber weights are certified to be exactly 1
    if (HUBER) { khsig = kh * sigma[row]; hub = !(khsig > 1.79769313486231570e308); }
  }
  __device__ __forceinline__ void group_begin(int j) {
    g.mbj = g.mb[j]; gz = 0.0; gs0 = 0.0; gwz = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) V[i] = 0.0;
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    const double* w = tile; const double* rw = tile + K * CELLS;
    double l0, l1;
    g.lmu_pair(w, ti, l0, l1);
    const double yd0 = (double)y0, yd1 = (double)y1;
    double mu0, s0, q0, mu1, s1, q1;
    if (g.fast_range(l0, l1)) {
      nb_sq_fast(yd0, l0, g.psi, g.mt->ex, &mu0, &s0, &q0);
      nb_sq_fast(yd1, l1, g.psi, g.mt->ex, &mu1, &s1, &q1);
    } else {
      SQ a = nb_sq_slow(yd0, l0, g.psi), b = nb_sq_slow(yd1, l1, g.psi);
      mu0 = a.mu; s0 = a.s; q0 = a.q; mu1 = b.mu; s1 = b.s; q1 = b.q;
    }
    double Z0 = fma(g.step, q0 - 1.0, l0), Z1 = fma(g.step, q1 - 1.0, l1);
    double wt0 = s0, wt1 = s1;
    if (HUBER && hub) {  // warp-uniform branch
      wt0 *= huber_w_half(khsig, nb_devhalf_t(y0, yd0, l0, g.logL(mu0), g.r, g.tr, g.mt));
      wt1 *= huber_w_half(khsig, nb_devhalf_t(y1, yd1, l1, g.logL(mu1), g.r, g.tr, g.mt));
    }
    if (HUBER && g.zi) {  // launch-uniform: the general (HUBER) instantiation also carries the ZINB weights
      const long long pos = (long long)c * CH + (ti & (CH - 1));
      if (m0) wt0 *= g.wz(y0, mu0, pos);
      if (m1) wt1 *= g.wz(y1, mu1, pos + 1);
    }
    if (!FULLCHUNK) { wt0 = m0 ? wt0 : 0.0; Z0 = m0 ? Z0 : 0.0; wt1 = m1 ? wt1 : 0.0; Z1 = m1 ? Z1 : 0.0; }
    gs0 += wt0 + wt1; gz += Z0 + Z1;
    if (wrow) {   // launch-uniform
      gwz = fma(wt1, Z1, fma(wt0, Z0, gwz));
      __stcs(reinterpret_cast<double2*>(wrow + (long long)c * CH + (ti & (CH - 1))), make_double2(wt0, wt1));
    }
    double2 r[K];
#pragma unroll
    for (int i = 0; i < K; ++i) r[i] = *reinterpret_cast<const double2*>(rw + i * CELLS + ti);
    int idx = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      double wi0 = wt0 * r[i].x, wi1 = wt1 * r[i].y;
      V[i] += wi0 + wi1;
      u[i] = fma(wi1, Z1, fma(wi0, Z0, u[i]));
#pragma unroll
      for (int jx = 0; jx <= i; ++jx) { A[idx] = fma(wi1, r[jx].y, fma(wi0, r[jx].x, A[idx])); ++idx; }
    }
  }
  __device__ __forceinline__ void group_end(int j) {
    double z = warp_sum(gz);
    if (lane == 0) ZS[(long long)row_ * st.m + j] = z;
    sw += gs0;
    if (wrow) {
      double a = warp_sum(gs0), q = warp_sum(gwz);
      if (lane == 0) { S0[(long long)row_ * st.m + j] = a; SWZ[(long long)row_ * st.m + j] = q; }
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {
      double v = warp_sum(V[i]);
      if (lane == 0) VS[((long long)row_ * st.m + j) * K + i] = v;
    }
  }
  __device__ __forceinline__ void row_end(int row) {
#pragma unroll
    for (int i = 0; i < KK; ++i) { double v = warp_sum(A[i]); if (lane == 0) A_out[(long long)row * KK + i] = v; }
#pragma unroll
    for (int i = 0; i < K; ++i) { double v = warp_sum(u[i]); if (lane == 0) u_out[(long long)row * K + i] = v; }
    double v = warp_sum(sw);
    if (lane == 0) sw_out[row] = v;
  }
};
// NW warps (= rows) per CTA.  Measured at cfg 2: NW = 6 (170 registers, no spill, 12 warps/SM) 16.7 ms vs NW = 8 (128
// registers, 60 bytes of spill, 16 warps/SM) 16.3 ms -- the sweep is bound by FP64 issue, not by the spills.
template <int K, bool HUBER, int NW>
__global__ void __launch_bounds__(NW * 32, 2) k_pass1(RowGeom geo, GeneState<K> st, const double* RMW,
                                                          const double* sigma, double kh, double* A_out, double* u_out,
                                                          double* sw_out, double* ZS, double* VS, double* Wc, double* S0, double* SWZ) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) { sh.wsets[0] = st.W; sh.wsets[1] = RMW; }
  constexpr int KK = K * (K + 1) / 2;
  const long long poff = (long long)(blockIdx.x % geo.nparts) * geo.Gl;   // part-major partial sums
  OpPass1<K, HUBER> op;
  op.st = st; op.sigma = sigma; op.kh = kh;
  op.A_out = A_out + poff * KK; op.u_out = u_out + poff * K; op.sw_out = sw_out + poff; op.ZS = ZS + poff * st.m; op.VS = VS + poff * st.m * K;
  op.Wc = Wc; op.S0 = Wc ? S0 + poff * st.m : nullptr; op.SWZ = Wc ? SWZ + poff * st.m : nullptr; op.Np = geo.Np;
  stream_rows<K, 2, OpPass1<K, HUBER>, NW>(geo, &sh, op, smem);
}

// ------------------------------------------------------------------------------------------------
// K_ll_w: the candidate's log-likelihood sweep (k_loglik, with the Huber certificate counters) FUSED with the element-wise half
// of the NEXT inner iteration's get.coef2 (R/ruvIIInb.R:303-354).  Both walk the same row with the same parameter set: when the
// candidate is accepted (:486-516) it becomes the current state, and its lmu, exp(lmu), sig.inv and working response Z are what
// the next iteration starts from -- one read of Y, one dot product, one exp per element instead of two.  The sweep leaves, per
// element, the IRLS weight w (= sig.inv: unit Huber weight) in Wc[g][pos] and w Z in WZc[g][pos], and per group sum Z (ZS).  The
// reductions that need w, Z and W together run from those two slabs, without counts and without transcendental functions:
// k_gram (Gram / score / group sums, below) and k_pass2w.  Splitting the sweep this way keeps each kernel's live registers small:
// with the 27 FP64 accumulators of the Gram in the same kernel as the exp / log arithmetic, 128 registers spilled and 168
// registers ran 12 warps per SM at 31 % of the FP64 pipe.
// Speculative in two ways: (i) a rejected candidate is thrown away with its slabs, (ii) every Huber weight is taken as 1; rows
// whose certificate fails in this very sweep are rewritten by k_pass1<HUBER = true> once their exact sigma is known.  The
// log-likelihood half keeps the arithmetic and the summation order of OpLoglik (same bits as the separate sweep).  NB only.
// ------------------------------------------------------------------------------------------------
template <int K>
struct OpLLW {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = false;
  static constexpr bool USES_PSI = true;
  __device__ __forceinline__ void psi_changed() { rlogr = g.r * log(g.r); }
  static constexpr int CELLS = TileCfg<K>::CELLS;
  GeneState<K> st;
  GeneRegs<K> g;
  // log-likelihood + certificate (OpLoglik)
  double acc, rlogr;
  double* ll_out; const double* hint; int* certc; int fastpath;
  int k_lo, k_hi, k_wl, k_wh, hmaxk, c1, c2, c3; bool cert;
  // element-wise working quantities -> slabs
  double *Wc, *WZc, *ZS; double *wrow, *zrow; double gz; long long Np;
  int row_, lane;
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row); row_ = row; lane = threadIdx.x & 31;
    wrow = Wc + (long long)row * Np; zrow = WZc + (long long)row * Np;
    acc = 0.0; rlogr = g.r * log(g.r);
    cert = false; c1 = c2 = c3 = 0; hmaxk = 0; k_lo = k_hi = k_wl = k_wh = 0;
    if (fastpath && certc) {
      double lo = hint[(long long)row * 3 + 0], hi = hint[(long long)row * 3 + 1], tg = hint[(long long)row * 3 + 2];
      cert = tg > 0.0 && lo <= hi;
      if (cert) {
        k_lo = OpLoglik<K>::okey(u_of_d(lo)); k_hi = OpLoglik<K>::okey(u_of_d(hi));
        k_wl = OpLoglik<K>::okey(u_of_d(lo - tg)); k_wh = OpLoglik<K>::okey(u_of_d(hi + tg));
      }
    }
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; gz = 0.0; }
  __device__ __forceinline__ void count(double h, double yd, double mu, bool m) {   // OpLoglik::count
    const int hi = __double2hiint(h);
    const int kraw = m ? (hi & 0x7fffffff) : 0;
    hmaxk = max(hmaxk, kraw);
    const int kpos = max(hi, 0);
    const int ks = (yd > mu) ? kpos : ~kpos;
    c1 += m & (ks <= k_lo); c2 += m & (ks >= k_hi); c3 += m & (ks >= k_wl) & (ks <= k_wh);
  }
  // OpLoglik::one_fast given mu = fexp_nc(lmu)
  __device__ __forceinline__ void ll_fast(int y, double yd, double lmu, double mu) {
    const double L = flog_nc(mu + g.r, g.mt->rc, g.mt->lc);
    acc += fma(-g.r, L, rlogr) + fma(yd, lmu - L, g.tl[y]);
    if (cert) count(fma(yd, g.mt->ly[y] - lmu, -(yd + g.r) * (g.tr[y] - L)), yd, mu, true);
  }
  // OpLoglik::one (plain NB): any count, any argument range
  __device__ __forceinline__ void ll_gen(int y, double yd, double lmu) {
    const double mu = g.expm(lmu);
    const double L = g.logL(mu);
    acc += nb_logpmf_t(y, yd, lmu, L, g.r, g.lgr, rlogr, g.tl, g.mt);
    if (cert) count(nb_devhalf_t(y, yd, lmu, L, g.r, g.tr, g.mt), yd, mu, true);
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    double l0, l1;
    g.lmu_pair(tile, ti, l0, l1);
    const bool fr = g.fast_range(l0, l1);
    const double yd0 = (double)y0, yd1 = (double)y1;
    double mu0, s0, q0, mu1, s1, q1;
    if (fr) {
      nb_sq_fast(yd0, l0, g.psi, g.mt->ex, &mu0, &s0, &q0);
      nb_sq_fast(yd1, l1, g.psi, g.mt->ex, &mu1, &s1, &q1);
    } else {
      SQ a = nb_sq_slow(yd0, l0, g.psi), b = nb_sq_slow(yd1, l1, g.psi);
      mu0 = a.mu; s0 = a.s; q0 = a.q; mu1 = b.mu; s1 = b.s; q1 = b.q;
    }
    if (FULLCHUNK && fr && (unsigned)(y0 | y1) < (unsigned)YT) {
      ll_fast(y0, yd0, l0, mu0);
      ll_fast(y1, yd1, l1, mu1);
    } else {
      if (m0) ll_gen(y0, yd0, l0);
      if (m1) ll_gen(y1, yd1, l1);
    }
    double Z0 = fma(g.step, q0 - 1.0, l0), Z1 = fma(g.step, q1 - 1.0, l1);
    if (!FULLCHUNK) { s0 = m0 ? s0 : 0.0; Z0 = m0 ? Z0 : 0.0; s1 = m1 ? s1 : 0.0; Z1 = m1 ? Z1 : 0.0; }
    gz += Z0 + Z1;
    const long long off = (long long)c * CH + (ti & (CH - 1));
    __stcs(reinterpret_cast<double2*>(wrow + off), make_double2(s0, s1));
    __stcs(reinterpret_cast<double2*>(zrow + off), make_double2(s0 * Z0, s1 * Z1));
  }
  __device__ __forceinline__ void group_end(int j) {
    double z = warp_sum(gz);
    if (lane == 0) ZS[(long long)row_ * st.m + j] = z;
  }
  __device__ __forceinline__ void row_end(int row) {
    double v = warp_sum(acc);
    if (lane == 0) ll_out[row] = v;
    if (certc) {
      int a1 = __reduce_add_sync(FULL, c1), a2 = __reduce_add_sync(FULL, c2), a3 = __reduce_add_sync(FULL, c3);
      const int hk = __reduce_max_sync(FULL, hmaxk);
      if (lane == 0) { int4 o = make_int4(a1, a2, a3, hk); reinterpret_cast<int4*>(certc)[row] = o; }
    }
  }
};
template <int K>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_ll_w(RowGeom geo, GeneState<K> st, double* ll_out /* [nparts][Gl] */, const double* hint,
                                                         int* certc /* [nparts][Gl][4] or nullptr */, int fastpath, double* Wc, double* WZc, double* ZS) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = st.W;
  const long long poff = (long long)(blockIdx.x % geo.nparts) * geo.Gl;   // part-major partial sums
  OpLLW<K> op;
  op.st = st; op.ll_out = ll_out + poff; op.hint = hint; op.certc = certc ? certc + poff * 4 : nullptr; op.fastpath = fastpath;
  op.Wc = Wc; op.WZc = WZc; op.ZS = ZS + poff * st.m; op.Np = geo.Np;
  stream_rows<K, 1>(geo, &sh, op, smem);
}

// ------------------------------------------------------------------------------------------------
// K_gram: get.coef2's weighted Gram / score (R/ruvIIInb.R:712-720) and the group sums of `projection` / rowWtAvg (:328-329,667-675)
// from the slabs of the element-wise sweep: w (Wc) and w Z (WZc) of every (gene, cell), RM_W from shared-memory tiles.
// Per gene: A = sum w RM_W RM_W' (K(K+1)/2), u = sum w Z RM_W (K), sw = sum w; per gene and group: VS = sum w RM_W (K),
// S0 = sum w, SWZ = sum w Z.  16 bytes and K(K+1)/2 + 3K + 2 FP64 instructions per element, no counts, no exp / log.
// ------------------------------------------------------------------------------------------------
template <int K, class Cfg>
struct OpGram {
  static constexpr int KK = K * (K + 1) / 2;
  static constexpr int CELLS = Cfg::CELLS;
  int m;
  double A[KK], u[K], V[K], sw, s0, swz;
  double *A_out, *u_out, *sw_out, *VS, *S0, *SWZ;
  int row_, lane;
  __device__ __forceinline__ void row_begin(int row, int) {
    row_ = row; lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < KK; ++i) A[i] = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) u[i] = 0.0;
    sw = 0.0;
  }
  __device__ __forceinline__ void group_begin(int) {
    s0 = 0.0; swz = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) V[i] = 0.0;
  }
  __device__ __forceinline__ void pair(int ti, const double* tile, const double* slab) {   // slab 0: w, slab 1: w Z (pads hold 0)
    const double2 w = *reinterpret_cast<const double2*>(slab + ti), z = *reinterpret_cast<const double2*>(slab + CELLS + ti);
    s0 += w.x + w.y; swz += z.x + z.y;
    double2 r[K];
#pragma unroll
    for (int i = 0; i < K; ++i) r[i] = *reinterpret_cast<const double2*>(tile + i * CELLS + ti);
    int idx = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const double wi0 = w.x * r[i].x, wi1 = w.y * r[i].y;
      V[i] += wi0 + wi1;
      u[i] = fma(z.y, r[i].y, fma(z.x, r[i].x, u[i]));
#pragma unroll
      for (int jx = 0; jx <= i; ++jx) { A[idx] = fma(wi1, r[jx].y, fma(wi0, r[jx].x, A[idx])); ++idx; }
    }
  }
  __device__ __forceinline__ void group_end(int j) {
    const long long o = (long long)row_ * m + j;
    double a = warp_sum(s0), b = warp_sum(swz);
    if (lane == 0) { S0[o] = a; SWZ[o] = b; }
    sw += s0;
#pragma unroll
    for (int i = 0; i < K; ++i) { double v = warp_sum(V[i]); if (lane == 0) VS[o * K + i] = v; }
  }
  __device__ __forceinline__ void row_end(int row) {
#pragma unroll
    for (int i = 0; i < KK; ++i) { double v = warp_sum(A[i]); if (lane == 0) A_out[(long long)row * KK + i] = v; }
#pragma unroll
    for (int i = 0; i < K; ++i) { double v = warp_sum(u[i]); if (lane == 0) u_out[(long long)row * K + i] = v; }
    double v = warp_sum(sw);
    if (lane == 0) sw_out[row] = v;
  }
};
// stage shapes (RUVNB_SLAB_CFG picks B for A/B runs).  Measured at cfg 2: 128-cell stages 4 deep 7.4 ms, 256-cell stages 3 deep with
// one CTA per SM 8.3 ms, 256-cell stages 2 deep with CTA barriers 6.5 ms -- the sweep is not short of bytes in flight: FP64 pipe 48 %,
// shared-memory wavefronts 45 %, DRAM 64 % of peak, so the cheaper stage bookkeeping wins
using GramCfgA = SlabCfg<2, 2>;   // 256-cell stages, 2 deep: 84 KB at k = 5, two CTAs per SM
using GramCfgB = SlabCfg<1, 4>;   // 128-cell stages, 4 deep: 84 KB
template <int K, class Cfg>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_gram(RowGeom geo, int m, const double* RMW, const double* Wc, const double* WZc, double* A_out,
                                                         double* u_out, double* sw_out, double* VS, double* S0, double* SWZ) {
  extern __shared__ double smem[];
  constexpr int KK = K * (K + 1) / 2;
  const long long poff = (long long)(blockIdx.x % geo.nparts) * geo.Gl;
  OpGram<K, Cfg> op;
  op.m = m;
  op.A_out = A_out + poff * KK; op.u_out = u_out + poff * K; op.sw_out = sw_out + poff; op.VS = VS + poff * m * K; op.S0 = S0 + poff * m; op.SWZ = SWZ + poff * m;
  const double* const slabs[2] = {Wc, WZc};
  stream_slabs<K, 2, Cfg>(geo, RMW, slabs, op, smem);
}

// finish pass 1 per gene: c = RM_W' diag(w) RM_Z with RM_Z = Z - groupmean(Z); normalise w by its
// mean (wtvec/mean(wtvec), :715); b = c - A adev_old (:718).  One thread per gene.
template <int K>
__global__ void k_alpha_finish(int Gl, int m, double Nreal, const int* group_n, double* A, const double* u,
                               const double* sw, const double* ZS, const double* VS, const double* adev_old,
                               double* cvec, double* Ab /* [Gl][KK+K]: A | b for the global sum */) {
  constexpr int KK = K * (K + 1) / 2;
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  double c[K];
#pragma unroll
  for (int i = 0; i < K; ++i) c[i] = u[(long long)g * K + i];
  for (int j = 0; j < m; ++j) {
    double zbar = ZS[(long long)g * m + j] / (double)group_n[j];
#pragma unroll
    for (int i = 0; i < K; ++i) c[i] -= zbar * VS[((long long)g * m + j) * K + i];
  }
  double scale = Nreal / sw[g];  // 1/mean(w)
  double Af[KK];
#pragma unroll
  for (int i = 0; i < KK; ++i) { Af[i] = A[(long long)g * KK + i] * scale; A[(long long)g * KK + i] = Af[i]; }
#pragma unroll
  for (int i = 0; i < K; ++i) { c[i] *= scale; cvec[(long long)g * K + i] = c[i]; }
  // b = c - A adev_old
  double ad[K];
#pragma unroll
  for (int i = 0; i < K; ++i) ad[i] = adev_old[(long long)g * K + i];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double s = c[i];
#pragma unroll
    for (int jx = 0; jx < K; ++jx) {
      int lo = i > jx ? i : jx, sm = i > jx ? jx : i;
      s -= Af[lo * (lo + 1) / 2 + sm] * ad[jx];
    }
    Ab[(long long)g * (KK + K) + KK + i] = s;
  }
#pragma unroll
  for (int i = 0; i < KK; ++i) Ab[(long long)g * (KK + K) + i] = Af[i];
}

// deterministic column sums of X[n][ncol] (ncol <= 64): one CTA, fixed summation order
static __global__ void __launch_bounds__(1024) k_colsum(const double* X, long long n, int ncol, double* out) {
  __shared__ double sm[1024];
  for (int cidx = 0; cidx < ncol; ++cidx) {
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) s += X[i * ncol + cidx];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[cidx] = sm[0];
    __syncthreads();
  }
}

// The same in two stages for the per-iteration sums over all genes (one CTA walking 20 strided columns of 20 000 rows took 0.24 ms):
// COLSUM_B CTAs each sum a contiguous slice of rows with coalesced loads -- thread t owns column t % ncol and strides by a multiple
// of ncol -- then one CTA adds the COLSUM_B partial rows in order.  Deterministic: the slices depend on n only.
#define COLSUM_B 64
static __global__ void __launch_bounds__(256) k_colsum_part(const double* X, long long n, int ncol, double* part /* [COLSUM_B][ncol] */) {
  __shared__ double sm[256];
  const int T = (256 / ncol) * ncol;                        // active threads: a multiple of ncol
  const long long per = (n + COLSUM_B - 1) / COLSUM_B;
  const long long r0 = min(n, (long long)blockIdx.x * per), r1 = min(n, r0 + per);
  const double* x = X + r0 * ncol;
  const long long ne = (r1 - r0) * ncol;
  double s = 0.0;
  if ((int)threadIdx.x < T) for (long long e = threadIdx.x; e < ne; e += T) s += x[e];
  sm[threadIdx.x] = s;
  __syncthreads();
  if ((int)threadIdx.x < ncol) {
    double t = 0.0;
    for (int q = threadIdx.x; q < T; q += ncol) t += sm[q];
    part[(long long)blockIdx.x * ncol + threadIdx.x] = t;
  }
}
static __global__ void k_colsum_fin(const double* part, int ncol, double* out) {
  const int c = threadIdx.x;
  if (c >= ncol) return;
  double t = 0.0;
  for (int b = 0; b < COLSUM_B; ++b) t += part[(long long)b * ncol + c];
  out[c] = t;
}

// small dense solvers ---------------------------------------------------------------------------
// Gaussian elimination with partial pivoting (what R's solve() / LAPACK dgesv does), n <= 9.
template <int NMAX>
__device__ __forceinline__ bool solve_lu(int n, double (*M)[NMAX], double* rhs) {
  for (int col = 0; col < n; ++col) {
    int piv = col; double best = fabs(M[col][col]);
    for (int r2 = col + 1; r2 < n; ++r2) if (fabs(M[r2][col]) > best) { best = fabs(M[r2][col]); piv = r2; }
    if (!(best > 0.0)) return false;
    if (piv != col) {
      for (int cc = 0; cc < n; ++cc) { double t = M[col][cc]; M[col][cc] = M[piv][cc]; M[piv][cc] = t; }
      double t = rhs[col]; rhs[col] = rhs[piv]; rhs[piv] = t;
    }
    for (int r2 = col + 1; r2 < n; ++r2) {
      double f = M[r2][col] / M[col][col];
      for (int cc = col; cc < n; ++cc) M[r2][cc] -= f * M[col][cc];
      rhs[r2] -= f * rhs[col];
    }
  }
  for (int r2 = n - 1; r2 >= 0; --r2) {
    double s = rhs[r2];
    for (int cc = r2 + 1; cc < n; ++cc) s -= M[r2][cc] * rhs[cc];
    rhs[r2] = s / M[r2][r2];
  }
  return true;
}

// alpha.mean = solve(sum_g A_g, sum_g b_g)  (R/ruvIIInb.R:351-354); one thread.
template <int K>
__global__ void k_alpha_mean(const double* red /* KK + K */, double* amean) {
  constexpr int KK = K * (K + 1) / 2;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double M[K][K], rhs[K];
  for (int i = 0; i < K; ++i) {
    for (int j = 0; j <= i; ++j) { M[i][j] = red[i * (i + 1) / 2 + j]; M[j][i] = M[i][j]; }
    rhs[i] = red[KK + i];
  }
  bool ok = solve_lu<K>(K, M, rhs);
  for (int i = 0; i < K; ++i) amean[i] = ok ? rhs[i] : __longlong_as_double(0x7ff8000000000000ll);
}

// alpha.dev_g = solve(A_g + lambda_a I, c_g - A_g alpha.mean) (R/ruvIIInb.R:722-730);
// alpha = alpha.dev + alpha.mean (:365); counts non-finite entries (:367).
template <int K>
__global__ void k_adev(int Gl, const double* A, const double* cvec, const double* amean, double lambda_a,
                       double* alpha_new, double* adev_new, int* bad) {
  constexpr int KK = K * (K + 1) / 2;
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  double M[K][K], rhs[K], am[K];
  for (int i = 0; i < K; ++i) am[i] = amean[i];
  for (int i = 0; i < K; ++i)
    for (int j = 0; j <= i; ++j) { double v = A[(long long)g * KK + i * (i + 1) / 2 + j]; M[i][j] = v; M[j][i] = v; }
  for (int i = 0; i < K; ++i) {
    double s = cvec[(long long)g * K + i];
    for (int j = 0; j < K; ++j) s -= M[i][j] * am[j];
    rhs[i] = s;
  }
  for (int i = 0; i < K; ++i) M[i][i] += lambda_a;
  bool ok = solve_lu<K>(K, M, rhs);
  int nb = 0;
  for (int i = 0; i < K; ++i) {
    double x = ok ? rhs[i] : __longlong_as_double(0x7ff8000000000000ll);
    double a = x + am[i];
    adev_new[(long long)g * K + i] = x;
    alpha_new[(long long)g * K + i] = a;
    if (!(fabs(a) <= 1.79769313486231570e308)) nb++;
  }
  if (nb) atomicAdd(bad, nb);
}

// if any alpha entry was non-finite, the reference reverts the WHOLE alpha (R/ruvIIInb.R:368)
static __global__ void k_revert_if(const int* flag, double* dst, const double* src, long long n) {
  if (*flag == 0) return;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// median / mad of strided vectors (columns of alpha: R/ruvIIInb.R:371-372), one CTA of RS_NT threads per vector:
// 12-bit radix descent with early gather (select.cuh: cta_two_middles); NaN anywhere -> NaN (R's median)
static __global__ void __launch_bounds__(RS_NT) k_strided_medmad(const double* base, long long vec_stride, long long elem_stride,
                                                         int len, double* med, double* mad) {
  __shared__ RowSelScratch sc;
  __shared__ int s_nan;
  const double* p = base + (long long)blockIdx.x * vec_stride;
  const int lenp = (len + RS_NT - 1) / RS_NT * RS_NT;
  const unsigned long long kinf = f64_key(__longlong_as_double(0x7ff0000000000000ll));
  if (threadIdx.x == 0) s_nan = 0;
  __syncthreads();
  int nn = 0;
  for (int i = threadIdx.x; i < len; i += RS_NT) { double v = p[(long long)i * elem_stride]; nn |= (v != v); }
  if (nn) atomicOr(&s_nan, 1);
  __syncthreads();
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  double me = qnan, ma = qnan;
  if (!s_nan && len > 0) {
    unsigned long long klo, khi;
    auto key1 = [&](int i) { return i < len ? f64_key(p[(long long)i * elem_stride]) : kinf; };
    cta_two_middles(lenp, (long long)len, key1, &sc, &klo, &khi);
    me = (len & 1) ? key_f64(khi) : (key_f64(klo) + key_f64(khi)) / 2.0;
    auto key2 = [&](int i) { return i < len ? f64_key(fabs(p[(long long)i * elem_stride] - me)) : kinf; };
    cta_two_middles(lenp, (long long)len, key2, &sc, &klo, &khi);
    ma = MAD_CONST * ((len & 1) ? key_f64(khi) : (key_f64(klo) + key_f64(khi)) / 2.0);
  }
  if (threadIdx.x == 0) { med[blockIdx.x] = me; mad[blockIdx.x] = ma; }
}

// clamp column l of X[n][K] to med[l] +- 4 mad[l]  (R/ruvIIInb.R:373-376; NaN bounds clamp nothing)
static __global__ void k_clamp_cols(double* X, long long n, int K, const double* med, const double* mad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * K) return;
  int l = (int)(i % K);
  double hi = med[l] + 4.0 * mad[l], lo = med[l] - 4.0 * mad[l];
  double v = X[i];
  if (v > hi) v = hi;
  if (v < lo) v = lo;
  X[i] = v;
}

// ------------------------------------------------------------------------------------------------
// W helpers: group means of W (projection of t(W), R/ruvIIInb.R:331-340), RM_W, norms
// ------------------------------------------------------------------------------------------------
// one warp per (factor l, group j): mean over the group's cells
static __global__ void k_wbar(const double* W, long long Np, int K, int m, const int* group_chunk0, const int* group_n,
                       const int* group_nch, const int* chunk_valid, double* Wbar /* [m][K] */) {
  int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= K * m) return;
  int l = wid / m, j = wid % m;
  const int c0 = group_chunk0[j];
  long long p0 = (long long)c0 * CH;
  int n = group_n[j];
  double s = 0.0;
  // the group's run of chunks (with batches: one padded sub-run per batch); pads are skipped
  for (int i = lane; i < group_nch[j] * CH; i += 32) if ((i & (CH - 1)) < chunk_valid[c0 + (i >> 7)]) s += W[(long long)l * Np + p0 + i];
  s = warp_sum(s);
  if (lane == 0) Wbar[j * K + l] = s / (double)n;
}
static __global__ void k_rmw(const double* W, long long Np, int K, const int* chunk_grp, const int* chunk_valid,
                      const double* Wbar, double* RMW) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np * K) return;
  int l = (int)(i / Np);
  long long pos = i - (long long)l * Np;
  int c = (int)(pos >> 7);
  bool valid = (int)(pos & (CH - 1)) < chunk_valid[c];
  RMW[i] = valid ? W[i] - Wbar[chunk_grp[c] * K + l] : 0.0;
}
// column sum of squares of W (one CTA per factor), deterministic
static __global__ void __launch_bounds__(1024) k_w_sumsq(const double* W, long long Np, double* out) {
  __shared__ double sm[1024];
  const double* w = W + (long long)blockIdx.x * Np;
  double s = 0.0;
  for (long long i = threadIdx.x; i < Np; i += 1024) s = fma(w[i], w[i], s);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}
// W[,l] /= sqrt(colNorm) (R/ruvIIInb.R:392-394); counts non-finite entries among real cells (:401)
static __global__ void k_w_scale(double* W, long long Np, int K, const double* sumsq, const int* chunk_valid, int* bad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np * K) return;
  int l = (int)(i / Np);
  long long pos = i - (long long)l * Np;
  bool valid = (int)(pos & (CH - 1)) < chunk_valid[pos >> 7];
  if (!valid) { W[i] = 0.0; return; }
  double v = W[i] / sqrt(sumsq[l]);
  W[i] = v;
  if (!(fabs(v) <= 1.79769313486231570e308)) atomicAdd(bad, 1);
}

// ------------------------------------------------------------------------------------------------
// K_w_update: getW for every cell (R/ruvIIInb.R:380-390,732-744).  One CTA owns `cpc` cells and all
// control genes: phase 1 signed deviances -> smem, phase 2 per-cell exact median/MAD (warp per cell),
// phase 3 weighted Gram/score over control genes, phase 4 the sequential one-regressor WLS
// (= forward substitution on the lower triangle of sum w a a').
// ------------------------------------------------------------------------------------------------
struct CtlPack {            // control-gene parameters, gathered (and, when sharded, all-reduced)
  const double* __restrict__ gmean;      // [n_ctl]   old gmean
  const double* __restrict__ psi;        // [nbatch][n_ctl]
  const double* __restrict__ step;       // [n_ctl]
  const double* __restrict__ alpha_old;  // [n_ctl][K]
  const double* __restrict__ alpha_new;  // [n_ctl][K]
  const double* __restrict__ Mb;         // [n_ctl][m] old Mb
  const double* pi0;        // [n_ctl] ZINB only
  const double* psi_w;      // [nbatch][n_ctl] ZINB only
  const uint8_t* zinf;      // [Np] or nullptr
  const int* chunk_batch;   // [nchunks] batch of every chunk (all 0 unless the dispersion is batch-specific)
  int zinb;
};
#define W_FAST_DYN_SMEM (128 * 64 * 4)   // k_w_update_fast: the counts of one gene tile x the CTA's 64 cells
template <int K, bool HUBER>
__global__ void __launch_bounds__(256) k_w_update(const int32_t* const* ctl_rows, int n_ctl, long long Np,
                                                  const int* chunk_grp, const int* chunk_valid, CtlPack cp, int m,
                                                  const double* W_old, double* W_out, int cpc, double kh,
                                                  long long pos_begin, long long pos_end, const int* cell_list,
                                                  const int* cell_count) {
  constexpr int KK = K * (K + 1) / 2;
  constexpr int NACC = KK + K;
  extern __shared__ double dyn[];
  __shared__ SelScratch sc[8];
  __shared__ double sigma_c[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cell_i = tid % cpc, slice = tid / cpc, nsl = 256 / cpc;
  // cells of this CTA: a contiguous run of positions, or entries of the fast kernel's failure list
  const long long ncell = cell_list ? (long long)*cell_count : pos_end - pos_begin;
  if ((long long)blockIdx.x * cpc >= ncell) return;
  auto cell_pos = [&](int ci) -> long long {
    long long idx = (long long)blockIdx.x * cpc + ci;
    if (idx >= ncell) return pos_end;
    return cell_list ? (long long)cell_list[idx] : pos_begin + idx;
  };
  const long long pos = cell_pos(cell_i);
  const bool inb = pos < pos_end;
  const int chunk = inb ? (int)(pos >> 7) : 0;
  const bool valid = inb && (int)(pos & (CH - 1)) < chunk_valid[chunk];
  const int j = chunk_grp[chunk];
  const double* psi_b = cp.psi + (long long)cp.chunk_batch[chunk] * n_ctl;      // psi[ctl, batch of this cell]
  const double* psiw_b = cp.psi_w + (long long)cp.chunk_batch[chunk] * n_ctl;
  const int nstr = n_ctl | 1;  // odd row stride (doubles): the per-cell deviance rows do not bank-conflict
  double wv[K];
#pragma unroll
  for (int l = 0; l < K; ++l) wv[l] = valid ? W_old[(long long)l * Np + pos] : 0.0;

  if (HUBER) {
    double* dsm = dyn;  // [cpc][nstr]
    if (valid) {
      for (int gi = slice; gi < n_ctl; gi += nsl) {
        int y = ctl_rows[gi][pos];
        double lmu = cp.gmean[gi] + cp.Mb[(long long)gi * m + j];
#pragma unroll
        for (int l = 0; l < K; ++l) lmu = fma(cp.alpha_old[(long long)gi * K + l], wv[l], lmu);
        double mu = exp(lmu);
        dsm[(long long)cell_i * nstr + gi] = nb_signdev(y, lmu, mu, 1.0 / psi_b[gi]);
      }
    }
    __syncthreads();
    for (int ci = warp; ci < cpc; ci += 8) {
      long long p2 = cell_pos(ci);
      bool v2 = p2 < pos_end && (int)(p2 & (CH - 1)) < chunk_valid[p2 >> 7];
      if (v2) {  // warp-uniform
        const double* d = dsm + (long long)ci * nstr;
        auto val = [&](int i) { return d[i]; };
        auto vld = [&](int) { return true; };
        double med, mad;
        group_median_mad<32>(n_ctl, val, vld, &sc[warp], lane, &med, &mad);
        if (lane == 0) sigma_c[ci] = mad / 0.6745;
      }
    }
    __syncthreads();
  }
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
  if (valid) {
    const double khsig = HUBER ? kh * sigma_c[cell_i] : 0.0;
    for (int gi = slice; gi < n_ctl; gi += nsl) {
      int y = ctl_rows[gi][pos];
      double gm = cp.gmean[gi];
      double lmu = gm + cp.Mb[(long long)gi * m + j];
#pragma unroll
      for (int l = 0; l < K; ++l) lmu = fma(cp.alpha_old[(long long)gi * K + l], wv[l], lmu);
      double mu = exp(lmu);
      double psi = psi_b[gi];
      double s = 1.0 / (1.0 / mu + psi);
      double z = lmu + cp.step[gi] * (((double)y + 0.01) / (mu + 0.01) - 1.0) - gm;  // (Z - gmean)[ctl], :381
      double wt = s;
      if (cp.zinb && y == 0 && cp.zinf[pos]) {  // wt.zinb of a zero count in a zero-inflated cell (:262,382)
        double rw = 1.0 / psiw_b[gi], p0 = cp.pi0[gi];
        double q = exp(-rw * (log(mu + rw) - log(rw)));
        wt *= 1.0 - p0 / (p0 + (1.0 - p0) * q);
      }
      if (HUBER) wt *= huber_w(khsig, dyn[(long long)cell_i * nstr + gi]);
      int idx = 0;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double wa = wt * cp.alpha_new[(long long)gi * K + i];
        acc[KK + i] = fma(wa, z, acc[KK + i]);
#pragma unroll
        for (int jx = 0; jx <= i; ++jx) { acc[idx] = fma(wa, cp.alpha_new[(long long)gi * K + jx], acc[idx]); ++idx; }
      }
    }
  }
  // reduce over the gene slices: lanes sharing a cell inside the warp, then the 8 warps (fixed order)
  for (int o = cpc; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
  }
  __syncthreads();
  double* red = dyn;  // [8][min(cpc,32)][NACC]  (reuses the deviance buffer)
  const int cw = cpc < 32 ? cpc : 32;  // distinct cells inside one warp
  if (lane < cw) {
    int cidx = (cpc <= 32) ? lane : (tid % cpc);
#pragma unroll
    for (int i = 0; i < NACC; ++i) red[((long long)warp * cpc + cidx) * NACC + i] = acc[i];
  }
  __syncthreads();
  if (tid < cpc) {
    double T[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) T[i] = 0.0;
    if (cpc <= 32) {
      for (int w = 0; w < 8; ++w)
#pragma unroll
        for (int i = 0; i < NACC; ++i) T[i] += red[((long long)w * cpc + tid) * NACC + i];
    } else {
      // cpc in {64,128,256}: a cell lives in exactly one warp
      int w = tid >> 5;
#pragma unroll
      for (int i = 0; i < NACC; ++i) T[i] = red[((long long)w * cpc + tid) * NACC + i];
    }
    long long p2 = cell_pos(tid);
    bool v2 = p2 < pos_end && (int)(p2 & (CH - 1)) < chunk_valid[p2 >> 7];
    if (p2 < pos_end) {
      double Wn[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double s = T[KK + i];
#pragma unroll
        for (int l = 0; l < i; ++l) s -= T[i * (i + 1) / 2 + l] * Wn[l];
        Wn[i] = s / T[i * (i + 1) / 2 + i];
        W_out[(long long)i * Np + p2] = v2 ? Wn[i] : 0.0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K_w_update_fast: getW (R/ruvIIInb.R:380-390,732-744) for the cells whose per-cell Huber weights are
// provably all 1 -- with k.huber = 100 that is every non-degenerate cell.  8 cells x 4 gene slices per
// warp (a 32-byte sector of each control-gene row per load); no per-cell staging of the n_ctl deviances:
//   sweep A  max|d| over the control genes  ->  t = max|d| / (k 1.4826/0.6745), bin width max|d|/64;
//   sweep B  recomputes d, fills a 128-bin histogram on [-max, max] and accumulates the Gram / score
//            with Huber weight 1;
//   certificate: 2t < bin width, so (med - t, med + t) meets at most 2 bins (3 with rounding of edge
//            elements); if no 3 adjacent bins hold more than (n-1)/2 elements then median|d - med| >= t,
//            i.e. k sigma >= max|d| and every weight is exactly 1.
// Cells that fail are appended to fail_list and redone by the exact kernel (k_w_update with a list).
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) k_w_update_fast(const int32_t* const* ctl_rows, int n_ctl, long long Np,
                                                       const int* chunk_grp, const int* chunk_valid, CtlPack cp,
                                                       const double* ctl_tabR, const MathTabs* gtabs, int m,
                                                       const double* W_old, double* W_out, double kh,
                                                       long long pos_begin, long long pos_end, int* fail_count,
                                                       int* fail_list) {
  constexpr int KK = K * (K + 1) / 2;
  constexpr int NACC = KK + K;
  __shared__ MathTabs mt;
  __shared__ unsigned hist[8][8][64];  // [warp][cell][128 bins x 16 bit]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const double* src = reinterpret_cast<const double*>(gtabs);
    double* dst = reinterpret_cast<double*>(&mt);
    for (int i = tid; i < (int)(sizeof(MathTabs) / 8); i += 256) dst[i] = src[i];
    unsigned* h = &hist[0][0][0];
    for (int i = tid; i < 8 * 8 * 64; i += 256) h[i] = 0;
  }
  __syncthreads();
  const int cell = lane & 7, slice = lane >> 3;
  const long long pos = pos_begin + ((long long)blockIdx.x * 8 + warp) * 8 + cell;
  const bool inb = pos < pos_end;
  const int chunk = inb ? (int)(pos >> 7) : 0;
  const bool valid = inb && (int)(pos & (CH - 1)) < chunk_valid[chunk];
  // the 64 positions of a CTA lie in one 128-cell chunk, hence in ONE replicate group: Mb[g, j] is a per-gene scalar here
  const int cta_chunk = (int)(min(pos_begin + (long long)blockIdx.x * 64, pos_end - 1) >> 7);
  const int j = chunk_grp[cta_chunk];
  // ... and in ONE batch: the dispersion column and the log(y + r) tables of the control genes are per CTA, too
  const double* psi_b = cp.psi + (long long)cp.chunk_batch[cta_chunk] * n_ctl;
  const double* tabR_b = ctl_tabR + (long long)cp.chunk_batch[cta_chunk] * n_ctl * YT;
  double wv[K];
#pragma unroll
  for (int l = 0; l < K; ++l) wv[l] = valid ? W_old[(long long)l * Np + pos] : 0.0;
  // control-gene parameters are staged GT genes at a time in shared memory (they were ~12 dependent global loads per
  // element): [0] gmean + Mb[g, j], [1] psi, [2] 1/psi, [3] step, [4] gmean, [5..) alpha_old, then alpha_new
  // The counts of the same GT genes x the CTA's 64 cells travel with them: 256 contiguous bytes per gene row, fetched with cp.async
  // by the whole CTA at once (32 KB in flight) -- each lane used to load its own 4-byte count one gene ahead and waited for DRAM
  // at every element (a warp-wide load touched four 32-byte sectors of four rows).
  constexpr int GT = 128, PW = 5 + 2 * K;
  __shared__ double gp[GT][PW];
  extern __shared__ __align__(16) int ybuf_dyn[];   // [GT][64]: W_FAST_DYN_SMEM bytes of dynamic shared memory
  int (*ybuf)[64] = reinterpret_cast<int (*)[64]>(ybuf_dyn);
  const long long cta_pos0 = pos_begin + (long long)blockIdx.x * 64;   // multiple of 64 cells; cta_pos0 + 64 <= Np (Np is a multiple of 512)
  const int ycol = warp * 8 + cell;
  auto stage = [&](int g0) {
    __syncthreads();
    for (int idx = tid; idx < GT * 16; idx += 256) {
      const int t = idx >> 4, piece = idx & 15;
      if (g0 + t < n_ctl) __pipeline_memcpy_async(&ybuf[t][piece * 4], ctl_rows[g0 + t] + cta_pos0 + piece * 4, 16);
    }
    __pipeline_commit();
    for (int idx = tid; idx < GT * PW; idx += 256) {
      const int t = idx / PW, f = idx - t * PW, gi = g0 + t;
      double v = 0.0;
      if (gi < n_ctl) {
        if (f == 0) v = cp.gmean[gi] + cp.Mb[(long long)gi * m + j];
        else if (f == 1) v = psi_b[gi];
        else if (f == 2) v = 1.0 / psi_b[gi];
        else if (f == 3) v = cp.step[gi];
        else if (f == 4) v = cp.gmean[gi];
        else if (f < 5 + K) v = cp.alpha_old[(long long)gi * K + (f - 5)];
        else v = cp.alpha_new[(long long)gi * K + (f - 5 - K)];
      }
      gp[t][f] = v;
    }
    __pipeline_wait_prior(0);
    __syncthreads();
  };
  // sweep A: max half deviance
  double hmax = 0.0; int bad = 0;
  for (int g0 = 0; g0 < n_ctl; g0 += GT) {
    stage(g0);
    if (valid) {
      const int nt = min(GT, n_ctl - g0);
      for (int t = slice; t < nt; t += 4) {
        const int y = ybuf[t][ycol];
        double lmu = gp[t][0];
#pragma unroll
        for (int l = 0; l < K; ++l) lmu = fma(gp[t][5 + l], wv[l], lmu);
        const double mu = fexp(lmu, mt.ex), r = gp[t][2];
        const double h = nb_devhalf_t(y, (double)y, lmu, flog(mu + r, mt.rc, mt.lc), r, tabR_b + (long long)(g0 + t) * YT, &mt);
        hmax = fmax(hmax, fmax(h, 0.0));
        bad |= !(h == h);
      }
    }
  }
  hmax = fmax(hmax, __shfl_xor_sync(FULL, hmax, 8)); hmax = fmax(hmax, __shfl_xor_sync(FULL, hmax, 16));
  bad |= __shfl_xor_sync(FULL, bad, 8); bad |= __shfl_xor_sync(FULL, bad, 16);
  const double C = MAD_CONST / 0.6745;
  const double mx = sqrt(2.0 * hmax);
  const double t_ = mx / (kh * C) * (1.0 + 1e-9);
  const double binw = mx / 64.0;
  const bool certp = valid && !bad && mx > 0.0 && mx < 1.79769313486231570e308 && (2.0 * t_ < 0.98 * binw) && n_ctl < 65536;
  const double scale = certp ? 64.0 / mx : 0.0;
  // sweep B: histogram + Gram / score with weight sig.inv
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
  unsigned* myh = hist[warp][cell];
  for (int g0 = 0; g0 < n_ctl; g0 += GT) {
    stage(g0);
    if (valid) {
      const int nt = min(GT, n_ctl - g0);
      for (int t = slice; t < nt; t += 4) {
        const int y = ybuf[t][ycol];
        const double yd = (double)y;
        double lmu = gp[t][0];
#pragma unroll
        for (int l = 0; l < K; ++l) lmu = fma(gp[t][5 + l], wv[l], lmu);
        const double mu = fexp(lmu, mt.ex);
        if (certp) {
          const double r = gp[t][2];
          const double h = nb_devhalf_t(y, yd, lmu, flog(mu + r, mt.rc, mt.lc), r, tabR_b + (long long)(g0 + t) * YT, &mt);
          const double d = nb_signdev_from_half(h, yd, mu);
          int b = (int)floor((d + mx) * scale);
          b = b < 0 ? 0 : (b > 127 ? 127 : b);
          atomicAdd(&myh[b >> 1], (b & 1) ? 65536u : 1u);
        }
        double sI, q;
        nb_siginv_ratio(yd, mu, gp[t][1], &sI, &q);
        const double z = fma(gp[t][3], q - 1.0, lmu) - gp[t][4];  // (Z - gmean)[ctl], :381
        int idx = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const double wa = sI * gp[t][5 + K + i];
          acc[KK + i] = fma(wa, z, acc[KK + i]);
#pragma unroll
          for (int jx = 0; jx <= i; ++jx) { acc[idx] = fma(wa, gp[t][5 + K + jx], acc[idx]); ++idx; }
        }
      }
    }
  }
  __syncwarp();
  // certificate: largest 3-adjacent-bin count of this cell; slice s scans bins [32 s, 32 s + 32)
  unsigned best = 0;
  if (certp) {
    const unsigned* myh = hist[warp][cell];
    auto cnt = [&](int b) -> unsigned { if (b < 0 || b > 127) return 0u; unsigned w = myh[b >> 1]; return (b & 1) ? (w >> 16) : (w & 0xffffu); };
    for (int b = slice * 32; b < slice * 32 + 32; ++b) { unsigned v = cnt(b - 1) + cnt(b) + cnt(b + 1); best = v > best ? v : best; }
  }
  { unsigned o = __shfl_xor_sync(FULL, best, 8); best = o > best ? o : best; o = __shfl_xor_sync(FULL, best, 16); best = o > best ? o : best; }
  const bool ok = certp && (long long)best <= ((long long)n_ctl - 1) / 2;
  if (valid && !ok && slice == 0) { int p = atomicAdd(fail_count, 1); fail_list[p] = (int)pos; }
  // reduce the gene slices, solve the sequential one-regressor WLS (forward substitution)
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] += __shfl_xor_sync(FULL, acc[i], 8); acc[i] += __shfl_xor_sync(FULL, acc[i], 16); }
  if (slice == 0 && inb) {
    double Wn[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      double sN = acc[KK + i];
#pragma unroll
      for (int l = 0; l < i; ++l) sN -= acc[i * (i + 1) / 2 + l] * Wn[l];
      Wn[i] = sN / acc[i * (i + 1) / 2 + i];
      W_out[(long long)i * Np + pos] = valid ? Wn[i] : 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K_pass2: per-gene per-group S0 = sum w, S1 = sum w (Z - alpha_new . W_new) for the gmean and Mb
// updates (R/ruvIIInb.R:405-420): Z, sig.inv and the Huber weights are those of the OLD parameters.
// ------------------------------------------------------------------------------------------------
template <int K, bool HUBER>
struct OpPass2 {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = !HUBER;   // certified rows: the four cells of a lane in one branch-free block
  static constexpr bool USES_PSI = true;
  __device__ __forceinline__ void psi_changed() {}
  static constexpr int CELLS = TileCfg<K>::CELLS;
  GeneState<K> st;
  GeneRegs<K> g;
  const double* alpha_new; double an[K];
  const double* sigma; double kh, khsig; bool hub;
  double s0, s1;
  double *S0, *S1;
  int row_, lane;
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row); row_ = row; lane = threadIdx.x & 31;
#pragma unroll
    for (int l = 0; l < K; ++l) an[l] = alpha_new[(long long)row * K + l];
    hub = false;
    if (HUBER) { khsig = kh * sigma[row]; hub = !(khsig > 1.79769313486231570e308); }
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; s0 = 0.0; s1 = 0.0; }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    const double* w = tile; const double* wn = tile + K * CELLS;
    double l0, l1;
    g.lmu_pair(w, ti, l0, l1);
    const double yd0 = (double)y0, yd1 = (double)y1;
    double mu0, sa, q0, mu1, sb, q1;
    if (g.fast_range(l0, l1)) {
      nb_sq_fast(yd0, l0, g.psi, g.mt->ex, &mu0, &sa, &q0);
      nb_sq_fast(yd1, l1, g.psi, g.mt->ex, &mu1, &sb, &q1);
    } else {
      SQ a = nb_sq_slow(yd0, l0, g.psi), b = nb_sq_slow(yd1, l1, g.psi);
      mu0 = a.mu; sa = a.s; q0 = a.q; mu1 = b.mu; sb = b.s; q1 = b.q;
    }
    double res0 = fma(g.step, q0 - 1.0, l0), res1 = fma(g.step, q1 - 1.0, l1);  // Z
    double wt0 = sa, wt1 = sb;
    if (HUBER && hub) {  // warp-uniform branch
      wt0 *= huber_w_half(khsig, nb_devhalf_t(y0, yd0, l0, g.logL(mu0), g.r, g.tr, g.mt));
      wt1 *= huber_w_half(khsig, nb_devhalf_t(y1, yd1, l1, g.logL(mu1), g.r, g.tr, g.mt));
    }
    if (HUBER && g.zi) {
      const long long pos = (long long)c * CH + (ti & (CH - 1));
      if (m0) wt0 *= g.wz(y0, mu0, pos);
      if (m1) wt1 *= g.wz(y1, mu1, pos + 1);
    }
#pragma unroll
    for (int l = 0; l < K; ++l) {
      double2 v = *reinterpret_cast<const double2*>(wn + l * CELLS + ti);
      res0 = fma(-an[l], v.x, res0);
      res1 = fma(-an[l], v.y, res1);
    }
    if (!FULLCHUNK) { wt0 = m0 ? wt0 : 0.0; res0 = m0 ? res0 : 0.0; wt1 = m1 ? wt1 : 0.0; res1 = m1 ? res1 : 0.0; }
    s0 += wt0 + wt1;
    s1 = fma(wt1, res1, fma(wt0, res0, s1));
  }
  __device__ __forceinline__ void quad(int y0, int y1, int y2, int y3, int c, int ti, const double* tile) {
    const double* w = tile; const double* wn = tile + K * CELLS;
    double l0, l1, l2, l3;
    g.lmu_pair(w, ti, l0, l1);
    g.lmu_pair(w, ti + 64, l2, l3);
    const double yd0 = (double)y0, yd1 = (double)y1, yd2 = (double)y2, yd3 = (double)y3;
    double mu, sa, sb, sc, sd, q0, q1, q2, q3;
    if (g.fast_range(l0, l1) & g.fast_range(l2, l3)) {
      nb_sq_fast(yd0, l0, g.psi, g.mt->ex, &mu, &sa, &q0);
      nb_sq_fast(yd1, l1, g.psi, g.mt->ex, &mu, &sb, &q1);
      nb_sq_fast(yd2, l2, g.psi, g.mt->ex, &mu, &sc, &q2);
      nb_sq_fast(yd3, l3, g.psi, g.mt->ex, &mu, &sd, &q3);
    } else {
      SQ a = nb_sq_slow(yd0, l0, g.psi), b = nb_sq_slow(yd1, l1, g.psi), cc = nb_sq_slow(yd2, l2, g.psi), d = nb_sq_slow(yd3, l3, g.psi);
      sa = a.s; q0 = a.q; sb = b.s; q1 = b.q; sc = cc.s; q2 = cc.q; sd = d.s; q3 = d.q;
    }
    double r0 = fma(g.step, q0 - 1.0, l0), r1 = fma(g.step, q1 - 1.0, l1), r2 = fma(g.step, q2 - 1.0, l2), r3 = fma(g.step, q3 - 1.0, l3);
#pragma unroll
    for (int l = 0; l < K; ++l) {
      double2 va = *reinterpret_cast<const double2*>(wn + l * CELLS + ti), vb = *reinterpret_cast<const double2*>(wn + l * CELLS + ti + 64);
      r0 = fma(-an[l], va.x, r0); r1 = fma(-an[l], va.y, r1); r2 = fma(-an[l], vb.x, r2); r3 = fma(-an[l], vb.y, r3);
    }
    s0 += (sa + sb) + (sc + sd);
    s1 = fma(sd, r3, fma(sc, r2, fma(sb, r1, fma(sa, r0, s1))));
  }
  __device__ __forceinline__ void group_end(int j) {
    double a = warp_sum(s0), b = warp_sum(s1);
    if (lane == 0) { S0[(long long)row_ * st.m + j] = a; S1[(long long)row_ * st.m + j] = b; }
  }
  __device__ __forceinline__ void row_end(int) {}
};
template <int K, bool HUBER>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_pass2(RowGeom geo, GeneState<K> st, const double* W_new,
                                                          const double* alpha_new, const double* sigma,
                                                          double kh, double* S0, double* S1) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) { sh.wsets[0] = st.W; sh.wsets[1] = W_new; }
  const long long poff = (long long)(blockIdx.x % geo.nparts) * geo.Gl;
  OpPass2<K, HUBER> op;
  op.st = st; op.alpha_new = alpha_new; op.sigma = sigma; op.kh = kh;
  op.S0 = S0 + poff * st.m; op.S1 = S1 + poff * st.m;
  stream_rows<K, 2>(geo, &sh, op, smem);
}

// ------------------------------------------------------------------------------------------------
// K_pass2w: the group sums of the gmean / Mb update (:405-420) from the weight slab.  The element-wise sweep of this parameter set
// left every element's weight w = huber x sig.inv x wt.zinb in Wc[g][pos]; k_gram already produced S0_j = sum_{c in j} w_c and
// SWZ_j = sum w Z; what needs the NEW W is T_j = sum_{c in j} w_c W_new[c, ]  (then S1_j = SWZ_j - alpha_new . T_j): K FMAs per
// element instead of a second dot product, exp, reciprocal and working response (42 -> 5 FP64 instructions), 8 bytes read per
// element and no counts at all -- an HBM-bound sweep.
// ------------------------------------------------------------------------------------------------
template <int K, class Cfg>
struct OpPass2W {
  static constexpr int CELLS = Cfg::CELLS;
  int m;
  double T[K];
  double* TS;
  int row_, lane;
  __device__ __forceinline__ void row_begin(int row, int) { row_ = row; lane = threadIdx.x & 31; }
  __device__ __forceinline__ void group_begin(int) {
#pragma unroll
    for (int l = 0; l < K; ++l) T[l] = 0.0;
  }
  __device__ __forceinline__ void pair(int ti, const double* tile, const double* slab) {   // (the cached weight of a pad is 0)
    const double2 w = *reinterpret_cast<const double2*>(slab + ti);
#pragma unroll
    for (int l = 0; l < K; ++l) {
      const double2 v = *reinterpret_cast<const double2*>(tile + l * CELLS + ti);
      T[l] = fma(w.y, v.y, fma(w.x, v.x, T[l]));
    }
  }
  __device__ __forceinline__ void group_end(int j) {
#pragma unroll
    for (int l = 0; l < K; ++l) { double v = warp_sum(T[l]); if (lane == 0) TS[((long long)row_ * m + j) * K + l] = v; }
  }
  __device__ __forceinline__ void row_end(int) {}
};
// Measured at cfg 2: 512-cell stages 2 deep 3.55 ms, 256-cell stages 4 deep 3.84 ms, 128-cell stages 5 deep with three CTAs per SM
// 4.1 ms.  The limit is the shared-memory pipe (76 % of its wavefront peak): every pair of cells reads 16 bytes of the slab and
// K x 16 bytes of the W tile back from shared memory, 48 bytes per 8-byte element -- as much time as HBM needs for the element.
using P2wCfgA = SlabCfg<4, 2>;   // 512-cell stages, 2 deep: 104 KB at k = 5, two CTAs per SM
using P2wCfgB = SlabCfg<2, 4>;   // 256-cell stages, 4 deep: 104 KB
template <int K, class Cfg>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_pass2w(RowGeom geo, int m, const double* W_new, const double* Wc, double* TS) {
  extern __shared__ double smem[];
  const long long poff = (long long)(blockIdx.x % geo.nparts) * geo.Gl;
  OpPass2W<K, Cfg> op;
  op.m = m; op.TS = TS + poff * m * K;
  const double* const slabs[1] = {Wc};
  stream_slabs<K, 1, Cfg>(geo, W_new, slabs, op, smem);
}
// S1[g][j] = SWZ[g][j] - alpha_new[g] . TS[g][j]  (sum_c w (Z - alpha_new . W_new) regrouped)
static __global__ void k_s1_combine(int Gl, int m, int K, const double* SWZ, const double* TS, const double* alpha_new, double* S1) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Gl * m) return;
  const long long g = i / m;
  double s = SWZ[i];
  for (int l = 0; l < K; ++l) s = fma(-alpha_new[g * K + l], TS[i * K + l], s);
  S1[i] = s;
}

// gmean (R/ruvIIInb.R:405-415) and Mb (:418-420, rowWtAvg :667-675) from the group sums.
static __global__ void k_gmean_mb(int Gl, int m, const double* S0, const double* S1, const double* Mb_old, double lambda_b,
                           double* gmean_new, double* Mb_new, int* nan_flag, int* bad) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  double num = 0.0, den = 0.0;
  for (int j = 0; j < m; ++j) {
    double s0 = S0[(long long)g * m + j];
    num += S1[(long long)g * m + j] - Mb_old[(long long)g * m + j] * s0;
    den += s0;
  }
  double gm = num / den;
  gmean_new[g] = gm;
  int nb = 0, nn = 0;
  for (int j = 0; j < m; ++j) {
    double s0 = S0[(long long)g * m + j];
    double v = (S1[(long long)g * m + j] - gm * s0) / (s0 + lambda_b);
    Mb_new[(long long)g * m + j] = v;
    if (v != v) nn++;
    if (!(fabs(v) <= 1.79769313486231570e308)) nb++;
  }
  if (nn) atomicAdd(nan_flag, nn);
  if (nb) atomicAdd(bad, nb);
}
// R median of m <= 64 finite values held two per lane (x0: item lane, x1: item lane + 32), by rank counting
__device__ __forceinline__ double warp_small_median(double x0, double x1, int m, int lane) {
  int lt0 = 0, eq0 = 0, lt1 = 0, eq1 = 0;
  for (int i = 0; i < m; ++i) {
    double a = __shfl_sync(FULL, x0, i & 31), b = __shfl_sync(FULL, x1, i & 31);
    double v = i < 32 ? a : b;
    lt0 += v < x0; eq0 += v == x0; lt1 += v < x1; eq1 += v == x1;
  }
  const int kl = (m - 1) / 2, ku = m / 2;
  double lo = 0.0, hi = 0.0;   // exactly one distinct value covers each rank; sum the (identical) candidates' max
  bool has0 = lane < m, has1 = lane + 32 < m;
  double clo = -1.79769313486231570e308, chi = -1.79769313486231570e308;
  if (has0 && lt0 <= kl && kl < lt0 + eq0) clo = x0;
  if (has1 && lt1 <= kl && kl < lt1 + eq1) clo = fmax(clo, x1);
  if (has0 && lt0 <= ku && ku < lt0 + eq0) chi = x0;
  if (has1 && lt1 <= ku && ku < lt1 + eq1) chi = fmax(chi, x1);
  lo = warp_max(clo); hi = warp_max(chi);
  return (m & 1) ? hi : (lo + hi) / 2.0;
}
// any NA -> whole Mb reverts (:424); NA/Inf -> 0 (:425); each row clamped to median +- 4 mad (:428-433).
// One warp per gene.
static __global__ void __launch_bounds__(256) k_mb_finalize(int Gl, int m, const int* nan_flag, const double* Mb_old,
                                                     double* Mb_new) {
  __shared__ SelScratch sc[8];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int g = blockIdx.x * 8 + warp;
  if (g >= Gl) return;
  double* row = Mb_new + (long long)g * m;
  const bool revert = *nan_flag != 0;
  for (int j = lane; j < m; j += 32) {
    double v = revert ? Mb_old[(long long)g * m + j] : row[j];
    if (!(fabs(v) <= 1.79769313486231570e308)) v = 0.0;
    row[j] = v;
  }
  __syncwarp();
  double med, mad;
  if (m <= 64) {
    // rank by counting: each lane holds up to two values, every value is broadcast once
    double v0 = lane < m ? row[lane] : 0.0, v1 = lane + 32 < m ? row[lane + 32] : 0.0;
    med = warp_small_median(v0, v1, m, lane);
    mad = MAD_CONST * warp_small_median(fabs(v0 - med), fabs(v1 - med), m, lane);
  } else {
    auto val = [&](int i) { return row[i]; };
    auto vld = [&](int) { return true; };
    group_median_mad<32>(m, val, vld, &sc[warp], lane, &med, &mad);
  }
  double hi = med + 4.0 * mad, lo = med - 4.0 * mad;
  for (int j = lane; j < m; j += 32) {
    double v = row[j];
    if (v > hi) v = hi;
    if (v < lo) v = lo;
    row[j] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K_init_coef: get.coef.wt (R/ruvIIInb.R:181-198,704-706): WLS of log(y+1) on [1, W], weights y+1
// ------------------------------------------------------------------------------------------------
template <int K>
struct OpInit {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = false;
  static constexpr int CELLS = TileCfg<K>::CELLS;
  GeneRegs<K> g;  // only the shared-table pointers are used
  static constexpr int P = K + 1;
  static constexpr int PP = P * (P + 1) / 2;
  double XtX[PP], Xty[P];
  double *gmean, *alpha;
  __device__ __forceinline__ void row_begin(int, int) {
#pragma unroll
    for (int i = 0; i < PP; ++i) XtX[i] = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) Xty[i] = 0.0;
  }
  __device__ __forceinline__ void group_begin(int) {}
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    if (m0) elem(y0, c, ti, tile);
    if (m1) elem(y1, c, ti + 1, tile);
  }
  __device__ __forceinline__ void elem(int y, int c, int ci, const double* tile) {
    double om = (double)y + 1.0;
    double ly = (y + 1 < YT) ? g.mt->ly[y + 1] : log(om);
    double x[P];
    x[0] = 1.0;
#pragma unroll
    for (int l = 0; l < K; ++l) x[l + 1] = tile[l * CELLS + ci];
    int idx = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      double wx = om * x[i];
      Xty[i] = fma(wx, ly, Xty[i]);
#pragma unroll
      for (int j = 0; j <= i; ++j) { XtX[idx] = fma(wx, x[j], XtX[idx]); ++idx; }
    }
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int row) {
#pragma unroll
    for (int i = 0; i < PP; ++i) XtX[i] = warp_sum(XtX[i]);
#pragma unroll
    for (int i = 0; i < P; ++i) Xty[i] = warp_sum(Xty[i]);
    if ((threadIdx.x & 31) == 0) {
      double M[P][P], rhs[P];
      for (int i = 0; i < P; ++i) {
        for (int j = 0; j <= i; ++j) { M[i][j] = XtX[i * (i + 1) / 2 + j]; M[j][i] = M[i][j]; }
        rhs[i] = Xty[i];
      }
      bool ok = solve_lu<P>(P, M, rhs);
      const double qnan = __longlong_as_double(0x7ff8000000000000ll);
      gmean[row] = ok ? rhs[0] : qnan;
      for (int l = 0; l < K; ++l) alpha[(long long)row * K + l] = ok ? rhs[l + 1] : qnan;
    }
  }
};
template <int K>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_init_coef(RowGeom geo, const double* W, double* gmean, double* alpha) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = W;
  OpInit<K> op;
  op.gmean = gmean; op.alpha = alpha;
  stream_rows<K, 1>(geo, &sh, op, smem);
}
// alpha.dev = alpha - colMeans(alpha) from the RAW alpha (:187-188), then NaN/Inf alpha -> 0 (:195)
static __global__ void k_init_adev(long long n, int K, const double* colsum, double Gtot, double* alpha, double* adev) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * K) return;
  int l = (int)(i % K);
  double a = alpha[i];
  adev[i] = a - colsum[l] / Gtot;
  if (!(fabs(a) <= 1.79769313486231570e308)) alpha[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
static __global__ void k_fill(double* p, long long n, double v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// step[g] *= fac where loglik_old[g] > loglik_new[g]  (R/ruvIIInb.R:490-491)
static __global__ void k_halve(int Gl, const double* ll_old, const double* ll_new, double fac, double* step) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < Gl && ll_old[g] > ll_new[g]) step[g] *= fac;
}
// gather rows: dst[i][ncol] = src[idx[i]][ncol]  (idx < 0 -> zeros)
static __global__ void k_gather_rows(const double* src, const int* idx, int n, int ncol, double* dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * ncol) return;
  int r = (int)(i / ncol), cidx = (int)(i % ncol);
  int s = idx[r];
  dst[i] = s >= 0 ? src[(long long)s * ncol + cidx] : 0.0;
}
// permute + pad a slab of count rows: dst[r][pos] = pos2cell[pos] >= 0 ? src(r, cell) : -1
static __global__ void k_permute_rows(const int32_t* src, long long rows, long long N, long long src_row_stride,
                               long long src_cell_stride, const int* pos2cell, long long Np, int32_t* dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Np) return;
  long long r = i / Np, pos = i - r * Np;
  int cidx = pos2cell[pos];
  dst[i] = cidx >= 0 ? src[r * src_row_stride + (long long)cidx * src_cell_stride] : -1;
}
static __global__ void k_fill_i32(int32_t* p, long long n, int32_t v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// R-layout slab (cells [c0, c0+nc) x all Gh host rows, column-major) -> dst[r][cell2pos[c]].
// 32 x 32 tile through shared memory: reads coalesced along genes, writes along positions of a row
// (cells of one group are contiguous after the permutation, so mostly coalesced as well).
static __global__ void k_scatter_cellmajor(const int32_t* stage, long long Gh, const int* rowidx, long long row0, long long rows,
                                    long long c0, long long nc, const int* cell2pos, long long Np, int32_t* dst) {
  __shared__ int32_t tile[32][33];
  const long long rb = (long long)blockIdx.x * 32, cb = (long long)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    long long c = cb + i, r = rb + threadIdx.x;
    if (c < nc && r < rows) {
      long long hr = rowidx ? (long long)rowidx[r] : row0 + r;
      tile[i][threadIdx.x] = stage[c * Gh + hr];
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    long long r = rb + i, c = cb + threadIdx.x;
    if (c < nc && r < rows) dst[r * Np + cell2pos[c0 + c]] = tile[threadIdx.x][i];
  }
}
// CSR expansion: zero the real cells / -1 the pads, then scatter the stored values (warp per row).
static __global__ void k_fill_pad(int32_t* dst, long long rows, long long Np, const int* pos2cell) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Np) return;
  dst[i] = pos2cell[i % Np] >= 0 ? 0 : -1;
}
static __global__ void k_csr_scatter(const long long* indptr, const int32_t* indices, const int32_t* values, long long rows,
                              long long base, const int* cell2pos, long long Np, int32_t* dst) {
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= rows) return;
  for (long long e = indptr[w] - base + lane; e < indptr[w + 1] - base; e += 32) dst[w * Np + cell2pos[indices[e]]] = values[e];
}

// a block of columns of a dgCMatrix (compressed sparse COLUMN: p / i / x slots, column = cell) scattered into the gene-major resident
// layout: one warp per column.  x holds the counts as doubles (R's numeric); ceiling as R/ruvIIInb.R:53 applies it.
static __global__ void k_csc_scatter(const int* colptr, const int* rowidx, const double* x, long long ncols, long long c0, int base, int Gl,
                                     const int* cell2pos, long long Np, int32_t* dst, int* bad) {
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= ncols) return;
  const long long pos = cell2pos[c0 + w];
  for (int e = colptr[w] - base + lane; e < colptr[w + 1] - base; e += 32) {
    const int g = rowidx[e];
    const double v = ceil(x[e]);
    if (g < 0 || g >= Gl || !(v >= 0.0) || v > 2147483647.0) { atomicAdd(bad, 1); continue; }
    dst[(long long)g * Np + pos] = (int32_t)v;
  }
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA peak of this device (the governing roofline of the sweeps): 8 independent
// DFMA chains per thread, 1024 x 8 x 2 flops per thread per call.
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_fp64_peak(double* out, double a, double b, int iters) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = a + (double)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 16; ++rep)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], b, a);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;  // keeps the chains alive
}

// counts beyond the per-gene tables, per row: rows that hold many of them cost ~2.5x per element in the log-likelihood
// sweep; the host launches those rows FIRST (RowGeom.rowlist) so that they do not form the tail of a sweep
static __global__ void __launch_bounds__(256) k_count_big(const int32_t* Yp, int rows, long long Np, int* nbig) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int4* y4 = reinterpret_cast<const int4*>(Yp + (long long)row * Np);
  int c = 0;
  for (long long i = lane; i < Np / 4; i += 32) { int4 v = y4[i]; c += (v.x >= YT) + (v.y >= YT) + (v.z >= YT) + (v.w >= YT); }
  c = __reduce_add_sync(FULL, c);
  if (lane == 0) nbig[row] = c;
}

static __global__ void k_scale_vec(double* x, long long n, double f) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= f;
}
static __global__ void k_count_nan(const double* x, long long n, int* count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && x[i] != x[i]) atomicAdd(count, 1);
}
