// disp.cuh -- H3: estimateDisp.par (reference R/estimateDisp.par.R:3-202) with design = 1 column of ones,
// offset = alpha W' + Mb[, grp] (no gmean) -- the call sites R/ruvIIInb.R:205-221,535-565.
//
// Device: the work that is O(G N) per dispersion value --
//   k_disp_init   start value of edgeR's one-group Newton-Raphson (log(sum y e^-o / n)), rowSums(y), mean e^o
//   k_disp_nr     one Newton-Raphson step of the intercept for ND dispersion values per sweep (edgeR C++
//                 glm_one_group: step = sum (y-mu)/(1+phi mu) / sum mu/(1+phi mu), tol 1e-10, maxit 50)
//   k_disp_apl    Cox-Reid adjusted profile likelihood for ND dispersion values per sweep (edgeR compute_apl):
//                 sum l_NB(y; mu, phi) - 1/2 log sum mu/(1+phi mu); lgamma(y + 1/phi_d) comes from a 21 x 128 table
//   k_disp_dev    NB deviance at a per-gene dispersion (for the prior d.f.)
//   k_tricube     local-constant tricube smoother of the 21 columns of l0 against AveLogCPM (locfit stand-in)
// Host: spline maximisation (FMM spline + edgeR interpolator::find_max), limma fitFDist moments, the control flow.
// The CPU checker restates the same arithmetic in its `edger` module (parity unpinned: edgeR/limma/locfit are not under /root/reference).
#pragma once
#include "kernels.cuh"

#define DISP_NG 21
#define DISP_ND 7   // dispersion values per sweep (3 sweeps cover the grid)
#define DISP_MAXP 16  // row parts of a scoring sweep over few rows

template <int K, int ND>
struct DispRowShared {
  RowShared base;
  double tg[ND][YT];  // lgamma(y + r_d) - lgamma(r_d) - lgamma(y + 1)
  FmRL rl[FM_NT];     // log tables interleaved (cached APL sweep)
};

struct DispArgs {
  const double* phi_grid;   // [ND] dispersions of this sweep, or nullptr
  const double* phi_row;    // [Gl] per-gene dispersion (ND == 1), or nullptr
  double* beta;             // [Gl][ldb] intercepts; this sweep uses columns d0 .. d0+ND-1
  int ldb, d0, ds;          // this sweep's columns are d0, d0 + ds, d0 + 2 ds, ... (ds = 3: every third grid point, so that
                            // the next sweep's points start from a converged neighbour one grid step away)
  const double* emean;      // [Gl] mean_c exp(offset) (ave_log_cpm prior) or nullptr
  double prior_count;       // aveLogCPM: y + p_c, offset log(e^o + 2 p_c), p_c = prior_count e^o / mean e^o
  double tol;
  int* notconv;             // set to 1 when some |step| >= tol
  double n;                 // number of cells
  const double* gm_real;    // [Gl] gmean of the parameter set: the ZINB weights wt.zinb need mu = exp(offset + gmean)
  double* info_out;         // [Gl][ldb] or nullptr: the sweep's Fisher information sum w mu / (1 + phi mu) per column -- at
                            // convergence the Cox-Reid term of the APL (adjustedProfileLik: 0.5 log(X' W X), one column)
  const double* Ecache;     // [Gl][Np] exp(offset) of every cell, written by k_disp_init, or nullptr (no room: recompute).  The
  long long Np;             // offsets are fixed for the whole dispersion step, so its ~40 sweeps stream 8 B per element
  int nchunks;              // instead of redoing k FMAs + one exp per element
  const double* syo_row;    // [Gl] sum y offset and
  const double* sy_row;     // [Gl] sum y of every row (k_disp_init): the APL needs no offset then
  double* step_prev;        // [Gl][ldb] last step of every column (0 before the first): Fisher scoring converges linearly, so
                            // rho = |step / step_prev| predicts the next step; a column whose NEXT step would be below tol
                            // stops now -- edgeR would spend one more sweep to apply a change < tol (1e-10) and stop.
                            // Only for |step| < 1e-8: info_out is the information BEFORE the step is applied
  int Gl = 0;               // rows of the shard (stride of the part buffers)
  int nparts = 1;           // row parts of this launch (cached sweeps over few rows); > 1: the sweep leaves partial S1, S2 per part in
  double *part_s1 = nullptr, *part_s2 = nullptr;  // [nparts][Gl][ND] and k_disp_nr_finish applies the Newton step
  int* done;                // [Gl] rows whose ND intercepts have all converged: later sweeps skip them (edgeR stops each
                            // fit at |step| < tol; rows converge after 4-8 steps, a few stragglers need 15-20)
};

// the row's slice of the exp(offset) cache, fetched one chunk ahead like the counts (stream_rows): a lane's two pairs of
// chunk c sit at c CH + 2 lane and c CH + 64 + 2 lane
struct ERow {
  const double* base; double2 na, nb, hold; int pre_c, nchunks;
  __device__ __forceinline__ void begin(const double* E, long long row, long long Np, int nch) {
    base = E ? E + row * Np : nullptr; pre_c = -1; nchunks = nch;
  }
  __device__ __forceinline__ double2 get(int c, int ti) {
    if (ti & 64) return hold;
    const int lane2 = 2 * (threadIdx.x & 31);
    double2 cur;
    if (pre_c == c) { cur = na; hold = nb; }
    else {
      cur = __ldcs(reinterpret_cast<const double2*>(base + (long long)c * CH + lane2));
      hold = __ldcs(reinterpret_cast<const double2*>(base + (long long)c * CH + 64 + lane2));
    }
    if (c + 1 < nchunks) {
      na = __ldcs(reinterpret_cast<const double2*>(base + (long long)(c + 1) * CH + lane2));
      nb = __ldcs(reinterpret_cast<const double2*>(base + (long long)(c + 1) * CH + 64 + lane2));
      pre_c = c + 1;
    }
    return cur;
  }
};

// ---- start values ---------------------------------------------------------------------------------------------
template <int K>
struct OpDispInit {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = false;
  GeneState<K> st; GeneRegs<K> g;
  double s_ye, s_y, s_e, s_w, s_yo; int nz;
  double *beta0, *rowsum, *emean, *syo_row; double n; const double* gm_real; double gmr;
  double* Ecache; long long Np; double* erow;
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row); s_ye = s_y = s_e = s_w = s_yo = 0.0; nz = 0; gmr = gm_real[row];
    erow = Ecache ? Ecache + (long long)row * Np : nullptr;
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; }
  __device__ __forceinline__ double one(int y, double off, bool m, long long pos) {
    if (!m) return 0.0;
    double E = g.expm(off);
    double yd = (double)y;
    double w = g.zi ? g.wz(y, E * exp(gmr), pos) : 1.0;   // weights = wt.zinb (R/ruvIIInb.R:211)
    s_e += E; s_y += yd; s_w += w; s_yo = fma(yd, off, s_yo);
    if (y > 0) { s_ye += yd / E * w; nz = 1; }
    return E;
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    double o0, o1;
    g.lmu_pair(tile, ti, o0, o1);
    const long long pos = (long long)c * CH + (ti & (CH - 1));
    double2 e;
    e.x = one(y0, o0, m0, pos); e.y = one(y1, o1, m1, pos + 1);
    if (erow) __stcs(reinterpret_cast<double2*>(erow + pos), e);
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int row) {
    double a = warp_sum(s_ye), b = warp_sum(s_y), e = warp_sum(s_e), tw = warp_sum(s_w), yo = warp_sum(s_yo);
    int any = __any_sync(FULL, nz);
    if ((threadIdx.x & 31) == 0) {
      beta0[row] = any ? log(a / tw) : -__longlong_as_double(0x7ff0000000000000ll);   // glm_one_group start
      rowsum[row] = b; emean[row] = e / n; syo_row[row] = yo;
    }
  }
};
template <int K>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_disp_init(RowGeom geo, GeneState<K> st, double* beta0, double* rowsum,
                                                              double* emean, double n, const double* gm_real, double* syo_row,
                                                              double* Ecache) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = st.W;
  OpDispInit<K> op;
  op.st = st; op.beta0 = beta0; op.rowsum = rowsum; op.emean = emean; op.n = n; op.gm_real = gm_real;
  op.syo_row = syo_row; op.Ecache = Ecache; op.Np = geo.Np;
  stream_rows<K, 1>(geo, &sh, op, smem);
}

// the Newton step of one (row, dispersion) from S1 = sum w y / (1 + phi mu), S2 = sum w E / (1 + phi mu); returns "not converged"
__device__ __forceinline__ bool disp_nr_update(const DispArgs& a, int row, int d, double s1, double s2, double eb) {
  s2 *= eb;
  double step = (s1 - s2) / s2;
  if (a.info_out) a.info_out[(long long)row * a.ldb + a.d0 + d * a.ds] = s2;
  double* bp = a.beta + (long long)row * a.ldb + a.d0 + d * a.ds;
  if (!(step == step)) return false;
  *bp += step;
  bool conv = fabs(step) < a.tol;
  if (a.step_prev) {
    double* sp = a.step_prev + (long long)row * a.ldb + a.d0 + d * a.ds;
    const double prev = *sp;
    *sp = step;
    if (!conv && prev != 0.0 && fabs(step) < 1e-8) { double rho = fabs(step / prev); conv = rho < 0.5 && rho * fabs(step) < a.tol; }
  }
  return !conv;
}

// ---- one Newton-Raphson step for ND dispersions ------------------------------------------------------------------
template <int K, int ND, bool CACHED = false>
struct OpDispNR {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = false;
  GeneState<K> st; GeneRegs<K> g; DispArgs a;
  // score = sum w (y - mu) / (1 + phi mu) = S1 - e^beta S2, information = e^beta S2, with S1 = sum w y / (1 + phi mu) and
  // S2 = sum w E / (1 + phi mu), mu = e^beta E: per cell and dispersion one FMA for the denominator, the reciprocal, two FMAs
  double eb[ND], ebphi[ND], dl[ND], info[ND];
  double yadd, escale, egm; bool live; ERow er;
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row);
    {   // the exp(offset) row must not be prefetched beyond this CTA's part of the row
      const int per = (a.nchunks + a.nparts - 1) / a.nparts;
      er.begin(a.Ecache, row, a.Np, min(a.nchunks, min(a.nchunks, (int)(blockIdx.x % a.nparts) * per) + per));
    }
    live = false;
    egm = g.zi ? exp(a.gm_real[row]) : 1.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      double b = a.beta[(long long)row * a.ldb + a.d0 + d * a.ds];
      eb[d] = exp(b);
      live |= (b > -1.79769313486231570e308) & (b < 1.79769313486231570e308);
      ebphi[d] = eb[d] * (a.phi_row ? a.phi_row[row] : a.phi_grid[d * a.ds]);
      dl[d] = 0.0; info[d] = 0.0;
    }
    if (a.done && a.done[row]) live = false;
    yadd = 0.0; escale = 1.0;
    if (a.emean) { double pcm = a.prior_count / a.emean[row]; yadd = pcm; escale = fma(2.0, pcm, 1.0); }
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; }
  __device__ __forceinline__ void one(int y, double E, bool m, long long pos) {
    if (!m) return;
    const double wzv = g.zi ? g.wz(y, E * egm, pos) : 1.0;
    double yd = fma(yadd, E, (double)y);
    E *= escale;
    if (g.zi) {
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        double w = wzv * frcp_nr(fma(E, ebphi[d], 1.0));
        dl[d] = fma(yd, w, dl[d]);
        info[d] = fma(E, w, info[d]);
      }
    } else {
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        double w = frcp_nr(fma(E, ebphi[d], 1.0));   // 1 + phi mu >= 1: inside the seed's range
        dl[d] = fma(yd, w, dl[d]);
        info[d] = fma(E, w, info[d]);
      }
    }
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    if (!live) return;  // all-zero gene: beta = -Inf stays (edgeR returns R_NegInf)
    const long long pos = (long long)c * CH + (ti & (CH - 1));
    if constexpr (CACHED) {
      const double2 e = er.get(c, ti);
      one(y0, e.x, m0, pos); one(y1, e.y, m1, pos + 1);
    } else if (er.base) {
      const double2 e = er.get(c, ti);
      one(y0, e.x, m0, pos); one(y1, e.y, m1, pos + 1);
    } else {
      double o0, o1;
      g.lmu_pair(tile, ti, o0, o1);
      one(y0, m0 ? g.expm(o0) : 0.0, m0, pos); one(y1, m1 ? g.expm(o1) : 0.0, m1, pos + 1);
    }
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int row) {
    if (!live) {   // nothing to iterate on: out of the later sweeps' row lists
      if (a.done && (threadIdx.x & 31) == 0) a.done[row] = 1;
      return;
    }
    if (a.nparts > 1) {   // partial sums of this part; the step is applied by k_disp_nr_finish
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        double s1 = warp_sum(dl[d]), s2 = warp_sum(info[d]);
        if ((threadIdx.x & 31) == 0) {
          const long long q = ((long long)(blockIdx.x % a.nparts) * a.Gl + row) * ND + d;
          a.part_s1[q] = s1; a.part_s2[q] = s2;
        }
      }
      return;
    }
    bool nc = false;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      double s1 = warp_sum(dl[d]), s2 = warp_sum(info[d]);
      if ((threadIdx.x & 31) == 0) nc |= disp_nr_update(a, row, d, s1, s2, eb[d]);
    }
    nc = __shfl_sync(FULL, nc, 0);
    if (nc) { if ((threadIdx.x & 31) == 0) atomicExch(a.notconv, 1); }
    else if (a.done && (threadIdx.x & 31) == 0) a.done[row] = 1;
  }
};
template <int K, int ND>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_disp_nr(RowGeom geo, GeneState<K> st, DispArgs a) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = st.W;
  OpDispNR<K, ND> op;
  op.st = st; op.a = a;
  stream_rows<K, 1>(geo, &sh, op, smem);
}

// the same sweep on the exp(offset) cache (a.Ecache != nullptr): nothing is taken from the W tiles
// (one dispersion per row -- the AveLogCPM and prior-d.f. fits -- is a pure stream of the exp(offset) slab: three CTAs per SM keep
// more of it in flight)
template <int K, int ND>
__global__ void __launch_bounds__(ROW_THREADS, (ND == 1 ? 3 : 2)) k_disp_nr_cached(RowGeom geo, GeneState<K> st, DispArgs a) {
  __shared__ RowShared sh;
  OpDispNR<K, ND, true> op;
  op.st = st; op.a = a;
  stream_rows_lean(geo, &sh, op);
}

// the Newton steps of a sweep launched with row parts: sums of the parts in a fixed order, then what the sweep's row end does
template <int ND>
static __global__ void k_disp_nr_finish(int nrows, const int* rowlist, DispArgs a) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nrows) return;
  const int row = rowlist ? rowlist[slot] : slot;
  if (a.done[row]) return;   // converged before this sweep, or nothing to iterate on (the sweep marked it)
  bool nc = false;
  for (int d = 0; d < ND; ++d) {
    double s1 = 0.0, s2 = 0.0;
    for (int p = 0; p < a.nparts; ++p) { const long long q = ((long long)p * a.Gl + row) * ND + d; s1 += a.part_s1[q]; s2 += a.part_s2[q]; }
    nc |= disp_nr_update(a, row, d, s1, s2, exp(a.beta[(long long)row * a.ldb + a.d0 + d * a.ds]));
  }
  if (nc) atomicExch(a.notconv, 1); else a.done[row] = 1;
}

// ---- adjusted profile likelihood for ND dispersions ---------------------------------------------------------------
template <int K, int ND, bool CACHED = false>
struct OpDispAPL {
  // log-lik of the row at dispersion 1/r and intercept beta:
  //   sum_c w_c [ lgamma(y + r) - lgamma(r) - lgamma(y + 1) + y log mu + r log r - (y + r) log(mu + r) ]
  // the lgamma term depends on the cell only through its count and w = 1 wherever y > 0 (wt.zinb only down-weights zeros),
  // so for y < YT it is a dot product of the row's count histogram with the grid's table (row end); y log mu =
  // y (beta + offset) is assembled at the row end too.  What is left per cell and dispersion is ONE logarithm.
  // The Cox-Reid term 0.5 log(sum w mu / (1 + phi mu)) is the information the converged NR sweep left in a.info_out.
  static constexpr bool PADS = false;
  static constexpr bool QUAD = false;
  GeneState<K> st; GeneRegs<K> g; DispArgs a;
  const double (*tg)[YT];   // shared: [ND][YT]
  const double* lgr_grid;   // [ND] lgamma(r_d)
  const int* hist;          // [Gl][YT] cells of the row with count y (0 < y < YT)
  double eb[ND], r[ND], ll[ND];
  double syo, sy, sw, egm;  // sum w y offset, sum w y, sum w
  double* l0; int ld0; ERow er;
  const FmRL* rl; bool rowok;   // cached sweep: every e^beta_d E + r_d of the row is a normal number while E < 1e280
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row);
    er.begin(a.Ecache, row, a.Np, a.nchunks);
    egm = g.zi ? exp(a.gm_real[row]) : 1.0; sw = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      eb[d] = exp(fmax(a.beta[(long long)row * a.ldb + a.d0 + d * a.ds], -1e8));  // mglmOneWay: pmax(beta, -1e8)
      r[d] = 1.0 / a.phi_grid[d * a.ds];
      ll[d] = 0.0;
    }
    rowok = true;
#pragma unroll
    for (int d = 0; d < ND; ++d) rowok &= (eb[d] < 1e20) & (r[d] > 1e-290) & (r[d] < 1e290);
    syo = 0.0; sy = 0.0;
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; }
  __device__ __forceinline__ void one(int y, double E, bool m, long long pos) {
    if (!m) return;
    double yd = (double)y;
    if (g.zi) {
      const double wzv = g.wz(y, E * egm, pos);
      sw += wzv;
      if (CACHED && rowok && E < 1e280) {
#pragma unroll
        for (int d = 0; d < ND; ++d) ll[d] = fma(-(yd + r[d]) * wzv, flog_nc2(fma(eb[d], E, r[d]), rl), ll[d]);
      } else {
#pragma unroll
        for (int d = 0; d < ND; ++d) ll[d] = fma(-(yd + r[d]) * wzv, flog(fma(eb[d], E, r[d]), g.mt->rc, g.mt->lc), ll[d]);
      }
    } else if (CACHED && rowok && E < 1e280) {   // no range checks inside, one table load per logarithm
#pragma unroll
      for (int d = 0; d < ND; ++d) ll[d] = fma(-(yd + r[d]), flog_nc2(fma(eb[d], E, r[d]), rl), ll[d]);
    } else {
#pragma unroll
      for (int d = 0; d < ND; ++d) ll[d] = fma(-(yd + r[d]), flog(fma(eb[d], E, r[d]), g.mt->rc, g.mt->lc), ll[d]);
    }
    if (y >= YT) {   // beyond the tables (rare): Stirling
#pragma unroll
      for (int d = 0; d < ND; ++d) ll[d] += nb_big_lgterm(yd, r[d], lgr_grid[d * a.ds], g.mt);
    }
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    const long long pos = (long long)c * CH + (ti & (CH - 1));
    if constexpr (CACHED) {
      const double2 e = er.get(c, ti);
      one(y0, e.x, m0, pos); one(y1, e.y, m1, pos + 1);
    } else if (er.base) {
      const double2 e = er.get(c, ti);
      one(y0, e.x, m0, pos); one(y1, e.y, m1, pos + 1);
    } else {
      double o0, o1;
      g.lmu_pair(tile, ti, o0, o1);
      if (m0) { syo = fma((double)y0, o0, syo); sy += (double)y0; }
      if (m1) { syo = fma((double)y1, o1, syo); sy += (double)y1; }
      one(y0, m0 ? g.expm(o0) : 0.0, m0, pos); one(y1, m1 ? g.expm(o1) : 0.0, m1, pos + 1);
    }
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int row) {
    double tyo = er.base ? a.syo_row[row] : warp_sum(syo), ty = er.base ? a.sy_row[row] : warp_sum(sy), tw = g.zi ? warp_sum(sw) : a.n;
    const int lane = threadIdx.x & 31;
    double hy[YT / 32];
#pragma unroll
    for (int q = 0; q < YT / 32; ++q) hy[q] = (double)hist[(long long)row * YT + q * 32 + lane];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      double t = ll[d];
#pragma unroll
      for (int q = 0; q < YT / 32; ++q) t = fma(hy[q], tg[d][q * 32 + lane], t);   // tg[d][0] = 0
      double s1 = warp_sum(t);
      if (lane == 0) {
        const double low = 1e-10;  // edgeR adj_coxreid low_value
        double s2 = a.info_out[(long long)row * a.ldb + a.d0 + d * a.ds];
        double adj = 0.5 * log(s2 < low ? low : s2);
        double beta = fmax(a.beta[(long long)row * a.ldb + a.d0 + d * a.ds], -1e8);
        double ylogmu = ty > 0.0 ? fma(beta, ty, tyo) : 0.0;
        l0[(long long)row * ld0 + a.d0 + d * a.ds] = s1 + ylogmu + tw * r[d] * log(r[d]) - adj;   // tw = n when all weights are 1
      }
    }
  }
};
template <int K, int ND>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_disp_apl(RowGeom geo, GeneState<K> st, DispArgs a, const double* tabG /* [ND][YT] */,
                                                             const double* lgr_grid, const int* hist, double* l0, int ld0) {
  extern __shared__ double smem[];
  __shared__ DispRowShared<K, ND> sh;
  if (threadIdx.x == 0) sh.base.wsets[0] = st.W;
  for (int i = threadIdx.x; i < ND * YT; i += ROW_THREADS) (&sh.tg[0][0])[i] = tabG[(long long)(i / YT) * a.ds * YT + (i % YT)];
  OpDispAPL<K, ND> op;
  op.st = st; op.a = a; op.tg = sh.tg; op.lgr_grid = lgr_grid; op.hist = hist; op.l0 = l0; op.ld0 = ld0; op.rl = sh.rl;
  stream_rows<K, 1>(geo, &sh.base, op, smem);
}

template <int K, int ND>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_disp_apl_cached(RowGeom geo, GeneState<K> st, DispArgs a, const double* tabG, const double* lgr_grid,
                                                                    const int* hist, double* l0, int ld0) {
  __shared__ DispRowShared<K, ND> sh;
  for (int i = threadIdx.x; i < ND * YT; i += ROW_THREADS) (&sh.tg[0][0])[i] = tabG[(long long)(i / YT) * a.ds * YT + (i % YT)];
  for (int i = threadIdx.x; i < FM_NT; i += ROW_THREADS) { sh.rl[i].rc = geo.mtabs->rc[i]; sh.rl[i].lc = geo.mtabs->lc[i]; }
  OpDispAPL<K, ND, true> op;
  op.st = st; op.a = a; op.tg = sh.tg; op.lgr_grid = lgr_grid; op.hist = hist; op.l0 = l0; op.ld0 = ld0; op.rl = sh.rl;
  stream_rows_lean(geo, &sh.base, op);
}

// counts of each value 1 .. YT-1 in every row (pads hold -1): one CTA per row, shared-memory histogram.  Depends on the
// counts only, so it is built once per upload / Winsorize.
static __global__ void __launch_bounds__(256) k_row_hist(const int32_t* __restrict__ Yp, long long Np, int* __restrict__ hist) {
  __shared__ int h[YT];
  const long long row = blockIdx.x;
  for (int i = threadIdx.x; i < YT; i += 256) h[i] = 0;
  __syncthreads();
  const int4* p = reinterpret_cast<const int4*>(Yp + row * Np);   // Np is a multiple of CH = 128
  for (long long i = threadIdx.x; i < Np / 4; i += 256) {
    int4 v = p[i];
    if ((unsigned)(v.x - 1) < (unsigned)(YT - 1)) atomicAdd(&h[v.x], 1);
    if ((unsigned)(v.y - 1) < (unsigned)(YT - 1)) atomicAdd(&h[v.y], 1);
    if ((unsigned)(v.z - 1) < (unsigned)(YT - 1)) atomicAdd(&h[v.z], 1);
    if ((unsigned)(v.w - 1) < (unsigned)(YT - 1)) atomicAdd(&h[v.w], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < YT; i += 256) hist[row * YT + i] = h[i];
}

// rows with flag == 0 -> compact ascending list (one CTA, ordered scan): the NR sweeps after the bulk has converged
static __global__ void __launch_bounds__(1024) k_list_zero(int n, const int* flag, int* list, int* count) {
  __shared__ int wsum[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int g0 = 0; g0 < n; g0 += 1024) {
    int g = g0 + threadIdx.x;
    bool f = g < n && flag[g] == 0;
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
  if (threadIdx.x == 0) *count = base;
}

// ---- NB deviance at a per-gene dispersion and intercept (prior d.f.) ------------------------------------------------
template <int K>
struct OpDispDev {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = false;
  GeneState<K> st; GeneRegs<K> g;
  const double *beta, *phi_row, *gm_real; double* dev;
  double b, eb, r, acc, egm;
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row); b = fmax(beta[row], -1e8); eb = exp(b); r = 1.0 / phi_row[row]; acc = 0.0;
    egm = g.zi ? exp(gm_real[row]) : 1.0;
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; }
  __device__ __forceinline__ void one(int y, double off, bool m, long long pos) {
    if (!m) return;
    double E = g.expm(off);
    double mu = eb * E;
    double yd = (double)y;
    const double wzv = g.zi ? g.wz(y, E * egm, pos) : 1.0;
    // unit deviance / 2 = y log(y/mu) - (y + r) log((y + r)/(mu + r)); the tables of the row were built for phi_row
    acc = fma(wzv, nb_devhalf_t(y, yd, b + off, flog(mu + r, g.mt->rc, g.mt->lc), r, g.tr, g.mt), acc);
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    double o0, o1;
    g.lmu_pair(tile, ti, o0, o1);
    const long long pos = (long long)c * CH + (ti & (CH - 1));
    one(y0, o0, m0, pos); one(y1, o1, m1, pos + 1);
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int row) {
    double s = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) dev[row] = 2.0 * s;
  }
};
template <int K>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_disp_dev(RowGeom geo, GeneState<K> st, const double* beta, const double* phi_row,
                                                             const double* gm_real, double* dev) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = st.W;
  OpDispDev<K> op;
  op.st = st; op.beta = beta; op.phi_row = phi_row; op.dev = dev; op.gm_real = gm_real;
  stream_rows<K, 1>(geo, &sh, op, smem);
}

// ---- local-constant tricube smoother (locfit deg = 0 stand-in) -------------------------------------------------------
// xs: covariate sorted ascending (n), perm: row of l0 for each sorted position; one CTA per sorted position i.  The q
// nearest neighbours of xs[i] are a contiguous window of the sorted order; bandwidth h = distance to the q-th nearest.
// idx (nullable): the sorted positions this launch evaluates (a gene shard computes the trend of its own genes only), else all n.
static __global__ void __launch_bounds__(128) k_tricube(const double* xs, const int* perm, int n, int q, const double* l0, int ld0,
                                                 double* m0 /* [n][DISP_NG] in sorted order */, const int* idx) {
  __shared__ double red[4][DISP_NG + 1];
  __shared__ int s_lo;
  __shared__ double s_h;
  const int i = idx ? idx[blockIdx.x] : (int)blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    const double xi = xs[i];
    // window [lo, lo+q) with lo in [max(0, i-q+1), min(i, n-q)] minimising max(xi - xs[lo], xs[lo+q-1] - xi)
    int a = max(0, i - q + 1), b = min(i, n - q);
    while (a < b) {  // the left distance decreases and the right distance increases with lo
      int mid = (a + b) >> 1;
      if (xi - xs[mid] > xs[mid + q - 1] - xi) a = mid + 1; else b = mid;
    }
    int lo = a;
    if (lo > max(0, i - q + 1)) {  // the optimum is at lo or lo-1
      double ca = fmax(xi - xs[lo], xs[lo + q - 1] - xi), cb = fmax(xi - xs[lo - 1], xs[lo + q - 2] - xi);
      if (cb < ca) lo = lo - 1;
    }
    s_lo = lo;
    s_h = fmax(xi - xs[lo], xs[lo + q - 1] - xi);
  }
  __syncthreads();
  const int lo = s_lo; const double h = s_h, xi = xs[i];
  double acc[DISP_NG + 1];
#pragma unroll
  for (int d = 0; d <= DISP_NG; ++d) acc[d] = 0.0;
  // points outside the window may tie at distance h (weight 0) -- only strictly closer points have weight
  for (int jj = lo + tid; jj < lo + q; jj += 128) {
    double dist = fabs(xs[jj] - xi);
    double w;
    if (h > 0.0) { double u = fmin(dist / h, 1.0); double v = 1.0 - u * u * u; w = v * v * v; }
    else w = 1.0;
    const double* row = l0 + (long long)perm[jj] * ld0;
#pragma unroll
    for (int d = 0; d < DISP_NG; ++d) acc[d] = fma(w, row[d], acc[d]);
    acc[DISP_NG] += w;
  }
#pragma unroll
  for (int d = 0; d <= DISP_NG; ++d) { double v = warp_sum(acc[d]); if (lane == 0) red[warp][d] = v; }
  __syncthreads();
  if (tid <= DISP_NG) {
    double v = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
    red[0][tid] = v;
  }
  __syncthreads();
  if (tid < DISP_NG) m0[(long long)i * DISP_NG + tid] = red[0][tid] / red[0][DISP_NG];
}

// ---- host: FMM spline + edgeR interpolator::find_max --------------------------------------------------------------------
static void fmm_spline_h(int n, const double* x, const double* y, double* b, double* c, double* d) {
  if (n < 3) { double t = y[1] - y[0]; b[0] = t / (x[1] - x[0]); b[1] = b[0]; c[0] = c[1] = d[0] = d[1] = 0.0; return; }
  const int nm1 = n - 1;
  d[0] = x[1] - x[0];
  c[1] = (y[1] - y[0]) / d[0];
  for (int i = 1; i < nm1; ++i) {
    d[i] = x[i + 1] - x[i];
    b[i] = 2.0 * (d[i - 1] + d[i]);
    c[i + 1] = (y[i + 1] - y[i]) / d[i];
    c[i] = c[i + 1] - c[i];
  }
  b[0] = -d[0]; b[nm1] = -d[nm1 - 1];
  c[0] = c[nm1] = 0.0;
  if (n > 3) {
    c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0]);
    c[nm1] = c[nm1 - 1] / (x[nm1] - x[nm1 - 2]) - c[nm1 - 2] / (x[nm1 - 1] - x[nm1 - 3]);
    c[0] = c[0] * d[0] * d[0] / (x[3] - x[0]);
    c[nm1] = -c[nm1] * d[nm1 - 1] * d[nm1 - 1] / (x[nm1] - x[nm1 - 3]);
  }
  for (int i = 1; i < n; ++i) { double t = d[i - 1] / b[i - 1]; b[i] -= t * d[i - 1]; c[i] -= t * c[i - 1]; }
  c[nm1] = c[nm1] / b[nm1];
  for (int i = nm1 - 1; i >= 0; --i) c[i] = (c[i] - d[i] * c[i + 1]) / b[i];
  b[nm1] = (y[nm1] - y[nm1 - 1]) / d[nm1 - 1] + d[nm1 - 1] * (c[nm1 - 1] + 2.0 * c[nm1]);
  for (int i = 0; i < nm1; ++i) {
    b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i]);
    d[i] = (c[i + 1] - c[i]) / d[i];
    c[i] = 3.0 * c[i];
  }
  c[nm1] = 3.0 * c[nm1];
  d[nm1] = d[nm1 - 1];
}
static bool quad_first_root(double a, double b, double c, double* sol) {
  double back = b * b - 4.0 * a * c;
  if (!(back >= 0.0) || a == 0.0) return false;
  *sol = -b / (2.0 * a) - sqrt(back) / (2.0 * a);
  return true;
}
static double maximize_interpolant_h(int n, const double* x, const double* y) {
  int at = 0;
  for (int i = 1; i < n; ++i) if (y[i] > y[at]) at = i;
  double maxed = y[at], xmax = x[at];
  double b[DISP_NG], c[DISP_NG], d[DISP_NG];
  fmm_spline_h(n, x, y, b, c, d);
  double s;
  if (at > 0 && quad_first_root(3 * d[at - 1], 2 * c[at - 1], b[at - 1], &s) && s > 0 && s < x[at] - x[at - 1]) {
    double t = ((d[at - 1] * s + c[at - 1]) * s + b[at - 1]) * s + y[at - 1];
    if (t > maxed) { maxed = t; xmax = s + x[at - 1]; }
  }
  if (at < n - 1 && quad_first_root(3 * d[at], 2 * c[at], b[at], &s) && s > 0 && s < x[at + 1] - x[at]) {
    double t = ((d[at] * s + c[at]) * s + b[at]) * s + y[at];
    if (t > maxed) { maxed = t; xmax = s + x[at]; }
  }
  return xmax;
}

// ---- host: digamma / trigamma / limma fitFDist ------------------------------------------------------------------------------
static double digamma_h(double x) {
  double r = 0.0;
  while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132)))));
}
static double trigamma_h(double x) {
  double r = 0.0;
  while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
  double f = 1.0 / (x * x);
  return r + 1.0 / x + f / 2 + f / x * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f * (1.0 / 30 - f * (5.0 / 66)))));
}
static double tetragamma_h(double x) {  // psi''(x)
  double r = 0.0;
  while (x < 6.0) { r -= 2.0 / (x * x * x); x += 1.0; }
  double f = 1.0 / (x * x);
  return r - f - f / x - f * f * (0.5 - f * (1.0 / 6 - f * (1.0 / 6 - f * (3.0 / 10))));
}
static double trigamma_inverse_h(double x) {  // limma::trigammaInverse
  if (x > 1e7) return 1.0 / sqrt(x);
  if (x < 1e-6) return 1.0 / x;
  double y = 0.5 + 1.0 / x;
  for (int it = 0; it < 50; ++it) {
    double tri = trigamma_h(y);
    double dif = tri * (1.0 - tri / x) / tetragamma_h(y);
    y += dif;
    if (-dif / y < 1e-8) break;
  }
  return y;
}
// limma::fitFDist(x, df1) without covariate -> df2 (Inf when the log-variances are no more variable than chi-square noise)
static double fit_f_dist_df2(const std::vector<double>& s2, double df1) {
  std::vector<double> x;
  for (double v : s2) if (std::isfinite(v) && v > -1e-15) x.push_back(std::max(v, 0.0));
  const size_t n = x.size();
  if (n == 0) return NAN;
  if (n == 1) return 0.0;
  std::vector<double> srt(x);
  std::sort(srt.begin(), srt.end());
  double m = (n & 1) ? srt[n / 2] : 0.5 * (srt[n / 2 - 1] + srt[n / 2]);
  if (m == 0.0) m = 1.0;
  double emean = 0.0;
  std::vector<double> e(n);
  const double cst = -digamma_h(df1 / 2.0) + log(df1 / 2.0);
  for (size_t i = 0; i < n; ++i) { e[i] = log(std::max(x[i], 1e-5 * m)) + cst; emean += e[i]; }
  emean /= (double)n;
  double evar = 0.0;
  for (size_t i = 0; i < n; ++i) evar += (e[i] - emean) * (e[i] - emean);
  evar = evar / (double)(n - 1) - trigamma_h(df1 / 2.0);
  if (evar > 0.0) return 2.0 * trigamma_inverse_h(evar);
  return INFINITY;
}

// ---- the north_star variant: per-gene Newton on log psi with digamma / trigamma differences ---------------------------
// NOT what the reference computes (it interpolates the 21-point APL grid and shrinks, see above).  Maximises over
// theta = log psi_g the NB log-likelihood of the gene at the CURRENT fitted means mu_gc = exp(gmean + Mb + alpha W):
//   dl/dr   = sum_c [ psi0(y + r) - psi0(r) + log r + 1 - log(mu + r) - (y + r)/(mu + r) ],
//   d2l/dr2 = sum_c [ psi1(y + r) - psi1(r) + 1/r - 2/(mu + r) + (y + r)/(mu + r)^2 ],       r = exp(-theta),
// with psi0(y + r) - psi0(r) = sum_{j<y} 1/(r + j) and psi1(y + r) - psi1(r) = -sum_{j<y} 1/(r + j)^2 tabulated per gene
// and per iteration for y < YT by a warp scan (asymptotic series beyond the table).  One sweep per Newton iteration.
__device__ __forceinline__ double digamma_d(double x) {
  double r = 0.0;
  while (x < 8.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132)))));
}
__device__ __forceinline__ double trigamma_d(double x) {
  double r = 0.0;
  while (x < 8.0) { r += 1.0 / (x * x); x += 1.0; }
  double f = 1.0 / (x * x);
  return r + 1.0 / x + f / 2 + f / x * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f * (1.0 / 30 - f * (5.0 / 66)))));
}
template <int K>
struct OpPsiNewton {
  static constexpr bool PADS = false;
  static constexpr bool QUAD = false;
  GeneState<K> st; GeneRegs<K> g;
  double *theta; int* notconv; double tol;
  double r, lr, ir, dg_r, tg_r, g1, g2; bool live;
  double *t0, *t1;   // this warp's tables (shared memory)
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row);
    const int lane = threadIdx.x & 31;
    double th = theta[row];
    live = (th == th);
    r = exp(-th); lr = -th; ir = 1.0 / r;
    dg_r = digamma_d(r); tg_r = trigamma_d(r);
    g1 = 0.0; g2 = 0.0;
    t0 = const_cast<double*>(g.tl); t1 = const_cast<double*>(g.tr);
    // exclusive prefix sums of 1/(r + j) and -1/(r + j)^2, j = 0..YT-1
    double c0 = 0.0, c1 = 0.0;
    for (int base = 0; base < YT; base += 32) {
      double inv = 1.0 / (r + (double)(base + lane));
      double a = inv, b = -inv * inv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double ta = __shfl_up_sync(FULL, a, o), tb = __shfl_up_sync(FULL, b, o);
        if (lane >= o) { a += ta; b += tb; }
      }
      t0[base + lane] = c0 + a - inv; t1[base + lane] = c1 + b + inv * inv;
      c0 += __shfl_sync(FULL, a, 31); c1 += __shfl_sync(FULL, b, 31);
    }
    __syncwarp();
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; }
  __device__ __forceinline__ void one(int y, double lmu, bool m) {
    if (!m) return;
    double mu = g.expm(lmu);
    double mr = mu + r, imr = 1.0 / mr, yd = (double)y;
    double d0, d1;
    if (y < YT) { d0 = t0[y]; d1 = t1[y]; } else { d0 = digamma_d(yd + r) - dg_r; d1 = trigamma_d(yd + r) - tg_r; }
    double q = (yd + r) * imr;
    g1 += d0 + lr + 1.0 - flog(mr, g.mt->rc, g.mt->lc) - q;
    g2 += d1 + ir - 2.0 * imr + q * imr;
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    if (!live) return;
    double l0, l1;
    g.lmu_pair(tile, ti, l0, l1);
    one(y0, l0, m0); one(y1, l1, m1);
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int row) {
    if (!live) return;
    double s1 = warp_sum(g1), s2 = warp_sum(g2);
    if ((threadIdx.x & 31) == 0) {
      double dth = -r * s1, d2th = r * r * s2 + r * s1;
      double step = (d2th < 0.0) ? -dth / d2th : (dth > 0.0 ? 1.0 : -1.0);
      if (step > 2.0) step = 2.0;
      if (step < -2.0) step = -2.0;
      if (step == step) {
        theta[row] += step;
        if (!(fabs(step) < tol)) atomicExch(notconv, 1);
      } else {
        theta[row] = step;  // NaN: give up on this gene
      }
    }
  }
};
template <int K>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_psi_newton(RowGeom geo, GeneState<K> st, double* theta, double tol, int* notconv) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = st.W;
  OpPsiNewton<K> op;
  op.st = st; op.theta = theta; op.tol = tol; op.notconv = notconv;
  stream_rows<K, 1>(geo, &sh, op, smem);
}
static __global__ void k_log_vec(const double* x, long long n, double* y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = log(x[i]);
}
static __global__ void k_exp_vec(const double* x, long long n, double* y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = exp(x[i]);
}
