// res.cuh -- get.res (reference R/get.res.R:12-129), NB branch with a vector psi: Pearson residuals, log-normalised
// counts and percentile-adjusted counts (PAC), block by block of `block.size` cells exactly as the reference streams
// them -- the per-gene mean caps (:38-45,77-80) and the Pearson clip limits (:63-66) are per-block quantities.
//   k_res_caps   per (gene, block): exp(median + 4 mad) caps of log mu, log mu.full, log mu.noUV (exact radix select on
//                the <= block.size values held in shared memory)                                              [R1]
//   k_res_elem   per element: Pearson / logcounts / PAC, 32 genes x 32 cells per CTA, transposed through shared memory so
//                that the block slab is written gene-fastest (R's column-major G x B)                            [R2-R4]
//   k_sel_*      exact type-7 quantiles (0.5 %, 99.5 %) of the G x B Pearson values of a block: 12-bit radix descent for
//                the four order statistics at once, CTA-private histograms flushed to a global one               [R2]
//   k_res_clip   clip to those limits
// PAC: mid-p = P(Y < y) + P(Y = y)/2 under (mu.full, r) by the pmf recurrence p(j+1) = p(j) (j + r)/(j + 1) mu/(mu + r),
// clipped to [.005, .995]; then qnbinom(p; mu.noUV, r) = smallest q with cdf(q) >= p (1 - 64 eps) (R nmath's left-
// continuity fuzz) by the same recurrence.  Included by res.cu.
#pragma once
#include "kernels.cuh"

#define RES_PEARSON 1
#define RES_LOGCOUNTS 2
#define RES_QUANTILE 3

struct ResArgs {
  int Gl, K, m, B;
  int ldB;                                      // leading dimension (cells) of mbG
  int Gout;                                     // rows of the output slab: live rows only (orow)
  long long N, Np, c0;
  const int32_t* Yp;
  const double *alpha, *gmean, *psi, *W, *Mb;   // current parameter set (Mb: [Gl][m])
  const double* mbG;                            // [Gl][ldB] per-cell Mb of this block, gene-major (fastruvIII.nb's Mb.all), or nullptr
  const int *cell2pos, *chunk_grp;
  const int* chunk_batch;                       // [nchunks] batch of every chunk: psi is [nbatch][Gl] (R/get.res.R:59-62,101-121: out$psi[, curr.batch])
  int ref_batch;                                // which.min(colMeans(out$psi)) (:113,120): the dispersion column of the quantile call
  const int* orow;                              // [Gl] output row of every resident row, -1 = row not reported (all-zero gene), or nullptr
  const double* pi0;                            // [Gl] zero-inflation probability of a ZINB fit, or nullptr (NB fit)
  const uint8_t* zflag;                         // [N] 1 where omega_gc = pi0_g applies: the reference reads out$pi0[, batch[c]], i.e. the
                                                //     zero-inflation flag of cell number batch[c] (R/get.res.R:53,84-85)
  double zavg;                                  // mean(zeroinf): rowMeans(out$pi0) = pi0_g zavg (:95)
  const double* wa_mean;                        // [Gl] a %*% colMeans(W)
  double* caps;                                 // [Gl][3] exp(median + 4 mad) of log mu, log mu.full, log mu.noUV over the block
  const MathTabs* mtabs;
};

__device__ __forceinline__ double res_mb(const ResArgs& a, int g, int c, long long pos) {
  return a.mbG ? a.mbG[(long long)g * a.ldB + c] : a.Mb[(long long)g * a.m + a.chunk_grp[pos >> 7]];
}

// R1: per (gene, block) caps exp(rowMedians(log mu) + 4 rowMads(log mu)) (R/get.res.R:38-45,77-80).  One CTA of RS_NT
// threads per gene; a W' and Mb of the block's cells sit in shared memory; both middle order statistics of a vector come
// out of ONE 12-bit radix descent (select.cuh cta_two_middles: 2-3 passes over the <= 12288 values instead of the 18 of
// an 8-bit select with a separate lower-middle pass).  which: bit 0 mu, bit 1 mu.full, bit 2 mu.noUV.
static __global__ void __launch_bounds__(RS_NT) k_res_caps(ResArgs a, int which) {
  extern __shared__ double dyn[];   // wa[Bp], mb[Bp]
  __shared__ RowSelScratch sc;
  const int g = blockIdx.x, tid = threadIdx.x;
  if (a.orow && a.orow[g] < 0) return;
  const int Bp = (a.B + RS_NT - 1) / RS_NT * RS_NT;
  double* wa = dyn; double* mb = dyn + Bp;
  const double gm = a.gmean[g], wam = a.wa_mean[g];
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  int nnan = 0;
  for (int c = tid; c < Bp; c += RS_NT) {
    double s = inf, b = 0.0;   // pads sort last in every vector
    if (c < a.B) {
      const long long pos = a.cell2pos[a.c0 + c];
      s = 0.0;
      for (int l = 0; l < a.K; ++l) s = fma(a.alpha[(long long)g * a.K + l], a.W[(long long)l * a.Np + pos], s);
      b = res_mb(a, g, c, pos);
      nnan |= (s != s) | (b != b);
    }
    wa[c] = s; mb[c] = b;
  }
  nnan = __syncthreads_or(nnan);
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  const long long n = a.B;
  auto med_mad = [&](auto val) -> double {   // median + 4 mad of val(i), i < B (R: NA if any NaN)
    if (nnan || gm != gm) return qnan;
    unsigned long long klo, khi;
    auto k1 = [&](int i) { return f64_key(val(i)); };
    cta_two_middles<48>(Bp, n, k1, &sc, &klo, &khi);
    const double med = (n & 1) ? key_f64(khi) : (key_f64(klo) + key_f64(khi)) / 2.0;
    auto k2 = [&](int i) { return i < a.B ? f64_key(fabs(val(i) - med)) : f64_key(inf); };
    cta_two_middles<48>(Bp, n, k2, &sc, &klo, &khi);
    const double mad = MAD_CONST * ((n & 1) ? key_f64(khi) : (key_f64(klo) + key_f64(khi)) / 2.0);
    return med + 4.0 * mad;
  };
  if (which & 1) { double v = med_mad([&](int i) { return wa[i] + gm; }); if (tid == 0) a.caps[(long long)g * 3 + 0] = exp(v); }
  if (which & 2) { double v = med_mad([&](int i) { return wa[i] + gm + mb[i]; }); if (tid == 0) a.caps[(long long)g * 3 + 1] = exp(v); }
  if (which & 4) { double v = med_mad([&](int i) { return i < a.B ? mb[i] + gm + wam : inf; }); if (tid == 0) a.caps[(long long)g * 3 + 2] = exp(v); }
}

// log((r)/(r + mu)) the way R's dnbinom_mu takes it at x = 0
__device__ __forceinline__ double res_log_p0(double r, double mu) {
  return r * (r < mu ? log(r / (r + mu)) : log1p(-mu / (r + mu)));
}
// percentile-adjusted count of one element.  omega / omega_avg: zero-inflation probabilities of a ZINB fit (0 for NB), with ZIM's
// definitions (ZIM 1.1.0, third-party, restated): dzinb = omega 1[x = 0] + (1 - omega) dnbinom, pzinb(q) = omega + (1 - omega) pnbinom(q)
// (also for q = -1, where it returns omega), qzinb(p) = qnbinom(max(0, (p - omega)/(1 - omega))).  R/get.res.R:84-99.
// r: size of the mid-p step (the cell's own batch), r_q: size of the quantile step (the reference batch; the same with one batch).
static __device__ double res_pac(int y, double mu_full, double mu_nouv, double r, double r_q, const MathTabs* mt, double omega, double omega_avg) {
  const double BIG = 1e250, LBIG = 575.64627324851142;  // log(1e250)
  // mid-p under (mu_full, r)
  double ls = res_log_p0(r, mu_full);
  const double ratio = mu_full / (mu_full + r);
  double t = 1.0, s = 0.0;
  for (int j = 0; j < y; ++j) {
    s += t;
    t *= ((double)j + r) * frcp((double)j + 1.0) * ratio;
    if (t > BIG) { t *= 1e-250; s *= 1e-250; ls += LBIG; }
  }
  double sc = fexp(ls, mt->ex);
  double lo = s * sc, hi = lo + t * sc;      // pnbinom(y - 1), + dnbinom(y)
  if (omega != 0.0) {                        // pzinb(y - 1), + dzinb(y)
    const double d = (hi - lo) * (1.0 - omega) + (y == 0 ? omega : 0.0);
    lo = omega + (1.0 - omega) * lo;
    hi = lo + d;
  }
  double p = (lo + hi) / 2.0;
  if (p > 0.995) p = 0.995;
  if (p < 0.005) p = 0.005;
  if (!(p == p)) return p;
  if (omega_avg != 0.0) { p = (p - omega_avg) / (1.0 - omega_avg); if (p < 0.0) p = 0.0; }   // qzinb
  // qnbinom(p; mu_nouv, r_q)
  if (mu_nouv == 0.0 || p == 0.0) return 0.0;
  const double target = p * (1.0 - 64.0 * 2.220446049250313e-16);
  r = r_q;
  ls = res_log_p0(r, mu_nouv);
  const double ratio2 = mu_nouv / (mu_nouv + r);
  sc = fexp(ls, mt->ex);
  t = 1.0;
  double cum = 1.0;
  long long q = 0;
  while (!(cum * sc >= target) && q < 2000000000ll) {
    t *= ((double)q + r) * frcp((double)q + 1.0) * ratio2;
    ++q;
    cum += t;
    if (t > BIG) { t *= 1e-250; cum *= 1e-250; ls += LBIG; sc = fexp(ls, mt->ex); }
    if (!(t == t)) return t;
  }
  return (double)q;
}

// Element kernel: warp = one gene x 32 consecutive cells, every warp on its own -- the pmf recurrences of the PAC run as long as
// the gene's counts are large, so warps of one CTA finish far apart and a CTA-wide barrier (the shared-memory transpose this
// kernel used to end with) left 2/3 of the issue slots waiting (ncu: stall_barrier 66 %).  Results go to a gene-major scratch
// tmp[g][ldB] (coalesced); k_res_out transposes it into the R-layout slab.
template <int TYPE>
static __global__ void __launch_bounds__(256) k_res_elem(ResArgs a, double* tmp, int* nan_count) {
  __shared__ MathTabs mt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const double* src = reinterpret_cast<const double*>(a.mtabs);
    double* dst = reinterpret_cast<double*>(&mt);
    for (int i = threadIdx.x; i < (int)(sizeof(MathTabs) / 8); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  const int g = blockIdx.x * 8 + warp, c = blockIdx.y * 32 + lane;
  if (g >= a.Gl || c >= a.B || (a.orow && a.orow[g] < 0)) return;
  double v;
  const long long pos = a.cell2pos[a.c0 + c];
  const int y = a.Yp[(long long)g * a.Np + pos];
  double wa = 0.0;
  for (int l = 0; l < a.K; ++l) wa = fma(a.alpha[(long long)g * a.K + l], a.W[(long long)l * a.Np + pos], wa);
  const double gm = a.gmean[g];
  const double omega = (a.pi0 && a.zflag[a.c0 + c]) ? a.pi0[g] : 0.0;       // out$pi0[, curr.batch], :53,69,84
  if (TYPE == RES_LOGCOUNTS) {
    v = log((double)y / (fexp(wa, mt.ex) * (1.0 - omega)) + 1.0);           // :68-74
  } else {
    const double mb = res_mb(a, g, c, pos);
    const double psi = a.psi[(long long)a.chunk_batch[pos >> 7] * a.Gl + g];
    double mf = fexp(wa + gm + mb, mt.ex);
    const double cap1 = a.caps[(long long)g * 3 + 1];
    if (mf > cap1) mf = cap1;                                               // :42-44
    if (TYPE == RES_PEARSON) {
      double mu = fexp(wa + gm, mt.ex);
      const double cap0 = a.caps[(long long)g * 3 + 0];
      if (mu > cap0) mu = cap0;                                             // :38-40
      v = ((double)y - mu * (1.0 - omega)) / sqrt(mf * (1.0 - omega) * (1.0 + mf * (psi + omega)));   // :53-57
      if (v != v) atomicAdd(nan_count, 1);
    } else {
      double mn = fexp(mb + gm + a.wa_mean[g], mt.ex);
      const double cap2 = a.caps[(long long)g * 3 + 2];
      if (mn > cap2) mn = cap2;                                             // :77-80
      v = res_pac(y, mf, mn, 1.0 / psi, 1.0 / a.psi[(long long)a.ref_batch * a.Gl + g], &mt, omega, a.pi0 ? a.pi0[g] * a.zavg : 0.0);   // :84-99,101-121
    }
  }
  tmp[(long long)g * a.ldB + c] = v;
}
// tmp[g][ldB] -> slab[c * Gout + orow[g]] (R's column-major G x B; rows without an output row are skipped).  OUT: 0 doubles,
// 1 int32 (PAC), 2 log(PAC + 1) as doubles (makeSCE's logPAC, R/makeSCE.R:36-38).  sum (nullable): every CTA adds the sum of
// what it wrote (the checksum of the sink that keeps nothing).  32 x 32 tiles, 256 threads.
template <int OUT>
static __global__ void __launch_bounds__(256) k_res_out(ResArgs a, const double* tmp, void* slab_, double* sum) {
  __shared__ double tile[32][33];
  __shared__ double s_part[8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int gb = blockIdx.x * 32, cb = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8) {
    const int g = gb + i, c = cb + tx;
    tile[i][tx] = (g < a.Gl && c < a.B && (!a.orow || a.orow[g] >= 0)) ? tmp[(long long)g * a.ldB + c] : 0.0;
  }
  __syncthreads();
  double w = 0.0;
  for (int i = ty; i < 32; i += 8) {
    const int c = cb + i, g = gb + tx;
    if (g < a.Gl && c < a.B) {
      const int o = a.orow ? a.orow[g] : g;
      if (o >= 0) {
        double v = tile[tx][i];
        if (OUT == 2) v = log(v + 1.0);
        if (OUT == 1) reinterpret_cast<int32_t*>(slab_)[(long long)c * a.Gout + o] = (v >= 2147483647.0) ? 2147483647 : (int32_t)v;
        else reinterpret_cast<double*>(slab_)[(long long)c * a.Gout + o] = v;
        w += (v == v) ? v : 0.0;
      }
    }
  }
  if (sum) {
    w = warp_sum(w);
    if (tx == 0) s_part[ty] = w;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < 8; ++i) t += s_part[i]; atomicAdd(sum, t); }
  }
}

// [B][Gl] (a block of R's column-major G x N matrix) -> gene-major [Gl][ldB]
static __global__ void k_res_transpose(const double* src, int Gl, int B, int ldB, double* dst) {
  __shared__ double tile[32][33];
  const int gb = blockIdx.x * 32, cb = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) { int c = cb + i, g = gb + threadIdx.x; if (c < B && g < Gl) tile[i][threadIdx.x] = src[(long long)c * Gl + g]; }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) { int g = gb + i, c = cb + threadIdx.x; if (c < B && g < Gl) dst[(long long)g * ldB + c] = tile[threadIdx.x][i]; }
}

// ---- exact quantiles of a slab: four order statistics by 12-bit radix descent ---------------------------------------
#define SEL_NT 4
#define SEL_BINS 4096
struct SelTargets {
  unsigned long long prefix[SEL_NT], pmask;
  long long rank[SEL_NT];     // 0-based rank inside the current bucket
  long long n;                // number of non-NaN values
  double h[2];                // type-7 interpolation weights of the two quantiles
  double lims[2];
};
// ranks of the type-7 quantiles .005 / .995 of n values (stats::quantile): index = 1 + (n-1) p, lo = floor, hi = ceil
static __global__ void k_sel_init(SelTargets* T, long long n_total, const int* nan_count) {
  if (threadIdx.x || blockIdx.x) return;
  long long n = n_total - (long long)*nan_count;
  T->n = n; T->pmask = 0;
  const double probs[2] = {0.005, 0.995};
  for (int q = 0; q < 2; ++q) {
    double index = 1.0 + (double)(n > 0 ? n - 1 : 0) * probs[q];
    double lo = floor(index), hi = ceil(index);
    T->rank[2 * q] = (long long)lo - 1; T->rank[2 * q + 1] = (long long)hi - 1;
    T->h[q] = index - lo;
    T->prefix[2 * q] = T->prefix[2 * q + 1] = 0;
  }
}
static __global__ void __launch_bounds__(256) k_sel_hist(const double* slab, long long len, int shift, int bits, const SelTargets* T,
                                                 unsigned* hist /* [SEL_NT][SEL_BINS] global */) {
  extern __shared__ unsigned sh[];  // [SEL_NT][SEL_BINS]
  const int nb = 1 << bits;
  for (int i = threadIdx.x; i < SEL_NT * SEL_BINS; i += 256) sh[i] = 0;
  __syncthreads();
  unsigned long long pf[SEL_NT];
  const unsigned long long pmask = T->pmask;
#pragma unroll
  for (int t = 0; t < SEL_NT; ++t) pf[t] = T->prefix[t];
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < len; i += (long long)gridDim.x * 256) {
    double v = slab[i];
    if (v != v) continue;   // quantile(na.rm = TRUE)
    unsigned long long key = f64_key(v);
    unsigned d = (unsigned)((key >> shift) & (unsigned long long)(nb - 1));
    unsigned long long kp = key & pmask;
#pragma unroll
    for (int t = 0; t < SEL_NT; ++t) {
      bool dup = false;  // targets sharing a prefix share a histogram (the first one's)
#pragma unroll
      for (int u = 0; u < t; ++u) dup |= (pf[u] == pf[t]);
      if (dup) continue;   // block-uniform
      // warp-aggregated update: the first levels put most of a warp into a handful of bins
      const bool ok = kp == pf[t];
      const unsigned act = __activemask();
      const unsigned peers = __match_any_sync(act, ok ? d : 0xffffu);   // one shared dummy: match cost grows with distinct values
      if (ok && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[t * SEL_BINS + d], (unsigned)__popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SEL_NT * SEL_BINS; i += 256) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
static __global__ void __launch_bounds__(SEL_NT * 32) k_sel_pick(SelTargets* T, unsigned* hist, int shift, int bits) {
  // one warp per target: locate the bucket holding its rank (serial scan by lane 0 over <= 4096 bins)
  const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = 1 << bits;
  __shared__ unsigned long long npf[SEL_NT];
  __shared__ long long nrk[SEL_NT];
  if (lane == 0) {
    int src = t;
    for (int u = 0; u < t; ++u) if (T->prefix[u] == T->prefix[t]) { src = u; break; }
    const unsigned* h = hist + src * SEL_BINS;
    long long acc = 0, rk = T->rank[t];
    int b = 0;
    for (; b < nb - 1; ++b) { if (acc + (long long)h[b] > rk) break; acc += h[b]; }
    npf[t] = T->prefix[t] | ((unsigned long long)b << shift);
    nrk[t] = rk - acc;
  }
  __syncthreads();
  if (lane == 0) { T->prefix[t] = npf[t]; T->rank[t] = nrk[t]; }
  __syncthreads();
  if (threadIdx.x == 0) T->pmask |= (unsigned long long)(nb - 1) << shift;
  for (int i = threadIdx.x; i < SEL_NT * SEL_BINS; i += SEL_NT * 32) hist[i] = 0;
}
static __global__ void k_sel_finish(SelTargets* T) {
  if (threadIdx.x || blockIdx.x) return;
  for (int q = 0; q < 2; ++q) {
    double lo = key_f64(T->prefix[2 * q]), hi = key_f64(T->prefix[2 * q + 1]);
    double v = lo;
    if (T->h[q] > 0.0 && hi != lo) v = (1.0 - T->h[q]) * lo + T->h[q] * hi;   // type 7
    T->lims[q] = v;
  }
}
static __global__ void __launch_bounds__(256) k_res_clip(double* slab, long long len, const SelTargets* T, double* sum) {
  __shared__ double s_part[8];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < len) {
    v = slab[i];
    if (v < T->lims[0]) v = T->lims[0];
    if (v > T->lims[1]) v = T->lims[1];
    slab[i] = v;
  }
  if (sum) {
    v = (v == v) ? v : 0.0;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int q = 0; q < 8; ++q) t += s_part[q]; atomicAdd(sum, t); }
  }
}

// colMeans(W) over the real cells (pads hold 0) and a %*% colMeans(W)
static __global__ void __launch_bounds__(1024) k_res_wmean(const double* W, long long Np, double n, double* wm) {
  __shared__ double sm[1024];
  const double* w = W + (long long)blockIdx.x * Np;
  double s = 0.0;
  for (long long i = threadIdx.x; i < Np; i += 1024) s += w[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) wm[blockIdx.x] = sm[0] / n;
}
static __global__ void k_res_wamean(int Gl, int K, const double* alpha, const double* wsum, double inv_n, double* out) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  double s = 0.0;
  for (int l = 0; l < K; ++l) s = fma(alpha[(long long)g * K + l], wsum[l] * inv_n, s);   // colMeans(W) = colSums / N (all ranks)
  out[g] = s;
}

// ------------------------------------------------------------------------------------------------
// sparse sink: the block slab [B][Gout] (R's G x B column-major) compacted into the slots of a dgCMatrix -- what makeSCE stores
// (`as(log(PAC + 1), "sparseMatrix")`, R/makeSCE.R:36-38).  One warp per cell column; entries keep their row order.
// ------------------------------------------------------------------------------------------------
static __global__ void k_csc_count(const double* slab, int Gout, int B, int* cnt) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B) return;
  const double* col = slab + (long long)w * Gout;
  int n = 0;
  for (int g = lane; g < Gout; g += 32) n += col[g] != 0.0;
  n = __reduce_add_sync(0xffffffffu, n);
  if (lane == 0) cnt[w] = n;
}
// exclusive scan of the B column counts (one CTA): colptr[c] = *total + sum_{c' < c} cnt[c'] for c = 0..B - 1, then *total += block nnz
// (colptr[B] of the last block is written by the host).  blk_nnz[0] = this block's count.
static __global__ void __launch_bounds__(1024) k_csc_scan(int B, const int* cnt, long long* total, long long* colptr, long long* blk_nnz) {
  __shared__ long long part[1024];
  const int t = threadIdx.x, per = (B + 1023) / 1024;
  long long s = 0;
  for (int q = 0; q < per; ++q) { const int c = t * per + q; if (c < B) s += cnt[c]; }
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan
    long long v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = *total + (t ? part[t - 1] : 0);
  for (int q = 0; q < per; ++q) { const int c = t * per + q; if (c < B) { colptr[c] = run; run += cnt[c]; } }
  __syncthreads();
  if (t == 0) { *blk_nnz = part[1023]; *total += part[1023]; }
}
static __global__ void k_csc_fill(const double* slab, int Gout, int B, const long long* colptr, long long base, int* ridx, double* val) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B) return;
  const double* col = slab + (long long)w * Gout;
  long long o = colptr[w] - base;
  for (int g0 = 0; g0 < Gout; g0 += 32) {
    const int g = g0 + lane;
    const double v = g < Gout ? col[g] : 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, v != 0.0);
    if (v != 0.0) { const long long q = o + __popc(bal & ((1u << lane) - 1u)); ridx[q] = g; val[q] = v; }
    o += __popc(bal);
  }
}
