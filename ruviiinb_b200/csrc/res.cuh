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
// continuity fuzz) by the same recurrence.  Included by ruvnb.cu only.
#pragma once
#include "kernels.cuh"

#define RES_PEARSON 1
#define RES_LOGCOUNTS 2
#define RES_QUANTILE 3

struct ResArgs {
  int Gl, K, m, B;
  long long N, Np, c0;
  const int32_t* Yp;
  const double *alpha, *gmean, *psi, *W, *Mb;   // current parameter set (Mb: [Gl][m])
  const double* mb_blk;                         // [B][Gl] per-cell Mb of this block (fastruvIII.nb's Mb.all) or nullptr
  const int *cell2pos, *chunk_grp;
  const double* wa_mean;                        // [Gl] a %*% colMeans(W)
  double* logcap;                               // [Gl][3]
};

__device__ __forceinline__ double res_mb(const ResArgs& a, int g, int c, long long pos) {
  return a.mb_blk ? a.mb_blk[(long long)c * a.Gl + g] : a.Mb[(long long)g * a.m + a.chunk_grp[pos >> 7]];
}

static __global__ void __launch_bounds__(256) k_res_caps(ResArgs a) {
  extern __shared__ double dyn[];   // wa[B], mb[B]
  __shared__ SelScratch sc;
  double* wa = dyn; double* mb = dyn + a.B;
  const int g = blockIdx.x, tid = threadIdx.x;
  const double gm = a.gmean[g], wam = a.wa_mean[g];
  for (int c = tid; c < a.B; c += 256) {
    long long pos = a.cell2pos[a.c0 + c];
    double s = 0.0;
    for (int l = 0; l < a.K; ++l) s = fma(a.alpha[(long long)g * a.K + l], a.W[(long long)l * a.Np + pos], s);
    wa[c] = s;
    mb[c] = res_mb(a, g, c, pos);
  }
  __syncthreads();
  auto vld = [&](int) { return true; };
  double med, mad;
  auto v0 = [&](int i) { return wa[i] + gm; };
  group_median_mad<256>(a.B, v0, vld, &sc, tid, &med, &mad);
  if (tid == 0) a.logcap[(long long)g * 3 + 0] = med + 4.0 * mad;
  __syncthreads();
  auto v1 = [&](int i) { return wa[i] + gm + mb[i]; };
  group_median_mad<256>(a.B, v1, vld, &sc, tid, &med, &mad);
  if (tid == 0) a.logcap[(long long)g * 3 + 1] = med + 4.0 * mad;
  __syncthreads();
  auto v2 = [&](int i) { return mb[i] + gm + wam; };
  group_median_mad<256>(a.B, v2, vld, &sc, tid, &med, &mad);
  if (tid == 0) a.logcap[(long long)g * 3 + 2] = med + 4.0 * mad;
}

// log((r)/(r + mu)) the way R's dnbinom_mu takes it at x = 0
__device__ __forceinline__ double res_log_p0(double r, double mu) {
  return r * (r < mu ? log(r / (r + mu)) : log1p(-mu / (r + mu)));
}
// percentile-adjusted count of one element
static __device__ double res_pac(int y, double mu_full, double mu_nouv, double r) {
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
  double sc = exp(ls);
  double lo = s * sc, hi = lo + t * sc;      // pnbinom(y - 1), + dnbinom(y)
  double p = (lo + hi) / 2.0;
  if (p > 0.995) p = 0.995;
  if (p < 0.005) p = 0.005;
  if (!(p == p)) return p;
  // qnbinom(p; mu_nouv, r)
  if (mu_nouv == 0.0) return 0.0;
  const double target = p * (1.0 - 64.0 * 2.220446049250313e-16);
  ls = res_log_p0(r, mu_nouv);
  const double ratio2 = mu_nouv / (mu_nouv + r);
  sc = exp(ls);
  t = 1.0;
  double cum = 1.0;
  long long q = 0;
  while (!(cum * sc >= target) && q < 2000000000ll) {
    t *= ((double)q + r) * frcp((double)q + 1.0) * ratio2;
    ++q;
    cum += t;
    if (t > BIG) { t *= 1e-250; cum *= 1e-250; ls += LBIG; sc = exp(ls); }
    if (!(t == t)) return t;
  }
  return (double)q;
}

// one CTA = 32 genes x 32 cells; warp w handles gene g0 + w, lane l cell cb + l; output slab[c * Gl + g]
template <int TYPE>
__global__ void __launch_bounds__(1024) k_res_elem(ResArgs a, double* slab, int* nan_count) {
  __shared__ double tile[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x * 32 + warp, c = blockIdx.y * 32 + lane;
  double v = 0.0;
  if (g < a.Gl && c < a.B) {
    const long long pos = a.cell2pos[a.c0 + c];
    const int y = a.Yp[(long long)g * a.Np + pos];
    double wa = 0.0;
    for (int l = 0; l < a.K; ++l) wa = fma(a.alpha[(long long)g * a.K + l], a.W[(long long)l * a.Np + pos], wa);
    const double gm = a.gmean[g];
    if (TYPE == RES_LOGCOUNTS) {
      v = log((double)y / exp(wa) + 1.0);                                   // :68-74
    } else {
      const double mb = res_mb(a, g, c, pos);
      const double psi = a.psi[g];
      double mf = exp(wa + gm + mb);
      const double cap1 = exp(a.logcap[(long long)g * 3 + 1]);
      if (mf > cap1) mf = cap1;                                               // :42-44
      if (TYPE == RES_PEARSON) {
        double mu = exp(wa + gm);
        const double cap0 = exp(a.logcap[(long long)g * 3 + 0]);
        if (mu > cap0) mu = cap0;                                             // :38-40
        v = ((double)y - mu) / sqrt(mf * (1.0 + mf * psi));                   // :57
        if (v != v) atomicAdd(nan_count, 1);
      } else {
        double mn = exp(mb + gm + a.wa_mean[g]);
        const double cap2 = exp(a.logcap[(long long)g * 3 + 2]);
        if (mn > cap2) mn = cap2;                                             // :77-80
        v = res_pac(y, mf, mn, 1.0 / psi);                                    // :88-99
      }
    }
  }
  tile[warp][lane] = v;
  __syncthreads();
  const int g2 = blockIdx.x * 32 + lane, c2 = blockIdx.y * 32 + warp;
  if (g2 < a.Gl && c2 < a.B) slab[(long long)c2 * a.Gl + g2] = tile[lane][warp];
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
static __global__ void k_res_clip(double* slab, long long len, const SelTargets* T) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double v = slab[i];
  if (v < T->lims[0]) v = T->lims[0];
  if (v > T->lims[1]) v = T->lims[1];
  slab[i] = v;
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
static __global__ void k_res_wamean(int Gl, int K, const double* alpha, const double* wm, double* out) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  double s = 0.0;
  for (int l = 0; l < K; ++l) s = fma(alpha[(long long)g * K + l], wm[l], s);
  out[g] = s;
}
