// winsor.cuh -- H0: ceiling(t(apply(Y, 1, DescTools::Winsorize, probs = c(0, 0.995)))) (R/ruvIIInb.R:53;
// rowQuantiles in R/fastruvIIInb.R:81-95) on the resident int32 counts, plus rowSums(Y > 0) for the zero-gene
// filter (:57).  Per gene: q = type-7 quantile at `prob` = (1-h) x[lo] + h x[hi], lo = floor(1 + (n-1) prob);
// counts above q become ceil(q), i.e. Y = min(Y, ceil(q)) for integer Y.  One CTA per gene row; the two order
// statistics come from a counting histogram when the row maximum is < 4096 (one pass), else from a 12/12/8-bit
// radix descent on the 32-bit value.  Included by core.cu (winsorize).
#pragma once
#include "common.cuh"

#define WZ_NT 512
#define WZ_BINS 4096

// kth (0-based) smallest non-negative value of the row by radix descent (pads are negative and skipped)
static __device__ unsigned wz_select(const int32_t* y, long long Np, long long kth, unsigned* hist, unsigned* s_pick) {
  const int tid = threadIdx.x;
  unsigned prefix = 0, pmask = 0;
  const int shifts[3] = {20, 8, 0}, bits[3] = {12, 12, 8};
  for (int lv = 0; lv < 3; ++lv) {
    const int nb = 1 << bits[lv], sh = shifts[lv];
    for (int i = tid; i < nb; i += WZ_NT) hist[i] = 0;
    __syncthreads();
    for (long long i = tid; i < Np; i += WZ_NT) {
      int v = y[i];
      if (v >= 0 && ((unsigned)v & pmask) == prefix) atomicAdd(&hist[((unsigned)v >> sh) & (nb - 1)], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      long long acc = 0; int b = 0;
      for (; b < nb - 1; ++b) { if (acc + hist[b] > kth) break; acc += hist[b]; }
      s_pick[0] = (unsigned)b; s_pick[1] = (unsigned)acc;
    }
    __syncthreads();
    prefix |= s_pick[0] << sh; pmask |= (unsigned)(nb - 1) << sh; kth -= s_pick[1];
    __syncthreads();
  }
  return prefix;
}

static __global__ void __launch_bounds__(WZ_NT) k_winsorize(int32_t* Yp, long long Np, long long n, double prob, double* cap_out,
                                                    long long* nz_out) {
  __shared__ unsigned hist[WZ_BINS];
  __shared__ unsigned s_pick[4];
  __shared__ int s_red[WZ_NT / 32];
  __shared__ long long s_nz[WZ_NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t* y = Yp + (long long)blockIdx.x * Np;
  // row maximum
  int mx = 0;
  for (long long i = tid; i < Np; i += WZ_NT) mx = max(mx, y[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = 0;
  for (int w = 0; w < WZ_NT / 32; ++w) mx = max(mx, s_red[w]);
  __syncthreads();
  // type-7 positions (stats::quantile): index = 1 + (n-1) p, 1-based
  const double index = 1.0 + (double)(n > 0 ? n - 1 : 0) * prob;
  const double lo = floor(index), hi = ceil(index);
  const long long klo = (long long)lo - 1, khi = (long long)hi - 1;
  unsigned vlo, vhi;
  if (mx < WZ_BINS) {
    for (int i = tid; i < WZ_BINS; i += WZ_NT) hist[i] = 0;
    __syncthreads();
    for (long long i = tid; i < Np; i += WZ_NT) { int v = y[i]; if (v >= 0) atomicAdd(&hist[v], 1u); }
    __syncthreads();
    if (tid == 0) {
      long long acc = 0; unsigned a = 0, b = 0; bool ga = false;
      for (int v = 0; v <= mx; ++v) {
        acc += hist[v];
        if (!ga && acc > klo) { a = (unsigned)v; ga = true; }
        if (acc > khi) { b = (unsigned)v; break; }
      }
      s_pick[2] = a; s_pick[3] = b;
    }
    __syncthreads();
    vlo = s_pick[2]; vhi = s_pick[3];
  } else {
    vlo = wz_select(y, Np, klo, hist, s_pick);
    vhi = (khi == klo) ? vlo : wz_select(y, Np, khi, hist, s_pick);
  }
  double q = (double)vlo;
  if (index > lo && vhi != vlo) { double h = index - lo; q = (1.0 - h) * (double)vlo + h * (double)vhi; }
  const double capd = ceil(q);
  const int cap = capd >= 2147483647.0 ? 2147483647 : (int)capd;
  long long nz = 0;
  for (long long i = tid; i < Np; i += WZ_NT) {
    int v = y[i];
    if (v > cap) { v = cap; y[i] = v; }
    nz += v > 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(FULL, nz, o);
  if (lane == 0) s_nz[warp] = nz;
  __syncthreads();
  if (tid == 0) {
    long long t = 0;
    for (int w = 0; w < WZ_NT / 32; ++w) t += s_nz[w];
    if (cap_out) cap_out[blockIdx.x] = capd;
    if (nz_out) nz_out[blockIdx.x] = t;
  }
}

// ---- the same over cell shards: histograms summed over the ranks (core.cu ruvnb_winsorize) ------------------------------------
// Pass A: hist[g][min(v, 4095)] -- resolves every gene whose two order statistics lie below the overflow bin (all but genes with
// counts in the thousands in > 0.5 % of the cells).  Pass B, unresolved genes only (cap[g] < 0), once per order statistic: the
// 12/12/8-bit radix descent of wz_select with every level's histogram summed over the ranks.
static __global__ void __launch_bounds__(WZ_NT) k_wz_hist(const int32_t* Yp, long long Np, int shift, int nb, unsigned pmask, const unsigned* prefix,
                                                         const double* cap, int direct, unsigned* hist) {
  __shared__ unsigned h[WZ_BINS];
  const int g = blockIdx.x, tid = threadIdx.x;
  if (!direct && !(cap[g] < 0.0)) return;
  for (int i = tid; i < WZ_BINS; i += WZ_NT) h[i] = 0;
  __syncthreads();
  const int32_t* y = Yp + (long long)g * Np;
  const unsigned pf = direct ? 0u : prefix[g];
  for (long long i = tid; i < Np; i += WZ_NT) {
    const int v = y[i];
    if (v < 0) continue;   // pad
    if (direct) atomicAdd(&h[v < WZ_BINS - 1 ? v : WZ_BINS - 1], 1u);
    else if (((unsigned)v & pmask) == pf) atomicAdd(&h[((unsigned)v >> shift) & (unsigned)(nb - 1)], 1u);
  }
  __syncthreads();
  for (int i = tid; i < WZ_BINS; i += WZ_NT) if (h[i]) hist[(long long)g * WZ_BINS + i] = h[i];
}
// type-7 positions (stats::quantile): index = 1 + (n-1) p; the two 0-based ranks and the interpolation weight
__device__ __forceinline__ void wz_ranks(long long n, double prob, long long* klo, long long* khi, double* frac) {
  const double index = 1.0 + (double)(n > 0 ? n - 1 : 0) * prob;
  const double lo = floor(index), hi = ceil(index);
  *klo = (long long)lo - 1; *khi = (long long)hi - 1; *frac = index - lo;
}
__device__ __forceinline__ double wz_cap_from(double a, double b, double frac) {
  double q = a;
  if (frac > 0.0 && b != a) q = (1.0 - frac) * a + frac * b;
  return ceil(q);
}
// pass A: one thread per gene
static __global__ void k_wz_pick_direct(int Gl, long long n, double prob, const unsigned* hist, double* cap, int* need) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  long long klo, khi; double frac;
  wz_ranks(n, prob, &klo, &khi, &frac);
  const unsigned* h = hist + (long long)g * WZ_BINS;
  long long acc = 0; int a = -1, b = -1;
  for (int v = 0; v < WZ_BINS; ++v) {
    acc += h[v];
    if (a < 0 && acc > klo) a = v;
    if (acc > khi) { b = v; break; }
  }
  if (a >= 0 && b >= 0 && b < WZ_BINS - 1) cap[g] = wz_cap_from((double)a, (double)b, frac);
  else { cap[g] = -1.0; atomicAdd(need, 1); }
}
// pass B: start of a descent for order statistic `which` (0 lower, 1 upper)
static __global__ void k_wz_desc_init(int Gl, long long n, double prob, int which, const double* cap, unsigned* prefix, long long* rank) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl || !(cap[g] < 0.0)) return;
  long long klo, khi; double frac;
  wz_ranks(n, prob, &klo, &khi, &frac);
  prefix[g] = 0u; rank[g] = which ? khi : klo;
}
static __global__ void k_wz_desc_pick(int Gl, const unsigned* hist, int shift, int nb, const double* cap, unsigned* prefix, long long* rank) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl || !(cap[g] < 0.0)) return;
  const unsigned* h = hist + (long long)g * WZ_BINS;
  long long acc = 0, rk = rank[g]; int b = 0;
  for (; b < nb - 1; ++b) { if (acc + (long long)h[b] > rk) break; acc += h[b]; }
  prefix[g] |= (unsigned)b << shift; rank[g] = rk - acc;
}
static __global__ void k_wz_desc_finish(int Gl, long long n, double prob, const unsigned* vlo, const unsigned* vhi, double* cap) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl || !(cap[g] < 0.0)) return;
  long long klo, khi; double frac;
  wz_ranks(n, prob, &klo, &khi, &frac);
  cap[g] = wz_cap_from((double)vlo[g], (double)vhi[g], frac);
}
static __global__ void __launch_bounds__(WZ_NT) k_wz_apply(int32_t* Yp, long long Np, const double* capd, long long* nz_out) {
  __shared__ long long s_nz[WZ_NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t* y = Yp + (long long)blockIdx.x * Np;
  const double cd = capd[blockIdx.x];
  const int cap = cd >= 2147483647.0 ? 2147483647 : (int)cd;
  long long nz = 0;
  for (long long i = tid; i < Np; i += WZ_NT) {
    int v = y[i];
    if (v > cap) { v = cap; y[i] = v; }
    nz += v > 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(FULL, nz, o);
  if (lane == 0) s_nz[warp] = nz;
  __syncthreads();
  if (tid == 0) { long long t = 0; for (int w = 0; w < WZ_NT / 32; ++w) t += s_nz[w]; nz_out[blockIdx.x] = t; }
}
// out[c][g] = Yp[g][pos[c]]: n cells gathered into R's column-major G x n (32 x 32 tiles through shared memory)
static __global__ void k_gather_cells(const int32_t* Yp, int Gl, long long Np, const int* pos, long long n, int32_t* out) {
  __shared__ int32_t tile[32][33];
  const long long gb = (long long)blockIdx.x * 32, cb = (long long)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) { long long g = gb + i, c = cb + threadIdx.x; if (g < Gl && c < n) tile[i][threadIdx.x] = Yp[g * Np + pos[c]]; }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) { long long c = cb + i, g = gb + threadIdx.x; if (g < Gl && c < n) out[c * Gl + g] = tile[threadIdx.x][i]; }
}

// inverse of the upload permutation: dst[r][cell] = Yp[r][cell2pos[cell]] (gene-major host layout)
static __global__ void k_unpermute_rows(const int32_t* Yp, long long rows, long long N, long long Np, const int* cell2pos, int32_t* dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * N) return;
  long long r = i / N, c = i - r * N;
  dst[i] = Yp[r * Np + cell2pos[c]];
}
