// winsor.cuh -- H0: ceiling(t(apply(Y, 1, DescTools::Winsorize, probs = c(0, 0.995)))) (R/ruvIIInb.R:53;
// rowQuantiles in R/fastruvIIInb.R:81-95) on the resident int32 counts, plus rowSums(Y > 0) for the zero-gene
// filter (:57).  Per gene: q = type-7 quantile at `prob` = (1-h) x[lo] + h x[hi], lo = floor(1 + (n-1) prob);
// counts above q become ceil(q), i.e. Y = min(Y, ceil(q)) for integer Y.  One CTA per gene row; the two order
// statistics come from a counting histogram when the row maximum is < 4096 (one pass), else from a 12/12/8-bit
// radix descent on the 32-bit value.  Included by ruvnb.cu only.
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

// inverse of the upload permutation: dst[r][cell] = Yp[r][cell2pos[cell]] (gene-major host layout)
static __global__ void k_unpermute_rows(const int32_t* Yp, long long rows, long long N, long long Np, const int* cell2pos, int32_t* dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * N) return;
  long long r = i / N, c = i - r * N;
  dst[i] = Yp[r * Np + cell2pos[c]];
}
