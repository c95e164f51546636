// svd.cuh -- H1: the initial W of ruvIII.nb (reference R/ruvIIInb.R:160-168),
//   W0 = irlba(log(Y[ctl,]/rowMeans(Y[ctl,]) + 1), nv = k)$v  -- the k leading right singular vectors of the n_ctl x N matrix A.
// Subspace iteration on A'A: V <- orth(A'(A V)) with A held as a dense FP64 slab (n_ctl x N, 0.8 GB at 1000 x 100k) for the
// duration of the call; then one Rayleigh-Ritz rotation.  irlba's Lanczos start is random and its tolerance is 1e-5, so the
// reference's own W0 is only defined up to that (and up to sign); parity runs inject W0.
// Sign convention: the entry of largest magnitude of every vector is positive.  Included by svd.cu.
#pragma once
#include "kernels.cuh"

// rowMeans(Y[ctl,]) over the real cells: one warp per control row
static __global__ void __launch_bounds__(256) k_svd_rowmean(const int32_t* const* ctl_rows, int n_ctl, long long Np, double n, double* rowmean) {
  const int gi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gi >= n_ctl) return;
  const int32_t* y = ctl_rows[gi];
  double s = 0.0;
  for (long long p = lane; p < Np; p += 32) { int v = y[p]; if (v > 0) s += (double)v; }   // pads are -1
  s = warp_sum(s);
  if (lane == 0) rowmean[gi] = s / n;
}
// A[gi][p] = log(Y[ctl_gi, p] / rowmean_gi + 1) as a dense FP64 slab [n_ctl][Np] (pads and zero counts 0), built once: the
// iterations then stream 8 B per element instead of paying a logarithm per element in both products
static __global__ void __launch_bounds__(256) k_svd_build(const int32_t* const* ctl_rows, int n_ctl, long long Np, const double* rowmean, double* A) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_ctl * Np) return;
  const int gi = (int)(i / Np); const long long p = i - (long long)gi * Np;
  const int v = ctl_rows[gi][p];
  A[i] = v > 0 ? log1p((double)v / rowmean[gi]) : 0.0;
}
// Upart[part][gi][l] = sum over the part's cells of A[gi,c] V[l][c].  A CTA holds 32 rows (4 per warp) over ONE cell range,
// so the V values its 8 warps load are the same lines (L1) and every V load feeds 4 rows; the parts fill the machine
// (n_ctl / 32 row blocks alone would not) and are summed in a fixed order afterwards.
#define SVD_RPW 4
template <int K>
__global__ void __launch_bounds__(256) k_svd_av(const double* __restrict__ A, int n_ctl, long long Np, long long seg,
                                               const double* __restrict__ V, double* __restrict__ Upart) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g0 = (blockIdx.x * 8 + warp) * SVD_RPW;
  const long long p0 = (long long)blockIdx.y * seg, p1 = p0 + seg < Np ? p0 + seg : Np;
  double acc[SVD_RPW][K];
#pragma unroll
  for (int r = 0; r < SVD_RPW; ++r)
#pragma unroll
    for (int l = 0; l < K; ++l) acc[r][l] = 0.0;
  const double* a[SVD_RPW];
#pragma unroll
  for (int r = 0; r < SVD_RPW; ++r) a[r] = A + (long long)(g0 + r < n_ctl ? g0 + r : n_ctl - 1) * Np;
  for (long long p = p0 + lane; p < p1; p += 32) {
    double v[K];
#pragma unroll
    for (int l = 0; l < K; ++l) v[l] = V[(long long)l * Np + p];
#pragma unroll
    for (int r = 0; r < SVD_RPW; ++r) {
      const double x = a[r][p];
#pragma unroll
      for (int l = 0; l < K; ++l) acc[r][l] = fma(x, v[l], acc[r][l]);
    }
  }
#pragma unroll
  for (int r = 0; r < SVD_RPW; ++r)
#pragma unroll
    for (int l = 0; l < K; ++l) {
      double t = warp_sum(acc[r][l]);
      if (lane == 0 && g0 + r < n_ctl) Upart[((long long)blockIdx.y * n_ctl + g0 + r) * K + l] = t;
    }
}
static __global__ void k_svd_sum_parts(const double* Upart, long long n, int parts, double* U) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int q = 0; q < parts; ++q) s += Upart[(long long)q * n + i];
  U[i] = s;
}
// Vout[l][p] = sum_gi A[gi,p] U[gi][l]: a CTA owns 32 cells, warp w the rows gi = w (mod 8); a row's 32 cells are one
// 256-byte read; the 8 partial sums are combined through shared memory in warp order
template <int K>
__global__ void __launch_bounds__(256) k_svd_atu(const double* __restrict__ A, int n_ctl, long long Np, const double* __restrict__ U,
                                                double* __restrict__ Vout) {
  __shared__ double sm[8][K][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pos = (long long)blockIdx.x * 32 + lane;   // Np is a multiple of 128
  double acc[K];
#pragma unroll
  for (int l = 0; l < K; ++l) acc[l] = 0.0;
#pragma unroll 4
  for (int gi = warp; gi < n_ctl; gi += 8) {
    const double x = A[(long long)gi * Np + pos];
#pragma unroll
    for (int l = 0; l < K; ++l) acc[l] = fma(x, U[(long long)gi * K + l], acc[l]);
  }
#pragma unroll
  for (int l = 0; l < K; ++l) sm[warp][l][lane] = acc[l];
  __syncthreads();
  for (int i = threadIdx.x; i < K * 32; i += 256) {
    const int l = i >> 5, c = i & 31;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][l][c];
    Vout[(long long)l * Np + (long long)blockIdx.x * 32 + c] = t;   // pad cells: A = 0 there, so 0
  }
}
// Gram of the K vectors V[l][.] (lower triangle, KK entries), deterministic: one CTA per entry
static __global__ void __launch_bounds__(1024) k_svd_gram(const double* V, long long Np, int K, double* G) {
  __shared__ double sm[1024];
  int e = blockIdx.x, i = 0;
  while ((i + 1) * (i + 2) / 2 <= e) ++i;
  const int j = e - i * (i + 1) / 2;
  const double* a = V + (long long)i * Np; const double* b = V + (long long)j * Np;
  double s = 0.0;
  for (long long p = threadIdx.x; p < Np; p += 1024) s = fma(a[p], b[p], s);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) G[e] = sm[0];
}
// V[.][p] <- V[.][p] T  (T: K x K row-major, applied from the right to the row vector of the K factors of a cell)
static __global__ void k_svd_apply(double* V, long long Np, int K, const double* T) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Np) return;
  double v[RUVNB_MAX_K], o[RUVNB_MAX_K];
  for (int l = 0; l < K; ++l) v[l] = V[(long long)l * Np + p];
  for (int c = 0; c < K; ++c) { double s = 0.0; for (int l = 0; l < K; ++l) s = fma(v[l], T[l * K + c], s); o[c] = s; }
  for (int l = 0; l < K; ++l) V[(long long)l * Np + p] = o[l];
}
// deterministic start: a hash of (cell, factor) mapped to (-1, 1); pads 0
static __global__ void k_svd_start(double* V, long long Np, int K, const int* pos2cell) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np * K) return;
  const int l = (int)(i / Np); const long long p = i - (long long)l * Np;
  const int c = pos2cell[p];
  if (c < 0) { V[i] = 0.0; return; }
  unsigned long long x = (unsigned long long)c * 0x9E3779B97F4A7C15ull + (unsigned long long)(l + 1) * 0xBF58476D1CE4E5B9ull;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
  V[i] = (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// host: symmetric K x K eigen-decomposition by cyclic Jacobi (eigenvalues descending, eigenvectors in the columns of Q)
static void jacobi_eig_h(int K, std::vector<double>& A, std::vector<double>& Q) {
  Q.assign(K * K, 0.0);
  for (int i = 0; i < K; ++i) Q[i * K + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < K; ++i) for (int j = 0; j < i; ++j) off += A[i * K + j] * A[i * K + j];
    if (off < 1e-300) break;
    for (int p = 0; p < K; ++p) for (int q = p + 1; q < K; ++q) {
      double apq = A[p * K + q];
      if (fabs(apq) < 1e-300) continue;
      double tau = (A[q * K + q] - A[p * K + p]) / (2.0 * apq);
      double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
      for (int r = 0; r < K; ++r) { double arp = A[r * K + p], arq = A[r * K + q]; A[r * K + p] = c * arp - s * arq; A[r * K + q] = s * arp + c * arq; }
      for (int r = 0; r < K; ++r) { double apr = A[p * K + r], aqr = A[q * K + r]; A[p * K + r] = c * apr - s * aqr; A[q * K + r] = s * apr + c * aqr; }
      for (int r = 0; r < K; ++r) { double qrp = Q[r * K + p], qrq = Q[r * K + q]; Q[r * K + p] = c * qrp - s * qrq; Q[r * K + q] = s * qrp + c * qrq; }
    }
  }
  // sort descending
  std::vector<int> idx(K);
  for (int i = 0; i < K; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return A[a * K + a] > A[b * K + b]; });
  std::vector<double> Q2(K * K), d(K);
  for (int c = 0; c < K; ++c) { d[c] = A[idx[c] * K + idx[c]]; for (int r = 0; r < K; ++r) Q2[r * K + c] = Q[r * K + idx[c]]; }
  Q = Q2;
  for (int i = 0; i < K; ++i) for (int j = 0; j < K; ++j) A[i * K + j] = (i == j) ? d[i] : 0.0;
}
