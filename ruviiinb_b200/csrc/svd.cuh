// svd.cuh -- H1: the initial W of ruvIII.nb (reference R/ruvIIInb.R:160-168),
//   W0 = irlba(log(Y[ctl,]/rowMeans(Y[ctl,]) + 1), nv = k)$v  -- the k leading right singular vectors of the n_ctl x N matrix A.
// Subspace iteration on A'A with the matrix never formed: V <- orth(A'(A V)), A[g,c] recomputed from the resident control
// rows in both products; then one Rayleigh-Ritz rotation.  irlba's Lanczos start is random and its tolerance is 1e-5, so the
// reference's own W0 is only defined up to that (and up to sign); parity runs inject W0.
// Sign convention: the entry of largest magnitude of every vector is positive.  Included by ruvnb.cu only.
#pragma once
#include "kernels.cuh"

// rowMeans(Y[ctl,]) over the real cells: one warp per control row
__global__ void __launch_bounds__(256) k_svd_rowmean(const int32_t* const* ctl_rows, int n_ctl, long long Np, double n, double* rowmean) {
  const int gi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gi >= n_ctl) return;
  const int32_t* y = ctl_rows[gi];
  double s = 0.0;
  for (long long p = lane; p < Np; p += 32) { int v = y[p]; if (v > 0) s += (double)v; }   // pads are -1
  s = warp_sum(s);
  if (lane == 0) rowmean[gi] = s / n;
}
// U[gi][l] = sum_c A[gi,c] V[l][c]; one warp per control row
__global__ void __launch_bounds__(256) k_svd_av(const int32_t* const* ctl_rows, int n_ctl, long long Np, int K, const double* rowmean,
                                               const double* V, double* U) {
  const int gi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gi >= n_ctl) return;
  const int32_t* y = ctl_rows[gi];
  const double irm = 1.0 / rowmean[gi];
  double acc[RUVNB_MAX_K];
  for (int l = 0; l < K; ++l) acc[l] = 0.0;
  for (long long p = lane; p < Np; p += 32) {
    int v = y[p];
    if (v <= 0) continue;                       // log(0/rm + 1) = 0; pads skipped
    double a = log1p((double)v * irm);
    for (int l = 0; l < K; ++l) acc[l] = fma(a, V[(long long)l * Np + p], acc[l]);
  }
  for (int l = 0; l < K; ++l) { double s = warp_sum(acc[l]); if (lane == 0) U[(long long)gi * K + l] = s; }
}
// Vout[l][p] = sum_gi A[gi,p] U[gi][l]; 8 cells x 4 gene slices per warp
template <int K>
__global__ void __launch_bounds__(256) k_svd_atu(const int32_t* const* ctl_rows, int n_ctl, long long Np, const int* chunk_valid,
                                                const double* rowmean, const double* U, double* Vout) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cell = lane & 7, slice = lane >> 3;
  const long long pos = ((long long)blockIdx.x * 8 + warp) * 8 + cell;
  const bool valid = pos < Np && (int)(pos & (CH - 1)) < chunk_valid[pos >> 7];
  double acc[K];
#pragma unroll
  for (int l = 0; l < K; ++l) acc[l] = 0.0;
  if (valid) {
    for (int gi = slice; gi < n_ctl; gi += 4) {
      int v = ctl_rows[gi][pos];
      if (v <= 0) continue;
      double a = log1p((double)v / rowmean[gi]);
#pragma unroll
      for (int l = 0; l < K; ++l) acc[l] = fma(a, U[(long long)gi * K + l], acc[l]);
    }
  }
#pragma unroll
  for (int l = 0; l < K; ++l) { acc[l] += __shfl_xor_sync(FULL, acc[l], 8); acc[l] += __shfl_xor_sync(FULL, acc[l], 16); }
  if (slice == 0 && pos < Np) {
#pragma unroll
    for (int l = 0; l < K; ++l) Vout[(long long)l * Np + pos] = valid ? acc[l] : 0.0;
  }
}
// Gram of the K vectors V[l][.] (lower triangle, KK entries), deterministic: one CTA per entry
__global__ void __launch_bounds__(1024) k_svd_gram(const double* V, long long Np, int K, double* G) {
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
__global__ void k_svd_apply(double* V, long long Np, int K, const double* T) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Np) return;
  double v[RUVNB_MAX_K], o[RUVNB_MAX_K];
  for (int l = 0; l < K; ++l) v[l] = V[(long long)l * Np + p];
  for (int c = 0; c < K; ++c) { double s = 0.0; for (int l = 0; l < K; ++l) s = fma(v[l], T[l * K + c], s); o[c] = s; }
  for (int l = 0; l < K; ++l) V[(long long)l * Np + p] = o[l];
}
// deterministic start: a hash of (cell, factor) mapped to (-1, 1); pads 0
__global__ void k_svd_start(double* V, long long Np, int K, const int* pos2cell) {
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
