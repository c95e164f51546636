// fast.cuh -- the full-data streaming passes of fastruvIII.nb (reference R/fastruvIIInb.R):
//   F5  W for ALL cells from the control genes (:632-716): per cell, lmu = gmean[ctl] + alpha_c W_c,
//       Z.c = alpha_c W_c + (y + .01)/(exp(lmu) + .01) - 1, then the sequential ridge regressions
//       W_j = sum_g wt_g a_gj (Z.c_g - sum_{l<j} a_gl W_l) / (lambda.a + sum_g wt_g a_gj^2)            (:642-649,675-682)
//       = forward substitution on the cell-independent Gram A_jl = sum_g wt_g a_gj a_gl with the per-cell score
//       b_j = sum_g wt_g a_gj Z.c_g; clamp to median +- 4 mad of Wsub (:655-661,688-694); repeated until
//       mean((|W - W.old|/|W.old|)^2) < 1e-5 or 8 sweeps (:709-712).  Cells are independent: the reference's blocks of
//       block.size cells do not change the result.
//   F6  Mb.all (:719-748): annotated cells take Mb[, group]; every other cell starts from Mb = Y (sic, :721) and runs
//       3 element-wise sweeps  lmu = gmean + alpha W + Mb;  Z = Mb + (y + .01)/(exp(lmu) + .01) - 1;
//       Mb = Z wt/(wt + lambda.b), wt = 1/(exp(-lmu) + psi); control genes 0; rows clamped to median +- 4 mad of the
//       annotated columns.  The three sweeps of an element run in registers.
// Included by fast.cu.
#pragma once
#include "kernels.cuh"

// cell-independent ridge Gram of the control genes: A[j][l] = sum_g wt_g a_gj a_gl (lower triangle), one warp
template <int K>
__global__ void k_fast_gram(int n_ctl, const double* alpha_c /* [n_ctl][K] */, const double* wt, double* A /* KK */) {
  constexpr int KK = K * (K + 1) / 2;
  double acc[KK];
#pragma unroll
  for (int i = 0; i < KK; ++i) acc[i] = 0.0;
  for (int g = threadIdx.x; g < n_ctl; g += 32) {
    int idx = 0;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int l = 0; l <= j; ++l) { acc[idx] = fma(wt[g] * alpha_c[(long long)g * K + j], alpha_c[(long long)g * K + l], acc[idx]); ++idx; }
  }
#pragma unroll
  for (int i = 0; i < KK; ++i) { double v = warp_sum(acc[i]); if (threadIdx.x == 0) A[i] = v; }
}

// one sweep over all cells: 8 cells x 4 gene slices per warp (as k_w_update_fast)
template <int K>
__global__ void __launch_bounds__(256) k_fast_w_sweep(const int32_t* const* ctl_rows, int n_ctl, long long Np, const int* chunk_valid,
                                                      const double* gmean_c, const double* alpha_c, const double* wt, const double* A,
                                                      double lambda_a, const double* med, const double* mad, const MathTabs* gtabs,
                                                      const double* W_in, double* W_out, double* crit_part /* [gridDim.x] */) {
  __shared__ MathTabs mt;
  __shared__ double s_crit[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const double* src = reinterpret_cast<const double*>(gtabs);
    double* dst = reinterpret_cast<double*>(&mt);
    for (int i = tid; i < (int)(sizeof(MathTabs) / 8); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  const int cell = lane & 7, slice = lane >> 3;
  const long long pos = ((long long)blockIdx.x * 8 + warp) * 8 + cell;
  const bool valid = pos < Np && (int)(pos & (CH - 1)) < chunk_valid[pos >> 7];
  double wv[K], b[K];
#pragma unroll
  for (int l = 0; l < K; ++l) { wv[l] = valid ? W_in[(long long)l * Np + pos] : 0.0; b[l] = 0.0; }
  if (valid) {
    for (int gi = slice; gi < n_ctl; gi += 4) {
      int y = ctl_rows[gi][pos];
      double aw = 0.0;
#pragma unroll
      for (int l = 0; l < K; ++l) aw = fma(alpha_c[(long long)gi * K + l], wv[l], aw);
      double mu = fexp(gmean_c[gi] + aw, mt.ex);
      double z = aw + ((double)y + 0.01) / (mu + 0.01) - 1.0;
      double wz = wt[gi] * z;
#pragma unroll
      for (int l = 0; l < K; ++l) b[l] = fma(wz, alpha_c[(long long)gi * K + l], b[l]);
    }
  }
#pragma unroll
  for (int l = 0; l < K; ++l) { b[l] += __shfl_xor_sync(FULL, b[l], 8); b[l] += __shfl_xor_sync(FULL, b[l], 16); }
  double cr = 0.0;
  if (valid && slice == 0) {
    double Wn[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double s = b[j];
#pragma unroll
      for (int l = 0; l < j; ++l) s -= A[j * (j + 1) / 2 + l] * Wn[l];
      double v = s / (lambda_a + A[j * (j + 1) / 2 + j]);
      const double hi = med[j] + 4.0 * mad[j], lo = med[j] - 4.0 * mad[j];   // :655-661
      Wn[j] = v;   // the sequential regressions use the unclamped W_l (the clamp comes after the block loop)
      double vc = v > hi ? hi : (v < lo ? lo : v);
      W_out[(long long)j * Np + pos] = vc;
      double d = fabs(vc - wv[j]) / fabs(wv[j]);
      cr += d * d;
    }
  }
  cr = warp_sum(cr);
  if (lane == 0) s_crit[warp] = cr;
  __syncthreads();
  if (tid == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += s_crit[w]; crit_part[blockIdx.x] = t; }
}

// F6: 32 genes x 32 cells per CTA, transposed store into the block slab [B][Gl]
struct FastMbArgs {
  int Gl, K, m, B;
  long long Np, c0;
  const int32_t* Yp;
  const double *alpha, *gmean, *psi, *W, *Mb;   // Mb: [Gl][m] of the subsample fit
  const int *cell2pos, *annot;                  // annot[c]: replicate group of cell c or -1
  const uint8_t* ctl_local;                     // [Gl]
  const double *rmed, *rmad;                    // [Gl] row median / mad of the annotated columns
  double lambda_b;
  int gene_major, ldB;                          // output [Gl][ldB] instead of the R-layout slab [B][Gl]
};
static __global__ void __launch_bounds__(1024) k_fast_mb(FastMbArgs a, double* slab) {
  __shared__ double tile[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x * 32 + warp, c = blockIdx.y * 32 + lane;
  double v = 0.0;
  if (g < a.Gl && c < a.B) {
    const long long cell = a.c0 + c, pos = a.cell2pos[cell];
    const int grp = a.annot[cell];
    if (a.ctl_local[g]) v = 0.0;                                             // :741
    else if (grp >= 0) v = a.Mb[(long long)g * a.m + grp];                   // :723,740
    else {
      const double y = (double)a.Yp[(long long)g * a.Np + pos];
      double base = a.gmean[g];
      for (int l = 0; l < a.K; ++l) base = fma(a.alpha[(long long)g * a.K + l], a.W[(long long)l * a.Np + pos], base);
      const double psi = a.psi[g];
      double mb = y;                                                         // Mb.all <- Y (:721)
      for (int it = 0; it < 3; ++it) {                                       // :727-738
        double lmu = base + mb;
        double Z = mb + ((y + 0.01) / (exp(lmu) + 0.01) - 1.0);
        double wt = 1.0 / (exp(-lmu) + psi);
        mb = Z * (wt / (wt + a.lambda_b));
      }
      v = mb;
    }
    const double hi = a.rmed[g] + 4.0 * a.rmad[g], lo = a.rmed[g] - 4.0 * a.rmad[g];   // clamp every row (:744-748)
    if (v > hi) v = hi;
    if (v < lo) v = lo;
  }
  if (a.gene_major) {   // the block stays on the device for get.res: [Gl][ldB], coalesced along the cells
    if (g < a.Gl && c < a.B) slab[(long long)g * a.ldB + c] = v;
    return;
  }
  tile[warp][lane] = v;
  __syncthreads();
  const int g2 = blockIdx.x * 32 + lane, c2 = blockIdx.y * 32 + warp;
  if (g2 < a.Gl && c2 < a.B) slab[(long long)c2 * a.Gl + g2] = tile[lane][warp];
}
