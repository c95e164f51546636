// fastfit.cuh -- F3/F4: the subsample fit of fastruvIII.nb (reference R/fastruvIIInb.R:223-623), nbatch == 1.
// The subsample is small by construction (<= 3000 annotated cells, :118-123), so the working response Z and the NB weight
// sig.inv are materialised once per iteration as dense [Gl][Np] doubles (the reference holds them too) and the remaining
// steps are plain per-gene / per-cell reductions over those two matrices.  The counts sit in the same permuted, padded
// gene-major tiles as everywhere else (cells of a replicate group contiguous).
//   k_ff_init_z     Z = log(Y + 1), gmean = rowMeans(Z), per-group means                                   (:224-227)
//   k_ff_working    lmu, sig.inv = 1/(exp(-lmu) + psi), Z = lmu + step_c ((y+.01)/(mu+.01) - 1); group sums of Z; sum_c sig.inv (:357-379,331)
//   k_ff_colmean    wt.cell = colMeans(sig.inv)                                                               (:405)
//   k_ff_b          b_g = sum_c RM_Z[g,c] wt.cell[c] RM_W[c,]   (RM_Z = Z - group mean)                       (:386-387,410)
//   k_ff_alpha      alpha_g = b_g Ginv  (Ginv = (RM_W' D RM_W + lambda.a I)^-1, k x k, from the host)         (:411)
//   k_ff_w          Wsub: sequential ridge regressions on the control genes with gene weights wt.ctl           (:443-457)
//   k_ff_gsums      S0[g,j] = sum_{c in j} sig.inv, S1[g,j] = sum sig.inv (Z - alpha W): gmean / Mb by k_gmean_mb (:478-490)
//   k_ff_cell_loglik  per-CELL log-likelihood sum_g log dnbinom (the fast fit halves steps per cell)          (:334-346,509-531)
// Included by fast.cu.
#pragma once
#include "kernels.cuh"

struct FFGeom {
  const int32_t* Yp; long long Np; int Gl, K, m;
  const int *chunk_grp, *chunk_valid;
};
__device__ __forceinline__ bool ff_valid(const FFGeom& g, long long pos) { return (int)(pos & (CH - 1)) < g.chunk_valid[pos >> 7]; }

// Z = log(y + 1) -> Zbuf; gmean = rowMeans(Z); ZS[g][j] = group sums.  One warp per gene.
static __global__ void __launch_bounds__(256) k_ff_init_z(FFGeom ge, double n, double* Zbuf, double* gmean, double* ZS) {
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= ge.Gl) return;
  const int32_t* y = ge.Yp + (long long)g * ge.Np;
  double* z = Zbuf + (long long)g * ge.Np;
  double tot = 0.0;
  for (int j = 0; j < ge.m; ++j) ZS[(long long)g * ge.m + j] = 0.0;
  __syncwarp();
  int curj = -1; double gs = 0.0;
  const int nch = (int)(ge.Np >> 7);
  for (int c = 0; c < nch; ++c) {
    const int nv = ge.chunk_valid[c], j = ge.chunk_grp[c];
    if (nv == 0) continue;
    if (j != curj) {
      if (curj >= 0) { double v = warp_sum(gs); if (lane == 0) ZS[(long long)g * ge.m + curj] = v; }
      curj = j; gs = 0.0;
    }
    for (int i = lane; i < CH; i += 32) {
      long long pos = (long long)c * CH + i;
      double v = 0.0;
      if (i < nv) { v = log((double)y[pos] + 1.0); gs += v; tot += v; }
      z[pos] = v;
    }
  }
  if (curj >= 0) { double v = warp_sum(gs); if (lane == 0) ZS[(long long)g * ge.m + curj] = v; }
  tot = warp_sum(tot);
  if (lane == 0) gmean[g] = tot / n;
}

// working quantities of the current parameters; one warp per gene
static __global__ void __launch_bounds__(256) k_ff_working(FFGeom ge, const double* gmean, const double* Mb, const double* alpha, const double* W,
                                                   const double* psi, const double* step_c, double* Zbuf, double* SIbuf, double* ZS,
                                                   double* si_rowsum) {
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= ge.Gl) return;
  const int32_t* y = ge.Yp + (long long)g * ge.Np;
  double* z = Zbuf + (long long)g * ge.Np; double* si = SIbuf + (long long)g * ge.Np;
  const double gm = gmean[g], ps = psi[g];
  double a[RUVNB_MAX_K];
  for (int l = 0; l < ge.K; ++l) a[l] = alpha[(long long)g * ge.K + l];
  int curj = -1; double gs = 0.0, sis = 0.0;
  const int nch = (int)(ge.Np >> 7);
  for (int c = 0; c < nch; ++c) {
    const int nv = ge.chunk_valid[c], j = ge.chunk_grp[c];
    if (nv == 0) continue;
    if (j != curj) {
      if (curj >= 0) { double v = warp_sum(gs); if (lane == 0) ZS[(long long)g * ge.m + curj] = v; }
      curj = j; gs = 0.0;
    }
    const double mb = Mb[(long long)g * ge.m + j];
    for (int i = lane; i < CH; i += 32) {
      long long pos = (long long)c * CH + i;
      double zv = 0.0, sv = 0.0;
      if (i < nv) {
        double lmu = gm + mb;
        for (int l = 0; l < ge.K; ++l) lmu = fma(a[l], W[(long long)l * ge.Np + pos], lmu);
        double mu = exp(lmu);
        sv = 1.0 / (exp(-lmu) + ps);
        zv = lmu + step_c[pos] * (((double)y[pos] + 0.01) / (mu + 0.01) - 1.0);
        gs += zv; sis += sv;
      }
      z[pos] = zv; si[pos] = sv;
    }
  }
  if (curj >= 0) { double v = warp_sum(gs); if (lane == 0) ZS[(long long)g * ge.m + curj] = v; }
  sis = warp_sum(sis);
  if (lane == 0) si_rowsum[g] = sis;
}

// wt.cell[pos] = mean over genes of sig.inv; one thread per position (coalesced across positions)
static __global__ void k_ff_colmean(FFGeom ge, const double* SIbuf, double* wt_cell) {
  long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= ge.Np) return;
  double s = 0.0;
  if (ff_valid(ge, pos)) { for (int g = 0; g < ge.Gl; ++g) s += SIbuf[(long long)g * ge.Np + pos]; s /= (double)ge.Gl; }
  wt_cell[pos] = s;
}
static __global__ void k_ff_cap(double* x, long long n, double cap) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && x[i] >= cap) x[i] = cap;
}

// b[g][l] = sum_c (Z - Zbar[g][grp c]) wt.cell[c] RM_W[c][l]; one warp per gene
static __global__ void __launch_bounds__(256) k_ff_b(FFGeom ge, const double* Zbuf, const double* ZS, const int* group_n, const double* wt_cell,
                                             const double* RMW, double* b) {
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= ge.Gl) return;
  const double* z = Zbuf + (long long)g * ge.Np;
  double acc[RUVNB_MAX_K];
  for (int l = 0; l < ge.K; ++l) acc[l] = 0.0;
  const int nch = (int)(ge.Np >> 7);
  for (int c = 0; c < nch; ++c) {
    const int nv = ge.chunk_valid[c], j = ge.chunk_grp[c];
    if (nv == 0) continue;
    const double zbar = ZS[(long long)g * ge.m + j] / (double)group_n[j];
    for (int i = lane; i < nv; i += 32) {
      long long pos = (long long)c * CH + i;
      double v = (z[pos] - zbar) * wt_cell[pos];
      for (int l = 0; l < ge.K; ++l) acc[l] = fma(v, RMW[(long long)l * ge.Np + pos], acc[l]);
    }
  }
  for (int l = 0; l < ge.K; ++l) { double v = warp_sum(acc[l]); if (lane == 0) b[(long long)g * ge.K + l] = v; }
}
// alpha_g = b_g Ginv (Ginv symmetric k x k, row-major); counts non-finite entries
static __global__ void k_ff_alpha(int Gl, int K, const double* b, const double* Ginv, double* alpha, int* bad) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  int nb = 0;
  for (int l = 0; l < K; ++l) {
    double s = 0.0;
    for (int q = 0; q < K; ++q) s = fma(b[(long long)g * K + q], Ginv[q * K + l], s);
    alpha[(long long)g * K + l] = s;
    if (!(fabs(s) <= 1.79769313486231570e308)) nb++;
  }
  if (nb) atomicAdd(bad, nb);
}

// Wsub for every subsample cell from the stored Z of the control genes (:443-457): 8 cells x 4 gene slices per warp
template <int K>
__global__ void __launch_bounds__(256) k_ff_w(FFGeom ge, const int* ctl_local, int n_ctl, const double* Zbuf, const double* gmean_old,
                                             const double* alpha_c /* [n_ctl][K] new */, const double* wt_ctl, const double* A /* KK */,
                                             double lambda_a, double* W_out, int* bad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cell = lane & 7, slice = lane >> 3;
  const long long pos = ((long long)blockIdx.x * 8 + warp) * 8 + cell;
  const bool valid = pos < ge.Np && ff_valid(ge, pos);
  double b[K];
#pragma unroll
  for (int l = 0; l < K; ++l) b[l] = 0.0;
  if (valid) {
    for (int gi = slice; gi < n_ctl; gi += 4) {
      const int g = ctl_local[gi];
      double wz = wt_ctl[gi] * (Zbuf[(long long)g * ge.Np + pos] - gmean_old[g]);
#pragma unroll
      for (int l = 0; l < K; ++l) b[l] = fma(wz, alpha_c[(long long)gi * K + l], b[l]);
    }
  }
#pragma unroll
  for (int l = 0; l < K; ++l) { b[l] += __shfl_xor_sync(FULL, b[l], 8); b[l] += __shfl_xor_sync(FULL, b[l], 16); }
  if (slice == 0 && pos < ge.Np) {
    double Wn[K]; int nb = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double s = b[j];
#pragma unroll
      for (int l = 0; l < j; ++l) s -= A[j * (j + 1) / 2 + l] * Wn[l];
      Wn[j] = s / (lambda_a + A[j * (j + 1) / 2 + j]);
      W_out[(long long)j * ge.Np + pos] = valid ? Wn[j] : 0.0;
      if (valid && !(fabs(Wn[j]) <= 1.79769313486231570e308)) nb++;
    }
    if (nb) atomicAdd(bad, nb);
  }
}

// S0[g][j] = sum_{c in j} sig.inv, S1[g][j] = sum sig.inv (Z - alpha_new W_new); one warp per gene
static __global__ void __launch_bounds__(256) k_ff_gsums(FFGeom ge, const double* Zbuf, const double* SIbuf, const double* alpha_new,
                                                 const double* W_new, double* S0, double* S1) {
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= ge.Gl) return;
  const double* z = Zbuf + (long long)g * ge.Np; const double* si = SIbuf + (long long)g * ge.Np;
  double a[RUVNB_MAX_K];
  for (int l = 0; l < ge.K; ++l) a[l] = alpha_new[(long long)g * ge.K + l];
  for (int j = lane; j < ge.m; j += 32) { S0[(long long)g * ge.m + j] = 0.0; S1[(long long)g * ge.m + j] = 0.0; }
  __syncwarp();
  int curj = -1; double s0 = 0.0, s1 = 0.0;
  const int nch = (int)(ge.Np >> 7);
  for (int c = 0; c < nch; ++c) {
    const int nv = ge.chunk_valid[c], j = ge.chunk_grp[c];
    if (nv == 0) continue;
    if (j != curj) {
      if (curj >= 0) { double v0 = warp_sum(s0), v1 = warp_sum(s1); if (lane == 0) { S0[(long long)g * ge.m + curj] = v0; S1[(long long)g * ge.m + curj] = v1; } }
      curj = j; s0 = 0.0; s1 = 0.0;
    }
    for (int i = lane; i < nv; i += 32) {
      long long pos = (long long)c * CH + i;
      double r = z[pos];
      for (int l = 0; l < ge.K; ++l) r = fma(-a[l], W_new[(long long)l * ge.Np + pos], r);
      s0 += si[pos]; s1 = fma(si[pos], r, s1);
    }
  }
  if (curj >= 0) { double v0 = warp_sum(s0), v1 = warp_sum(s1); if (lane == 0) { S0[(long long)g * ge.m + curj] = v0; S1[(long long)g * ge.m + curj] = v1; } }
}
static __global__ void k_ff_zero_ctl_rows(int Gl, int m, const uint8_t* ctl_local, double* Mb) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Gl * m) return;
  if (ctl_local[i / m]) Mb[i] = 0.0;
}
// unweighted group means of the residual for the initial Mb (:270: projection(..., Wmat = NULL) ignores lambda)
static __global__ void __launch_bounds__(256) k_ff_init_mb(FFGeom ge, const double* Zbuf, const double* gmean, const double* alpha, const double* W,
                                                   const int* group_n, const uint8_t* ctl_local, double* Mb) {
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= ge.Gl) return;
  const double* z = Zbuf + (long long)g * ge.Np;
  double a[RUVNB_MAX_K];
  for (int l = 0; l < ge.K; ++l) a[l] = alpha[(long long)g * ge.K + l];
  const double gm = gmean[g];
  int curj = -1; double s = 0.0;
  const int nch = (int)(ge.Np >> 7);
  auto flush = [&](int j) { double v = warp_sum(s); if (lane == 0) Mb[(long long)g * ge.m + j] = ctl_local[g] ? 0.0 : v / (double)group_n[j]; };
  for (int c = 0; c < nch; ++c) {
    const int nv = ge.chunk_valid[c], j = ge.chunk_grp[c];
    if (nv == 0) continue;
    if (j != curj) { if (curj >= 0) flush(curj); curj = j; s = 0.0; }
    for (int i = lane; i < nv; i += 32) {
      long long pos = (long long)c * CH + i;
      double r = z[pos] - gm;
      for (int l = 0; l < ge.K; ++l) r = fma(-a[l], W[(long long)l * ge.Np + pos], r);
      s += r;
    }
  }
  if (curj >= 0) flush(curj);
}

// per-CELL log-likelihood sum_g log dnbinom(y; mu, 1/psi_g), non-finite -> log(DBL_MIN); one thread per position
static __global__ void k_ff_cell_loglik(FFGeom ge, const double* gmean, const double* Mb, const double* alpha, const double* W, const double* psi,
                                 const double* tabL /* [Gl][YT] */, const double* lgr, double* out) {
  long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= ge.Np) return;
  double acc = 0.0;
  if (ff_valid(ge, pos)) {
    const int j = ge.chunk_grp[pos >> 7];
    double w[RUVNB_MAX_K];
    for (int l = 0; l < ge.K; ++l) w[l] = W[(long long)l * ge.Np + pos];
    for (int g = 0; g < ge.Gl; ++g) {
      const int y = ge.Yp[(long long)g * ge.Np + pos];
      double lmu = gmean[g] + Mb[(long long)g * ge.m + j];
      for (int l = 0; l < ge.K; ++l) lmu = fma(alpha[(long long)g * ge.K + l], w[l], lmu);
      const double r = 1.0 / psi[g], mu = exp(lmu), L = log(mu + r), yd = (double)y;
      double v = r * log(r) - r * L;
      if (y > 0) v += ((y < YT) ? tabL[(long long)g * YT + y] : (lgamma(yd + r) - lgr[g] - lgamma(yd + 1.0))) + yd * (lmu - L);
      if (!(fabs(v) <= 1.79769313486231570e308)) v = LOG_DBL_MIN;
      acc += v;
    }
  }
  out[pos] = acc;
}
// step[c] *= fac where loglik_old[c] > loglik_new[c] (per cell, :538-539)
static __global__ void k_ff_halve(long long Np, const double* ll_old, const double* ll_new, double fac, double* step) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np && ll_old[i] > ll_new[i]) step[i] *= fac;
}
