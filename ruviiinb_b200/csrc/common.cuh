// common.cuh -- shared device helpers for libruvnb_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <stdint.h>
#include <math.h>
#include "fastmath.cuh"

#define CH 128          // cells per chunk: one warp-wide pass of 4 cells per lane
#define CHPAD 4         // the chunk count is padded to a multiple of this (tiles of up to 4 chunks per stage)
#define NWARP 8         // warps (= gene rows) per CTA in the row-streaming kernels
#define YT 128          // per-gene count tables cover y < YT (larger counts take the libdevice path)
#define ROW_THREADS (NWARP * 32)
#define LOG_DBL_MIN (-708.3964185322641)  // log(.Machine$double.xmin), R/ruvIIInb.R:290

#define FULL 0xffffffffu

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Order-preserving map double -> uint64 (total order; NaN sorts above +Inf, callers count NaN apart).
__device__ __forceinline__ unsigned long long f64_key(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// Position of the 4 cells a lane owns inside a 128-cell chunk: {2l, 2l+1, 64+2l, 65+2l}.
// Two 8-byte loads per lane are each a contiguous 256-byte warp access, and the matching double2
// reads of the staged W tile are bank-conflict free (16-byte lane stride).
__device__ __forceinline__ int lane_cell(int lane, int e) { return ((e >> 1) << 6) + 2 * lane + (e & 1); }

// ---- per-element NB working quantities (R/ruvIIInb.R:303-321) -----------------------------------
// signed deviance residual: sign(y-mu) sqrt(2 (l_sat - l)); the lgamma terms of the two dnbinom
// calls cancel, leaving 2 [ y log(y/mu) - (y+r) log((y+r)/(mu+r)) ].
__device__ __forceinline__ double nb_signdev(int y, double lmu, double mu, double r) {
  double yd = (double)y;
  double L = log(mu + r);
  double t = -(yd + r) * (log(yd + r) - L);
  if (y > 0) t += yd * (log(yd) - lmu);
  double d2 = 2.0 * t;
  double d = sqrt(fmax(d2, 0.0));
  if (d2 != d2) d = d2;  // keep NaN
  return (yd > mu) ? d : ((yd < mu) ? -d : (mu != mu ? mu : 0.0));
}
// sig.inv = 1/(exp(-lmu) + psi)
__device__ __forceinline__ double nb_siginv(double lmu, double psi) { return 1.0 / (exp(-lmu) + psi); }
// log dnbinom(y; mu, size = r) with +-Inf/NaN -> log(DBL_MIN) (R/ruvIIInb.R:286,290)
__device__ __forceinline__ double nb_logpmf(int y, double lmu, double mu, double r, double lgr, double rlogr) {
  double yd = (double)y;
  double L = log(mu + r);
  double v;
  if (y == 0) {
    v = rlogr - r * L;
  } else {
    v = lgamma(yd + r) - lgr - lgamma(yd + 1.0) + rlogr - r * L + yd * (lmu - L);
  }
  if (!(fabs(v) <= 1.79769313486231570e308)) v = LOG_DBL_MIN;  // NaN or +-Inf
  return v;
}

// ---- table-driven per-element quantities ----------------------------------
// ly: log(y) for y < YT (MathTabs.ly, shared by all genes); tr: the gene's log(y + r) for y < YT.
// Half deviance  y (log y - lmu) - (y + r)(log(y + r) - L),  L = log(mu + r): the same conditioned form
// as nb_signdev above with the two logs of integer arguments looked up.
// out-of-table counts (y >= YT): table-driven logs + Stirling (lgamma_big); libdevice only outside their range
__device__ __forceinline__ void nb_big_logs(double yd, double r, const MathTabs* mt, double* lyr, double* lyy) {
  *lyr = flog(yd + r, mt->rc, mt->lc);
  *lyy = flog(yd, mt->rc, mt->lc);
}
static __device__ __noinline__ double nb_slow_lgterm(double yd, double r, double lgr) { return lgamma(yd + r) - lgr - lgamma(yd + 1.0); }
__device__ __forceinline__ double nb_big_lgterm(double yd, double r, double lgr, const MathTabs* mt) {
  if (!(r < 1e100 && yd < 1e15)) return nb_slow_lgterm(yd, r, lgr);
  double a = yd + r, b = yd + 1.0;
  return lgamma_big(a, flog_nc(a, mt->rc, mt->lc)) - lgr - lgamma_big(b, flog_nc(b, mt->rc, mt->lc));
}
__device__ __forceinline__ double nb_devhalf_t(int y, double yd, double lmu, double L, double r,
                                               const double* __restrict__ tr, const MathTabs* mt) {
  double lyr, lyy;
  if (y < YT) { lyr = tr[y]; lyy = mt->ly[y]; } else { nb_big_logs(yd, r, mt, &lyr, &lyy); }
  double t = -(yd + r) * (lyr - L);
  if (y > 0) t = fma(yd, lyy - lmu, t);
  return t;
}
// signed deviance residual from the half deviance (same sign / NaN conventions as nb_signdev)
__device__ __forceinline__ double nb_signdev_from_half(double t, double yd, double mu) {
  double d2 = 2.0 * t;
  double d = sqrt(fmax(d2, 0.0));
  if (d2 != d2) d = d2;
  return (yd > mu) ? d : ((yd < mu) ? -d : (mu != mu ? mu : 0.0));
}
// MASS::psi.huber weight pmin(1, k sigma / |d|) from the half deviance, NaN-propagating like R's pmin
__device__ __forceinline__ double huber_w_half(double khsig, double t) {
  double d2 = 2.0 * t;
  double q = khsig * rsqrt(fmax(d2, 0.0));
  if (d2 != d2) q = d2;
  return (q < 1.0) ? q : ((q != q) ? q : 1.0);
}
// log dnbinom(y; mu, size = r): tl = the gene's lgamma(y+r) - lgamma(r) - lgamma(y+1) for y < YT
__device__ __forceinline__ double nb_logpmf_t(int y, double yd, double lmu, double L, double r, double lgr, double rlogr,
                                              const double* __restrict__ tl, const MathTabs* mt) {
  double v = fma(-r, L, rlogr);
  if (y > 0) {
    double g = (y < YT) ? tl[y] : nb_big_lgterm(yd, r, lgr, mt);
    v += fma(yd, lmu - L, g);
  }
  if (!(fabs(v) <= 1.79769313486231570e308)) v = LOG_DBL_MIN;  // NaN or +-Inf (R/ruvIIInb.R:290)
  return v;
}
// sig.inv = 1/(exp(-lmu) + psi) = mu/(1 + psi mu) and the working-response ratio (y + .01)/(mu + .01)
// (R/ruvIIInb.R:305,321) with ONE reciprocal: both share the denominator (1 + psi mu)(mu + .01).
struct SQ { double mu, s, q; };
// the reference's own formulas with libdevice: out-of-range arguments (rare), kept out of line
static __device__ __noinline__ SQ nb_sq_slow(double yd, double lmu, double psi) {
  SQ o;
  o.mu = exp(lmu);
  o.s = 1.0 / (1.0 / o.mu + psi);
  o.q = (yd + 0.01) / (o.mu + 0.01);
  return o;
}
// fast form; the caller guarantees |lmu| < 230 and psi in [1e-100, 1e100]
__device__ __forceinline__ void nb_sq_fast(double yd, double lmu, double psi, const double* __restrict__ ex, double* mu,
                                           double* s, double* q) {
  double m = fexp_nc(lmu, ex);
  double a = fma(psi, m, 1.0), b = m + 0.01;
  double inv = frcp(a * b);
  *mu = m;
  *s = (m * b) * inv;
  *q = ((yd + 0.01) * a) * inv;
}
__device__ __forceinline__ void nb_siginv_ratio(double yd, double mu, double psi, double* s, double* q) {
  if (!(mu < 1e100)) {  // overflowed / NaN mu: the reference's own formulas (rare, warp-divergent)
    *s = 1.0 / (1.0 / mu + psi);
    *q = (yd + 0.01) / (mu + 0.01);
    return;
  }
  double a = fma(psi, mu, 1.0), b = mu + 0.01;
  double inv = frcp(a * b);
  *s = (mu * b) * inv;
  *q = ((yd + 0.01) * a) * inv;
}
