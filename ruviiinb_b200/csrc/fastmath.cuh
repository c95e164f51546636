// fastmath.cuh -- table-driven FP64 exp / log / reciprocal for the row-streaming kernels.
//
// The IRLS sweeps are bound by the FP64 pipe, and exp/log dominate the per-element
// instruction count when taken from libdevice (degree-11 / degree-~20 polynomials).  These versions move
// part of the range reduction into 128-entry tables staged in shared memory:
//   fexp(x):  x = (128 e + j) ln2/128 + r, |r| <= ln2/256;  exp(x) = 2^e * T[j] * (1 + r q(r)), q degree 4
//             (truncation r^6/720 < 6e-19): 10 FP64 instructions, relative error < 3e-16.
//   flog(v):  v = 2^e m, m in [1,2), j = top 7 mantissa bits, u = m * RC[j] - 1, |u| <= 2^-8;
//             log v = e ln2 + LC[j] + u q(u), q degree 5 (truncation u^7/7 < 2e-18): 10 FP64 instructions,
//             absolute error < 3e-16 (1 + |e|/8).
//   frcp(x):  MUFU.RCP64H seed (20 bits) + one cubic Newton step: 3 FMA, relative error < 3e-16.
// Everything is __host__ __device__ so tests/test_fastmath.py checks the SAME source on the CPU against
// libm (ruvnb_selftest_math in core.cu); arguments outside the fast range fall back to exp()/log().
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define FM_NT 128
struct MathTabs {
  double ex[FM_NT];  // 2^(j/128)
  double rc[FM_NT];  // ~ 1 / (1 + (j + 0.5)/128)
  double lc[FM_NT];  // -log(rc[j])
  double ly[FM_NT];  // log(j), ly[0] = 0  (log of an integer count)
};

static inline void build_math_tabs(MathTabs* t) {
  for (int j = 0; j < FM_NT; ++j) {
    t->ex[j] = (double)exp2l((long double)j / 128.0L);
    double c = 1.0 + ((double)j + 0.5) / 128.0;
    t->rc[j] = 1.0 / c;
    t->lc[j] = (double)(-logl((long double)t->rc[j]));
    t->ly[j] = j ? (double)logl((long double)j) : 0.0;
  }
}

#if defined(__CUDA_ARCH__)
#define FM_HI(x) __double2hiint(x)
#define FM_LO(x) __double2loint(x)
#define FM_MK(h, l) __hiloint2double(h, l)
#else
static inline int fm_hi_(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int fm_lo_(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffll); }
static inline double fm_mk_(int h, int l) { int64_t b = (int64_t)(((uint64_t)(uint32_t)h << 32) | (uint32_t)l); double x; memcpy(&x, &b, 8); return x; }
#define FM_HI(x) fm_hi_(x)
#define FM_LO(x) fm_lo_(x)
#define FM_MK(h, l) fm_mk_(h, l)
#endif

#if defined(__CUDACC__)
#define FM_NOINLINE __noinline__
#else
#define FM_NOINLINE
#endif
__host__ __device__ FM_NOINLINE static double fm_slow_exp(double x) { return exp(x); }
__host__ __device__ FM_NOINLINE static double fm_slow_log(double v) { return log(v); }

__host__ __device__ __forceinline__ double fexp_nc(double x, const double* __restrict__ ex);
__host__ __device__ __forceinline__ double fexp(double x, const double* __restrict__ ex) {
  if (!(fabs(x) < 700.0)) return fm_slow_exp(x);  // overflow / underflow / NaN: libm path (rare), out of line
  return fexp_nc(x, ex);
}
// no range check: the caller guarantees |x| < 700
__host__ __device__ __forceinline__ double fexp_nc(double x, const double* __restrict__ ex) {
  const double MAGIC = 6755399441055744.0;     // 1.5 * 2^52: the integer lands in the low word
  const double L2E128 = 184.6649652337873;     // 128 / ln 2
  const double C_HI = 0.005415212333900854;    // ln2/128 with the low 24 mantissa bits zero (0x3f762e42fe000000)
  const double C_LO = 1.4223718738313642e-11;  // ln2/128 - C_HI
  double kd = fma(x, L2E128, MAGIC);
  int n = FM_LO(kd);                           // |n| < 2^17
  double kf = kd - MAGIC;
  double r = fma(kf, -C_HI, x);                // exact: kf has <= 18 bits, C_HI 29
  r = fma(kf, -C_LO, r);
  double s = ex[n & (FM_NT - 1)];
#ifdef FM_ESTRIN
  // Estrin: e^r - 1 = r + r^2 (1/2 + r/6 + r^2 (1/24 + r/120)); the same six operations as Horner, four levels deep instead of six
  const double r2 = r * r;
  const double qa = fma(r, 1.0 / 6.0, 0.5), qb = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  double v = fma(s, fma(r2, fma(r2, qb, qa), r), s);
#else
  double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  q = fma(r, q, 1.0 / 6.0);
  q = fma(r, q, 0.5);
  q = fma(r, q, 1.0);
  double v = fma(s, r * q, s);                 // 2^(j/128) e^r, in [1, 2.01)
#endif
  int e = n >> 7;                              // floor division; |e| <= 1010 -> result stays normal
  return FM_MK(FM_HI(v) + (e << 20), FM_LO(v));
}

// log(1 + u) / u for |u| <= 2^-8: 1 - u/2 + u^2/3 - u^3/4 + u^4/5 - u^5/6
__host__ __device__ __forceinline__ double fm_log_poly(double u) {
#ifdef FM_ESTRIN
  // Estrin: (1 - u/2) + u^2 (1/3 - u/4) + u^4 (1/5 - u/6): seven operations three levels deep (Horner: five, five deep)
  const double u2 = u * u;
  const double a = fma(u, -0.5, 1.0), b = fma(u, -0.25, 1.0 / 3.0), c = fma(u, -1.0 / 6.0, 0.2);
  return fma(u2 * u2, c, fma(u2, b, a));
#else
  double q = fma(u, -1.0 / 6.0, 0.2);
  q = fma(u, q, -0.25);
  q = fma(u, q, 1.0 / 3.0);
  q = fma(u, q, -0.5);
  return fma(u, q, 1.0);
#endif
}
__host__ __device__ __forceinline__ double flog_nc(double v, const double* __restrict__ rc, const double* __restrict__ lc);
__host__ __device__ __forceinline__ double flog(double v, const double* __restrict__ rc, const double* __restrict__ lc) {
  if (!(v > 1e-300 && v < 1e300)) return fm_slow_log(v);  // zero, subnormal, negative, Inf, NaN: libm path
  return flog_nc(v, rc, lc);
}
// no range check: the caller guarantees a normal positive finite v
__host__ __device__ __forceinline__ double flog_nc(double v, const double* __restrict__ rc, const double* __restrict__ lc) {
  int hi = FM_HI(v);
  int e = (hi >> 20) - 1023;
  int j = (hi >> 13) & (FM_NT - 1);
  double m = FM_MK((hi & 0x000fffff) | 0x3ff00000, FM_LO(v));
  double u = fma(m, rc[j], -1.0);
  double q = fm_log_poly(u);
  double ed = FM_MK(0x43300000, (int)(0x80000000u ^ (unsigned)e)) - 4503601774854144.0;  // (double)e without I2F
  return fma(ed, 0.6931471805599453, fma(u, q, lc[j]));
}

// the same with the two table entries of an interval side by side ({rc[j], lc[j]}): one 16-byte load per logarithm
struct FmRL { double rc, lc; };
__host__ __device__ __forceinline__ double flog_nc2(double v, const FmRL* __restrict__ rl) {
  int hi = FM_HI(v);
  int e = (hi >> 20) - 1023;
  const FmRL t = rl[(hi >> 13) & (FM_NT - 1)];
  double m = FM_MK((hi & 0x000fffff) | 0x3ff00000, FM_LO(v));
  double u = fma(m, t.rc, -1.0);
  double q = fm_log_poly(u);
  double ed = FM_MK(0x43300000, (int)(0x80000000u ^ (unsigned)e)) - 4503601774854144.0;  // (double)e without I2F
  return fma(ed, 0.6931471805599453, fma(u, q, t.lc));
}

// 1/x for normal positive x (callers guard the range)
__host__ __device__ __forceinline__ double frcp(double x) {
#if defined(__CUDA_ARCH__)
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  double e = fma(-x, y0, 1.0);
  double t = fma(e, e, e);
  double y1 = fma(y0, t, y0);   // error e^3 ~ 2^-60
  e = fma(-x, y1, 1.0);         // one linear polish: rounding-level result even if the seed has < 20 good bits
  return fma(y1, e, y1);
#else
  return 1.0 / x;
#endif
}

// 1/x to ~2^-58 (seed + the cubic step, no polish): for sums that only steer an iteration (Newton score / information)
__host__ __device__ __forceinline__ double frcp_nr(double x) {
#if defined(__CUDA_ARCH__)
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  double e = fma(-x, y0, 1.0);
  double t = fma(e, e, e);
  return fma(y0, t, y0);
#else
  return 1.0 / x;
#endif
}

// lgamma(x) for x >= 128 from its logarithm lx = log(x): Stirling series, truncation 1/(1680 x^7) < 2e-18
__host__ __device__ __forceinline__ double lgamma_big(double x, double lx) {
  double ix = frcp(x), ix2 = ix * ix;
  double ser = fma(ix2, fma(ix2, 1.0 / 1260.0, -1.0 / 360.0), 1.0 / 12.0) * ix;
  return fma(x - 0.5, lx, -x) + (0.91893853320467274178 + ser);
}
