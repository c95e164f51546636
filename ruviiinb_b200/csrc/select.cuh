// select.cuh -- exact order statistics (median, MAD) by 8-bit radix select on order-preserving
// 64-bit keys.  R semantics restated: stats::median (mean of the two middle order statistics for
// even n, NA if any NaN) and stats::mad = 1.4826 * median(|x - median(x)|), as used at
// R/ruvIIInb.R:371-372,409,428-429,713,723,734.
#pragma once
#include "common.cuh"

#define MAD_CONST 1.4826

template <int NT>
__device__ __forceinline__ void group_sync() {
  if (NT == 32) __syncwarp(); else __syncthreads();
}

// Scratch a group needs in shared memory.
struct SelScratch {
  unsigned int hist[256];
  unsigned long long s0, s1;
  unsigned long long red[32];
};

template <int NT>
__device__ __forceinline__ unsigned long long group_max_ull(unsigned long long v, SelScratch* sc, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(FULL, v, o);
    v = t > v ? t : v;
  }
  if (NT == 32) return v;
  group_sync<NT>();
  if ((tid & 31) == 0) sc->red[tid >> 5] = v;
  group_sync<NT>();
  unsigned long long r = 0;
  for (int w = 0; w < NT / 32; ++w) r = sc->red[w] > r ? sc->red[w] : r;
  group_sync<NT>();
  return r;
}
template <int NT>
__device__ __forceinline__ unsigned long long group_sum_ull(unsigned long long v, SelScratch* sc, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  if (NT == 32) return v;
  group_sync<NT>();
  if ((tid & 31) == 0) sc->red[tid >> 5] = v;
  group_sync<NT>();
  unsigned long long r = 0;
  for (int w = 0; w < NT / 32; ++w) r += sc->red[w];
  group_sync<NT>();
  return r;
}

// kth (0-based) smallest key among the valid items i in [0,len); get(i, key) returns false for
// items to skip.  All NT threads of the group call it with the same arguments.
template <int NT, class Get>
__device__ unsigned long long group_select_kth(int len, long long kth, Get get, SelScratch* sc, int tid) {
  unsigned long long prefix = 0, pmask = 0;
  for (int pass = 7; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int i = tid; i < 256; i += NT) sc->hist[i] = 0;
    group_sync<NT>();
    for (int i0 = 0; i0 < len; i0 += NT) {
      int i = i0 + tid;
      unsigned long long key = 0;
      bool ok = (i < len) && get(i, key) && ((key & pmask) == prefix);
      int b = ok ? (int)((key >> shift) & 255ull) : 256;  // one shared dummy for the lanes that sit out (match cost grows with distinct values)
      // warp-aggregated histogram update: one atomic per distinct bucket in the warp
      unsigned peers = __match_any_sync(FULL, b);
      if (ok && (__ffs(peers) - 1) == (tid & 31)) atomicAdd(&sc->hist[b], (unsigned)__popc(peers));
    }
    group_sync<NT>();
    if (tid == 0) {
      long long acc = 0;
      int b = 0;
      for (; b < 255; ++b) {
        if (acc + (long long)sc->hist[b] > kth) break;
        acc += sc->hist[b];
      }
      sc->s0 = (unsigned long long)b;
      sc->s1 = (unsigned long long)acc;
    }
    group_sync<NT>();
    prefix |= sc->s0 << shift;
    pmask |= 0xffull << shift;
    kth -= (long long)sc->s1;
    group_sync<NT>();
  }
  return prefix;
}

// R median of the valid items (n = number of valid items; caller guarantees no NaN among them).
template <int NT, class Get>
__device__ double group_median(int len, long long n, Get get, SelScratch* sc, int tid) {
  if (n <= 0) return __longlong_as_double(0x7ff8000000000000ll);
  unsigned long long up = group_select_kth<NT>(len, n / 2, get, sc, tid);
  double hi = key_f64(up);
  if (n & 1) return hi;
  // lower middle: the largest key below `up` if exactly n/2 keys are below it, else `up` itself
  unsigned long long cl = 0, ml = 0;
  for (int i = tid; i < len; i += NT) {
    unsigned long long key;
    if (get(i, key) && key < up) { cl++; ml = key > ml ? key : ml; }
  }
  cl = group_sum_ull<NT>(cl, sc, tid);
  ml = group_max_ull<NT>(ml, sc, tid);
  double lo = (cl == (unsigned long long)(n / 2)) ? key_f64(ml) : hi;
  return (lo + hi) / 2.0;  // R: mean(c(lo, hi))
}

// median and mad (already x1.4826) of items val(i), i in [0,len) with valid(i); NaN if any NaN.
template <int NT, class Val, class Valid>
__device__ void group_median_mad(int len, Val val, Valid valid, SelScratch* sc, int tid, double* med_out,
                                 double* mad_out) {
  unsigned long long n = 0, nnan = 0;
  for (int i = tid; i < len; i += NT) {
    if (valid(i)) { n++; double v = val(i); if (v != v) nnan++; }
  }
  n = group_sum_ull<NT>(n, sc, tid);
  nnan = group_sum_ull<NT>(nnan, sc, tid);
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  if (nnan > 0 || n == 0) { *med_out = qnan; *mad_out = qnan; return; }
  auto get1 = [&](int i, unsigned long long& key) { if (!valid(i)) return false; key = f64_key(val(i)); return true; };
  double med = group_median<NT>(len, (long long)n, get1, sc, tid);
  auto get2 = [&](int i, unsigned long long& key) { if (!valid(i)) return false; key = f64_key(fabs(val(i) - med)); return true; };
  double mad = group_median<NT>(len, (long long)n, get2, sc, tid);
  *med_out = med;
  *mad_out = MAD_CONST * mad;
}

// ------------------------------------------------------------------------------------------------
// CTA-wide exact selection over a long global array (the exact-MAD path of a gene row, N ~ 1e5):
// 12-bit radix descent that stops as soon as the bucket holding the wanted rank has <= RS_CAP keys,
// gathers that bucket into shared memory and ranks it there.  Typical cost: 2 histogram passes +
// 1 gather pass over the row (the previous 8-bit select took 9).  Returns BOTH middle order
// statistics (ranks (n-1)/2 and n/2) so that R's even-n median needs no extra pass.
// ------------------------------------------------------------------------------------------------
#define RS_NT 512
#define RS_BINS 4096
#define RS_CAP 1024
struct RowSelScratch {
  unsigned hist[RS_BINS];
  unsigned long long buf[RS_CAP];
  unsigned wtot[RS_NT / 32];
  unsigned long long red[RS_NT / 32];
  unsigned bucket, below, cnt, nbuf;
  unsigned long long out_hi, out_lo;
};

// key(i) for i in [0, len), len a multiple of RS_NT; n = number of real items (the rest are +Inf pads
// that sort last); ranks are 0-based.  All RS_NT threads call with identical arguments.
// CAP: bucket size at which the descent stops and the bucket is ranked in shared memory (cost ~ CAP^2 / RS_NT compare steps).
// Rows streamed from HBM (a pass costs a sweep over the row) stop early, at RS_CAP; vectors that already sit in shared memory
// (get.res caps: a pass is ~10 iterations per thread) keep descending to a few dozen keys.  AGG: warp-aggregated histogram
// updates (one atomic per distinct bin and warp) -- pays at the first level, where most keys share a few exponent bins.
template <int CAP = RS_CAP, class KeyFn>
__device__ void cta_two_middles(int len, long long n, KeyFn key, RowSelScratch* sc, unsigned long long* klo_out,
                                unsigned long long* khi_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long prefix = 0, pmask = 0;
  long long rank = n / 2;
  unsigned cnt = 0;
  int shift = 64;
  for (int level = 0; level < 6; ++level) {
    const int bits = (level < 5) ? 12 : 4;
    shift -= bits;
    const int nb = 1 << bits;
    for (int i = tid; i < nb; i += RS_NT) sc->hist[i] = 0;
    __syncthreads();
    // four independent loads in flight per thread (the pass is latency-bound: load -> match -> atomic)
    for (int i0 = 0; i0 < len; i0 += 4 * RS_NT) {
      unsigned long long kx[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) kx[q] = (i0 + q * RS_NT < len) ? key(i0 + q * RS_NT + tid) : ~pmask;  // ~pmask never matches a non-empty prefix mask
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        bool ok = (i0 + q * RS_NT < len) && (kx[q] & pmask) == prefix;
        int b = ok ? (int)((kx[q] >> shift) & (unsigned long long)(nb - 1)) : nb;   // one shared dummy for the lanes that sit out
        if (level == 0) {
          unsigned peers = __match_any_sync(FULL, b);
          if (ok && (__ffs(peers) - 1) == lane) atomicAdd(&sc->hist[b], (unsigned)__popc(peers));
        } else if (ok) {
          atomicAdd(&sc->hist[b], 1u);   // deeper levels spread the surviving keys over the bins: no contention to aggregate
        }
      }
    }
    __syncthreads();
    // locate the bucket holding `rank`: thread t owns bins [t*per, (t+1)*per)
    const int per = (nb + RS_NT - 1) / RS_NT;  // 8 (or 1 for the 16-bin level; threads >= nb own nothing)
    unsigned loc = 0;
    for (int q = 0; q < per; ++q) { int bi = tid * per + q; if (bi < nb) loc += sc->hist[bi]; }
    unsigned inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sc->wtot[warp] = inc;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < warp; ++w) base += sc->wtot[w];
    unsigned excl = base + inc - loc;
    if ((long long)excl <= rank && rank < (long long)excl + loc) {
      unsigned acc = excl;
      for (int q = 0; q < per; ++q) {
        int bi = tid * per + q;
        unsigned h = (bi < nb) ? sc->hist[bi] : 0;
        if (rank < (long long)acc + h) { sc->bucket = (unsigned)bi; sc->below = acc; sc->cnt = h; break; }
        acc += h;
      }
    }
    __syncthreads();
    prefix |= (unsigned long long)sc->bucket << shift;
    pmask |= (unsigned long long)(nb - 1) << shift;
    rank -= (long long)sc->below;
    cnt = sc->cnt;
    __syncthreads();
    if (cnt <= CAP) break;
  }
  // gather the bucket; largest key below it
  if (tid == 0) sc->nbuf = 0;
  __syncthreads();
  unsigned long long bmax = 0;
  const bool fits = cnt <= CAP;
  for (int i0 = 0; i0 < len; i0 += RS_NT) {
    unsigned long long kx = key(i0 + tid);
    unsigned long long mk = kx & pmask;
    if (mk == prefix) {
      if (fits) { unsigned p = atomicAdd(&sc->nbuf, 1u); sc->buf[p] = kx; }
    } else if (mk < prefix) {
      bmax = kx > bmax ? kx : bmax;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(FULL, bmax, o); bmax = t > bmax ? t : bmax; }
  if (lane == 0) sc->red[warp] = bmax;
  __syncthreads();
  bmax = 0;
  for (int w = 0; w < RS_NT / 32; ++w) bmax = sc->red[w] > bmax ? sc->red[w] : bmax;
  if (!fits) {
    // > CAP identical keys: all 64 bits are resolved and the bucket is one value
    *khi_out = prefix;
    *klo_out = rank >= 1 ? prefix : bmax;
    __syncthreads();
    return;
  }
  if (tid == 0) { sc->out_lo = bmax; sc->out_hi = 0; }
  __syncthreads();
  for (unsigned i = tid; i < cnt; i += RS_NT) {
    unsigned long long x = sc->buf[i];
    unsigned lt = 0, eq = 0;
    for (unsigned j = 0; j < cnt; ++j) { unsigned long long yk = sc->buf[j]; lt += yk < x; eq += yk == x; }
    if ((long long)lt <= rank && rank < (long long)lt + eq) sc->out_hi = x;
    if (rank >= 1 && (long long)lt <= rank - 1 && rank - 1 < (long long)lt + eq) sc->out_lo = x;
  }
  __syncthreads();
  *khi_out = sc->out_hi;
  *klo_out = sc->out_lo;
  __syncthreads();
}
