// zinb.cuh -- H4: the zero-inflation probability pi0_g (reference R/ruvIIInb.R:237-251,437-451,770-772):
//   pi0_g = logistic( optim(p = -2, fn = logl.zinb, y, mu, size)$par ),
//   logl.zinb(p) = - sum_{c in zeroinf} log( omega 1[y = 0] + (1 - omega) dnbinom(y; mu, size) ),  omega = logistic(p),
// optim's default method being Nelder-Mead (R src/appl/optim.c:nmmin; alpha 1, beta .5, gamma 2, reltol
// sqrt(DBL_EPSILON), maxit 500).  The objective depends on the data only through q_c = dnbinom(0; mu_c, size) of the
// zero counts, the number n_pos of positive counts and S_pos = sum log dnbinom(y > 0):
//   f(p) = - [ sum_{y = 0} log(omega + (1 - omega) q_c) + n_pos log(1 - omega) + S_pos ].
// k_zinb_q writes q_c once (one sweep of the counts); every Nelder-Mead evaluation is then a pass over q only, all genes
// in lock-step (k_nm_eval), with the simplex logic of each gene as a small state machine (k_nm_step).
// The E-step weights (H5) are never stored: the sweeps recompute them (GeneRegs::wz).  Included by irls.cu.
#pragma once
#include "kernels.cuh"

#define ZQ_SKIP 2.0   // sentinel in q: not a zero count of a zero-inflated cell

template <int K>
struct OpZinbQ {
  static constexpr bool PADS = true;
  static constexpr bool QUAD = false;
  static constexpr bool USES_PSI = true;
  __device__ __forceinline__ void psi_changed() { rlogr = g.r * log(g.r); }
  GeneState<K> st; GeneRegs<K> g;
  double* Q; long long Np; double* qrow;
  double spos, rlogr; int npos, row_;
  double* S_pos; int* n_pos;
  __device__ __forceinline__ void row_begin(int row, int) {
    g.load(st, row); qrow = Q + (long long)row * Np; spos = 0.0; npos = 0; rlogr = g.r * log(g.r); row_ = row;
  }
  __device__ __forceinline__ void group_begin(int j) { g.mbj = g.mb[j]; }
  __device__ __forceinline__ double one(int y, double lmu, bool m, long long pos) {
    if (!m || !g.zinf[pos]) return ZQ_SKIP;
    double mu = g.expm(lmu);
    double L = g.logL(mu);
    if (y == 0) return fexp(fma(-g.r, L, rlogr), g.mt->ex);   // dnbinom(0; mu, r) = (r/(r + mu))^r
    spos += nb_logpmf_t(y, (double)y, lmu, L, g.r, g.lgr, rlogr, g.tl, g.mt);
    npos++;
    return ZQ_SKIP;
  }
  template <bool FULLCHUNK>
  __device__ __forceinline__ void pair(int y0, int y1, int c, int ti, const double* tile, bool m0, bool m1) {
    double l0, l1;
    g.lmu_pair(tile, ti, l0, l1);
    const long long pos = (long long)c * CH + (ti & (CH - 1));
    double2 q;
    q.x = one(y0, l0, m0, pos);
    q.y = one(y1, l1, m1, pos + 1);
    *reinterpret_cast<double2*>(qrow + pos) = q;
  }
  __device__ __forceinline__ void group_end(int) {}
  __device__ __forceinline__ void row_end(int row) {
    double s = warp_sum(spos);
    int n = __reduce_add_sync(FULL, npos);
    if ((threadIdx.x & 31) == 0) { S_pos[row] = s; n_pos[row] = n; }
  }
};
template <int K>
__global__ void __launch_bounds__(ROW_THREADS, 2) k_zinb_q(RowGeom geo, GeneState<K> st, double* Q, double* S_pos, int* n_pos) {
  extern __shared__ double smem[];
  __shared__ RowShared sh;
  if (threadIdx.x == 0) sh.wsets[0] = st.W;
  OpZinbQ<K> op;
  op.st = st; op.Q = Q; op.Np = geo.Np; op.S_pos = S_pos; op.n_pos = n_pos;
  stream_rows<K, 1>(geo, &sh, op, smem);
}

// ---- Nelder-Mead, one dimension, one state machine per gene ---------------------------------------------------------------
enum { NM_INIT = 0, NM_CALCVERT, NM_REFLECT, NM_EXPAND, NM_CONTRACT, NM_DONE };
struct NMState {
  double P[2], F[2];     // the two vertices and their function values
  double C, Cpt;         // centroid; the reflected point kept during an expansion
  double B;              // point being evaluated
  double VL, VH, VR, convtol, oldsize;
  int L, H;              // 0-based indices of the lowest / highest vertex
  int phase, funcount, fail, err;
};
#define NM_BIG 1.0e35
#define NM_RELTOL 1.4901161193847656e-08   // sqrt(.Machine$double.eps)
#define NM_MAXIT 500

static __global__ void k_nm_init(int Gl, NMState* S, double* xeval) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  NMState s;
  memset(&s, 0, sizeof s);
  s.phase = NM_INIT; s.B = -2.0;
  S[g] = s;
  xeval[g] = -2.0;
}

// f(x) of every gene still running: one warp per gene over its q row
static __global__ void __launch_bounds__(256) k_nm_eval(int Gl, long long Np, const double* Q, const double* S_pos, const int* n_pos,
                                                const NMState* S, const double* xeval, const MathTabs* gt, double* feval) {
  __shared__ MathTabs mt;
  {
    const double* src = reinterpret_cast<const double*>(gt);
    double* dst = reinterpret_cast<double*>(&mt);
    for (int i = threadIdx.x; i < (int)(sizeof(MathTabs) / 8); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= Gl || S[g].phase == NM_DONE) return;
  const double om = 1.0 / (1.0 + exp(-xeval[g])), om1 = 1.0 - om;
  const double2* q2 = reinterpret_cast<const double2*>(Q + (long long)g * Np);
  double acc = 0.0;
  for (long long i = lane; i < Np / 2; i += 32) {
    double2 q = q2[i];
    if (q.x <= 1.0) acc += flog(fma(om1, q.x, om), mt.rc, mt.lc);
    if (q.y <= 1.0) acc += flog(fma(om1, q.y, om), mt.rc, mt.lc);
  }
  acc = warp_sum(acc);
  if (lane == 0) feval[g] = -(acc + (double)n_pos[g] * log(om1) + S_pos[g]);
}

__device__ __forceinline__ bool nm_finite(double f) { return fabs(f) <= 1.79769313486231570e308; }
// top of nmmin's do-loop for n = 1 after the vertices are valued: pick L/H, test convergence, reflect
__device__ __forceinline__ void nm_loop_top(NMState& s) {
  // for (j = 1..n1) if (j != L) {...}: with two vertices only the other one is compared
  s.VL = s.F[s.L]; s.VH = s.VL; s.H = s.L;
  const int j = 1 - s.L;
  const double f = s.F[j];
  if (f < s.VL) { s.L = j; s.VL = f; }
  if (f > s.VH) { s.H = j; s.VH = f; }
  if (s.VH <= s.VL + s.convtol) { s.phase = NM_DONE; return; }   // abstol = -Inf never triggers
  double temp = -s.P[s.H];
  temp += s.P[0]; temp += s.P[1];
  s.C = temp / 1.0;
  s.B = (1.0 + 1.0) * s.C - 1.0 * s.P[s.H];
  s.phase = NM_REFLECT;
}
__device__ __forceinline__ void nm_while_check(NMState& s, bool calcvert) {
  if (!(s.funcount <= NM_MAXIT)) { s.fail = 1; s.phase = NM_DONE; return; }
  if (calcvert) { s.B = s.P[1 - s.L]; s.phase = NM_CALCVERT; return; }
  nm_loop_top(s);
}
static __global__ void k_nm_step(int Gl, NMState* S, const double* feval, double* xeval, int* n_active) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  NMState s = S[g];
  if (s.phase == NM_DONE) return;
  double f = feval[g];
  switch (s.phase) {
    case NM_INIT:
      if (!nm_finite(f)) { s.err = 1; s.phase = NM_DONE; break; }   // optim(): error -> tryCatch -> NA
      s.funcount = 1;
      s.convtol = NM_RELTOL * (fabs(f) + NM_RELTOL);
      s.P[0] = -2.0; s.F[0] = f; s.L = 0;
      {
        double step = 0.1 * 2.0;        // 0.1 |par|
        double trystep = step;
        s.P[1] = -2.0;
        while (s.P[1] == -2.0) { s.P[1] = -2.0 + trystep; trystep *= 10.0; }
        s.oldsize = trystep;            // size += trystep after the *= 10
      }
      s.B = s.P[1]; s.phase = NM_CALCVERT;
      break;
    case NM_CALCVERT:
      if (!nm_finite(f)) f = NM_BIG;
      s.funcount++;
      s.F[1 - s.L] = f;
      nm_loop_top(s);
      break;
    case NM_REFLECT:
      if (!nm_finite(f)) f = NM_BIG;
      s.funcount++;
      s.VR = f;
      if (s.VR < s.VL) {
        double t = 2.0 * s.B + (1.0 - 2.0) * s.C;
        s.Cpt = s.B; s.B = t; s.phase = NM_EXPAND;
      } else {
        if (s.VR < s.VH) { s.P[s.H] = s.B; s.F[s.H] = s.VR; }
        s.B = (1.0 - 0.5) * s.P[s.H] + 0.5 * s.C;
        s.phase = NM_CONTRACT;
      }
      break;
    case NM_EXPAND:
      if (!nm_finite(f)) f = NM_BIG;
      s.funcount++;
      if (f < s.VR) { s.P[s.H] = s.B; s.F[s.H] = f; }
      else { s.P[s.H] = s.Cpt; s.F[s.H] = s.VR; }
      nm_while_check(s, false);
      break;
    case NM_CONTRACT: {
      if (!nm_finite(f)) f = NM_BIG;
      s.funcount++;
      bool calcvert = false;
      if (f < s.F[s.H]) { s.P[s.H] = s.B; s.F[s.H] = f; }
      else if (s.VR >= s.VH) {
        calcvert = true;
        const int j = 1 - s.L;
        s.P[j] = 0.5 * (s.P[j] - s.P[s.L]) + s.P[s.L];
        double size = fabs(s.P[j] - s.P[s.L]);
        if (size < s.oldsize) s.oldsize = size;
        else { s.fail = 10; s.phase = NM_DONE; break; }
      }
      nm_while_check(s, calcvert);
      break;
    }
    default: break;
  }
  if (s.phase != NM_DONE) { xeval[g] = s.B; atomicAdd(n_active, 1); }
  S[g] = s;
}
// pi0 <- logistic(par) where optim succeeded, else the previous value (:251,451); counts non-finite entries (:453)
static __global__ void k_nm_finish(int Gl, const NMState* S, const double* pi0_old, double* pi0_new, int* bad) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Gl) return;
  const NMState& s = S[g];
  double v = pi0_old[g];
  if (!s.err) {
    double t = 1.0 / (1.0 + exp(-s.P[s.L]));
    if (t == t) v = t;
  }
  pi0_new[g] = v;
  if (!(fabs(v) <= 1.79769313486231570e308)) atomicAdd(bad, 1);
}
