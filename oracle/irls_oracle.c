/*
 * irls_oracle.c -- C/OpenMP restatement of ONE inner IRLS iteration of ruvIII.nb (NB, nbatch == 1).
 *
 * ORACLE / CPU BASELINE ONLY (test infrastructure): built into oracle/_ref/libirls_oracle.so by
 * oracle/Makefile, loaded by oracle/cport.py, used by tests/ (cross-checked against the NumPy oracle)
 * and by bench.py's cpu_baseline / --impl reference legs.  Nothing under ruviiinb_b200/ links it.
 * PARITY UNPINNED: the reference has no tests or golden vectors and R is not installed here.
 *
 * Follows the reference line by line, with the same dense G x N intermediates it builds
 * (lmu.hat, sig.inv, signdev, Z: R/ruvIIInb.R:303-321) and one task per gene / per cell where the
 * reference forks through BiocParallel::bpmapply (R/ruvIIInb.R:350,356,385):
 *   working quantities            R/ruvIIInb.R:303-321
 *   projection / RM_Z, RM_W       R/ruvIIInb.R:328-340, 639-665
 *   get.coef2 + alpha.mean        R/ruvIIInb.R:346-354, 712-720
 *   get.adev + clamp              R/ruvIIInb.R:356-376, 722-730
 *   getW + normalise              R/ruvIIInb.R:380-402, 732-744
 *   gmean                         R/ruvIIInb.R:405-415
 *   Mb (rowWtAvg) + clamp         R/ruvIIInb.R:418-433, 667-675
 *   candidate log-likelihood      R/ruvIIInb.R:469-484
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG_DBL_MIN (-708.3964185322641) /* log(.Machine$double.xmin), R/ruvIIInb.R:290 */
#define MAXK 8

/* ---- stats::median / stats::mad (quickselect on a scratch copy) -------------------------------- */
static double kth_smallest(double* a, long n, long k) {
  long lo = 0, hi = n - 1;
  while (lo < hi) {
    double x = a[(lo + hi) / 2];
    long i = lo, j = hi;
    do {
      while (a[i] < x) i++;
      while (x < a[j]) j--;
      if (i <= j) { double t = a[i]; a[i] = a[j]; a[j] = t; i++; j--; }
    } while (i <= j);
    if (j < k) lo = i;
    if (k < i) hi = j;
  }
  return a[k];
}
static double r_median_scratch(double* a, long n) { /* destroys a; NaN if any NaN */
  if (n <= 0) return NAN;
  for (long i = 0; i < n; ++i) if (a[i] != a[i]) return NAN;
  double hi = kth_smallest(a, n, n / 2);
  if (n & 1) return hi;
  double lo = a[0]; /* after selection everything left of n/2 is <= hi: the lower middle is their max */
  for (long i = 1; i < n / 2; ++i) if (a[i] > lo) lo = a[i];
  return (lo + hi) / 2.0;
}
static double r_mad(const double* x, long n, long stride, double* scratch) {
  for (long i = 0; i < n; ++i) scratch[i] = x[i * stride];
  double med = r_median_scratch(scratch, n);
  if (med != med) return NAN;
  for (long i = 0; i < n; ++i) scratch[i] = fabs(x[i * stride] - med);
  return 1.4826 * r_median_scratch(scratch, n);
}
static double r_median(const double* x, long n, long stride, double* scratch) {
  for (long i = 0; i < n; ++i) scratch[i] = x[i * stride];
  return r_median_scratch(scratch, n);
}
/* MASS::psi.huber(u, k) = pmin(1, k/|u|), NaN kept */
static inline double psi_huber(double u, double k) {
  double q = k / fabs(u);
  return (q < 1.0) ? q : ((q != q) ? q : 1.0);
}
/* solve(M, rhs) by Gaussian elimination with partial pivoting (what R's solve() does); 0 on success */
static int solve_small(int n, double M[MAXK + 1][MAXK + 1], double* rhs) {
  for (int c = 0; c < n; ++c) {
    int p = c; double best = fabs(M[c][c]);
    for (int r = c + 1; r < n; ++r) if (fabs(M[r][c]) > best) { best = fabs(M[r][c]); p = r; }
    if (!(best > 0.0)) return 1;
    if (p != c) { for (int j = 0; j < n; ++j) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; } double t = rhs[c]; rhs[c] = rhs[p]; rhs[p] = t; }
    for (int r = c + 1; r < n; ++r) {
      double f = M[r][c] / M[c][c];
      for (int j = c; j < n; ++j) M[r][j] -= f * M[c][j];
      rhs[r] -= f * rhs[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = rhs[r];
    for (int j = r + 1; j < n; ++j) s -= M[r][j] * rhs[j];
    rhs[r] = s / M[r][r];
  }
  return 0;
}
/* log dnbinom(y; mu, size = r) through lgamma; +-Inf/NaN -> log(DBL_MIN) (R/ruvIIInb.R:286,290) */
static inline double nb_logpmf(double y, double lmu, double mu, double r) {
  int sg;
  double L = log(mu + r), v;
  if (y == 0.0) v = r * (log(r) - L);
  else v = lgamma_r(y + r, &sg) - lgamma_r(r, &sg) - lgamma_r(y + 1.0, &sg) + r * (log(r) - L) + y * (lmu - L);
  if (!(fabs(v) <= 1.79769313486231570e308)) v = LOG_DBL_MIN;
  return v;
}
/* sign(y-mu) sqrt(2 (dnbinom(y; mu=y) - dnbinom(y; mu))) : the lgamma terms cancel (R/ruvIIInb.R:317) */
static inline double nb_signdev(double y, double lmu, double mu, double r) {
  double t = -(y + r) * (log(y + r) - log(mu + r));
  if (y > 0.0) t += y * (log(y) - lmu);
  double d2 = 2.0 * t, d = sqrt(d2 > 0.0 ? d2 : 0.0);
  if (d2 != d2) d = d2;
  return (y > mu) ? d : ((y < mu) ? -d : (mu != mu ? mu : 0.0));
}

/*
 * One inner iteration, in place.  Row-major arrays: Y [G][N], Mb [G][m], alpha/adev [G][k], W [N][k].
 * loglik_out [G] receives the candidate per-gene log-likelihood; bad[3] the problematic-entry counts.
 * Returns 0, or 1 on allocation failure.
 */
int irls_oracle_iteration(const int32_t* Y, long G, long N, int k, int m, const int32_t* grp, const uint8_t* ctl,
                          double* gmean, double* Mb, double* alpha, double* adev, double* W, const double* psi,
                          const double* step, double k_huber, double lambda_a, double lambda_b, double* loglik_out,
                          int* bad, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const int nth = omp_get_max_threads();
#else
  const int nth = 1;
#endif
  const long GN = G * N;
  double* Z = (double*)malloc(sizeof(double) * GN);
  double* SI = (double*)malloc(sizeof(double) * GN);   /* sig.inv */
  double* SD = (double*)malloc(sizeof(double) * GN);   /* signdev */
  double* HW = (double*)malloc(sizeof(double) * GN);   /* wt: Huber weights stored by the gmean loop (:410) */
  double* RMW = (double*)malloc(sizeof(double) * N * k);
  double* Wbar = (double*)calloc((size_t)m * k, sizeof(double));
  long* gn = (long*)calloc(m, sizeof(long));
  double* Ab = (double*)malloc(sizeof(double) * G * (k * k + k));
  double* Ag = (double*)malloc(sizeof(double) * G * k * k);
  double* cg = (double*)malloc(sizeof(double) * G * k);
  double* scratch = (double*)malloc(sizeof(double) * (size_t)nth * (N > G ? N : G));
  double* alpha_new = (double*)malloc(sizeof(double) * G * k);
  double* adev_new = (double*)malloc(sizeof(double) * G * k);
  double* W_new = (double*)malloc(sizeof(double) * N * k);
  double* sigma = (double*)malloc(sizeof(double) * G);
  double* zbar = (double*)malloc(sizeof(double) * (size_t)nth * m);
  if (!Z || !SI || !SD || !HW || !RMW || !Wbar || !gn || !Ab || !Ag || !cg || !scratch || !alpha_new || !adev_new || !W_new || !sigma || !zbar) return 1;
  const long sstride = (N > G ? N : G);
  long nctl = 0;
  for (long g = 0; g < G; ++g) nctl += ctl[g] ? 1 : 0;
  long* cidx = (long*)malloc(sizeof(long) * (nctl > 0 ? nctl : 1));
  { long q = 0; for (long g = 0; g < G; ++g) if (ctl[g]) cidx[q++] = g; }

  /* :303-321 working quantities */
#pragma omp parallel for schedule(static)
  for (long g = 0; g < G; ++g) {
    const double r = 1.0 / psi[g];
    for (long c = 0; c < N; ++c) {
      double lmu = gmean[g] + Mb[g * m + grp[c]];
      for (int l = 0; l < k; ++l) lmu += alpha[g * k + l] * W[c * k + l];
      double mu = exp(lmu), y = (double)Y[g * N + c];
      SI[g * N + c] = 1.0 / (exp(-lmu) + psi[g]);
      SD[g * N + c] = nb_signdev(y, lmu, mu, r);
      Z[g * N + c] = lmu + step[g] * ((y + 0.01) / (mu + 0.01) - 1.0);
    }
  }
  /* :331-340 RM_W = W - groupmean(W) */
  for (long c = 0; c < N; ++c) { gn[grp[c]]++; for (int l = 0; l < k; ++l) Wbar[grp[c] * k + l] += W[c * k + l]; }
  for (int j = 0; j < m; ++j) for (int l = 0; l < k; ++l) Wbar[j * k + l] /= (double)gn[j];
  for (long c = 0; c < N; ++c) for (int l = 0; l < k; ++l) RMW[c * k + l] = W[c * k + l] - Wbar[grp[c] * k + l];

  /* :346-350 get.coef2 per gene: sigma, weights, A_g, c_g = RM_W' diag(w) RM_Z ; b_g = c_g - A_g adev_old */
#pragma omp parallel for schedule(dynamic, 4)
  for (long g = 0; g < G; ++g) {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    double* sc = scratch + (long)t * sstride;
    double* zb = zbar + (long)t * m;
    const double sig = r_mad(SD + g * N, N, 1, sc) / 0.6745;  /* :713 */
    sigma[g] = sig;
    for (int j = 0; j < m; ++j) zb[j] = 0.0;
    double sw = 0.0;
    for (long c = 0; c < N; ++c) {
      zb[grp[c]] += Z[g * N + c];
      double w = psi_huber(SD[g * N + c], k_huber * sig) * SI[g * N + c];  /* :714 (wt.zinb == 1) */
      sc[c] = w; sw += w;
    }
    for (int j = 0; j < m; ++j) zb[j] /= (double)gn[j];  /* rowAvg :657-665 */
    const double mean_w = sw / (double)N;                 /* :715 */
    double A[MAXK][MAXK], cv[MAXK];
    for (int a = 0; a < k; ++a) { cv[a] = 0.0; for (int b = 0; b < k; ++b) A[a][b] = 0.0; }
    for (long c = 0; c < N; ++c) {
      const double w = sc[c] / mean_w, rz = Z[g * N + c] - zb[grp[c]];
      for (int a = 0; a < k; ++a) {
        const double wa = w * RMW[c * k + a];
        cv[a] += wa * rz;
        for (int b = 0; b < k; ++b) A[a][b] += wa * RMW[c * k + b];
      }
    }
    for (int a = 0; a < k; ++a) {
      cg[g * k + a] = cv[a];
      double s = cv[a];
      for (int b = 0; b < k; ++b) { Ag[(g * k + a) * k + b] = A[a][b]; s -= A[a][b] * adev[g * k + b]; Ab[g * (k * k + k) + a * k + b] = A[a][b]; }
      Ab[g * (k * k + k) + k * k + a] = s;  /* :718 */
    }
  }
  /* :351-354 alpha.mean = solve(sum A, sum b) */
  double amean[MAXK];
  {
    double M[MAXK + 1][MAXK + 1], rhs[MAXK + 1];
    for (int a = 0; a < k; ++a) { rhs[a] = 0.0; for (int b = 0; b < k; ++b) M[a][b] = 0.0; }
    for (long g = 0; g < G; ++g) for (int a = 0; a < k; ++a) {
      rhs[a] += Ab[g * (k * k + k) + k * k + a];
      for (int b = 0; b < k; ++b) M[a][b] += Ab[g * (k * k + k) + a * k + b];
    }
    int fail = solve_small(k, M, rhs);
    for (int a = 0; a < k; ++a) amean[a] = fail ? NAN : rhs[a];
  }
  /* :356-365 get.adev */
  int nbad_alpha = 0;
#pragma omp parallel for schedule(static) reduction(+ : nbad_alpha)
  for (long g = 0; g < G; ++g) {
    double M[MAXK + 1][MAXK + 1], rhs[MAXK + 1];
    for (int a = 0; a < k; ++a) {
      double s = cg[g * k + a];
      for (int b = 0; b < k; ++b) { M[a][b] = Ag[(g * k + a) * k + b]; s -= M[a][b] * amean[b]; }
      rhs[a] = s;
    }
    for (int a = 0; a < k; ++a) M[a][a] += lambda_a;
    int fail = solve_small(k, M, rhs);
    for (int a = 0; a < k; ++a) {
      double x = fail ? NAN : rhs[a];
      adev_new[g * k + a] = x;
      alpha_new[g * k + a] = x + amean[a];
      if (!(fabs(alpha_new[g * k + a]) <= 1.79769313486231570e308)) nbad_alpha++;
    }
  }
  if (nbad_alpha) memcpy(alpha_new, alpha, sizeof(double) * G * k);  /* :368 */
  /* :371-376 clamp columns */
  for (int l = 0; l < k; ++l) {
    double med = r_median(alpha_new + l, G, k, scratch), mad = r_mad(alpha_new + l, G, k, scratch);
    double hi = med + 4.0 * mad, lo = med - 4.0 * mad;
    for (long g = 0; g < G; ++g) { double v = alpha_new[g * k + l]; if (v > hi) v = hi; if (v < lo) v = lo; alpha_new[g * k + l] = v; }
  }
  /* :380-390 getW per cell over control genes: sequential one-regressor WLS */
#pragma omp parallel for schedule(static)
  for (long c = 0; c < N; ++c) {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    double* sc = scratch + (long)t * sstride;  /* first nctl: |.| scratch; we need 2*nctl <= sstride */
    double dv[4096], wv[4096];
    double* d = nctl <= 4096 ? dv : (double*)malloc(sizeof(double) * nctl);
    double* w = nctl <= 4096 ? wv : (double*)malloc(sizeof(double) * nctl);
    for (long q = 0; q < nctl; ++q) d[q] = SD[cidx[q] * N + c];
    const double sig = r_mad(d, nctl, 1, sc) / 0.6745;  /* :734 */
    double sw = 0.0;
    for (long q = 0; q < nctl; ++q) { w[q] = psi_huber(d[q], k_huber * sig) * SI[cidx[q] * N + c]; sw += w[q]; }
    const double mw = sw / (double)nctl;
    double Wc[MAXK];
    for (int j = 0; j < k; ++j) {
      double num = 0.0, den = 0.0;
      for (long q = 0; q < nctl; ++q) {
        const long g = cidx[q];
        double resid = Z[g * N + c] - gmean[g];  /* (Z - gmean)[ctl], :381 */
        for (int l = 0; l < j; ++l) resid -= alpha_new[g * k + l] * Wc[l];
        const double a = alpha_new[g * k + j], ww = w[q] / mw;
        num += ww * a * resid; den += ww * a * a;
      }
      Wc[j] = num / den;  /* Rfast::lmfit with one regressor, :737-741 */
    }
    for (int j = 0; j < k; ++j) W_new[c * k + j] = Wc[j];
    if (nctl > 4096) { free(d); free(w); }
  }
  /* :392-402 normalise, revert on NaN/Inf */
  int nbad_W = 0;
  for (int l = 0; l < k; ++l) {
    double ss = 0.0;
    for (long c = 0; c < N; ++c) ss += W_new[c * k + l] * W_new[c * k + l];
    const double nr = sqrt(ss);
    for (long c = 0; c < N; ++c) { W_new[c * k + l] /= nr; if (!(fabs(W_new[c * k + l]) <= 1.79769313486231570e308)) nbad_W++; }
  }
  if (nbad_W) memcpy(W_new, W, sizeof(double) * N * k);
  /* :405-433 gmean then Mb */
  int nbad_Mb = 0, nan_Mb = 0;
  double* Mb_new = (double*)malloc(sizeof(double) * G * m);
  double* gmean_new = (double*)malloc(sizeof(double) * G);
#pragma omp parallel for schedule(static) reduction(+ : nbad_Mb, nan_Mb)
  for (long g = 0; g < G; ++g) {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    double* sc = scratch + (long)t * sstride;
    double* zb = zbar + (long)t * m;  /* reused: per-group sum of weights */
    double s0j[512], s1j[512];
    double* S0 = m <= 512 ? s0j : (double*)malloc(sizeof(double) * m);
    double* S1 = m <= 512 ? s1j : (double*)malloc(sizeof(double) * m);
    (void)zb;
    const double sig = sigma[g];  /* same mad(signdev_g)/0.6745 as :409 */
    double sw = 0.0, acc = 0.0;
    for (long c = 0; c < N; ++c) {
      double hw = psi_huber(SD[g * N + c], k_huber * sig);  /* :410 */
      HW[g * N + c] = hw;
      double wt = hw * SI[g * N + c];                       /* :411 */
      double wa = 0.0;
      for (int l = 0; l < k; ++l) wa += alpha_new[g * k + l] * W_new[c * k + l];
      sc[c] = Z[g * N + c] - wa;  /* Z - alpha W' */
      sw += wt; acc += (sc[c] - Mb[g * m + grp[c]]) * wt;   /* Z.res of :405 */
    }
    gmean_new[g] = acc / sw;                                /* :412-413 */
    for (int j = 0; j < m; ++j) { S0[j] = 0.0; S1[j] = 0.0; }
    for (long c = 0; c < N; ++c) {
      double wt = HW[g * N + c] * SI[g * N + c];
      S0[grp[c]] += wt; S1[grp[c]] += (sc[c] - gmean_new[g]) * wt;  /* :418-420 */
    }
    for (int j = 0; j < m; ++j) {
      double v = S1[j] / (S0[j] + lambda_b);  /* rowWtAvg :667-675 */
      Mb_new[g * m + j] = v;
      if (v != v) nan_Mb++;
      if (!(fabs(v) <= 1.79769313486231570e308)) nbad_Mb++;
    }
    if (m > 512) { free(S0); free(S1); }
  }
  if (nan_Mb) memcpy(Mb_new, Mb, sizeof(double) * G * m);  /* :424 */
  for (long i = 0; i < G * m; ++i) if (!(fabs(Mb_new[i]) <= 1.79769313486231570e308)) Mb_new[i] = 0.0;  /* :425 */
#pragma omp parallel for schedule(static)
  for (long g = 0; g < G; ++g) {  /* :428-433 */
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    double* sc = scratch + (long)t * sstride;
    double med = r_median(Mb_new + g * m, m, 1, sc), mad = r_mad(Mb_new + g * m, m, 1, sc);
    double hi = med + 4.0 * mad, lo = med - 4.0 * mad;
    for (int j = 0; j < m; ++j) { double v = Mb_new[g * m + j]; if (v > hi) v = hi; if (v < lo) v = lo; Mb_new[g * m + j] = v; }
  }
  /* :469-484 candidate log-likelihood */
#pragma omp parallel for schedule(static)
  for (long g = 0; g < G; ++g) {
    const double r = 1.0 / psi[g];
    double s = 0.0;
    for (long c = 0; c < N; ++c) {
      double lmu = gmean_new[g] + Mb_new[g * m + grp[c]];
      for (int l = 0; l < k; ++l) lmu += alpha_new[g * k + l] * W_new[c * k + l];
      s += nb_logpmf((double)Y[g * N + c], lmu, exp(lmu), r);
    }
    loglik_out[g] = s;
  }
  memcpy(alpha, alpha_new, sizeof(double) * G * k);
  memcpy(adev, adev_new, sizeof(double) * G * k);
  memcpy(W, W_new, sizeof(double) * N * k);
  memcpy(gmean, gmean_new, sizeof(double) * G);
  memcpy(Mb, Mb_new, sizeof(double) * G * m);
  bad[0] = nbad_alpha; bad[1] = nbad_W; bad[2] = nbad_Mb;
  free(Z); free(SI); free(SD); free(HW); free(RMW); free(Wbar); free(gn); free(Ab); free(Ag); free(cg); free(scratch);
  free(alpha_new); free(adev_new); free(W_new); free(sigma); free(zbar); free(cidx); free(Mb_new); free(gmean_new);
  return 0;
}

int irls_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
