// limma_host.h -- host-side restatement of what `estimateDisp.par` asks of limma (reference R/estimateDisp.par.R:166-168):
//     s2.fit <- limma::squeezeVar(s2, df = df.residual, covariate = AveLogCPM[sel], robust = robust, winsor.tail.p = c(0.05, 0.1))
//     prior.df <- s2.fit$df.prior                      # one value (robust = FALSE) or one per gene (robust = TRUE)
// i.e. limma::fitFDist with a covariate (natural-spline trend of log s2, moment estimate of df2) and limma::fitFDistRobustly
// (lowess trend, Winsorized moments matched by Gauss-Legendre quadrature + Brent root, per-gene shrunken df2 for outliers).
// limma, stats::lowess, splines::ns and R's distribution functions are third-party code that is not under /root/reference:
// this is their published algorithm restated (PARITY UNPINNED); the CPU checker of the test-suite takes its
// F / chi-square functions from SciPy, so the special functions below are tested against an independent implementation
// (tests/test_limma_host.py through ruvnb_selftest_squeeze_var -- no device needed).  Plain C++, no CUDA.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

namespace limma_host {

static const double kInf = std::numeric_limits<double>::infinity();
static const double kNaN = std::numeric_limits<double>::quiet_NaN();

// ---- gamma-family functions ----------------------------------------------------------------------------------------
inline double digamma(double x) {
  double r = 0.0;
  while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + std::log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132)))));
}
inline double trigamma(double x) {
  double r = 0.0;
  while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + 1.0 / x + f / 2 + f / x * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f * (1.0 / 30 - f * (5.0 / 66)))));
}
inline double tetragamma(double x) {   // psi''(x)
  double r = 0.0;
  while (x < 6.0) { r -= 2.0 / (x * x * x); x += 1.0; }
  const double f = 1.0 / (x * x);
  return r - f - f / x - f * f * (0.5 - f * (1.0 / 6 - f * (1.0 / 6 - f * (3.0 / 10 - f * (5.0 / 6)))));
}
inline double trigamma_inverse(double x) {   // limma::trigammaInverse
  if (!(x == x)) return kNaN;
  if (x > 1e7) return 1.0 / std::sqrt(x);
  if (x < 1e-6) return 1.0 / x;
  double y = 0.5 + 1.0 / x;
  for (int it = 0; it < 50; ++it) {
    const double tri = trigamma(y);
    const double dif = tri * (1.0 - tri / x) / tetragamma(y);
    y += dif;
    if (-dif / y < 1e-8) break;
  }
  return y;
}

// ---- regularized incomplete beta / gamma in log space ----------------------------------------------------------------
// continued fraction of I_x(a, b) (modified Lentz); converges fast for x < (a + 1)/(a + b + 2)
inline double betacf(double a, double b, double x) {
  const double tiny = 1e-300, eps = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (std::fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 200000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < eps) break;
  }
  return h;
}
// log I_x(a, b) with x given as (x, 1 - x) so that neither end loses digits
inline double log_ibeta(double a, double b, double x, double xc) {
  if (x <= 0.0) return -kInf;
  if (xc <= 0.0) return 0.0;
  const double lfront = a * std::log(x) + b * std::log(xc) - (std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b));
  if (x < (a + 1.0) / (a + b + 2.0)) return lfront + std::log(betacf(a, b, x) / a);
  const double lother = lfront + std::log(betacf(b, a, xc) / b);   // log I_{1-x}(b, a)
  return std::log1p(-std::exp(lother));
}
// log P(F > f) for F ~ F(d1, d2), d2 finite
inline double pf_upper_log(double f, double d1, double d2) {
  if (!(f > 0.0)) return 0.0;
  if (f == kInf) return -kInf;
  const double den = d2 + d1 * f;
  return log_ibeta(0.5 * d2, 0.5 * d1, d2 / den, d1 * f / den);
}
inline double pf_lower(double f, double d1, double d2) {
  if (!(f > 0.0)) return 0.0;
  const double den = d2 + d1 * f;
  return std::exp(log_ibeta(0.5 * d1, 0.5 * d2, d1 * f / den, d2 / den));
}
inline double df_density(double f, double d1, double d2) {
  if (!(f > 0.0)) return 0.0;
  const double a = 0.5 * d1, b = 0.5 * d2, den = d2 + d1 * f;
  const double lx = std::log(d1 * f / den), lxc = std::log(d2 / den);
  const double lbeta = std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b);
  // dbeta(x; a, b) dx/df,  dx/df = d1 d2 / den^2
  return std::exp((a - 1.0) * lx + (b - 1.0) * lxc - lbeta + std::log(d1) + std::log(d2) - 2.0 * std::log(den));
}
// regularized gamma: log Q(a, x) (upper) and P(a, x) (lower)
inline double gamma_p_series(double a, double x) {   // P(a, x), x < a + 1
  double sum = 1.0 / a, del = sum, ap = a;
  for (int n = 0; n < 1000000; ++n) { ap += 1.0; del *= x / ap; sum += del; if (std::fabs(del) < std::fabs(sum) * 1e-16) break; }
  return sum * std::exp(-x + a * std::log(x) - std::lgamma(a));
}
inline double gamma_q_cf_log(double a, double x) {   // log Q(a, x), x >= a + 1
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 1000000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return -x + a * std::log(x) - std::lgamma(a) + std::log(h);
}
inline double pchisq_upper(double x, double df) {
  if (!(x > 0.0)) return 1.0;
  const double a = 0.5 * df, h = 0.5 * x;
  if (h < a + 1.0) return 1.0 - gamma_p_series(a, h);
  return std::exp(gamma_q_cf_log(a, h));
}
inline double pchisq_lower(double x, double df) {
  if (!(x > 0.0)) return 0.0;
  const double a = 0.5 * df, h = 0.5 * x;
  if (h < a + 1.0) return gamma_p_series(a, h);
  return 1.0 - std::exp(gamma_q_cf_log(a, h));
}
inline double dchisq(double x, double df) {
  if (!(x > 0.0)) return 0.0;
  const double a = 0.5 * df;
  return std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - std::lgamma(a)) * 0.5;
}
// quantile of a continuous distribution on (0, inf) by bisection in log f + secant polish; cdf must be increasing
template <class Cdf>
inline double quantile_pos(Cdf cdf, double p) {
  double lo = -700.0, hi = 700.0;   // log f
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (cdf(std::exp(mid)) < p) lo = mid; else hi = mid;
    if (hi - lo < 1e-15 * std::max(1.0, std::fabs(mid))) break;
  }
  return std::exp(0.5 * (lo + hi));
}
inline double qf_lower(double p, double d1, double d2) { return quantile_pos([&](double f) { return pf_lower(f, d1, d2); }, p); }
inline double qchisq_lower(double p, double df) { return quantile_pos([&](double x) { return pchisq_lower(x, df); }, p); }

// ---- small numerical tools ---------------------------------------------------------------------------------------------
// statmod::gauss.quad.prob(n, "uniform"): Gauss-Legendre nodes / weights mapped to (0, 1)
inline void gauss_legendre_01(int n, std::vector<double>& x, std::vector<double>& w) {
  x.assign(n, 0.0); w.assign(n, 0.0);
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = std::cos(pi * (i + 0.75) / (n + 0.5)), pp = 0.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double z1 = z;
      z = z1 - p1 / pp;
      if (std::fabs(z - z1) < 1e-16) break;
    }
    x[i] = 0.5 * (1.0 - z); x[n - 1 - i] = 0.5 * (1.0 + z);
    w[i] = w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}
// R's zeroin (Brent) on [ax, bx] with known end values
template <class F>
inline double brent_root(F f, double ax, double bx, double fa, double fb, double tol, int maxit = 1000) {
  double a = ax, b = bx, c = a, fc = fa;
  const double EPS = 2.220446049250313e-16;
  for (int it = 0; it <= maxit; ++it) {
    const double prev_step = b - a;
    if (std::fabs(fc) < std::fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    const double tol_act = 2.0 * EPS * std::fabs(b) + tol / 2.0;
    double new_step = (c - b) / 2.0;
    if (std::fabs(new_step) <= tol_act || fb == 0.0) return b;
    if (std::fabs(prev_step) >= tol_act && std::fabs(fa) > std::fabs(fb)) {
      double p, q; const double cb = c - b;
      if (a == c) { const double t1 = fb / fa; p = cb * t1; q = 1.0 - t1; }
      else { const double qq = fa / fc, t1 = fb / fc, t2 = fb / fa; p = t2 * (cb * qq * (qq - t1) - (b - a) * (t1 - 1.0)); q = (qq - 1.0) * (t1 - 1.0) * (t2 - 1.0); }
      if (p > 0.0) q = -q; else p = -p;
      if (p < (0.75 * cb * q - std::fabs(tol_act * q) / 2.0) && p < std::fabs(prev_step * q / 2.0)) new_step = p / q;
    }
    if (std::fabs(new_step) < tol_act) new_step = new_step > 0.0 ? tol_act : -tol_act;
    a = b; fa = fb;
    b += new_step; fb = f(b);
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) { c = a; fc = fa; }
  }
  return b;
}
// type-7 quantile of a SORTED vector
inline double quantile7_sorted(const std::vector<double>& s, double p) {
  const size_t n = s.size();
  const double index = 1.0 + (double)(n - 1) * p, lo = std::floor(index), hi = std::ceil(index);
  double q = s[(size_t)lo - 1];
  if (index > lo && s[(size_t)hi - 1] != q) { const double h = index - lo; q = (1.0 - h) * q + h * s[(size_t)hi - 1]; }
  return q;
}
inline double median_of(std::vector<double> v) {
  const size_t n = v.size();
  std::sort(v.begin(), v.end());
  return (n & 1) ? v[n / 2] : 0.5 * (v[n / 2 - 1] + v[n / 2]);
}
// R's rank(): average ranks for ties (1-based)
inline std::vector<double> rank_average(const std::vector<double>& v) {
  const size_t n = v.size();
  std::vector<size_t> o(n);
  std::iota(o.begin(), o.end(), 0);
  std::stable_sort(o.begin(), o.end(), [&](size_t a, size_t b) { return v[a] < v[b]; });
  std::vector<double> r(n);
  size_t i = 0;
  while (i < n) {
    size_t j = i;
    while (j + 1 < n && v[o[j + 1]] == v[o[i]]) ++j;
    const double rk = 0.5 * ((double)i + (double)j) + 1.0;
    for (size_t q = i; q <= j; ++q) r[o[q]] = rk;
    i = j + 1;
  }
  return r;
}
// least squares of y on the n x p matrix X (row-major, p <= 8) by Householder QR: fitted values and residual sum of squares;
// returns the numerical rank
inline int lstsq_fitted(std::vector<double> X, int n, int p, std::vector<double> y, std::vector<double>* fitted, double* rss) {
  std::vector<double> X0 = X, y0 = y;
  std::vector<double> beta(p, 0.0), rdiag(p, 0.0);
  int rank = p;
  for (int k = 0; k < p; ++k) {
    double nrm = 0.0;
    for (int i = k; i < n; ++i) nrm = std::hypot(nrm, X[(size_t)i * p + k]);
    if (nrm == 0.0) { rdiag[k] = 0.0; continue; }
    if (X[(size_t)k * p + k] < 0) nrm = -nrm;
    for (int i = k; i < n; ++i) X[(size_t)i * p + k] /= nrm;
    X[(size_t)k * p + k] += 1.0;
    for (int j = k + 1; j < p; ++j) {
      double s = 0.0;
      for (int i = k; i < n; ++i) s += X[(size_t)i * p + k] * X[(size_t)i * p + j];
      s = -s / X[(size_t)k * p + k];
      for (int i = k; i < n; ++i) X[(size_t)i * p + j] += s * X[(size_t)i * p + k];
    }
    double s = 0.0;
    for (int i = k; i < n; ++i) s += X[(size_t)i * p + k] * y[i];
    s = -s / X[(size_t)k * p + k];
    for (int i = k; i < n; ++i) y[i] += s * X[(size_t)i * p + k];
    rdiag[k] = -nrm;
  }
  double rmax = 0.0;
  for (int k = 0; k < p; ++k) rmax = std::max(rmax, std::fabs(rdiag[k]));
  for (int k = p - 1; k >= 0; --k) {
    if (std::fabs(rdiag[k]) <= 1e-7 * rmax) { beta[k] = 0.0; --rank; continue; }   // lm.fit's tolerance
    double s = y[k];
    for (int j = k + 1; j < p; ++j) s -= X[(size_t)k * p + j] * beta[j];
    beta[k] = s / rdiag[k];
  }
  fitted->assign(n, 0.0);
  double r2 = 0.0;
  for (int i = 0; i < n; ++i) {
    double f = 0.0;
    for (int j = 0; j < p; ++j) f += X0[(size_t)i * p + j] * beta[j];
    (*fitted)[i] = f;
    r2 += (y0[i] - f) * (y0[i] - f);
  }
  *rss = r2;
  return rank;
}
// a basis of the space splines::ns(x, df, intercept = TRUE) spans (natural cubic splines, boundary knots at the range, df - 2
// interior knots at the type-7 quantiles j/(df - 1)): truncated-power form on x scaled to [0, 1]
inline std::vector<double> ns_basis(const std::vector<double>& x, int df) {
  const int n = (int)x.size();
  std::vector<double> s(x);
  std::sort(s.begin(), s.end());
  const double lo = s.front(), hi = s.back();
  const int nik = df - 2, Kn = nik + 2;
  std::vector<double> kn(Kn);
  kn[0] = 0.0; kn[Kn - 1] = 1.0;
  for (int j = 0; j < nik; ++j) kn[j + 1] = (quantile7_sorted(s, (j + 1.0) / (nik + 1.0)) - lo) / (hi - lo);
  auto cube = [](double v) { return v > 0.0 ? v * v * v : 0.0; };
  std::vector<double> X((size_t)n * df);
  for (int i = 0; i < n; ++i) {
    const double t = (x[i] - lo) / (hi - lo);
    auto d = [&](int j) { return (cube(t - kn[j]) - cube(t - kn[Kn - 1])) / (kn[Kn - 1] - kn[j]); };
    X[(size_t)i * df + 0] = 1.0;
    if (df > 1) X[(size_t)i * df + 1] = t;
    for (int j = 0; j < Kn - 2; ++j) X[(size_t)i * df + 2 + j] = d(j) - d(Kn - 2);
  }
  return X;
}

// ---- limma::fitFDist(x, df1, covariate) --------------------------------------------------------------------------------------
struct FDistFit { std::vector<double> scale; double df2 = kNaN; };   // scale: one value per x when a covariate is given, else one
inline FDistFit fit_f_dist(std::vector<double> x, double df1, const std::vector<double>* covariate) {
  FDistFit out;
  const int n = (int)x.size();
  if (n == 0) return out;
  if (n == 1) { out.scale.assign(1, x[0]); out.df2 = 0.0; return out; }
  int splinedf = 1;
  if (covariate) {
    std::vector<double> u(*covariate);
    std::sort(u.begin(), u.end());
    const int nuniq = (int)(std::unique(u.begin(), u.end()) - u.begin());
    splinedf = std::min(1 + (n >= 3) + (n >= 6) + (n >= 30), nuniq);
    if (splinedf < 2) covariate = nullptr;
  }
  for (auto& v : x) v = std::max(v, 0.0);
  double m = median_of(x);
  if (m == 0.0) m = 1.0;
  for (auto& v : x) v = std::max(v, 1e-5 * m);
  std::vector<double> e(n);
  const double corr = -digamma(0.5 * df1) + std::log(0.5 * df1);
  for (int i = 0; i < n; ++i) e[i] = std::log(x[i]) + corr;
  std::vector<double> emean(n);
  double evar;
  if (!covariate) {
    double s = 0.0; for (double v : e) s += v;
    const double mu = s / n;
    double q = 0.0; for (double v : e) q += (v - mu) * (v - mu);
    evar = q / (n - 1.0);
    emean.assign(n, mu);
  } else {
    double rss = 0.0;
    const int rank = lstsq_fitted(ns_basis(*covariate, splinedf), n, splinedf, e, &emean, &rss);
    evar = rss / (double)(n - rank);
  }
  evar -= trigamma(0.5 * df1);
  if (evar > 0.0) {
    out.df2 = 2.0 * trigamma_inverse(evar);
    const double c2 = digamma(0.5 * out.df2) - std::log(0.5 * out.df2);
    out.scale.resize(covariate ? n : 1);
    for (size_t i = 0; i < out.scale.size(); ++i) out.scale[i] = std::exp(emean[i] + c2);
  } else {
    out.df2 = kInf;
    if (covariate) { out.scale.resize(n); for (int i = 0; i < n; ++i) out.scale[i] = std::exp(emean[i]); }
    else { double s = 0.0; for (double v : x) s += v; out.scale.assign(1, s / n); }
  }
  return out;
}

// ---- stats::lowess (clowess), x ascending -------------------------------------------------------------------------------------
inline void lowess(const std::vector<double>& x, const std::vector<double>& y, double f, int nsteps, double delta, std::vector<double>* ys_out) {
  const int n = (int)x.size();
  std::vector<double>& ys = *ys_out;
  ys.assign(n, 0.0);
  if (n < 2) { if (n == 1) ys[0] = y[0]; return; }
  std::vector<double> rw(n, 1.0), res(n, 0.0), w(n, 0.0);
  const int ns = std::max(2, std::min(n, (int)(f * n + 1e-7)));
  const double range = x[n - 1] - x[0];
  for (int iter = 1; iter <= nsteps + 1; ++iter) {
    int nleft = 0, nright = ns - 1, last = -1, i = 0;
    for (;;) {
      if (nright < n - 1) {
        const double d1 = x[i] - x[nleft], d2 = x[nright + 1] - x[i];
        if (d1 > d2) { ++nleft; ++nright; continue; }
      }
      // lowest(): fitted value at x[i]
      bool ok;
      {
        const double xs = x[i];
        const double h = std::max(xs - x[nleft], x[nright] - xs), h9 = 0.999 * h, h1 = 0.001 * h;
        double a = 0.0;
        int j = nleft;
        while (j < n) {
          w[j] = 0.0;
          const double r = std::fabs(x[j] - xs);
          if (r <= h9) {
            if (r <= h1) w[j] = 1.0; else { const double q = r / h, c3 = 1.0 - q * q * q; w[j] = c3 * c3 * c3; }
            if (iter > 1) w[j] *= rw[j];
            a += w[j];
          } else if (x[j] > xs) break;
          ++j;
        }
        const int nrt = j - 1;
        if (a <= 0.0) ok = false;
        else {
          ok = true;
          for (j = nleft; j <= nrt; ++j) w[j] /= a;
          if (h > 0.0) {
            a = 0.0;
            for (j = nleft; j <= nrt; ++j) a += w[j] * x[j];
            double b = xs - a, c = 0.0;
            for (j = nleft; j <= nrt; ++j) c += w[j] * (x[j] - a) * (x[j] - a);
            if (std::sqrt(c) > 0.001 * range) { b /= c; for (j = nleft; j <= nrt; ++j) w[j] *= (b * (x[j] - a) + 1.0); }
          }
          double v = 0.0;
          for (j = nleft; j <= nrt; ++j) v += w[j] * y[j];
          ys[i] = v;
        }
      }
      if (!ok) ys[i] = y[i];
      if (last < i - 1) {
        const double denom = x[i] - x[last];
        for (int j = last + 1; j < i; ++j) { const double alpha = (x[j] - x[last]) / denom; ys[j] = alpha * ys[i] + (1.0 - alpha) * ys[last]; }
      }
      last = i;
      const double cut = x[last] + delta;
      for (i = last + 1; i < n; ++i) {
        if (x[i] > cut) break;
        if (x[i] == x[last]) { ys[i] = ys[last]; last = i; }
      }
      i = std::max(last + 1, i - 1);
      if (last >= n - 1) break;
    }
    for (int q = 0; q < n; ++q) res[q] = y[q] - ys[q];
    double sc = 0.0;
    for (int q = 0; q < n; ++q) sc += std::fabs(res[q]);
    sc /= n;
    if (iter > nsteps) break;
    for (int q = 0; q < n; ++q) rw[q] = std::fabs(res[q]);
    std::vector<double> srt(rw);
    std::sort(srt.begin(), srt.end());
    const int m1 = n / 2;
    const double cmad = (n % 2 == 0) ? 3.0 * (srt[m1] + srt[n - m1 - 1]) : 6.0 * srt[m1];
    if (cmad < 1e-7 * sc) break;
    const double c9 = 0.999 * cmad, c1 = 0.001 * cmad;
    for (int q = 0; q < n; ++q) {
      const double r = std::fabs(res[q]);
      if (r <= c1) rw[q] = 1.0;
      else if (r <= c9) { const double t = 1.0 - (r / cmad) * (r / cmad); rw[q] = t * t; }
      else rw[q] = 0.0;
    }
  }
}
// limma::loessFit(z, covariate, span = 0.4, iterations = 4) without weights = lowess on the points ordered by x
inline std::vector<double> loess_fit_trend(const std::vector<double>& z, const std::vector<double>& cov) {
  const size_t n = z.size();
  std::vector<size_t> o(n);
  std::iota(o.begin(), o.end(), 0);
  std::stable_sort(o.begin(), o.end(), [&](size_t a, size_t b) { return cov[a] < cov[b]; });
  std::vector<double> xs(n), ysrt(n), fit;
  for (size_t i = 0; i < n; ++i) { xs[i] = cov[o[i]]; ysrt[i] = z[o[i]]; }
  lowess(xs, ysrt, 0.4, 3, 0.01 * (xs[n - 1] - xs[0]), &fit);
  std::vector<double> out(n);
  for (size_t i = 0; i < n; ++i) out[o[i]] = fit[i];
  return out;
}

// ---- limma::fitFDistRobustly(x, df1 (constant), covariate, winsor.tail.p) ----------------------------------------------------
struct RobustFit { std::vector<double> scale, df2_shrunk; double df2 = kNaN; };
inline RobustFit fit_f_dist_robustly(std::vector<double> x, double df1, const std::vector<double>* covariate, double tail_lo, double tail_hi) {
  RobustFit out;
  const int n = (int)x.size();
  if (n < 2) { out.df2_shrunk.assign(n, kNaN); return out; }
  auto from_nonrobust = [&](const FDistFit& nr) {
    out.df2 = nr.df2; out.scale = nr.scale; out.df2_shrunk.assign(n, nr.df2);
    if ((int)out.scale.size() == 1) out.scale.assign(n, nr.scale[0]);
  };
  if (n == 2) { from_nonrobust(fit_f_dist(x, df1, covariate)); return out; }
  const double m = median_of(x);
  if (!(m > 0.0)) { out.df2_shrunk.assign(n, kNaN); return out; }   // limma stops: "x has non-positive median"
  for (auto& v : x) if (v < m * 1e-12) v = m * 1e-12;
  const FDistFit nr = fit_f_dist(x, df1, covariate);
  if (tail_lo < 1.0 / n && tail_hi < 1.0 / n) { from_nonrobust(nr); return out; }
  std::vector<double> z(n), ztrend(n), zresid(n);
  for (int i = 0; i < n; ++i) z[i] = std::log(x[i]);
  if (!covariate) {   // mean(z, trim = winsor.tail.p[2])
    std::vector<double> s(z);
    std::sort(s.begin(), s.end());
    const int lo = (int)std::floor(n * tail_hi);
    double t = 0.0;
    for (int i = lo; i < n - lo; ++i) t += s[i];
    ztrend.assign(n, t / (n - 2 * lo));
  } else {
    ztrend = loess_fit_trend(z, *covariate);
  }
  for (int i = 0; i < n; ++i) zresid[i] = z[i] - ztrend[i];
  std::vector<double> srt(zresid);
  std::sort(srt.begin(), srt.end());
  const double q0 = quantile7_sorted(srt, tail_lo), q1 = quantile7_sorted(srt, 1.0 - tail_hi);
  double zwmean = 0.0;
  std::vector<double> zw(n);
  for (int i = 0; i < n; ++i) { zw[i] = std::min(std::max(zresid[i], q0), q1); zwmean += zw[i]; }
  zwmean /= n;
  double zwvar = 0.0;
  for (int i = 0; i < n; ++i) zwvar += (zw[i] - zwmean) * (zw[i] - zwmean);
  zwvar = zwvar / n * n / (n - 1.0);
  std::vector<double> gx, gw;
  gauss_legendre_01(128, gx, gw);
  auto moments = [&](double df2, double* mean, double* var) {
    if (df2 > 1e6) df2 = kInf;   // beyond this the F quantiles / density are the chi-square ones to 1e-6, and the continued fractions slow
    double fq0, fq1;
    if (df2 == kInf) { fq0 = qchisq_lower(tail_lo, df1) / df1; fq1 = qchisq_lower(1.0 - tail_hi, df1) / df1; }
    else { fq0 = qf_lower(tail_lo, df1, df2); fq1 = qf_lower(1.0 - tail_hi, df1, df2); }
    const double zq0 = std::log(fq0), zq1 = std::log(fq1), l0 = fq0 / (1.0 + fq0), l1 = fq1 / (1.0 + fq1), q21 = l1 - l0;
    std::vector<double> zn(128), fd(128);
    double s1 = 0.0;
    for (int i = 0; i < 128; ++i) {
      const double node = l0 + q21 * gx[i], fn = node / (1.0 - node);
      zn[i] = std::log(fn);
      const double dens = df2 == kInf ? dchisq(fn * df1, df1) * df1 : df_density(fn, df1, df2);
      fd[i] = dens / ((1.0 - node) * (1.0 - node));
      s1 += gw[i] * fd[i] * zn[i];
    }
    const double wm = q21 * s1 + zq0 * tail_lo + zq1 * tail_hi;
    double s2 = 0.0;
    for (int i = 0; i < 128; ++i) s2 += gw[i] * fd[i] * (zn[i] - wm) * (zn[i] - wm);
    *mean = wm;
    *var = q21 * s2 + (zq0 - wm) * (zq0 - wm) * tail_lo + (zq1 - wm) * (zq1 - wm) * tail_hi;
  };
  double mom_mean, mom_var;
  moments(kInf, &mom_mean, &mom_var);
  const double funval_inf = std::log(zwvar / mom_var);
  out.scale.resize(n); out.df2_shrunk.resize(n);
  std::vector<double> fstat(n), ltp(n);
  auto order_by = [&](const std::vector<double>& key) {
    std::vector<size_t> o(n);
    std::iota(o.begin(), o.end(), 0);
    std::stable_sort(o.begin(), o.end(), [&](size_t a, size_t b) { return key[a] < key[b]; });
    return o;
  };
  if (funval_inf <= 0.0) {
    out.df2 = kInf;
    for (int i = 0; i < n; ++i) { const double ztc = ztrend[i] + zwmean - mom_mean; out.scale[i] = std::exp(ztc); fstat[i] = std::exp(z[i] - ztc); }
    const std::vector<double> r = rank_average(fstat);
    std::vector<double> tailp(n);
    for (int i = 0; i < n; ++i) {
      tailp[i] = pchisq_upper(fstat[i] * df1, df1);
      const double emp = ((double)n - r[i] + 0.5) / n, pno = std::min(tailp[i] / emp, 1.0);
      out.df2_shrunk[i] = pno < 1.0 ? pno * n * df1 : kInf;
    }
    const std::vector<size_t> o = order_by(tailp);
    double cm = -kInf;
    for (size_t q = 0; q < o.size(); ++q) { cm = std::max(cm, out.df2_shrunk[o[q]]); out.df2_shrunk[o[q]] = cm; }
    return out;
  }
  if (nr.df2 == kInf) { from_nonrobust(nr); return out; }
  auto fun = [&](double u) { double mm, vv; moments(u / (1.0 - u), &mm, &vv); return std::log(zwvar / vv); };
  const double rbx = nr.df2 / (1.0 + nr.df2);
  const double f_low = fun(rbx);
  double df2;
  if (f_low >= 0.0) df2 = nr.df2;
  else { const double u = brent_root(fun, rbx, 1.0, f_low, funval_inf, 1e-8); df2 = u / (1.0 - u); }
  out.df2 = df2;
  if (df2 == kInf || !(df2 == df2)) {   // the root sat at the upper end: the chi-square limit
    moments(kInf, &mom_mean, &mom_var);
    for (int i = 0; i < n; ++i) out.scale[i] = std::exp(ztrend[i] + zwmean - mom_mean);
    out.df2 = kInf; out.df2_shrunk.assign(n, kInf);
    return out;
  }
  moments(df2, &mom_mean, &mom_var);
  double min_ltp = kInf, fmax = -kInf;
  for (int i = 0; i < n; ++i) {
    const double ztc = ztrend[i] + zwmean - mom_mean;
    out.scale[i] = std::exp(ztc);
    fstat[i] = std::exp(z[i] - ztc);
    ltp[i] = pf_upper_log(fstat[i], df1, df2);
    min_ltp = std::min(min_ltp, ltp[i]); fmax = std::max(fmax, fstat[i]);
  }
  const std::vector<double> r = rank_average(fstat);
  std::vector<double> lpno(n);
  bool any = false;
  for (int i = 0; i < n; ++i) { lpno[i] = std::min(ltp[i] - (std::log((double)n - r[i] + 0.5) - std::log((double)n)), 0.0); any |= lpno[i] < 0.0; }
  if (!any) { out.df2_shrunk.assign(n, df2); return out; }
  if (min_ltp == -kInf) {
    for (int i = 0; i < n; ++i) out.df2_shrunk[i] = std::exp(lpno[i]) * df2;
  } else {
    double df2_out = std::log(0.5) / min_ltp * df2;
    const double new_ltp = pf_upper_log(fmax, df1, df2_out);
    df2_out = std::log(0.5) / new_ltp * df2_out;
    for (int i = 0; i < n; ++i) out.df2_shrunk[i] = std::exp(lpno[i]) * df2 + (-std::expm1(lpno[i])) * df2_out;
  }
  const std::vector<size_t> o = order_by(ltp);
  std::vector<double> so(n);
  for (int i = 0; i < n; ++i) so[i] = out.df2_shrunk[o[i]];
  double cs = 0.0, best = kInf; int imin = 0;
  for (int i = 0; i < n; ++i) { cs += so[i]; const double mm = cs / (i + 1.0); if (mm < best) { best = mm; imin = i; } }
  for (int i = 0; i <= imin; ++i) so[i] = best;
  double cm = -kInf;
  for (int i = 0; i < n; ++i) { cm = std::max(cm, so[i]); out.df2_shrunk[o[i]] = cm; }
  return out;
}

// df.prior of limma::squeezeVar(var = s2, df = df1, covariate, robust, winsor.tail.p): n values (all equal when !robust)
inline std::vector<double> squeeze_var_df_prior(const std::vector<double>& s2, double df1, const std::vector<double>* covariate, bool robust,
                                                double tail_lo = 0.05, double tail_hi = 0.1) {
  if (robust) return fit_f_dist_robustly(s2, df1, covariate, tail_lo, tail_hi).df2_shrunk;
  return std::vector<double>(s2.size(), fit_f_dist(s2, df1, covariate).df2);
}

}  // namespace limma_host
