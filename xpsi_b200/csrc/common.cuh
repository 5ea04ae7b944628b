// Device-side numerical primitives shared by the xpsi_b200 kernels (sm_100a).
//
// Every routine restates the arithmetic the reference obtains from GSL through
// xpsi/include/GSL.pxd (Steffen / Akima interpolants, interval search, exact
// piecewise-cubic integrals) in a stateless, per-thread form: no accelerator
// caches, no allocation, coefficients rebuilt in registers where they are used.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace xb {

// xpsi/global_imports.py:70-80
constexpr double kC = 2.99792458e8;
constexpr double kKeV = 1.60217662e-16;
constexpr double kKB = 1.38064852e-23;
constexpr double kHKeV = 4.135667662e-18;
constexpr double kPi = 3.14159265358979323846;
constexpr double kHalfPi = 1.57079632679489661923;
constexpr double kTwoPi = 6.28318530717958647692;
constexpr double kLn10 = 2.30258509299404568402;
constexpr double kKBOverKeV = kKB / kKeV;
constexpr double kErg = 1.0e-7;
constexpr double kPlanckDistConst = 5.040366110812353e22;   // hot_BB.pyx:28

// status codes shared with include/xpsi_b200.h
constexpr int kOk = 0;
constexpr int kNumericalError = 1;      // the reference's (ERROR, None)
constexpr int kUnsupported = 3;         // configuration outside what the kernels cover

// xpsi/tools/core.pyx:118-122
__device__ __forceinline__ bool are_equal(double x, double y, double eps = 1.0e-12) {
  return fabs(x - y) < eps;
}

// xpsi/cellmesh/rays.pyx:53-57
__device__ __forceinline__ double eval_image_deflection(int order, double psi) {
  if (order & 1) return (double)(order + 1) * kPi - psi;
  return (double)order * kPi + psi;
}

// Strided read-only view, so reversed ray arrays need no copy
// (integrator_for_azimuthal_invariance.pyx:213-226 builds reversed copies).
struct View {
  const double* p;
  int stride;
  __device__ __forceinline__ double operator[](int i) const { return p[(long)i * stride]; }
};

// Largest i in [0, n-2] with x[i] <= q   (gsl_interp_bsearch contract).
template <class X>
__device__ __forceinline__ int interval_search(const X& x, int n, double q) {
  int lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    int mid = (hi + lo) >> 1;
    if (x[mid] > q) hi = mid; else lo = mid;
  }
  return lo;
}

// Same contract, starting from a guess and walking (for near-uniform grids).
template <class X>
__device__ __forceinline__ int interval_walk(const X& x, int n, double q, int guess) {
  int i = guess < 0 ? 0 : (guess > n - 2 ? n - 2 : guess);
  while (i > 0 && x[i] > q) --i;
  while (i < n - 2 && x[i + 1] <= q) ++i;
  return i;
}

// ---- Steffen (1990) monotone cubic ---------------------------------------
__device__ __forceinline__ double steffen_sign(double s) { return s < 0.0 ? -1.0 : 1.0; }

template <class X, class Y>
__device__ __forceinline__ double steffen_node_slope(const X& x, const Y& y, int n, int i) {
  if (i == 0) return (y[1] - y[0]) / (x[1] - x[0]);
  if (i == n - 1) return (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]);
  const double hi = x[i + 1] - x[i], him1 = x[i] - x[i - 1];
  const double si = (y[i + 1] - y[i]) / hi, sim1 = (y[i] - y[i - 1]) / him1;
  const double pi = (sim1 * hi + si * him1) / (him1 + hi);
  const double m = fmin(fabs(sim1), fmin(fabs(si), 0.5 * fabs(pi)));
  return (steffen_sign(sim1) + steffen_sign(si)) * m;
}

// value and first derivative at q lying in interval i
template <class X, class Y>
__device__ __forceinline__ void steffen_eval(const X& x, const Y& y, int n, int i, double q,
                                             double* val, double* der) {
  const double h = x[i + 1] - x[i];
  const double s = (y[i + 1] - y[i]) / h;
  const double y0 = steffen_node_slope(x, y, n, i);
  const double y1 = steffen_node_slope(x, y, n, i + 1);
  const double a = (y0 + y1 - 2.0 * s) / h / h;
  const double b = (3.0 * s - 2.0 * y0 - y1) / h;
  const double d = q - x[i];
  *val = y[i] + d * (y0 + d * (b + d * a));
  *der = y0 + d * (2.0 * b + d * 3.0 * a);
}

// ---- Akima (1970) ----------------------------------------------------------
// Coefficients (b,c,d) of interval i for nodes (x,y), n nodes; periodic uses the
// wrapped ghost slopes, natural the extrapolated ones.
template <class X, class Y>
__device__ __forceinline__ double akima_slope(const X& x, const Y& y, int n, int i, bool periodic) {
  // interval slope m_i for i in [-2, n]; ghosts per end condition
  if (i >= 0 && i <= n - 2) return (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
  if (periodic) {
    const int w = i < 0 ? i + (n - 1) : i - (n - 1);
    return (y[w + 1] - y[w]) / (x[w + 1] - x[w]);
  }
  const double m0 = (y[1] - y[0]) / (x[1] - x[0]);
  const double m1 = (y[2] - y[1]) / (x[2] - x[1]);
  const double mn2 = (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]);
  const double mn3 = (y[n - 2] - y[n - 3]) / (x[n - 2] - x[n - 3]);
  if (i == -1) return 2.0 * m0 - m1;
  if (i == -2) return 3.0 * m0 - 2.0 * m1;
  if (i == n - 1) return 2.0 * mn2 - mn3;
  return 3.0 * mn2 - 2.0 * mn3;   // i == n
}

__device__ __forceinline__ void akima_from_slopes(double mm2, double mm1, double m0, double mp1,
                                                  double mp2, double h, double* b, double* c,
                                                  double* d) {
  const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
  if (NE == 0.0) { *b = m0; *c = 0.0; *d = 0.0; return; }
  const double NE_next = fabs(mp2 - mp1) + fabs(m0 - mm1);
  const double alpha = fabs(mm1 - mm2) / NE;
  double tL;
  if (NE_next == 0.0) tL = m0;
  else {
    const double alpha1 = fabs(m0 - mm1) / NE_next;
    tL = (1.0 - alpha1) * m0 + alpha1 * mp1;
  }
  const double bb = (1.0 - alpha) * mm1 + alpha * m0;
  *b = bb;
  *c = (3.0 * m0 - 2.0 * bb - tL) / h;
  *d = (bb + tL - 2.0 * m0) / (h * h);
}

// same, with the interval width supplied as its reciprocal
__device__ __forceinline__ void akima_from_slopes_ih(double mm2, double mm1, double m0, double mp1,
                                                     double mp2, double ih, double* b, double* c,
                                                     double* d) {
  const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
  if (NE == 0.0) { *b = m0; *c = 0.0; *d = 0.0; return; }
  const double NE_next = fabs(mp2 - mp1) + fabs(m0 - mm1);
  const double alpha = fabs(mm1 - mm2) / NE;
  double tL;
  if (NE_next == 0.0) tL = m0;
  else {
    const double alpha1 = fabs(m0 - mm1) / NE_next;
    tL = (1.0 - alpha1) * m0 + alpha1 * mp1;
  }
  const double bb = (1.0 - alpha) * mm1 + alpha * m0;
  *b = bb;
  *c = (3.0 * m0 - 2.0 * bb - tL) * ih;
  *d = (bb + tL - 2.0 * m0) * (ih * ih);
}

template <class X, class Y>
__device__ __forceinline__ void akima_coeffs(const X& x, const Y& y, int n, int i, bool periodic,
                                             double* b, double* c, double* d) {
  akima_from_slopes(akima_slope(x, y, n, i - 2, periodic), akima_slope(x, y, n, i - 1, periodic),
                    akima_slope(x, y, n, i, periodic), akima_slope(x, y, n, i + 1, periodic),
                    akima_slope(x, y, n, i + 2, periodic), x[i + 1] - x[i], b, c, d);
}

// the same with the interval slopes m_0 .. m_{n-2} computed once (sm[i], the identical expression) instead of five
// times per interval: the ghost slopes at the two ends still go through akima_slope
template <class X, class Y>
__device__ __forceinline__ void akima_coeffs_cached(const double* sm, const X& x, const Y& y, int n, int i, bool periodic,
                                                    double* b, double* c, double* d) {
  auto m = [&](int k) -> double { return (k >= 0 && k <= n - 2) ? sm[k] : akima_slope(x, y, n, k, periodic); };
  akima_from_slopes(m(i - 2), m(i - 1), sm[i], m(i + 1), m(i + 2), x[i + 1] - x[i], b, c, d);
}

// Steffen in the same (b,c,d) form: y = y_i + t(b + t(c + t d))
template <class X, class Y>
__device__ __forceinline__ void steffen_coeffs(const X& x, const Y& y, int n, int i, double* b,
                                               double* c, double* d) {
  const double h = x[i + 1] - x[i];
  const double s = (y[i + 1] - y[i]) / h;
  const double y0 = steffen_node_slope(x, y, n, i);
  const double y1 = steffen_node_slope(x, y, n, i + 1);
  *b = y0;
  *c = (3.0 * s - 2.0 * y0 - y1) / h;
  *d = (y0 + y1 - 2.0 * s) / h / h;
}

// phase/energy interpolant selector (xpsi/tools/core.pyx:21): 0 Akima, 1 Steffen
enum Interp { kAkima = 0, kSteffen = 1, kCubic = 2 };

template <class X, class Y>
__device__ __forceinline__ void interp_coeffs(int kind, bool periodic, const X& x, const Y& y, int n,
                                              int i, double* b, double* c, double* d) {
  if (kind == kSteffen) steffen_coeffs(x, y, n, i, b, c, d);
  else akima_coeffs(x, y, n, i, periodic, b, c, d);
}

// ---- C2 cubic spline (GSL cspline / cspline_periodic) -----------------------------------------
// Unlike Akima / Steffen this interpolant is global: the second-derivative-like coefficients c_i solve a
// tridiagonal system (natural end conditions c_0 = c_{n-1} = 0) or a cyclic one (periodic: c_0 = c_{n-1}),
// solved here by one thread with the Thomas algorithm (plus a Sherman-Morrison correction for the cyclic
// corner entries).  The interval polynomial is y_i + t (b_i + t (c_i + t d_i)) with
// b_i = dy/dx - dx (c_{i+1} + 2 c_i) / 3,  d_i = (c_{i+1} - c_i) / (3 dx).
constexpr int kMaxCubicNodes = 260;

template <class X, class Y>
__device__ void cspline_second(const X& x, const Y& y, int n, bool periodic, double* c) {
  double cp[kMaxCubicNodes], q[kMaxCubicNodes];
  if (!periodic) {
    c[0] = 0.0; c[n - 1] = 0.0;
    const int m = n - 2;
    if (m <= 0) return;
    // row i (unknown c_{i+1}): sub h_i, diag 2 (h_i + h_{i+1}), super h_{i+1}, rhs 3 (dy_{i+1}/h_{i+1} - dy_i/h_i)
    for (int i = 0; i < m; ++i) {
      const double hi = x[i + 1] - x[i], hi1 = x[i + 2] - x[i + 1];
      const double g = 3.0 * ((y[i + 2] - y[i + 1]) / hi1 - (y[i + 1] - y[i]) / hi);
      const double diag = 2.0 * (hi1 + hi);
      if (i == 0) { cp[0] = hi1 / diag; c[1] = g / diag; }
      else {
        const double den = diag - hi * cp[i - 1];
        cp[i] = hi1 / den;
        c[i + 1] = (g - hi * c[i]) / den;
      }
    }
    for (int i = m - 2; i >= 0; --i) c[i + 1] -= cp[i] * c[i + 2];
    return;
  }
  const int m = n - 1;                      // unknowns u_i = c_{i+1}, i = 0 .. m-1 (c_0 = c_{n-1})
  if (m < 3) { for (int i = 0; i < n; ++i) c[i] = 0.0; return; }
  auto hh = [&](int i) -> double { return (i < m) ? x[i + 1] - x[i] : x[1] - x[0]; };          // h_i, cyclic
  auto dy = [&](int i) -> double { return (i < m) ? y[i + 1] - y[i] : y[1] - y[0]; };
  const double h0 = hh(0);
  const double diag0 = 2.0 * (hh(0) + hh(1));
  const double gam = -diag0;
  // Thomas on T (diag_0 - gam, diag_{m-1} - h0 h0 / gam) for rhs g -> c[1..m] and for u = (gam, 0, ..., 0, h0) -> q
  for (int i = 0; i < m; ++i) {
    const double hi = hh(i), hi1 = hh(i + 1);
    const double g = 3.0 * (dy(i + 1) / hi1 - dy(i) / hi);
    double diag = 2.0 * (hi + hi1);
    if (i == 0) diag -= gam;
    if (i == m - 1) diag -= h0 * h0 / gam;
    const double u = (i == 0) ? gam : ((i == m - 1) ? h0 : 0.0);
    if (i == 0) { cp[0] = hi1 / diag; c[1] = g / diag; q[0] = u / diag; }
    else {
      const double den = diag - hi * cp[i - 1];
      cp[i] = hi1 / den;
      c[i + 1] = (g - hi * c[i]) / den;
      q[i] = (u - hi * q[i - 1]) / den;
    }
  }
  for (int i = m - 2; i >= 0; --i) { c[i + 1] -= cp[i] * c[i + 2]; q[i] -= cp[i] * q[i + 1]; }
  const double fact = (c[1] + h0 * c[m] / gam) / (1.0 + q[0] + h0 * q[m - 1] / gam);
  for (int i = 0; i < m; ++i) c[i + 1] -= fact * q[i];
  c[0] = c[m];
}

// coefficient quads (y, b, c, d) of every interval, written with stride `qs` doubles between intervals
template <class X, class Y>
__device__ void cspline_quads(const X& x, const Y& y, int n, bool periodic, double* out, int qs) {
  double c[kMaxCubicNodes];
  cspline_second(x, y, n, periodic, c);
  for (int i = 0; i + 1 < n; ++i) {
    const double dx = x[i + 1] - x[i], dyv = y[i + 1] - y[i];
    double* o = out + (long)i * qs;
    o[0] = y[i];
    o[1] = dyv / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    o[2] = c[i];
    o[3] = (c[i + 1] - c[i]) / (3.0 * dx);
  }
}

// exact integral of y0 + t(b + t(c + t d)) for t in [r1, r2]
__device__ __forceinline__ double cubic_piece_integral(double y0, double b, double c, double d,
                                                       double r1, double r2) {
  const double r12 = r1 + r2;
  const double q = r1 * r1 + r2 * r2;
  return (r2 - r1) * (y0 + 0.5 * b * r12 + (1.0 / 3.0) * c * (q + r1 * r2) + 0.25 * d * r12 * q);
}

// ---- 4-point Lagrange weights (hot_Num4D.pyx:356-409) ------------------------
// base node b = clamp(j-1, 0, n-4) for p[j] <= v <= p[j+1]  
template <class P>
__device__ __forceinline__ int lagrange_base(const P& p, int n, double v) {
  int j = interval_search(p, n, v);
  int b = j - 1;
  if (b < 0) b = 0;
  if (b > n - 4) b = n - 4;
  return b;
}

template <class P>
__device__ __forceinline__ void lagrange_weights(const P& p, int b, double v, double w[4]) {
  const double p0 = p[b], p1 = p[b + 1], p2 = p[b + 2], p3 = p[b + 3];
  const double d0 = v - p0, d1 = v - p1, d2 = v - p2, d3 = v - p3;
  w[0] = d1 * d2 * d3 * (1.0 / (p0 - p1) / (p0 - p2) / (p0 - p3));
  w[1] = d0 * d2 * d3 * (1.0 / (p1 - p0) / (p1 - p2) / (p1 - p3));
  w[2] = d0 * d1 * d3 * (1.0 / (p2 - p0) / (p2 - p1) / (p2 - p3));
  w[3] = d0 * d1 * d2 * (1.0 / (p3 - p0) / (p3 - p1) / (p3 - p2));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace xb
