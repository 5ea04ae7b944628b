// GPU "embed": cell meshes of hot-region members and of the closed surface, and their rays, from model parameters.
//
// Replaces the per-parameter-vector producer of all integrator inputs for hot regions made of
// circular members -- plain spots, superseding members with an omission hole, ceding members masked by their
// superseding region (the two sharing the region's cell budget), members covering a rotational pole (polar
// variant, cellmesh/polar_mesh.pyx) -- and for the closed equal-area mesh of Elsewhere / Everywhere
// (cellmesh/global_mesh.pyx):
//   xpsi/HotRegion.py:774-870  __construct_cellMesh -> mesh_tools.allocate_cells (:892-1000),
//                              mesh.construct_spot_cellMesh (mesh.pyx:18-417)
//   xpsi/HotRegion.py:903-951  __compute_rays -> rays.compute_rays (rays.pyx:249-390)
//   xpsi/HotRegion.py:880-901  __calibrate_lag;  :953-1017 __compute_cellParamVecs
//   surface_radiation_field/effective_gravity_universal.pyx
// The reference evaluates the underlying 1-D integrals with adaptive GSL rules (CQUAD at
// epsrel 1e-8 for the mesh, QAG-GK61 at 1e-12 for rays); here they are fixed-order
// Gauss-Legendre rules on intervals split at the integrand's kinks, with a sqrt
// substitution where the spot boundary is tangent to a parallel -- converged to <= 1e-12,
// which is tighter than the reference's own tolerance.  Discretisation choices that are
// *not* converged quantities are reproduced exactly: the 1000-node area table and its
// Steffen inverse interpolation for the ring parallels, the equal-area ring/cell layout,
// the 5-point boundary-cell test, the cos(alpha) ray grid.
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace xb {

__constant__ double c_gl16_x[16] = {-9.89400934991649939e-01, -9.44575023073232600e-01, -8.65631202387831755e-01, -7.55404408355002999e-01, -6.17876244402643771e-01, -4.58016777657227370e-01, -2.81603550779258915e-01, -9.50125098376374405e-02, 9.50125098376374405e-02, 2.81603550779258915e-01, 4.58016777657227370e-01, 6.17876244402643771e-01, 7.55404408355002999e-01, 8.65631202387831755e-01, 9.44575023073232600e-01, 9.89400934991649939e-01};
__constant__ double c_gl16_w[16] = {2.71524594117541762e-02, 6.22535239386474565e-02, 9.51585116824926053e-02, 1.24628971255534071e-01, 1.49595988816576708e-01, 1.69156519395002647e-01, 1.82603415044923639e-01, 1.89450610455068641e-01, 1.89450610455068641e-01, 1.82603415044923639e-01, 1.69156519395002647e-01, 1.49595988816576708e-01, 1.24628971255534071e-01, 9.51585116824926053e-02, 6.22535239386474565e-02, 2.71524594117541762e-02};
__constant__ double c_gl32_x[32] = {-9.97263861849481570e-01, -9.85611511545268382e-01, -9.64762255587506390e-01, -9.34906075937739667e-01, -8.96321155766052091e-01, -8.49367613732569970e-01, -7.94483795967942386e-01, -7.32182118740289711e-01, -6.63044266930215231e-01, -5.87715757240762304e-01, -5.06899908932229359e-01, -4.21351276130635333e-01, -3.31868602282127667e-01, -2.39287362252137065e-01, -1.44471961582796488e-01, -4.83076656877383243e-02, 4.83076656877383243e-02, 1.44471961582796488e-01, 2.39287362252137065e-01, 3.31868602282127667e-01, 4.21351276130635333e-01, 5.06899908932229359e-01, 5.87715757240762304e-01, 6.63044266930215231e-01, 7.32182118740289711e-01, 7.94483795967942386e-01, 8.49367613732569970e-01, 8.96321155766052091e-01, 9.34906075937739667e-01, 9.64762255587506390e-01, 9.85611511545268382e-01, 9.97263861849481570e-01};
__constant__ double c_gl32_w[32] = {7.01861000947050576e-03, 1.62743947309057432e-02, 2.53920653092620241e-02, 3.42738629130217645e-02, 4.28358980222268357e-02, 5.09980592623760914e-02, 5.86840934785355650e-02, 6.58222227763616829e-02, 7.23457941088483381e-02, 7.81938957870702278e-02, 8.33119242269467070e-02, 8.76520930044037833e-02, 9.11738786957637798e-02, 9.38443990808045109e-02, 9.56387200792747083e-02, 9.65400885147276594e-02, 9.65400885147276594e-02, 9.56387200792747083e-02, 9.38443990808045109e-02, 9.11738786957637798e-02, 8.76520930044037833e-02, 8.33119242269467070e-02, 7.81938957870702278e-02, 7.23457941088483381e-02, 6.58222227763616829e-02, 5.86840934785355650e-02, 5.09980592623760914e-02, 4.28358980222268357e-02, 3.42738629130217645e-02, 2.53920653092620241e-02, 1.62743947309057432e-02, 7.01861000947050576e-03};

// ---- oblate surface (AlGendy & Morsink 2014), mesh_tools.pyx:18-120 --------------------
__device__ __forceinline__ double radius_normalised(double mu, double eps, double zeta) {
  return 1.0 + eps * (-0.788 + 1.030 * zeta) * mu * mu;
}
__device__ __forceinline__ double f_theta(double mu, double rn, double eps, double zeta) {
  const double d = -2.0 * eps * (-0.788 + 1.030 * zeta) * mu * sqrt(1.0 - mu * mu);
  return d / (rn * sqrt(1.0 - 2.0 * zeta / rn));
}
// surface-area element per unit azimuth / R_eq^2 (mesh_tools.pyx:99-120); av=1 weights by theta.  sin and cos of
// theta come from one sincos (sqrt(1 - mu^2) = sin theta on [0, pi]) and f_theta's rn sqrt(1 - 2 zeta / rn) is formed
// as sqrt(rn (rn - 2 zeta)): one reciprocal square root and one square root instead of three square roots and two
// divisions -- the same value to rounding, and this function is 40 % of the mesh kernel's samples
__device__ __forceinline__ double area_element_sc(double theta, double st, double mu, double eps, double zeta, int av) {
  if (are_equal(theta, 0.0)) return 0.0;
  const double rn = radius_normalised(mu, eps, zeta);
  const double f = (-2.0 * eps * (-0.788 + 1.030 * zeta) * mu * st) * rsqrt(rn * (rn - 2.0 * zeta));
  const double v = rn * rn * sqrt(1.0 + f * f) * st;
  return av ? theta * v : v;
}
__device__ __forceinline__ double area_element(double theta, double eps, double zeta, int av) {
  double st, mu;
  sincos(theta, &st, &mu);
  return area_element_sc(theta, st, mu, eps, zeta, av);
}
__device__ __forceinline__ double integrate_area(double lo, double hi, double eps, double zeta, int av) {
  const double h = 0.5 * (hi - lo), m = 0.5 * (hi + lo);
  double s = 0.0;
#pragma unroll 4
  for (int k = 0; k < 32; ++k) s += c_gl32_w[k] * area_element(m + h * c_gl32_x[k], eps, zeta, av);
  return h * s;
}
// effective_gravity_universal.pyx:12-71
__device__ __forceinline__ double effective_gravity(double mu, double R_eq, double x, double eps) {
  const double g_0 = x * kC * kC / (R_eq * sqrt(1.0 - 2.0 * x));
  const double esq = eps, esqsq = eps * eps;
  const double c_e = -0.791 + 0.776 * x, c_p = 1.138 - 1.431 * x;
  const double d_e = (-1.315 + 2.431 * x) * esq * x, d_p = (0.653 - 2.864 * x) * esq * x;
  const double d_60 = (13.47 - 27.13 * x) * esq * x;
  const double f_e = -1.172 * x * esqsq, f_p = 0.975 * x * esqsq;
  double g = 1.0;
  g += (c_e + d_e + f_e) * esq * (1.0 - mu * mu);
  g += (c_p + d_p + f_p - d_60) * esq * mu * mu;
  g += d_60 * esq * fabs(mu);
  return log10(g * g_0) + 2.0;
}
__device__ __forceinline__ double eval_psi(double theta, double phi, double THETA) {   // mesh_tools.pyx:156-160
  return acos(cos(THETA) * cos(theta) + sin(THETA) * sin(theta) * cos(phi));
}
// half-width in azimuth of the spot at colatitude theta (mesh_tools.pyx:265-274)
__device__ __forceinline__ double spot_halfwidth_sc(double st, double ct, double cosT, double sinT, double cos_rho) {
  double c = (cos_rho - cosT * ct) / (sinT * st);
  if (c > 1.0) c = 1.0;
  if (c < -1.0) c = -1.0;
  return acos(c);
}
__device__ __forceinline__ double spot_halfwidth(double theta, double cosT, double sinT, double cos_rho) {
  double c = (cos_rho - cosT * cos(theta)) / (sinT * sin(theta));
  if (c > 1.0) c = 1.0;
  if (c < -1.0) c = -1.0;
  return acos(c);
}

// integral over theta in [s0, s1] of g(theta) with optional sqrt substitution at either end
template <class G>
__device__ __forceinline__ double gl16_piece(const G& g, double s0, double s1, bool sing0, bool sing1) {
  double tot = 0.0;
  if (!sing0 && !sing1) {
    const double h = 0.5 * (s1 - s0), m = 0.5 * (s1 + s0);
    for (int k = 0; k < 16; ++k) tot += c_gl16_w[k] * g(m + h * c_gl16_x[k]);
    return h * tot;
  }
  const double mid = (sing0 && sing1) ? 0.5 * (s0 + s1) : (sing0 ? s1 : s0);
  if (sing0) {             // theta = s0 + t^2
    const double L = sqrt(mid - s0), h = 0.5 * L;
    double s = 0.0;
    for (int k = 0; k < 16; ++k) { const double t = h + h * c_gl16_x[k]; s += c_gl16_w[k] * g(s0 + t * t) * 2.0 * t; }
    tot += h * s;
  }
  if (sing1) {             // theta = s1 - t^2
    const double L = sqrt(s1 - mid), h = 0.5 * L;
    double s = 0.0;
    for (int k = 0; k < 16; ++k) { const double t = h + h * c_gl16_x[k]; s += c_gl16_w[k] * g(s1 - t * t) * 2.0 * t; }
    tot += h * s;
  }
  return tot;
}

// area of cell [tha,thb] x [pa,pb] inside the spot / R_eq^2 (mesh_tools.pyx:303-388,429-473, superRadius = 0)
__device__ __noinline__ double spot_cell_area(double tha, double thb, double pa, double pb, double eps, double zeta,
                                 double TH, double rho) {
  const double cosT = cos(TH), sinT = sin(TH), cos_rho = cos(rho);
  const double lo = fmax(tha, TH - rho), hi = fmin(thb, TH + rho);
  if (!(hi > lo)) return 0.0;
  auto g = [&](double th) -> double {
    double st, ct;
    sincos(th, &st, &ct);
    const double a = spot_halfwidth_sc(st, ct, cosT, sinT, cos_rho);
    const double ov = fmin(pb, a) - fmax(pa, -a);
    return ov > 0.0 ? ov * area_element_sc(th, st, ct, eps, zeta, 0) : 0.0;
  };
  double bp[6];
  int nb = 0;
  bp[nb++] = lo; bp[nb++] = hi;
  for (int s = 0; s < 2; ++s) {          // colatitudes where the spot boundary crosses phi = pa, pb
    const double ph = s == 0 ? pa : pb;
    const double A = cosT, B = sinT * cos(ph), Rn = sqrt(A * A + B * B);
    if (fabs(cos_rho) <= Rn) {
      const double base = atan2(B, A), d = acos(cos_rho / Rn);
      if (base - d > lo && base - d < hi) bp[nb++] = base - d;
      if (base + d > lo && base + d < hi) bp[nb++] = base + d;
    }
  }
  for (int i = 1; i < nb; ++i) {         // insertion sort
    const double v = bp[i];
    int j = i - 1;
    while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; --j; }
    bp[j + 1] = v;
  }
  double tot = 0.0;
  for (int i = 0; i + 1 < nb; ++i) {
    const double s0 = bp[i], s1 = bp[i + 1];
    if (!(s1 - s0 > 1.0e-15)) continue;
    const bool sing0 = fabs(s0 - (TH - rho)) < 1.0e-14 || fabs(s0 - (TH + rho)) < 1.0e-14;
    const bool sing1 = fabs(s1 - (TH - rho)) < 1.0e-14 || fabs(s1 - (TH + rho)) < 1.0e-14;
    tot += gl16_piece(g, s0, s1, sing0, sing1);
  }
  return tot;
}

// ---- generic regions: omission hole / superseding mask, ceding partner, polar caps -----------------
// mesh_tools.pyx:262-271
__device__ __forceinline__ double eval_phi(double theta, double THETA, double psi) {
  double c = cos(psi) - cos(THETA) * cos(theta);
  c /= sin(THETA) * sin(theta);
  if (!(-1.0 <= c && c <= 1.0)) return -1.0;
  return acos(c);
}
// mesh_tools.pyx:273-301: length of [LB, UB] inside [a, b]
__device__ __forceinline__ double get_interval(double a, double b, double LB, double UB) {
  int ac = (LB <= a && a <= UB), bc = (LB <= b && b <= UB);
  if (ac == 0) ac = 2 * (int)(a > UB);
  if (bc == 0) bc = 2 * (int)(b > UB);
  if (ac == bc) return ac == 1 ? b - a : 0.0;
  if (bc == 1 && ac == 0) return b - LB;
  if (bc == 2 && ac == 1) return UB - a;
  if (bc == 2 && ac == 0) return UB - LB;
  return 0.0;
}
struct Region {            // the region being meshed and the region masking it (mesh.pyx naming: "cede" and "super")
  double colat, radius, hRadius, hAzi, hColat;
};
// azimuthal width of the region minus its mask at colatitude theta, restricted to [pa, pb] when cell != 0
// (cell_integrand mesh_tools.pyx:303-388, spot_integrand :690-771)
__device__ __noinline__ double region_width(const Region& g, double theta, double pa, double pb, int cell) {
  if (are_equal(theta, 0.0)) return 0.0;
  double aLB, aUB, cLB, cUB;
  aUB = eval_phi(theta, g.colat, g.radius);
  if (are_equal(aUB, -1.0)) {
    if (theta + g.colat < g.radius) { aUB = kPi; aLB = -kPi; }
    else return 0.0;
  } else aLB = -aUB;
  if (fabs(theta - g.hColat) > g.hRadius) cLB = cUB = 0.0;
  else {
    cUB = eval_phi(theta, g.hColat, g.hRadius);
    if (are_equal(cUB, -1.0)) return 0.0;
    cLB = -cUB; cUB += g.hAzi; cLB += g.hAzi;
    if (cUB > kPi) cUB -= kTwoPi;
    if (cLB < -kPi) cLB += kTwoPi;
  }
  auto iv = [&](double LB, double UB) -> double { return cell ? get_interval(pa, pb, LB, UB) : UB - LB; };
  if (cUB >= cLB) {
    if (cLB >= aUB || cUB <= aLB) return iv(aLB, aUB);
    if (cLB >= aLB && cUB <= aUB) return iv(aLB, cLB) + iv(cUB, aUB);
    if (cLB < aLB && aLB < cUB && cUB <= aUB) return iv(cUB, aUB);
    if (aLB <= cLB && cLB < aUB && aUB < cUB) return iv(aLB, cLB);
    return 0.0;
  }
  if (cUB >= aUB || cLB <= aLB) return 0.0;
  if (cUB >= aLB && cLB <= aUB) return iv(cUB, cLB);
  if (cUB < aLB && aLB < cLB && cLB <= aUB) return iv(aLB, cLB);
  if (aLB <= cUB && cUB < aUB && aUB < cLB) return iv(cUB, aUB);
  return iv(aLB, aUB);
}

__constant__ double c_k15_x[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                  0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                  0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                  0.207784955007898467600689403773245, 0.0};
__constant__ double c_k15_w[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                  0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                  0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                  0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
__constant__ double c_g7_w[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                                 0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

// Adaptive Gauss-Kronrod (7,15) by depth-first bisection: the reference integrates the same integrands with
// CQUAD at epsrel 1e-8 (mesh_tools.pyx:465-473,852-860); kinks and square-root end points are resolved by
// bisection until the panel's |K15 - G7| is below tol_abs.
template <class F>
__device__ __noinline__ double adaptive_gk15(const F& f, double A, double B, double tol_abs) {
  if (!(B > A)) return 0.0;
  double sa[44], sb[44];
  int sd[44];
  int top = 0;
  sa[0] = A; sb[0] = B; sd[0] = 0; top = 1;
  double total = 0.0;
  while (top > 0) {
    --top;
    const double a = sa[top], b = sb[top];
    const int depth = sd[top];
    const double h = 0.5 * (b - a), m = 0.5 * (a + b);
    const double fc = f(m);
    double K = c_k15_w[7] * fc, G = c_g7_w[3] * fc;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const double dx = h * c_k15_x[k];
      const double v = f(m - dx) + f(m + dx);
      K += c_k15_w[k] * v;
      if (k & 1) G += c_g7_w[k >> 1] * v;
    }
    K *= h; G *= h;
    if (fabs(K - G) <= tol_abs || depth >= 40 || top + 2 > 44) total += K;
    else {
      sa[top] = a; sb[top] = m; sd[top] = depth + 1; ++top;
      sa[top] = m; sb[top] = b; sd[top] = depth + 1; ++top;
    }
  }
  return total;
}

// area of the region (minus mask) between colatitudes lo and hi / R_eq^2, integrated by one warp
// (integrateSpot, mesh_tools.pyx:773-860, Lorentz = 0; it only feeds the cell allocation).  The range is split
// at the parallels tangent to either boundary circle, where the width behaves like a square root and a
// theta = s + t^2 substitution makes the integrand smooth; every lane then takes a slice of each piece.
__device__ __noinline__ double warp_region_area(const Region& g, double lo, double hi, double eps, double zeta, int lane) {
  double bp[6];
  int nb = 0;
  bp[nb++] = lo; bp[nb++] = hi;
  double tg[4];
  int nt = 0;
  tg[nt++] = fabs(g.colat - g.radius); tg[nt++] = g.colat + g.radius;
  if (g.hRadius > 0.0) { tg[nt++] = fabs(g.hColat - g.hRadius); tg[nt++] = g.hColat + g.hRadius; }
  for (int k = 0; k < nt; ++k) if (tg[k] > lo && tg[k] < hi) bp[nb++] = tg[k];
  for (int i = 1; i < nb; ++i) {
    const double v = bp[i];
    int j = i - 1;
    while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; --j; }
    bp[j + 1] = v;
  }
  auto f = [&](double th) -> double { return region_width(g, th, 0.0, 0.0, 0) * area_element(th, eps, zeta, 0); };
  auto is_tangent = [&](double x) -> bool {
    for (int k = 0; k < nt; ++k) if (fabs(x - tg[k]) < 1.0e-14) return true;
    return false;
  };
  const double tol = 1.0e-11;
  double part = 0.0;
  for (int i = 0; i + 1 < nb; ++i) {
    const double s0 = bp[i], s1 = bp[i + 1];
    if (!(s1 - s0 > 1.0e-15)) continue;
    const bool sing0 = is_tangent(s0), sing1 = is_tangent(s1);
    if (!sing0 && !sing1) {
      const double w = (s1 - s0) / 32.0;
      part += adaptive_gk15(f, s0 + w * lane, s0 + w * (lane + 1), tol);
      continue;
    }
    const double mid = (sing0 && sing1) ? 0.5 * (s0 + s1) : (sing0 ? s1 : s0);
    if (sing0) {
      const double w = sqrt(mid - s0) / 32.0;
      part += adaptive_gk15([&](double t) -> double { return f(s0 + t * t) * 2.0 * t; }, w * lane, w * (lane + 1), tol);
    }
    if (sing1) {
      const double w = sqrt(s1 - mid) / 32.0;
      part += adaptive_gk15([&](double t) -> double { return f(s1 - t * t) * 2.0 * t; }, w * lane, w * (lane + 1), tol);
    }
  }
  return warp_sum(part);
}

// colatitudes at which the small circle (centre colatitude TH at azimuth 0, angular radius rho) crosses the
// meridian at azimuth ph, appended to bp when inside (lo, hi)
__device__ __forceinline__ void circle_meridian_crossings(double TH, double rho, double ph, double lo, double hi,
                                                          double* bp, int* nb) {
  const double A = cos(TH), B = sin(TH) * cos(ph), Rn = sqrt(A * A + B * B), cr = cos(rho);
  if (fabs(cr) <= Rn && Rn > 0.0) {
    const double base = atan2(B, A), d = acos(cr / Rn);
    for (int s = -1; s <= 1; s += 2) {
      double t = base + s * d;
      if (t < 0.0) t = -t;                       // the same great circle continued through the pole
      if (t > kPi) t = kTwoPi - t;
      if (t > lo && t < hi) bp[(*nb)++] = t;
    }
  }
}

// integrateCell (mesh_tools.pyx:429-473) for a region with a mask: the colatitude range is first split where
// either boundary circle crosses the cell's meridians or touches a parallel, so every piece is either empty
// or smooth up to square-root end points (no sliver can fall between quadrature nodes); the pieces are then
// integrated adaptively (the two circles' mutual intersections are left to the bisection).
__device__ __noinline__ double region_cell_area(const Region& g, double l, double u, double pa, double pb, double eps,
                                   double zeta, double tol_abs) {
  double bp[24];
  int nb = 0;
  bp[nb++] = l; bp[nb++] = u;
  circle_meridian_crossings(g.colat, g.radius, pa, l, u, bp, &nb);
  circle_meridian_crossings(g.colat, g.radius, pb, l, u, bp, &nb);
  // parallels tangent to a boundary circle: the azimuthal width behaves like a square root there
  double tg[4];
  int nt = 0;
  tg[nt++] = fabs(g.colat - g.radius); tg[nt++] = g.colat + g.radius;
  if (g.hRadius > 0.0) {
    circle_meridian_crossings(g.hColat, g.hRadius, pa - g.hAzi, l, u, bp, &nb);
    circle_meridian_crossings(g.hColat, g.hRadius, pb - g.hAzi, l, u, bp, &nb);
    tg[nt++] = fabs(g.hColat - g.hRadius); tg[nt++] = g.hColat + g.hRadius;
  }
  for (int k = 0; k < nt; ++k) if (tg[k] > l && tg[k] < u) bp[nb++] = tg[k];
  for (int i = 1; i < nb; ++i) {         // insertion sort
    const double v = bp[i];
    int j = i - 1;
    while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; --j; }
    bp[j + 1] = v;
  }
  auto f = [&](double th) -> double { return region_width(g, th, pa, pb, 1) * area_element(th, eps, zeta, 0); };
  auto is_tangent = [&](double x) -> bool {
    for (int k = 0; k < nt; ++k) if (fabs(x - tg[k]) < 1.0e-14) return true;
    return false;
  };
  double tot = 0.0;
  for (int i = 0; i + 1 < nb; ++i) {
    const double s0 = bp[i], s1 = bp[i + 1];
    if (!(s1 - s0 > 1.0e-15)) continue;
    const bool sing0 = is_tangent(s0), sing1 = is_tangent(s1);
    if (!sing0 && !sing1) { tot += adaptive_gk15(f, s0, s1, tol_abs); continue; }
    // theta = s0 + t^2 (or s1 - t^2) turns the square-root end point into a smooth integrand
    const double mid = (sing0 && sing1) ? 0.5 * (s0 + s1) : (sing0 ? s1 : s0);
    if (sing0) tot += adaptive_gk15([&](double t) -> double { return f(s0 + t * t) * 2.0 * t; }, 0.0, sqrt(mid - s0), tol_abs);
    if (sing1) tot += adaptive_gk15([&](double t) -> double { return f(s1 - t * t) * 2.0 * t; }, 0.0, sqrt(s1 - mid), tol_abs);
  }
  return tot;
}

// The same adaptive rule evaluated by a whole warp: lane k < 15 owns Kronrod node k of the current panel, the
// sums are warp reductions and every lane takes the same accept / bisect decision (no divergence, and the
// serial chain is 15 times shorter than one thread per cell, which left 255 threads of a polar-cap CTA
// waiting for the few that owned boundary cells).
template <class F>
__device__ __noinline__ double warp_adaptive_gk15(const F& f, double A, double B, double tol_abs, int lane) {
  if (!(B > A)) return 0.0;
  double sa[44], sb[44];
  int sd[44];
  int top = 1;
  sa[0] = A; sb[0] = B; sd[0] = 0;
  // node / weights of this lane: lanes 0..6 left of centre, 7 centre, 8..14 right of centre, 15.. idle
  const int kk = lane < 7 ? lane : (lane == 7 ? 7 : (lane < 15 ? 14 - lane : 7));
  const double xk = lane < 7 ? -c_k15_x[kk] : (lane < 15 ? c_k15_x[kk] : 0.0);
  const double wk = lane < 15 ? c_k15_w[kk] : 0.0;
  const double wg = (lane < 15 && ((kk & 1) || kk == 7)) ? c_g7_w[kk == 7 ? 3 : (kk >> 1)] : 0.0;
  double total = 0.0;
  while (top > 0) {
    --top;
    const double a = sa[top], b = sb[top];
    const int depth = sd[top];
    const double h = 0.5 * (b - a), m = 0.5 * (a + b);
    const double fv = lane < 15 ? f(m + h * xk) : 0.0;
    const double K = warp_sum(wk * fv) * h, G = warp_sum(wg * fv) * h;
    if (fabs(K - G) <= tol_abs || depth >= 40 || top + 2 > 44) total += K;
    else {
      sa[top] = a; sb[top] = m; sd[top] = depth + 1; ++top;
      sa[top] = m; sb[top] = b; sd[top] = depth + 1; ++top;
    }
  }
  return total;
}

// region_cell_area evaluated by a warp (all lanes pass the same arguments and receive the same result)
__device__ __noinline__ double warp_region_cell_area(const Region& g, double l, double u, double pa, double pb, double eps,
                                        double zeta, double tol_abs, int lane) {
  double bp[24];
  int nb = 0;
  bp[nb++] = l; bp[nb++] = u;
  circle_meridian_crossings(g.colat, g.radius, pa, l, u, bp, &nb);
  circle_meridian_crossings(g.colat, g.radius, pb, l, u, bp, &nb);
  double tg[4];
  int nt = 0;
  tg[nt++] = fabs(g.colat - g.radius); tg[nt++] = g.colat + g.radius;
  if (g.hRadius > 0.0) {
    circle_meridian_crossings(g.hColat, g.hRadius, pa - g.hAzi, l, u, bp, &nb);
    circle_meridian_crossings(g.hColat, g.hRadius, pb - g.hAzi, l, u, bp, &nb);
    tg[nt++] = fabs(g.hColat - g.hRadius); tg[nt++] = g.hColat + g.hRadius;
  }
  for (int k = 0; k < nt; ++k) if (tg[k] > l && tg[k] < u) bp[nb++] = tg[k];
  for (int i = 1; i < nb; ++i) {
    const double v = bp[i];
    int j = i - 1;
    while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; --j; }
    bp[j + 1] = v;
  }
  auto f = [&](double th) -> double { return region_width(g, th, pa, pb, 1) * area_element(th, eps, zeta, 0); };
  auto is_tangent = [&](double x) -> bool {
    for (int k = 0; k < nt; ++k) if (fabs(x - tg[k]) < 1.0e-14) return true;
    return false;
  };
  double tot = 0.0;
  for (int i = 0; i + 1 < nb; ++i) {
    const double s0 = bp[i], s1 = bp[i + 1];
    if (!(s1 - s0 > 1.0e-15)) continue;
    const bool sing0 = is_tangent(s0), sing1 = is_tangent(s1);
    if (!sing0 && !sing1) { tot += warp_adaptive_gk15(f, s0, s1, tol_abs, lane); continue; }
    const double mid = (sing0 && sing1) ? 0.5 * (s0 + s1) : (sing0 ? s1 : s0);
    if (sing0) tot += warp_adaptive_gk15([&](double t) -> double { return f(s0 + t * t) * 2.0 * t; }, 0.0, sqrt(mid - s0), tol_abs, lane);
    if (sing1) tot += warp_adaptive_gk15([&](double t) -> double { return f(s1 - t * t) * 2.0 * t; }, 0.0, sqrt(s1 - mid), tol_abs, lane);
  }
  return tot;
}

// geometry of one member's bounding mesh: polar caps use the whole azimuth and start at the pole
// (polar_mesh.pyx:53-61, mesh_tools.pyx:925-941)
struct MeshFrame { double lo, hi, bphi; int polar, invert; };
__device__ __forceinline__ MeshFrame mesh_frame(Region& g) {
  MeshFrame m;
  m.polar = (g.colat - g.radius < 0.0 || g.colat + g.radius > kPi);
  m.invert = 0;
  if (m.polar) {
    m.lo = 0.0;
    if (g.colat + g.radius > kPi) { m.invert = 1; m.hi = kPi - g.colat + g.radius; g.colat = kPi - g.colat; g.hColat = kPi - g.hColat; }
    else m.hi = g.colat + g.radius;
    m.bphi = kPi;
  } else {
    m.lo = g.colat - g.radius; m.hi = g.colat + g.radius;
    m.bphi = asin(sin(g.radius) / sin(g.colat));
  }
  return m;
}

#ifndef XB_MESH_THREADS
#define XB_MESH_THREADS 256
#endif
constexpr int kMeshThreads = XB_MESH_THREADS;
constexpr int kAreaNodes = 1000;          // mesh.pyx:40,88

// One CTA per member instance q.
#ifndef XB_MESH_CTAS
#define XB_MESH_CTAS 2
#endif
__global__ void __launch_bounds__(kMeshThreads, XB_MESH_CTAS) k_spot_mesh(EmbedArgs a) {
  const int q = blockIdx.x, b = q / a.M, tid = threadIdx.x;
  const double R_eq = a.R_eq[b], eps = a.epsilon[b], zeta = a.zeta[b], r_s = a.r_s[b];
  const int m_idx = q - b * a.M;
  Region g;
  g.colat = a.colatitude[q]; g.radius = a.ang_radius[q];
  g.hRadius = a.hole_radius ? a.hole_radius[q] : 0.0;
  g.hColat = a.hole_radius ? a.hole_colatitude[q] : g.colat;
  g.hAzi = a.hole_radius ? a.hole_azimuth[q] : 0.0;
  const int partner = a.partner ? a.partner[m_idx] : -1;
  __shared__ double s_area[kAreaNodes], s_colat[kAreaNodes];
  __shared__ double s_par[128], s_theta[128];
  __shared__ double s_boxA, s_spotA;
  __shared__ int s_n, s_nb;
  constexpr int kBListCap = 4096;
  __shared__ unsigned short s_blist[kBListCap];            // boundary cells of a generic member (second pass)
  if (!(g.radius > 0.0)) {
    if (tid == 0) { a.n_rings[q] = 0; a.n_azi[q] = 0; atomicExch(a.status + b, kUnsupported); }
    return;
  }
  const MeshFrame fr = mesh_frame(g);              // a southern polar cap is meshed mirrored (g is transformed)
  const double TH = g.colat, rho = g.radius;
  const double lo = fr.lo, hi = fr.hi;
  const bool generic = fr.polar || g.hRadius > 0.0 || partner >= 0;
  const double bphi = fr.bphi;                                      // mesh.pyx:59 / pi for a polar cap
  // ---- allocate_cells (mesh_tools.pyx:892-1099): bounding-mesh area vs region area ------------
  const int cells_num = a.member_cells ? a.member_cells[3 * m_idx] : a.num_cells;
  const double cells_min = a.member_cells ? (double)a.member_cells[3 * m_idx + 1] : a.min_sqrt;
  const double cells_max = a.member_cells ? (double)a.member_cells[3 * m_idx + 2] : a.max_sqrt;
  if (tid < 32) {
    const double h = 0.5 * (hi - lo), m = 0.5 * (hi + lo);
    double v = c_gl32_w[tid] * area_element(m + h * c_gl32_x[tid], eps, zeta, 0);
    v = warp_sum(v) * h;
    if (!generic) {
      // spot area: azimuthal width 2a(theta); sqrt substitution at both tangent parallels
      const double cosT = cos(TH), sinT = sin(TH), cos_rho = cos(rho);
      const double L = sqrt(m - lo), hh = 0.5 * L;
      const double t = hh + hh * c_gl32_x[tid];
      double s = c_gl32_w[tid] * 2.0 * t *
                 (2.0 * spot_halfwidth(lo + t * t, cosT, sinT, cos_rho) * area_element(lo + t * t, eps, zeta, 0) +
                  2.0 * spot_halfwidth(hi - t * t, cosT, sinT, cos_rho) * area_element(hi - t * t, eps, zeta, 0));
      s = warp_sum(s) * hh;
      if (tid == 0) {
        s_boxA = 2.0 * bphi * v;
        s_spotA = s;
        double spotA = s;
        if (are_equal(spotA * R_eq * R_eq, 0.0)) spotA = s_boxA / 1000.0;
        double sq = ceil(sqrt((double)cells_num * s_boxA / spotA));
        if (sq < cells_min) sq = cells_min; else if (sq > cells_max) sq = cells_max;
        int n = (int)sq;
        if (n % 2 != 0) n += 1;
        s_n = n;
      }
    } else {
      const double boxA = 2.0 * bphi * v;
      double ownA = warp_region_area(g, lo, hi, eps, zeta, tid);
      if (are_equal(ownA * R_eq * R_eq, 0.0)) ownA = boxA / 1000.0;
      double numCell = (double)cells_num;
      if (partner >= 0) {            // superseding + ceding members share num_cells (:1040-1060)
        const int qp = b * a.M + partner;
        Region gp;
        gp.colat = a.colatitude[qp]; gp.radius = a.ang_radius[qp];
        gp.hRadius = a.hole_radius ? a.hole_radius[qp] : 0.0;
        gp.hColat = a.hole_radius ? a.hole_colatitude[qp] : gp.colat;
        gp.hAzi = a.hole_radius ? a.hole_azimuth[qp] : 0.0;
        const MeshFrame fp = mesh_frame(gp);
        const double hp = 0.5 * (fp.hi - fp.lo), mp = 0.5 * (fp.hi + fp.lo);
        double vp = c_gl32_w[tid] * area_element(mp + hp * c_gl32_x[tid], eps, zeta, 0);
        vp = warp_sum(vp) * hp;
        double partA = warp_region_area(gp, fp.lo, fp.hi, eps, zeta, tid);
        if (are_equal(partA * R_eq * R_eq, 0.0)) partA = 2.0 * fp.bphi * vp / 1000.0;
        const int cede = a.is_cede[m_idx];
        const double superA = cede ? partA : ownA, cedeA = cede ? ownA : partA;
        const double y = cedeA / superA - 1.0;
        const double superN = are_equal(y, 0.0) ? 0.5 * numCell : numCell * ((sqrt(1.0 + y) - 1.0) / y);
        numCell = cede ? (are_equal(y, 0.0) ? 0.5 * numCell : numCell - superN) : superN;
      }
      if (tid == 0) {
        s_boxA = boxA; s_spotA = ownA;
        double sq = ceil(sqrt(numCell * boxA / ownA));
        if (sq < cells_min) sq = cells_min; else if (sq > cells_max) sq = cells_max;
        int n = (int)sq;
        if (n % 2 != 0) n += 1;
        s_n = n;
      }
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n > a.max_rings || n > 128) {
    if (tid == 0) { a.n_rings[q] = 0; a.n_azi[q] = 0; atomicExch(a.status + b, kUnsupported); }
    return;
  }
  const double cellA = s_boxA / (double)(n * n);                    // per R_eq^2
  const double dphi = 2.0 * bphi / (double)n;
  const double eta = cellA / dphi;
  // ---- equal-area parallels: 1000-node area table + Steffen inverse (mesh.pyx:88-131) -------
  for (int i = tid; i < kAreaNodes; i += kMeshThreads) {
    const double c = hi + (lo - hi) * ((double)i / (double)(kAreaNodes - 1));    // linspace(hi, lo, 1000)
    s_colat[i] = (i == kAreaNodes - 1) ? lo : c;
    s_area[i] = integrate_area(s_colat[i], hi, eps, zeta, 0);
  }
  __syncthreads();
  for (int i = tid; i < n - 1; i += kMeshThreads) {
    const double target = (double)n * eta - (double)(i + 1) * eta;
    // the reference accumulates i_eta -= eta; reproduce the same rounding sequence
    double ie = (double)n * eta;
    for (int k = 0; k <= i; ++k) ie -= eta;
    (void)target;
    const int idx = interval_search(s_area, kAreaNodes, ie);
    double val, der;
    steffen_eval(s_area, s_colat, kAreaNodes, idx, ie, &val, &der);
    s_par[i] = val;
  }
  __syncthreads();
  // ---- rings (mesh.pyx:226-254) -----------------------------------------------------------------
  const long ring0 = (long)q * a.max_rings;
  for (int i = tid; i < n; i += kMeshThreads) {
    const double l = (i == 0) ? lo : s_par[i - 1];
    const double u = (i == n - 1) ? hi : s_par[i];
    const double th = integrate_area(l, u, eps, zeta, 1) / eta;
    s_theta[i] = th;
    const double mu = cos(th);
    const double rn = radius_normalised(mu, eps, zeta);
    const double f = f_theta(mu, rn, eps, zeta);
    const double cg = 1.0 / sqrt(1.0 + f * f);
    a.theta[ring0 + i] = fr.invert ? kPi - th : th;                   // polar_mesh.pyx:419-422
    a.radial[ring0 + i] = rn * R_eq;
    a.r_s_over_r[ring0 + i] = r_s / (rn * R_eq);                   // HotRegion.py:913
    a.cos_gamma[ring0 + i] = cg;
    a.maxAlpha[ring0 + i] = kHalfPi + acos(cg);                     // mesh.pyx:253
    const int np_ = a.n_params;
    double* sp_ = a.srcParams + (ring0 + i) * np_;
    sp_[0] = a.temperature[q];                                      // HotRegion.py:965-994
    sp_[1] = effective_gravity(mu, R_eq, zeta, eps);
    // further uniform local variables of a custom hot region, e.g. the beaming parameters of
    // examples_modeling_tutorial/modules/CustomHotRegion_Beaming.py:149-178
    for (int x = 2; x < np_; ++x) sp_[x] = a.extra_params ? a.extra_params[(long)q * (np_ - 2) + (x - 2)] : 0.0;
    if (a.corrParams) {      // elsewhere parameters on the spot's mesh (HotRegion.py:1019-1031, Elsewhere.py:308-331)
      double* cp_ = a.corrParams + (ring0 + i) * np_;
      cp_[0] = a.else_temperature[b];
      cp_[1] = sp_[1];
      for (int x = 2; x < np_; ++x) cp_[x] = 0.0;
    }
  }
  if (tid == 0) { a.n_rings[q] = n; a.n_azi[q] = n; }
  __syncthreads();
  // ---- cells (mesh.pyx:255-352 with superRadius = 0; mirror symmetry in azimuth) -----------------
  const double phi_shift = a.phi_shift[q];
  const int half = n / 2;
  if (generic) {
    // mask boundary points tangent to iso-coordinate curves (mesh.pyx:130-185, polar_mesh.pyx:129-178)
    double sp[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    if (g.hRadius > 0.0) {
      if (g.hColat - g.hRadius < 0.0) {
        sp[0][0] = g.hRadius - g.hColat; sp[0][1] = g.hAzi + kPi;
        if (sp[0][1] > kPi) sp[0][1] -= kTwoPi;
        sp[1][0] = g.hColat + g.hRadius; sp[1][1] = g.hAzi;
      } else if (g.hColat + g.hRadius > kPi) {
        sp[0][0] = g.hColat - g.hRadius; sp[0][1] = g.hAzi;
        sp[1][0] = kTwoPi - g.hColat - g.hRadius; sp[1][1] = g.hAzi + kPi;
        if (sp[1][1] > kPi) sp[1][1] -= kTwoPi;
      } else {
        sp[0][0] = g.hColat - g.hRadius; sp[0][1] = g.hAzi;
        sp[1][0] = acos(cos(g.hColat) / cos(g.hRadius));
        sp[1][1] = g.hAzi + asin(sin(g.hRadius) / sin(g.hColat));
        if (sp[1][1] > kPi) sp[1][1] -= kTwoPi;
        sp[2][0] = g.hColat + g.hRadius; sp[2][1] = g.hAzi;
        sp[3][0] = sp[1][0];
        sp[3][1] = g.hAzi - asin(sin(g.hRadius) / sin(g.hColat));
        if (sp[3][1] < -kPi) sp[3][1] += kTwoPi;
      }
    }
    const bool mirror = are_equal(g.hColat, g.colat) && are_equal(g.hAzi, 0.0);     // mesh.pyx:187-191
    const int j_max = mirror ? half : n;
    if (tid == 0) s_nb = 0;
    __syncthreads();
    for (int t = tid; t < n * j_max; t += kMeshThreads) {
      const int i = t / j_max, j = t - i * j_max;
      const double l = (i == 0) ? lo : s_par[i - 1];
      const double u = (i == n - 1) ? hi : s_par[i];
      double lft = -bphi;
      for (int k = 0; k < j; ++k) lft += dphi;           // the reference accumulates leftmost += delta_phi
      const double right = lft + dphi;
      const double thc = s_theta[i], pc = lft + 0.5 * dphi;
      int p1, p2, p3, p4, p5;
      auto inside = [&](double th, double ph) -> int {
        if (!(eval_psi(th, ph, g.colat) <= g.radius)) return 2;
        return (g.hRadius <= eval_psi(th, ph - g.hAzi, g.hColat)) ? 1 : 0;
      };
      if (fr.polar && i == 0) {                           // polar_mesh.pyx:222-229: both upper corners are the pole
        p1 = (g.colat <= g.radius) ? ((g.hRadius <= g.hColat) ? 1 : 0) : 2;
        p3 = p1;
      } else { p1 = inside(l, lft); p3 = inside(l, right); }
      p2 = inside(u, lft); p4 = inside(u, right); p5 = inside(thc, pc);
      bool integrate = !(p1 == p2 && p2 == p3 && p3 == p4 && p4 == p5);
      if (!integrate && g.hRadius > 0.0)
        for (int k = 0; k < 4; ++k)
          if (l <= sp[k][0] && sp[k][0] <= u && lft <= sp[k][1] && sp[k][1] <= right) integrate = true;
      if (integrate) {                                    // boundary cell: second pass, one warp each
        const int slot = atomicAdd(&s_nb, 1);
        if (slot < kBListCap) { s_blist[slot] = (unsigned short)t; continue; }
      }
      const double area = integrate ? region_cell_area(g, l, u, lft, right, eps, zeta, 1.0e-11 * cellA) * R_eq * R_eq
                                    : ((p5 == 1) ? cellA * R_eq * R_eq : 0.0);
      const long row = (ring0 + i) * a.max_azi;
      a.cellArea[row + j] = area;
      if (mirror) a.cellArea[row + n - 1 - j] = area;
    }
    __syncthreads();
    // second pass: one warp per boundary cell (integrateCell, mesh_tools.pyx:429-473)
    const int lane = tid & 31, warp = tid >> 5;
    const int n_list = min(s_nb, kBListCap);
    for (int w = warp; w < n_list; w += kMeshThreads / 32) {
      const int t = s_blist[w];
      const int i = t / j_max, j = t - i * j_max;
      const double l = (i == 0) ? lo : s_par[i - 1];
      const double u = (i == n - 1) ? hi : s_par[i];
      double lft = -bphi;
      for (int k = 0; k < j; ++k) lft += dphi;
      const double area = warp_region_cell_area(g, l, u, lft, lft + dphi, eps, zeta, 1.0e-11 * cellA, lane) * R_eq * R_eq;
      if (lane == 0) {
        const long row = (ring0 + i) * a.max_azi;
        a.cellArea[row + j] = area;
        if (mirror) a.cellArea[row + n - 1 - j] = area;
      }
    }
  } else
  for (int t = tid; t < n * half; t += kMeshThreads) {
    const int i = t / half, j = t - i * half;
    const double l = (i == 0) ? lo : s_par[i - 1];
    const double u = (i == n - 1) ? hi : s_par[i];
    const double left = -bphi + dphi * (double)j;      // the reference accumulates leftmost += delta_phi
    double lft = -bphi;
    for (int k = 0; k < j; ++k) lft += dphi;
    (void)left;
    const double right = lft + dphi;
    const int p1 = eval_psi(l, lft, TH) <= rho, p2 = eval_psi(u, lft, TH) <= rho;
    const int p3 = eval_psi(l, right, TH) <= rho, p4 = eval_psi(u, right, TH) <= rho;
    const int p5 = eval_psi(s_theta[i], lft + 0.5 * dphi, TH) <= rho;
    double area = 0.0;
    if (!(p1 == p2 && p2 == p3 && p3 == p4 && p4 == p5)) area = spot_cell_area(l, u, lft, right, eps, zeta, TH, rho);
    else if (p5) area = cellA;
    area *= R_eq * R_eq;
    const long row = (ring0 + i) * a.max_azi;
    a.cellArea[row + j] = area;
    a.cellArea[row + n - 1 - j] = area;
  }
  // azimuths: linspace(-bphi + dphi/2, bphi - dphi/2, n) (+ pi if antiphased), HotRegion.py:834-837
  for (int t = tid; t < n * n; t += kMeshThreads) {
    const int i = t / n, j = t - i * n;
    const double start = -bphi + 0.5 * dphi, stop = bphi - 0.5 * dphi;
    const double step = (stop - start) / (double)(n - 1);
    double ph = (j == n - 1) ? stop : start + step * (double)j;     // numpy.linspace
    a.phi[(ring0 + i) * a.max_azi + j] = ph + phi_shift;
  }
  // zero the padding of this instance's cell rows
  for (int t = tid; t < a.max_rings * a.max_azi; t += kMeshThreads) {
    const int i = t / a.max_azi, j = t - i * a.max_azi;
    if (i >= n || j >= n) a.cellArea[ring0 * a.max_azi + t] = 0.0;
  }
}

// ---- rays (rays.pyx:61-247) ---------------------------------------------------------------------
// The reference integrates four integrands with QAG (61-point rule, epsrel 1e-12, <= 100 bisections).  Here
// every pair (deflection, lag) that shares its square root is integrated by one Gauss-Legendre panel; for
// rays whose impact parameter is within a factor of the photon-sphere value the 32-point panel is checked
// against the 16-point one and, where they disagree beyond 1e-13, the interval is bisected adaptively
// towards the near-singular end (rays grazing the photon sphere: the integrand tends to 1/x there and the
// deflection grows like -log(1 - b / b_ph); stars with R -> 3 r_g reach 6 pi and more).
#ifndef XB_FAST_RAYS
#define XB_FAST_RAYS 1
#endif
// p1 (MODE 0) / p0 (MODE 1) arrive as the reciprocal 1 / (R / r_s - 1) resp. 1 / (r_c / r_s - 1), formed once per ray:
// one reciprocal square root and one division per node instead of a square root and three divisions
template <int MODE>
__device__ __forceinline__ void ray_integrand(double x, double p0, double p1, double* fd, double* fl) {
  const double o = 1.0 - x * x;
#if XB_FAST_RAYS
  if (MODE == 0) {            // outDef / outLag (:75-90): p0 = sin^2 alpha, p1 = 1 / (R / r_s - 1)
    const double r = rsqrt(1.0 - p0 + x * x * p0 * (2.0 - x * x - o * o * p1));      // 1 / f
    *fd = x * r;
    *fl = x * r * r / (r + 1.0);                                                     // x / (f + f^2)
  } else {                    // inDef / inLag (:62-73): p0 = 1 / (r_c / r_s - 1)
    const double X = rsqrt(2.0 - x * x - o * o * p0);
    *fd = X;
    *fl = X * X / (x + X);
  }
#else
  if (MODE == 0) {
    const double f = sqrt(1.0 - p0 + x * x * p0 * (2.0 - x * x - o * o * p1));
    *fd = x / f;
    *fl = x / (f + f * f);
  } else {
    const double X = 1.0 / sqrt(2.0 - x * x - o * o * p0);
    *fd = X;
    *fl = X * X / (x + X);
  }
#endif
}
template <int N, int MODE>
__device__ __forceinline__ void ray_panel(double lo, double hi, double p0, double p1, double* d, double* l) {
  const double h = 0.5 * (hi - lo), m = 0.5 * (hi + lo);
  double sd = 0.0, sl = 0.0;
#pragma unroll 4
  for (int k = 0; k < N; ++k) {
    const double x = m + h * (N == 32 ? c_gl32_x[k] : c_gl16_x[k]);
    const double w = (N == 32 ? c_gl32_w[k] : c_gl16_w[k]);
    double fd, fl;
    ray_integrand<MODE>(x, p0, p1, &fd, &fl);
    sd += w * fd;
    sl += w * fl;
  }
  *d = h * sd; *l = h * sl;
}
template <int MODE>
__device__ __noinline__ void ray_adaptive(double lo, double hi, double p0, double p1, double scale_d, double scale_l,
                                          double* d, double* l) {
  constexpr int kDepth = 64;
  double s_lo[kDepth], s_hi[kDepth];
  int sp = 0;
  s_lo[0] = lo; s_hi[0] = hi; sp = 1;
  double sd = 0.0, sl = 0.0;
  const double tol_d = 2.0e-14 * scale_d, tol_l = 2.0e-14 * scale_l, min_w = 1.0e-15 * (hi - lo);
  while (sp > 0) {
    --sp;
    const double a = s_lo[sp], b = s_hi[sp];
    double d32, l32, d16, l16;
    ray_panel<32, MODE>(a, b, p0, p1, &d32, &l32);
    ray_panel<16, MODE>(a, b, p0, p1, &d16, &l16);
    const bool ok = (fabs(d32 - d16) <= tol_d && fabs(l32 - l16) <= tol_l) || (b - a) <= min_w || sp + 2 > kDepth;
    if (ok) { sd += d32; sl += l32; }
    else {
      const double mid = 0.5 * (a + b);
      s_lo[sp] = mid; s_hi[sp] = b; ++sp;      // far half first on the stack: the near-singular end is usually lo
      s_lo[sp] = a; s_hi[sp] = mid; ++sp;
    }
  }
  *d = sd; *l = sl;
}
template <int MODE>
__device__ __forceinline__ void ray_pair(bool guarded, double lo, double hi, double p0, double p1, double* d, double* l) {
  ray_panel<32, MODE>(lo, hi, p0, p1, d, l);
  if (!guarded) return;
  double d16, l16;
  ray_panel<16, MODE>(lo, hi, p0, p1, &d16, &l16);
  if (fabs(*d - d16) <= 1.0e-13 * fabs(*d) && fabs(*l - l16) <= 1.0e-13 * fabs(*l)) return;
  ray_adaptive<MODE>(lo, hi, p0, p1, fabs(*d), fabs(*l), d, l);
}
__device__ __forceinline__ void ray_integrals(double cos_alpha, double r_s, double u, double* defl, double* lag) {
  const double sas = 1.0 - cos_alpha * cos_alpha;
  const double sa = sqrt(sas);
  const double alpha = acos(cos_alpha);
  const double b = sa / (u * sqrt(1.0 - u));
  const double b_ph = 3.0 * sqrt(3.0) / 2.0;
  // one 32-point panel is converged to 1e-13 except within 3e-3 of the photon-sphere impact parameter from below
  // (outgoing branch, R -> 3 r_g only) and 3e-5 from above (turning-point branch): measured against a graded
  // composite rule over u = r_s / R in [0.15, 0.66]; the guard keeps a tenfold margin
  const bool guarded = (b > 0.95 * b_ph) && (b < 1.02 * b_ph);
  if (b <= b_ph) {
    double sd, sl;
    ray_pair<0>(guarded, 0.0, 1.0, sas, 1.0 / (1.0 / u - 1.0), &sd, &sl);
    *defl = sd * 2.0 * b * u;
    *lag = sl * 2.0 * b * b * u * r_s;
  } else {
    double rc, wR;
    if (alpha <= kHalfPi && are_equal(sa, 1.0)) { rc = 1.0 / u; wR = 0.0; }
    else {
      rc = 2.0 * b * cos((atan(sqrt(4.0 * b * b / 27.0 - 1.0)) - kPi) / 3.0) / sqrt(3.0);
      wR = sqrt(1.0 - rc * u);
      if (wR != wR) wR = 0.0;
    }
    double d0, l0;                                   // integrals over [wR, 1]
    const double irc = 1.0 / (rc - 1.0);
    ray_pair<1>(guarded, wR, 1.0, irc, 0.0, &d0, &l0);
    if (alpha <= kHalfPi) {
      *defl = d0 * 2.0 * b / rc;
      *lag = l0 * 2.0 * b * b * r_s / rc;
    } else {
      double d1, l1;                                 // integrals over [0, 1]
      ray_pair<1>(guarded, 0.0, 1.0, irc, 0.0, &d1, &l1);
      *defl = 2.0 * b * (2.0 * d1 - d0) / rc;
      *lag = 2.0 * b * b * r_s * (2.0 * l1 - l0) / rc;
      *lag += 2.0 * r_s * (1.0 / u - rc + log((1.0 / u - 1.0) / (rc - 1.0)));
    }
  }
  *lag /= kC;
}

// grid (ring, q); threads over rays
#ifndef XB_RAYS_CTAS
#define XB_RAYS_CTAS 10     // 48 registers (embed 3.20 -> 2.88 ms)
#endif
__global__ void __launch_bounds__(128, XB_RAYS_CTAS) k_rays(EmbedArgs a) {
  const int i = blockIdx.x, q = blockIdx.y, b = q / a.M;
  const int n = a.n_rings[q];
  if (i >= n) return;
  const long ring = (long)q * a.max_rings + i;
  const int N_R = a.n_rays;
  const double u = a.r_s_over_r[ring], r_s = a.r_s[b];
  double extreme = kHalfPi;                                        // rays.pyx:310-318
  if (u < 2.0 / 3.0) extreme = kPi - asin(sqrt(1.0 - u) * (3.0 * sqrt(3.0) / 2.0) * u);
  else if (!are_equal(u, 2.0 / 3.0)) extreme = asin(sqrt(1.0 - u) * (3.0 * sqrt(3.0) / 2.0) * u);
  double maxAlpha = a.maxAlpha[ring];
  if (maxAlpha >= (1.0 - 1.0e-8) * extreme) maxAlpha = (1.0 - 1.0e-8) * extreme;
  const double inc = (1.0 - cos(maxAlpha)) / ((double)N_R - 1.0);
  // lag calibration (HotRegion.py:880-901)
  const double R_i = a.R_eq[b] / r_s;
  const double Ccal = ((1.0 / u - R_i) + log((1.0 / u - 1.0) / (R_i - 1.0))) * (r_s / kC);
  const double two_pi_f = kTwoPi * a.mode_frequency;
  __shared__ int s_bad;
  __shared__ double s_last[2];
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < N_R; j += blockDim.x) {
    double ca = 1.0 - (double)j * inc, d = 0.0, l = 0.0;
    if (j == 0) ca = 1.0;
    else ray_integrals(ca, r_s, u, &d, &l);
    if (d != d || l != l || d < 0.0) s_bad = 1;                    // the reference patches such rays linearly
    a.cos_alpha[ring * N_R + j] = ca;
    a.deflection[ring * N_R + j] = d;
    if (a.lag) a.lag[ring * N_R + j] = (l - Ccal) * two_pi_f;
    if (j >= N_R - 2) s_last[j - (N_R - 2)] = d;
    if (j == N_R - 1) a.maxDeflection[ring] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0 && (s_bad || !(s_last[1] > s_last[0]))) atomicExch(a.status + b, kNumericalError);   // rays.pyx:327-331
}

// ---- closed mesh of the whole surface (cellmesh/global_mesh.pyx:18-147), for Elsewhere ------------------
// One CTA per parameter vector.  Equal-area cells: n x n, ring colatitudes from the same 1000-node
// area table + Steffen inverse the reference uses, southern rings mirrored (:118-123).
__global__ void __launch_bounds__(kMeshThreads) k_closed_mesh(ClosedMeshArgs a) {
  const int b = blockIdx.x, tid = threadIdx.x, n = a.n, nc = a.n / 2;
  const double R_eq = a.R_eq[b], eps = a.epsilon[b], zeta = a.zeta[b], r_s = a.r_s[b];
  __shared__ double s_area[kAreaNodes], s_colat[kAreaNodes];
  __shared__ double s_par[128], s_theta[128];
  // :55-60  cellArea = 2 * 2 pi * A(0, pi/2) / numCell;  eta = cellArea * n / 2 pi
  const double cellA = 2.0 * kTwoPi * integrate_area(0.0, kHalfPi, eps, zeta, 0) / (double)(n * n);
  const double eta = cellA * (double)n / kTwoPi;
  for (int i = tid; i < kAreaNodes; i += kMeshThreads) {
    const double c = kHalfPi + (0.0 - kHalfPi) * ((double)i / (double)(kAreaNodes - 1));   // linspace(pi/2, 0, 1000)
    s_colat[i] = (i == kAreaNodes - 1) ? 0.0 : c;
    s_area[i] = integrate_area(s_colat[i], kHalfPi, eps, zeta, 0);
  }
  __syncthreads();
  for (int i = tid; i < nc - 1; i += kMeshThreads) {
    double ie = (double)nc * eta;                      // the reference accumulates i_eta -= eta (:83-87)
    for (int k = 0; k <= i; ++k) ie -= eta;
    const int idx = interval_search(s_area, kAreaNodes, ie);
    double val, der;
    steffen_eval(s_area, s_colat, kAreaNodes, idx, ie, &val, &der);
    s_par[i] = val;
  }
  __syncthreads();
  const long ring0 = (long)b * n;
  for (int i = tid; i < nc; i += kMeshThreads) {       // northern rings (:92-116)
    const double l = (i == 0) ? 0.0 : s_par[i - 1];
    const double u = (i == nc - 1) ? kHalfPi : s_par[i];
    const double th = integrate_area(l, u, eps, zeta, 1) / eta;
    const double mu = cos(th);
    const double rn = radius_normalised(mu, eps, zeta);
    const double f = f_theta(mu, rn, eps, zeta);
    const double cg = 1.0 / sqrt(1.0 + f * f);
    const double g = effective_gravity(mu, R_eq, zeta, eps);
    const int im = n - 1 - i;                          // mirrored southern ring (:118-123)
    s_theta[i] = th; s_theta[im] = kPi - th;
    a.radial[ring0 + i] = rn * R_eq; a.radial[ring0 + im] = rn * R_eq;
    a.r_s_over_r[ring0 + i] = r_s / (rn * R_eq); a.r_s_over_r[ring0 + im] = r_s / (rn * R_eq);   // Elsewhere.py:282
    a.cos_gamma[ring0 + i] = cg; a.cos_gamma[ring0 + im] = cg;
    a.maxAlpha[ring0 + i] = kHalfPi + acos(cg); a.maxAlpha[ring0 + im] = kHalfPi + acos(cg);
    a.ring_gravity[ring0 + i] = g; a.ring_gravity[ring0 + im] = g;
  }
  if (tid == 0) { a.cellArea[b] = cellA * R_eq * R_eq; a.n_rings[b] = n; }
  __syncthreads();
  const double dphi = kTwoPi / (double)n;
  const double start = -kPi + 0.5 * dphi, stop = kPi - 0.5 * dphi;
  const double step = (stop - start) / (double)(n - 1);
  const double T = a.temperature[b];
  for (int t = tid; t < n * n; t += kMeshThreads) {
    const int i = t / n, j = t - i * n;
    const long c = (ring0 + i) * n + j;
    a.theta[c] = s_theta[i];
    a.phi[c] = (j == n - 1) ? stop : start + step * (double)j;      // numpy.linspace (:116-117)
    a.srcParams[2 * c] = T;                                          // Elsewhere.py:308-331
    a.srcParams[2 * c + 1] = a.ring_gravity[ring0 + i];
  }
}

cudaError_t launch_embed_closed(ClosedMeshArgs a, cudaStream_t stream) {
  if (a.n < 4 || a.n > 128 || (a.n & 1)) return cudaErrorInvalidValue;
  k_closed_mesh<<<a.B, kMeshThreads, 0, stream>>>(a);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  EmbedArgs r;                                   // rays of the n rings (Elsewhere.py:270-300): no lag needed
  memset(&r, 0, sizeof(r));
  r.B = a.B; r.M = 1; r.max_rings = a.n; r.n_rays = a.n_rays;
  r.R_eq = a.R_eq; r.r_s = a.r_s; r.n_rings = a.n_rings; r.r_s_over_r = a.r_s_over_r; r.maxAlpha = a.maxAlpha;
  r.deflection = a.deflection; r.cos_alpha = a.cos_alpha; r.lag = nullptr; r.maxDeflection = a.maxDeflection;
  r.status = a.status;
  dim3 grid(a.n, a.B);
  k_rays<<<grid, 128, 0, stream>>>(r);
  return cudaGetLastError();
}

cudaError_t launch_embed_spots(EmbedArgs a, cudaStream_t stream) {
  if (a.max_rings > 128 || a.n_params < 2) return cudaErrorInvalidValue;
  k_spot_mesh<<<a.B * a.M, kMeshThreads, 0, stream>>>(a);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  dim3 grid(a.max_rings, a.B * a.M);
  k_rays<<<grid, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace xb
