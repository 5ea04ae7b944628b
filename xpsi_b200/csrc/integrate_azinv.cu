// Pulse integrator exploiting azimuthal invariance of a hot-region member.
//
// Replaces xpsi/cellmesh/integrator_for_azimuthal_invariance.pyx:70-665 (the
// reference's dominant hot loop) together with the atmosphere evaluations it
// calls: hot_BB.pyx:54-98 and hot_Num4D.pyx:248-460.
//
// B200 mapping (not a translation of the OpenMP loop nest) -- two kernels:
//
// k_azinv_geometry   one CTA per (member instance q, ring).  The ring's ray row
//   and its cos(psi) image are staged in shared memory; one thread per (image
//   order, half-leaf) evaluates deflection / cos(alpha) / lag through Steffen
//   pieces rebuilt in registers, then Doppler, redshift and the solid-angle
//   Jacobian.  The reference's sequential visibility state machine is replayed
//   verbatim by one thread per image order.  Output per (q, ring, image): four
//   N_L-vectors PHASE, Z (log10 Z for Num4D), mu*eta, GEOM (0 = leaf dark), plus a
//   ring header (image orders to integrate, (T,g) Lagrange stencil, redshift range).
//
// k_azinv_slab       Num4D only, one CTA per (q, ring): contracts the 4-D table over
//   (log T, log g) -- constant on a ring -- into a (mu, E) slab restricted to the
//   energy rows the ring can reach, written to an L2-resident workspace.
//
// k_azinv_flux<ATM>  one CTA per (q, ring, chunk of 8 energies): thousands of
//   small CTAs, ~42 KB of shared memory each, 5 resident per SM.
//   * Num4D: the (mu, energy-row) tile of the ring's slab this chunk can reach
//     (67 x ~22 doubles) is fetched into shared memory by one TMA tensor copy
//     (cp.async.bulk.tensor.2d, completion on an mbarrier); every intensity is
//     then a 4x4 stencil on-chip with precomputed Lagrange denominators.
//   * the leaf profile and its phase-spline (Akima periodic / Steffen)
//     coefficients are built in shared memory.
//   * accumulation over the ring's cells uses interval moments: for an output
//     phase k the cells falling in one leaf interval m share the cubic, so
//     sum_j A_j S(x_j) = c0 W0 + c1 W1 + c2 W2 + c3 W3 with W_p = sum_j A_j d_j^p.
//     One thread owns phase k and all 8 energies: it walks the cells once,
//     and per touched interval spends 4 FMAs per energy instead of 4 per cell.
//     The reference adds a cell only where the spline is positive
//     (pyx:593); intervals whose cubic is not provably non-negative (Bernstein
//     coefficients) are flagged per energy and evaluated cell by cell.
//   * rings/chunks are combined with fp64 RED atomics into flux[q, E, P].
#include <string.h>
#include <cuda.h>            // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no -lcuda)
#include <cuda/ptx>

#include "common.cuh"
#include "kernels.h"
#include "azinv_shared.cuh"

namespace xb {

constexpr int kNEC = 8;            // energies per flux CTA
constexpr int kGeomThreads = 128;
constexpr int kFluxThreads = 128;


// accretion-disc occultation, Ibragimov & Poutanen (2009) (common_functions.pyx:110-138): 1 = ray not blocked
__device__ __forceinline__ int disk_block(double R_in, double cos_i, double cos_psi, double cos_theta_i,
                                          double r_s_over_r_i, double radius, double sin_alpha,
                                          double theta_i_over_pi) {
  const double cos_psi_d = (cos_i * cos_psi - cos_theta_i) /
                           sqrt(cos_i * cos_i + cos_theta_i * cos_theta_i - 2 * cos_i * cos_theta_i * cos_psi);
  const double sin_psi_d = sqrt(1 - cos_psi_d * cos_psi_d);
  const double r_s_i = r_s_over_r_i * radius;
  const double impact_b = radius * sin_alpha / sqrt(1 - r_s_over_r_i);
  const double r_psi_d = sqrt((r_s_i * r_s_i * (1 - cos_psi_d) * (1 - cos_psi_d)) / (4 * (1 + cos_psi_d) * (1 + cos_psi_d)) +
                              ((impact_b * impact_b) / (sin_psi_d * sin_psi_d))) -
                         (r_s_i * (1 - cos_psi_d)) / (2 * (1 + cos_psi_d));
  return (theta_i_over_pi < 0.5 || (theta_i_over_pi > 0.5 && r_psi_d < R_in)) ? 1 : 0;
}

// order-preserving map double <-> unsigned 64 (for shared-memory atomicMin/Max)
__device__ __forceinline__ unsigned long long order_key(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_order(unsigned long long k) {
  const unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// ===========================================================================
// geometry
// ===========================================================================
template <int ATM>
#ifndef XB_GEOM_CTAS
#define XB_GEOM_CTAS 8      // 64 registers: the kernel waits on its serial visibility replay, more resident CTAs hide it (19.38 -> 18.98 ms integrate)
#endif
__global__ void __launch_bounds__(kGeomThreads, XB_GEOM_CTAS) k_azinv_geometry(AzinvArgs a) {
  const int i = blockIdx.x;                 // ring
  const int q = blockIdx.y;                 // member instance
  const int tid = threadIdx.x;
  const int R_ = a.n_rings_q ? a.n_rings_q[q] : a.n_rings;
  const int A_ = a.n_azi_q ? a.n_azi_q[q] : a.n_azi;
  const long ring = (long)q * a.n_rings + i;           // padded ring index
  if (i >= R_) { if (tid == 0) a.ws_ihdr[ring * kIHdr] = 0; return; }
  const int N_R = a.n_rays, N_L = a.n_leaves;
  const long cell0 = ring * a.n_azi;
  const int leaf_lim = (N_L % 2 == 0) ? N_L / 2 : (N_L + 1) / 2;
  const int n_img_max = a.n_img_max;

  extern __shared__ double smem[];
  __shared__ int s_J, s_jhalf, s_nimg;
  __shared__ int s_inv2[kMaxImages], s_dom[kMaxImages], s_mono[kMaxImages];
  __shared__ unsigned long long s_zlo[kMaxImages], s_zhi[kMaxImages];   // order-preserving keys

  // ---- does the ring radiate? (pyx:286-296); a null mask means cellArea > 0 (HotRegion.py:965)
  if (tid == 0) { s_J = A_; s_jhalf = N_R - 1; }
  if (tid < kMaxImages) { s_inv2[tid] = 0; s_dom[tid] = 0; s_mono[tid] = 0; s_zlo[tid] = ~0ull; s_zhi[tid] = 0ull; }
  __syncthreads();
  for (int j = tid; j < A_; j += kGeomThreads) {
    const bool rad = a.radiates ? (a.radiates[cell0 + j] == 1) : (a.cellArea[cell0 + j] > 0.0);
    if (rad) atomicMin(&s_J, j);
  }
  __syncthreads();
  const int J = s_J;
  if (J >= A_) { if (tid == 0) a.ws_ihdr[ring * kIHdr] = 0; return; }

  double* sp = smem;
  double* s_defl = sp; sp += N_R;
  double* s_calpha = sp; sp += N_R;
  double* s_lag = sp; sp += N_R;
  double* s_cosd = sp; sp += N_R;
  double* s_ptrue = sp; sp += (long)n_img_max * N_L;
  double* s_phase = sp; sp += (long)n_img_max * N_L;
  double* s_gZ = nullptr; double* s_gA = nullptr;       // general integrator: Z and mu*eta are splined too
  if (a.general) { s_gZ = sp; sp += (long)n_img_max * N_L; s_gA = sp; sp += (long)n_img_max * N_L; }
  int* s_vis = reinterpret_cast<int*>(sp);              // [n_img_max][N_L]

  const double* g_defl = a.deflection + ring * N_R;
  const double* g_calpha = a.cos_alpha + ring * N_R;
  const double* g_lag = a.lag + ring * N_R;
  for (int r = tid; r < N_R; r += kGeomThreads) {
    const double d = g_defl[r];
    s_defl[r] = d; s_calpha[r] = g_calpha[r]; s_lag[r] = g_lag[r];
    s_cosd[r] = cos(d);                                   // pyx:216-218
    if (d > kHalfPi) atomicMin(&s_jhalf, r);              // pyx:301-303
  }

  // ---- ring constants (pyx:318-331) ---------------------------------------------
  const double inclination = a.inclination[q];
  const double omega = a.omega[q];
  const double sin_i = sin(inclination), cos_i = cos(inclination);
  const double radius = a.radial[ring];
  const double Grav_z = sqrt(1.0 - a.r_s_over_r[ring]);
  const double cos_gamma = a.cos_gamma[ring];
  const double sin_gamma = sqrt(1.0 - cos_gamma * cos_gamma);
  const double theta_i = a.theta[ring * a.theta_ring_stride];
  const double cos_theta_i = cos(theta_i), sin_theta_i = sin(theta_i);
  const double theta_i_over_pi = theta_i / kPi;
  const double beta = radius * omega * sin_theta_i / (kC * Grav_z);
  const double Lorentz = sqrt(1.0 - beta * beta);
  const double maxDefl = a.maxDeflection[ring];
  int n_img = a.image_order_limit > 0 ? a.image_order_limit : (int)ceil(maxDefl / kPi);
  const bool clamped = n_img > n_img_max;      // more image orders inferred than the kernels keep
  if (clamped) n_img = n_img_max;
  __syncthreads();
  const int jhalf = s_jhalf;        // first ray with deflection > pi/2 (clamped)

  // ---- geometry of every (image, half-leaf) (pyx:342-441) ----------------------------
  View vDefl{s_defl, 1}, vCalpha{s_calpha, 1}, vLag{s_lag, 1};
  View vAltX{s_cosd + jhalf, -1}, vAltY{s_calpha + jhalf, -1};   // pyx:305-308
  const int n_alt = jhalf + 1;
  const double alt_xmin = s_cosd[jhalf];
  for (int t = tid; t < n_img * leaf_lim; t += kGeomThreads) {
    const int I = t / leaf_lim, k = t - I * leaf_lim;
    double* W = leaf_ptr(a.ws_leaf, ring, n_img_max, I, N_L);
    double* wZ = W + N_L; double* wAbb = W + 2 * N_L; double* wGeom = W + 3 * N_L;
    const double leaf_k = a.leaves[k];
    double cos_psi = cos_i * cos_theta_i + sin_i * sin_theta_i * cos(leaf_k);
    double psi = eval_image_deflection(I, acos(cos_psi));
    double sin_psi = sin(psi);
    if (!are_equal(psi, 0.0) && are_equal(sin_psi, 0.0)) {    // pole singularity nudge
      const double _i = cos_i >= 0.0 ? inclination + inclination * 1.0e-6
                                     : inclination - inclination * 1.0e-6;
      cos_psi = cos(_i) * cos_theta_i + sin(_i) * sin_theta_i * cos(leaf_k);
      psi = eval_image_deflection(I, acos(cos_psi));
      sin_psi = sin(psi);
    }
    int visible = 0;
    double cos_alpha = -1.0, sin_alpha = 0.0, mu = 0.0, deriv = -1.0;
    const bool use_alt = (psi <= kHalfPi) && (cos_psi >= alt_xmin);
    int idx = 0;
    if (psi <= maxDefl) {
      if (psi < s_defl[0] || psi > s_defl[N_R - 1]) { atomicExch(&s_dom[I], 1); }   // pyx:362-368
      else {
        if (use_alt) {
          idx = interval_search(vAltX, n_alt, cos_psi);
          steffen_eval(vAltX, vAltY, n_alt, idx, cos_psi, &cos_alpha, &deriv);
        } else {
          idx = interval_search(vDefl, N_R, psi);
          steffen_eval(vDefl, vCalpha, N_R, idx, psi, &cos_alpha, &deriv);
          deriv = exp(log(fabs(deriv)) - log(fabs(sin_psi)));
        }
        sin_alpha = sqrt(1.0 - cos_alpha * cos_alpha);
        mu = cos_alpha * cos_gamma;
        if (!are_equal(psi, 0.0)) {
          const double cos_delta = (cos_i - cos_theta_i * cos_psi) / (sin_theta_i * sin_psi);
          if (theta_i_over_pi < 0.5) mu += sin_alpha * sin_gamma * cos_delta;
          else mu -= sin_alpha * sin_gamma * cos_delta;
        }
        if (mu > 0.0)                   // pyx:390-396
          visible = (a.R_in < 1.0e6) ? disk_block(a.R_in, cos_i, cos_psi, cos_theta_i, a.r_s_over_r[ring], radius,
                                                   sin_alpha, theta_i_over_pi)
                                     : 1;
      }
    }
    double lagv = 0.0;
    if (visible) {
      const int il = use_alt ? interval_search(vDefl, N_R, psi) : idx;
      double dl;
      steffen_eval(vDefl, vLag, N_R, il, psi, &lagv, &dl);
    }
    for (int ks = 0; ks < 2; ++ks) {
      const bool take = (0 < k && k < leaf_lim - 1) || (k == 0 && ks == 0) ||
                        (k == leaf_lim - 1 && (N_L % 2 == 1) && ks == 0) ||
                        (k == leaf_lim - 1 && (N_L % 2 == 0));
      if (!take) continue;
      const int kdx = ks == 0 ? k : N_L - 1 - k;
      s_vis[I * N_L + kdx] = visible;
      if (!visible) { wGeom[kdx] = 0.0; wZ[kdx] = 1.0; wAbb[kdx] = 0.0; continue; }
      double superlum, eta;
      if (!are_equal(psi, 0.0)) {
        const double cos_xi = sin_alpha * sin_i * sin(a.leaves[kdx]) / sin_psi;
        superlum = 1.0 + beta * cos_xi;
        eta = Lorentz / superlum;
      } else { superlum = 1.0; eta = Lorentz; }
      const double Z = eta * Grav_z;
      if (a.general) {                 // integrator.pyx:394-397: raw Z and mu*eta, splined over the leaves later
        s_gZ[I * N_L + kdx] = Z; s_gA[I * N_L + kdx] = mu * eta;
        wGeom[kdx] = mu * fabs(deriv) * Grav_z * eta * eta * eta / superlum;
        s_ptrue[I * N_L + kdx] = a.leaves[kdx] + lagv;
        continue;
      }
      const double zstore = (ATM == 2) ? log10(Z) : Z;
      wZ[kdx] = zstore;
      atomicMin(&s_zlo[I], order_key(zstore));
      atomicMax(&s_zhi[I], order_key(zstore));
      wAbb[kdx] = mu * eta;
      wGeom[kdx] = mu * fabs(deriv) * Grav_z * eta * eta * eta / superlum;
      s_ptrue[I * N_L + kdx] = a.leaves[kdx] + lagv;
    }
  }
  __syncthreads();

  // ---- visibility state machine, one thread per image (pyx:339-563, verbatim order) ----
  if (tid < n_img) {
    double* W = leaf_ptr(a.ws_leaf, ring, n_img_max, tid, N_L);
    double* wZ = W + N_L; double* wAbb = W + 2 * N_L; double* wGeom = W + 3 * N_L;
    double* PH = s_phase + tid * N_L;
    const double* PT = s_ptrue + tid * N_L;
    int* vis = s_vis + tid * N_L;
    int Inv = 2, k0 = 0;
    if (a.general) {
      // integrator.pyx:316-520: the leaf loop runs over ALL leaves (the upper half only drives the
      // flags and re-fills ranges), and Z / mu*eta are stepped linearly across dark ranges
      double* Zs = s_gZ + tid * N_L;
      double* As = s_gA + tid * N_L;
      for (int k = 0; k < N_L; ++k) {
        if (vis[k]) {
          if (k < leaf_lim) {
            PH[k] = PT[k];
            const bool mirror = (0 < k && k < leaf_lim - 1) || (k == leaf_lim - 1 && N_L % 2 == 0);
            if (mirror) PH[N_L - 1 - k] = PT[N_L - 1 - k];
          }
          if (k == 0) {
            PH[N_L - 1] = PH[0] + kTwoPi;
            Zs[N_L - 1] = Zs[0]; As[N_L - 1] = As[0]; wGeom[N_L - 1] = wGeom[0];
            vis[N_L - 1] = 1;
          } else if (Inv == 2) {
            const double step = a.leaves[k] / (double)k;
            const double zs = (Zs[k] - Zs[N_L - k - 1]) / (2.0 * (double)k);
            const double as = (As[k] - As[N_L - k - 1]) / (2.0 * (double)k);
            for (int m = N_L - k; m < N_L; ++m) {
              PH[m] = PH[m - 1] + step; Zs[m] = Zs[m - 1] + zs; As[m] = As[m - 1] + as; wGeom[m] = 0.0;
            }
            PH[0] = PH[N_L - 1] - kTwoPi;
            Zs[0] = Zs[N_L - 1]; As[0] = As[N_L - 1]; wGeom[0] = wGeom[N_L - 1];
            for (int m = 1; m < k; ++m) {
              PH[m] = PH[m - 1] + step; Zs[m] = Zs[m - 1] + zs; As[m] = As[m - 1] + as; wGeom[m] = 0.0;
            }
          } else if (Inv == 1) {
            const double den = (double)(k - k0 + 1);
            double step = (PH[k] - PH[k0 - 1]) / den;
            double zs = (Zs[k] - Zs[k0 - 1]) / den, as = (As[k] - As[k0 - 1]) / den;
            for (int m = k0; m < k; ++m) { PH[m] = PH[m - 1] + step; Zs[m] = Zs[m - 1] + zs; As[m] = As[m - 1] + as; }
            step = (PH[N_L - k0] - PH[N_L - 1 - k]) / den;
            zs = (Zs[N_L - k0] - Zs[N_L - 1 - k]) / den; as = (As[N_L - k0] - As[N_L - 1 - k]) / den;
            for (int m = N_L - k; m < N_L - k0; ++m) {
              PH[m] = PH[m - 1] + step; Zs[m] = Zs[m - 1] + zs; As[m] = As[m - 1] + as;
            }
          }
          Inv = 0;
        } else {
          if (k == 0) { vis[N_L - 1] = 0; wGeom[N_L - 1] = 0.0; }
          if (Inv == 0) {
            if (N_L - k > k) {          // past the half-way leaf the reference's range is empty
              const double den = (double)(N_L - 2 * k + 1);
              const double step = (PH[N_L - k] - PH[k - 1]) / den;
              const double zs = (Zs[N_L - k] - Zs[k - 1]) / den, as = (As[N_L - k] - As[k - 1]) / den;
              for (int m = k; m < N_L - k; ++m) {
                PH[m] = PH[m - 1] + step; Zs[m] = Zs[m - 1] + zs; As[m] = As[m - 1] + as; wGeom[m] = 0.0;
              }
            }
            Inv = 1; k0 = k;
          }
        }
      }
    } else
    for (int k = 0; k < leaf_lim; ++k) {
      if (vis[k]) {
        PH[k] = PT[k];
        const bool mirror = (0 < k && k < leaf_lim - 1) || (k == leaf_lim - 1 && N_L % 2 == 0);
        if (mirror) PH[N_L - 1 - k] = PT[N_L - 1 - k];
        if (k == 0) {
          PH[N_L - 1] = PH[0] + kTwoPi;            // leaf N_L-1 is a copy of leaf 0 (pyx:480-484)
          wZ[N_L - 1] = wZ[0]; wAbb[N_L - 1] = wAbb[0]; wGeom[N_L - 1] = wGeom[0];
          vis[N_L - 1] = 1;
        } else if (Inv == 2) {
          const double step = a.leaves[k] / (double)k;
          for (int m = N_L - k; m < N_L; ++m) PH[m] = PH[m - 1] + step;
          PH[0] = PH[N_L - 1] - kTwoPi;
          for (int m = 1; m < k; ++m) PH[m] = PH[m - 1] + step;
        } else if (Inv == 1) {
          double step = (PH[k] - PH[k0 - 1]) / (double)(k - k0 + 1);
          for (int m = k0; m < k; ++m) PH[m] = PH[m - 1] + step;
          step = (PH[N_L - k0] - PH[N_L - 1 - k]) / (double)(k - k0 + 1);
          for (int m = N_L - k; m < N_L - k0; ++m) PH[m] = PH[m - 1] + step;
        }
        Inv = 0;
      } else {
        if (k == 0) { vis[N_L - 1] = 0; wGeom[N_L - 1] = 0.0; wZ[N_L - 1] = 1.0; wAbb[N_L - 1] = 0.0; }
        if (Inv == 0) {
          const double step = (PH[N_L - k] - PH[k - 1]) / (double)(N_L - 2 * k + 1);
          for (int m = k; m < N_L - k; ++m) PH[m] = PH[m - 1] + step;
          Inv = 1; k0 = k;
        }
      }
    }
    s_inv2[tid] = (Inv == 2);
  } else if (tid == 32 || tid == 33) {
    // While warp 0 replays the visibility state machines, two idle threads prepare the ring's atmosphere
    // headers (their table-axis searches are chains of dependent global loads): thread 32 the hot
    // atmosphere, thread 33 the elsewhere atmosphere of the correction.  They do not depend on the replay;
    // a ring that turns out dark is skipped through ih[0] = 0 whatever its header holds.
    double* dh = a.ws_hdr + ring * kDHdr;
    int* ih = a.ws_ihdr + ring * kIHdr;
    if (tid == 32) {
      const double* VEC = a.srcParams + (a.params_per_cell ? (cell0 + J) : ring) * a.n_params;
      const double kT = kKBOverKeV * pow(10.0, VEC[0]);
      dh[10] = VEC[0]; dh[12] = kT; dh[13] = log10(kT);
      dh[14] = (ATM == 2) ? (kErg / kHKeV) * pow(10.0, 3.0 * VEC[0]) : kErg * kPlanckDistConst;
      if (ATM == 2) {              // (T,g) stencil of the ring, hot_Num4D.pyx:295-409
        View vT{a.hot.logT, 1}, vG{a.hot.logg, 1};
        const int bT = lagrange_base(vT, a.hot.nT, VEC[0]);
        const int bG = lagrange_base(vG, a.hot.ng, VEC[1]);
        double w[4];
        lagrange_weights(vT, bT, VEC[0], w);
        for (int x = 0; x < 4; ++x) dh[2 + x] = w[x];
        lagrange_weights(vG, bG, VEC[1], w);
        for (int x = 0; x < 4; ++x) dh[6 + x] = w[x];
        ih[2] = bT; ih[3] = bG;
        dh[11] = VEC[1];
      }
    } else if (a.corrParams) {     // elsewhere atmosphere at the ring's correction parameters
      const double* CV = a.corrParams + (a.params_per_cell ? (cell0 + J) : ring) * a.n_params;
      const double kTc = kKBOverKeV * pow(10.0, CV[0]);
      dh[kCorrD + 10] = CV[0]; dh[kCorrD + 12] = kTc; dh[kCorrD + 13] = log10(kTc);
      dh[kCorrD + 14] = (a.else_atm_ext == 2) ? kErg / kHKeV * pow(10.0, 3.0 * CV[0]) : kErg * kPlanckDistConst;
      if (a.else_atm_ext == 2) {
        View vT{a.els.logT, 1}, vG{a.els.logg, 1};
        const int bT = lagrange_base(vT, a.els.nT, CV[0]);
        const int bG = lagrange_base(vG, a.els.ng, CV[1]);
        double w[4];
        lagrange_weights(vT, bT, CV[0], w);
        for (int x = 0; x < 4; ++x) dh[kCorrD + 2 + x] = w[x];
        lagrange_weights(vG, bG, CV[1], w);
        for (int x = 0; x < 4; ++x) dh[kCorrD + 6 + x] = w[x];
        ih[6] = bT; ih[7] = bG;
        dh[kCorrD + 11] = CV[1];
      }
    }
  }
  __syncthreads();
  // lagged phases must increase strictly (pyx:556-563): checked by all threads
  for (int t = tid; t < n_img * N_L; t += kGeomThreads) {
    const int I = t / N_L, m = t - I * N_L;
    if (m > 0 && !s_inv2[I] && s_phase[t] <= s_phase[t - 1]) s_mono[I] = 1;
  }
  __syncthreads();
  if (tid == 0) {
    // The reference walks image orders in sequence: the first never-visible order ends the
    // loop (pyx:553-554), so failures on orders it would not have reached are not failures.
    int n = 0, bad = 0;
    while (n < n_img && !s_inv2[n]) { bad |= s_dom[n] | s_mono[n]; ++n; }
    if (n < n_img) bad |= s_dom[n];          // that order's leaf loop did run
    if (bad) { atomicExch(a.status + q, kNumericalError); n = 0; }
    // every kept order was visible and the reference would have gone on to the next one: refuse, never truncate
    else if (clamped && n == n_img) { atomicExch(a.status + q, kUnsupported); n = 0; }
    if (a.work && !bad) {        // algorithmic-work counters for the roofline
      const int reached = (n < n_img) ? n + 1 : n;
      unsigned long long V = 0, Kc = 0;
      for (int I = 0; I < n; ++I)
        for (int l = 0; l < N_L; ++l) V += s_vis[I * N_L + l] ? 1 : 0;
      for (int j = 0; j < A_; ++j)
        Kc += (a.radiates ? (a.radiates[cell0 + j] == 1) : (a.cellArea[cell0 + j] > 0.0)) ? 1 : 0;
      atomicAdd(a.work + 0, (unsigned long long)reached * leaf_lim);
      atomicAdd(a.work + 1, V);
      atomicAdd(a.work + 2, (unsigned long long)n);
      atomicAdd(a.work + 3, Kc * n);
    }
    s_nimg = n;
    int* ih = a.ws_ihdr + ring * kIHdr;
    double* dh = a.ws_hdr + ring * kDHdr;
    ih[0] = n; ih[1] = J;
    double zlo = 1e300, zhi = -1e300;
    for (int I = 0; I < n; ++I) {
      if (s_zhi[I] == 0ull) continue;
      zlo = fmin(zlo, key_order(s_zlo[I])); zhi = fmax(zhi, key_order(s_zhi[I]));
    }
    dh[0] = zlo; dh[1] = zhi;
  }
  __syncthreads();
  n_img = s_nimg;
  for (int t = tid; t < n_img * N_L; t += kGeomThreads) {
    const int I = t / N_L, l = t - I * N_L;
    double* W = leaf_ptr(a.ws_leaf, ring, n_img_max, I, N_L);
    W[l] = s_phase[t];
    if (a.general) { W[N_L + l] = s_gZ[t]; W[2 * N_L + l] = s_gA[t]; }
  }
}

// ===========================================================================
// slab (Num4D)
// ===========================================================================
// energy rows of the table needed for log10(E'/kT) in [vlo, vhi]: [elo, ehi)
template <class P>
__device__ __forceinline__ void row_range(const P& axis, int nE, double vlo, double vhi, int* elo, int* ehi) {
  *elo = lagrange_base(axis, nE, vlo - 1.0e-9);
  *ehi = lagrange_base(axis, nE, vhi + 1.0e-9) + 4;
}

constexpr int kSlabThreads = 256;
constexpr int kGSpan = 8;         // log g nodes the rings of one member may touch in the shared contraction

// All lit rings of member q share log T and their log g stencils fit in kGSpan nodes?  Then the table is
// contracted over T once per member (k_azinv_slab_member) instead of once per ring.  Block-uniform result.
__device__ __forceinline__ bool member_uniform(const AzinvArgs& a, int q, int which, int tid, int nthreads) {
  const int ho = which ? kCorrD : 0;
  const long ring0 = (long)q * a.n_rings;
  int first = -1;
  for (int r = 0; r < a.n_rings; ++r) if (a.ws_ihdr[(ring0 + r) * kIHdr] != 0) { first = r; break; }
  if (first < 0) return false;
  const double T0 = a.ws_hdr[(ring0 + first) * kDHdr + ho + 10];
  const int g0 = a.ws_ihdr[(ring0 + first) * kIHdr + (which ? 7 : 3)];
  int ok = 1;
  for (int r = tid; r < a.n_rings; r += nthreads) {
    if (a.ws_ihdr[(ring0 + r) * kIHdr] == 0) continue;
    if (a.ws_hdr[(ring0 + r) * kDHdr + ho + 10] != T0) ok = 0;
    const int g = a.ws_ihdr[(ring0 + r) * kIHdr + (which ? 7 : 3)];
    if (g - g0 > (kGSpan - 4) / 2 || g0 - g > (kGSpan - 4) / 2) ok = 0;      // => max - min + 4 <= kGSpan
  }
  return __syncthreads_and(ok) != 0;
}

// which = 0: hot atmosphere at the ring's parameters; 1: elsewhere atmosphere at the
// ring's correction parameters (pyx:469-476)
__global__ void __launch_bounds__(kSlabThreads) k_azinv_slab(AzinvArgs a, int which) {
  const int i = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
  const long ring = (long)q * a.n_rings + i;
  int* ih = a.ws_ihdr + ring * kIHdr;
  if (ih[0] == 0) return;
  const double* dh = a.ws_hdr + ring * kDHdr;
  const AtmTable& T = which ? a.els : a.hot;
  const int ho = which ? kCorrD : 0;             // header offset of this atmosphere's block
  __shared__ int s_elo, s_nrows;
  __shared__ double s_wT[4], s_wG[4];
  const double zlo = dh[0], zhi = dh[1], log_kT = dh[ho + 13];
  const int n_chunks = (a.n_energies + kNEC - 1) / kNEC;
  View vE{T.logE, 1};
  if (tid == 0) {
    int elo = 0, ehi = 4;
    if (zlo <= zhi)
      row_range(vE, T.nE, a.log10_energies[0] - zhi - log_kT,
                a.log10_energies[a.n_energies - 1] - zlo - log_kT, &elo, &ehi);
    if (ehi - elo > a.slab_rows_ring) { atomicExch(a.status + q, kUnsupported); ih[0] = 0; ehi = elo; }
    ih[which ? 8 : 4] = elo; ih[which ? 9 : 5] = ehi - elo;
    s_elo = elo; s_nrows = ehi - elo;
    for (int x = 0; x < 4; ++x) { s_wT[x] = dh[ho + 2 + x]; s_wG[x] = dh[ho + 6 + x]; }
  }
  // energy rows each 8-energy chunk of this ring reaches (first table row, count)
  int2* chunk_tab = reinterpret_cast<int2*>(a.ws_chunk) + ((long)which * a.Q * a.n_rings + ring) * n_chunks;
  for (int c = tid; c < n_chunks; c += kSlabThreads) {
    const int e0 = c * kNEC, e1 = min(e0 + kNEC, a.n_energies) - 1;
    int lo = 0, hi = 4;
    if (zlo <= zhi)
      row_range(vE, T.nE, a.log10_energies[e0] - zhi - log_kT, a.log10_energies[e1] - zlo - log_kT, &lo, &hi);
    int2 r; r.x = lo; r.y = hi - lo;
    chunk_tab[c] = r;
  }
  __syncthreads();
  const int elo = s_elo, nrows = s_nrows, nmu = T.nmu;
  if (nrows == 0) return;
  if (member_uniform(a, q, which, tid, kSlabThreads)) return;   // k_azinv_slab_member contracts for all rings at once
  const int bT = ih[which ? 6 : 2], bG = ih[which ? 7 : 3];
  const long S0 = (long)T.ng * nmu * T.nE, S1 = (long)nmu * T.nE, S2 = T.nE;
  double* out = (which ? a.ws_slab2 : a.ws_slab) + ring * (long)nmu * a.slab_rows_ring;
  const int tw = tid & 31, wid = tid >> 5;        // a warp per mu row, lanes along the energy rows
  for (int m = wid; m < nmu; m += kSlabThreads / 32)
  for (int e = tw; e < nrows; e += 32) {
    const double* base = T.buf + (long)bT * S0 + (long)bG * S1 + (long)m * S2 + elo + e;
    double acc = 0.0;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      double inner = 0.0;
#pragma unroll
      for (int y = 0; y < 4; ++y) inner += s_wG[y] * __ldg(base + x * S0 + y * S1);
      acc += s_wT[x] * inner;
    }
    out[(long)m * a.slab_rows_ring + e] = acc;
  }
}

// One CTA per (mu row, member): V[g'] = sum_x wT[x] buf[bT+x][g'][mu][e] for the few log g nodes the member's
// rings touch (4 loads each), then every ring's slab row is sum_y wG_ring[y] V[bG_ring + y] -- the table is
// read once per member instead of once per ring (the rings differ only in log g and in the rows they reach).
constexpr int kSlabMThreads = 192;

constexpr int kSlabMRows = 4;           // mu rows per CTA: the ring headers are staged once for all of them

#ifndef XB_SLABM_CTAS
#define XB_SLABM_CTAS 6     // 56 registers: store-latency bound, more resident CTAs (integrate 18.97 -> 18.39 ms)
#endif
__global__ void __launch_bounds__(kSlabMThreads, XB_SLABM_CTAS) k_azinv_slab_member(AzinvArgs a, int which) {
  const int q = blockIdx.y, tid = threadIdx.x;
  const AtmTable& T = which ? a.els : a.hot;
  const int ho = which ? kCorrD : 0;
  const long ring0 = (long)q * a.n_rings;
  extern __shared__ double smem[];
  double* s_V = smem;                                          // [kGSpan][kSlabMThreads]
  double* s_wG = s_V + kGSpan * kSlabMThreads;                 // [n_rings][4]
  int* s_row = reinterpret_cast<int*>(s_wG + 4 * a.n_rings);   // [n_rings][3]: first row, rows, g base
  __shared__ int s_gmin, s_gmax, s_first, s_nlit;
  if (!member_uniform(a, q, which, tid, kSlabMThreads)) return;
  if (tid == 0) { s_gmin = 1 << 30; s_gmax = -1; s_first = -1; s_nlit = 0; }
  __syncthreads();
  // lit rings, compacted: slot -> (first row, rows, g base), weights and the ring index
  int* s_ring = s_row + 3 * a.n_rings;                         // [n_rings]
  for (int r = tid; r < a.n_rings; r += kSlabMThreads) {
    const int* ih = a.ws_ihdr + (ring0 + r) * kIHdr;
    if (ih[0] == 0 || ih[which ? 9 : 5] <= 0) continue;
    const int slot = atomicAdd(&s_nlit, 1);
    const int g = ih[which ? 7 : 3];
    s_row[3 * slot] = ih[which ? 8 : 4]; s_row[3 * slot + 1] = ih[which ? 9 : 5]; s_row[3 * slot + 2] = g;
    s_ring[slot] = r;
    atomicMin(&s_gmin, g); atomicMax(&s_gmax, g); atomicMax(&s_first, r);
    const double* dh = a.ws_hdr + (ring0 + r) * kDHdr;
    for (int y = 0; y < 4; ++y) s_wG[4 * slot + y] = dh[ho + 6 + y];
  }
  __syncthreads();
  if (s_first < 0) return;
  const int gmin = s_gmin, n_lit = s_nlit;
  const int span = s_gmax - gmin + 4;                          // <= kGSpan by member_uniform
  const int bT = a.ws_ihdr[(ring0 + s_first) * kIHdr + (which ? 6 : 2)];
  const double* dh0 = a.ws_hdr + (ring0 + s_first) * kDHdr;
  const double wT0 = dh0[ho + 2], wT1 = dh0[ho + 3], wT2 = dh0[ho + 4], wT3 = dh0[ho + 5];
  const long S0 = (long)T.ng * T.nmu * T.nE, S1 = (long)T.nmu * T.nE, S2 = T.nE;
  const int rows_stride = a.slab_rows_ring;
  for (int mi = 0; mi < kSlabMRows; ++mi) {
    const int m = blockIdx.x * kSlabMRows + mi;
    if (m >= T.nmu) break;
    double* out_base = (which ? a.ws_slab2 : a.ws_slab) + ring0 * (long)T.nmu * rows_stride + (long)m * rows_stride;
    for (int e0 = 0; e0 < T.nE; e0 += kSlabMThreads) {
      const int e = e0 + tid;
      if (e >= T.nE) continue;
      const double* base = T.buf + (long)bT * S0 + (long)gmin * S1 + (long)m * S2 + e;
      for (int k = 0; k < span; ++k) {                         // own column of s_V: no barrier needed
        const double* b = base + k * S1;
        s_V[k * kSlabMThreads + tid] = wT0 * __ldg(b) + wT1 * __ldg(b + S0) + wT2 * __ldg(b + 2 * S0) + wT3 * __ldg(b + 3 * S0);
      }
      for (int sl = 0; sl < n_lit; ++sl) {
        const int rel = e - s_row[3 * sl];
        if (rel < 0 || rel >= s_row[3 * sl + 1]) continue;
        const double* V = s_V + (s_row[3 * sl + 2] - gmin) * kSlabMThreads + tid;
        const double* w = s_wG + 4 * sl;
        // same contraction as k_azinv_slab with the sums exchanged (equal up to rounding)
        out_base[(long)s_ring[sl] * T.nmu * rows_stride + rel] =
            w[0] * V[0] + w[1] * V[kSlabMThreads] + w[2] * V[2 * kSlabMThreads] + w[3] * V[3 * kSlabMThreads];
      }
    }
  }
}


// Walk a ring's radiating cells (ascending azimuth) for output phase phk and hand every touched leaf
// interval m to the visitor together with the area moments W_p = sum_j A_j d_j^p of the cells inside it
// (d_j = phase + azimuth - PHASE[m], brought into [first,last] by whole turns, pyx:575-583).
template <class Visitor>
__device__ __forceinline__ void walk_cells(double phk, const double* s_PH, int N_L, const double* s_cphi,
                                           const double* s_carea, int n_cells, int* status, Visitor&& visit) {
  const double ph_first = s_PH[0], ph_last = s_PH[N_L - 1];
  // cells ascend in azimuth, so the whole-turn offset only ever steps down by 2 pi along the walk
  int c = 0;
  int m = -1;
  double off = 0.0;
  bool searched = false;
  while (c < n_cells) {
    double x = phk + s_cphi[c] + off;
    if (x > ph_last || x < ph_first) {
      double xr = phk + s_cphi[c];
      if (xr > ph_last) { while (xr > ph_last) xr -= kTwoPi; }
      else if (xr < ph_first) { while (xr < ph_first) xr += kTwoPi; }
      if (xr < ph_first || xr > ph_last) { atomicExch(status, kNumericalError); ++c; continue; }
      off = xr - (phk + s_cphi[c]);
      x = xr;
      searched = false;
    }
    if (!searched) { m = interval_search(s_PH, N_L, x); searched = true; }
    else m = interval_walk(s_PH, N_L, x, m);
    const double xm = s_PH[m], xn = s_PH[m + 1];
    const bool last_iv = (m == N_L - 2);
    double W0 = 0.0, W1 = 0.0, W2 = 0.0, W3 = 0.0;
    const int c_start = c;
    for (;;) {                                // absorb the cells that fall in interval m
      const double d = x - xm;
      const double A = s_carea[c];
      W0 += A;
      double t = A * d; W1 += t;
      t *= d; W2 += t;
      t *= d; W3 += t;
      ++c;
      if (c >= n_cells) break;
      x = phk + s_cphi[c] + off;
      if (!(x >= xm && (x < xn || (last_iv && x <= xn)))) break;   // next interval, or wrap (x > last)
    }
    visit(m, W0, W1, W2, W3, c_start, c);
  }
}

// compact list of a ring's radiating cells (one warp: keeps azimuth order); returns the count to lane 0..31
__device__ __forceinline__ int compact_cells(const AzinvArgs& a, long cell0, int A_, int lane, double* s_cphi,
                                             double* s_carea) {
  int base = 0;
  for (int j0 = 0; j0 < A_; j0 += 32) {
    const int j = j0 + lane;
    bool rad = false;
    double area = 0.0, phi = 0.0;
    if (j < A_) {
      area = a.cellArea[cell0 + j]; phi = a.phi[cell0 + j];
      rad = a.radiates ? (a.radiates[cell0 + j] == 1) : (area > 0.0);
    }
    const unsigned m = __ballot_sync(0xffffffffu, rad);
    if (rad) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      s_cphi[pos] = phi; s_carea[pos] = area;
    }
    base += __popc(m);
  }
  return base;
}

// the compact list alone, for launches without a moments workspace (one warp per ring)
__global__ void __launch_bounds__(32) k_azinv_cells(AzinvArgs a) {
  const int i = blockIdx.x, q = blockIdx.y;
  const long ring = (long)q * a.n_rings + i;
  if (a.ws_ihdr[ring * kIHdr] == 0) return;
  const int A_ = a.n_azi_q ? a.n_azi_q[q] : a.n_azi;
  double* gc = a.ws_cells + ring * 2 * (long)a.n_azi;
  const int n = compact_cells(a, ring * a.n_azi, A_, threadIdx.x, gc, gc + a.n_azi);
  if (threadIdx.x == 0) a.ws_ihdr[ring * kIHdr + 10] = n;
}

// ===========================================================================
// moments: the cell walk is identical for all energy chunks of a ring, so it is done once per
// (ring, image order) here and handed to the flux CTAs through an L2-resident workspace
// ===========================================================================
constexpr int kMomThreads = 128;

__global__ void __launch_bounds__(kMomThreads) k_azinv_moments(AzinvArgs a) {
  const int i = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
  const long ring = (long)q * a.n_rings + i;
  const int n_img = a.ws_ihdr[ring * kIHdr];
  if (n_img == 0) return;
  const int A_ = a.n_azi_q ? a.n_azi_q[q] : a.n_azi;
  const int N_L = a.n_leaves, N_P = a.n_phases, cap = a.mom_cap;
  extern __shared__ double smem[];
  __shared__ int s_ncell;
  double* s_cphi = smem;
  double* s_carea = s_cphi + a.n_azi;
  double* s_PH = s_carea + a.n_azi;
  unsigned char* s_lit = reinterpret_cast<unsigned char*>(s_PH + N_L);      // [N_L] leaf lit, then [N_L] interval live
  unsigned char* s_live = s_lit + N_L;
  if (tid < 32) { const int n = compact_cells(a, ring * a.n_azi, A_, tid, s_cphi, s_carea); if (tid == 0) s_ncell = n; }
  const int k = tid;
  const double phk = (k < N_P) ? a.phases[k] : 0.0;
  __syncthreads();
  // the compact list is also what the flux CTAs of this ring need (slow path only): published once here
  // instead of being rebuilt by each of the ring's energy chunks
  {
    double* gc = a.ws_cells + ring * 2 * (long)a.n_azi;
    const int n = s_ncell;
    for (int j = tid; j < n; j += kMomThreads) { gc[j] = s_cphi[j]; gc[a.n_azi + j] = s_carea[j]; }
    if (tid == 0) a.ws_ihdr[ring * kIHdr + 10] = n;
  }
  for (int I = 0; I < n_img; ++I) {
    __syncthreads();
    const double* W = leaf_ptr(a.ws_leaf, ring, a.n_img_max, I, N_L);
    for (int l = tid; l < N_L; l += kMomThreads) { s_PH[l] = W[l]; s_lit[l] = (W[3 * N_L + l] != 0.0) ? 1 : 0; }
    __syncthreads();
    // An interval whose six surrounding leaves (m-2 .. m+3, periodic) are all dark carries an identically zero
    // spline for either interpolant (Akima's node slopes use two intervals on each side): its entries are
    // dropped, which is most of the work of the higher image orders.
    for (int m = tid; m < N_L - 1; m += kMomThreads) {
      int live = 0;
      for (int dlt = -2; dlt <= 3; ++dlt) {
        int l = m + dlt;
        if (l < 0) l += N_L - 1; else if (l > N_L - 1) l -= N_L - 1;
        live |= s_lit[l];
      }
      s_live[m] = (unsigned char)live;
    }
    __syncthreads();
    if (k >= N_P) continue;
    const long slot = ring * a.n_img_max + I;
    double* mom = a.ws_mom + slot * (long)cap * 4 * N_P;
    int2* meta = a.ws_meta + slot * (long)cap * N_P;
    int cnt = 0;
    walk_cells(phk, s_PH, N_L, s_cphi, s_carea, s_ncell, a.status + q,
               [&](int m, double W0, double W1, double W2, double W3, int c0, int c1) {
                 if (!s_live[m]) return;
                 if (cnt < cap) {
                   double* e = mom + (long)cnt * 4 * N_P + k;
                   e[0] = W0; e[N_P] = W1; e[2 * N_P] = W2; e[3 * N_P] = W3;
                   meta[(long)cnt * N_P + k] = make_int2(m, c0 | (c1 << 16));
                 }
                 ++cnt;
               });
    a.ws_cnt[slot * N_P + k] = (cnt <= cap) ? cnt : -1;      // -1: too many intervals, the flux CTA walks itself
  }
}

// ===========================================================================
// moments as tensor-core operands: the same walk, written as dense tiles.  For one image and one tile of 8
// consecutive output phases the touched leaf intervals form one cyclically contiguous run (cells ascend in
// azimuth, phases ascend), so the tile is a band: step s of the tile is interval (m_start + s) mod (N_L - 1) and
// holds, for each of its 8 phases, the four moments W_0..W_3 (zeros where the phase does not reach the interval).
// One step is then the B operand of one m8n8k4 DMMA: D[energy][phase] += C[energy][p] * W[p][phase].
// ===========================================================================
constexpr int kTilePhases = 8;

__global__ void __launch_bounds__(kMomThreads) k_azinv_tiles(AzinvArgs a) {
  const int i = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
  const long ring = (long)q * a.n_rings + i;
  int* ih = a.ws_ihdr + ring * kIHdr;
  const int n_img = ih[0];
  if (n_img == 0) return;
  const int A_ = a.n_azi_q ? a.n_azi_q[q] : a.n_azi;
  const int N_L = a.n_leaves, N_P = a.n_phases, cap = a.tile_cap, NI = N_L - 1;
  const int n_tiles = (N_P + kTilePhases - 1) / kTilePhases;
  extern __shared__ double smem[];
  __shared__ int s_ncell, s_over;
  double* s_cphi = smem;
  double* s_carea = s_cphi + a.n_azi;
  double* s_PH = s_carea + a.n_azi;
  unsigned char* s_lit = reinterpret_cast<unsigned char*>(s_PH + N_L);
  unsigned char* s_live = s_lit + N_L;
  if (tid < 32) { const int n = compact_cells(a, ring * a.n_azi, A_, tid, s_cphi, s_carea); if (tid == 0) { s_ncell = n; s_over = 0; } }
  const int k = tid, tile = k >> 3, kk = k & 7;
  const bool active = k < N_P, in_tile = tile < n_tiles;
  const double phk = active ? a.phases[k] : 0.0;
  __syncthreads();
  {
    double* gc = a.ws_cells + ring * 2 * (long)a.n_azi;
    const int n = s_ncell;
    for (int j = tid; j < n; j += kMomThreads) { gc[j] = s_cphi[j]; gc[a.n_azi + j] = s_carea[j]; }
    if (tid == 0) ih[10] = n;
  }
  const int n_cells = s_ncell;
  for (int I = 0; I < n_img; ++I) {
    __syncthreads();
    const double* W = leaf_ptr(a.ws_leaf, ring, a.n_img_max, I, N_L);
    for (int l = tid; l < N_L; l += kMomThreads) { s_PH[l] = W[l]; s_lit[l] = (W[3 * N_L + l] != 0.0) ? 1 : 0; }
    __syncthreads();
    for (int m = tid; m < NI; m += kMomThreads) {       // see k_azinv_moments: dark neighbourhoods carry zero cubics
      int live = 0;
      for (int dlt = -2; dlt <= 3; ++dlt) {
        int l = m + dlt;
        if (l < 0) l += NI; else if (l > NI) l -= NI;
        live |= s_lit[l];
      }
      s_live[m] = (unsigned char)live;
    }
    __syncthreads();
    // interval of this phase's first cell; the tile starts at the one of its first phase
    int m_first = 0;
    if (active && n_cells > 0) {
      const double ph_first = s_PH[0], ph_last = s_PH[N_L - 1];
      double xr = phk + s_cphi[0];
      if (xr > ph_last) { while (xr > ph_last) xr -= kTwoPi; }
      else if (xr < ph_first) { while (xr < ph_first) xr += kTwoPi; }
      if (xr >= ph_first && xr <= ph_last) m_first = interval_search(s_PH, N_L, xr);
    }
    const int m_start = __shfl_sync(0xffffffffu, m_first, (tid & 31) & ~(kTilePhases - 1));
    const long slot = ring * a.n_img_max + I;
    const long tbase = (slot * n_tiles + (in_tile ? tile : 0)) * cap;
    double* tp = a.ws_tiles + tbase * 32 + kk * 4;
    int* mp = a.ws_tmeta + tbase * kTilePhases + kk;
    int next = 0, u_off = 0, prev_m = -1;
    bool over = false;
    auto put = [&](int s, double W0, double W1, double W2, double W3, int cells) {
      double2* d = reinterpret_cast<double2*>(tp + (long)s * 32);
      d[0] = make_double2(W0, W1); d[1] = make_double2(W2, W3);
      mp[(long)s * kTilePhases] = cells;
    };
    if (active && in_tile)
      walk_cells(phk, s_PH, N_L, s_cphi, s_carea, n_cells, a.status + q,
                 [&](int m, double W0, double W1, double W2, double W3, int c0, int c1) {
                   if (prev_m < 0) u_off = (m < m_start) ? NI : 0;       // this phase has wrapped, the tile's first one not yet
                   else if (m <= prev_m) u_off += NI;                     // whole-turn wrap of the walk
                   prev_m = m;
                   if (!s_live[m] || over) return;
                   const int st = m + u_off - m_start;
                   if (st < next || st >= cap) { over = true; return; }
                   for (; next < st; ++next) put(next, 0.0, 0.0, 0.0, 0.0, 0);
                   // moments of the Hermite basis on the interval (s = delta / h): what multiplies
                   // (y_m, y_m+1, t_m, t_m+1) in sum_j A_j f(delta_j)
                   const double ih = 1.0 / (s_PH[m + 1] - s_PH[m]);
                   const double u2 = W2 * ih, u3 = W3 * (ih * ih);
                   const double H01 = ih * (3.0 * u2 - 2.0 * u3);
                   put(st, W0 - H01, H01, W1 - 2.0 * u2 + u3, u3 - u2, c0 | (c1 << 16));
                   next = st + 1;
                 });
    int ns = over ? cap + 1 : next;
#pragma unroll
    for (int o = 1; o < kTilePhases; o <<= 1) ns = max(ns, __shfl_xor_sync(0xffffffffu, ns, o));
    if (ns <= cap) ns = min((ns + 3) & ~3, cap);      // whole groups of four steps (zero steps add nothing); cap % 4 == 0
    if (in_tile) {
      if (ns > cap) { if (kk == 0) s_over = 1; }
      else for (; next < ns; ++next) put(next, 0.0, 0.0, 0.0, 0.0, 0);
      if (kk == 0) a.ws_thdr[slot * n_tiles + tile] = make_int2(m_start, ns > cap ? -1 : ns);
    }
  }
  __syncthreads();
  if (tid == 0) {
    ih[11] = s_over;       // 1: some tile of this ring overflowed -> the ring goes to the scalar flux kernel
    if (s_over && a.ws_ovf) {            // ... through the compact list the scalar kernel strides over
      a.ws_ovf[1 + a.Q * a.n_rings + ring] = 1;
      a.ws_ovf[1 + atomicAdd(a.ws_ovf, 1)] = (int)ring;
    }
  }
}

// ===========================================================================
// flux
// ===========================================================================
// shared-memory view of one Num4D atmosphere inside a flux CTA
struct SlabCtx {
  double* axE; double* invden; double* axMu; double* slab;
  int nrows, elo_tab, nE, nmu;            // nrows is even (rows are copied 16 bytes at a time)
  double inv_dE, log_kT;
  const double* mu_invden;            // global, [nmu-3][4]
};

// The slab comes first in dynamic shared memory (TMA destinations want 128-byte alignment); the small axis
// arrays are carved later
__device__ __forceinline__ double* slab_ctx_carve_slab(SlabCtx& c, double* sp, int rows_max, int nmu) {
  c.slab = sp; sp += ((long)nmu * rows_max + 15) & ~15l;
  return sp;
}
__device__ __forceinline__ double* slab_ctx_carve(SlabCtx& c, double* sp, int N_L, int rows_max, int nmu) {
  c.axE = sp; sp += rows_max;                       // rows_max is even: every piece stays 16-byte aligned
  c.invden = sp; sp += 4 * rows_max;
  c.axMu = sp; sp += (nmu + 1) & ~1;
  return sp;
}

// Rows of the ring's slab this chunk reaches: the (mu, energy-row) tile [nmu][rows_max] starting at the chunk's first
// row is fetched by ONE TMA tensor copy (cp.async.bulk.tensor.2d) issued by thread 0 -- columns past the ring's slab
// row are zero-filled by the copy engine and are never read (the chunk's row range already holds every stencil).
template <int NT = kFluxThreads>
__device__ __forceinline__ void slab_ctx_load(SlabCtx& c, const AtmTable& T, const CUtensorMap* tmap, long ring,
                                              int elo_ring, int2 chunk_rows, int rows_max, int tid,
                                              uint64_t* mbar) {
  c.nrows = rows_max; c.elo_tab = chunk_rows.x; c.nE = T.nE; c.nmu = T.nmu;
  // the tile must start on a 16-byte boundary of global memory: an even row of the ring's slab
  if ((c.elo_tab - elo_ring) & 1) --c.elo_tab;
  if (tid == 0) {
    const int coords[2] = {c.elo_tab - elo_ring, (int)(ring * T.nmu)};
    cuda::ptx::cp_async_bulk_tensor(cuda::ptx::space_cluster, cuda::ptx::space_global, c.slab, tmap, coords, mbar);
  }
  for (int m = tid; m < T.nmu; m += NT) c.axMu[m] = T.mu[m];
  // tile rows past the end of the table get an unreachable axis value
  for (int r = tid; r < c.nrows; r += NT) c.axE[r] = (c.elo_tab + r < T.nE) ? T.logE[c.elo_tab + r] : 1.0e300;
  // Lagrange denominators per base row: copied from the table's precomputed list
  const int nv = min(c.nrows, c.nE - c.elo_tab);                     // rows inside the table
  for (int r = tid; r < 4 * (nv - 3); r += NT) c.invden[r] = __ldg(T.E_invden + 4 * c.elo_tab + r);
}

__device__ __forceinline__ void slab_ctx_finish(SlabCtx& c, const AtmTable& T, int tid) {     // after a barrier
  const int nv = min(c.nrows, c.nE - c.elo_tab);                     // rows inside the table
  c.mu_invden = T.mu_invden;
  // mean spacing of the axis segment: first guess of the energy stencil (then walked)
  c.inv_dE = (nv > 1) ? (double)(nv - 1) / (c.axE[nv - 1] - c.axE[0]) : 0.0;
}

// mu stencil of one leaf, kept in registers by the thread that owns the leaf
// clamp_low: beam_opt 3 only -- its mu sweep leaves the reference's stencil state at the top of the axis, so a
// query below the table is walked down and clamped to the first node on every call (hot_Num4D.pyx:301-323)
struct MuStencil { int b; double w[4]; };
__device__ __forceinline__ MuStencil slab_ctx_mu_stencil(const SlabCtx& c, double v, bool clamp_low = false) {
  if (clamp_low && v < c.axMu[0]) v = c.axMu[0];
  MuStencil m;
  m.b = lagrange_base(c.axMu, c.nmu, v);
  const double d0 = v - c.axMu[m.b], d1 = v - c.axMu[m.b + 1], d2 = v - c.axMu[m.b + 2], d3 = v - c.axMu[m.b + 3];
  const double* iv = c.mu_invden + 4 * m.b;
  m.w[0] = d1 * d2 * d3 * __ldg(iv); m.w[1] = d0 * d2 * d3 * __ldg(iv + 1);
  m.w[2] = d0 * d1 * d3 * __ldg(iv + 2); m.w[3] = d0 * d1 * d2 * __ldg(iv + 3);
  return m;
}

// I / T^3 at log10(E'/kT) = v with the leaf's mu stencil: 4x4 (mu, E) stencil on the slab (hot_Num4D.pyx:416-437)
__device__ __forceinline__ double slab_ctx_eval(const SlabCtx& c, double v, const MuStencil& ms) {
  const int j = interval_walk(c.axE, c.nrows, v, (int)((v - c.axE[0]) * c.inv_dE));
  int bE = j - 1;                                                   // base node (App. C.5)
  if (c.elo_tab + bE < 0) bE = -c.elo_tab;
  if (c.elo_tab + bE > c.nE - 4) bE = c.nE - 4 - c.elo_tab;
  const double d0 = v - c.axE[bE], d1 = v - c.axE[bE + 1], d2 = v - c.axE[bE + 2], d3 = v - c.axE[bE + 3];
  const double* iv = c.invden + 4 * bE;
  const double wE0 = d1 * d2 * d3 * iv[0], wE1 = d0 * d2 * d3 * iv[1],
               wE2 = d0 * d1 * d3 * iv[2], wE3 = d0 * d1 * d2 * iv[3];
  const double* row = c.slab + (long)ms.b * c.nrows + bE;
  double sum = 0.0;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const double* r = row + x * c.nrows;
    sum += ms.w[x] * (wE0 * r[0] + wE1 * r[1] + wE2 * r[2] + wE3 * r[3]);
  }
  return sum < 0.0 ? 0.0 : sum;                                     // hot_Num4D.pyx:436-437
}

// the same for an arbitrary mu (beam_opt 3 integrates over mu, hot_wrapper.pyx:173-192)
__device__ __noinline__ double slab_ctx_eval_mu(const SlabCtx& c, double v, double mu) {
  const int j = interval_walk(c.axE, c.nrows, v, (int)((v - c.axE[0]) * c.inv_dE));
  int bE = j - 1;
  if (c.elo_tab + bE < 0) bE = -c.elo_tab;
  if (c.elo_tab + bE > c.nE - 4) bE = c.nE - 4 - c.elo_tab;
  double wE[4], wM[4];
  lagrange_weights(c.axE, bE, v, wE);
  const int bM = lagrange_base(c.axMu, c.nmu, mu);
  lagrange_weights(c.axMu, bM, mu, wM);
  const double* row = c.slab + (long)bM * c.nrows + bE;
  double sum = 0.0;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const double* r = row + x * c.nrows;
    sum += wM[x] * (wE[0] * r[0] + wE[1] * r[1] + wE[2] * r[2] + wE[3] * r[3]);
  }
  return sum < 0.0 ? 0.0 : sum;
}

// beaming of one profile value; zstore is Z (blackbody) or log10 Z (Num4D), logT the ring's log10 T
template <int ATM>
__device__ __noinline__ double profile_beaming(int beam_opt, const SlabCtx& hot, double I_E, double E, double logE,
                                               double zstore, double mu, double kT, double log_kT, double logT,
                                               const double* BV) {
  const double Ep = (ATM == 2) ? E * exp10(-zstore) : E / zstore;
  const double t3 = (ATM == 2) ? pow(10.0, 3.0 * logT) : 1.0;
  const double v = logE - zstore - log_kT;
  return apply_beaming(beam_opt, I_E * t3, Ep, mu, BV, [&](double mu_i) -> double {
           return (ATM == 2) ? slab_ctx_eval_mu(hot, v, mu_i) * t3 : bb_intensity(Ep, kT);
         }) / t3;
}

// a / s for 0 <= a <= s, s > 0: the Akima weight of stage 2.  Bit-identical to the plain quotient.
__device__ __forceinline__ double ratio_scaled(double a, double s) {
  // a zero numerator (a locally straight profile) is what sends the IEEE division out of line: divide s by s instead
  const double q = ((a == 0.0) ? s : a) / s;
  return (a == 0.0) ? 0.0 : q;
}

// ATM: hot atmosphere (1 BB, 2 Num4D).  CORR: elsewhere correction (0 none, 1 BB, 2 Num4D)
// BEAM: 0 = no beaming code at all (keeps the common instantiation free of the call's register pressure)
// CUBIC: 1 = the global C2 phase spline (its solver needs 6 KB of per-thread local memory, kept out of the
// default instantiation)
// NLP: 0, or the compile-time value of n_leaves = n_phases (common grids get instantiations whose shared-memory
// and workspace offsets are immediates instead of per-access integer arithmetic)
template <int ATM, int CORR, int BEAM, int CUBIC, int NLP>
__device__ __forceinline__ void azinv_flux_body(const AzinvArgs& a, const CUtensorMap* tm_hot, const CUtensorMap* tm_els,
                                                const int i, const int chunk, const int q, int* s_mbar_live) {
  const int n_chunks = (a.n_energies + kNEC - 1) / kNEC;
  const int tid = threadIdx.x;
  const long ring = (long)q * a.n_rings + i;
  const int* ih = a.ws_ihdr + ring * kIHdr;
  // chunk handed back by the tensor-core kernel (degenerate Akima node)?  Requested with the headers
  const int redo = (a.ws_tiles && a.ws_redo) ? a.ws_redo[ring * n_chunks + chunk] : 0;
  // everything the set-up needs from the headers is requested at once (one L2 round trip instead of a chain:
  // the loads behind the early exits would otherwise wait for the ones in front of them)
  const int n_img = ih[0];
  const int n_cells = ih[10];
  int2 cr_hot = make_int2(0, 0), cr_els = make_int2(0, 0);
  if (ATM == 2) cr_hot = reinterpret_cast<const int2*>(a.ws_chunk)[ring * n_chunks + chunk];
  if (CORR == 2) cr_els = reinterpret_cast<const int2*>(a.ws_chunk)[((long)a.Q * a.n_rings + ring) * n_chunks + chunk];
  const int elo_hot = (ATM == 2) ? ih[4] : 0, elo_els = (CORR == 2) ? ih[8] : 0;
  if (n_img == 0) return;
  // this ring's tiles fit: k_azinv_flux_mma integrates it -- unless it handed this chunk back (degenerate Akima node)
  if (a.ws_tiles && ih[11] == 0 && !redo) return;
  const double* dh = a.ws_hdr + ring * kDHdr;
  const int N_E = a.n_energies, N_L = NLP ? NLP : a.n_leaves, N_P = NLP ? NLP : a.n_phases;
  const long cell0 = ring * a.n_azi;
  const int e0 = chunk * kNEC;
  const int ne = min(kNEC, N_E - e0);
  const int n_img_max = a.n_img_max;

  extern __shared__ __align__(128) double smem[];
  __shared__ double s_E[kNEC], s_logE[kNEC];
  __shared__ __align__(8) uint64_t s_mbar;                // completion barrier of the slab copies
  constexpr int kLitWords = 8;                    // lit-leaf bits of the current image (used when N_L <= 256)
  __shared__ unsigned s_litmask[kLitWords];
  const double kT = dh[12], log_kT = dh[13], norm = dh[14];
  const double kT_c = dh[kCorrD + 12], log_kT_c = dh[kCorrD + 13], norm_c = dh[kCorrD + 14];
  // layout: the arrays of the accumulation loop first, at offsets that depend on N_L only (their addresses are
  // re-formed inside the loop rather than held in registers), then the TMA tile(s) on a 128-byte boundary
  double* sp = smem;
  SlabCtx hot, els;
  // radiating cells of the ring in azimuth order (azimuths, then areas at + n_azi): published once per ring by
  // k_azinv_moments / k_azinv_cells; only the slow paths below read them.  The address is formed where it is
  // used: no registers across the hot loop
  auto cells = [&]() -> const double* { return a.ws_cells + ring * 2 * (long)a.n_azi; };
  double* s_PH = sp; sp += N_L;
  double* s_aux = sp; sp += N_L;          // 1/h of the leaf intervals
  // cubic pieces of the chunk's energies in two planes of 16-byte pairs: s_lo[e][l] = (y, b), s_hi[e][l] = (c, d).
  // Threads of a warp read consecutive l in the accumulation stage, so each plane is read at a 16-byte stride
  // (one wavefront per 8 lanes); interleaved (y, b, c, d) quads put lanes i and i+4 on the same banks.
  double* s_coef = sp; sp += (long)kNEC * N_L * 4;
  double2* s_lo = reinterpret_cast<double2*>(s_coef);
  double2* s_hi = s_lo + (long)kNEC * N_L;
  unsigned char* s_flag = reinterpret_cast<unsigned char*>(sp);     // [N_L][kNEC]: 1 = cubic may dip below zero
  sp += N_L;                                                        // kNEC = 8 flag bytes = one double per leaf
  sp = smem + (((sp - smem) + 15) & ~15l);
  if (ATM == 2) sp = slab_ctx_carve_slab(hot, sp, a.slab_ne_max, a.hot.nmu);
  if (CORR == 2) sp = slab_ctx_carve_slab(els, sp, a.slab_ne_max, a.els.nmu);
  if (ATM == 2) sp = slab_ctx_carve(hot, sp, N_L, a.slab_ne_max, a.hot.nmu);
  if (CORR == 2) sp = slab_ctx_carve(els, sp, N_L, a.slab_ne_max, a.els.nmu);

  // ---- compact list of the ring's radiating cells (one warp: keeps azimuth order) ------
  if (tid < kNEC) {
    const int e = tid;
    s_E[e] = a.energies[e0 + (e < ne ? e : 0)];
    s_logE[e] = a.log10_energies[e0 + (e < ne ? e : 0)];
  }

  // ---- Num4D: fetch the tile of the ring's slab(s) this chunk reaches (TMA) ------------------------
  if (n_cells == 0) return;                   // (before any copy is in flight)
  // rows needed, counting the one in front when the chunk starts on an odd row of the ring's slab (tile alignment)
  if ((ATM == 2 && cr_hot.y + ((cr_hot.x - elo_hot) & 1) > a.slab_ne_max) ||
      (CORR == 2 && cr_els.y + ((cr_els.x - elo_els) & 1) > a.slab_ne_max)) {
    if (tid == 0) atomicExch(a.status + q, kUnsupported);   // budget too small: refuse, never clamp
    return;
  }
  if ((ATM == 2 || CORR == 2) && tid == 0) {
    // list mode runs the body several times per CTA: the barrier of the previous work item is invalidated first
    if (*s_mbar_live) asm volatile("mbarrier.inval.shared.b64 [%0];" :: "r"((unsigned)__cvta_generic_to_shared(&s_mbar)) : "memory");
    *s_mbar_live = 1;
    cuda::ptx::mbarrier_init(&s_mbar, 1);
    cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
    const unsigned bytes = (unsigned)(((ATM == 2) ? a.hot.nmu : 0) + ((CORR == 2) ? a.els.nmu : 0)) *
                           (unsigned)a.slab_ne_max * (unsigned)sizeof(double);
    cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared,
                                         &s_mbar, bytes);
  }
  if (ATM == 2) {
    hot.log_kT = log_kT;
    slab_ctx_load(hot, a.hot, tm_hot, ring, elo_hot, cr_hot, a.slab_ne_max, tid, &s_mbar);
  }
  if (CORR == 2) {
    els.log_kT = log_kT_c;
    slab_ctx_load(els, a.els, tm_els, ring, elo_els, cr_els, a.slab_ne_max, tid, &s_mbar);
  }
  __syncthreads();
  if (ATM == 2) slab_ctx_finish(hot, a.hot, tid);
  if (CORR == 2) slab_ctx_finish(els, a.els, tid);

  const int interp_kind = a.phase_interp;
  const int k = tid;                         // output phase owned in the accumulation stage
  const double phk = (k < N_P) ? a.phases[k] : 0.0;
  double acc[kNEC];
#pragma unroll
  for (int g = 0; g < kNEC; ++g) acc[g] = 0.0;

  for (int I = 0; I < n_img; ++I) {
    __syncthreads();
    // ---- leaf arrays of this image: the lagged phases go to shared memory (every later stage needs them),
    // redshift / mu*eta / geometry factor stay in the registers of the thread that owns the leaf ---------
    const double* W = leaf_ptr(a.ws_leaf, ring, n_img_max, I, N_L);
    for (int l = tid; l < N_L; l += kFluxThreads) s_PH[l] = W[l];
    // this thread's first leaf (geometry factor, redshift, mu*eta): requested before the barrier
    const bool pre_ok = tid < N_L;
    const double pre_geom = pre_ok ? W[3 * N_L + tid] : 0.0;
    const double pre_zst = pre_ok ? W[N_L + tid] : 1.0, pre_abb = pre_ok ? W[2 * N_L + tid] : 0.0;
    __syncthreads();
    for (int l = tid; l < N_L - 1; l += kFluxThreads) s_aux[l] = 1.0 / (s_PH[l + 1] - s_PH[l]);
    if ((ATM == 2 || CORR == 2) && I == 0)                        // the slab tile has landed (first use: stage 1);
      while (!cuda::ptx::mbarrier_try_wait_parity(&s_mbar, 0u, 2000u)) {}   // suspended waits, not a hot spin
    // ---- (1) leaf profile (pyx:445-478): thread = leaf, the mu stencil is shared by the chunk's energies -----
    for (int lb = 0; lb < N_L; lb += kFluxThreads) {
      const int l = lb + tid;
      const double geom = (lb == 0) ? pre_geom : ((l < N_L) ? W[3 * N_L + l] : 0.0);
      {                                     // which leaves are lit: lets stage 2 skip blocks of dark intervals
        const unsigned lit = __ballot_sync(0xffffffffu, geom != 0.0);
        if ((tid & 31) == 0 && (l >> 5) < kLitWords) s_litmask[l >> 5] = lit;
      }
      if (l >= N_L) continue;
      if (geom == 0.0) {
#pragma unroll
        for (int e = 0; e < kNEC; ++e) s_coef[((long)e * N_L + l) * 2] = 0.0;
        continue;
      }
      const double zst = (lb == 0) ? pre_zst : W[N_L + l];          // Z for a blackbody hot atmosphere, log10 Z for Num4D
      const double abb = (lb == 0) ? pre_abb : W[2 * N_L + l];      // mu * eta
      MuStencil ms_hot, ms_els;
      if (ATM == 2) ms_hot = slab_ctx_mu_stencil(hot, abb, BEAM && a.beam_opt == 3);
      if (CORR == 2) ms_els = slab_ctx_mu_stencil(els, abb);
      const double Zlin = (CORR == 1 && ATM == 2) ? exp10(zst) : zst;
      const double Zlog = (CORR == 2 && ATM != 2) ? log10(zst) : zst;
#pragma unroll
      for (int e = 0; e < kNEC; ++e) {
        double I_E;
        if (ATM == 1) I_E = bb_intensity(s_E[e] / zst, kT);
        else I_E = slab_ctx_eval(hot, s_logE[e] - zst - log_kT, ms_hot);
        if (BEAM)                         // hot_wrapper.pyx:155-199 (options 1-3); kept out of line: rarely used
          I_E = profile_beaming<ATM>(a.beam_opt, hot, I_E, s_E[e], s_logE[e], zst, abb, kT, log_kT, dh[10],
                                     a.srcParams + (a.params_per_cell ? (cell0 + ih[1]) : ring) * a.n_params);
        double corr = 0.0;
        if (CORR == 1) corr = bb_intensity(s_E[e] / Zlin, kT_c) * norm_c;
        else if (CORR == 2) corr = slab_ctx_eval(els, s_logE[e] - Zlog - log_kT_c, ms_els) * norm_c;
        s_coef[((long)e * N_L + l) * 2] = (I_E * norm - corr) * geom;  // pyx:478 (energies past ne repeat the first)
      }
    }
    __syncthreads();
    // the accumulation stage's first loads (entry count, first moment entry of this thread's phase) are issued
    // here, so their L2 latency is spent under stage 2 instead of at the head of stage 3
    const long slot = ring * n_img_max + I;
    int cnt = -1;
    double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
    int2 nm = make_int2(0, 0);
    if (k < N_P && a.ws_mom) {
      const double* mom = a.ws_mom + slot * (long)a.mom_cap * 4 * N_P + k;
      cnt = a.ws_cnt[slot * N_P + k];
      n0 = mom[0]; n1 = mom[N_P]; n2 = mom[2 * N_P]; n3 = mom[3 * N_P];
      nm = a.ws_meta[slot * (long)a.mom_cap * N_P + k];
    }
    // ---- (2) phase-spline coefficients + positivity flags (pyx:566-569) ----------------------
    // thread = (energy, block of consecutive leaf intervals): the five interval slopes of Akima's rule slide
    // along the block in registers, so an interval costs one new slope and one division (the weight
    // alpha of its right node, reused as the left node of the next interval) -- same arithmetic as GSL's
    // akima_calc, a third of the instructions of evaluating every interval from scratch
    {
      constexpr int kBlk = kFluxThreads / kNEC;            // blocks per energy
      const int e = tid / kBlk, blk = tid - e * kBlk;
      const int per = (N_L - 1 + kBlk - 1) / kBlk;
      const int l0 = blk * per, l1 = min(l0 + per, N_L - 1);
      // a block whose leaves l0-2 .. l1+2 (periodic) are all dark holds identically zero cubics for every
      // interpolant (Akima's node slopes reach two intervals to each side): nothing to compute
      bool dark = (!CUBIC && l0 < l1 && N_L <= 32 * kLitWords);
      for (int l = l0 - 2; dark && l <= l1 + 2; ++l) {
        int lw = l;
        if (lw < 0) lw += N_L - 1; else if (lw > N_L - 1) lw -= N_L - 1;
        if ((s_litmask[lw >> 5] >> (lw & 31)) & 1u) dark = false;
      }
      if (dark) {
        for (int l = l0; l < l1; ++l) {
          s_coef[((long)e * N_L + l) * 2 + 1] = 0.0;
          s_hi[(long)e * N_L + l] = make_double2(0.0, 0.0);
          s_flag[l * kNEC + e] = 0;
        }
      } else if (l0 < l1) {          // all kNEC energies (a short last chunk repeats its first energy)
        const View y{s_coef + (long)e * N_L * 2, 2};          // node values: slot 0 of every (y, b) pair of this energy
        auto emit = [&](int l, double b, double c, double d) {
          const double y0 = y[l];
          s_coef[((long)e * N_L + l) * 2 + 1] = b;            // slot 0 (y) is left alone: neighbours read it
          s_hi[(long)e * N_L + l] = make_double2(c, d);
          if (CORR == 0) {
            // Bernstein coefficients of the cubic on [0,h] (end values are the nodes themselves):
            // all >= 0  =>  the spline is >= 0 on the interval.  With the correction active the
            // reference adds every cell whatever its sign (pyx:593), so nothing is flagged.
            const double h = s_PH[l + 1] - s_PH[l];
            const double B1 = y0 + b * h * (1.0 / 3.0);
            const double B2 = y0 + h * ((2.0 / 3.0) * b + c * h * (1.0 / 3.0));
            const double y1 = y[l + 1];
            bool neg = (y0 < 0.0 || y1 < 0.0);
            if (!neg && (B1 < 0.0 || B2 < 0.0)) {
              // Bernstein is only sufficient: look at the cubic's interior critical points before giving up
              // the fast path (b + 2 c t + 3 d t^2 = 0)
              auto below = [&](double t) -> bool {
                return t > 0.0 && t < h && (y0 + t * (b + t * (c + t * d))) < 0.0;
              };
              if (d != 0.0) {
                const double disc = c * c - 3.0 * b * d;
                if (disc >= 0.0) {
                  const double sq = sqrt(disc), i3d = 1.0 / (3.0 * d);
                  neg = below((-c - sq) * i3d) || below((-c + sq) * i3d);
                }
              } else if (c != 0.0) neg = below(-b / (2.0 * c));
            }
            s_flag[l * kNEC + e] = neg ? 1 : 0;
          } else s_flag[l * kNEC + e] = 0;
        };
        if (CUBIC) {
          // global C2 spline (cspline_periodic): one thread per energy solves the cyclic system
          if (blk == 0) {
            double cc[kMaxCubicNodes];
            cspline_second(s_PH, y, N_L, true, cc);
            for (int l = 0; l < N_L - 1; ++l) {
              const double dx = s_PH[l + 1] - s_PH[l], dyv = y[l + 1] - y[l];
              emit(l, dyv / dx - dx * (cc[l + 1] + 2.0 * cc[l]) / 3.0, cc[l], (cc[l + 1] - cc[l]) / (3.0 * dx));
            }
          }
        } else if (interp_kind == kSteffen) {
          for (int l = l0; l < l1; ++l) {
            double b, c, d;
            steffen_coeffs(s_PH, y, N_L, l, &b, &c, &d);
            emit(l, b, c, d);
          }
        } else {
          // Akima (periodic ghosts) with interval slopes taken as dy * (1/h)
          auto slope = [&](int ii) -> double {
            if (ii < 0) ii += N_L - 1; else if (ii > N_L - 2) ii -= N_L - 1;
            return (y[ii + 1] - y[ii]) * s_aux[ii];
          };
          double mm2 = slope(l0 - 2), mm1 = slope(l0 - 1), m0 = slope(l0), mp1 = slope(l0 + 1);
          double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
          double alpha = (NE != 0.0) ? ratio_scaled(fabs(mm1 - mm2), NE) : 0.0;
          for (int l = l0; l < l1; ++l) {
            const double mp2 = slope(l + 2);
            const double NE_next = fabs(mp2 - mp1) + fabs(m0 - mm1);
            const double alpha1 = (NE_next != 0.0) ? ratio_scaled(fabs(m0 - mm1), NE_next) : 0.0;
            double b, c, d;
            if (NE == 0.0) { b = m0; c = 0.0; d = 0.0; }
            else {
              const double tL = (NE_next == 0.0) ? m0 : (1.0 - alpha1) * m0 + alpha1 * mp1;
              const double ih = s_aux[l];
              b = (1.0 - alpha) * mm1 + alpha * m0;
              c = (3.0 * m0 - 2.0 * b - tL) * ih;
              d = (b + tL - 2.0 * m0) * (ih * ih);
            }
            emit(l, b, c, d);
            mm2 = mm1; mm1 = m0; m0 = mp1; mp1 = mp2; NE = NE_next; alpha = alpha1;
          }
        }
      }
    }
    __syncthreads();
    // ---- (3) interval moments x spline coefficients: 4 FMAs per (interval, energy) (pyx:571-596) -----
    if (k < N_P) {
      const double ph_first = s_PH[0], ph_last = s_PH[N_L - 1];
      auto flush = [&](int m, double W0, double W1, double W2, double W3, int c_start, int c_end) {
        static_assert(kNEC == 8, "one flag byte per energy, read as one 64-bit word");
        const unsigned long long fl = *reinterpret_cast<const unsigned long long*>(s_flag + (long)m * kNEC);
        const double2* lop = s_lo + m;
        const double2* hip = s_hi + m;
        if (fl == 0ull) {
#pragma unroll
          for (int g = 0; g < kNEC; ++g) {
            const double2 lo = lop[g * N_L];
            const double2 hi = hip[g * N_L];
            acc[g] = fma(hi.y, W3, fma(hi.x, W2, fma(lo.y, W1, fma(lo.x, W0, acc[g]))));      // 4 DFMA
          }
        } else {
          // some energy's cubic may dip below zero on this interval: the reference adds a cell
          // only where the spline is positive (pyx:593), so go cell by cell for those energies
          const double xm = s_PH[m];
          const double* s_cphi = cells();
          for (int cc = c_start; cc < c_end; ++cc) {
            double xr = phk + s_cphi[cc];
            if (xr > ph_last) { while (xr > ph_last) xr -= kTwoPi; }
            else if (xr < ph_first) { while (xr < ph_first) xr += kTwoPi; }
            const double d = xr - xm;
            const double A = s_cphi[a.n_azi + cc];
#pragma unroll
            for (int g = 0; g < kNEC; ++g) {
              if ((fl >> (8 * g)) & 1ull) {
                const double2 lo = lop[g * N_L], hi = hip[g * N_L];
                const double f = lo.x + d * (lo.y + d * (hi.x + d * hi.y));
                if (f > 0.0) acc[g] += A * f;
              }
            }
          }
#pragma unroll
          for (int g = 0; g < kNEC; ++g) {
            if (!((fl >> (8 * g)) & 1ull)) {
              const double2 lo = lop[g * N_L], hi = hip[g * N_L];
              acc[g] += lo.x * W0 + lo.y * W1 + hi.x * W2 + hi.y * W3;
            }
          }
        }
      };
      if (cnt >= 0) {
        // moments prepared once per (ring, image) by k_azinv_moments; loads are coalesced over k
        const double* mom = a.ws_mom + slot * (long)a.mom_cap * 4 * N_P + k;
        const int2* meta = a.ws_meta + slot * (long)a.mom_cap * N_P + k;
        // software pipeline: entry t+1 is in flight (L2) while entry t is consumed; the two pointers are bumped
        // (no per-entry 64-bit index arithmetic)
        for (int t = 1; t <= cnt; ++t) {
          const double W0 = n0, W1 = n1, W2 = n2, W3 = n3;
          const int2 mt = nm;
          mom += 4 * N_P; meta += N_P;
          if (t < cnt) { n0 = mom[0]; n1 = mom[N_P]; n2 = mom[2 * N_P]; n3 = mom[3 * N_P]; nm = meta[0]; }
          flush(mt.x, W0, W1, W2, W3, mt.y & 0xffff, mt.y >> 16);
        }
      } else {
        walk_cells(phk, s_PH, N_L, cells(), cells() + a.n_azi, n_cells, a.status + q, flush);
      }
    }
  }
  // ---- ring/chunk contribution -> flux[q, e, k] -------------------------------------------------
  if (k < N_P) {
    if (a.flux_part) {                      // deterministic mode: this CTA's slot of the partial sums
      double* part = a.flux_part + (ring * (long)N_E) * N_P;
#pragma unroll
      for (int g = 0; g < kNEC; ++g)
        if (g < ne) part[(long)(e0 + g) * N_P + k] = acc[g];
    } else {
      double* flux_q = a.flux + (long)q * N_E * N_P;
#pragma unroll
      for (int g = 0; g < kNEC; ++g)
        if (g < ne && acc[g] != 0.0) atomicAdd(flux_q + (long)(e0 + g) * N_P + k, acc[g]);
    }
  }
}


// Grid: (ring x energy chunk, member instance) -- or, with a.ovf_list_mode, (energy chunk, G) CTAs that stride over
// the compact list of rings the tensor-core kernel does not cover (tiles overflowed, or a chunk was handed back): the
// list is short (about 1 % of the rings), and a million CTAs that only find "nothing to do" cost 1.3-1.7 ms.
template <int ATM, int CORR, int BEAM, int CUBIC, int NLP>
__global__ void __launch_bounds__(kFluxThreads, (CORR == 2) ? 3 : 5)
k_azinv_flux(AzinvArgs a, const __grid_constant__ CUtensorMap tm_hot, const __grid_constant__ CUtensorMap tm_els) {
  __shared__ int s_mbar_live;
  if (threadIdx.x == 0) s_mbar_live = 0;
  if (a.ovf_list_mode) {
    const int cnt = a.ws_ovf[0];
    for (int k = blockIdx.y; k < cnt; k += gridDim.y) {
      const int ring = a.ws_ovf[1 + k];
      __syncthreads();                       // the previous work item is done with shared memory (and s_mbar_live is visible)
      azinv_flux_body<ATM, CORR, BEAM, CUBIC, NLP>(a, &tm_hot, &tm_els, ring % a.n_rings, (int)blockIdx.x, ring / a.n_rings,
                                                   &s_mbar_live);
    }
  } else {
    const int n_chunks = (a.n_energies + kNEC - 1) / kNEC;
    const int i = blockIdx.x / n_chunks;
    __syncthreads();
    azinv_flux_body<ATM, CORR, BEAM, CUBIC, NLP>(a, &tm_hot, &tm_els, i, (int)(blockIdx.x - i * n_chunks), (int)blockIdx.y,
                                                 &s_mbar_live);
  }
}

// ---------------------------------------------------------------------------------------------------------
// The same integrator with the accumulation stage on the fp64 tensor cores.
//
// Stage 3 of k_azinv_flux spends 4 FMAs per (phase, interval, energy) on 32 bytes of per-thread coefficient
// reads: the 128 B/clk shared-memory path caps it at a quarter of the DFMA rate.  Here a warp owns tiles of 8
// output phases; for every step of a tile (k_azinv_tiles) one m8n8k4 DMMA multiplies the interval's cubic pieces
// for the chunk's 8 energies (A: [energy][p], one conflict-free 256-byte row of shared memory) by the 8 phases'
// moments (B: [p][phase], one coalesced 256-byte line from L2): 1 byte of shared memory per FMA instead of 8,
// one issue slot per 256 FMAs instead of 8.  The accumulators are the D fragments (2 doubles per lane and tile),
// carried across image orders.  Intervals whose cubic may dip below zero for some energy (pyx:593 adds a cell only
// where the spline is positive) enter the DMMA with that energy's row zeroed and are redone cell by cell by the
// warp afterwards.  Rings with a tile that does not fit tile_cap steps are left to k_azinv_flux (ih[11]).
// ---------------------------------------------------------------------------------------------------------
// Node-form (Hermite) coefficient rows of the tensor-core flux kernel: per leaf [8 energies: value y][8 energies: node
// slope t][flag word][pad].  The cubic on interval m is y_m h00 + y_{m+1} h01 + t_m h10 + t_{m+1} h11, so the A operand
// of step m is (y_m, y_{m+1}, t_m, t_{m+1}) per energy -- two rows of 16 doubles instead of one row of 32 monomial
// coefficients: the plane shrinks from 26.4 KB to 16 KB at 100 leaves and a sixth CTA fits on the SM.  Row stride 20
// = 4 (mod 16): rows m and m + 1 of a fragment read fall on complementary banks (2 wavefronts, the minimum for 256
// bytes), and the stage-2 accesses of 4 node blocks x 8 energies per warp cover every bank exactly twice.
constexpr int kRowH = 20;


// Build-time parameters of k_azinv_flux_mma (dev/build_variants.sh + dev/time_step.py time variants side by side on
// the GPU box; profiles/r02f_variants.txt, r02j_variants.txt).  With the node-form rows the kernel needs 31 KB of
// shared memory: 7 CTAs fit an SM within 72 registers, 6 within 80.  Both take the rolled tile loop (the accumulators of
// the tiles a warp is not working on live in local memory) and a stage-1 energy loop unrolled by 2; 6 CTAs with two
// independent DMMA chains per tile (80 registers) are 3 % faster than 7 CTAs with one (72).
#ifndef XB_S1_UNROLL
#define XB_S1_UNROLL 2
#endif
constexpr int kS1Unroll = XB_S1_UNROLL;
#ifndef XB_ROLL_TILES
#define XB_ROLL_TILES 1
#endif
// threads per CTA and resident CTAs per SM (every stage is written for any multiple of 32 threads: stage 1 strides
// over leaves, stage 2 deals kMmaThreads / 8 node blocks per energy, stage 3 deals the 8-phase tiles over the warps)
#ifndef XB_S2_UNROLL
#define XB_S2_UNROLL 1
#endif
constexpr int kS2Unroll = XB_S2_UNROLL;
#ifndef XB_B_LOAD
#define XB_B_LOAD 0          // B rows of the tensor-core stage: 0 plain ld.global, 1 .cg (L2 only), 2 .cs (streaming), 3 .nc
#endif
#ifndef XB_DUAL_ACC
#define XB_DUAL_ACC 1        // even and odd steps of a tile on two independent DMMA chains
#endif
#ifndef XB_FLAG_BITWALK
#define XB_FLAG_BITWALK 1    // flagged-interval correction walks the set bits of the mask instead of every step of the tile
#endif
#ifndef XB_MMA_THREADS
#define XB_MMA_THREADS 128
#endif
#ifndef XB_MMA_CTAS
#define XB_MMA_CTAS 6
#endif
constexpr int kMmaThreads = XB_MMA_THREADS, kMmaCtas = XB_MMA_CTAS;
// exact test behind an inconclusive Bernstein test: does y0 + t (b + t (c + t d)) go below zero inside (0, h)?
// Rare, and kept out of line: the flux kernel is large enough for instruction fetch to show in its stall reasons.
__device__ __noinline__ bool cubic_dips_below(double y0, double b, double c, double d, double h) {
  auto below = [&](double t) -> bool { return t > 0.0 && t < h && (y0 + t * (b + t * (c + t * d))) < 0.0; };
  if (d != 0.0) {
    const double disc = c * c - 3.0 * b * d;
    if (disc >= 0.0) {
      const double sq = sqrt(disc), i3d = 1.0 / (3.0 * d);
      return below((-c - sq) * i3d) || below((-c + sq) * i3d);
    }
    return false;
  }
  if (c != 0.0) return below(-b / (2.0 * c));
  return false;
}

// cell-by-cell correction of one 8-phase tile for the (interval, energy) pairs whose cubic may dip below zero
// (lane = (energy pair eg, eg + 4; phase kk of the tile)); out of line for the same reason.  The monomial pieces are
// rebuilt from the node form where they are needed.
__device__ __noinline__ double2 flagged_tile_correction(const double* s_coef, const double* s_PH, int N_L, int N_P, int tile,
                                                         int2 th, int lane, unsigned long long fm0, unsigned long long fm1,
                                                         unsigned long long fm2, unsigned long long fm3,
                                                         const double* phases, const double* cl, int n_azi, const int* mt) {
  const int kk = lane & 7, eg = lane >> 3, NI = N_L - 1, ns = th.y;
  const int k = tile * kTilePhases + kk;
  const double phk = (k < N_P) ? phases[k] : 0.0;
  const double ph_first = s_PH[0], ph_last = s_PH[N_L - 1];
  mt += kk;
  double s_lo = 0.0, s_hi = 0.0;
#if XB_FLAG_BITWALK
  // the flagged intervals are few: visit the set bits (warp-uniform) instead of testing every step of the tile
#pragma unroll 1
  for (int w = 0; w < 4; ++w) {
    unsigned long long bits = (w == 0) ? fm0 : (w == 1) ? fm1 : (w == 2) ? fm2 : fm3;
    while (bits) {
      const int m = w * 64 + __ffsll((long long)bits) - 1;
      bits &= bits - 1ull;
      int s = m - th.x;
      if (s < 0) s += NI;
      if (s >= ns) continue;
      {
#else
  int m = th.x;
  for (int s = 0; s < ns; ++s) {
    if (((m < 64 ? fm0 : m < 128 ? fm1 : m < 192 ? fm2 : fm3) >> (m & 63)) & 1ull) {     // N_L <= 256 (launcher)
#endif
      const double* row = s_coef + (long)m * kRowH;
      const unsigned long long fw = *reinterpret_cast<const unsigned long long*>(row + 16);
      const bool f_lo = (fw >> (8 * eg)) & 1ull, f_hi = (fw >> (8 * (eg + 4))) & 1ull;
      const int cells = mt[(long)s * kTilePhases];
      const int c_start = cells & 0xffff, c_end = cells >> 16;
      if ((f_lo || f_hi) && k < N_P) {
        const double xm = s_PH[m];
        const double ih = 1.0 / (s_PH[m + 1] - xm);
        double yl = 0.0, bl = 0.0, cl2 = 0.0, dl = 0.0, yh = 0.0, bh = 0.0, ch2 = 0.0, dh2 = 0.0;
        if (f_lo) {
          const double y1 = row[kRowH + eg], t1 = row[kRowH + 8 + eg];
          yl = row[eg]; bl = row[8 + eg];
          const double sl = (y1 - yl) * ih;
          cl2 = (3.0 * sl - 2.0 * bl - t1) * ih; dl = (bl + t1 - 2.0 * sl) * (ih * ih);
        }
        if (f_hi) {
          const double y1 = row[kRowH + eg + 4], t1 = row[kRowH + 8 + eg + 4];
          yh = row[eg + 4]; bh = row[8 + eg + 4];
          const double sl = (y1 - yh) * ih;
          ch2 = (3.0 * sl - 2.0 * bh - t1) * ih; dh2 = (bh + t1 - 2.0 * sl) * (ih * ih);
        }
        for (int cc = c_start; cc < c_end; ++cc) {
          double xr = phk + cl[cc];
          if (xr > ph_last) { while (xr > ph_last) xr -= kTwoPi; }
          else if (xr < ph_first) { while (xr < ph_first) xr += kTwoPi; }
          const double d = xr - xm;
          const double A = cl[n_azi + cc];
          if (f_lo) { const double f = yl + d * (bl + d * (cl2 + d * dl)); if (f < 0.0) s_lo -= A * f; }
          if (f_hi) { const double f = yh + d * (bh + d * (ch2 + d * dh2)); if (f < 0.0) s_hi -= A * f; }
        }
      }
    }
#if XB_FLAG_BITWALK
    }
  }
#else
    if (++m == NI) m = 0;
  }
#endif
  return make_double2(s_lo, s_hi);
}

template <int ATM, int CORR, int NLP>
__global__ void __launch_bounds__(kMmaThreads, (CORR == 2) ? 3 : kMmaCtas)
k_azinv_flux_mma(AzinvArgs a, const __grid_constant__ CUtensorMap tm_hot, const __grid_constant__ CUtensorMap tm_els) {
  const int n_chunks = (a.n_energies + kNEC - 1) / kNEC;
  const int i = blockIdx.x / n_chunks;
  const int chunk = blockIdx.x - i * n_chunks;
  const int q = blockIdx.y;
  const int tid = threadIdx.x;
  const long ring = (long)q * a.n_rings + i;
  const int* ih = a.ws_ihdr + ring * kIHdr;
  const int n_img = ih[0];
  const int n_cells = ih[10];
  const int scalar_ring = ih[11];
  int2 cr_hot = make_int2(0, 0), cr_els = make_int2(0, 0);
  if (ATM == 2) cr_hot = reinterpret_cast<const int2*>(a.ws_chunk)[ring * n_chunks + chunk];
  if (CORR == 2) cr_els = reinterpret_cast<const int2*>(a.ws_chunk)[((long)a.Q * a.n_rings + ring) * n_chunks + chunk];
  const int elo_hot = (ATM == 2) ? ih[4] : 0, elo_els = (CORR == 2) ? ih[8] : 0;
  if (n_img == 0 || scalar_ring) return;
  const double* dh = a.ws_hdr + ring * kDHdr;
  const int N_E = a.n_energies, N_L = NLP ? NLP : a.n_leaves, N_P = NLP ? NLP : a.n_phases;
  const int NI = N_L - 1;
  const int n_tiles = (N_P + kTilePhases - 1) / kTilePhases;
  const int e0 = chunk * kNEC;
  const int ne = min(kNEC, N_E - e0);
  const int n_img_max = a.n_img_max;

  extern __shared__ __align__(128) double smem[];
  __shared__ double s_E[kNEC], s_logE[kNEC];
  __shared__ __align__(8) uint64_t s_mbar;
  constexpr int kLitWords = 8;
  __shared__ unsigned s_litmask[kLitWords];
  __shared__ unsigned long long s_fmask[4];          // bit m: some energy's cubic may dip below zero on interval m
  const double kT = dh[12], log_kT = dh[13], norm = dh[14];
  const double kT_c = dh[kCorrD + 12], log_kT_c = dh[kCorrD + 13], norm_c = dh[kCorrD + 14];
  double* sp = smem;
  SlabCtx hot, els;
  double* s_PH = sp; sp += N_L;
  double* s_aux = sp; sp += N_L;
  double* s_coef = sp; sp += (long)N_L * kRowH;      // [leaf][y x 8 | t x 8 | flag bytes | pad]
  sp = smem + (((sp - smem) + 15) & ~15l);
  if (ATM == 2) sp = slab_ctx_carve_slab(hot, sp, a.slab_ne_max, a.hot.nmu);
  if (CORR == 2) sp = slab_ctx_carve_slab(els, sp, a.slab_ne_max, a.els.nmu);
  if (ATM == 2) sp = slab_ctx_carve(hot, sp, N_L, a.slab_ne_max, a.hot.nmu);
  if (CORR == 2) sp = slab_ctx_carve(els, sp, N_L, a.slab_ne_max, a.els.nmu);

  if (tid < kNEC) {
    const int e = tid;
    s_E[e] = a.energies[e0 + (e < ne ? e : 0)];
    s_logE[e] = a.log10_energies[e0 + (e < ne ? e : 0)];
  }
  if (n_cells == 0) return;
  if ((ATM == 2 && cr_hot.y + ((cr_hot.x - elo_hot) & 1) > a.slab_ne_max) ||
      (CORR == 2 && cr_els.y + ((cr_els.x - elo_els) & 1) > a.slab_ne_max)) {
    if (tid == 0) atomicExch(a.status + q, kUnsupported);
    return;
  }
  if ((ATM == 2 || CORR == 2) && tid == 0) {
    cuda::ptx::mbarrier_init(&s_mbar, 1);
    cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
    const unsigned bytes = (unsigned)(((ATM == 2) ? a.hot.nmu : 0) + ((CORR == 2) ? a.els.nmu : 0)) *
                           (unsigned)a.slab_ne_max * (unsigned)sizeof(double);
    cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared,
                                         &s_mbar, bytes);
  }
  if (ATM == 2) { hot.log_kT = log_kT; slab_ctx_load<kMmaThreads>(hot, a.hot, &tm_hot, ring, elo_hot, cr_hot, a.slab_ne_max, tid, &s_mbar); }
  if (CORR == 2) { els.log_kT = log_kT_c; slab_ctx_load<kMmaThreads>(els, a.els, &tm_els, ring, elo_els, cr_els, a.slab_ne_max, tid, &s_mbar); }
  __syncthreads();
  if (ATM == 2) slab_ctx_finish(hot, a.hot, tid);
  if (CORR == 2) slab_ctx_finish(els, a.els, tid);

  const int interp_kind = a.phase_interp;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kMmaThreads / 32, kTilesPerWarp = (kMmaThreads / kTilePhases + kWarps - 1) / kWarps;
  double acc[kTilesPerWarp][2];            // D fragments: energy lane / 4, phases 2 (lane % 4) + {0, 1} of the tile
#pragma unroll
  for (int t = 0; t < kTilesPerWarp; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }

  for (int I = 0; I < n_img; ++I) {
    __syncthreads();
    const double* W = leaf_ptr(a.ws_leaf, ring, n_img_max, I, N_L);
    for (int l = tid; l < N_L; l += kMmaThreads) s_PH[l] = W[l];
    const bool pre_ok = tid < N_L;
    const double pre_geom = pre_ok ? W[3 * N_L + tid] : 0.0;
    const double pre_zst = pre_ok ? W[N_L + tid] : 1.0, pre_abb = pre_ok ? W[2 * N_L + tid] : 0.0;
    __syncthreads();
    for (int l = tid; l < NI; l += kMmaThreads) s_aux[l] = 1.0 / (s_PH[l + 1] - s_PH[l]);
    if (tid < 4) s_fmask[tid] = 0ull;
    if ((ATM == 2 || CORR == 2) && I == 0)
      while (!cuda::ptx::mbarrier_try_wait_parity(&s_mbar, 0u, 2000u)) {}
    // ---- (1) leaf profile: thread = leaf ------------------------------------------------------------------
    for (int lb = 0; lb < N_L; lb += kMmaThreads) {
      const int l = lb + tid;
      const double geom = (lb == 0) ? pre_geom : ((l < N_L) ? W[3 * N_L + l] : 0.0);
      {
        const unsigned lit = __ballot_sync(0xffffffffu, geom != 0.0);
        if ((tid & 31) == 0 && (l >> 5) < kLitWords) s_litmask[l >> 5] = lit;
      }
      if (l >= N_L) continue;
      double* row = s_coef + (long)l * kRowH;
      if (geom == 0.0) {
#pragma unroll
        for (int e = 0; e < kNEC; e += 2) *reinterpret_cast<double2*>(row + e) = make_double2(0.0, 0.0);
        continue;
      }
      const double zst = (lb == 0) ? pre_zst : W[N_L + l];
      const double abb = (lb == 0) ? pre_abb : W[2 * N_L + l];
      MuStencil ms_hot, ms_els;
      if (ATM == 2) ms_hot = slab_ctx_mu_stencil(hot, abb, false);
      if (CORR == 2) ms_els = slab_ctx_mu_stencil(els, abb);
      const double Zlin = (CORR == 1 && ATM == 2) ? exp10(zst) : zst;
      const double Zlog = (CORR == 2 && ATM != 2) ? log10(zst) : zst;
#pragma unroll kS1Unroll
      for (int e = 0; e < kNEC; ++e) {
        double I_E;
        if (ATM == 1) I_E = bb_intensity(s_E[e] / zst, kT);
        else I_E = slab_ctx_eval(hot, s_logE[e] - zst - log_kT, ms_hot);
        double corr = 0.0;
        if (CORR == 1) corr = bb_intensity(s_E[e] / Zlin, kT_c) * norm_c;
        else if (CORR == 2) corr = slab_ctx_eval(els, s_logE[e] - Zlog - log_kT_c, ms_els) * norm_c;
        row[e] = (I_E * norm - corr) * geom;
      }
    }
    __syncthreads();
    // ---- (2) node slopes + positivity flags: thread = (block of consecutive nodes, energy), energy fastest -----
    // Akima and Steffen are Hermite interpolants: the piece on interval l is fixed by (y_l, y_l+1) and the node
    // slopes (t_l, t_l+1).  GSL's Akima takes a different slope on the two sides of a node in one degenerate case
    // (gsl akima.c: interval l has NE != 0, interval l + 1 has NE == 0 and the two interval slopes differ -- two
    // exactly straight segments meeting at the node); such a chunk is handed to the scalar kernel, which keeps
    // per-interval pieces (ws_redo).
    int irregular = a.force_redo;            // test hook: hand every chunk back (exercises the ws_redo plumbing)
    {
      constexpr int kBlk = kMmaThreads / kNEC;
      const int e = tid & (kNEC - 1), blk = tid / kNEC;
      const int per = (N_L + kBlk - 1) / kBlk;
      const int l0 = blk * per, l1 = min(l0 + per, N_L);           // nodes [l0, l1), intervals [l0, min(l1, NI))
      bool dark = (l0 < l1 && N_L <= 32 * kLitWords);
      for (int l = l0 - 2; dark && l <= l1 + 3; ++l) {
        int lw = l;
        if (lw < 0) lw += NI; else if (lw > NI) lw -= NI;
        if ((s_litmask[lw >> 5] >> (lw & 31)) & 1u) dark = false;
      }
      double* col = s_coef + e;                                    // y at [l * kRowH], t at [l * kRowH + 8]
      auto flag_of = [&](int l) -> unsigned char* { return reinterpret_cast<unsigned char*>(s_coef + (long)l * kRowH + 16) + e; };
      if (dark) {
        for (int l = l0; l < l1; ++l) { col[(long)l * kRowH + 8] = 0.0; *flag_of(l) = 0; }
      } else if (l0 < l1) {
        const View y{col, kRowH};
        // interval l - 1 is complete once the slope of node l is known
        auto finish = [&](int li, double t0, double t1) {
          bool neg = false;
          if (CORR == 0) {
            const double y0 = y[li], y1 = y[li + 1];
            const double h = s_PH[li + 1] - s_PH[li];
            const double B1 = y0 + t0 * h * (1.0 / 3.0);
            const double B2 = y1 - t1 * h * (1.0 / 3.0);
            neg = (y0 < 0.0 || y1 < 0.0);
            if (!neg && (B1 < 0.0 || B2 < 0.0)) {
              const double ih = s_aux[li], sl = (y1 - y0) * ih;
              neg = cubic_dips_below(y0, t0, (3.0 * sl - 2.0 * t0 - t1) * ih, (t0 + t1 - 2.0 * sl) * (ih * ih), h);
            }
          }
          *flag_of(li) = neg ? 1 : 0;
          if (neg) atomicOr(&s_fmask[li >> 6], 1ull << (li & 63));
        };
        const int l_end = min(l1, NI);                              // last node whose slope this thread needs
        if (interp_kind == kSteffen) {
          double t_prev = 0.0;
          for (int l = l0; l <= l_end; ++l) {
            const double t = steffen_node_slope(s_PH, y, N_L, l);
            if (l > l0) finish(l - 1, t_prev, t);
            if (l < l1) col[(long)l * kRowH + 8] = t;
            t_prev = t;
          }
        } else {
          auto slope = [&](int ii) -> double {
            if (ii < 0) ii += NI; else if (ii > N_L - 2) ii -= NI;
            return (y[ii + 1] - y[ii]) * s_aux[ii];
          };
          double mm2 = slope(l0 - 2), mm1 = slope(l0 - 1), m0 = slope(l0);
          double t_prev = 0.0, NE_prev = 0.0;
#pragma unroll kS2Unroll
          for (int l = l0; l <= l_end; ++l) {
            const double mp1 = slope(l + 1);
            const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
            const double alpha = (NE != 0.0) ? ratio_scaled(fabs(mm1 - mm2), NE) : 0.0;
            const double t = (NE == 0.0) ? m0 : (1.0 - alpha) * mm1 + alpha * m0;      // gsl: b of interval l
            if (l > l0) {
              if (NE_prev != 0.0 && NE == 0.0 && mm1 != m0) irregular = 1;
              finish(l - 1, t_prev, t);
            }
            if (l < l1) col[(long)l * kRowH + 8] = t;
            t_prev = t; NE_prev = NE;
            mm2 = mm1; mm1 = m0; m0 = mp1;
          }
        }
      }
    }
    if (__syncthreads_or(irregular)) {
      // nothing of this chunk has left the CTA yet (the accumulators are written after the last image)
      if (tid == 0) {
        a.ws_redo[ring * n_chunks + chunk] = 1;
        if (a.ws_ovf && atomicExch(a.ws_ovf + 1 + a.Q * a.n_rings + ring, 1) == 0)
          a.ws_ovf[1 + atomicAdd(a.ws_ovf, 1)] = (int)ring;
      }
      return;
    }
    // ---- (3) tiles x cubic pieces on the tensor cores ------------------------------------------------------
    // Every interval enters with its true cubic (no masking in the loop); where a cubic may dip below zero the
    // cells on its negative part are taken out again afterwards: sum_j A_j max(f, 0) = sum_j A_j f - sum_{f<0} A_j f.
    const long slot = ring * n_img_max + I;
    const int fe = lane >> 2;                                  // energy row of this lane's A element / D fragment
    const unsigned long long fm0 = s_fmask[0], fm1 = s_fmask[1], fm2 = s_fmask[2], fm3 = s_fmask[3];
    const bool any_flag = (fm0 | fm1 | fm2 | fm3) != 0ull;
    // A fragment of step m: lane (energy fe = lane / 4, p = lane % 4) reads y_m, y_m+1, t_m, t_m+1 for p = 0..3
    const unsigned row_bytes = kRowH * 8u, wrap_bytes = (unsigned)NI * row_bytes;
    const unsigned coef0 = (unsigned)__cvta_generic_to_shared(s_coef) +
                           (unsigned)(((lane & 1) * kRowH + ((lane >> 1) & 1) * 8 + fe) * 8);
    // XB_ROLL_TILES: one copy of the tile body instead of kTilesPerWarp; the accumulators of the tiles a warp is
    // not working on then live in local memory (2 loads + 2 stores per tile and image, L1 hits)
#if XB_ROLL_TILES
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int t = 0; t < kTilesPerWarp; ++t) {
      const int tile = warp + t * kWarps;
      if (tile >= n_tiles) break;
      double acc0 = acc[t][0], acc1 = acc[t][1];
      const int2 th = a.ws_thdr[slot * n_tiles + tile];        // (first interval, steps: a multiple of 4)
      const int ns = th.y;
      const double* bp = a.ws_tiles + ((slot * n_tiles + tile) * (long)a.tile_cap) * 32 + lane;
      unsigned off = coef0 + (unsigned)th.x * row_bytes;
      const unsigned end = coef0 + wrap_bytes;
#if XB_DUAL_ACC
      double accB0 = 0.0, accB1 = 0.0;          // odd steps: a second, independent DMMA chain
#endif
      for (int s0 = 0; s0 < ns; s0 += 4) {
#if XB_B_LOAD == 1
        const double b0 = __ldcg(bp), b1 = __ldcg(bp + 32), b2 = __ldcg(bp + 64), b3 = __ldcg(bp + 96);
#elif XB_B_LOAD == 2
        const double b0 = __ldcs(bp), b1 = __ldcs(bp + 32), b2 = __ldcs(bp + 64), b3 = __ldcs(bp + 96);
#elif XB_B_LOAD == 3
        const double b0 = __ldg(bp), b1 = __ldg(bp + 32), b2 = __ldg(bp + 64), b3 = __ldg(bp + 96);
#else
        const double b0 = bp[0], b1 = bp[32], b2 = bp[64], b3 = bp[96];
#endif
        bp += 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double av;
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(av) : "r"(off));
          const double bv = (j == 0) ? b0 : (j == 1) ? b1 : (j == 2) ? b2 : b3;
#if XB_DUAL_ACC
          if (j & 1)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(accB0), "+d"(accB1) : "d"(av), "d"(bv));
          else
#endif
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc0), "+d"(acc1) : "d"(av), "d"(bv));
          off += row_bytes;
          if (off >= end) off -= wrap_bytes;
        }
      }
#if XB_DUAL_ACC
      acc0 += accB0; acc1 += accB1;
#endif
      if (any_flag) {
        const double2 corr2 = flagged_tile_correction(
            s_coef, s_PH, N_L, N_P, tile, th, lane, fm0, fm1, fm2, fm3, a.phases, a.ws_cells + ring * 2 * (long)a.n_azi,
            a.n_azi, a.ws_tmeta + ((slot * n_tiles + tile) * (long)a.tile_cap) * kTilePhases);
        const double s_lo = corr2.x, s_hi = corr2.y;
        // hand the corrections to the lanes that own D[energy][phase]: energy fe = lane / 4, phases 2 (lane % 4) + {0, 1}
        const int src = (fe & 3) * 8 + 2 * (lane & 3);
        const double a0 = __shfl_sync(0xffffffffu, s_lo, src), a1 = __shfl_sync(0xffffffffu, s_lo, src + 1);
        const double b0 = __shfl_sync(0xffffffffu, s_hi, src), b1 = __shfl_sync(0xffffffffu, s_hi, src + 1);
        acc0 += (fe < 4) ? a0 : b0;
        acc1 += (fe < 4) ? a1 : b1;
      }
      acc[t][0] = acc0; acc[t][1] = acc1;
    }
  }
  // ---- ring/chunk contribution -> flux[q, e, k] (RED) or its slot of the partial sums (deterministic mode) ----
  {
    const int fe = lane >> 2;
    double* out = a.flux_part ? a.flux_part + (ring * (long)N_E) * N_P : a.flux + (long)q * N_E * N_P;
#pragma unroll
    for (int t = 0; t < kTilesPerWarp; ++t) {
      const int tile = warp + t * kWarps;
      if (tile >= n_tiles || fe >= ne) continue;
      const int k0 = tile * kTilePhases + 2 * (lane & 3);
      double* o = out + (long)(e0 + fe) * N_P + k0;
      if (a.flux_part) {
        if (k0 < N_P) o[0] = acc[t][0];
        if (k0 + 1 < N_P) o[1] = acc[t][1];
      } else {
        if (k0 < N_P && acc[t][0] != 0.0) atomicAdd(o, acc[t][0]);
        if (k0 + 1 < N_P && acc[t][1] != 0.0) atomicAdd(o + 1, acc[t][1]);
      }
    }
  }
}

// deterministic mode: flux[q, e, k] = sum over the member's lit rings, in ring order, of the partial sums
__global__ void __launch_bounds__(128) k_azinv_reduce_rings(AzinvArgs a) {
  const int e = blockIdx.x, q = blockIdx.y, k = threadIdx.x;
  if (k >= a.n_phases) return;
  const int R_ = a.n_rings_q ? a.n_rings_q[q] : a.n_rings;
  double sum = 0.0;
  for (int r = 0; r < R_; ++r) {
    const long ring = (long)q * a.n_rings + r;
    const int* ih = a.ws_ihdr + ring * kIHdr;
    if (ih[0] == 0 || ih[10] == 0) continue;
    sum += a.flux_part[(ring * (long)a.n_energies + e) * a.n_phases + k];
  }
  a.flux[((long)q * a.n_energies + e) * a.n_phases + k] = sum;
}

// flux[q, e, k] /= E_e keV   (pyx:610-612)
__global__ void k_scale_flux(double* flux, const double* energies, int Q, int N_E, int N_P) {
  const long n = (long)Q * N_E * N_P;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    const int e = (int)((t / N_P) % N_E);
    flux[t] = flux[t] / (energies[e] * kKeV);
  }
}

static size_t geom_smem_bytes(const AzinvArgs& a) {
  return (4ul * a.n_rays + (a.general ? 4ul : 2ul) * a.n_img_max * a.n_leaves) * sizeof(double) +
         (size_t)a.n_img_max * a.n_leaves * sizeof(int);
}

cudaError_t launch_azinv_geometry(const AzinvArgs& a, cudaStream_t stream) {
  const size_t gsm = geom_smem_bytes(a);
  if (gsm > 227 * 1024) return cudaErrorInvalidValue;
  dim3 ggrid(a.n_rings, a.Q);
  cudaError_t err;
  if (a.hot_atm_ext == 1) {
    if ((err = cudaFuncSetAttribute(k_azinv_geometry<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm)) != cudaSuccess) return err;
    k_azinv_geometry<1><<<ggrid, kGeomThreads, gsm, stream>>>(a);
  } else {
    if ((err = cudaFuncSetAttribute(k_azinv_geometry<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm)) != cudaSuccess) return err;
    k_azinv_geometry<2><<<ggrid, kGeomThreads, gsm, stream>>>(a);
  }
  return cudaGetLastError();
}

static size_t flux_smem_bytes(const AzinvArgs& a, int atm, int corr) {
  size_t d = 2ul * a.n_leaves + (size_t)kNEC * a.n_leaves * 4 + a.n_leaves;     // phases, 1/h, cubic pieces, flags
  d = (d + 15) & ~15ul;
  if (atm == 2) d += 5ul * a.slab_ne_max + ((a.hot.nmu + 1) & ~1) + (((size_t)a.hot.nmu * a.slab_ne_max + 15) & ~15ul);
  if (corr == 2) d += 5ul * a.slab_ne_max + ((a.els.nmu + 1) & ~1) + (((size_t)a.els.nmu * a.slab_ne_max + 15) & ~15ul);
  return d * sizeof(double);
}

// Doppler spread of log10 Z over one ring allowed for when sizing buffers:
// log10((1+b)/(1-b)) at |beta| = 0.23 (700 Hz, 16 km).  Rings beyond it are refused
// (status 3), never clamped.
constexpr double kDopplerDex = 0.2;

void azinv_moment_sizes(const AzinvArgs& a, size_t* mom_doubles, size_t* meta_int2, size_t* cnt_ints) {
  const size_t slots = (size_t)a.Q * a.n_rings * a.n_img_max;
  *mom_doubles = slots * a.mom_cap * 4 * a.n_phases;
  *meta_int2 = slots * a.mom_cap * a.n_phases;
  *cnt_ints = slots * a.n_phases;
}

void azinv_tile_sizes(const AzinvArgs& a, size_t* tile_doubles, size_t* tmeta_ints, size_t* thdr_int2) {
  const size_t slots = (size_t)a.Q * a.n_rings * a.n_img_max;
  const size_t n_tiles = (size_t)(a.n_phases + kTilePhases - 1) / kTilePhases;
  *tile_doubles = slots * n_tiles * a.tile_cap * 32;
  *tmeta_ints = slots * n_tiles * a.tile_cap * kTilePhases;
  *thdr_int2 = slots * n_tiles;
}

void azinv_workspace_sizes(const AzinvArgs& a, size_t* leaf_doubles, size_t* hdr_doubles, size_t* ihdr_ints,
                           size_t* slab_doubles) {
  const size_t rings = (size_t)a.Q * a.n_rings;
  // ihdr_ints also covers the per-chunk row table: [rings][n_chunks] int2 after the headers
  *leaf_doubles = rings * a.n_img_max * 4 * a.n_leaves;
  *hdr_doubles = rings * kDHdr;
  *ihdr_ints = rings * kIHdr + 2 * rings * 2 * (size_t)((a.n_energies + kNEC - 1) / kNEC);   // hot + correction tables
  const int nmu = a.hot.nmu > a.els.nmu ? a.hot.nmu : a.els.nmu;
  *slab_doubles = (a.hot_atm_ext == 2 || a.else_atm_ext == 2) ? rings * (size_t)nmu * a.slab_rows_ring : 0;
}

void azinv_slab_budgets(const AtmTable& t, const double* energies, int n_energies, int* rows_chunk,
                        int* rows_ring) {
  if (t.min_dlogE <= 0.0) { int rc = t.nE + 1; while ((rc & 3) != 2) ++rc; *rows_chunk = rc; *rows_ring = (t.nE + 1) & ~1; return; }
  double span = 0.0;
  for (int e0 = 0; e0 < n_energies; e0 += kNEC) {
    const int e1 = (e0 + kNEC < n_energies ? e0 + kNEC : n_energies) - 1;
    const double s = log10(energies[e1] / energies[e0]);
    if (s > span) span = s;
  }
  // chunk rows: + 1 for starting on an even row (16-byte aligned copies), then up to the next count = 2 (mod 4)
  // (the shared-memory row stride, slab_ctx_load); the ring budget is even
  int rc = (int)ceil((span + kDopplerDex) / t.min_dlogE) + 6;
  int rr = (int)ceil((log10(energies[n_energies - 1] / energies[0]) + kDopplerDex) / t.min_dlogE) + 6;
  if (rc > t.nE) rc = t.nE;
  rc += 1;
  while ((rc & 3) != 2) ++rc;
  const int nE_even = (t.nE + 1) & ~1;
  rr = (rr + 1) & ~1;
  *rows_chunk = rc;
  *rows_ring = rr > nE_even ? nE_even : rr;
}

// 2-D tensor map over a slab workspace viewed as [Q * n_rings * nmu rows][slab_rows_ring doubles]; one box is the
// (mu, energy-row) tile [nmu][slab_ne_max] a flux CTA keeps in shared memory
static cudaError_t encode_slab_map(CUtensorMap* tm, const double* ws, const AzinvArgs& a, int nmu) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || !ptr) return cudaErrorNotSupported;
    fn = reinterpret_cast<EncodeFn>(ptr);
  }
  if (nmu > 256 || a.slab_ne_max > 256 || (a.slab_rows_ring & 1) || (a.slab_ne_max & 1)) return cudaErrorNotSupported;
  const cuuint64_t gdim[2] = {(cuuint64_t)a.slab_rows_ring, (cuuint64_t)a.Q * a.n_rings * nmu};
  const cuuint64_t gstride[1] = {(cuuint64_t)a.slab_rows_ring * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)a.slab_ne_max, (cuuint32_t)nmu};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ws), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int ATM, int CORR, int BEAM, int CUBIC, int NLP>
static cudaError_t launch_flux_n(const AzinvArgs& a, dim3 grid, size_t smem, cudaStream_t stream,
                                 const CUtensorMap& tm_hot, const CUtensorMap& tm_els) {
  cudaError_t err = cudaFuncSetAttribute(k_azinv_flux<ATM, CORR, BEAM, CUBIC, NLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  k_azinv_flux<ATM, CORR, BEAM, CUBIC, NLP><<<grid, kFluxThreads, smem, stream>>>(a, tm_hot, tm_els);
  return cudaGetLastError();
}

static size_t flux_mma_smem_bytes(const AzinvArgs& a, int atm, int corr) {
  size_t d = 2ul * a.n_leaves + (size_t)a.n_leaves * kRowH;                    // phases, 1/h, node-form rows
  d = (d + 15) & ~15ul;
  if (atm == 2) d += 5ul * a.slab_ne_max + ((a.hot.nmu + 1) & ~1) + (((size_t)a.hot.nmu * a.slab_ne_max + 15) & ~15ul);
  if (corr == 2) d += 5ul * a.slab_ne_max + ((a.els.nmu + 1) & ~1) + (((size_t)a.els.nmu * a.slab_ne_max + 15) & ~15ul);
  return d * sizeof(double);
}

template <int ATM, int CORR, int NLP>
static cudaError_t launch_flux_mma_n(const AzinvArgs& a, dim3 grid, size_t smem, cudaStream_t stream,
                                     const CUtensorMap& tm_hot, const CUtensorMap& tm_els) {
  cudaError_t err = cudaFuncSetAttribute(k_azinv_flux_mma<ATM, CORR, NLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  k_azinv_flux_mma<ATM, CORR, NLP><<<grid, kMmaThreads, smem, stream>>>(a, tm_hot, tm_els);
  return cudaGetLastError();
}

// the tensor-core flux kernel (rings whose tiles fit); the scalar kernel launched behind it takes the other rings
template <int ATM, int CORR>
static cudaError_t launch_flux_mma(const AzinvArgs& a, dim3 grid, cudaStream_t stream) {
  CUtensorMap tm_hot, tm_els;
  memset(&tm_hot, 0, sizeof(tm_hot)); memset(&tm_els, 0, sizeof(tm_els));
  cudaError_t err;
  if (ATM == 2 && (err = encode_slab_map(&tm_hot, a.ws_slab, a, a.hot.nmu)) != cudaSuccess) return err;
  if (CORR == 2 && (err = encode_slab_map(&tm_els, a.ws_slab2, a, a.els.nmu)) != cudaSuccess) return err;
#ifdef XB_PAD_SMEM
  const size_t smem = flux_mma_smem_bytes(a, ATM, CORR) + XB_PAD_SMEM;       // occupancy experiment (dev/build_variants.sh)
#else
  const size_t smem = flux_mma_smem_bytes(a, ATM, CORR);
#endif
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  if (a.n_leaves == a.n_phases && a.n_leaves == 100) return launch_flux_mma_n<ATM, CORR, 100>(a, grid, smem, stream, tm_hot, tm_els);
  if (a.n_leaves == a.n_phases && a.n_leaves == 64) return launch_flux_mma_n<ATM, CORR, 64>(a, grid, smem, stream, tm_hot, tm_els);
  return launch_flux_mma_n<ATM, CORR, 0>(a, grid, smem, stream, tm_hot, tm_els);
}

template <int ATM, int CORR, int BEAM, int CUBIC>
static cudaError_t launch_flux_b(const AzinvArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  CUtensorMap tm_hot, tm_els;
  memset(&tm_hot, 0, sizeof(tm_hot)); memset(&tm_els, 0, sizeof(tm_els));
  cudaError_t err;
  if (ATM == 2 && (err = encode_slab_map(&tm_hot, a.ws_slab, a, a.hot.nmu)) != cudaSuccess) return err;
  if (CORR == 2 && (err = encode_slab_map(&tm_els, a.ws_slab2, a, a.els.nmu)) != cudaSuccess) return err;
  // the plain configurations get instantiations for the usual leaf / phase grids
  if (!BEAM && !CUBIC && a.n_leaves == a.n_phases) {
    if (a.n_leaves == 100) return launch_flux_n<ATM, CORR, 0, 0, 100>(a, grid, smem, stream, tm_hot, tm_els);
    if (a.n_leaves == 64) return launch_flux_n<ATM, CORR, 0, 0, 64>(a, grid, smem, stream, tm_hot, tm_els);
  }
  return launch_flux_n<ATM, CORR, BEAM, CUBIC, 0>(a, grid, smem, stream, tm_hot, tm_els);
}
template <int ATM, int CORR>
static cudaError_t launch_flux(const AzinvArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  const bool cubic = a.phase_interp == kCubic;
  if (a.beam_opt != 0)
    return cubic ? launch_flux_b<ATM, CORR, 1, 1>(a, grid, smem, stream) : launch_flux_b<ATM, CORR, 1, 0>(a, grid, smem, stream);
  return cubic ? launch_flux_b<ATM, CORR, 0, 1>(a, grid, smem, stream) : launch_flux_b<ATM, CORR, 0, 0>(a, grid, smem, stream);
}

cudaError_t launch_integrate_azinv(AzinvArgs a, cudaStream_t stream) {
  if (a.n_phases > kFluxThreads) return cudaErrorInvalidValue;
  if (a.phase_interp == kCubic && a.n_leaves > kMaxCubicNodes) return cudaErrorInvalidValue;
  if (a.n_img_max > kMaxImages || a.n_img_max < 1) return cudaErrorInvalidValue;
  if (!a.ws_leaf || !a.ws_ihdr || !a.ws_hdr || !a.log10_energies) return cudaErrorInvalidValue;
  a.ws_chunk = a.ws_ihdr + (size_t)a.Q * a.n_rings * kIHdr;      // 8-byte aligned: kIHdr is even
  const int atm = a.hot_atm_ext;
  const int corr = a.corrParams ? a.else_atm_ext : 0;
  if (atm != 1 && atm != 2) return cudaErrorNotSupported;
  if (corr != 0 && corr != 1 && corr != 2) return cudaErrorNotSupported;
  if (a.beam_opt < 0 || a.beam_opt > 3 || (a.beam_opt != 0 && a.n_params < 7)) return cudaErrorNotSupported;
  if (a.R_in <= 0.0) a.R_in = 1.0e6;
  if (!a.corrParams) a.else_atm_ext = 0;
  if ((atm == 2 && !a.ws_slab) || (corr == 2 && !a.ws_slab2)) return cudaErrorInvalidValue;
  if ((atm == 2 || corr == 2) && (a.slab_ne_max < 4 || a.slab_rows_ring < 4)) return cudaErrorInvalidValue;
  const size_t gsm = geom_smem_bytes(a), fsm = flux_smem_bytes(a, atm, corr);
  if (gsm > 227 * 1024 || fsm > 227 * 1024) return cudaErrorInvalidValue;
  const int n_chunks = (a.n_energies + kNEC - 1) / kNEC;
  dim3 ggrid(a.n_rings, a.Q), fgrid(a.n_rings * n_chunks, a.Q);
  cudaError_t err;
  a.general = 0;
  if ((err = launch_azinv_geometry(a, stream)) != cudaSuccess) return err;
  const size_t msm = ((size_t)kGSpan * kSlabMThreads + 4ul * a.n_rings) * sizeof(double) + 4ul * a.n_rings * sizeof(int);
  if (atm == 2) {
    k_azinv_slab<<<ggrid, kSlabThreads, 0, stream>>>(a, 0);
    k_azinv_slab_member<<<dim3((a.hot.nmu + kSlabMRows - 1) / kSlabMRows, a.Q), kSlabMThreads, msm, stream>>>(a, 0);
  }
  if (corr == 2) {
    k_azinv_slab<<<ggrid, kSlabThreads, 0, stream>>>(a, 1);
    k_azinv_slab_member<<<dim3((a.els.nmu + kSlabMRows - 1) / kSlabMRows, a.Q), kSlabMThreads, msm, stream>>>(a, 1);
  }
  if (!a.ws_cells) return cudaErrorInvalidValue;
  // tensor-core accumulation: plain configurations only (no beaming, not the global C2 spline)
  if (a.ws_tiles && (a.beam_opt != 0 || a.phase_interp == kCubic || !a.ws_tmeta || !a.ws_thdr || a.tile_cap < 8 ||
                     (a.tile_cap & 3) || a.n_leaves > 256 || a.n_azi > 0xffff)) {
    a.ws_tiles = nullptr;
  }
  if (a.ws_tiles && !a.ws_redo) a.ws_tiles = nullptr;
  if (a.ws_tiles) a.ws_mom = nullptr;       // rings whose tiles overflow walk their cells inside the scalar kernel
  if (a.ws_tiles) {
    cudaError_t em = cudaMemsetAsync(a.ws_redo, 0, (size_t)a.Q * a.n_rings * n_chunks * sizeof(int), stream);
    if (em != cudaSuccess) return em;
    // ws_ovf = [count | ring list (Q * n_rings) | listed flags (Q * n_rings)]: count and flags start at zero
    if (a.ws_ovf) {
      if ((em = cudaMemsetAsync(a.ws_ovf, 0, sizeof(int), stream)) != cudaSuccess) return em;
      if ((em = cudaMemsetAsync(a.ws_ovf + 1 + (size_t)a.Q * a.n_rings, 0, (size_t)a.Q * a.n_rings * sizeof(int), stream)) != cudaSuccess) return em;
    }
  }
  a.ovf_list_mode = 0;
  if (a.ws_tiles) {
    const size_t msm = (2ul * a.n_azi + a.n_leaves) * sizeof(double) + 2ul * a.n_leaves;
    k_azinv_tiles<<<ggrid, kMomThreads, msm, stream>>>(a);
  } else if (a.ws_mom) {
    if (!a.ws_meta || !a.ws_cnt || a.mom_cap < 1 || a.n_azi > 0xffff) return cudaErrorInvalidValue;
    const size_t msm = (2ul * a.n_azi + a.n_leaves) * sizeof(double) + 2ul * a.n_leaves;
    k_azinv_moments<<<ggrid, kMomThreads, msm, stream>>>(a);
  } else k_azinv_cells<<<ggrid, 32, 0, stream>>>(a);
  if ((err = cudaGetLastError()) != cudaSuccess) return err;
  if (a.ev_flux[0]) cudaEventRecord(a.ev_flux[0], stream);
  if (a.ws_tiles) {
    if (atm == 1 && corr == 0) err = launch_flux_mma<1, 0>(a, fgrid, stream);
    else if (atm == 1 && corr == 1) err = launch_flux_mma<1, 1>(a, fgrid, stream);
    else if (atm == 1 && corr == 2) err = launch_flux_mma<1, 2>(a, fgrid, stream);
    else if (atm == 2 && corr == 0) err = launch_flux_mma<2, 0>(a, fgrid, stream);
    else if (atm == 2 && corr == 1) err = launch_flux_mma<2, 1>(a, fgrid, stream);
    else err = launch_flux_mma<2, 2>(a, fgrid, stream);
    if (err != cudaSuccess) return err;
  }
  if (a.ws_tiles && a.ws_ovf) {       // the scalar kernel only has the listed rings to do
    a.ovf_list_mode = 1;
    int G = a.Q * a.n_rings;
    if (G > 148 * 4) G = 148 * 4;
    fgrid = dim3(n_chunks, G);
  }
  if (atm == 1 && corr == 0) err = launch_flux<1, 0>(a, fgrid, fsm, stream);
  else if (atm == 1 && corr == 1) err = launch_flux<1, 1>(a, fgrid, fsm, stream);
  else if (atm == 1 && corr == 2) err = launch_flux<1, 2>(a, fgrid, fsm, stream);
  else if (atm == 2 && corr == 0) err = launch_flux<2, 0>(a, fgrid, fsm, stream);
  else if (atm == 2 && corr == 1) err = launch_flux<2, 1>(a, fgrid, fsm, stream);
  else err = launch_flux<2, 2>(a, fgrid, fsm, stream);
  if (err != cudaSuccess) return err;
  if (a.ev_flux[1]) cudaEventRecord(a.ev_flux[1], stream);
  err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  if (a.flux_part) {
    k_azinv_reduce_rings<<<dim3(a.n_energies, a.Q), 128, 0, stream>>>(a);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
  }
  if (a.scale_by_energy) {
    const long n = (long)a.Q * a.n_energies * a.n_phases;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_scale_flux<<<blocks, 256, 0, stream>>>(a.flux, a.energies, a.Q, a.n_energies, a.n_phases);
    err = cudaGetLastError();
  }
  return err;
}

}  // namespace xb
