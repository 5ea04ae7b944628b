// Pulse integrator exploiting azimuthal invariance of a hot-region member.
//
// Replaces xpsi/cellmesh/integrator_for_azimuthal_invariance.pyx:70-665 (the
// reference's dominant hot loop) together with the atmosphere evaluations it
// calls: hot_BB.pyx:54-98 and hot_Num4D.pyx:248-460.
//
// B200 mapping (not a translation of the OpenMP loop nest):
//   * one CTA per (member instance q, mesh ring i); the batch axis q = theta x
//     member is just the second grid dimension;
//   * ring constants, the ring's ray row and its cos(psi) image live in shared
//     memory; the three Steffen splines of the reference are never
//     materialised -- each half-leaf thread rebuilds the two node slopes it
//     needs in registers;
//   * geometry of all image orders is done at once (one thread per (image,
//     half-leaf)); the reference's sequential visibility state machine is
//     replayed verbatim by one thread per image;
//   * Num4D: the 4-D table is contracted over (log T, log g) -- constant on a
//     ring -- into a (mu, E) slab restricted to the reachable energy rows and
//     kept in shared memory, so every intensity is a 4x4 stencil on-chip;
//   * energies are processed in tiles; for each tile the leaf profile and its
//     phase-spline coefficients are built in shared memory, then each thread
//     owns (output phase, 4 energies), walks the ring's radiating cells and
//     accumulates in registers, reloading spline coefficients only when the
//     cell crosses into the next leaf interval;
//   * rings are combined with fp64 RED atomics into flux[q, E, P].
#include "common.cuh"
#include "kernels.h"

namespace xb {

constexpr int kET = 8;        // energies per tile
constexpr int kEG = 4;        // energies per thread item
constexpr int kThreads = 256;

struct LeafSet {              // per image order, N_L entries each
  double* phase;              // PHASE as left by the visibility state machine
  double* ptrue;              // leaf + lag for visible leaves
  double* Z;                  // total redshift eta * Grav_z
  double* abb;                // mu * eta
  double* geom;               // mu |deriv| Grav_z eta^3 / (1 + beta cos xi)
  double* muw;                // Num4D: 4 Lagrange weights in mu
  int* mub;                   // Num4D: base node in mu
  int* vis;                   // leaf carries signal
};

// leaf arrays are laid out [image][field][N_L]; resolved arithmetically so the
// image index never forces the pointer table into local memory
__device__ __forceinline__ LeafSet leaf_set(double* dbase, int* ibase, int I, int N_L, bool num4d) {
  const int nf = num4d ? 9 : 5;
  double* d = dbase + (long)I * nf * N_L;
  int* ii = ibase + (long)I * 2 * N_L;
  LeafSet S;
  S.phase = d; S.ptrue = d + N_L; S.Z = d + 2 * N_L; S.abb = d + 3 * N_L; S.geom = d + 4 * N_L;
  S.muw = num4d ? d + 5 * N_L : nullptr;
  S.vis = ii; S.mub = ii + N_L;
  return S;
}

__device__ __forceinline__ double bb_intensity(double E, double kT) {
  return E * E * E / (exp(E / kT) - 1.0);     // hot_BB.pyx:85-87
}

template <int ATM>
__global__ void __launch_bounds__(kThreads)
k_integrate_azinv(AzinvArgs a) {
  const int i = blockIdx.x;                 // ring
  const int q = blockIdx.y;                 // member instance
  const int tid = threadIdx.x;
  const int R_ = a.n_rings_q ? a.n_rings_q[q] : a.n_rings;
  const int A_ = a.n_azi_q ? a.n_azi_q[q] : a.n_azi;
  if (i >= R_) return;
  const int N_R = a.n_rays, N_E = a.n_energies, N_L = a.n_leaves, N_P = a.n_phases;
  const long ring = (long)q * a.n_rings + i;           // padded ring index
  const long cell0 = ring * a.n_azi;                    // padded cell row
  const int leaf_lim = (N_L % 2 == 0) ? N_L / 2 : (N_L + 1) / 2;

  extern __shared__ double smem[];
  __shared__ int s_J, s_jhalf, s_nimg, s_fail, s_elo, s_ne;
  __shared__ int s_inv2[kMaxImages], s_dom[kMaxImages], s_mono[kMaxImages];
  __shared__ double s_wT[4], s_wG[4];
  __shared__ int s_bT, s_bG;

  // ---- does the ring radiate? (pyx:286-296) --------------------------------
  if (tid == 0) { s_J = A_; s_jhalf = N_R - 1; s_fail = 0; }
  if (tid < kMaxImages) { s_inv2[tid] = 0; s_dom[tid] = 0; s_mono[tid] = 0; }
  __syncthreads();
  // CELL_RADIATES (HotRegion.py:965); a null mask means cellArea > 0
  auto radiates = [&](int j) -> bool {
    return a.radiates ? (a.radiates[cell0 + j] == 1) : (a.cellArea[cell0 + j] > 0.0);
  };
  for (int j = tid; j < A_; j += kThreads)
    if (radiates(j)) atomicMin(&s_J, j);
  __syncthreads();
  const int J = s_J;
  if (J >= A_) return;

  // ---- shared-memory carve --------------------------------------------------
  double* sp = smem;
  double* s_defl = sp; sp += N_R;
  double* s_calpha = sp; sp += N_R;
  double* s_lag = sp; sp += N_R;
  double* s_cosd = sp; sp += N_R;
  double* s_phi = sp; sp += a.n_azi;
  double* s_area = sp; sp += a.n_azi;
  double* s_E = sp; sp += N_E;
  double* s_logE = sp; sp += N_E;
  double* s_y = sp; sp += kET * N_L;
  double* s_coef = sp; sp += kET * N_L * 4;
  const int n_img_max = a.n_img_max;
  double* s_leafd = sp; sp += (long)n_img_max * (ATM == 2 ? 9 : 5) * N_L;
  double* s_axE = nullptr; double* s_axMu = nullptr; double* s_slab = nullptr;
  if (ATM == 2) {
    s_axE = sp; sp += a.hot.nE;
    s_axMu = sp; sp += a.hot.nmu;
    s_slab = sp; sp += (long)a.hot.nmu * a.slab_ne_max;
  }
  int* s_leafi = reinterpret_cast<int*>(sp);

  // ---- stage the ring ---------------------------------------------------------
  const double* g_defl = a.deflection + ring * N_R;
  const double* g_calpha = a.cos_alpha + ring * N_R;
  const double* g_lag = a.lag + ring * N_R;
  for (int r = tid; r < N_R; r += kThreads) {
    const double d = g_defl[r];
    s_defl[r] = d; s_calpha[r] = g_calpha[r]; s_lag[r] = g_lag[r];
    s_cosd[r] = cos(d);                                   // pyx:216-218
    if (d > kHalfPi) atomicMin(&s_jhalf, r);              // pyx:301-303
  }
  for (int j = tid; j < A_; j += kThreads) {
    s_phi[j] = a.phi[cell0 + j];
    s_area[j] = radiates(j) ? a.cellArea[cell0 + j] : -1.0;   // <0: cell is dark
  }
  for (int e = tid; e < N_E; e += kThreads) {
    const double E = a.energies[e];
    s_E[e] = E; s_logE[e] = log10(E);
  }
  if (ATM == 2) {
    for (int e = tid; e < a.hot.nE; e += kThreads) s_axE[e] = a.hot.logE[e];
    for (int m = tid; m < a.hot.nmu; m += kThreads) s_axMu[m] = a.hot.mu[m];
  }

  // ---- ring constants (pyx:318-331) ---------------------------------------------
  const double inclination = a.inclination[q];
  const double omega = a.omega[q];
  const double sin_i = sin(inclination), cos_i = cos(inclination);
  const double radius = a.radial[ring];
  const double rsr = a.r_s_over_r[ring];
  const double Grav_z = sqrt(1.0 - rsr);
  const double cos_gamma = a.cos_gamma[ring];
  const double sin_gamma = sqrt(1.0 - cos_gamma * cos_gamma);
  const double theta_i = a.theta[ring * a.theta_ring_stride];
  const double cos_theta_i = cos(theta_i), sin_theta_i = sin(theta_i);
  const double theta_i_over_pi = theta_i / kPi;
  const double beta = radius * omega * sin_theta_i / (kC * Grav_z);
  const double Lorentz = sqrt(1.0 - beta * beta);
  const double maxDefl = a.maxDeflection[ring];
  const double* VEC = a.srcParams + (a.params_per_cell ? (cell0 + J) : ring) * a.n_params;
  const double logT = VEC[0];
  const double kT = kKBOverKeV * pow(10.0, logT);
  const double log_kT = log10(kT);
  int n_img = a.image_order_limit > 0 ? a.image_order_limit : (int)ceil(maxDefl / kPi);
  if (n_img > n_img_max) n_img = n_img_max;
  __syncthreads();
  const int jhalf = s_jhalf;        // first ray with deflection > pi/2 (clamped)

  // ---- Num4D: (T,g) stencil + reachable energy rows, then contract the slab --------
  if (ATM == 2) {
    if (tid == 0) {
      View vT{a.hot.logT, 1}, vG{a.hot.logg, 1};
      s_bT = lagrange_base(vT, a.hot.nT, logT);
      s_bG = lagrange_base(vG, a.hot.ng, VEC[1]);
      double w[4];
      lagrange_weights(vT, s_bT, logT, w);
      for (int k = 0; k < 4; ++k) s_wT[k] = w[k];
      lagrange_weights(vG, s_bG, VEC[1], w);
      for (int k = 0; k < 4; ++k) s_wG[k] = w[k];
      // log10(E'/kT) spans [logE_0 - log Zmax, logE_last - log Zmin] - log kT
      const double Zmax = Lorentz / (1.0 - fabs(beta)) * Grav_z;
      const double Zmin = Lorentz / (1.0 + fabs(beta)) * Grav_z;
      const double vlo = s_logE[0] - log10(Zmax) - log_kT - 1.0e-9;
      const double vhi = s_logE[N_E - 1] - log10(Zmin) - log_kT + 1.0e-9;
      const int elo = lagrange_base(s_axE, a.hot.nE, vlo);
      const int ehi = lagrange_base(s_axE, a.hot.nE, vhi) + 4;      // exclusive
      if (ehi - elo > a.slab_ne_max) {      // budget too small for this ring: refuse, never clamp
        atomicExch(a.status + q, kUnsupported);
        s_fail = 1;
      }
      s_elo = elo; s_ne = ehi - elo;
    }
    __syncthreads();
    if (s_fail) return;
    const int elo = s_elo, ne = s_ne, nmu = a.hot.nmu;
    const long S0 = (long)a.hot.ng * nmu * a.hot.nE, S1 = (long)nmu * a.hot.nE, S2 = a.hot.nE;
    for (int t = tid; t < nmu * ne; t += kThreads) {
      const int m = t / ne, e = t - m * ne;
      const double* base = a.hot.buf + (long)s_bT * S0 + (long)s_bG * S1 + (long)m * S2 + elo + e;
      double acc = 0.0;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        double inner = 0.0;
#pragma unroll
        for (int y = 0; y < 4; ++y) inner += s_wG[y] * __ldg(base + x * S0 + y * S1);
        acc += s_wT[x] * inner;
      }
      s_slab[t] = acc;
    }
  }

  // ---- geometry of every (image, half-leaf) (pyx:342-441) ----------------------------
  View vDefl{s_defl, 1}, vCalpha{s_calpha, 1}, vLag{s_lag, 1};
  View vAltX{s_cosd + jhalf, -1}, vAltY{s_calpha + jhalf, -1};   // pyx:305-308
  const int n_alt = jhalf + 1;
  const double alt_xmin = s_cosd[jhalf];
  for (int t = tid; t < n_img * leaf_lim; t += kThreads) {
    const int I = t / leaf_lim, k = t - I * leaf_lim;
    const LeafSet S = leaf_set(s_leafd, s_leafi, I, N_L, ATM == 2);
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
        if (mu > 0.0) visible = 1;      // R_in >= 1e6: no disc (pyx:390-396)
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
      S.vis[kdx] = visible;
      if (!visible) continue;
      double superlum, eta;
      if (!are_equal(psi, 0.0)) {
        const double cos_xi = sin_alpha * sin_i * sin(a.leaves[kdx]) / sin_psi;
        superlum = 1.0 + beta * cos_xi;
        eta = Lorentz / superlum;
      } else { superlum = 1.0; eta = Lorentz; }
      S.Z[kdx] = eta * Grav_z;
      S.abb[kdx] = mu * eta;
      S.geom[kdx] = mu * fabs(deriv) * Grav_z * eta * eta * eta / superlum;
      S.ptrue[kdx] = a.leaves[kdx] + lagv;
    }
  }
  __syncthreads();

  // ---- visibility state machine, one thread per image (pyx:339-563, verbatim order) ----
  if (tid < n_img) {
    const LeafSet S = leaf_set(s_leafd, s_leafi, tid, N_L, ATM == 2);
    double* PH = S.phase;
    int Inv = 2, k0 = 0;
    for (int k = 0; k < leaf_lim; ++k) {
      if (S.vis[k]) {
        PH[k] = S.ptrue[k];
        const bool mirror = (0 < k && k < leaf_lim - 1) || (k == leaf_lim - 1 && N_L % 2 == 0);
        if (mirror) PH[N_L - 1 - k] = S.ptrue[N_L - 1 - k];
        if (k == 0) {
          PH[N_L - 1] = PH[0] + kTwoPi;            // leaf N_L-1 is a copy of leaf 0
          S.Z[N_L - 1] = S.Z[0]; S.abb[N_L - 1] = S.abb[0]; S.geom[N_L - 1] = S.geom[0];
          S.vis[N_L - 1] = 1;
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
        if (k == 0) S.vis[N_L - 1] = 0;
        if (Inv == 0) {
          const double step = (PH[N_L - k] - PH[k - 1]) / (double)(N_L - 2 * k + 1);
          for (int m = k; m < N_L - k; ++m) PH[m] = PH[m - 1] + step;
          Inv = 1; k0 = k;
        }
      }
    }
    s_inv2[tid] = (Inv == 2);
    if (Inv != 2) {
      for (int m = 1; m < N_L; ++m)
        if (PH[m] <= PH[m - 1]) { s_mono[tid] = 1; break; }          // pyx:556-563
    }
  }
  __syncthreads();
  if (tid == 0) {
    // The reference walks image orders in sequence: the first never-visible order ends the
    // loop (pyx:553-554), so failures on orders it would not have reached are not failures.
    int n = 0, bad = 0;
    while (n < n_img && !s_inv2[n]) { bad |= s_dom[n] | s_mono[n]; ++n; }
    if (n < n_img) bad |= s_dom[n];          // that order's leaf loop did run
    if (bad) atomicExch(a.status + q, kNumericalError);
    s_nimg = n; s_fail = bad;
  }
  __syncthreads();
  if (s_fail) return;
  n_img = s_nimg;
  if (a.work && tid == 0) {        // algorithmic-work counters for the roofline
    int reached = (n_img < ((a.image_order_limit > 0) ? a.image_order_limit : n_img_max)) ? n_img + 1 : n_img;
    unsigned long long V = 0, Kc = 0;
    for (int I = 0; I < n_img; ++I) {
      const LeafSet S = leaf_set(s_leafd, s_leafi, I, N_L, ATM == 2);
      for (int l = 0; l < N_L; ++l) V += S.vis[l] ? 1 : 0;
    }
    for (int j = 0; j < A_; ++j) Kc += (s_area[j] >= 0.0) ? 1 : 0;
    atomicAdd(a.work + 0, (unsigned long long)reached * leaf_lim);
    atomicAdd(a.work + 1, V);
    atomicAdd(a.work + 2, (unsigned long long)n_img);
    atomicAdd(a.work + 3, Kc * n_img);
  }
  if (n_img == 0) return;

  // ---- Num4D: per-leaf mu stencils -------------------------------------------------------
  if (ATM == 2) {
    for (int t = tid; t < n_img * N_L; t += kThreads) {
      const int I = t / N_L, l = t - I * N_L;
      const LeafSet S = leaf_set(s_leafd, s_leafi, I, N_L, true);
      if (!S.vis[l]) continue;
      const double v = S.abb[l];
      const int b = lagrange_base(s_axMu, a.hot.nmu, v);
      double w[4];
      lagrange_weights(s_axMu, b, v, w);
      S.mub[l] = b;
#pragma unroll
      for (int k = 0; k < 4; ++k) S.muw[4 * l + k] = w[k];
      S.Z[l] = log10(S.Z[l]);                // only log10 Z is needed from here on
    }
    __syncthreads();
  }

  const double norm = (ATM == 2) ? (kErg / kHKeV) * pow(10.0, 3.0 * logT)
                                 : kErg * kPlanckDistConst;
  const int interp_kind = a.phase_interp;
  const bool periodic = (interp_kind != kSteffen);
  double* flux_q = a.flux + (long)q * N_E * N_P;

  // ---- energy tiles ------------------------------------------------------------------------
  for (int e0 = 0; e0 < N_E; e0 += kET) {
    const int et = min(kET, N_E - e0);
    const int n_groups = (et + kEG - 1) / kEG;
    const int n_items = n_groups * N_P;
    // every thread owns at most two (phase, energy-group) items per tile
    double acc[2][kEG];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int g = 0; g < kEG; ++g) acc[s][g] = 0.0;

    for (int I = 0; I < n_img; ++I) {
      const LeafSet S = leaf_set(s_leafd, s_leafi, I, N_L, ATM == 2);
      // (1) leaf profile for the tile (pyx:445-478)
      for (int t = tid; t < et * N_L; t += kThreads) {
        const int e = t / N_L, l = t - e * N_L;
        double val = 0.0;
        if (S.vis[l]) {
          if (ATM == 1) {
            val = bb_intensity(s_E[e0 + e] / S.Z[l], kT) * norm * S.geom[l];
          } else {
            const double v = s_logE[e0 + e] - S.Z[l] - log_kT;     // log10(E'/kT)
            int bE = lagrange_base(s_axE, a.hot.nE, v);
            double wE[4];
            lagrange_weights(s_axE, bE, v, wE);
            bE -= s_elo;                       // inside the slab by construction of [elo, ehi)
            const double* row = s_slab + (long)S.mub[l] * s_ne + bE;
            const double* wM = S.muw + 4 * l;
            double sum = 0.0;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const double* r = row + x * s_ne;
              sum += wM[x] * (wE[0] * r[0] + wE[1] * r[1] + wE[2] * r[2] + wE[3] * r[3]);
            }
            if (sum < 0.0) sum = 0.0;                              // hot_Num4D.pyx:436-437
            val = sum * norm * S.geom[l];
          }
        }
        s_y[e * N_L + l] = val;
      }
      __syncthreads();
      // (2) phase-spline coefficients (pyx:566-569)
      View vPH{S.phase, 1};
      for (int t = tid; t < et * (N_L - 1); t += kThreads) {
        const int e = t / (N_L - 1), l = t - e * (N_L - 1);
        View vY{s_y + e * N_L, 1};
        double b, c, d;
        interp_coeffs(interp_kind, periodic, vPH, vY, N_L, l, &b, &c, &d);
        double* o = s_coef + ((long)e * N_L + l) * 4;
        o[0] = vY[l]; o[1] = b; o[2] = c; o[3] = d;
      }
      __syncthreads();
      // (3) evaluate at cell-shifted phases and accumulate (pyx:571-596)
      const double ph_first = S.phase[0], ph_last = S.phase[N_L - 1];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int item = tid + s * kThreads;
        if (item >= n_items) break;
        const int g = item / N_P, k = item - g * N_P;
        const int eb = g * kEG;
        const int ng = min(kEG, et - eb);
        const double phk = a.phases[k];
        int idx = -1;
        double xprev = 0.0;
        double c0[kEG], c1[kEG], c2[kEG], c3[kEG];
        int loaded = -1;
        for (int j = 0; j < A_; ++j) {
          const double area = s_area[j];
          if (area < 0.0) continue;
          double x = phk + s_phi[j];
          if (x > ph_last) { while (x > ph_last) x -= kTwoPi; }
          else if (x < ph_first) { while (x < ph_first) x += kTwoPi; }
          if (x < ph_first || x > ph_last) { atomicExch(a.status + q, kNumericalError); continue; }
          if (idx < 0 || x < xprev) idx = interval_search(vPH, N_L, x);
          else idx = interval_walk(vPH, N_L, x, idx);
          xprev = x;
          if (idx != loaded) {
#pragma unroll
            for (int gg = 0; gg < kEG; ++gg) {
              if (gg < ng) {
                const double2* cp =
                    reinterpret_cast<const double2*>(s_coef + ((long)(eb + gg) * N_L + idx) * 4);
                const double2 lo = cp[0], hi = cp[1];
                c0[gg] = lo.x; c1[gg] = lo.y; c2[gg] = hi.x; c3[gg] = hi.y;
              }
            }
            loaded = idx;
          }
          const double d = x - S.phase[idx];
#pragma unroll
          for (int gg = 0; gg < kEG; ++gg) {
            if (gg < ng) {
              const double f = c0[gg] + d * (c1[gg] + d * (c2[gg] + d * c3[gg]));
              if (f > 0.0) acc[s][gg] += area * f;
            }
          }
        }
      }
      __syncthreads();
    }
    // (4) ring contribution -> flux[q, e, k]
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int item = tid + s * kThreads;
      if (item >= n_items) break;
      const int g = item / N_P, k = item - g * N_P;
#pragma unroll
      for (int gg = 0; gg < kEG; ++gg) {
        const int e = e0 + g * kEG + gg;
        if (g * kEG + gg < et && acc[s][gg] != 0.0) atomicAdd(flux_q + (long)e * N_P + k, acc[s][gg]);
      }
    }
  }
}

// flux[q, e, k] /= E_e keV   (pyx:610-612)
__global__ void k_scale_flux(double* flux, const double* energies, int Q, int N_E, int N_P) {
  const long n = (long)Q * N_E * N_P;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    const int e = (int)((t / N_P) % N_E);
    flux[t] = flux[t] / (energies[e] * kKeV);
  }
}

size_t azinv_smem_bytes(const AzinvArgs& a, int atm) {
  size_t d = 4ul * a.n_rays + 2ul * a.n_azi + 2ul * a.n_energies + (size_t)kET * a.n_leaves * 5;
  d += (size_t)a.n_img_max * a.n_leaves * (5 + (atm == 2 ? 4 : 0));
  if (atm == 2) d += a.hot.nE + a.hot.nmu + (size_t)a.hot.nmu * a.slab_ne_max;
  size_t bytes = d * sizeof(double);
  bytes += ((size_t)a.n_img_max * a.n_leaves * 2) * sizeof(int);
  return bytes;
}

cudaError_t launch_integrate_azinv(AzinvArgs a, cudaStream_t stream) {
  if (a.n_phases * ((kET + kEG - 1) / kEG) > 2 * kThreads) return cudaErrorInvalidValue;
  if (a.n_img_max > kMaxImages || a.n_img_max < 1) return cudaErrorInvalidValue;
  const int atm = a.hot_atm_ext;
  if (atm == 2) {
    if (a.slab_ne_max <= 0 || a.slab_ne_max > a.hot.nE) a.slab_ne_max = a.hot.nE;
    if (a.slab_ne_max < 8) a.slab_ne_max = 8;
  }
  size_t smem = azinv_smem_bytes(a, atm);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  dim3 grid(a.n_rings, a.Q);
  cudaError_t err;
  if (atm == 1) {
    err = cudaFuncSetAttribute(k_integrate_azinv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k_integrate_azinv<1><<<grid, kThreads, smem, stream>>>(a);
  } else if (atm == 2) {
    err = cudaFuncSetAttribute(k_integrate_azinv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k_integrate_azinv<2><<<grid, kThreads, smem, stream>>>(a);
  } else {
    return cudaErrorNotSupported;
  }
  err = cudaGetLastError();
  if (err != cudaSuccess) return err;
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
