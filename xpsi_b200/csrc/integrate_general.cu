// Pulse integrator WITHOUT azimuthal invariance (HotRegion(symmetry=False),
// Everywhere(time_invariant=False)).
//
// Replaces xpsi/cellmesh/integrator.pyx:48-667.  The leaf geometry (and the
// reference's visibility state machine, which here also steps Z and mu*eta
// across dark leaf ranges, pyx:406-520) comes from k_azinv_geometry in its
// `general` mode.  This file holds the second stage (pyx:537-592): per (ring,
// image) GEOM is splined against the lagged leaf phase with the phase
// interpolant, Z and mu*eta with Steffen, and the atmosphere is evaluated per
// (cell, output phase, energy) with the CELL's own parameters.
//
// B200 mapping: one CTA per (ring, chunk of 32 energies, cell group, instance).
//   * the spline coefficients of every image of the ring sit in shared memory;
//   * Num4D: the table is contracted over the cell's (log T, log g) stencil
//     into a (mu, E-row) slab in shared memory holding only the rows these 32
//     energies can reach; it is rebuilt only when the parameters change from
//     one cell to the next (never, for a uniform spot);
//   * a warp owns a block of <= 32 output phases: each lane first prepares one
//     (cell, image, phase) item -- phase wrap, interval search, GEOM / Z /
//     mu*eta, mu stencil -- then the warp replays the 32 items with lane =
//     energy, so the per-item work is amortised over 32 atmosphere evaluations;
//   * per-CTA sums live in shared memory and leave with one fp64 RED per
//     (phase, energy).
#include "azinv_shared.cuh"

namespace xb {

constexpr int kGenThreads = 128;
constexpr int kGenWarps = kGenThreads / 32;
constexpr int kGenE = 32;                    // energies per CTA: one per lane

struct GenSlab {
  double* axE;      // full log10(E/kT) axis of the table
  double* axMu;     // full mu axis
  double* slab;     // [nmu][rows]
  int rows_max;
};

// contract the table over the (T,g) stencil of VEC for energy rows [elo, elo+nrows)
__device__ __forceinline__ void gen_build_slab(const AtmTable& T, const GenSlab& S, const double* VEC, int elo,
                                               int nrows, int tid) {
  View vT{T.logT, 1}, vG{T.logg, 1};
  const int bT = lagrange_base(vT, T.nT, VEC[0]);
  const int bG = lagrange_base(vG, T.ng, VEC[1]);
  double wT[4], wG[4];
  lagrange_weights(vT, bT, VEC[0], wT);
  lagrange_weights(vG, bG, VEC[1], wG);
  const long S0 = (long)T.ng * T.nmu * T.nE, S1 = (long)T.nmu * T.nE, S2 = T.nE;
  const int total = T.nmu * nrows;
  for (int t = tid; t < total; t += kGenThreads) {
    const int m = t / nrows, e = t - m * nrows;
    const double* base = T.buf + (long)bT * S0 + (long)bG * S1 + (long)m * S2 + elo + e;
    double acc = 0.0;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      double inner = 0.0;
#pragma unroll
      for (int y = 0; y < 4; ++y) inner += wG[y] * __ldg(base + x * S0 + y * S1);
      acc += wT[x] * inner;
    }
    S.slab[m * nrows + e] = acc;
  }
}

// I / T^3 at log10(E'/kT) = v with the mu stencil (bM, wM) already known (hot_Num4D.pyx:416-437)
__device__ __forceinline__ double gen_slab_eval(const GenSlab& S, int nE_tab, int elo, int nrows, double v, int bM,
                                                const double wM[4], int* bad) {
  const int bE = lagrange_base(S.axE, nE_tab, v);
  const int loc = bE - elo;
  if (loc < 0 || loc + 4 > nrows) { *bad = 1; return 0.0; }
  double wE[4];
  lagrange_weights(S.axE, bE, v, wE);
  const double* row = S.slab + (long)bM * nrows + loc;
  double sum = 0.0;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const double* r = row + x * nrows;
    sum += wM[x] * (wE[0] * r[0] + wE[1] * r[1] + wE[2] * r[2] + wE[3] * r[3]);
  }
  return sum < 0.0 ? 0.0 : sum;
}

// ATM: hot atmosphere (1 BB, 2 Num4D).  CORR: elsewhere correction (0 none, 1 BB, 2 Num4D)
template <int ATM, int CORR>
__global__ void __launch_bounds__(kGenThreads) k_general_flux(AzinvArgs a, int n_groups) {
  const int n_chunks = (a.n_energies + kGenE - 1) / kGenE;
  int bx = blockIdx.x;
  const int grp = bx % n_groups; bx /= n_groups;
  const int chunk = bx % n_chunks;
  const int i = bx / n_chunks;
  const int q = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long ring = (long)q * a.n_rings + i;
  const int n_img = a.ws_ihdr[ring * kIHdr];
  if (n_img == 0) return;
  const int A_ = a.n_azi_q ? a.n_azi_q[q] : a.n_azi;
  const int N_E = a.n_energies, N_L = a.n_leaves, N_P = a.n_phases;
  const long cell0 = ring * a.n_azi;
  const int e0 = chunk * kGenE;
  const int ne = min(kGenE, N_E - e0);
  const int n_img_max = a.n_img_max;

  extern __shared__ double smem[];
  double* sp = smem;
  double* s_acc = sp; sp += (long)N_P * kGenE;                      // [phase][energy]
  double* s_tab = sp; sp += (long)n_img_max * 13 * N_L;             // per image: PH, G[4], Z[4], A[4]
  GenSlab hot{}, els{};
  if (ATM == 2) { hot.axE = sp; sp += a.hot.nE; hot.axMu = sp; sp += a.hot.nmu; hot.slab = sp; sp += (long)a.hot.nmu * a.slab_ne_max; }
  if (CORR == 2) { els.axE = sp; sp += a.els.nE; els.axMu = sp; sp += a.els.nmu; els.slab = sp; sp += (long)a.els.nmu * a.slab_ne_max; }
  __shared__ int s_bad;

  if (tid == 0) s_bad = 0;
  for (int t = tid; t < N_P * kGenE; t += kGenThreads) s_acc[t] = 0.0;
  if (ATM == 2) {
    for (int t = tid; t < a.hot.nE; t += kGenThreads) hot.axE[t] = a.hot.logE[t];
    for (int t = tid; t < a.hot.nmu; t += kGenThreads) hot.axMu[t] = a.hot.mu[t];
  }
  if (CORR == 2) {
    for (int t = tid; t < a.els.nE; t += kGenThreads) els.axE[t] = a.els.logE[t];
    for (int t = tid; t < a.els.nmu; t += kGenThreads) els.axMu[t] = a.els.mu[t];
  }
  // ---- leaf arrays of every image, then spline coefficients (pyx:537-542) ----------------------
  for (int t = tid; t < n_img * N_L; t += kGenThreads) {
    const int I = t / N_L, l = t - I * N_L;
    const double* W = leaf_ptr(a.ws_leaf, ring, n_img_max, I, N_L);
    double* tb = s_tab + (long)I * 13 * N_L;
    tb[l] = W[l];                         // PHASE
    tb[(1 + 0) * N_L + l] = W[3 * N_L + l];       // GEOM nodes
    tb[(5 + 0) * N_L + l] = W[N_L + l];           // Z nodes
    tb[(9 + 0) * N_L + l] = W[2 * N_L + l];       // mu*eta nodes
  }
  __syncthreads();
  for (int t = tid; t < n_img * (N_L - 1); t += kGenThreads) {
    const int I = t / (N_L - 1), l = t - I * (N_L - 1);
    double* tb = s_tab + (long)I * 13 * N_L;
    const double* PH = tb;
    double b, c, d;
    if (a.phase_interp != kCubic) {
      interp_coeffs(a.phase_interp, true, PH, tb + 1 * N_L, N_L, l, &b, &c, &d);
      tb[2 * N_L + l] = b; tb[3 * N_L + l] = c; tb[4 * N_L + l] = d;
    } else if (l == 0) {             // global C2 spline: one thread per image solves the cyclic system
      double cc[kMaxCubicNodes];
      const double* yg = tb + 1 * N_L;
      cspline_second(PH, yg, N_L, true, cc);
      for (int i = 0; i < N_L - 1; ++i) {
        const double dx = PH[i + 1] - PH[i], dyv = yg[i + 1] - yg[i];
        tb[2 * N_L + i] = dyv / dx - dx * (cc[i + 1] + 2.0 * cc[i]) / 3.0;
        tb[3 * N_L + i] = cc[i];
        tb[4 * N_L + i] = (cc[i + 1] - cc[i]) / (3.0 * dx);
      }
    }
    steffen_coeffs(PH, tb + 5 * N_L, N_L, l, &b, &c, &d);
    tb[6 * N_L + l] = b; tb[7 * N_L + l] = c; tb[8 * N_L + l] = d;
    steffen_coeffs(PH, tb + 9 * N_L, N_L, l, &b, &c, &d);
    tb[10 * N_L + l] = b; tb[11 * N_L + l] = c; tb[12 * N_L + l] = d;
  }
  // ---- ring constants: the redshift range bounds the table rows these energies can reach ------
  const double Grav_z = sqrt(1.0 - a.r_s_over_r[ring]);
  const double theta_i = a.theta[ring * a.theta_ring_stride];
  const double beta = fabs(a.radial[ring] * a.omega[q] * sin(theta_i) / (kC * Grav_z));
  const double Lorentz = sqrt(1.0 - beta * beta);
  const double lzmax = log10(Lorentz / (1.0 - beta) * Grav_z) + 1.0e-6;
  const double lzmin = log10(Lorentz / (1.0 + beta) * Grav_z) - 1.0e-6;
  const double E_lane = a.energies[e0 + (lane < ne ? lane : 0)];
  const double logE_lane = a.log10_energies[e0 + (lane < ne ? lane : 0)];
  const double logE_lo = a.log10_energies[e0], logE_hi = a.log10_energies[e0 + ne - 1];
  // phases owned by this warp
  const int kb = (N_P + kGenWarps - 1) / kGenWarps;
  const int k_lo = warp * kb, k_hi = min(N_P, k_lo + kb);
  const int my_k = k_lo + lane;
  const double phk = (my_k < k_hi) ? a.phases[my_k] : 0.0;
  __syncthreads();

  double cur_T = nan(""), cur_g = nan(""), curc_T = nan(""), curc_g = nan("");
  int elo = 0, nrows = 0, elo_c = 0, nrows_c = 0;
  for (int j = grp; j < A_; j += n_groups) {
    const bool rad = a.radiates ? (a.radiates[cell0 + j] == 1) : (a.cellArea[cell0 + j] > 0.0);
    if (!rad) continue;                                  // CTA-uniform
    const double* VEC = a.srcParams + (a.params_per_cell ? (cell0 + j) : ring) * a.n_params;
    const double* CV = a.corrParams ? a.corrParams + (a.params_per_cell ? (cell0 + j) : ring) * a.n_params : nullptr;
    const double area = a.cellArea[cell0 + j], phi_j = a.phi[cell0 + j];
    const double kT = kKBOverKeV * pow(10.0, VEC[0]);
    const double log_kT = log10(kT);
    const double norm = (ATM == 2) ? (kErg / kHKeV) * pow(10.0, 3.0 * VEC[0]) : kErg * kPlanckDistConst;
    double kT_c = 1.0, log_kT_c = 0.0, norm_c = 0.0;
    if (CORR != 0) {
      kT_c = kKBOverKeV * pow(10.0, CV[0]); log_kT_c = log10(kT_c);
      norm_c = (CORR == 2) ? (kErg / kHKeV) * pow(10.0, 3.0 * CV[0]) : kErg * kPlanckDistConst;
    }
    // ---- Num4D: (re)build the slab(s) when this cell's parameters differ from the previous cell's ----
    bool rebuild = false, rebuild_c = false;
    if (ATM == 2) rebuild = !(VEC[0] == cur_T && VEC[1] == cur_g);
    if (CORR == 2) rebuild_c = !(CV[0] == curc_T && CV[1] == curc_g);
    if (rebuild || rebuild_c) {
      __syncthreads();                                   // every warp is done with the old slab
      if (rebuild) {
        cur_T = VEC[0]; cur_g = VEC[1];
        elo = lagrange_base(hot.axE, a.hot.nE, logE_lo - lzmax - log_kT - 1.0e-9);
        nrows = lagrange_base(hot.axE, a.hot.nE, logE_hi - lzmin - log_kT + 1.0e-9) + 4 - elo;
        if (nrows > a.slab_ne_max) { if (tid == 0) atomicExch(a.status + q, kUnsupported); return; }
        gen_build_slab(a.hot, hot, VEC, elo, nrows, tid);
      }
      if (rebuild_c) {
        curc_T = CV[0]; curc_g = CV[1];
        elo_c = lagrange_base(els.axE, a.els.nE, logE_lo - lzmax - log_kT_c - 1.0e-9);
        nrows_c = lagrange_base(els.axE, a.els.nE, logE_hi - lzmin - log_kT_c + 1.0e-9) + 4 - elo_c;
        if (nrows_c > a.slab_ne_max) { if (tid == 0) atomicExch(a.status + q, kUnsupported); return; }
        gen_build_slab(a.els, els, CV, elo_c, nrows_c, tid);
      }
      __syncthreads();
    }
    // ---- (image, phase) items of this cell: lane = item, then lane = energy (pyx:548-592) --------
    for (int I = 0; I < n_img; ++I) {
      const double* tb = s_tab + (long)I * 13 * N_L;
      const double* PH = tb;
      double G = 0.0, Zv = 1.0, Av = 0.0;
      int bM = 0, bMc = 0;
      double wM[4] = {0, 0, 0, 0}, wMc[4] = {0, 0, 0, 0};
      if (my_k < k_hi) {
        double x = phk + phi_j;
        const double first = PH[0], last = PH[N_L - 1];
        if (x > last) { while (x > last) x -= kTwoPi; }
        else if (x < first) { while (x < first) x += kTwoPi; }
        if (x < first || x > last) atomicExch(a.status + q, kNumericalError);      // pyx:558-563
        else {
          const int m = interval_search(PH, N_L, x);
          const double d = x - PH[m];
          G = tb[1 * N_L + m] + d * (tb[2 * N_L + m] + d * (tb[3 * N_L + m] + d * tb[4 * N_L + m]));
          if (G > 0.0) {
            Zv = tb[5 * N_L + m] + d * (tb[6 * N_L + m] + d * (tb[7 * N_L + m] + d * tb[8 * N_L + m]));
            Av = tb[9 * N_L + m] + d * (tb[10 * N_L + m] + d * (tb[11 * N_L + m] + d * tb[12 * N_L + m]));
            if (ATM == 2) {
              // beam_opt 3: a query below the table is clamped to its first node on every call (see integrate_azinv.cu)
              const double Aq = (a.beam_opt == 3 && Av < hot.axMu[0]) ? hot.axMu[0] : Av;
              bM = lagrange_base(hot.axMu, a.hot.nmu, Aq); lagrange_weights(hot.axMu, bM, Aq, wM);
            }
            if (CORR == 2) { bMc = lagrange_base(els.axMu, a.els.nmu, Av); lagrange_weights(els.axMu, bMc, Av, wMc); }
          } else G = 0.0;
        }
      }
      const unsigned lit = __ballot_sync(0xffffffffu, G > 0.0);
      for (unsigned rest = lit; rest; rest &= rest - 1) {
        const int t = __ffs(rest) - 1;
        const double g = __shfl_sync(0xffffffffu, G, t);
        const double z = __shfl_sync(0xffffffffu, Zv, t);
        const double abb = __shfl_sync(0xffffffffu, Av, t);
        int bm = 0, bmc = 0;
        double w[4], wc[4];
        if (ATM == 2) {
          bm = __shfl_sync(0xffffffffu, bM, t);
#pragma unroll
          for (int x = 0; x < 4; ++x) w[x] = __shfl_sync(0xffffffffu, wM[x], t);
        }
        if (CORR == 2) {
          bmc = __shfl_sync(0xffffffffu, bMc, t);
#pragma unroll
          for (int x = 0; x < 4; ++x) wc[x] = __shfl_sync(0xffffffffu, wMc[x], t);
        }
        if (lane >= ne) continue;
        const double Ep = E_lane / z;
        int bad = 0;
        double I_E;
        if (ATM == 1) I_E = bb_intensity(Ep, kT);
        else I_E = gen_slab_eval(hot, a.hot.nE, elo, nrows, logE_lane - log10(z) - log_kT, bm, w, &bad);
        if (a.beam_opt != 0) {            // hot_wrapper.pyx:155-199 (options 1-3), the cell's own parameters
          const double t3 = (ATM == 2) ? pow(10.0, 3.0 * VEC[0]) : 1.0;
          const double v = logE_lane - log10(z) - log_kT;
          I_E = apply_beaming(a.beam_opt, I_E * t3, Ep, abb, VEC, [&](double mu_i) -> double {
                  if (ATM != 2) return bb_intensity(Ep, kT);
                  double wi[4];
                  const int bi = lagrange_base(hot.axMu, a.hot.nmu, mu_i);
                  lagrange_weights(hot.axMu, bi, mu_i, wi);
                  return gen_slab_eval(hot, a.hot.nE, elo, nrows, v, bi, wi, &bad) * t3;
                }) / t3;
        }
        double corr = 0.0;
        if (CORR == 1) corr = bb_intensity(Ep, kT_c) * norm_c;
        else if (CORR == 2)
          corr = gen_slab_eval(els, a.els.nE, elo_c, nrows_c, logE_lane - log10(z) - log_kT_c, bmc, wc, &bad) * norm_c;
        if (bad) s_bad = 1;
        s_acc[(k_lo + t) * kGenE + lane] += area * (I_E * norm - corr) * g;       // pyx:592
      }
    }
  }
  __syncthreads();
  if (s_bad && tid == 0) atomicExch(a.status + q, kUnsupported);     // a row outside the staged slab: refuse
  double* flux_q = a.flux + (long)q * N_E * N_P;
  for (int t = tid; t < N_P * ne; t += kGenThreads) {
    const int e = t / N_P, k = t - e * N_P;
    const double v = s_acc[k * kGenE + e];
    if (v != 0.0) atomicAdd(flux_q + (long)(e0 + e) * N_P + k, v);
  }
}

static size_t general_smem_bytes(const AzinvArgs& a, int atm, int corr) {
  size_t d = (size_t)a.n_phases * kGenE + (size_t)a.n_img_max * 13 * a.n_leaves;
  if (atm == 2) d += a.hot.nE + a.hot.nmu + (size_t)a.hot.nmu * a.slab_ne_max;
  if (corr == 2) d += a.els.nE + a.els.nmu + (size_t)a.els.nmu * a.slab_ne_max;
  return d * sizeof(double);
}

// energy rows 32 consecutive energies can reach, with the Doppler spread of one ring
int general_slab_rows(const AtmTable& t, const double* energies, int n_energies) {
  if (t.min_dlogE <= 0.0) return t.nE;
  double span = 0.0;
  for (int e0 = 0; e0 < n_energies; e0 += kGenE) {
    const int e1 = (e0 + kGenE < n_energies ? e0 + kGenE : n_energies) - 1;
    const double s = log10(energies[e1] / energies[e0]);
    if (s > span) span = s;
  }
  const int rows = (int)ceil((span + 0.2) / t.min_dlogE) + 8;
  return rows > t.nE ? t.nE : rows;
}

template <int ATM, int CORR>
static cudaError_t launch_general_flux(const AzinvArgs& a, dim3 grid, size_t smem, int n_groups, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(k_general_flux<ATM, CORR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  k_general_flux<ATM, CORR><<<grid, kGenThreads, smem, stream>>>(a, n_groups);
  return cudaGetLastError();
}

__global__ void k_general_scale(double* flux, const double* energies, int Q, int N_E, int N_P) {
  const long n = (long)Q * N_E * N_P;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    const int e = (int)((t / N_P) % N_E);
    flux[t] = flux[t] / (energies[e] * kKeV);            // pyx:607-609
  }
}

cudaError_t launch_integrate_general(AzinvArgs a, cudaStream_t stream) {
  if (a.n_phases > 32 * kGenWarps) return cudaErrorInvalidValue;
  if (a.phase_interp == kCubic && a.n_leaves > kMaxCubicNodes) return cudaErrorInvalidValue;
  if (a.n_img_max > kMaxImages || a.n_img_max < 1) return cudaErrorInvalidValue;
  if (!a.ws_leaf || !a.ws_ihdr || !a.ws_hdr || !a.log10_energies) return cudaErrorInvalidValue;
  const int atm = a.hot_atm_ext;
  const int corr = a.corrParams ? a.else_atm_ext : 0;
  if (atm != 1 && atm != 2) return cudaErrorNotSupported;
  if (corr != 0 && corr != 1 && corr != 2) return cudaErrorNotSupported;
  if (a.beam_opt < 0 || a.beam_opt > 3 || (a.beam_opt != 0 && a.n_params < 7)) return cudaErrorNotSupported;
  if ((atm == 2 || corr == 2) && a.slab_ne_max < 4) return cudaErrorInvalidValue;
  if (!a.corrParams) a.else_atm_ext = 0;
  a.general = 1;
  a.R_in = 1.0e6;                       // integrator.pyx takes R_in but never applies the disc
  a.work = nullptr;
  const size_t fsm = general_smem_bytes(a, atm, corr);
  if (fsm > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t err;
  if ((err = launch_azinv_geometry(a, stream)) != cudaSuccess) return err;
  const int n_chunks = (a.n_energies + kGenE - 1) / kGenE;
  // split a ring's cells over CTAs until the grid covers the 148 SMs a few times
  long base = (long)a.n_rings * n_chunks * a.Q;
  int n_groups = (int)((148L * 8 + base - 1) / base);
  if (n_groups < 1) n_groups = 1;
  if (n_groups > a.n_azi) n_groups = a.n_azi;
  dim3 grid((unsigned)(a.n_rings * n_chunks * n_groups), a.Q);
  if (atm == 1 && corr == 0) err = launch_general_flux<1, 0>(a, grid, fsm, n_groups, stream);
  else if (atm == 1 && corr == 1) err = launch_general_flux<1, 1>(a, grid, fsm, n_groups, stream);
  else if (atm == 1 && corr == 2) err = launch_general_flux<1, 2>(a, grid, fsm, n_groups, stream);
  else if (atm == 2 && corr == 0) err = launch_general_flux<2, 0>(a, grid, fsm, n_groups, stream);
  else if (atm == 2 && corr == 1) err = launch_general_flux<2, 1>(a, grid, fsm, n_groups, stream);
  else err = launch_general_flux<2, 2>(a, grid, fsm, n_groups, stream);
  if (err != cudaSuccess) return err;
  if (a.scale_by_energy) {
    const long n = (long)a.Q * a.n_energies * a.n_phases;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_general_scale<<<blocks, 256, 0, stream>>>(a.flux, a.energies, a.Q, a.n_energies, a.n_phases);
    err = cudaGetLastError();
  }
  return err;
}

}  // namespace xb
