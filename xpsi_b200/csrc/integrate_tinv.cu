// Time-invariant integrator over a closed mesh (Elsewhere / Everywhere time_invariant).
//
// Replaces xpsi/cellmesh/integrator_for_time_invariance.pyx:59-338 and the atmosphere
// dispatch of elsewhere_wrapper.pyx:23-81 (atm_ext 1 blackbody, 2 Num4D).
//
// One CTA per (instance q, mesh ring); a warp takes one azimuthal cell at a time: the
// cell's geometry (psi, cos alpha and its derivative through Steffen pieces rebuilt in
// registers, Doppler, redshift, Jacobian) is evaluated once per image order, then the
// lanes split the energies and keep their partial sums in registers for every cell the
// warp visits -- no atomics until the ring is finished.  Num4D: when all cells of the
// ring share (log T, log g) -- the Elsewhere/Everywhere default -- the table is
// contracted once per ring into a (mu, E-row) slab in shared memory; otherwise each
// cell contracts the 4-D stencil straight from the L2-resident table.
#include "common.cuh"
#include "kernels.h"

namespace xb {

constexpr int kTinvThreads = 128;
constexpr int kTinvEPerLane = 8;         // supports up to 256 energies

template <class P>
__device__ __forceinline__ void tinv_row_range(const P& axis, int nE, double vlo, double vhi, int* elo, int* ehi) {
  *elo = lagrange_base(axis, nE, vlo - 1.0e-9);
  *ehi = lagrange_base(axis, nE, vhi + 1.0e-9) + 4;
}

template <int ATM>
__global__ void __launch_bounds__(kTinvThreads) k_tinv(TinvArgs a) {
  const int i = blockIdx.x, q = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.sqrt_numPix, N_R = a.n_rays, N_E = a.n_energies;
  const long ring = (long)q * n + i;
  const long cell0 = ring * n;

  extern __shared__ double smem[];
  double* sp = smem;
  double* s_defl = sp; sp += N_R;
  double* s_calpha = sp; sp += N_R;
  double* s_cosd = sp; sp += N_R;
  double* s_E = sp; sp += N_E;
  double* s_logE = sp; sp += N_E;
  double* s_red = sp; sp += (kTinvThreads / 32) * N_E;
  double* s_axE = nullptr; double* s_axMu = nullptr; double* s_slab = nullptr;
  if (ATM == 2) { s_axE = sp; sp += a.atm.nE; s_axMu = sp; sp += a.atm.nmu; s_slab = sp; }
  __shared__ int s_jhalf, s_uniform, s_elo, s_nrows, s_bT, s_bG, s_fail;
  __shared__ double s_wT[4], s_wG[4];

  if (tid == 0) { s_jhalf = N_R - 1; s_uniform = 1; s_fail = 0; }
  __syncthreads();
  const double* g_defl = a.deflection + ring * N_R;
  const double* g_calpha = a.cos_alpha + ring * N_R;
  for (int r = tid; r < N_R; r += kTinvThreads) {
    const double d = g_defl[r];
    s_defl[r] = d; s_calpha[r] = g_calpha[r];
    s_cosd[r] = cos(d);                                   // pyx:136-140
    if (d > kHalfPi) atomicMin(&s_jhalf, r);              // pyx:187-189
  }
  for (int e = tid; e < N_E; e += kTinvThreads) { s_E[e] = a.energies[e]; s_logE[e] = log10(a.energies[e]); }
  const double* P0 = a.srcParams + cell0 * a.n_params;
  for (int j = tid; j < n; j += kTinvThreads) {
    const double* Pj = a.srcParams + (cell0 + j) * a.n_params;
    if (Pj[0] != P0[0] || (ATM == 2 && Pj[1] != P0[1])) s_uniform = 0;
  }
  if (ATM == 2) {
    for (int e = tid; e < a.atm.nE; e += kTinvThreads) s_axE[e] = a.atm.logE[e];
    for (int m = tid; m < a.atm.nmu; m += kTinvThreads) s_axMu[m] = a.atm.mu[m];
  }

  // ---- ring constants (pyx:199-211) -------------------------------------------------------
  const double inclination = a.inclination[q], omega = a.omega[q];
  const double sin_i = sin(inclination), cos_i = cos(inclination);
  const double radius = a.radial[ring];
  const double Grav_z = sqrt(1.0 - a.r_s_over_r[ring]);
  const double cos_gamma = a.cos_gamma[ring];
  const double sin_gamma = sqrt(1.0 - cos_gamma * cos_gamma);
  const double theta_i = a.theta[cell0];
  const double cos_theta_i = cos(theta_i), sin_theta_i = sin(theta_i);
  const double theta_i_over_pi = theta_i / kPi;
  const double beta = radius * omega * sin_theta_i / (kC * Grav_z);
  const double Lorentz = sqrt(1.0 - beta * beta);
  const double maxDefl = a.maxDeflection[ring];
  const int _IO = a.image_order_limit > 0 ? a.image_order_limit : (int)ceil(maxDefl / kPi);
  __syncthreads();
  const int jhalf = s_jhalf;
  const bool uniform = s_uniform != 0;

  // ---- Num4D, ring with a single (T,g): contract the slab once ---------------------------------
  if (ATM == 2 && uniform) {
    if (tid == 0) {
      View vT{a.atm.logT, 1}, vG{a.atm.logg, 1};
      s_bT = lagrange_base(vT, a.atm.nT, P0[0]);
      s_bG = lagrange_base(vG, a.atm.ng, P0[1]);
      double w[4];
      lagrange_weights(vT, s_bT, P0[0], w);
      for (int x = 0; x < 4; ++x) s_wT[x] = w[x];
      lagrange_weights(vG, s_bG, P0[1], w);
      for (int x = 0; x < 4; ++x) s_wG[x] = w[x];
      const double log_kT = log10(kKBOverKeV * pow(10.0, P0[0]));
      const double Zmax = Lorentz / (1.0 - fabs(beta)) * Grav_z, Zmin = Lorentz / (1.0 + fabs(beta)) * Grav_z;
      int elo, ehi;
      tinv_row_range(s_axE, a.atm.nE, s_logE[0] - log10(Zmax) - log_kT, s_logE[N_E - 1] - log10(Zmin) - log_kT,
                     &elo, &ehi);
      if (ehi - elo > a.slab_rows) { atomicExch(a.status + q, kUnsupported); s_fail = 1; }
      s_elo = elo; s_nrows = ehi - elo;
    }
    __syncthreads();
    if (s_fail) return;
    const int elo = s_elo, nrows = s_nrows, nmu = a.atm.nmu;
    const long S0 = (long)a.atm.ng * nmu * a.atm.nE, S1 = (long)nmu * a.atm.nE, S2 = a.atm.nE;
    for (int m = warp; m < nmu; m += kTinvThreads / 32)
      for (int e = lane; e < nrows; e += 32) {
        const double* base = a.atm.buf + (long)s_bT * S0 + (long)s_bG * S1 + (long)m * S2 + elo + e;
        double acc = 0.0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          double inner = 0.0;
#pragma unroll
          for (int y = 0; y < 4; ++y) inner += s_wG[y] * __ldg(base + x * S0 + y * S1);
          acc += s_wT[x] * inner;
        }
        s_slab[m * nrows + e] = acc;
      }
    __syncthreads();
  }

  View vDefl{s_defl, 1}, vCalpha{s_calpha, 1};
  View vAltX{s_cosd + jhalf, -1}, vAltY{s_calpha + jhalf, -1};   // pyx:190-194
  const int n_alt = jhalf + 1;
  const double alt_xmin = s_cosd[jhalf];
  double acc[kTinvEPerLane];
#pragma unroll
  for (int g = 0; g < kTinvEPerLane; ++g) acc[g] = 0.0;

  // ---- cells of the ring: one warp per cell (pyx:221-303) ------------------------------------------
  for (int j = warp; j < n; j += kTinvThreads / 32) {
    const double phi_j = a.phi[cell0 + j];
    const double* VEC = a.srcParams + (cell0 + j) * a.n_params;
    const double _cos_psi = cos_i * cos_theta_i + sin_i * sin_theta_i * cos(phi_j);
    const double _psi = acos(_cos_psi);
    const double kT = kKBOverKeV * pow(10.0, VEC[0]);
    const double log_kT = log10(kT);
    for (int I = 0; I < _IO; ++I) {
      double cos_psi = _cos_psi;
      double psi = eval_image_deflection(I, _psi);
      double sin_psi = sin(psi);
      if (!are_equal(psi, 0.0) && are_equal(sin_psi, 0.0)) {      // pole singularity nudge
        const double _i = cos_i >= 0.0 ? inclination + inclination * 1.0e-6 : inclination - inclination * 1.0e-6;
        cos_psi = cos(_i) * cos_theta_i + sin(_i) * sin_theta_i * cos(phi_j);
        psi = eval_image_deflection(I, acos(cos_psi));
        sin_psi = sin(psi);
      }
      if (psi > maxDefl) break;                                   // higher orders not visible
      if (psi < s_defl[0] || psi > s_defl[N_R - 1]) { if (lane == 0) atomicExch(a.status + q, kNumericalError); break; }
      const bool use_alt = (psi <= kHalfPi) && (cos_psi >= alt_xmin);
      double cos_alpha, deriv;
      if (use_alt) {
        const int idx = interval_search(vAltX, n_alt, cos_psi);
        steffen_eval(vAltX, vAltY, n_alt, idx, cos_psi, &cos_alpha, &deriv);
      } else {
        const int idx = interval_search(vDefl, N_R, psi);
        steffen_eval(vDefl, vCalpha, N_R, idx, psi, &cos_alpha, &deriv);
        deriv = exp(log(fabs(deriv)) - log(fabs(sin_psi)));
      }
      const double sin_alpha = sqrt(1.0 - cos_alpha * cos_alpha);
      double mu = cos_alpha * cos_gamma;
      if (!are_equal(psi, 0.0)) {
        const double cos_delta = (cos_i - cos_theta_i * cos_psi) / (sin_theta_i * sin_psi);
        if (theta_i_over_pi < 0.5) mu += sin_alpha * sin_gamma * cos_delta;
        else mu -= sin_alpha * sin_gamma * cos_delta;
      }
      if (!(mu > 0.0)) continue;
      double eta;
      if (!are_equal(sin_psi, 0.0)) {
        const double cos_xi = sin_alpha * sin_i * sin(phi_j) / sin_psi;
        eta = Lorentz / (1.0 + beta * cos_xi);
      } else eta = Lorentz;
      const double Z = eta * Grav_z, ABB = mu * eta;
      const double GEOM = mu * fabs(deriv) * Grav_z * eta * eta * eta;       // no 1/superlum (pyx:291)
      if (ATM == 1) {
#pragma unroll
        for (int g = 0; g < kTinvEPerLane; ++g) {
          const int e = lane + 32 * g;
          if (e < N_E) { const double Ep = s_E[e] / Z; acc[g] += Ep * Ep * Ep / (exp(Ep / kT) - 1.0) * GEOM; }
        }
      } else {
        const double logZ = log10(Z);
        const int bM = lagrange_base(s_axMu, a.atm.nmu, ABB);
        double wM[4];
        lagrange_weights(s_axMu, bM, ABB, wM);
        const double t3 = pow(10.0, 3.0 * VEC[0]);
        int bT = 0, bG = 0;
        double wT[4], wG[4];
        if (!uniform) {
          View vT{a.atm.logT, 1}, vG{a.atm.logg, 1};
          bT = lagrange_base(vT, a.atm.nT, VEC[0]); lagrange_weights(vT, bT, VEC[0], wT);
          bG = lagrange_base(vG, a.atm.ng, VEC[1]); lagrange_weights(vG, bG, VEC[1], wG);
        }
#pragma unroll
        for (int g = 0; g < kTinvEPerLane; ++g) {
          const int e = lane + 32 * g;
          if (e >= N_E) continue;
          const double v = s_logE[e] - logZ - log_kT;                         // log10(E'/kT)
          const int bE = lagrange_base(s_axE, a.atm.nE, v);
          double wE[4];
          lagrange_weights(s_axE, bE, v, wE);
          double sum = 0.0;
          if (uniform) {
            const double* row = s_slab + (long)bM * s_nrows + (bE - s_elo);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const double* r = row + x * s_nrows;
              sum += wM[x] * (wE[0] * r[0] + wE[1] * r[1] + wE[2] * r[2] + wE[3] * r[3]);
            }
          } else {
            const long S0 = (long)a.atm.ng * a.atm.nmu * a.atm.nE, S1 = (long)a.atm.nmu * a.atm.nE, S2 = a.atm.nE;
            for (int x = 0; x < 4; ++x)
              for (int y = 0; y < 4; ++y) {
                const double wxy = wT[x] * wG[y];
                const double* base = a.atm.buf + (long)(bT + x) * S0 + (long)(bG + y) * S1 + (long)bM * S2 + bE;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                  const double* r = base + m * S2;
                  sum += wxy * wM[m] * (wE[0] * __ldg(r) + wE[1] * __ldg(r + 1) + wE[2] * __ldg(r + 2) + wE[3] * __ldg(r + 3));
                }
              }
          }
          if (sum < 0.0) sum = 0.0;                                            // hot_Num4D.pyx:436-437
          acc[g] += sum * t3 * GEOM;
        }
      }
    }
  }
  // ---- ring sum over the CTA's warps, then one RED per energy ----------------------------------------
#pragma unroll
  for (int g = 0; g < kTinvEPerLane; ++g) {
    const int e = lane + 32 * g;
    if (e < N_E) s_red[warp * N_E + e] = acc[g];
  }
  __syncthreads();
  for (int e = tid; e < N_E; e += kTinvThreads) {
    double s = 0.0;
    for (int w = 0; w < kTinvThreads / 32; ++w) s += s_red[w * N_E + e];
    if (s != 0.0) atomicAdd(a.flux + (long)q * N_E + e, s);
  }
}

// flux[q][e] *= cellArea * norm / (E keV)   (pyx:317-318)
__global__ void k_tinv_scale(TinvArgs a, double norm) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= (long)a.Q * a.n_energies) return;
  const int q = (int)(t / a.n_energies), e = (int)(t - (long)q * a.n_energies);
  a.flux[t] = a.flux[t] * (a.cellArea[q] * norm / (a.energies[e] * kKeV));
}

cudaError_t launch_integrate_tinv(TinvArgs a, cudaStream_t stream) {
  if (a.n_energies > 32 * kTinvEPerLane) return cudaErrorInvalidValue;
  if (a.atm_ext != 1 && a.atm_ext != 2) return cudaErrorNotSupported;
  size_t d = 3ul * a.n_rays + 2ul * a.n_energies + (size_t)(kTinvThreads / 32) * a.n_energies;
  if (a.atm_ext == 2) d += a.atm.nE + a.atm.nmu + (size_t)a.atm.nmu * a.slab_rows;
  const size_t smem = d * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  dim3 grid(a.sqrt_numPix, a.Q);
  cudaError_t err;
  if (a.atm_ext == 1) {
    if ((err = cudaFuncSetAttribute(k_tinv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
    k_tinv<1><<<grid, kTinvThreads, smem, stream>>>(a);
  } else {
    if ((err = cudaFuncSetAttribute(k_tinv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
    k_tinv<2><<<grid, kTinvThreads, smem, stream>>>(a);
  }
  if ((err = cudaGetLastError()) != cudaSuccess) return err;
  const double norm = (a.atm_ext == 2) ? kErg / kHKeV : kErg * kPlanckDistConst;   // elsewhere_wrapper.pyx:62-81
  const long nt = (long)a.Q * a.n_energies;
  k_tinv_scale<<<(unsigned)((nt + 127) / 128), 128, 0, stream>>>(a, norm);
  return cudaGetLastError();
}

int tinv_slab_rows(const AtmTable& t, const double* energies, int n_energies) {
  if (t.min_dlogE <= 0.0) return t.nE;
  const int rows = (int)ceil((log10(energies[n_energies - 1] / energies[0]) + 0.42) / t.min_dlogE) + 8;
  return rows > t.nE ? t.nE : rows;
}

}  // namespace xb
