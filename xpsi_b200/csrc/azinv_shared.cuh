// Pieces shared by the pulse integrators (integrate_azinv.cu, integrate_general.cu): the leaf
// workspace written by k_azinv_geometry and the per-ring headers.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace xb {

__device__ __forceinline__ double bb_intensity(double E, double kT) {
  return E * E * E / (exp(E / kT) - 1.0);     // hot_BB.pyx:85-87
}

// per-ring headers written by the geometry kernel
//   ints   [0] image orders to integrate  [1] first radiating cell  [2],[3] (T,g) base nodes
//          [4],[5] first row / row count of the ring's slab (written by k_azinv_slab)
//   doubles [0],[1] min/max of Z (log10 Z for Num4D) over lit leaves  [2..5] T weights
//          [6..9] g weights  [10] log10 T  [11] log10 g  [12] kT (keV)  [13] log10 kT
//          [14] intensity normalisation (hot_BB.pyx:98 / hot_Num4D.pyx:436-460)
//   elsewhere correction (pyx:257-268): the same block of doubles again at +kCorrD for the ring's
//   correction parameters; ints [6],[7] its (T,g) base nodes, [8],[9] its slab rows
constexpr int kIHdr = 12, kDHdr = 32, kCorrD = 16;

// leaf workspace layout: [ring][image][4][N_L]
__device__ __forceinline__ double* leaf_ptr(double* ws, long ring, int n_img_max, int I, int N_L) {
  return ws + ((ring * n_img_max + I) * 4) * (long)N_L;
}

// beaming modification of the hot intensity (hot_wrapper.pyx:155-199).  I_hot and eval_mu(mu') are in the
// reference's eval_hot units (Num4D: including the 10^(3 log T) factor) because option 3 compares its
// normalisation integral with zero at an absolute 1e-12 (tools/core.pyx:118-122).
template <class EvalMu>
__device__ __forceinline__ double apply_beaming(int beam_opt, double I_hot, double Ep, double mu, const double* BV,
                                                EvalMu&& eval_mu) {
  const double ab = BV[2], bb = BV[3], cb = BV[4], db = BV[5];
  const double Ec = pow(Ep, cb), Ed = pow(Ep, db);
  const double f = 1.0 + ab * Ec * mu + bb * Ed * mu * mu;
  double I = 0.0;
  if (beam_opt == 1) I = f * I_hot;
  else if (beam_opt == 2) I = 0.5 / (0.5 + (1.0 / 3.0) * ab * Ec + (1.0 / 4.0) * bb * Ed) * f * I_hot;
  else if (beam_opt == 3) {              // trapezoid over mu, :173-192
    const double nimu = BV[6];
    const long n = (long)nimu;
    double mu_i = 0.0, nom = 0.0, den = 0.0;
    for (long im = 0; im < n; ++im) {
      mu_i = mu_i + (1.0 / nimu);
      const double dmu = (im == 0 || (double)im == nimu - 1) ? (0.5 / nimu) : (1.0 / nimu);
      const double Ii = eval_mu(mu_i);
      const double fi = 1.0 + ab * Ec * mu_i + bb * Ed * mu_i * mu_i;
      den = den + mu_i * fi * Ii * dmu;
      nom = nom + mu_i * Ii * dmu;
    }
    I = are_equal(den, 0.0) ? 0.0 : (nom / den) * f * I_hot;
  }
  return I < 0.0 ? 0.0 : I;
}

// geometry stage alone (used by the general integrator, which brings its own flux kernel)
cudaError_t launch_azinv_geometry(const AzinvArgs& a, cudaStream_t stream);

}  // namespace xb
