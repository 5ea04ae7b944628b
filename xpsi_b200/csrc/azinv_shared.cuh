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

// geometry stage alone (used by the general integrator, which brings its own flux kernel)
cudaError_t launch_azinv_geometry(const AzinvArgs& a, cudaStream_t stream);

}  // namespace xb
