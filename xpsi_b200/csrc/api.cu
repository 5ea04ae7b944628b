// C ABI of the B200-native X-PSI likelihood hot path (include/xpsi_b200.h).
// Host-side plumbing only: device buffers, one stream, handles for the
// theta-independent constants, and the batched pipeline that chains the kernels.
#include "../../include/xpsi_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "kernels.h"

namespace {

thread_local std::string g_err;
cudaStream_t g_stream = nullptr;
long long g_launches = 0, g_h2d = 0, g_d2h = 0;

int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? XPSI_B200_ENODEVICE : XPSI_B200_ECUDA;
}
#define CK(call)                                              \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return cuda_fail(_e, #call);       \
  } while (0)

int ensure_stream() {
  if (g_stream) return 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_err = "xpsi_b200: no usable CUDA device (the hot path has no CPU fallback)";
    return XPSI_B200_ENODEVICE;
  }
  CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  return 0;
}

// device buffer with explicit lifetime
template <class T>
struct Dev {
  T* p = nullptr;
  size_t n = 0;
  Dev() {}
  Dev(const Dev&) = delete;
  Dev& operator=(const Dev&) = delete;
  ~Dev() { if (p) cudaFree(p); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; }
    n = count;
    return cudaMalloc(&p, (count ? count : 1) * sizeof(T));
  }
  cudaError_t upload(const T* h, size_t count) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess) return e;
    g_h2d += (long long)(count * sizeof(T));
    return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, g_stream);
  }
  cudaError_t download(T* h, size_t count) const {
    g_d2h += (long long)(count * sizeof(T));
    return cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, g_stream);
  }
};

// per channel tile, the span of input intervals holding any non-zero response entry
std::vector<int> response_k_ranges(const double* m, int n_chan, int ld, int in0, int n_in) {
  const int T = xb::fold_tile_rows();
  const int nt = (n_chan + T - 1) / T;
  std::vector<int> r(2 * nt);
  for (int t = 0; t < nt; ++t) {
    int lo = n_in, hi = 0;
    for (int c = t * T; c < n_chan && c < (t + 1) * T; ++c) {
      const double* row = m + (size_t)c * ld + in0;
      int a = 0, b = n_in;
      while (a < n_in && row[a] == 0.0) ++a;
      while (b > a && row[b - 1] == 0.0) --b;
      if (a < lo) lo = a;
      if (b > hi) hi = b;
    }
    if (hi < lo) { lo = 0; hi = 0; }
    r[2 * t] = lo; r[2 * t + 1] = hi;
  }
  return r;
}

}  // namespace

struct xpsi_b200_atmosphere {
  Dev<double> logT, logg, mu, logE, buf, mu_invden, E_invden;
  xb::AtmTable view;
};

extern "C" {

const char* xpsi_b200_last_error(void) { return g_err.c_str(); }

int xpsi_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int xpsi_b200_set_device(int device) {
  CK(cudaSetDevice(device));
  return 0;
}

void xpsi_b200_counters(long long* k, long long* h2d, long long* d2h) {
  if (k) *k = g_launches;
  if (h2d) *h2d = g_h2d;
  if (d2h) *d2h = g_d2h;
}

int xpsi_b200_fp64_peak_tflops(double* tflops) {
  int rc = ensure_stream();
  if (rc) return rc;
  cudaError_t e = xb::measure_fp64_peak(tflops, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "measure_fp64_peak");
  g_launches += 6;
  return 0;
}

void* xpsi_b200_stream(void) { return ensure_stream() == 0 ? (void*)g_stream : nullptr; }

xpsi_b200_atmosphere* xpsi_b200_atmosphere_create(const double* logT, int nT, const double* logg,
                                                  int ng, const double* mu, int nmu,
                                                  const double* logE, int nE, const double* buf) {
  if (ensure_stream() != 0) return nullptr;
  if (nT < 4 || ng < 4 || nmu < 4 || nE < 4) { g_err = "atmosphere axes need >= 4 nodes"; return nullptr; }
  xpsi_b200_atmosphere* a = new xpsi_b200_atmosphere();
  cudaError_t e = a->logT.upload(logT, nT);
  if (e == cudaSuccess) e = a->logg.upload(logg, ng);
  if (e == cudaSuccess) e = a->mu.upload(mu, nmu);
  if (e == cudaSuccess) e = a->logE.upload(logE, nE);
  if (e == cudaSuccess) e = a->buf.upload(buf, (size_t)nT * ng * nmu * nE);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
  if (e != cudaSuccess) { cuda_fail(e, "atmosphere_create"); delete a; return nullptr; }
  double dmin = 1e300;
  for (int i = 1; i < nE; ++i) dmin = fmin(dmin, logE[i] - logE[i - 1]);
  // inverse Lagrange denominators per base node of the mu and E axes (hot_Num4D.pyx:386-409 "SPACE"), so the
  // kernels' stencil weights need no divisions
  auto invden = [](const double* p, int n) {
    std::vector<double> v((size_t)(n - 3) * 4);
    for (int b = 0; b + 3 < n; ++b) {
      const double p0 = p[b], p1 = p[b + 1], p2 = p[b + 2], p3 = p[b + 3];
      v[4 * b + 0] = 1.0 / (p0 - p1) / (p0 - p2) / (p0 - p3);
      v[4 * b + 1] = 1.0 / (p1 - p0) / (p1 - p2) / (p1 - p3);
      v[4 * b + 2] = 1.0 / (p2 - p0) / (p2 - p1) / (p2 - p3);
      v[4 * b + 3] = 1.0 / (p3 - p0) / (p3 - p1) / (p3 - p2);
    }
    return v;
  };
  std::vector<double> im = invden(mu, nmu), ie = invden(logE, nE);
  e = a->mu_invden.upload(im.data(), im.size());
  if (e == cudaSuccess) e = a->E_invden.upload(ie.data(), ie.size());
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
  if (e != cudaSuccess) { cuda_fail(e, "atmosphere_create"); delete a; return nullptr; }
  a->view = xb::AtmTable{a->logT.p, a->logg.p, a->mu.p, a->logE.p, a->buf.p, nT, ng, nmu, nE, dmin,
                         a->mu_invden.p, a->E_invden.p};
  return a;
}

void xpsi_b200_atmosphere_destroy(xpsi_b200_atmosphere* atm) { delete atm; }

static int integrate_member(
    int general,
    double R, double omega, double r_s, double inclination, int n_rings, int n_azi,
    const double* cellArea, const double* radial, const double* r_s_over_r, const double* theta,
    const double* phi, const double* srcCellParams, int n_params, const int* CELL_RADIATES,
    const double* correction_srcCellParams, int numRays, const double* deflection,
    const double* cos_alpha, const double* lag, const double* maxDeflection,
    const double* cos_gammaArray, int n_energies, const double* energies, int n_leaves,
    const double* leaves, int n_phases, const double* phases,
    const xpsi_b200_atmosphere* hot_atmosphere, const xpsi_b200_atmosphere* elsewhere_atmosphere,
    int hot_atm_ext, int else_atm_ext, int beam_opt, int image_order_limit, double R_in,
    int phase_interpolant, double* flux_out) {
  (void)R; (void)r_s;
  int rc = ensure_stream();
  if (rc) return rc;
  if (correction_srcCellParams) {
    if (else_atm_ext != XPSI_B200_ATM_BB && else_atm_ext != XPSI_B200_ATM_NUM4D)
      return fail(XPSI_B200_EUNSUPPORTED, "else_atm_ext must be 1 (BB) or 2 (Num4D)");
    if (else_atm_ext == XPSI_B200_ATM_NUM4D && !elsewhere_atmosphere)
      return fail(XPSI_B200_EINVAL, "Num4D elsewhere correction needs a preloaded atmosphere");
  }
  if (beam_opt < 0 || beam_opt > 3) return fail(XPSI_B200_EINVAL, "beam_opt must be 0-3");
  if (beam_opt != 0 && n_params < 7) return fail(XPSI_B200_EINVAL, "beam_opt needs srcCellParams[..., 2:7]");
  if (hot_atm_ext != XPSI_B200_ATM_BB && hot_atm_ext != XPSI_B200_ATM_NUM4D)
    return fail(XPSI_B200_EUNSUPPORTED, "hot_atm_ext must be 1 (BB) or 2 (Num4D)");
  if (hot_atm_ext == XPSI_B200_ATM_NUM4D && !hot_atmosphere)
    return fail(XPSI_B200_EINVAL, "Num4D needs a preloaded atmosphere");
  if (phase_interpolant < 0 || phase_interpolant > 2)
    return fail(XPSI_B200_EINVAL, "phase interpolant must be Akima (0), Steffen (1) or Cubic (2)");
  if (phase_interpolant == 2 && n_leaves > 260) return fail(XPSI_B200_EUNSUPPORTED, "Cubic interpolant: at most 260 leaves");
  if (n_rings < 1 || n_azi < 1 || numRays < 3 || n_energies < 1 || n_leaves < 5 || n_phases < 1 || n_params < 1)
    return fail(XPSI_B200_EINVAL, "bad dimensions");

  const size_t nc = (size_t)n_rings * n_azi, nr = (size_t)n_rings * numRays;
  Dev<double> d_area, d_radial, d_rsr, d_theta, d_phi, d_par, d_defl, d_ca, d_lag, d_maxd, d_cg, d_E, d_L,
      d_P, d_flux, d_scal;
  Dev<int> d_rad, d_status;
  Dev<double> d_corr;
  if (correction_srcCellParams) CK(d_corr.upload(correction_srcCellParams, nc * n_params));
  const double scal[2] = {omega, inclination};
  CK(d_scal.upload(scal, 2));
  CK(d_area.upload(cellArea, nc)); CK(d_radial.upload(radial, n_rings)); CK(d_rsr.upload(r_s_over_r, n_rings));
  CK(d_theta.upload(theta, nc)); CK(d_phi.upload(phi, nc)); CK(d_par.upload(srcCellParams, nc * n_params));
  CK(d_rad.upload(CELL_RADIATES, nc));
  CK(d_defl.upload(deflection, nr)); CK(d_ca.upload(cos_alpha, nr)); CK(d_lag.upload(lag, nr));
  CK(d_maxd.upload(maxDeflection, n_rings)); CK(d_cg.upload(cos_gammaArray, n_rings));
  CK(d_E.upload(energies, n_energies)); CK(d_L.upload(leaves, n_leaves)); CK(d_P.upload(phases, n_phases));
  std::vector<double> l10E(n_energies);
  for (int i = 0; i < n_energies; ++i) l10E[i] = log10(energies[i]);
  Dev<double> d_l10E;
  CK(d_l10E.upload(l10E.data(), n_energies));
  CK(d_flux.alloc((size_t)n_energies * n_phases));
  CK(cudaMemsetAsync(d_flux.p, 0, (size_t)n_energies * n_phases * sizeof(double), g_stream));
  CK(d_status.alloc(1));
  CK(cudaMemsetAsync(d_status.p, 0, sizeof(int), g_stream));

  xb::AzinvArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = 1; a.n_rings = n_rings; a.n_azi = n_azi; a.n_rays = numRays; a.n_energies = n_energies;
  a.n_leaves = n_leaves; a.n_phases = n_phases; a.n_params = n_params;
  a.omega = d_scal.p; a.inclination = d_scal.p + 1;
  a.cellArea = d_area.p; a.phi = d_phi.p; a.theta = d_theta.p; a.theta_ring_stride = n_azi;
  a.radial = d_radial.p; a.r_s_over_r = d_rsr.p; a.srcParams = d_par.p; a.params_per_cell = 1;
  a.radiates = d_rad.p; a.deflection = d_defl.p; a.cos_alpha = d_ca.p; a.lag = d_lag.p;
  a.maxDeflection = d_maxd.p; a.cos_gamma = d_cg.p; a.energies = d_E.p; a.leaves = d_L.p; a.phases = d_P.p;
  a.log10_energies = d_l10E.p;
  a.hot_atm_ext = hot_atm_ext;
  if (hot_atm_ext == XPSI_B200_ATM_NUM4D) {
    a.hot = hot_atmosphere->view;
    xb::azinv_slab_budgets(a.hot, energies, n_energies, &a.slab_ne_max, &a.slab_rows_ring);
    if (general) a.slab_ne_max = xb::general_slab_rows(a.hot, energies, n_energies);
  }
  if (correction_srcCellParams) {
    a.corrParams = d_corr.p; a.else_atm_ext = else_atm_ext;
    if (else_atm_ext == XPSI_B200_ATM_NUM4D) {
      a.els = elsewhere_atmosphere->view;
      int rc2 = 0, rr2 = 0;
      xb::azinv_slab_budgets(a.els, energies, n_energies, &rc2, &rr2);
      if (general) rc2 = xb::general_slab_rows(a.els, energies, n_energies);
      if (rc2 > a.slab_ne_max) a.slab_ne_max = rc2;
      if (rr2 > a.slab_rows_ring) a.slab_rows_ring = rr2;
    }
  }
  a.image_order_limit = image_order_limit > 0 ? image_order_limit : 0;
  a.n_img_max = image_order_limit > 0 ? image_order_limit : xb::kMaxImages;
  if (a.n_img_max > xb::kMaxImages) return fail(XPSI_B200_EUNSUPPORTED, "image_order_limit > 6");
  a.beam_opt = beam_opt; a.R_in = R_in;
  a.phase_interp = phase_interpolant;
  a.scale_by_energy = 1;
  a.flux = d_flux.p; a.status = d_status.p;
  Dev<double> d_ws, d_wh, d_wslab, d_wslab2, d_wcells; Dev<int> d_wi;
  {
    size_t nl, nh, ni, ns;
    xb::azinv_workspace_sizes(a, &nl, &nh, &ni, &ns);
    CK(d_ws.alloc(nl)); CK(d_wh.alloc(nh)); CK(d_wi.alloc(ni));
    CK(cudaMemsetAsync(d_wi.p, 0, ni * sizeof(int), g_stream));
    if (!general) {
      CK(d_wcells.alloc(2ul * n_rings * n_azi)); a.ws_cells = d_wcells.p;
      CK(d_wslab.alloc(ns));
      if (a.else_atm_ext == XPSI_B200_ATM_NUM4D) CK(d_wslab2.alloc(ns));
    }
  }
  a.ws_leaf = d_ws.p; a.ws_hdr = d_wh.p; a.ws_ihdr = d_wi.p; a.ws_slab = d_wslab.p; a.ws_slab2 = d_wslab2.p;
  cudaError_t e = general ? xb::launch_integrate_general(a, g_stream) : xb::launch_integrate_azinv(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, general ? "launch_integrate_general" : "launch_integrate_azinv");
  g_launches += general ? 3 : ((hot_atm_ext == XPSI_B200_ATM_NUM4D) ? 6 : 4);
  int status = 0;
  CK(d_flux.download(flux_out, (size_t)n_energies * n_phases));
  CK(d_status.download(&status, 1));
  CK(cudaStreamSynchronize(g_stream));
  if (status != 0) return fail(status, status == 1 ? "numerical error in pulse integration" : "unsupported configuration");
  return 0;
}

int xpsi_b200_integrate_azimuthal_invariance(
    double R, double omega, double r_s, double inclination, int n_rings, int n_azi,
    const double* cellArea, const double* radial, const double* r_s_over_r, const double* theta,
    const double* phi, const double* srcCellParams, int n_params, const int* CELL_RADIATES,
    const double* correction_srcCellParams, int numRays, const double* deflection,
    const double* cos_alpha, const double* lag, const double* maxDeflection,
    const double* cos_gammaArray, int n_energies, const double* energies, int n_leaves,
    const double* leaves, int n_phases, const double* phases,
    const xpsi_b200_atmosphere* hot_atmosphere, const xpsi_b200_atmosphere* elsewhere_atmosphere,
    int hot_atm_ext, int else_atm_ext, int beam_opt, int image_order_limit, double R_in,
    int phase_interpolant, double* flux_out) {
  return integrate_member(0, R, omega, r_s, inclination, n_rings, n_azi, cellArea, radial, r_s_over_r, theta, phi,
                          srcCellParams, n_params, CELL_RADIATES, correction_srcCellParams, numRays, deflection,
                          cos_alpha, lag, maxDeflection, cos_gammaArray, n_energies, energies, n_leaves, leaves,
                          n_phases, phases, hot_atmosphere, elsewhere_atmosphere, hot_atm_ext, else_atm_ext,
                          beam_opt, image_order_limit, R_in, phase_interpolant, flux_out);
}

int xpsi_b200_integrate_general(
    double R, double omega, double r_s, double inclination, int n_rings, int n_azi,
    const double* cellArea, const double* radial, const double* r_s_over_r, const double* theta,
    const double* phi, const double* srcCellParams, int n_params, const int* CELL_RADIATES,
    const double* correction_srcCellParams, int numRays, const double* deflection,
    const double* cos_alpha, const double* lag, const double* maxDeflection,
    const double* cos_gammaArray, int n_energies, const double* energies, int n_leaves,
    const double* leaves, int n_phases, const double* phases,
    const xpsi_b200_atmosphere* hot_atmosphere, const xpsi_b200_atmosphere* elsewhere_atmosphere,
    int hot_atm_ext, int else_atm_ext, int beam_opt, int image_order_limit, double R_in,
    int phase_interpolant, double* flux_out) {
  return integrate_member(1, R, omega, r_s, inclination, n_rings, n_azi, cellArea, radial, r_s_over_r, theta, phi,
                          srcCellParams, n_params, CELL_RADIATES, correction_srcCellParams, numRays, deflection,
                          cos_alpha, lag, maxDeflection, cos_gammaArray, n_energies, energies, n_leaves, leaves,
                          n_phases, phases, hot_atmosphere, elsewhere_atmosphere, hot_atm_ext, else_atm_ext,
                          beam_opt, image_order_limit, R_in, phase_interpolant, flux_out);
}

int xpsi_b200_integrate_time_invariance(
    double R, double omega, double r_s, double inclination, int sqrt_numPix, double cellArea,
    const double* radial, const double* r_s_over_r, const double* theta, const double* phi,
    const double* srcCellParams, int n_params, int numRays, const double* deflection,
    const double* cos_alpha, const double* maxDeflection, const double* cos_gammaArray, int n_energies,
    const double* energies, const xpsi_b200_atmosphere* atmosphere, int atm_ext, int image_order_limit,
    double* flux_out) {
  (void)R; (void)r_s;
  int rc = ensure_stream();
  if (rc) return rc;
  if (atm_ext != XPSI_B200_ATM_BB && atm_ext != XPSI_B200_ATM_NUM4D)
    return fail(XPSI_B200_EUNSUPPORTED, "atm_ext must be 1 (BB) or 2 (Num4D)");
  if (atm_ext == XPSI_B200_ATM_NUM4D && !atmosphere) return fail(XPSI_B200_EINVAL, "Num4D needs a preloaded atmosphere");
  if (sqrt_numPix < 1 || numRays < 3 || n_energies < 1 || n_energies > 256 || n_params < 1)
    return fail(XPSI_B200_EINVAL, "bad dimensions (at most 256 energies)");
  const size_t n = sqrt_numPix, nc = n * n, nr = n * numRays;
  Dev<double> d_scal, d_radial, d_rsr, d_theta, d_phi, d_par, d_defl, d_ca, d_maxd, d_cg, d_E, d_flux;
  Dev<int> d_status;
  const double scal[3] = {omega, inclination, cellArea};
  CK(d_scal.upload(scal, 3));
  CK(d_radial.upload(radial, n)); CK(d_rsr.upload(r_s_over_r, n)); CK(d_theta.upload(theta, nc));
  CK(d_phi.upload(phi, nc)); CK(d_par.upload(srcCellParams, nc * n_params));
  CK(d_defl.upload(deflection, nr)); CK(d_ca.upload(cos_alpha, nr));
  CK(d_maxd.upload(maxDeflection, n)); CK(d_cg.upload(cos_gammaArray, n)); CK(d_E.upload(energies, n_energies));
  CK(d_flux.alloc(n_energies)); CK(cudaMemsetAsync(d_flux.p, 0, n_energies * sizeof(double), g_stream));
  CK(d_status.alloc(1)); CK(cudaMemsetAsync(d_status.p, 0, sizeof(int), g_stream));
  xb::TinvArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = 1; a.sqrt_numPix = sqrt_numPix; a.n_rays = numRays; a.n_energies = n_energies; a.n_params = n_params;
  a.omega = d_scal.p; a.inclination = d_scal.p + 1; a.cellArea = d_scal.p + 2;
  a.radial = d_radial.p; a.r_s_over_r = d_rsr.p; a.theta = d_theta.p; a.phi = d_phi.p; a.srcParams = d_par.p;
  a.deflection = d_defl.p; a.cos_alpha = d_ca.p; a.maxDeflection = d_maxd.p; a.cos_gamma = d_cg.p;
  a.energies = d_E.p; a.atm_ext = atm_ext; a.image_order_limit = image_order_limit > 0 ? image_order_limit : 0;
  if (atm_ext == XPSI_B200_ATM_NUM4D) { a.atm = atmosphere->view; a.slab_rows = xb::tinv_slab_rows(a.atm, energies, n_energies); }
  a.flux = d_flux.p; a.status = d_status.p;
  cudaError_t e = xb::launch_integrate_tinv(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_integrate_tinv");
  g_launches += 2;
  int status = 0;
  CK(d_flux.download(flux_out, n_energies)); CK(d_status.download(&status, 1));
  CK(cudaStreamSynchronize(g_stream));
  if (status != 0) return fail(status, status == 1 ? "numerical error in time-invariant integration" : "unsupported configuration");
  return 0;
}

int xpsi_b200_energy_integrator(const double* signal, int n_energies, int n_phases,
                                const double* log10_energies, const double* log10_edges, int n_in,
                                int phase_interpolant, double* out) {
  int rc = ensure_stream();
  if (rc) return rc;
  if (n_energies < 5 || n_phases < 1 || n_in < 1) return fail(XPSI_B200_EINVAL, "bad dimensions");
  Dev<double> d_sig, d_x, d_edges, d_out;
  CK(d_sig.upload(signal, (size_t)n_energies * n_phases));
  CK(d_x.upload(log10_energies, n_energies));
  CK(d_edges.upload(log10_edges, n_in + 1));
  CK(d_out.alloc((size_t)n_in * n_phases));
  xb::EnergyIntegArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = 1; a.n_energies = n_energies; a.n_phases = n_phases; a.n_in = n_in;
  a.signal = d_sig.p; a.log10_energies = d_x.p; a.log10_edges = d_edges.p; a.interp = phase_interpolant;
  a.out = d_out.p; a.q_per_b = 1;
  cudaError_t e = xb::launch_energy_integrator(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_energy_integrator");
  g_launches += 1;
  std::vector<double> tmp((size_t)n_in * n_phases);
  CK(d_out.download(tmp.data(), tmp.size()));
  CK(cudaStreamSynchronize(g_stream));
  for (int p = 0; p < n_phases; ++p)           // device layout [p][j] -> reference layout [j][p]
    for (int j = 0; j < n_in; ++j) out[(size_t)j * n_phases + p] = tmp[(size_t)p * n_in + j];
  return 0;
}

int xpsi_b200_instrument_fold(const double* matrix, int n_rows, int n_cols, int i0, int i1, int o0,
                              int o1, const double* signal, int n_phases, double* out) {
  int rc = ensure_stream();
  if (rc) return rc;
  if (i0 < 0 || i1 > n_cols || i1 <= i0 || o0 < 0 || o1 > n_rows || o1 <= o0 || n_phases < 1)
    return fail(XPSI_B200_EINVAL, "bad ranges");
  const int n_in = i1 - i0, n_chan = o1 - o0;
  Dev<double> d_m, d_x, d_out;
  CK(d_m.upload(matrix + (size_t)o0 * n_cols, (size_t)n_chan * n_cols));
  // device operand layout is [phase][input interval]
  std::vector<double> xt((size_t)n_phases * n_in);
  for (int j = 0; j < n_in; ++j)
    for (int p = 0; p < n_phases; ++p) xt[(size_t)p * n_in + j] = signal[(size_t)j * n_phases + p];
  CK(d_x.upload(xt.data(), xt.size()));
  CK(d_out.alloc((size_t)n_chan * n_phases));
  std::vector<int> kr = response_k_ranges(matrix + (size_t)o0 * n_cols, n_chan, n_cols, i0, n_in);
  Dev<int> d_kr;
  CK(d_kr.upload(kr.data(), kr.size()));
  xb::FoldArgs a;
  a.n_cols = 1; a.n_phases = n_phases; a.n_in = n_in; a.n_chan = n_chan;
  a.matrix = d_m.p; a.ld_matrix = n_cols; a.in0 = i0; a.x = d_x.p; a.out = d_out.p; a.k_range = d_kr.p;
  cudaError_t e = xb::launch_fold(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_fold");
  g_launches += 1;
  CK(d_out.download(out, (size_t)n_chan * n_phases));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

static int row_spline_call(const double* x, int n_nodes, const double* y, int n_rows, long yrs, long yns,
                           size_t y_count, const double* q, int n_q, int n_out, int op, double shift,
                           double scale, int allow_negative, int interp, int periodic, double* out,
                           long ors, long ocs) {
  int rc = ensure_stream();
  if (rc) return rc;
  if (interp < 0 || interp > 2) return fail(XPSI_B200_EINVAL, "interpolant must be Akima (0), Steffen (1) or Cubic (2)");
  if (interp == 2 && n_nodes > 260) return fail(XPSI_B200_EUNSUPPORTED, "Cubic interpolant: at most 260 nodes");
  if (n_nodes < (interp == 0 ? 5 : 3) || n_rows < 1 || n_out < 1) return fail(XPSI_B200_EINVAL, "bad dimensions");
  Dev<double> d_x, d_y, d_q, d_o;
  CK(d_x.upload(x, n_nodes)); CK(d_y.upload(y, y_count)); CK(d_q.upload(q, n_q));
  CK(d_o.alloc((size_t)n_rows * n_out));
  xb::RowSplineArgs a;
  a.n_rows = n_rows; a.n_nodes = n_nodes; a.n_out = n_out; a.x = d_x.p; a.y = d_y.p;
  a.y_row_stride = yrs; a.y_node_stride = yns; a.q = d_q.p; a.op = op; a.shift = shift; a.scale = scale;
  a.allow_negative = allow_negative; a.interp = interp; a.periodic = periodic; a.out = d_o.p;
  a.out_row_stride = ors; a.out_col_stride = ocs;
  cudaError_t e = xb::launch_row_spline(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_row_spline");
  g_launches += 1;
  CK(d_o.download(out, (size_t)n_rows * n_out));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_phase_integrator(double exposure_time, const double* phases, int n_bins, const double* signal,
                               int n_rows, const double* signal_phases, int n_phases, double phase_shift,
                               int allow_negative, int phase_interpolant, double* out) {
  return row_spline_call(signal_phases, n_phases, signal, n_rows, n_phases, 1, (size_t)n_rows * n_phases, phases,
                         n_bins + 1, n_bins, 0, phase_shift, exposure_time, allow_negative, phase_interpolant, 1,
                         out, n_bins, 1);
}

int xpsi_b200_phase_interpolator(const double* new_phases, int n_new, const double* phases, int n_phases,
                                 const double* signal, int n_rows, double phase_shift, int allow_negative,
                                 int phase_interpolant, double* out) {
  return row_spline_call(phases, n_phases, signal, n_rows, n_phases, 1, (size_t)n_rows * n_phases, new_phases,
                         n_new, n_new, 1, phase_shift, 1.0, allow_negative, phase_interpolant, 1, out, n_new, 1);
}

int xpsi_b200_energy_interpolator(const double* signal, int n_energies, int n_phases, const double* log10_energies,
                                  const double* new_log10_energies, int n_new, int energy_interpolant,
                                  double* out) {
  // one spline per phase column; result laid out [n_new][n_phases] like the reference's transposed return
  return row_spline_call(log10_energies, n_energies, signal, n_phases, 1, n_phases, (size_t)n_energies * n_phases,
                         new_log10_energies, n_new, n_new, 2, 0.0, 1.0, 1, energy_interpolant, 0, out, 1, n_phases);
}

int xpsi_b200_intensity(int n, const double* energies, const double* mu, const double* local_variables,
                        int n_vars, const xpsi_b200_atmosphere* atmosphere, int region_extension, int atm_ext,
                        int beam_opt, double* out) {
  int rc = ensure_stream();
  if (rc) return rc;
  if (n < 1 || n_vars < 1) return fail(XPSI_B200_EINVAL, "bad dimensions");
  if (atm_ext != XPSI_B200_ATM_BB && atm_ext != XPSI_B200_ATM_NUM4D)
    return fail(XPSI_B200_EUNSUPPORTED, "atmosphere extension must be BB (1) or Num4D (2)");
  if (atm_ext == XPSI_B200_ATM_NUM4D && !atmosphere) return fail(XPSI_B200_EINVAL, "Num4D needs a preloaded atmosphere");
  if (atm_ext == XPSI_B200_ATM_NUM4D && n_vars < 2) return fail(XPSI_B200_EINVAL, "Num4D needs (log T, log g)");
  if (region_extension != 0 && region_extension != 1) return fail(XPSI_B200_EINVAL, "region must be 0 (hot) or 1 (elsewhere)");
  if (beam_opt < 0 || beam_opt > 3) return fail(XPSI_B200_EINVAL, "beam_opt must be 0-3");
  if (region_extension == 0 && beam_opt != 0 && n_vars < 7)
    return fail(XPSI_B200_EINVAL, "beaming needs 7 local variables per point");
  Dev<double> d_E, d_mu, d_v, d_o;
  CK(d_E.upload(energies, n)); CK(d_mu.upload(mu, n)); CK(d_v.upload(local_variables, (size_t)n * n_vars));
  CK(d_o.alloc(n));
  xb::IntensityArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.n_vars = n_vars; a.energies = d_E.p; a.mu = d_mu.p; a.vars = d_v.p;
  if (atmosphere) a.atm = atmosphere->view;
  a.atm_ext = atm_ext; a.region = region_extension; a.beam_opt = beam_opt; a.out = d_o.p;
  cudaError_t e = xb::launch_intensity(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_intensity");
  g_launches += 1;
  CK(d_o.download(out, n));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_interstellar_attenuate(const double* attenuation, int n_rows, int n_cols, double* signal) {
  int rc = ensure_stream();
  if (rc) return rc;
  if (n_rows < 1 || n_cols < 1) return fail(XPSI_B200_EINVAL, "bad dimensions");
  Dev<double> d_a, d_s;
  CK(d_a.upload(attenuation, n_rows)); CK(d_s.upload(signal, (size_t)n_rows * n_cols));
  cudaError_t e = xb::launch_attenuate(d_a.p, n_rows, n_cols, d_s.p, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_attenuate");
  g_launches += 1;
  CK(d_s.download(signal, (size_t)n_rows * n_cols));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_precomputation(const int* counts, int n_chan, int n_bins, double* out) {
  int rc = ensure_stream();
  if (rc) return rc;
  Dev<int> d_c; Dev<double> d_o;
  CK(d_c.upload(counts, (size_t)n_chan * n_bins));
  CK(d_o.alloc(n_chan));
  cudaError_t e = xb::launch_precomputation(d_c.p, n_chan, n_bins, d_o.p, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_precomputation");
  g_launches += 1;
  CK(d_o.download(out, n_chan));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

}  // extern "C"

namespace {
// components with their own phase grids -> padded device arrays: pulses [n_comp][n_chan][Pmax], grids [n_comp][Pmax]
struct CompUpload { Dev<double> pulses, grids; Dev<int> n_nodes, allow; int Pmax = 0; bool uniform = true; };

int upload_components(CompUpload& u, const double* const* components, int n_comp, int n_chan,
                      const double* const* component_phases, const int* n_phases, const int* allow_negative) {
  u.Pmax = 0;
  for (int c = 0; c < n_comp; ++c) {
    if (n_phases[c] < 5 && n_phases[c] != 1) return fail(XPSI_B200_EINVAL, "a component needs 1 (time-invariant) or >= 5 phases");
    if (n_phases[c] > u.Pmax) u.Pmax = n_phases[c];
  }
  u.uniform = true;
  for (int c = 1; c < n_comp; ++c) {
    if (n_phases[c] != n_phases[0]) { u.uniform = false; break; }
    for (int i = 0; i < n_phases[0]; ++i)
      if (component_phases[c][i] != component_phases[0][i]) { u.uniform = false; break; }
  }
  const size_t P = (size_t)u.Pmax, np = (size_t)n_chan * P;
  CK(u.pulses.alloc(np * n_comp));
  CK(u.grids.alloc(P * n_comp));
  if (!u.uniform) {
    CK(cudaMemsetAsync(u.pulses.p, 0, np * n_comp * sizeof(double), g_stream));
    CK(cudaMemsetAsync(u.grids.p, 0, P * n_comp * sizeof(double), g_stream));
  }
  for (int c = 0; c < n_comp; ++c) {
    const size_t w = (size_t)n_phases[c] * sizeof(double);
    g_h2d += (long long)(w * (n_chan + 1));
    CK(cudaMemcpy2DAsync(u.pulses.p + c * np, P * sizeof(double), components[c], w, w, n_chan, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(u.grids.p + c * P, component_phases[c], w, cudaMemcpyHostToDevice, g_stream));
  }
  CK(u.n_nodes.upload(n_phases, n_comp));
  if (allow_negative) CK(u.allow.upload(allow_negative, n_comp));
  return 0;
}
}  // namespace

extern "C" {
int xpsi_b200_eval_marginal_likelihood(
    double exposure_time, const double* phases, int n_bins, const double* counts, int n_chan,
    const double* const* components, int n_comp, const double* const* component_phases, const int* n_phases,
    const double* phase_shifts, const double* precomp, const double* support, double epsilon,
    double sigmas, double llzero, const int* allow_negative, double slim, const double* background,
    int phase_interpolant, double* lnL, double* expected_counts, double* mcl_background,
    double* mcl_background_given_support) {
  int rc = ensure_stream();
  if (rc) return rc;
  if (n_bins < 1 || n_bins > xb::marginal_max_bins()) return fail(XPSI_B200_EUNSUPPORTED, "1..128 phase bins supported");
  if (n_comp < 1 || n_chan < 1 || !component_phases || !n_phases) return fail(XPSI_B200_EINVAL, "bad dimensions");
  Dev<double> d_sh, d_dph, d_cnt, d_pre, d_sup, d_bg, d_clnl, d_exp, d_mb, d_mbs, d_lnl;
  Dev<int> d_cst, d_st;
  CompUpload cu;
  rc = upload_components(cu, components, n_comp, n_chan, component_phases, n_phases, allow_negative);
  if (rc) return rc;
  const int n_phases_max = cu.Pmax;
  CK(d_sh.upload(phase_shifts, n_comp));
  CK(d_dph.upload(phases, n_bins + 1)); CK(d_cnt.upload(counts, (size_t)n_chan * n_bins));
  CK(d_pre.upload(precomp, n_chan)); CK(d_sup.upload(support, (size_t)n_chan * 2));
  if (background) CK(d_bg.upload(background, (size_t)n_chan * n_bins));
  CK(d_clnl.alloc(n_chan)); CK(d_cst.alloc(n_chan)); CK(d_exp.alloc((size_t)n_chan * n_bins));
  CK(d_mb.alloc(n_chan)); CK(d_mbs.alloc(n_chan)); CK(d_lnl.alloc(1)); CK(d_st.alloc(1));
  CK(cudaMemsetAsync(d_st.p, 0, sizeof(int), g_stream));
  xb::MarginalArgs a;
  memset(&a, 0, sizeof(a));
  a.B = 1; a.n_comp = n_comp; a.n_chan = n_chan; a.n_phases = n_phases_max; a.n_bins = n_bins;
  a.pulses = cu.pulses.p; a.comp_phases = cu.grids.p; a.phase_shifts = d_sh.p; a.data_phases = d_dph.p;
  if (!cu.uniform) a.comp_n_phases = cu.n_nodes.p;
  if (allow_negative) a.comp_allow_negative = cu.allow.p;
  a.counts = d_cnt.p; a.precomp = d_pre.p; a.support = d_sup.p; a.background = background ? d_bg.p : nullptr;
  a.exposure_time = exposure_time; a.epsilon = epsilon; a.sigmas = sigmas; a.llzero = llzero; a.slim = slim;
  a.allow_negative = 0; a.interp = phase_interpolant;
  a.chan_lnL = d_clnl.p; a.chan_status = d_cst.p; a.expected = d_exp.p; a.mcl_bg = d_mb.p;
  a.mcl_bg_support = d_mbs.p; a.lnL = d_lnl.p; a.status = d_st.p;
  cudaError_t e = xb::launch_marginal(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_marginal");
  g_launches += 2;
  int status = 0;
  CK(d_lnl.download(lnL, 1)); CK(d_st.download(&status, 1));
  if (expected_counts) CK(d_exp.download(expected_counts, (size_t)n_chan * n_bins));
  if (mcl_background) CK(d_mb.download(mcl_background, n_chan));
  if (mcl_background_given_support) CK(d_mbs.download(mcl_background_given_support, n_chan));
  CK(cudaStreamSynchronize(g_stream));
  if (status != 0) return fail(status, status == 11 ? "model exceeds data by more than slim sigma" : "marginal integral failed");
  return 0;
}

int xpsi_b200_poisson_likelihood_given_background(
    double exposure_time, const double* phases, int n_bins, const double* counts, int n_chan,
    const double* const* components, int n_comp, const double* const* component_phases, const int* n_phases,
    const double* phase_shifts, const double* background, const double* neg_sum_ln_data_factorial,
    const int* allow_negative, int phase_interpolant, double* lnL, double* expected_counts) {
  int rc = ensure_stream();
  if (rc) return rc;
  if (n_bins < 1 || n_bins > xb::marginal_max_bins()) return fail(XPSI_B200_EUNSUPPORTED, "1..128 phase bins supported");
  if (n_comp < 1 || n_chan < 1 || !component_phases || !n_phases || !background) return fail(XPSI_B200_EINVAL, "bad arguments");
  Dev<double> d_sh, d_dph, d_cnt, d_pre, d_bg, d_clnl, d_exp, d_lnl, d_sup;
  Dev<int> d_cst, d_st;
  CompUpload cu;
  rc = upload_components(cu, components, n_comp, n_chan, component_phases, n_phases, allow_negative);
  if (rc) return rc;
  const int n_phases_max = cu.Pmax;
  CK(d_sh.upload(phase_shifts, n_comp));
  CK(d_dph.upload(phases, n_bins + 1));
  if (counts) CK(d_cnt.upload(counts, (size_t)n_chan * n_bins));
  if (neg_sum_ln_data_factorial) CK(d_pre.upload(neg_sum_ln_data_factorial, n_chan));
  CK(d_bg.upload(background, (size_t)n_chan * n_bins));
  std::vector<double> sup(2 * (size_t)n_chan, 0.0);
  CK(d_sup.upload(sup.data(), sup.size()));
  CK(d_clnl.alloc(n_chan)); CK(d_cst.alloc(n_chan)); CK(d_exp.alloc((size_t)n_chan * n_bins));
  CK(d_lnl.alloc(1)); CK(d_st.alloc(1));
  CK(cudaMemsetAsync(d_st.p, 0, sizeof(int), g_stream));
  xb::MarginalArgs a;
  memset(&a, 0, sizeof(a));
  a.B = 1; a.n_comp = n_comp; a.n_chan = n_chan; a.n_phases = n_phases_max; a.n_bins = n_bins;
  a.pulses = cu.pulses.p; a.comp_phases = cu.grids.p; a.phase_shifts = d_sh.p; a.data_phases = d_dph.p;
  if (!cu.uniform) a.comp_n_phases = cu.n_nodes.p;
  if (allow_negative) a.comp_allow_negative = cu.allow.p;
  a.counts = counts ? d_cnt.p : nullptr; a.precomp = neg_sum_ln_data_factorial ? d_pre.p : nullptr;
  a.support = d_sup.p; a.background = d_bg.p; a.exposure_time = exposure_time; a.slim = -1.0;
  a.allow_negative = 0; a.interp = phase_interpolant; a.given_background = 1;
  a.chan_lnL = d_clnl.p; a.chan_status = d_cst.p; a.expected = d_exp.p; a.lnL = d_lnl.p; a.status = d_st.p;
  cudaError_t e = xb::launch_marginal(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_marginal");
  g_launches += 2;
  int status = 0;
  if (lnL) CK(d_lnl.download(lnL, 1));
  CK(d_st.download(&status, 1));
  if (expected_counts) CK(d_exp.download(expected_counts, (size_t)n_chan * n_bins));
  CK(cudaStreamSynchronize(g_stream));
  if (status != 0 && counts) return fail(XPSI_B200_EQUADRATURE, "zero expectation in a bin that holds counts");
  return 0;
}

}  // extern "C"

// ===========================================================================
// batched pipeline
// ===========================================================================
struct xpsi_b200_pipeline {
  xpsi_b200_pipeline_config cfg;
  int max_batch = 0, slab_rows_chunk = 0, slab_rows_ring = 0;
  std::vector<int> member_component;
  const xpsi_b200_atmosphere* atm = nullptr;
  // constants
  Dev<double> energies, log10E, leaves, phases, phase_cycles, log10_edges, response, data_phases, counts,
      support, precomp;
  Dev<int> col_of_q, k_range; Dev<int2> ei_span;
  // per-batch inputs
  Dev<double> omega, inclination, d_sq, shifts, omega_q, incl_q, cellArea, phi, theta, radial, rsr, params,
      defl, calpha, lag, maxd, cgamma;
  Dev<int> n_rings, n_azi;
  // intermediates / outputs
  Dev<double> flux, xin, folded, chan_lnL, expected, lnL;
  Dev<int> chan_status, status_q, status;
  Dev<unsigned long long> work;
  Dev<double> ws_leaf, ws_hdr, ws_slab, ws_cells, ws_tiles, flux_part; Dev<int> ws_ihdr, ws_tmeta, ws_redo, ws_ovf; Dev<int2> ws_thdr;
  int tile_cap = 0;
  int deterministic = 0;             // ring sums by a two-stage ordered reduction instead of fp64 atomics
  // embed inputs / scratch
  Dev<double> e_Req, e_rs, e_eps, e_zeta, e_colat, e_rad, e_temp, e_phish, e_maxAlpha, e_hrad, e_hcolat, e_hazi, e_extra;
  Dev<int> e_partner, e_iscede, e_member_cells;
  int count_work = 0;
  int embed_status_valid = 0;        // status[] already carries embed failures for this batch
  xb::EmbedArgs embed_args;          // last uploaded spot batch (device pointers), for resident re-runs
  int embed_ready = 0;
  cudaEvent_t ev_embed[2] = {nullptr, nullptr};
  float embed_ms = 0.f; int embed_timed = 0;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_flux[2] = {nullptr, nullptr};
  float stage_ms[4] = {0, 0, 0, 0};
  // optional components (xpsi_b200_pipeline_set_extras)
  xpsi_b200_pipeline_extras ex = {};
  int else_slab_rows = 0, slab2_chunk = 0, slab2_ring = 0;
  Dev<double> att_base, att_power, corr, ws_slab2;
  int att_power_valid = 0, else_arrays_valid = 0, corr_valid = 0, else_temp_valid = 0;
  int extras_B = 0;                  // batch size the per-batch extras on the device were uploaded for
  Dev<double> x_temp, x_area, x_radial, x_rsr, x_theta, x_phi, x_params, x_defl, x_calpha, x_maxd, x_cgamma,
      x_maxAlpha, x_grav, x_flux;
  Dev<int> x_nrings, x_status;
  // further signals (instruments / data sets) registered from the same photosphere signal
  // (xpsi/Likelihood.py:346-420 loops over the signals of a photosphere): xpsi_b200_pipeline_add_signal
  struct SignalPart {
    int n_in = 0, n_chan = 0, n_bins = 0, allow_negative = 0;
    double exposure_time = 0, epsilon = 0, sigmas = 0, llzero = 0, slim = 0;
    Dev<double> log10_edges, response, data_phases, counts, support, precomp, att_base;
    Dev<int> k_range; Dev<int2> ei_span;
    Dev<double> xin, folded, chan_lnL, expected, lnL; Dev<int> chan_status;
  };
  std::vector<std::unique_ptr<SignalPart>> more;
  Dev<double> sig_shift, sig_shifts_s;   // [B][n_signals] instrument phase shifts (xpsi/Signal.py:581-583), scratch [B][C]
  int sig_shift_valid = 0;
  int tinv_only = 0;                  // the star is an Everywhere(time_invariant=True) surface: one phase column
  // sweep store: N parameter vectors uploaded once, evaluated block by block (xpsi_b200_pipeline_sweep_*)
  struct SpotStore {
    size_t N = 0;
    Dev<double> omega, incl, d_sq, shifts, Req, rs, eps, zeta, colat, rad, temp, phish, hrad, hcolat, hazi, extra,
        att_power, else_temp, lnL, sig_shift;
    Dev<int> status;
    int has_hole = 0, has_partner = 0, has_extra = 0, has_att = 0, has_else = 0, has_sig_shift = 0, has_member_cells = 0;
    double mode_frequency = 0.0;
    int num_cells = 0, min_sqrt = 0, max_sqrt = 0;
  } store;
};

namespace {

__global__ void k_expand_scalars(const double* omega, const double* incl, int B, int M, double* omega_q,
                                 double* incl_q) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B * M) return;
  omega_q[q] = omega[q / M]; incl_q[q] = incl[q / M];
}

__global__ void k_member_status(const int* status_q, int B, int M, int* status, int keep) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int s = keep ? status[b] : 0;
  for (int m = 0; m < M; ++m) if (status_q[b * M + m] != 0 && s == 0) s = status_q[b * M + m];
  status[b] = s;
}

// Photosphere.py:589-592: add the (already normalised) elsewhere spectrum to every phase column of member 0;
// flux holds raw ring sums that are later divided by E keV, so the spectrum is multiplied back first
__global__ void k_add_spectrum(const double* spectrum, const double* energies, int B, int M, int N_E, int N_P,
                               double* flux) {
  const long n = (long)B * N_E * N_P;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    const int b = (int)(t / ((long)N_E * N_P));
    const long r = t - (long)b * N_E * N_P;
    const int e = (int)(r / N_P);
    flux[(long)b * M * N_E * N_P + r] += spectrum[(long)b * N_E + e] * (energies[e] * 1.60217662e-16);   // keV, xpsi/global_imports.py:71
  }
}

// shifts of signal s = hot-region shifts + the signal's own phase shift (xpsi/Signal.py:581-583)
__global__ void k_signal_shifts(const double* shifts, const double* sig_shift, int B, int C, int S, int s, double* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < B * C) out[t] = shifts[t] + sig_shift[(t / C) * S + s];
}

// joint log-likelihood: sum over the signals (xpsi/Likelihood.py:494-500)
__global__ void k_add_lnL(const double* part, int B, double* lnL) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) lnL[b] += part[b];
}

__global__ void k_merge_status(const int* src, int B, int* dst) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && dst[b] == 0 && src[b] != 0) dst[b] = src[b];
}

int pipeline_upload(xpsi_b200_pipeline* p, int B, const xpsi_b200_batch* h) {
  p->embed_status_valid = 0;
  const xpsi_b200_pipeline_config& c = p->cfg;
  const size_t Q = (size_t)B * c.n_members, R = c.max_rings, A = c.max_azi;
  CK(p->omega.upload(h->omega, B)); CK(p->inclination.upload(h->inclination, B));
  CK(p->d_sq.upload(h->d_sq, B)); CK(p->shifts.upload(h->phase_shifts, (size_t)B * c.n_components));
  CK(p->n_rings.upload(h->n_rings, Q)); CK(p->n_azi.upload(h->n_azi, Q));
  CK(p->cellArea.upload(h->cellArea, Q * R * A)); CK(p->phi.upload(h->phi, Q * R * A));
  CK(p->theta.upload(h->theta, Q * R)); CK(p->radial.upload(h->radial, Q * R));
  CK(p->rsr.upload(h->r_s_over_r, Q * R)); CK(p->params.upload(h->srcParams, Q * R * c.n_params));
  CK(p->defl.upload(h->deflection, Q * R * c.n_rays)); CK(p->calpha.upload(h->cos_alpha, Q * R * c.n_rays));
  CK(p->lag.upload(h->lag, Q * R * c.n_rays)); CK(p->maxd.upload(h->maxDeflection, Q * R));
  CK(p->cgamma.upload(h->cos_gamma, Q * R));
  return 0;
}

int pipeline_run(xpsi_b200_pipeline* p, int B) {
  const xpsi_b200_pipeline_config& c = p->cfg;
  const int M = c.n_members, C = c.n_components, Q = B * M;
  const size_t nflux = (size_t)Q * c.n_energies * c.n_phases;
  // per-batch extras (attenuation powers, elsewhere / correction arrays, signal shifts) belong to ONE batch: a run
  // with another batch size would read stale or uninitialised entries
  if ((p->att_power_valid || p->else_arrays_valid || p->else_temp_valid || p->sig_shift_valid) && p->extras_B != B)
    return fail(XPSI_B200_EINVAL, "the per-batch extras on the device were uploaded for a different batch size");
  CK(cudaEventRecord(p->ev[0], g_stream));
  CK(cudaMemsetAsync(p->flux.p, 0, nflux * sizeof(double), g_stream));
  CK(cudaMemsetAsync(p->status_q.p, 0, Q * sizeof(int), g_stream));
  // time-invariant spectrum of the closed surface mesh: Elsewhere (xpsi/Elsewhere.py:393-442) or, for a
  // pipeline created with one phase column, Everywhere(time_invariant=True) (xpsi/Everywhere.py:577-601)
  auto run_closed_surface = [&]() -> int {
    if (!p->else_arrays_valid || (!p->tinv_only && !p->corr_valid))
      return fail(XPSI_B200_EINVAL, "the closed surface (elsewhere / everywhere) is enabled but its per-batch inputs were not provided");
    xb::TinvArgs t;
    memset(&t, 0, sizeof(t));
    t.Q = B; t.sqrt_numPix = p->ex.else_sqrt_num_cells; t.n_rays = p->ex.else_num_rays;
    t.n_energies = c.n_energies; t.n_params = 2;
    t.omega = p->omega.p; t.inclination = p->inclination.p; t.cellArea = p->x_area.p;
    t.radial = p->x_radial.p; t.r_s_over_r = p->x_rsr.p; t.theta = p->x_theta.p; t.phi = p->x_phi.p;
    t.srcParams = p->x_params.p; t.deflection = p->x_defl.p; t.cos_alpha = p->x_calpha.p;
    t.maxDeflection = p->x_maxd.p; t.cos_gamma = p->x_cgamma.p; t.energies = p->energies.p;
    t.atm_ext = p->ex.else_atm_ext;
    if (t.atm_ext == XPSI_B200_ATM_NUM4D) { t.atm = p->ex.elsewhere_atmosphere->view; t.slab_rows = p->else_slab_rows; }
    t.image_order_limit = p->ex.else_image_order_limit > 0 ? p->ex.else_image_order_limit : 0;
    t.flux = p->x_flux.p; t.status = p->x_status.p;
    CK(cudaMemsetAsync(p->x_flux.p, 0, (size_t)B * c.n_energies * sizeof(double), g_stream));
    CK(cudaMemsetAsync(p->x_status.p, 0, B * sizeof(int), g_stream));
    cudaError_t et = xb::launch_integrate_tinv(t, g_stream);
    if (et != cudaSuccess) return cuda_fail(et, "launch_integrate_tinv");
    g_launches += 2;
    return 0;
  };
  auto add_spectrum = [&]() {
    int blocks = (int)(((long)B * c.n_energies * c.n_phases + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_add_spectrum<<<blocks, 256, 0, g_stream>>>(p->x_flux.p, p->energies.p, B, M, c.n_energies, c.n_phases, p->flux.p);
    g_launches += 1;
  };
  cudaError_t e = cudaSuccess;
  if (p->tinv_only) {
    // Photosphere.py:560-566: the star's signal is the Everywhere spectrum, one phase column
    if (!p->ex.elsewhere) return fail(XPSI_B200_EINVAL, "a time-invariant pipeline needs the closed-surface settings (set_extras)");
    int rc = run_closed_surface();
    if (rc) return rc;
    add_spectrum();
    CK(cudaEventRecord(p->ev_flux[0], g_stream)); CK(cudaEventRecord(p->ev_flux[1], g_stream));
  } else {
  k_expand_scalars<<<(Q + 127) / 128, 128, 0, g_stream>>>(p->omega.p, p->inclination.p, B, M, p->omega_q.p, p->incl_q.p);
  xb::AzinvArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = Q; a.n_rings = c.max_rings; a.n_azi = c.max_azi; a.n_rays = c.n_rays; a.n_energies = c.n_energies;
  a.n_leaves = c.n_leaves; a.n_phases = c.n_phases; a.n_params = c.n_params;
  a.n_rings_q = p->n_rings.p; a.n_azi_q = p->n_azi.p;
  a.omega = p->omega_q.p; a.inclination = p->incl_q.p;
  a.cellArea = p->cellArea.p; a.phi = p->phi.p; a.theta = p->theta.p; a.theta_ring_stride = 1;
  a.radial = p->radial.p; a.r_s_over_r = p->rsr.p; a.srcParams = p->params.p; a.params_per_cell = 0;
  a.radiates = nullptr; a.deflection = p->defl.p; a.cos_alpha = p->calpha.p; a.lag = p->lag.p;
  a.maxDeflection = p->maxd.p; a.cos_gamma = p->cgamma.p;
  a.energies = p->energies.p; a.leaves = p->leaves.p; a.phases = p->phases.p;
  a.log10_energies = p->log10E.p;
  a.hot_atm_ext = c.hot_atm_ext;
  if (c.hot_atm_ext == XPSI_B200_ATM_NUM4D) {
    a.hot = p->atm->view; a.slab_ne_max = p->slab_rows_chunk; a.slab_rows_ring = p->slab_rows_ring;
  }
  a.image_order_limit = c.image_order_limit > 0 ? c.image_order_limit : 0;
  a.n_img_max = c.image_order_limit > 0 ? c.image_order_limit : xb::kMaxImages;
  a.phase_interp = c.phase_interpolant;
  a.R_in = 1.0e6;
  a.scale_by_energy = 0;
  a.beam_opt = p->ex.beam_opt;
  if (p->ex.elsewhere) {
    // ---- Elsewhere: time-invariant spectrum of the closed mesh, then the correction in the hot members ----
    int rc = run_closed_surface();
    if (rc) return rc;
    a.corrParams = p->corr.p; a.else_atm_ext = p->ex.else_atm_ext;
    if (a.else_atm_ext == XPSI_B200_ATM_NUM4D) {
      a.els = p->ex.elsewhere_atmosphere->view; a.ws_slab2 = p->ws_slab2.p;
      if (p->slab2_chunk > a.slab_ne_max) a.slab_ne_max = p->slab2_chunk;
      if (p->slab2_ring > a.slab_rows_ring) a.slab_rows_ring = p->slab2_ring;
    }
  }
  a.flux = p->flux.p; a.status = p->status_q.p;
  a.ws_leaf = p->ws_leaf.p; a.ws_hdr = p->ws_hdr.p; a.ws_ihdr = p->ws_ihdr.p; a.ws_slab = p->ws_slab.p;
  a.ws_tiles = p->ws_tiles.p; a.ws_tmeta = p->ws_tmeta.p; a.ws_thdr = p->ws_thdr.p; a.tile_cap = p->tile_cap;
  a.ws_redo = p->ws_redo.p; a.ws_ovf = p->ws_ovf.p;
  { const char* fr = getenv("XPSI_B200_FORCE_REDO"); a.force_redo = (fr && fr[0] == '1') ? 1 : 0; }
  a.ws_cells = p->ws_cells.p;
  a.flux_part = p->deterministic ? p->flux_part.p : nullptr;
  a.work = p->count_work ? p->work.p : nullptr;       // accumulates over evaluations; reset by work_counters()
  a.ev_flux[0] = p->ev_flux[0]; a.ev_flux[1] = p->ev_flux[1];
  e = xb::launch_integrate_azinv(a, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_integrate_azinv");
  if (p->ex.elsewhere) {
    add_spectrum();
    g_launches += (p->ex.else_atm_ext == XPSI_B200_ATM_NUM4D ? 2 : 0);
  }
  }
  CK(cudaEventRecord(p->ev[1], g_stream));

  // ---- per signal: energy integration onto the instrument's input intervals (+ interstellar attenuation),
  // response fold, marginal likelihood.  Signal 0 is the one of the pipeline configuration; the others were
  // registered with xpsi_b200_pipeline_add_signal and read the same photosphere signal (p->flux).
  const int P_sig = c.n_phases;
  const int S = 1 + (int)p->more.size();
  k_member_status<<<(B + 127) / 128, 128, 0, g_stream>>>(p->status_q.p, B, M, p->status.p, p->embed_status_valid);
  if (p->ex.elsewhere) k_merge_status<<<(B + 127) / 128, 128, 0, g_stream>>>(p->x_status.p, B, p->status.p);
  p->embed_status_valid = 0;
  for (int s = 0; s < S; ++s) {
    xpsi_b200_pipeline::SignalPart* sp = s ? p->more[s - 1].get() : nullptr;
    const int n_in = sp ? sp->n_in : c.n_in, n_chan = sp ? sp->n_chan : c.n_chan, n_bins = sp ? sp->n_bins : c.n_bins;
    double* xin = sp ? sp->xin.p : p->xin.p;
    double* folded = sp ? sp->folded.p : p->folded.p;
    xb::EnergyIntegArgs ei;
    memset(&ei, 0, sizeof(ei));
    ei.Q = Q; ei.n_energies = c.n_energies; ei.n_phases = P_sig; ei.n_in = n_in;
    ei.signal = p->flux.p; ei.raw_energies = p->energies.p; ei.div_b = p->d_sq.p; ei.q_per_b = M;
    ei.log10_energies = p->log10E.p; ei.log10_edges = sp ? sp->log10_edges.p : p->log10_edges.p;
    ei.interp = c.phase_interpolant;
    ei.span = sp ? sp->ei_span.p : p->ei_span.p;
    ei.col_of_q = p->col_of_q.p; ei.accumulate = (M > C) ? 1 : 0; ei.out = xin;
    const double* att = sp ? sp->att_base.p : p->att_base.p;
    if (att) { ei.attenuation = att; ei.att_power = p->att_power_valid ? p->att_power.p : nullptr; }
    if (ei.accumulate)
      CK(cudaMemsetAsync(xin, 0, (size_t)B * C * P_sig * n_in * sizeof(double), g_stream));
    e = xb::launch_energy_integrator(ei, g_stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_energy_integrator");
    if (s == 0) CK(cudaEventRecord(p->ev[2], g_stream));

    xb::FoldArgs f;
    f.n_cols = B * C; f.n_phases = P_sig; f.n_in = n_in; f.n_chan = n_chan;
    f.matrix = sp ? sp->response.p : p->response.p; f.ld_matrix = n_in; f.in0 = 0; f.x = xin; f.out = folded;
    f.k_range = sp ? sp->k_range.p : p->k_range.p;
    e = xb::launch_fold(f, g_stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_fold");
    if (s == 0) CK(cudaEventRecord(p->ev[3], g_stream));

    const double* shifts = p->shifts.p;
    if (p->sig_shift_valid) {
      k_signal_shifts<<<(B * C + 127) / 128, 128, 0, g_stream>>>(p->shifts.p, p->sig_shift.p, B, C, S, s, p->sig_shifts_s.p);
      shifts = p->sig_shifts_s.p;
      g_launches += 1;
    }
    xb::MarginalArgs m;
    memset(&m, 0, sizeof(m));
    m.B = B; m.n_comp = C; m.n_chan = n_chan; m.n_phases = P_sig; m.n_bins = n_bins;
    m.pulses = folded; m.comp_phases = p->phase_cycles.p; m.phase_shifts = shifts;
    m.data_phases = sp ? sp->data_phases.p : p->data_phases.p; m.counts = sp ? sp->counts.p : p->counts.p;
    m.precomp = sp ? sp->precomp.p : p->precomp.p; m.support = sp ? sp->support.p : p->support.p;
    m.exposure_time = sp ? sp->exposure_time : c.exposure_time; m.epsilon = sp ? sp->epsilon : c.epsilon;
    m.sigmas = sp ? sp->sigmas : c.sigmas; m.llzero = sp ? sp->llzero : c.llzero;
    m.slim = sp ? sp->slim : c.slim; m.allow_negative = sp ? sp->allow_negative : c.allow_negative;
    m.interp = c.phase_interpolant;
    m.chan_lnL = sp ? sp->chan_lnL.p : p->chan_lnL.p; m.chan_status = sp ? sp->chan_status.p : p->chan_status.p;
    m.expected = sp ? sp->expected.p : p->expected.p;
    m.lnL = sp ? sp->lnL.p : p->lnL.p; m.status = p->status.p;      // a status raised by an earlier signal is kept
    e = xb::launch_marginal(m, g_stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_marginal");
    if (sp) {
      k_add_lnL<<<(B + 127) / 128, 128, 0, g_stream>>>(sp->lnL.p, B, p->lnL.p);
      g_launches += 5;        // energy, fold, marginal, channel-sum, add
    }
  }
  CK(cudaEventRecord(p->ev[4], g_stream));     // with several signals the "marginal" stage time covers signals 1..S-1 whole
  // expand, geometry, [slab, slab-member], tiles, flux (tensor-core + scalar for overflow rings), [ring reduction],
  // energy, fold, member-status, marginal, channel-sum
  g_launches += 10 + (c.hot_atm_ext == XPSI_B200_ATM_NUM4D ? 2 : 0) + (p->deterministic ? 1 : 0);
  return 0;
}

int pipeline_embed_launch(xpsi_b200_pipeline* p) {
  CK(cudaEventRecord(p->ev_embed[0], g_stream));
  if (p->ex.elsewhere) {
    if (!p->else_temp_valid) return fail(XPSI_B200_EINVAL, "elsewhere is enabled but else_temperature was not uploaded");
    p->embed_args.else_temperature = p->x_temp.p; p->embed_args.corrParams = p->corr.p;
    xb::ClosedMeshArgs m;
    memset(&m, 0, sizeof(m));
    m.B = p->embed_args.B; m.n = p->ex.else_sqrt_num_cells; m.n_rays = p->ex.else_num_rays;
    m.R_eq = p->embed_args.R_eq; m.r_s = p->embed_args.r_s; m.epsilon = p->embed_args.epsilon; m.zeta = p->embed_args.zeta;
    m.temperature = p->x_temp.p; m.cellArea = p->x_area.p; m.theta = p->x_theta.p; m.phi = p->x_phi.p;
    m.srcParams = p->x_params.p; m.radial = p->x_radial.p; m.r_s_over_r = p->x_rsr.p; m.cos_gamma = p->x_cgamma.p;
    m.maxAlpha = p->x_maxAlpha.p; m.ring_gravity = p->x_grav.p; m.deflection = p->x_defl.p; m.cos_alpha = p->x_calpha.p;
    m.maxDeflection = p->x_maxd.p; m.n_rings = p->x_nrings.p; m.status = p->embed_args.status;
    cudaError_t ec = xb::launch_embed_closed(m, g_stream);
    if (ec != cudaSuccess) return cuda_fail(ec, "launch_embed_closed");
    p->else_arrays_valid = 1; p->corr_valid = 1;
    g_launches += 2;
  }
  if (!p->tinv_only) {
    cudaError_t e = xb::launch_embed_spots(p->embed_args, g_stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_embed_spots");
    g_launches += 2;
  }
  CK(cudaEventRecord(p->ev_embed[1], g_stream));
  p->embed_status_valid = 1;
  p->embed_timed = 1;
  return 0;
}

}  // namespace

extern "C" {

xpsi_b200_pipeline* xpsi_b200_pipeline_create(const xpsi_b200_pipeline_config* cfg, int max_batch) {
  if (ensure_stream() != 0) return nullptr;
  const xpsi_b200_pipeline_config& c = *cfg;
  if (max_batch < 1 || c.n_components < 1 || c.n_members < c.n_components || c.n_bins < 1 || c.n_bins > xb::marginal_max_bins() ||
      c.n_energies < 5 || (c.n_phases != 1 && (c.n_leaves < 5 || c.n_phases < 5)) ||
      (c.n_phases == 1 && (c.n_members != 1 || c.n_components != 1)) ||
      (c.hot_atm_ext != XPSI_B200_ATM_BB && c.hot_atm_ext != XPSI_B200_ATM_NUM4D) ||
      (c.hot_atm_ext == XPSI_B200_ATM_NUM4D && !c.hot_atmosphere)) {
    g_err = "pipeline_create: invalid configuration";
    return nullptr;
  }
  xpsi_b200_pipeline* p = new xpsi_b200_pipeline();
  p->cfg = c;
  p->max_batch = max_batch;
  p->tinv_only = (c.n_phases == 1);       // Everywhere(time_invariant=True): no hot regions, one phase column
  if (p->tinv_only) { p->cfg.max_rings = 0; p->cfg.max_azi = 0; p->cfg.n_leaves = 0; }
  p->atm = c.hot_atmosphere;
  p->member_component.assign(c.member_component, c.member_component + c.n_members);
  p->cfg.member_component = p->member_component.data();
  const int M = c.n_members, C = c.n_components;
  const size_t B = max_batch, Q = B * M, R = p->cfg.max_rings, A = p->cfg.max_azi;
  std::vector<double> l10E(c.n_energies), l10edges(c.n_in + 1), cyc(c.n_phases);
  for (int i = 0; i < c.n_energies; ++i) l10E[i] = log10(c.energies[i]);
  for (int i = 0; i <= c.n_in; ++i) l10edges[i] = log10(c.energy_edges[i]);
  for (int i = 0; i < c.n_phases; ++i) cyc[i] = c.phases[i] / (2.0 * M_PI);
  std::vector<int> colq(Q);
  for (size_t b = 0; b < B; ++b)
    for (int m = 0; m < M; ++m) colq[b * M + m] = (int)(b * C + c.member_component[m]);
  std::vector<int> icounts((size_t)c.n_chan * c.n_bins);
  for (size_t i = 0; i < icounts.size(); ++i) icounts[i] = (int)c.counts[i];
  Dev<int> d_ic;
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  ok(p->energies.upload(c.energies, c.n_energies)); ok(p->log10E.upload(l10E.data(), c.n_energies));
  if (!p->tinv_only) ok(p->leaves.upload(c.leaves, c.n_leaves));
  ok(p->phases.upload(c.phases, c.n_phases));
  ok(p->phase_cycles.upload(cyc.data(), c.n_phases)); ok(p->log10_edges.upload(l10edges.data(), c.n_in + 1));
  {
    std::vector<int2> span(c.n_in);
    xb::energy_span_table(l10E.data(), c.n_energies, l10edges.data(), c.n_in, span.data());
    ok(p->ei_span.upload(span.data(), span.size()));
  }
  ok(p->response.upload(c.response, (size_t)c.n_chan * c.n_in));
  {
    std::vector<int> kr = response_k_ranges(c.response, c.n_chan, c.n_in, 0, c.n_in);
    ok(p->k_range.upload(kr.data(), kr.size()));
    ok(cudaStreamSynchronize(g_stream));       // kr is a temporary
  }
  ok(p->data_phases.upload(c.data_phases, c.n_bins + 1));
  ok(p->counts.upload(c.counts, (size_t)c.n_chan * c.n_bins)); ok(p->support.upload(c.support, (size_t)c.n_chan * 2));
  ok(p->col_of_q.upload(colq.data(), Q)); ok(d_ic.upload(icounts.data(), icounts.size()));
  ok(p->precomp.alloc(c.n_chan));
  if (e == cudaSuccess) e = xb::launch_precomputation(d_ic.p, c.n_chan, c.n_bins, p->precomp.p, g_stream);
  ok(p->omega.alloc(B)); ok(p->inclination.alloc(B)); ok(p->d_sq.alloc(B)); ok(p->shifts.alloc(B * C));
  ok(p->omega_q.alloc(Q)); ok(p->incl_q.alloc(Q)); ok(p->n_rings.alloc(Q)); ok(p->n_azi.alloc(Q));
  ok(p->cellArea.alloc(Q * R * A)); ok(p->phi.alloc(Q * R * A)); ok(p->theta.alloc(Q * R));
  ok(p->radial.alloc(Q * R)); ok(p->rsr.alloc(Q * R)); ok(p->params.alloc(Q * R * c.n_params));
  ok(p->defl.alloc(Q * R * c.n_rays)); ok(p->calpha.alloc(Q * R * c.n_rays)); ok(p->lag.alloc(Q * R * c.n_rays));
  ok(p->maxd.alloc(Q * R)); ok(p->cgamma.alloc(Q * R));
  // the padded mesh and ray arrays are zeroed once: rings and cells a spot does not use are never read by the
  // integrators, but fetch_embed copies whole arrays to the host
  if (e == cudaSuccess) {
    ok(cudaMemsetAsync(p->n_rings.p, 0, Q * sizeof(int), g_stream)); ok(cudaMemsetAsync(p->n_azi.p, 0, Q * sizeof(int), g_stream));
    ok(cudaMemsetAsync(p->cellArea.p, 0, Q * R * A * sizeof(double), g_stream)); ok(cudaMemsetAsync(p->phi.p, 0, Q * R * A * sizeof(double), g_stream));
    ok(cudaMemsetAsync(p->theta.p, 0, Q * R * sizeof(double), g_stream)); ok(cudaMemsetAsync(p->radial.p, 0, Q * R * sizeof(double), g_stream));
    ok(cudaMemsetAsync(p->params.p, 0, Q * R * c.n_params * sizeof(double), g_stream));
    ok(cudaMemsetAsync(p->defl.p, 0, Q * R * c.n_rays * sizeof(double), g_stream)); ok(cudaMemsetAsync(p->calpha.p, 0, Q * R * c.n_rays * sizeof(double), g_stream));
    ok(cudaMemsetAsync(p->lag.p, 0, Q * R * c.n_rays * sizeof(double), g_stream));
    ok(cudaMemsetAsync(p->maxd.p, 0, Q * R * sizeof(double), g_stream)); ok(cudaMemsetAsync(p->cgamma.p, 0, Q * R * sizeof(double), g_stream));
  }
  ok(p->flux.alloc(Q * c.n_energies * c.n_phases)); ok(p->xin.alloc(B * C * c.n_phases * c.n_in));
  ok(p->folded.alloc(B * C * c.n_chan * c.n_phases)); ok(p->chan_lnL.alloc(B * c.n_chan));
  ok(p->chan_status.alloc(B * c.n_chan)); ok(p->expected.alloc(B * c.n_chan * c.n_bins));
  ok(p->lnL.alloc(B)); ok(p->status_q.alloc(Q)); ok(p->status.alloc(B)); ok(p->work.alloc(4));
  if (!p->tinv_only) {
    xb::AzinvArgs w;
    memset(&w, 0, sizeof(w));
    w.Q = (int)Q; w.n_rings = c.max_rings; w.n_leaves = c.n_leaves; w.hot_atm_ext = c.hot_atm_ext;
    w.n_energies = c.n_energies;
    w.n_img_max = c.image_order_limit > 0 ? c.image_order_limit : xb::kMaxImages;
    if (c.hot_atm_ext == XPSI_B200_ATM_NUM4D) {
      w.hot = p->atm->view;
      xb::azinv_slab_budgets(w.hot, c.energies, c.n_energies, &p->slab_rows_chunk, &p->slab_rows_ring);
      w.slab_rows_ring = p->slab_rows_ring;
    }
    size_t nl, nh, ni, ns;
    xb::azinv_workspace_sizes(w, &nl, &nh, &ni, &ns);
    ok(p->ws_leaf.alloc(nl)); ok(p->ws_hdr.alloc(nh)); ok(p->ws_ihdr.alloc(ni)); ok(p->ws_slab.alloc(ns));
    // header words of rings a spot does not use are read (and ignored) before the ring's image count is looked at
    if (e == cudaSuccess) ok(cudaMemsetAsync(p->ws_ihdr.p, 0, ni * sizeof(int), g_stream));
    // interval-moment tiles (B operands of the tensor-core accumulation): 48 steps per 8-phase tile cover rings
    // whose cells reach ~41 leaf intervals (99 % of the ST-U prior's rings at 100 leaves); wider rings are
    // integrated by the scalar flux kernel, which walks the cells itself
    w.n_phases = c.n_phases; w.tile_cap = 48;
    size_t td, tm, tq;
    xb::azinv_tile_sizes(w, &td, &tm, &tq);
    ok(p->ws_tiles.alloc(td)); ok(p->ws_tmeta.alloc(tm)); ok(p->ws_thdr.alloc(tq));
    ok(p->ws_redo.alloc(Q * c.max_rings * (size_t)((c.n_energies + 7) / 8)));
    ok(p->ws_ovf.alloc(1 + 2 * Q * c.max_rings));
    ok(p->ws_cells.alloc(Q * c.max_rings * 2 * (size_t)c.max_azi));
    p->tile_cap = w.tile_cap;
  }
  for (int i = 0; i < 5; ++i) ok(cudaEventCreate(&p->ev[i]));
  for (int i = 0; i < 2; ++i) ok(cudaEventCreate(&p->ev_embed[i]));
  for (int i = 0; i < 2; ++i) ok(cudaEventCreate(&p->ev_flux[i]));
  ok(cudaMemsetAsync(p->work.p, 0, 4 * sizeof(unsigned long long), g_stream));
  ok(cudaStreamSynchronize(g_stream));
  if (e != cudaSuccess) { cuda_fail(e, "pipeline_create"); delete p; return nullptr; }
  g_launches += 1;
  return p;
}

void xpsi_b200_pipeline_destroy(xpsi_b200_pipeline* p) {
  if (!p) return;
  for (int i = 0; i < 5; ++i) if (p->ev[i]) cudaEventDestroy(p->ev[i]);
  for (int i = 0; i < 2; ++i) if (p->ev_embed[i]) cudaEventDestroy(p->ev_embed[i]);
  for (int i = 0; i < 2; ++i) if (p->ev_flux[i]) cudaEventDestroy(p->ev_flux[i]);
  delete p;
}

int xpsi_b200_pipeline_set_deterministic(xpsi_b200_pipeline* p, int on) {
  if (!p) return fail(XPSI_B200_EINVAL, "null pipeline");
  if (on) {
    const xpsi_b200_pipeline_config& c = p->cfg;
    CK(p->flux_part.alloc((size_t)p->max_batch * c.n_members * c.max_rings * c.n_energies * c.n_phases));
  }
  p->deterministic = on ? 1 : 0;
  return 0;
}

int xpsi_b200_pipeline_set_extras(xpsi_b200_pipeline* p, const xpsi_b200_pipeline_extras* x) {
  if (!p || !x) return fail(XPSI_B200_EINVAL, "null argument");
  const xpsi_b200_pipeline_config& c = p->cfg;
  if (x->beam_opt < 0 || x->beam_opt > 3) return fail(XPSI_B200_EINVAL, "beam_opt must be 0-3");
  if (x->beam_opt != 0 && c.n_params < 7) return fail(XPSI_B200_EINVAL, "beam_opt needs n_params >= 7");
  p->ex = *x;
  p->ex.attenuation = nullptr;
  p->att_power_valid = 0; p->else_arrays_valid = 0; p->corr_valid = 0; p->else_temp_valid = 0;
  if (x->attenuation) CK(p->att_base.upload(x->attenuation, c.n_in));
  else p->att_base.release();
  CK(p->att_power.alloc(p->max_batch));
  if (x->elsewhere) {
    const int n = x->else_sqrt_num_cells, nr = x->else_num_rays;
    if (n < 4 || n > 128 || (n & 1) || nr < 3) return fail(XPSI_B200_EINVAL, "bad elsewhere mesh dimensions");
    if (x->else_atm_ext != XPSI_B200_ATM_BB && x->else_atm_ext != XPSI_B200_ATM_NUM4D)
      return fail(XPSI_B200_EUNSUPPORTED, "else_atm_ext must be 1 (BB) or 2 (Num4D)");
    if (x->else_atm_ext == XPSI_B200_ATM_NUM4D && !x->elsewhere_atmosphere)
      return fail(XPSI_B200_EINVAL, "Num4D elsewhere needs a preloaded atmosphere");
    if (c.n_energies > 256) return fail(XPSI_B200_EINVAL, "elsewhere supports at most 256 energies");
    const size_t B = p->max_batch, Q = B * c.n_members;
    CK(p->x_temp.alloc(B)); CK(p->x_area.alloc(B)); CK(p->x_radial.alloc(B * n)); CK(p->x_rsr.alloc(B * n));
    CK(p->x_theta.alloc(B * n * n)); CK(p->x_phi.alloc(B * n * n)); CK(p->x_params.alloc(B * n * n * 2));
    CK(p->x_defl.alloc(B * n * nr)); CK(p->x_calpha.alloc(B * n * nr)); CK(p->x_maxd.alloc(B * n));
    CK(p->x_cgamma.alloc(B * n)); CK(p->x_maxAlpha.alloc(B * n)); CK(p->x_grav.alloc(B * n));
    CK(p->x_flux.alloc(B * c.n_energies)); CK(p->x_nrings.alloc(B)); CK(p->x_status.alloc(B));
    CK(p->corr.alloc(Q * c.max_rings * c.n_params));
    if (x->else_atm_ext == XPSI_B200_ATM_NUM4D) {
      const xb::AtmTable& t = x->elsewhere_atmosphere->view;
      p->else_slab_rows = xb::tinv_slab_rows(t, c.energies, c.n_energies);
      xb::azinv_slab_budgets(t, c.energies, c.n_energies, &p->slab2_chunk, &p->slab2_ring);
      // the hot and the correction slabs share the per-ring stride: size both workspaces for the larger one
      const int rows = p->slab2_ring > p->slab_rows_ring ? p->slab2_ring : p->slab_rows_ring;
      const int nmu = (c.hot_atm_ext == XPSI_B200_ATM_NUM4D && p->atm->view.nmu > t.nmu) ? p->atm->view.nmu : t.nmu;
      CK(p->ws_slab2.alloc(Q * c.max_rings * (size_t)nmu * rows));
      if (c.hot_atm_ext == XPSI_B200_ATM_NUM4D && rows > p->slab_rows_ring)
        CK(p->ws_slab.alloc(Q * c.max_rings * (size_t)nmu * rows));
    }
  }
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}


// ---- further signals of the same photosphere (xpsi/Likelihood.py:346-420, :494-500) --------------------------
int xpsi_b200_pipeline_add_signal(xpsi_b200_pipeline* p, const xpsi_b200_signal_config* sc) {
  if (!p || !sc) return fail(XPSI_B200_EINVAL, "null argument");
  const xpsi_b200_pipeline_config& c = p->cfg;
  if (sc->n_in < 1 || sc->n_chan < 1 || sc->n_bins < 1 || !sc->energy_edges || !sc->response || !sc->data_phases ||
      !sc->counts || !sc->support)
    return fail(XPSI_B200_EINVAL, "add_signal: incomplete signal configuration");
  if (sc->n_bins > xb::marginal_max_bins())
    return fail(XPSI_B200_EUNSUPPORTED, "add_signal: more data phase bins than the marginal kernel covers");
  std::unique_ptr<xpsi_b200_pipeline::SignalPart> sp(new xpsi_b200_pipeline::SignalPart());
  sp->n_in = sc->n_in; sp->n_chan = sc->n_chan; sp->n_bins = sc->n_bins; sp->allow_negative = sc->allow_negative;
  sp->exposure_time = sc->exposure_time; sp->epsilon = sc->epsilon; sp->sigmas = sc->sigmas; sp->llzero = sc->llzero;
  sp->slim = sc->slim;
  std::vector<double> l10E(c.n_energies), l10edges(sc->n_in + 1);
  for (int i = 0; i < c.n_energies; ++i) l10E[i] = log10(c.energies[i]);
  for (int i = 0; i <= sc->n_in; ++i) l10edges[i] = log10(sc->energy_edges[i]);
  std::vector<int2> span(sc->n_in);
  xb::energy_span_table(l10E.data(), c.n_energies, l10edges.data(), sc->n_in, span.data());
  std::vector<int> kr = response_k_ranges(sc->response, sc->n_chan, sc->n_in, 0, sc->n_in);
  std::vector<int> icounts((size_t)sc->n_chan * sc->n_bins);
  for (size_t i = 0; i < icounts.size(); ++i) icounts[i] = (int)sc->counts[i];
  Dev<int> d_ic;
  CK(sp->log10_edges.upload(l10edges.data(), l10edges.size())); CK(sp->ei_span.upload(span.data(), span.size()));
  CK(sp->response.upload(sc->response, (size_t)sc->n_chan * sc->n_in)); CK(sp->k_range.upload(kr.data(), kr.size()));
  CK(sp->data_phases.upload(sc->data_phases, sc->n_bins + 1));
  CK(sp->counts.upload(sc->counts, (size_t)sc->n_chan * sc->n_bins)); CK(sp->support.upload(sc->support, (size_t)sc->n_chan * 2));
  if (sc->attenuation) CK(sp->att_base.upload(sc->attenuation, sc->n_in));
  CK(d_ic.upload(icounts.data(), icounts.size())); CK(sp->precomp.alloc(sc->n_chan));
  cudaError_t e = xb::launch_precomputation(d_ic.p, sc->n_chan, sc->n_bins, sp->precomp.p, g_stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_precomputation");
  const size_t B = p->max_batch, C = c.n_components;
  CK(sp->xin.alloc(B * C * c.n_phases * sc->n_in)); CK(sp->folded.alloc(B * C * sc->n_chan * c.n_phases));
  CK(sp->chan_lnL.alloc(B * sc->n_chan)); CK(sp->chan_status.alloc(B * sc->n_chan));
  CK(sp->expected.alloc(B * sc->n_chan * sc->n_bins)); CK(sp->lnL.alloc(B));
  CK(p->att_power.alloc(p->max_batch));
  CK(cudaStreamSynchronize(g_stream));          // the host temporaries go away
  g_launches += 1;
  p->more.push_back(std::move(sp));
  p->sig_shift_valid = 0;
  return (int)p->more.size();                    // index of the new signal (the configuration's signal is 0)
}

int xpsi_b200_pipeline_n_signals(xpsi_b200_pipeline* p) { return p ? 1 + (int)p->more.size() : 0; }

int xpsi_b200_pipeline_upload_signal_shifts(xpsi_b200_pipeline* p, int B, const double* shifts) {
  if (!p || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  if (!shifts) { p->sig_shift_valid = 0; return 0; }
  if (p->extras_B != B) { p->att_power_valid = 0; p->else_arrays_valid = 0; p->corr_valid = 0; p->else_temp_valid = 0; }
  p->extras_B = B;
  const size_t S = 1 + p->more.size();
  CK(p->sig_shift.upload(shifts, (size_t)B * S));
  CK(p->sig_shifts_s.alloc((size_t)p->max_batch * p->cfg.n_components));
  CK(cudaStreamSynchronize(g_stream));
  p->sig_shift_valid = 1;
  return 0;
}

int xpsi_b200_pipeline_fetch_signal(xpsi_b200_pipeline* p, int signal, int B, double* folded, double* expected, double* lnL) {
  if (!p || B < 1 || B > p->max_batch || signal < 0 || signal > (int)p->more.size())
    return fail(XPSI_B200_EINVAL, "bad signal index or batch size");
  const xpsi_b200_pipeline_config& c = p->cfg;
  if (signal == 0) {
    if (folded) CK(p->folded.download(folded, (size_t)B * c.n_components * c.n_chan * c.n_phases));
    if (expected) CK(p->expected.download(expected, (size_t)B * c.n_chan * c.n_bins));
    if (lnL) {
      // the configuration's signal alone: the joint value minus the other signals' terms is not kept; recompute
      // from the channel terms (ordered sum, as k_sum_channels)
      std::vector<double> ch((size_t)B * c.n_chan);
      CK(p->chan_lnL.download(ch.data(), ch.size()));
      CK(cudaStreamSynchronize(g_stream));
      for (int b = 0; b < B; ++b) { double s = 0.0; for (int k = 0; k < c.n_chan; ++k) s += ch[(size_t)b * c.n_chan + k]; lnL[b] = s; }
    }
  } else {
    xpsi_b200_pipeline::SignalPart* sp = p->more[signal - 1].get();
    if (folded) CK(sp->folded.download(folded, (size_t)B * c.n_components * sp->n_chan * c.n_phases));
    if (expected) CK(sp->expected.download(expected, (size_t)B * sp->n_chan * sp->n_bins));
    if (lnL) CK(sp->lnL.download(lnL, B));
  }
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_pipeline_sweep_upload_signal_shifts(xpsi_b200_pipeline* p, long long N, const double* shifts) {
  if (!p || N < 1 || (size_t)N != p->store.N) return fail(XPSI_B200_EINVAL, "signal shifts must match the uploaded sweep");
  if (!shifts) { p->store.has_sig_shift = 0; return 0; }
  const size_t S = 1 + p->more.size();
  CK(p->store.sig_shift.upload(shifts, (size_t)N * S));
  CK(p->sig_shifts_s.alloc((size_t)p->max_batch * p->cfg.n_components));
  CK(cudaStreamSynchronize(g_stream));
  p->store.has_sig_shift = 1;
  return 0;
}

int xpsi_b200_pipeline_upload_extras(xpsi_b200_pipeline* p, int B, const xpsi_b200_batch_extras* h) {
  if (!p || !h || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  const xpsi_b200_pipeline_config& c = p->cfg;
  if (p->extras_B != B) { p->att_power_valid = 0; p->else_arrays_valid = 0; p->corr_valid = 0; p->else_temp_valid = 0; p->sig_shift_valid = 0; }
  p->extras_B = B;
  if (h->att_power) { CK(p->att_power.upload(h->att_power, B)); p->att_power_valid = 1; }
  if (p->ex.elsewhere) {
    const size_t n = p->ex.else_sqrt_num_cells, nr = p->ex.else_num_rays;
    if (h->else_temperature) { CK(p->x_temp.upload(h->else_temperature, B)); p->else_temp_valid = 1; }
    if (h->else_cellArea) {
      if (!h->else_radial || !h->else_r_s_over_r || !h->else_theta || !h->else_phi || !h->else_srcParams ||
          !h->else_deflection || !h->else_cos_alpha || !h->else_maxDeflection || !h->else_cos_gamma)
        return fail(XPSI_B200_EINVAL, "incomplete elsewhere arrays");
      CK(p->x_area.upload(h->else_cellArea, B)); CK(p->x_radial.upload(h->else_radial, B * n));
      CK(p->x_rsr.upload(h->else_r_s_over_r, B * n)); CK(p->x_theta.upload(h->else_theta, B * n * n));
      CK(p->x_phi.upload(h->else_phi, B * n * n)); CK(p->x_params.upload(h->else_srcParams, B * n * n * 2));
      CK(p->x_defl.upload(h->else_deflection, B * n * nr)); CK(p->x_calpha.upload(h->else_cos_alpha, B * n * nr));
      CK(p->x_maxd.upload(h->else_maxDeflection, B * n)); CK(p->x_cgamma.upload(h->else_cos_gamma, B * n));
      p->else_arrays_valid = 1;
    }
    if (h->correction_srcParams) {
      CK(p->corr.upload(h->correction_srcParams, (size_t)B * c.n_members * c.max_rings * c.n_params));
      p->corr_valid = 1;
    }
  }
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_pipeline_fetch_elsewhere(xpsi_b200_pipeline* p, int B, double* spectrum) {
  if (!p || !p->ex.elsewhere || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "no elsewhere spectrum");
  CK(p->x_flux.download(spectrum, (size_t)B * p->cfg.n_energies));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_pipeline_upload(xpsi_b200_pipeline* p, int B, const xpsi_b200_batch* h) {
  if (!p || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  int rc = pipeline_upload(p, B, h);
  if (rc) return rc;
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_pipeline_eval_resident(xpsi_b200_pipeline* p, int B) {
  if (!p || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  return pipeline_run(p, B);
}

int xpsi_b200_pipeline_download(xpsi_b200_pipeline* p, int B, double* lnL, int* status) {
  if (!p || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  CK(p->lnL.download(lnL, B));
  CK(p->status.download(status, B));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_pipeline_eval(xpsi_b200_pipeline* p, int B, const xpsi_b200_batch* h, double* lnL, int* status) {
  if (!p || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  int rc = pipeline_upload(p, B, h);
  if (rc) return rc;
  rc = pipeline_run(p, B);
  if (rc) return rc;
  return xpsi_b200_pipeline_download(p, B, lnL, status);
}

// device-side embed arguments over the pipeline's own per-batch arrays
static int set_embed_args(xpsi_b200_pipeline* p, int B, double mode_frequency, int num_cells, int min_sqrt,
                          int max_sqrt, bool hole, bool partner, bool extra, bool member_cells = false) {
  const xpsi_b200_pipeline_config& c = p->cfg;
  const size_t M = c.n_members;
  CK(p->e_maxAlpha.alloc((size_t)p->max_batch * M * c.max_rings));
  xb::EmbedArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.M = (int)M; a.max_rings = c.max_rings; a.max_azi = c.max_azi; a.n_rays = c.n_rays; a.n_params = c.n_params;
  a.num_cells = num_cells; a.min_sqrt = min_sqrt; a.max_sqrt = max_sqrt;
  a.mode_frequency = mode_frequency;
  a.R_eq = p->e_Req.p; a.r_s = p->e_rs.p; a.epsilon = p->e_eps.p; a.zeta = p->e_zeta.p;
  a.colatitude = p->e_colat.p; a.ang_radius = p->e_rad.p; a.temperature = p->e_temp.p; a.phi_shift = p->e_phish.p;
  if (hole) { a.hole_radius = p->e_hrad.p; a.hole_colatitude = p->e_hcolat.p; a.hole_azimuth = p->e_hazi.p; }
  if (partner) { a.partner = p->e_partner.p; a.is_cede = p->e_iscede.p; }
  if (extra) a.extra_params = p->e_extra.p;
  if (member_cells) a.member_cells = p->e_member_cells.p;
  a.n_rings = p->n_rings.p; a.n_azi = p->n_azi.p; a.cellArea = p->cellArea.p; a.phi = p->phi.p; a.theta = p->theta.p;
  a.radial = p->radial.p; a.r_s_over_r = p->rsr.p; a.srcParams = p->params.p; a.cos_gamma = p->cgamma.p;
  a.maxAlpha = p->e_maxAlpha.p; a.deflection = p->defl.p; a.cos_alpha = p->calpha.p; a.lag = p->lag.p;
  a.maxDeflection = p->maxd.p; a.status = p->status.p;
  p->embed_args = a; p->embed_ready = 1;
  return 0;
}

static int check_partner(const xpsi_b200_spot_batch* h, size_t M) {
  if (!h->is_cede) return fail(XPSI_B200_EINVAL, "partner needs is_cede");
  for (size_t m = 0; m < M; ++m)
    if (h->partner[m] >= (int)M || h->partner[m] == (int)m) return fail(XPSI_B200_EINVAL, "bad partner index");
  return 0;
}

int xpsi_b200_pipeline_embed_spots(xpsi_b200_pipeline* p, int B, const xpsi_b200_spot_batch* h) {
  if (!p || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  const xpsi_b200_pipeline_config& c = p->cfg;
  if (c.n_params < 2) return fail(XPSI_B200_EUNSUPPORTED, "spot embed needs n_params >= 2 (log T, log g, ...)");
  const size_t M = c.n_members, Q = (size_t)B * M;
  CK(p->omega.upload(h->omega, B)); CK(p->inclination.upload(h->inclination, B)); CK(p->d_sq.upload(h->d_sq, B));
  CK(p->shifts.upload(h->phase_shifts, (size_t)B * c.n_components));
  CK(p->e_Req.upload(h->R_eq, B)); CK(p->e_rs.upload(h->r_s, B)); CK(p->e_eps.upload(h->epsilon, B));
  CK(p->e_zeta.upload(h->zeta, B)); CK(p->e_colat.upload(h->colatitude, Q)); CK(p->e_rad.upload(h->ang_radius, Q));
  CK(p->e_temp.upload(h->temperature, Q)); CK(p->e_phish.upload(h->phi_shift, Q));
  if (h->hole_radius) {
    if (!h->hole_colatitude || !h->hole_azimuth) return fail(XPSI_B200_EINVAL, "incomplete hole arrays");
    CK(p->e_hrad.upload(h->hole_radius, Q)); CK(p->e_hcolat.upload(h->hole_colatitude, Q));
    CK(p->e_hazi.upload(h->hole_azimuth, Q));
  }
  if (h->extra_params && c.n_params > 2) CK(p->e_extra.upload(h->extra_params, Q * (c.n_params - 2)));
  if (h->partner) {
    int rc = check_partner(h, M);
    if (rc) return rc;
    CK(p->e_partner.upload(h->partner, M)); CK(p->e_iscede.upload(h->is_cede, M));
  }
  if (h->member_cells) CK(p->e_member_cells.upload(h->member_cells, 3 * M));
  CK(cudaMemsetAsync(p->status.p, 0, B * sizeof(int), g_stream));
  int rc = set_embed_args(p, B, h->mode_frequency, h->num_cells, h->min_sqrt_num_cells, h->max_sqrt_num_cells,
                          h->hole_radius != nullptr, h->partner != nullptr, h->extra_params && c.n_params > 2,
                          h->member_cells != nullptr);
  if (rc) return rc;
  return pipeline_embed_launch(p);
}

// ---- sweep: N parameter vectors resident on the device, evaluated in blocks of <= max_batch ------------------
int xpsi_b200_pipeline_sweep_upload(xpsi_b200_pipeline* p, long long N, const xpsi_b200_spot_batch* h,
                                    const double* att_power, const double* else_temperature) {
  if (!p || !h || N < 1) return fail(XPSI_B200_EINVAL, "bad sweep size");
  const xpsi_b200_pipeline_config& c = p->cfg;
  if (c.n_params < 2) return fail(XPSI_B200_EUNSUPPORTED, "spot embed needs n_params >= 2 (log T, log g, ...)");
  if (p->ex.elsewhere && !else_temperature) return fail(XPSI_B200_EINVAL, "elsewhere is enabled: else_temperature[N] is required");
  auto& s = p->store;
  const size_t n = (size_t)N, M = c.n_members, Q = n * M;
  CK(s.omega.upload(h->omega, n)); CK(s.incl.upload(h->inclination, n)); CK(s.d_sq.upload(h->d_sq, n));
  CK(s.shifts.upload(h->phase_shifts, n * c.n_components));
  CK(s.Req.upload(h->R_eq, n)); CK(s.rs.upload(h->r_s, n)); CK(s.eps.upload(h->epsilon, n)); CK(s.zeta.upload(h->zeta, n));
  CK(s.colat.upload(h->colatitude, Q)); CK(s.rad.upload(h->ang_radius, Q)); CK(s.temp.upload(h->temperature, Q));
  CK(s.phish.upload(h->phi_shift, Q));
  s.has_hole = h->hole_radius != nullptr;
  if (s.has_hole) {
    if (!h->hole_colatitude || !h->hole_azimuth) return fail(XPSI_B200_EINVAL, "incomplete hole arrays");
    CK(s.hrad.upload(h->hole_radius, Q)); CK(s.hcolat.upload(h->hole_colatitude, Q)); CK(s.hazi.upload(h->hole_azimuth, Q));
  }
  s.has_extra = h->extra_params && c.n_params > 2;
  if (s.has_extra) CK(s.extra.upload(h->extra_params, Q * (c.n_params - 2)));
  s.has_partner = h->partner != nullptr;
  if (s.has_partner) {
    int rc = check_partner(h, M);
    if (rc) return rc;
    CK(p->e_partner.upload(h->partner, M)); CK(p->e_iscede.upload(h->is_cede, M));
  }
  s.has_member_cells = h->member_cells != nullptr;
  if (s.has_member_cells) CK(p->e_member_cells.upload(h->member_cells, 3 * M));
  s.has_att = att_power != nullptr;
  if (s.has_att) CK(s.att_power.upload(att_power, n));
  s.has_else = else_temperature != nullptr;
  if (s.has_else) CK(s.else_temp.upload(else_temperature, n));
  s.has_sig_shift = 0;
  CK(s.lnL.alloc(n)); CK(s.status.alloc(n));
  s.mode_frequency = h->mode_frequency; s.num_cells = h->num_cells;
  s.min_sqrt = h->min_sqrt_num_cells; s.max_sqrt = h->max_sqrt_num_cells;
  s.N = n;
  CK(cudaStreamSynchronize(g_stream));          // the host arrays may go away after the call
  return 0;
}

int xpsi_b200_pipeline_sweep_run(xpsi_b200_pipeline* p, long long first, long long count) {
  if (!p || first < 0 || count < 1 || (size_t)(first + count) > p->store.N)
    return fail(XPSI_B200_EINVAL, "sweep range outside the uploaded parameter vectors");
  const xpsi_b200_pipeline_config& c = p->cfg;
  auto& s = p->store;
  const size_t M = c.n_members, Cn = c.n_components;
  // rows of a store array starting at element `at` -> the pipeline's per-batch array (sized for max_batch once);
  // `per` = elements per parameter vector
  auto d2d = [&](Dev<double>& dst, const Dev<double>& src, size_t at, size_t rows, size_t per) -> cudaError_t {
    cudaError_t e = dst.alloc(per * (size_t)p->max_batch);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(dst.p, src.p + at, rows * per * sizeof(double), cudaMemcpyDeviceToDevice, g_stream);
  };
  for (long long off = first; off < first + count; off += p->max_batch) {
    const size_t B = (size_t)((first + count - off) < p->max_batch ? (first + count - off) : p->max_batch);
    const size_t o = (size_t)off;
    const size_t X = (size_t)(c.n_params > 2 ? c.n_params - 2 : 0);
    CK(d2d(p->omega, s.omega, o, B, 1)); CK(d2d(p->inclination, s.incl, o, B, 1)); CK(d2d(p->d_sq, s.d_sq, o, B, 1));
    CK(d2d(p->shifts, s.shifts, o * Cn, B, Cn));
    CK(d2d(p->e_Req, s.Req, o, B, 1)); CK(d2d(p->e_rs, s.rs, o, B, 1)); CK(d2d(p->e_eps, s.eps, o, B, 1));
    CK(d2d(p->e_zeta, s.zeta, o, B, 1));
    CK(d2d(p->e_colat, s.colat, o * M, B, M)); CK(d2d(p->e_rad, s.rad, o * M, B, M));
    CK(d2d(p->e_temp, s.temp, o * M, B, M)); CK(d2d(p->e_phish, s.phish, o * M, B, M));
    if (s.has_hole) {
      CK(d2d(p->e_hrad, s.hrad, o * M, B, M)); CK(d2d(p->e_hcolat, s.hcolat, o * M, B, M));
      CK(d2d(p->e_hazi, s.hazi, o * M, B, M));
    }
    if (s.has_extra) CK(d2d(p->e_extra, s.extra, o * M * X, B, M * X));
    p->extras_B = (int)B;
    if (s.has_att) { CK(d2d(p->att_power, s.att_power, o, B, 1)); p->att_power_valid = 1; }
    if (s.has_else) { CK(d2d(p->x_temp, s.else_temp, o, B, 1)); p->else_temp_valid = 1; }
    if (s.has_sig_shift) { const size_t S = 1 + p->more.size(); CK(d2d(p->sig_shift, s.sig_shift, o * S, B, S)); p->sig_shift_valid = 1; }
    CK(cudaMemsetAsync(p->status.p, 0, B * sizeof(int), g_stream));
    int rc = set_embed_args(p, (int)B, s.mode_frequency, s.num_cells, s.min_sqrt, s.max_sqrt, s.has_hole,
                            s.has_partner, s.has_extra, s.has_member_cells);
    if (rc) return rc;
    rc = pipeline_embed_launch(p);
    if (rc) return rc;
    rc = pipeline_run(p, (int)B);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s.lnL.p + o, p->lnL.p, B * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    CK(cudaMemcpyAsync(s.status.p + o, p->status.p, B * sizeof(int), cudaMemcpyDeviceToDevice, g_stream));
  }
  return 0;
}

int xpsi_b200_pipeline_sweep_download(xpsi_b200_pipeline* p, long long first, long long count, double* lnL, int* status) {
  if (!p || first < 0 || count < 1 || (size_t)(first + count) > p->store.N)
    return fail(XPSI_B200_EINVAL, "sweep range outside the uploaded parameter vectors");
  g_d2h += (long long)(count * (sizeof(double) + sizeof(int)));
  CK(cudaMemcpyAsync(lnL, p->store.lnL.p + first, count * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(status, p->store.status.p + first, count * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_pipeline_sweep_results(xpsi_b200_pipeline* p, double** lnL, int** status) {
  if (!p || !p->store.N) return fail(XPSI_B200_EINVAL, "no sweep uploaded");
  if (lnL) *lnL = p->store.lnL.p;
  if (status) *status = p->store.status.p;
  return 0;
}

int xpsi_b200_pipeline_eval_spots_resident(xpsi_b200_pipeline* p, int B) {
  if (!p || !p->embed_ready || B != p->embed_args.B) return fail(XPSI_B200_EINVAL, "no resident spot batch of this size");
  CK(cudaMemsetAsync(p->status.p, 0, B * sizeof(int), g_stream));
  int rc = pipeline_embed_launch(p);
  if (rc) return rc;
  return pipeline_run(p, B);
}

int xpsi_b200_pipeline_eval_spots(xpsi_b200_pipeline* p, int B, const xpsi_b200_spot_batch* h, double* lnL,
                                  int* status) {
  int rc = xpsi_b200_pipeline_embed_spots(p, B, h);
  if (rc) return rc;
  rc = pipeline_run(p, B);
  if (rc) return rc;
  return xpsi_b200_pipeline_download(p, B, lnL, status);
}

int xpsi_b200_pipeline_fetch_embed(xpsi_b200_pipeline* p, int B, int* n_rings, double* cellArea, double* phi,
                                   double* theta, double* radial, double* srcParams, double* cos_gamma,
                                   double* deflection, double* cos_alpha, double* lag, double* maxDeflection) {
  if (!p || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  const xpsi_b200_pipeline_config& c = p->cfg;
  const size_t Q = (size_t)B * c.n_members, R = c.max_rings, A = c.max_azi;
  if (n_rings) CK(p->n_rings.download(n_rings, Q));
  if (cellArea) CK(p->cellArea.download(cellArea, Q * R * A));
  if (phi) CK(p->phi.download(phi, Q * R * A));
  if (theta) CK(p->theta.download(theta, Q * R));
  if (radial) CK(p->radial.download(radial, Q * R));
  if (srcParams) CK(p->params.download(srcParams, Q * R * c.n_params));
  if (cos_gamma) CK(p->cgamma.download(cos_gamma, Q * R));
  if (deflection) CK(p->defl.download(deflection, Q * R * c.n_rays));
  if (cos_alpha) CK(p->calpha.download(cos_alpha, Q * R * c.n_rays));
  if (lag) CK(p->lag.download(lag, Q * R * c.n_rays));
  if (maxDeflection) CK(p->maxd.download(maxDeflection, Q * R));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_pipeline_fetch(xpsi_b200_pipeline* p, int B, double* flux, double* folded, double* expected) {
  if (!p || B < 1 || B > p->max_batch) return fail(XPSI_B200_EINVAL, "bad batch size");
  const xpsi_b200_pipeline_config& c = p->cfg;
  if (flux) CK(p->flux.download(flux, (size_t)B * c.n_members * c.n_energies * c.n_phases));
  if (folded) CK(p->folded.download(folded, (size_t)B * c.n_components * c.n_chan * c.n_phases));
  if (expected) CK(p->expected.download(expected, (size_t)B * c.n_chan * c.n_bins));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int xpsi_b200_pipeline_work_counters(xpsi_b200_pipeline* p, int enable, unsigned long long out[4]) {
  if (!p) return fail(XPSI_B200_EINVAL, "null pipeline");
  if (out && p->count_work) {
    CK(cudaMemcpyAsync(out, p->work.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
  }
  if (enable) CK(cudaMemsetAsync(p->work.p, 0, 4 * sizeof(unsigned long long), g_stream));   // counts accumulate from here
  p->count_work = enable;
  return 0;
}

int xpsi_b200_pipeline_stage_ms(xpsi_b200_pipeline* p, float ms[6]) {
  if (!p) return fail(XPSI_B200_EINVAL, "null pipeline");
  CK(cudaEventSynchronize(p->ev[4]));
  for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&ms[i], p->ev[i], p->ev[i + 1]));
  ms[4] = 0.f;
  if (p->embed_timed) CK(cudaEventElapsedTime(&ms[4], p->ev_embed[0], p->ev_embed[1]));
  CK(cudaEventElapsedTime(&ms[5], p->ev_flux[0], p->ev_flux[1]));      // the flux kernel alone (inside integrate)
  return 0;
}

}  // extern "C"
