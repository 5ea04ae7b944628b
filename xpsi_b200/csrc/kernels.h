// Kernel argument blocks and launchers (internal; the public surface is
// include/xpsi_b200.h).  All pointers are DEVICE pointers.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace xb {

constexpr int kMaxImages = 6;

// a5: the reference's _preloaded struct (surface_radiation_field/preload.pxd:3-14)
struct AtmTable {
  const double* logT; const double* logg; const double* mu; const double* logE;
  const double* buf;                 // C-order [T][g][mu][E]
  int nT, ng, nmu, nE;
  double min_dlogE;                  // smallest spacing of the logE axis (host-computed)
  const double* mu_invden;           // [nmu-3][4] inverse Lagrange denominators per base node (host-computed)
  const double* E_invden;            // [nE-3][4]
};

// a1: integrator_for_azimuthal_invariance.integrate, batched over Q member instances.
// Per-instance arrays are padded to (n_rings, n_azi); true sizes in n_rings_q/n_azi_q
// (nullptr => every instance uses the padded size).
struct AzinvArgs {
  int Q, n_rings, n_azi, n_rays, n_energies, n_leaves, n_phases, n_params;
  const int* n_rings_q; const int* n_azi_q;
  const double* omega; const double* inclination;                  // [Q]
  const double* cellArea; const double* phi;                       // [Q][R][A]
  const double* theta; int theta_ring_stride;                      // theta[ring*stride] = ring colatitude
  const double* radial; const double* r_s_over_r;                  // [Q][R]
  const double* srcParams; int params_per_cell;                    // [Q][R][A][n] or [Q][R][n]
  const int* radiates;                                             // [Q][R][A]
  const double* corrParams;          // nullptr, or the layout of srcParams: elsewhere correction (pyx:257-268)
  AtmTable els; int else_atm_ext;    // elsewhere atmosphere for the correction (elsewhere_wrapper.pyx:23-81)
  const double* deflection; const double* cos_alpha; const double* lag;   // [Q][R][N_R]
  const double* maxDeflection; const double* cos_gamma;            // [Q][R]
  const double* energies; const double* leaves; const double* phases;
  const double* log10_energies;                                    // [N_E]
  AtmTable hot;
  int hot_atm_ext;                   // 1 blackbody, 2 Num4D (hot_wrapper.pyx:72-252)
  int image_order_limit;             // 0 => infer ceil(maxDeflection/pi)
  int beam_opt;                      // hot_wrapper.pyx:155-172: 0 none, 1, 2 (parameters VEC[2..5])
  double R_in;                       // inner disc radius [m]; >= 1e6 means no disc (pyx:390-396)
  int n_img_max;
  int phase_interp;                  // tools/core.pyx:21  0 Akima(periodic) 1 Steffen
  int slab_ne_max;                   // Num4D: energy rows a flux CTA may hold in shared memory
  int slab_rows_ring;                // Num4D: energy rows per ring in the slab workspace
  int scale_by_energy;               // apply flux /= E keV (pyx:610-612)
  int general;                       // set by the launchers: 1 = geometry for integrator.pyx (no azimuthal invariance)
  double* flux;                      // [Q][N_E][N_P], zero-initialised by the caller
  int* status;                       // [Q]
  // workspaces (sizes from azinv_workspace_sizes)
  double* ws_leaf;                   // [Q][n_rings][n_img_max][4][N_L]   geometry -> flux
  double* ws_hdr; int* ws_ihdr;      // per-ring headers (+ per-chunk row table behind ws_ihdr)
  int* ws_chunk;                     // set by the launcher
  double* ws_slab;                   // Num4D: [Q][n_rings][nmu][slab_rows_ring]
  double* ws_slab2;                  // same for a Num4D elsewhere correction
  // optional: interval moments of the cell walk, shared by a ring's energy chunks (azinv_moment_sizes)
  double* ws_mom; int2* ws_meta; int* ws_cnt; int mom_cap;
  double* ws_cells;                  // with ws_mom: [Q][n_rings][2][n_azi] compact (azimuth, area) lists of the radiating cells
  unsigned long long* work;          // nullptr or [4]: H half-leaf visits, V visible leaves,
                                     //   RI (ring,image) pairs reaching the phase stage, K radiating cells over RI
  // optional (instead of ws_mom): the same interval moments as dense per-(image, 8-phase tile) B operands of the
  // tensor-core accumulation stage (azinv_tile_sizes): ws_tiles [slot][tile][tile_cap][32] doubles (lane = phase-in-
  // tile * 4 + Hermite-basis moment H00, H01, H10, H11), ws_tmeta [slot][tile][tile_cap][8] ints (cell range of each phase in that interval),
  // ws_thdr [slot][tile] = (first interval, steps or -1 when the tile needs more than tile_cap steps)
  double* ws_tiles; int* ws_tmeta; int2* ws_thdr; int tile_cap;
  int* ws_redo;                      // with ws_tiles: [Q][n_rings][n_chunks], set by the tensor-core kernel for a chunk it
                                     // hands back to the scalar kernel (zeroed by the launcher)
  int force_redo;                    // test hook (XPSI_B200_FORCE_REDO=1): every chunk is handed back
  int* ws_ovf;                       // with ws_tiles: [1 + 2 Q n_rings]: count, compact list of the rings the scalar kernel
                                     // has to visit (overflowed tiles or a handed-back chunk), per-ring "listed" flags
  int ovf_list_mode;                 // set by the launcher: the scalar kernel strides over that list
  // optional: deterministic two-stage ring reduction.  Every (ring, energy chunk) CTA stores its sum into
  // flux_part [Q][n_rings][N_E][N_P] and k_azinv_reduce_rings adds the lit rings in index order (no atomics)
  double* flux_part;
  cudaEvent_t ev_flux[2];            // host side only: recorded around the flux kernel when non-null (roofline timing)
};
void azinv_tile_sizes(const AzinvArgs& a, size_t* tile_doubles, size_t* tmeta_ints, size_t* thdr_int2);
cudaError_t launch_integrate_azinv(AzinvArgs a, cudaStream_t stream);
// a2: cellmesh/integrator.pyx (no azimuthal invariance); same argument block, parameters read per cell.
// Uses ws_leaf / ws_hdr / ws_ihdr only; slab_ne_max = energy rows a CTA may hold (general_slab_rows)
cudaError_t launch_integrate_general(AzinvArgs a, cudaStream_t stream);
int general_slab_rows(const AtmTable& t, const double* host_energies, int n_energies);
void azinv_workspace_sizes(const AzinvArgs& a, size_t* leaf_doubles, size_t* hdr_doubles, size_t* ihdr_ints,
                           size_t* slab_doubles);
void azinv_moment_sizes(const AzinvArgs& a, size_t* mom_doubles, size_t* meta_int2, size_t* cnt_ints);
void azinv_slab_budgets(const AtmTable& t, const double* host_energies, int n_energies, int* rows_chunk,
                        int* rows_ring);

// a3: integrator_for_time_invariance.integrate, batched over Q instances (square meshes n x n)
struct TinvArgs {
  int Q, sqrt_numPix, n_rays, n_energies, n_params;
  const double* omega; const double* inclination; const double* cellArea;       // [Q]
  const double* radial; const double* r_s_over_r;                               // [Q][n]
  const double* theta; const double* phi;                                       // [Q][n][n]
  const double* srcParams;                                                      // [Q][n][n][n_params]
  const double* deflection; const double* cos_alpha;                            // [Q][n][N_R]
  const double* maxDeflection; const double* cos_gamma;                         // [Q][n]
  const double* energies;
  AtmTable atm; int atm_ext;         // elsewhere_wrapper.pyx:23-81: 1 blackbody, 2 Num4D
  int image_order_limit;
  int slab_rows;                     // Num4D: energy rows budgeted for the per-ring slab
  double* flux;                      // [Q][N_E], zero-initialised by the caller
  int* status;                       // [Q]
};
cudaError_t launch_integrate_tinv(TinvArgs a, cudaStream_t stream);
int tinv_slab_rows(const AtmTable& t, const double* host_energies, int n_energies);

// f1: embed of circular hot spots (mesh + rays) straight into the integrator's batch arrays
struct EmbedArgs {
  int B, M, max_rings, max_azi, n_rays, n_params;
  int num_cells; double min_sqrt, max_sqrt;      // HotRegion(sqrt_num_cells^2, min_sqrt_num_cells, max_sqrt_num_cells)
  double mode_frequency;                          // Photosphere['mode_frequency'] (HotRegion.py:887)
  const double* R_eq; const double* r_s; const double* epsilon; const double* zeta;      // [B] Spacetime.py:110-188
  const double* colatitude; const double* ang_radius; const double* temperature;        // [B*M]
  const double* phi_shift;                        // [B*M] added to cell azimuths (pi if antiphased)
  // optional (nullptr: plain circular spots): the region masking each member -- the omission hole of a
  // superseding member, or the superseding region inside a ceding member (HotRegion.py:819-865) -- and the
  // pairing of superseding / ceding members that share num_cells (mesh_tools.pyx:1040-1060)
  const double* hole_radius; const double* hole_colatitude; const double* hole_azimuth;   // [B*M]
  const int* partner; const int* is_cede;         // [M]: partner member index or -1; 1 = ceding member
  const int* member_cells;                        // nullptr, or [M][3]: (num_cells, min_sqrt, max_sqrt) of the hot
                                                  // region each member belongs to (regions may differ)
  const double* extra_params;                     // nullptr, or [B*M][n_params-2]: local variables after (log T, log g)
  const double* else_temperature;                 // nullptr, or [B]: log10 T of Elsewhere (for corrParams)
  double* corrParams;                             // nullptr, or [B*M][max_rings][n_params]: correction parameter rows
  // outputs: the integrator's per-instance inputs (padded layout of AzinvArgs)
  int* n_rings; int* n_azi;
  double* cellArea; double* phi; double* theta; double* radial; double* r_s_over_r; double* srcParams;
  double* cos_gamma; double* maxAlpha;
  double* deflection; double* cos_alpha; double* lag; double* maxDeflection;
  int* status;                                    // [B]
};
cudaError_t launch_embed_spots(EmbedArgs a, cudaStream_t stream);

// f1: embed of the closed whole-surface mesh of Elsewhere (global_mesh.pyx + rays) into TinvArgs arrays
struct ClosedMeshArgs {
  int B, n, n_rays;                               // n = sqrt_num_cells (even)
  const double* R_eq; const double* r_s; const double* epsilon; const double* zeta;      // [B]
  const double* temperature;                      // [B] log10 T
  double* cellArea;                               // [B]
  double* theta; double* phi; double* srcParams;  // [B][n][n], [B][n][n], [B][n][n][2]
  double* radial; double* r_s_over_r; double* cos_gamma; double* maxAlpha; double* ring_gravity;   // [B][n]
  double* deflection; double* cos_alpha; double* maxDeflection;                          // [B][n][n_rays], [B][n]
  int* n_rings;                                   // [B] scratch
  int* status;                                    // [B]
};
cudaError_t launch_embed_closed(ClosedMeshArgs a, cudaStream_t stream);

// a9: tools/energy_integrator.pyx:27-114, one spline per (signal q, phase column)
struct EnergyIntegArgs {
  int Q, n_energies, n_phases, n_in;
  const double* signal;              // [Q][N_E][N_P]
  const double* raw_energies;        // nullptr, or [N_E] keV: signal is a raw integrator sum and is
                                     //   first divided by E keV (integrator pyx:610-612)
  const double* div_b; int q_per_b;  // nullptr, or [Q/q_per_b]: then divided by d_sq (Likelihood.py:361-364)
  const double* log10_energies;      // [N_E]
  const double* log10_edges;         // [n_in + 1]
  int interp;                        // phase interpolant (!), energy_integrator.pyx:54
  // out[(col_of_q[q]*N_P + p) * n_in + j]  (+= when accumulate)
  const int* col_of_q;               // nullptr => q
  int accumulate;
  const double* attenuation;         // nullptr or [n_in] (Interstellar.__call__)
  const double* att_power;           // nullptr, or [Q/q_per_b]: the factor is attenuation[j] ** att_power[b]
  const int2* span;                  // nullptr, or [n_in]: energy intervals holding the (clamped) ends of input
                                     //   interval j -- the grids are theta-independent, so the pipeline looks them
                                     //   up once on the host (energy_span_table) instead of per (phase, q) on the device
  double* out;
};
// span table of EnergyIntegArgs (host): same clamping and search contract as the kernel's own lookup
void energy_span_table(const double* log10_energies, int n_energies, const double* log10_edges, int n_in, int2* out);
cudaError_t launch_energy_integrator(EnergyIntegArgs a, cudaStream_t stream);

// a11: Instrument.__call__  C[col][chan][p] = sum_in R[chan][in] X[(col,p)][in]
struct FoldArgs {
  int n_cols;                        // number of (theta, component) signals
  int n_phases, n_in, n_chan;
  const double* matrix;              // [n_chan][ld_matrix], first used column = in0
  int ld_matrix, in0;
  const double* x;                   // [n_cols * n_phases][n_in]
  double* out;                       // [n_cols][n_chan][n_phases]
  const int* k_range;                // nullptr, or per fold_tile_rows()-channel tile: [k_begin, k_end)
                                     //   of input intervals (relative to in0) with a non-zero entry
};
cudaError_t launch_fold(FoldArgs a, cudaStream_t stream);
int fold_tile_rows();

// a12-a14: expected counts + background-marginalised likelihood
struct MarginalArgs {
  int B, n_comp, n_chan, n_phases, n_bins;
  const double* pulses;              // [B][n_comp][n_chan][n_phases]  count rate
  const double* comp_phases;         // [n_phases] cycles shared by the components, or with comp_n_phases
                                     //   [n_comp][n_phases]: one grid per component (n_phases = the longest,
                                     //   compute_expected_counts.pyx:66-197 takes component_phases[i] per component)
  const int* comp_n_phases;          // nullptr, or [n_comp] nodes of each component's grid (1 = time-invariant)
  const int* comp_allow_negative;    // nullptr (allow_negative applies to all), or [n_comp]
  const double* phase_shifts;        // [B][n_comp]
  const double* data_phases;         // [n_bins + 1]
  const double* counts;              // [n_chan][n_bins]
  const double* precomp;             // [n_chan]
  const double* support;             // [n_chan][2]
  const double* background;          // nullptr or [n_chan][n_bins]
  double exposure_time, epsilon, sigmas, llzero, slim;
  int allow_negative, interp;
  int given_background;              // 1: _poisson_likelihood_given_background.pyx:14-113 (background required,
                                     //    in count rate; expected = T (star + background); no marginalisation)
  double* chan_lnL;                  // [B][n_chan]
  int* chan_status;                  // [B][n_chan]
  double* expected;                  // nullptr or [B][n_chan][n_bins]
  double* mcl_bg;                    // nullptr or [B][n_chan]
  double* mcl_bg_support;            // nullptr or [B][n_chan]
  double* lnL;                       // [B]
  int* status;                       // [B]
};
cudaError_t launch_marginal(MarginalArgs a, cudaStream_t stream);
int marginal_max_bins();           // lane j of a warp owns data phase bins j, j+32, ...: up to 128

// surface_radiation_field.intensity (core.pyx:125-308): point-wise intensities from local variables
struct IntensityArgs {
  int n, n_vars;
  const double* energies; const double* mu; const double* vars;     // [n], [n], [n][n_vars]
  AtmTable atm; int atm_ext;         // 1 blackbody, 2 Num4D
  int region;                        // 0 hot (beaming applies), 1 elsewhere
  int beam_opt;                      // 0-3 (hot_wrapper.pyx:155-199)
  double* out;                       // [n] photons/s/keV/cm^2/sr
};
cudaError_t launch_intensity(IntensityArgs a, cudaStream_t stream);

// a10: Interstellar.__call__
cudaError_t launch_attenuate(const double* att, int n_rows, int n_cols, double* signal, cudaStream_t stream);

// a12: row-wise spline tools (phase_integrator / phase_interpolator / energy_interpolator)
struct RowSplineArgs {
  int n_rows, n_nodes, n_out;
  const double* x;                   // [n_nodes] node abscissae
  const double* y; long y_row_stride, y_node_stride;
  const double* q;                   // op 0: [n_out+1] bin edges; op 1,2: [n_out] points
  int op;                            // 0 integrate phase bins, 1 interpolate in phase, 2 interpolate in energy
  double shift, scale;
  int allow_negative, interp, periodic;
  double* out; long out_row_stride, out_col_stride;
};
cudaError_t launch_row_spline(RowSplineArgs a, cudaStream_t stream);

// peak fp64 FMA rate of the device, measured with a register-resident DFMA chain
cudaError_t measure_fp64_peak(double* tflops, cudaStream_t stream);

// a13: precomputation
cudaError_t launch_precomputation(const int* counts, int n_chan, int n_bins, double* out,
                                  cudaStream_t stream);

}  // namespace xb
