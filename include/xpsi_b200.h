/*
 * xpsi_b200.h -- C ABI of the B200-native X-PSI likelihood hot path.
 *
 * The reference (X-PSI v3.3.0) has no FFI: its seam is "a Python callable that
 * takes and returns numpy arrays" (SURVEY.md s8b).  Each entry point below is
 * what a ctypes/cffi binding of that seam calls; the reference interface it
 * replaces is cited as <reference-relative path>:<lines>.  The Python mirror
 * of those callables lives in xpsi_b200/ (same names, argument order and
 * return conventions as the reference), INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; all arrays C-contiguous float64
 *     unless stated (int = 32-bit).  Pointers are HOST memory unless the
 *     function name ends in _device.
 *   - return value: 0 success; 1 the reference's numerical ERROR (its
 *     "(1, None)" return); <0 API/CUDA failure (message via
 *     xpsi_b200_last_error()).  Nothing is retained past a call except through
 *     explicit *_create / *_destroy handles.
 *   - there is no CPU fallback: every call fails with XPSI_B200_ENODEVICE when
 *     no CUDA device is usable.
 */
#ifndef XPSI_B200_H
#define XPSI_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XPSI_B200_OK 0
#define XPSI_B200_ENUMERICAL 1      /* reference returns (1, None) / raises PulseError */
#define XPSI_B200_EUNSUPPORTED 3    /* configuration outside the kernels' coverage   */
#define XPSI_B200_ESLIM 11          /* eval_marginal_likelihood slim early exit       */
#define XPSI_B200_EQUADRATURE 12    /* marginal integral non-positive                 */
#define XPSI_B200_EINVAL (-1)
#define XPSI_B200_ECUDA (-2)
#define XPSI_B200_ENODEVICE (-3)

/* atmosphere extension ids, xpsi/surface_radiation_field/hot_wrapper.pyx:72-252 */
#define XPSI_B200_ATM_BB 1
#define XPSI_B200_ATM_NUM4D 2
/* global interpolant ids, xpsi/tools/core.pyx:21 */
#define XPSI_B200_INTERP_AKIMA 0
#define XPSI_B200_INTERP_STEFFEN 1

const char* xpsi_b200_last_error(void);
int xpsi_b200_device_count(void);
int xpsi_b200_set_device(int device);
/* counters for bench.py: kernels launched / bytes moved by this library so far */
void xpsi_b200_counters(long long* kernel_launches, long long* h2d_bytes, long long* d2h_bytes);
void* xpsi_b200_stream(void);   /* cudaStream_t all work is issued on */
/* measured fp64 FMA peak of the current device (register-resident DFMA chains): the
 * roofline denominator bench.py reports against */
int xpsi_b200_fp64_peak_tflops(double* tflops);

/* ---- preloaded atmosphere table ------------------------------------------
 * replaces init_preload / _preloaded, surface_radiation_field/preload.pyx:6-44
 * (tuple (logT, logg, mu, logE, buf) of xpsi/Photosphere.py:208-217).        */
typedef struct xpsi_b200_atmosphere xpsi_b200_atmosphere;
xpsi_b200_atmosphere* xpsi_b200_atmosphere_create(const double* logT, int nT, const double* logg,
                                                  int ng, const double* mu, int nmu,
                                                  const double* logE, int nE, const double* buf);
void xpsi_b200_atmosphere_destroy(xpsi_b200_atmosphere* atm);

/* ---- cellmesh.integrator_for_azimuthal_invariance.integrate -----------------
 * replaces xpsi/cellmesh/integrator_for_azimuthal_invariance.pyx:70-665
 * (call site xpsi/HotRegion.py:1169-1197).  flux_out is [n_energies][n_phases].
 * correction_srcCellParams (NULL or [n_rings][n_azi][n_params]) activates the elsewhere
 * correction with else_atm_ext / elsewhere_atmosphere.  beam_opt 0-3 (hot_wrapper.pyx:155-199,
 * parameters in srcCellParams[..., 2:7]) and the disc occultation for R_in < 1e6
 * (common_functions.pyx:110-138) are covered. */
int xpsi_b200_integrate_azimuthal_invariance(
    double R, double omega, double r_s, double inclination,
    int n_rings, int n_azi,
    const double* cellArea, const double* radialCoords_of_parallels, const double* r_s_over_r,
    const double* theta, const double* phi,
    const double* srcCellParams, int n_params, const int* CELL_RADIATES,
    const double* correction_srcCellParams,
    int numRays, const double* deflection, const double* cos_alpha, const double* lag,
    const double* maxDeflection, const double* cos_gammaArray,
    int n_energies, const double* energies, int n_leaves, const double* leaves,
    int n_phases, const double* phases,
    const xpsi_b200_atmosphere* hot_atmosphere, const xpsi_b200_atmosphere* elsewhere_atmosphere,
    int hot_atm_ext, int else_atm_ext, int beam_opt, int image_order_limit, double R_in,
    int phase_interpolant, double* flux_out);

/* ---- cellmesh.integrator.integrate (no azimuthal invariance) ------------------
 * replaces xpsi/cellmesh/integrator.pyx:48-667 (bound by HotRegion.symmetry = False,
 * xpsi/HotRegion.py:567-569, and Everywhere(time_invariant=False), xpsi/Everywhere.py:330-332).
 * Same arguments as above; the atmosphere is evaluated with each CELL's srcCellParams /
 * correction_srcCellParams.  R_in is accepted and ignored, as in the reference.  At most
 * 128 phases. */
int xpsi_b200_integrate_general(
    double R, double omega, double r_s, double inclination,
    int n_rings, int n_azi,
    const double* cellArea, const double* radialCoords_of_parallels, const double* r_s_over_r,
    const double* theta, const double* phi,
    const double* srcCellParams, int n_params, const int* CELL_RADIATES,
    const double* correction_srcCellParams,
    int numRays, const double* deflection, const double* cos_alpha, const double* lag,
    const double* maxDeflection, const double* cos_gammaArray,
    int n_energies, const double* energies, int n_leaves, const double* leaves,
    int n_phases, const double* phases,
    const xpsi_b200_atmosphere* hot_atmosphere, const xpsi_b200_atmosphere* elsewhere_atmosphere,
    int hot_atm_ext, int else_atm_ext, int beam_opt, int image_order_limit, double R_in,
    int phase_interpolant, double* flux_out);

/* ---- cellmesh.integrator_for_time_invariance.integrate ------------------------------
 * replaces xpsi/cellmesh/integrator_for_time_invariance.pyx:59-338 (call sites
 * xpsi/Elsewhere.py:418-438, xpsi/Everywhere.py:581-601).  theta/phi [n][n],
 * srcCellParams [n][n][n_params], rays [n][numRays]; flux_out [n_energies].            */
int xpsi_b200_integrate_time_invariance(
    double R, double omega, double r_s, double inclination, int sqrt_numPix, double cellArea,
    const double* radialCoords_of_parallels, const double* r_s_over_r, const double* theta,
    const double* phi, const double* srcCellParams, int n_params, int numRays,
    const double* deflection, const double* cos_alpha, const double* maxDeflection,
    const double* cos_gammaArray, int n_energies, const double* energies,
    const xpsi_b200_atmosphere* atmosphere, int atm_ext, int image_order_limit, double* flux_out);

/* ---- tools.energy_integrator -------------------------------------------------
 * replaces xpsi/tools/energy_integrator.pyx:27-114.  signal [n_energies][n_phases],
 * out [n_in][n_phases] (the reference's transposed return).                    */
int xpsi_b200_energy_integrator(const double* signal, int n_energies, int n_phases,
                                const double* log10_energies, const double* log10_edges, int n_in,
                                int phase_interpolant, double* out);

/* ---- tools.phase_integrator / phase_interpolator / energy_interpolator ----------------
 * replace xpsi/tools/phase_integrator.pyx:23-121 (out [n_rows][n_bins]),
 * phase_interpolator.pyx:25-98 (out [n_rows][n_new]) and energy_interpolator.pyx:27-125
 * (signal [n_energies][n_phases], out [n_new][n_phases]; the *energy* interpolant).     */
int xpsi_b200_phase_integrator(double exposure_time, const double* phases, int n_bins, const double* signal,
                               int n_rows, const double* signal_phases, int n_phases, double phase_shift,
                               int allow_negative, int phase_interpolant, double* out);
int xpsi_b200_phase_interpolator(const double* new_phases, int n_new, const double* phases, int n_phases,
                                 const double* signal, int n_rows, double phase_shift, int allow_negative,
                                 int phase_interpolant, double* out);
int xpsi_b200_energy_interpolator(const double* signal, int n_energies, int n_phases, const double* log10_energies,
                                  const double* new_log10_energies, int n_new, int energy_interpolant,
                                  double* out);

/* ---- surface_radiation_field.intensity -----------------------------------------
 * replaces xpsi/surface_radiation_field/core.pyx:125-308: photon specific intensity
 * [photons/s/keV/cm^2/sr] at n points (energy keV, mu, local variables row) evaluated
 * directly by the atmosphere extension: atm_ext 1 (hot_BB.pyx:54-98) or 2 (hot_Num4D.pyx:248-460);
 * region_extension 0 = hot (beam_opt 0-3 of hot_wrapper.pyx:155-199, needs 7 variables per
 * point when beam_opt != 0), 1 = elsewhere (beam_opt ignored, elsewhere_wrapper.pyx:50-68). */
int xpsi_b200_intensity(int n, const double* energies, const double* mu, const double* local_variables,
                        int n_vars, const xpsi_b200_atmosphere* atmosphere, int region_extension, int atm_ext,
                        int beam_opt, double* out);

/* ---- Interstellar.__call__ ------------------------------------------------------------
 * replaces the in-place row scaling of xpsi/Interstellar.py:27-58:
 * signal[i][:] *= attenuation[i], signal [n_rows][n_cols] modified in place.           */
int xpsi_b200_interstellar_attenuate(const double* attenuation, int n_rows, int n_cols, double* signal);

/* ---- Instrument.__call__ -------------------------------------------------------
 * replaces numpy.dot(matrix[o0:o1, i0:i1], signal), xpsi/Instrument.py:192-197.
 * matrix is the full [n_rows][n_cols] response; signal [i1-i0][n_phases];
 * out [o1-o0][n_phases].                                                         */
int xpsi_b200_instrument_fold(const double* matrix, int n_rows, int n_cols, int i0, int i1, int o0,
                              int o1, const double* signal, int n_phases, double* out);

/* ---- likelihoods.precomputation / eval_marginal_likelihood -----------------------
 * replace xpsi/likelihoods/default_background_marginalisation.pyx:38-68 and :450-734.
 * components: n_comp arrays, component c is [n_chan][n_phases[c]] on its own phase
 * grid component_phases[c][n_phases[c]] (cycles), as compute_expected_counts.pyx:66-197
 * takes them (n_phases[c] = 1: a time-invariant component); allow_negative: NULL (no
 * component may contribute negatively) or one flag per component.  background may be NULL.  Returns 0,
 * XPSI_B200_ESLIM or XPSI_B200_EQUADRATURE (the two cases where the reference
 * returns a random near-llzero value); *lnL is then unspecified.               */
int xpsi_b200_precomputation(const int* counts, int n_chan, int n_bins, double* out);
int xpsi_b200_eval_marginal_likelihood(
    double exposure_time, const double* phases, int n_bins, const double* counts, int n_chan,
    const double* const* components, int n_comp, const double* const* component_phases, const int* n_phases,
    const double* phase_shifts, const double* neg_sum_ln_data_factorial, const double* support,
    double epsilon, double sigmas, double llzero, const int* allow_negative, double slim,
    const double* background, int phase_interpolant,
    double* lnL, double* expected_counts, double* mcl_background,
    double* mcl_background_given_support);

/* ---- likelihoods.poisson_likelihood_given_background / expected counts ----------------
 * replaces xpsi/likelihoods/_poisson_likelihood_given_background.pyx:14-113 and
 * tools/compute_expected_counts.pyx:200-315 (also the expected-count half of
 * tools/synthesise.pyx): expected = T (star + background), background in count rate
 * [n_chan][n_bins].  counts / neg_sum_ln_data_factorial / lnL may be NULL when only the
 * expected counts are wanted.  Returns XPSI_B200_EQUADRATURE where the reference adds a
 * random near-llzero penalty (zero expectation in a bin with counts).                  */
int xpsi_b200_poisson_likelihood_given_background(
    double exposure_time, const double* phases, int n_bins, const double* counts, int n_chan,
    const double* const* components, int n_comp, const double* const* component_phases, const int* n_phases,
    const double* phase_shifts, const double* background, const double* neg_sum_ln_data_factorial,
    const int* allow_negative, int phase_interpolant, double* lnL, double* expected_counts);

/* ---- batched likelihood pipeline (additional API) -------------
 * One handle holds every theta-independent constant on the device; eval takes a
 * batch of B parameter vectors already reduced to integrator inputs (mesh +
 * rays per hot-region member, the outputs of Star.update / HotRegion.embed,
 * xpsi/HotRegion.py:1033-1070) and returns lnL[B], status[B].                  */
typedef struct xpsi_b200_pipeline xpsi_b200_pipeline;

typedef struct {
  int n_components;          /* hot regions (likelihood signal components)              */
  int n_members;             /* integrator calls per theta (>= n_components)            */
  const int* member_component; /* [n_members] component each member's flux is added to  */
  int max_rings, max_azi;    /* padded mesh dims                                        */
  int n_rays, n_params;
  int n_energies; const double* energies;        /* keV */
  int n_leaves; const double* leaves;            /* rad */
  int n_phases; const double* phases;            /* rad; cycles = phases / 2 pi.  n_phases == 1: the star is an
                                                  * Everywhere(time_invariant=True) surface (xpsi/Everywhere.py:
                                                  * 577-601, Photosphere.py:560-566) -- no hot regions (n_members =
                                                  * n_components = 1, mesh fields ignored); the closed-surface
                                                  * settings come with set_extras (the "elsewhere" fields) and its
                                                  * temperature with else_temperature                             */
  int hot_atm_ext; const xpsi_b200_atmosphere* hot_atmosphere;
  int image_order_limit;
  int phase_interpolant;
  /* instrument */
  int n_in; const double* energy_edges;          /* [n_in+1] keV                          */
  int n_chan; const double* response;            /* [n_chan][n_in]                        */
  /* data + likelihood settings */
  int n_bins; const double* data_phases;         /* [n_bins+1] cycles                     */
  const double* counts;                          /* [n_chan][n_bins]                      */
  const double* support;                         /* [n_chan][2]                           */
  double exposure_time, epsilon, sigmas, llzero, slim;
  int allow_negative;
} xpsi_b200_pipeline_config;

typedef struct {
  /* per theta [B] */
  const double* omega; const double* inclination; const double* d_sq;
  const double* phase_shifts;                    /* [B][n_components] cycles              */
  /* per member instance q = b*n_members + m, padded to (max_rings, max_azi) */
  const int* n_rings; const int* n_azi;          /* [B*M]                                 */
  const double* cellArea; const double* phi;     /* [B*M][max_rings][max_azi]             */
  const double* theta;                           /* [B*M][max_rings] ring colatitude      */
  const double* radial; const double* r_s_over_r;/* [B*M][max_rings]                      */
  const double* srcParams;                       /* [B*M][max_rings][n_params]            */
  const double* deflection; const double* cos_alpha; const double* lag; /* [B*M][max_rings][n_rays] */
  const double* maxDeflection; const double* cos_gamma;                 /* [B*M][max_rings] */
} xpsi_b200_batch;

/* Parameter-level inputs for hot regions made of circular members (ST, ST-U, and with the optional
 * fields CST / EST / PST / CDT / EDT / PDT, polar caps included): the library then runs the embed
 * (mesh + rays, replacing HotRegion.embed, xpsi/HotRegion.py:1033-1070, mesh.pyx / polar_mesh.pyx /
 * mesh_tools.pyx / rays.pyx) on the GPU as well.                                                     */
typedef struct {
  /* per theta [B]: derived spacetime scalars, xpsi/Spacetime.py:110-188 */
  const double* R_eq; const double* r_s; const double* epsilon; const double* zeta;
  const double* omega; const double* inclination; const double* d_sq;
  const double* phase_shifts;                    /* [B][n_components] cycles                     */
  /* per member instance [B*M] */
  const double* colatitude; const double* ang_radius; const double* temperature;
  const double* phi_shift;                       /* radians added to cell azimuths (pi: antiphased) */
  double mode_frequency;                         /* Hz                                            */
  int num_cells, min_sqrt_num_cells, max_sqrt_num_cells;
  /* optional (NULL: plain circular spots).  The region masking each member: the omission hole of a
   * superseding member (omit_radius, omit_colatitude, omit_azimuth) or the superseding region inside a
   * ceding member (super_radius, super_colatitude, -cede_azimuth), as HotRegion passes them to the mesh
   * routines (xpsi/HotRegion.py:819-865); radius 0 = none.  Members covering a pole are meshed by the
   * polar variant (cellmesh/polar_mesh.pyx) automatically.                                          */
  const double* hole_radius; const double* hole_colatitude; const double* hole_azimuth;   /* [B*M] */
  /* optional (NULL: every member allocates num_cells alone): superseding / ceding members of one hot
   * region share num_cells (mesh_tools.pyx:1040-1060): partner[m] = index of the other member or -1,
   * is_cede[m] = 1 for the ceding member                                                            */
  const int* partner; const int* is_cede;         /* [M] */
  /* optional (NULL: zeros) when the pipeline was created with n_params > 2: the uniform local variables
   * that follow (log T, log g) in srcCellParams, e.g. the beaming parameters (abb, bbb, cbb, dbb, nimu) of
   * examples_modeling_tutorial/modules/CustomHotRegion_Beaming.py:149-178                            */
  const double* extra_params;                     /* [B*M][n_params-2] */
  /* optional (NULL: num_cells / min / max above apply to every member): the cell budget of the hot region each
   * member belongs to, [M][3] = (sqrt_num_cells^2, min_sqrt_num_cells, max_sqrt_num_cells) -- hot regions of one
   * model may be constructed with different resolutions (xpsi/HotRegion.py:122-140)                       */
  const int* member_cells;
} xpsi_b200_spot_batch;

/* ---- optional model components of the batched pipeline ---------------------------------------------
 * Set once after create.  elsewhere != 0 adds xpsi.Elsewhere (xpsi/Elsewhere.py): its time-invariant
 * spectrum (integrator_for_time_invariance.pyx) is added to every phase column of member 0's signal
 * (xpsi/Photosphere.py:589-592) and the hot members are integrated with the elsewhere correction
 * (integrator_for_azimuthal_invariance.pyx:257-268,469-478).  attenuation != NULL applies
 * xpsi.Interstellar.__call__ (xpsi/Interstellar.py:27-58) after the energy integration with the factor
 * attenuation[j] ** att_power[b] (att_power from the batch extras; 1 when absent).  beam_opt is passed
 * to the hot-region integrator (needs n_params >= 7 when non-zero).                                   */
typedef struct {
  int elsewhere;
  int else_sqrt_num_cells, else_num_rays, else_atm_ext, else_image_order_limit;
  const xpsi_b200_atmosphere* elsewhere_atmosphere;
  const double* attenuation;                     /* NULL or [n_in], copied                           */
  int beam_opt;
} xpsi_b200_pipeline_extras;
int xpsi_b200_pipeline_set_extras(xpsi_b200_pipeline* p, const xpsi_b200_pipeline_extras* extras);
/* on != 0: the ring / energy-chunk contributions to a member's signal are combined by a two-stage reduction in
 * fixed order (no floating-point atomics): lnL is then bitwise reproducible from run to run and independent of
 * the position of a parameter vector in its batch (samplers that resume or replay compare lnL bitwise).  Costs
 * one extra pass over [B*M][max_rings][n_energies][n_phases] doubles of workspace.                          */
int xpsi_b200_pipeline_set_deterministic(xpsi_b200_pipeline* p, int on);

/* Per-batch inputs of the optional components, uploaded for the next evaluation of B parameter vectors.
 * Parameter-level path (eval_spots): else_temperature [B] and att_power [B] are enough -- the closed
 * Elsewhere mesh and its rays are embedded on the device (cellmesh/global_mesh.pyx, rays.pyx).
 * Mesh-level path (eval): the caller passes the arrays Elsewhere.embed produced, in the layout of
 * xpsi_b200_integrate_time_invariance with a leading [B] axis, plus the correction parameter rows
 * of the hot members [B*M][max_rings][n_params].  Unused pointers may be NULL.                        */
typedef struct {
  const double* att_power;                       /* [B]                                              */
  const double* else_temperature;                /* [B] log10 K (parameter-level path)               */
  const double* else_cellArea;                   /* [B]                                              */
  const double* else_radial; const double* else_r_s_over_r;          /* [B][n]                       */
  const double* else_theta; const double* else_phi;                  /* [B][n][n]                    */
  const double* else_srcParams;                  /* [B][n][n][2]                                     */
  const double* else_deflection; const double* else_cos_alpha;       /* [B][n][else_num_rays]        */
  const double* else_maxDeflection; const double* else_cos_gamma;    /* [B][n]                       */
  const double* correction_srcParams;            /* [B*M][max_rings][n_params]                       */
} xpsi_b200_batch_extras;
int xpsi_b200_pipeline_upload_extras(xpsi_b200_pipeline* p, int B, const xpsi_b200_batch_extras* host);
/* Elsewhere spectrum [B][n_energies] of the last evaluation (photons/cm^2/s/keV before 1/d^2) */
int xpsi_b200_pipeline_fetch_elsewhere(xpsi_b200_pipeline* p, int B, double* spectrum);

/* ---- several signals per likelihood (xpsi/Likelihood.py:346-420,494-500; docs/source/Instrument_synergy.ipynb) ----
 * The reference integrates the photosphere once at the energies the signals share (Likelihood.py:102-107,316-318)
 * and registers that signal with every Signal object: energy integration onto the instrument's input intervals,
 * the signal's interstellar attenuation (Signal.py:431-439), the response fold (Instrument.py:146-197) and the
 * signal's own likelihood call; the joint log-likelihood is the sum.  add_signal registers one more
 * (instrument, data) pair behind the same integrator stage and returns its index (the signal of the pipeline
 * configuration is 0), or a negative error code.  A parameter vector whose evaluation fails in any signal keeps
 * the first non-zero status.  Arrays are copied.                                                              */
typedef struct {
  int n_in; const double* energy_edges;          /* [n_in+1] keV                                     */
  int n_chan; const double* response;            /* [n_chan][n_in]                                   */
  int n_bins; const double* data_phases;         /* [n_bins+1] cycles                                */
  const double* counts;                          /* [n_chan][n_bins]                                 */
  const double* support;                         /* [n_chan][2]                                      */
  double exposure_time, epsilon, sigmas, llzero, slim;
  int allow_negative;
  const double* attenuation;                     /* NULL or [n_in]: attenuation[j] ** att_power[b]   */
} xpsi_b200_signal_config;
int xpsi_b200_pipeline_add_signal(xpsi_b200_pipeline* p, const xpsi_b200_signal_config* signal);
int xpsi_b200_pipeline_n_signals(xpsi_b200_pipeline* p);
/* per-signal phase shifts [B][n_signals] in cycles, added to the hot regions' shifts for that signal
 * (Signal.shifts, xpsi/Signal.py:581-583); NULL switches them off again */
int xpsi_b200_pipeline_upload_signal_shifts(xpsi_b200_pipeline* p, int B, const double* shifts);
int xpsi_b200_pipeline_sweep_upload_signal_shifts(xpsi_b200_pipeline* p, long long N, const double* shifts);
/* folded signal [B][n_components][n_chan][n_phases], expected counts [B][n_chan][n_bins] and log-likelihood [B]
 * of one signal of the last evaluation (any pointer may be NULL) */
int xpsi_b200_pipeline_fetch_signal(xpsi_b200_pipeline* p, int signal, int B, double* folded, double* expected,
                                    double* lnL);

/* embed B parameter vectors on the device (inputs of the next eval_resident) */
int xpsi_b200_pipeline_embed_spots(xpsi_b200_pipeline* p, int B, const xpsi_b200_spot_batch* host);
/* re-run embed + the four stages on the spot batch already on the device (kernels only) */
int xpsi_b200_pipeline_eval_spots_resident(xpsi_b200_pipeline* p, int B);
/* embed + evaluate + download: theta-level end-to-end call */
int xpsi_b200_pipeline_eval_spots(xpsi_b200_pipeline* p, int B, const xpsi_b200_spot_batch* host,
                                  double* lnL, int* status);
/* Sweep of N parameter vectors (the reference's scatter / evaluate / gather loop over importance samples or
 * live points, xpsi/Sample.py:287-336, for one rank's share): the parameter-level inputs are uploaded ONCE
 * (att_power[N] / else_temperature[N] as in the batch extras, NULL when unused), sweep_run evaluates rows
 * [first, first+count) in blocks of <= max_batch without synchronising (embed + four stages per block, results
 * kept on the device), sweep_download copies lnL / status of a row range back.  sweep_results hands out the
 * device arrays lnL[N] / status[N] so that a collective can read them in place.                            */
int xpsi_b200_pipeline_sweep_upload(xpsi_b200_pipeline* p, long long N, const xpsi_b200_spot_batch* host,
                                    const double* att_power, const double* else_temperature);
int xpsi_b200_pipeline_sweep_run(xpsi_b200_pipeline* p, long long first, long long count);
int xpsi_b200_pipeline_sweep_download(xpsi_b200_pipeline* p, long long first, long long count, double* lnL,
                                      int* status);
int xpsi_b200_pipeline_sweep_results(xpsi_b200_pipeline* p, double** lnL, int** status);
/* fetch the embedded integrator inputs of the last embed (any pointer may be NULL) */
int xpsi_b200_pipeline_fetch_embed(xpsi_b200_pipeline* p, int B, int* n_rings, double* cellArea, double* phi,
                                   double* theta, double* radial, double* srcParams, double* cos_gamma,
                                   double* deflection, double* cos_alpha, double* lag, double* maxDeflection);

xpsi_b200_pipeline* xpsi_b200_pipeline_create(const xpsi_b200_pipeline_config* cfg, int max_batch);
void xpsi_b200_pipeline_destroy(xpsi_b200_pipeline* p);
/* host buffers: copies in, runs, copies lnL/status out (the e2e path) */
int xpsi_b200_pipeline_eval(xpsi_b200_pipeline* p, int B, const xpsi_b200_batch* host_batch,
                            double* lnL, int* status);
/* stage a host batch on the device once; eval_resident then times kernels only */
int xpsi_b200_pipeline_upload(xpsi_b200_pipeline* p, int B, const xpsi_b200_batch* host_batch);
int xpsi_b200_pipeline_eval_resident(xpsi_b200_pipeline* p, int B);
int xpsi_b200_pipeline_download(xpsi_b200_pipeline* p, int B, double* lnL, int* status);
/* optional: fetch intermediate device results of the last eval (NULL to skip) */
int xpsi_b200_pipeline_fetch(xpsi_b200_pipeline* p, int B, double* flux /*[B*M][E][P] raw*/,
                             double* folded /*[B][C][chan][P]*/, double* expected /*[B][chan][bins]*/);
/* algorithmic-work counters of the integrator: enable!=0 resets the counters and makes
 * later evals count (accumulating over evals); out (if non-NULL and counting was on) receives H, V, RI, K
 * summed over the evals since then */
int xpsi_b200_pipeline_work_counters(xpsi_b200_pipeline* p, int enable, unsigned long long out[4]);
/* per-stage device time of the last evaluation in ms: integrate, energy, fold, marginal, embed, and the
 * flux kernel alone (part of integrate) */
int xpsi_b200_pipeline_stage_ms(xpsi_b200_pipeline* p, float ms[6]);

#ifdef __cplusplus
}
#endif
#endif
