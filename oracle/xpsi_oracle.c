/*
 * xpsi_oracle.c -- CPU restatement of the X-PSI likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call this file; the
 * product (xpsi_b200/) must never route through it.
 *
 * Plain sequential C that follows the reference's control flow function by
 * function (citations are reference-relative file:line), on top of the
 * from-scratch GSL-subset shim in oracle/gslshim (the reference gets these
 * primitives from GNU GSL, an un-vendored, unpinned dependency).
 *
 * Pinning: tests/test_oracle.py checks every function here against the golden
 * vectors in tests/golden/ (recorded from the reference's own unmodified
 * sources built by oracle/build_ref.py), which in turn reproduce the
 * reference's published known answer lnL = -47881.27817666349
 * (xpsi/tests/test_likelihood.py:134) to 2.3e-12 relative.
 *
 * One deliberate difference: hot_Num4D keeps a per-thread stencil cache whose
 * base node depends on evaluation history (hot_Num4D.pyx:301-351); this
 * restatement uses the stateless rule b = clamp(j-1, 0, N-4) for
 * p[j] <= x <= p[j+1], identical for in-table queries.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gsl/gslshim.h"

#define ORACLE_OK 0
#define ORACLE_ERROR 1

/* xpsi/global_imports.py:70-80 */
static const double C_LIGHT = 2.99792458e8;
static const double KEV = 1.60217662e-16;
static const double K_B = 1.38064852e-23;
static const double H_KEV = 4.135667662e-18;

static int are_equal(double x, double y) { return fabs(x - y) < 1.0e-12; }   /* tools/core.pyx:118-122 */

static double eval_image_deflection(int order, double psi) {                  /* rays.pyx:53-57 */
  if (order % 2 != 0) return (double)(order + 1) * M_PI + pow(-1.0, (double)order) * psi;
  return (double)order * M_PI + pow(-1.0, (double)order) * psi;
}

static const gsl_interp_type *phase_interpolant(int kind) {                   /* tools/core.pyx:84-99 */
  if (kind == 1) return gsl_interp_steffen;
  if (kind == 2) return gsl_interp_cspline_periodic;
  return gsl_interp_akima_periodic;
}

/* ---------------------------------------------------------------- atmospheres */
typedef struct {
  const double *axis[4];     /* logT, logg, mu, logE */
  int n[4];
  const double *buf;         /* C-order [T][g][mu][E] */
} atm_table;

static double eval_hot_BB(double E, double logT) {                            /* hot_BB.pyx:54-87 */
  double temp = (K_B / KEV) * pow(10.0, logT);
  return E * E * E / (exp(E / temp) - 1.0);
}

static int base_node(const double *p, int n, double x) {
  int j = (int)gsl_interp_bsearch(p, x, 0, (size_t)n - 1);
  int b = j - 1;
  if (b < 0) b = 0;
  if (b > n - 4) b = n - 4;
  return b;
}

static double eval_hot_Num4D(const atm_table *t, double E, double mu, const double *VEC) {
  /* hot_Num4D.pyx:248-439 */
  double vec[4], w[4][4];
  int b[4];
  double E_eff = (K_B / KEV) * pow(10.0, VEC[0]);
  vec[0] = VEC[0]; vec[1] = VEC[1]; vec[2] = mu; vec[3] = log10(E / E_eff);
  for (int i = 0; i < 4; i++) {
    const double *p = t->axis[i];
    b[i] = base_node(p, t->n[i], vec[i]);
    const double *q = p + b[i];
    for (int a = 0; a < 4; a++) {
      double diff = 1.0, space = 1.0;
      for (int c = 0; c < 4; c++) if (c != a) { diff *= vec[i] - q[c]; space /= (q[a] - q[c]); }
      w[i][a] = diff * space;
    }
  }
  size_t S0 = (size_t)t->n[1] * t->n[2] * t->n[3], S1 = (size_t)t->n[2] * t->n[3], S2 = (size_t)t->n[3];
  double I = 0.0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 4; k++)
        for (int l = 0; l < 4; l++)
          I += w[0][i] * w[1][j] * w[2][k] * w[3][l] *
               t->buf[(b[0] + i) * S0 + (b[1] + j) * S1 + (b[2] + k) * S2 + (b[3] + l)];
  if (I < 0.0) return 0.0;
  return I * pow(10.0, 3.0 * vec[0]);
}

static double eval_hot(int atm_ext, const atm_table *t, double E, double mu, const double *VEC) {
  return atm_ext == 2 ? eval_hot_Num4D(t, E, mu, VEC) : eval_hot_BB(E, VEC[0]);  /* hot_wrapper.pyx:130-200 */
}
static double eval_hot_norm(int atm_ext) {
  return atm_ext == 2 ? 1.0e-7 / H_KEV : 1.0e-7 * 5.040366110812353e22;          /* hot_Num4D.pyx:460, hot_BB.pyx:98 */
}

/* hot_wrapper.pyx:110-199: intensity with the beaming modifications (options 1-3) */
static double eval_hot_I(int beam_opt, int atm_ext, const atm_table *t, double E, double mu, const double *VEC) {
  /* Option 3 sweeps mu up to 1 before every call, which leaves the Num4D stencil state at the top of the mu
   * axis; the next query below the table then walks down to the first node and is clamped to it
   * (hot_Num4D.pyx:301-323) -- with the sweep in between this happens on every call, not only the first. */
  double mu_q = (beam_opt == 3 && atm_ext == 2 && mu < t->axis[2][0]) ? t->axis[2][0] : mu;
  double I_hot = eval_hot(atm_ext, t, E, mu_q, VEC);
  if (beam_opt == 0) return I_hot;
  double abb = VEC[2], bbb = VEC[3], cbb = VEC[4], dbb = VEC[5], nimu = VEC[6];
  double beam = 0.0;
  if (beam_opt == 1) beam = (1.0 + abb * pow(E, cbb) * mu + bbb * pow(E, dbb) * mu * mu) * I_hot;
  if (beam_opt == 2) {
    double anorm = 0.5 / (0.5 + (1.0 / 3.0) * abb * pow(E, cbb) + (1.0 / 4.0) * bbb * pow(E, dbb));
    beam = anorm * (1.0 + abb * pow(E, cbb) * mu + bbb * pow(E, dbb) * mu * mu) * I_hot;
  }
  if (beam_opt == 3) {                                  /* trapezoid over mu, :173-192 */
    double mu_imu = 0.0, I_nom = 0.0, I_denom = 0.0;
    for (size_t imu = 0; imu < (size_t)nimu; imu++) {
      mu_imu = mu_imu + (1.0 / nimu);
      double dmu = (imu == 0 || imu == nimu - 1) ? (0.5 / nimu) : (1.0 / nimu);
      double Ii = eval_hot(atm_ext, t, E, mu_imu, VEC);
      double f = 1.0 + abb * pow(E, cbb) * mu_imu + bbb * pow(E, dbb) * mu_imu * mu_imu;
      I_denom = I_denom + mu_imu * f * Ii * dmu;
      I_nom = I_nom + mu_imu * Ii * dmu;
    }
    beam = are_equal(I_denom, 0.0) ? 0.0
                                   : (I_nom / I_denom) * (1.0 + abb * pow(E, cbb) * mu + bbb * pow(E, dbb) * mu * mu) * I_hot;
  }
  return beam < 0.0 ? 0.0 : beam;
}

/* cellmesh/common_functions.pyx:110-138 */
static int disk_block(double R_in, double cos_i, double cos_psi, double cos_theta_i, double r_s_over_r_i,
                      double radius, double sin_alpha, double theta_i_over_pi) {
  double cos_psi_d = (cos_i * cos_psi - cos_theta_i) /
                     sqrt(cos_i * cos_i + cos_theta_i * cos_theta_i - 2 * cos_i * cos_theta_i * cos_psi);
  double sin_psi_d = sqrt(1 - cos_psi_d * cos_psi_d);
  double r_s_i = r_s_over_r_i * radius;
  double impact_b = radius * sin_alpha / sqrt(1 - r_s_over_r_i);
  double r_psi_d = sqrt((r_s_i * r_s_i * (1 - cos_psi_d) * (1 - cos_psi_d)) / (4 * (1 + cos_psi_d) * (1 + cos_psi_d)) +
                        ((impact_b * impact_b) / (sin_psi_d * sin_psi_d))) -
                   (r_s_i * (1 - cos_psi_d)) / (2 * (1 + cos_psi_d));
  return (theta_i_over_pi < 0.5 || (theta_i_over_pi > 0.5 && r_psi_d < R_in)) ? 1 : 0;
}

/* ------------------------------------- integrator_for_azimuthal_invariance.pyx:70-665 */
int oracle_integrate_azinv(
    double omega, double inclination, int n_rings, int n_azi, const double *cellArea,
    const double *radial, const double *r_s_over_r, const double *theta, const double *phi,
    const double *srcCellParams, int n_params, const int *CELL_RADIATES, int N_R,
    const double *deflection, const double *cos_alpha, const double *lag, const double *maxDeflection,
    const double *cos_gammaArray, int N_E, const double *energies, int N_L, const double *leaves,
    int N_P, const double *phases, int hot_atm_ext, const double *logT, int nT, const double *logg,
    int ng, const double *mu_ax, int nmu, const double *logE, int nE, const double *buf,
    int image_order_limit, int phase_interp, double *flux,
    /* elsewhere correction (pyx:257-268): NULL, or [n_rings][n_azi][n_params] + its atmosphere */
    const double *correction, int else_atm_ext, const double *c_logT, int c_nT, const double *c_logg,
    int c_ng, const double *c_mu, int c_nmu, const double *c_logE, int c_nE, const double *c_buf,
    /* disc inner radius (>= 1e6: none, pyx:390-396) and beaming option 0-2 (hot_wrapper.pyx:155-172) */
    double R_in, int beam_opt) {
  atm_table tab = {{logT, logg, mu_ax, logE}, {nT, ng, nmu, nE}, buf};
  atm_table ctab = {{c_logT, c_logg, c_mu, c_logE}, {c_nT, c_ng, c_nmu, c_nE}, c_buf};
  const int perform_correction = correction != NULL;
  const double sin_i = sin(inclination), cos_i = cos(inclination);
  const size_t leaf_lim = (N_L % 2 == 0) ? N_L / 2 : (N_L + 1) / 2;
  int terminate = 0;
  double *PHASE = malloc(sizeof(double) * N_L);
  double *PROFILE = malloc(sizeof(double) * (size_t)N_E * N_L);
  double *cos_deflection = malloc(sizeof(double) * N_R);
  double *cos_alpha_alt = malloc(sizeof(double) * N_R);
  gsl_interp_accel *acc_a = gsl_interp_accel_alloc(), *acc_alt = gsl_interp_accel_alloc(),
                   *acc_l = gsl_interp_accel_alloc(), *acc_p = gsl_interp_accel_alloc();
  gsl_interp *interp_alpha = gsl_interp_alloc(gsl_interp_steffen, N_R);
  gsl_interp *interp_lag = gsl_interp_alloc(gsl_interp_steffen, N_R);
  gsl_interp *interp_PROFILE = gsl_interp_alloc(phase_interpolant(phase_interp), N_L);
  memset(flux, 0, sizeof(double) * (size_t)N_E * N_P);

  for (int i = 0; i < n_rings && !terminate; i++) {
    int J = -1;
    for (int j = 0; j < n_azi; j++) if (CELL_RADIATES[i * n_azi + j] == 1) { J = j; break; }   /* :286-296 */
    if (J < 0) continue;
    const double *defl = deflection + (size_t)i * N_R, *calpha = cos_alpha + (size_t)i * N_R,
                 *lagr = lag + (size_t)i * N_R;
    for (int j = 0; j < N_R; j++) {                                             /* :213-226 */
      cos_deflection[j] = cos(defl[N_R - j - 1]);
      cos_alpha_alt[j] = calpha[N_R - j - 1];
    }
    int jh = 0;
    while (jh < N_R - 1 && defl[jh] <= M_PI / 2.0) jh++;                        /* :301-303 */
    const double *defl_alt_ptr = cos_deflection + (N_R - jh - 1), *alpha_alt_ptr = cos_alpha_alt + (N_R - jh - 1);
    gsl_interp *interp_alt = gsl_interp_alloc(gsl_interp_steffen, jh + 1);
    gsl_interp_init(interp_alt, defl_alt_ptr, alpha_alt_ptr, jh + 1);
    gsl_interp_accel_reset(acc_alt); gsl_interp_accel_reset(acc_a); gsl_interp_accel_reset(acc_l);
    gsl_interp_init(interp_alpha, defl, calpha, N_R);
    gsl_interp_init(interp_lag, defl, lagr, N_R);

    const double radius = radial[i];                                            /* :318-331 */
    const double Grav_z = sqrt(1.0 - r_s_over_r[i]);
    const double cos_gamma = cos_gammaArray[i];
    const double sin_gamma = sqrt(1.0 - cos_gamma * cos_gamma);
    const double cos_theta_i = cos(theta[i * n_azi]), sin_theta_i = sin(theta[i * n_azi]);
    const double theta_i_over_pi = theta[i * n_azi] / M_PI;
    const double beta = radius * omega * sin_theta_i / (C_LIGHT * Grav_z);
    const double Lorentz = sqrt(1.0 - beta * beta);
    const double *VEC = srcCellParams + ((size_t)i * n_azi + J) * n_params;
    double _cos_alpha = -1.0, deriv = -1.0;
    const int _IO = image_order_limit > 0 ? image_order_limit : (int)ceil(maxDeflection[i] / M_PI);

    for (int I = 0; I < _IO && !terminate; I++) {
      int InvisFlag = 2;
      size_t _InvisPhase = 0;
      for (size_t k = 0; k < leaf_lim; k++) {
        double cos_psi = cos_i * cos_theta_i + sin_i * sin_theta_i * cos(leaves[k]);
        double psi = eval_image_deflection(I, acos(cos_psi));
        double sin_psi = sin(psi), sin_alpha = 0.0, mu = 0.0;
        int calc = 0;
        if (!are_equal(psi, 0.0) && are_equal(sin_psi, 0.0)) {                  /* :346-359 */
          double _i = cos_i >= 0.0 ? inclination + inclination * 1.0e-6 : inclination - inclination * 1.0e-6;
          cos_psi = cos(_i) * cos_theta_i + sin(_i) * sin_theta_i * cos(leaves[k]);
          psi = eval_image_deflection(I, acos(cos_psi));
          sin_psi = sin(psi);
        }
        const int use_alt = (psi <= M_PI / 2.0 && cos_psi >= interp_alt->xmin);
        if (psi <= maxDeflection[i]) {
          if (psi < interp_alpha->xmin || psi > interp_alpha->xmax) { terminate = 1; break; }
          if (use_alt) _cos_alpha = gsl_interp_eval(interp_alt, defl_alt_ptr, alpha_alt_ptr, cos_psi, acc_alt);
          else _cos_alpha = gsl_interp_eval(interp_alpha, defl, calpha, psi, acc_a);
          sin_alpha = sqrt(1.0 - _cos_alpha * _cos_alpha);
          mu = _cos_alpha * cos_gamma;
          if (!are_equal(psi, 0.0)) {
            double cos_delta = (cos_i - cos_theta_i * cos_psi) / (sin_theta_i * sin_psi);
            if (theta_i_over_pi < 0.5) mu = mu + sin_alpha * sin_gamma * cos_delta;
            else mu = mu - sin_alpha * sin_gamma * cos_delta;
          }
          if (mu > 0.0)
            calc = (R_in < 1e6) ? disk_block(R_in, cos_i, cos_psi, cos_theta_i, r_s_over_r[i], radius, sin_alpha,
                                             theta_i_over_pi)
                                : 1;
        }
        if (calc) {
          if (use_alt) deriv = gsl_interp_eval_deriv(interp_alt, defl_alt_ptr, alpha_alt_ptr, cos_psi, acc_alt);
          else {
            deriv = gsl_interp_eval_deriv(interp_alpha, defl, calpha, psi, acc_a);
            deriv = exp(log(fabs(deriv)) - log(fabs(sin_psi)));
          }
          if (psi < interp_lag->xmin || psi > interp_lag->xmax) { terminate = 1; break; }
          double _phase_lag = gsl_interp_eval(interp_lag, defl, lagr, psi, acc_l);
          for (int ks = 0; ks < 2; ks++) {
            if ((0 < k && k < leaf_lim - 1) || (k == 0 && ks == 0) ||
                (k == leaf_lim - 1 && N_L % 2 == 1 && ks == 0) || (k == leaf_lim - 1 && N_L % 2 == 0)) {
              size_t _kdx = ks == 0 ? k : (size_t)N_L - 1 - k;
              double superlum, eta;
              if (!are_equal(psi, 0.0)) {
                double cos_xi = sin_alpha * sin_i * sin(leaves[_kdx]) / sin_psi;
                superlum = 1.0 + beta * cos_xi;
                eta = Lorentz / superlum;
              } else { superlum = 1.0; eta = Lorentz; }
              double _Z = eta * Grav_z, _ABB = mu * eta;
              double _GEOM = mu * fabs(deriv) * Grav_z * eta * eta * eta / superlum;
              PHASE[_kdx] = leaves[_kdx] + _phase_lag;
              for (int p = 0; p < N_E; p++) {
                double E_prime = energies[p] / _Z;
                double I_E = eval_hot_I(beam_opt, hot_atm_ext, &tab, E_prime, _ABB, VEC);
                double correction_I_E = 0.0;
                if (perform_correction)                                          /* :469-476 */
                  correction_I_E = eval_hot(else_atm_ext, &ctab, E_prime, _ABB,
                                            correction + ((size_t)i * n_azi + J) * n_params) * eval_hot_norm(else_atm_ext);
                PROFILE[(size_t)p * N_L + _kdx] = (I_E * eval_hot_norm(hot_atm_ext) - correction_I_E) * _GEOM;
              }
            }
          }
          if (k == 0) {                                                          /* :480-484 */
            PHASE[N_L - 1] = PHASE[0] + 2.0 * M_PI;
            for (int p = 0; p < N_E; p++) PROFILE[(size_t)p * N_L + N_L - 1] = PROFILE[(size_t)p * N_L];
          } else if (InvisFlag == 2) {                                           /* :487-511 */
            double step = leaves[k] / (double)k;
            for (size_t m = N_L - k; m < (size_t)N_L; m++) PHASE[m] = PHASE[m - 1] + step;
            for (int p = 0; p < N_E; p++) for (size_t m = N_L - k; m < (size_t)N_L; m++) PROFILE[(size_t)p * N_L + m] = 0.0;
            PHASE[0] = PHASE[N_L - 1] - 2.0 * M_PI;
            for (size_t m = 1; m < k; m++) PHASE[m] = PHASE[m - 1] + step;
            for (int p = 0; p < N_E; p++) for (size_t m = 0; m < k; m++) PROFILE[(size_t)p * N_L + m] = 0.0;
          } else if (InvisFlag == 1) {                                           /* :513-524 */
            double step = (PHASE[k] - PHASE[_InvisPhase - 1]) / (double)(k - _InvisPhase + 1);
            for (size_t m = _InvisPhase; m < k; m++) PHASE[m] = PHASE[m - 1] + step;
            step = (PHASE[N_L - _InvisPhase] - PHASE[N_L - 1 - k]) / (double)(k - _InvisPhase + 1);
            for (size_t m = N_L - k; m < N_L - _InvisPhase; m++) PHASE[m] = PHASE[m - 1] + step;
          }
          InvisFlag = 0;
        } else if (InvisFlag == 0) {                                             /* :529-549 */
          double step = (PHASE[N_L - k] - PHASE[k - 1]) / (double)(N_L - 2 * k + 1);
          for (size_t m = k; m < N_L - k; m++) PHASE[m] = PHASE[m - 1] + step;
          for (int p = 0; p < N_E; p++) for (size_t m = k; m < N_L - k; m++) PROFILE[(size_t)p * N_L + m] = 0.0;
          InvisFlag = 1; _InvisPhase = k;
        }
      }
      if (terminate) break;
      if (InvisFlag == 2) break;                                                 /* :553-554 */
      for (int m = 1; m < N_L; m++) if (PHASE[m] <= PHASE[m - 1]) { terminate = 1; break; }
      if (terminate) break;
      for (int p = 0; p < N_E && !terminate; p++) {                              /* :566-596 */
        double *profile_ptr = PROFILE + (size_t)p * N_L;
        gsl_interp_accel_reset(acc_p);
        gsl_interp_init(interp_PROFILE, PHASE, profile_ptr, N_L);
        for (int j = 0; j < n_azi && !terminate; j++) {
          if (CELL_RADIATES[i * n_azi + j] != 1) continue;
          double phi_shift = phi[i * n_azi + j];
          for (int k = 0; k < N_P; k++) {
            double x = phases[k] + phi_shift;
            if (x > PHASE[N_L - 1]) { while (x > PHASE[N_L - 1]) x -= 2.0 * M_PI; }
            else if (x < PHASE[0]) { while (x < PHASE[0]) x += 2.0 * M_PI; }
            if (x < interp_PROFILE->xmin || x > interp_PROFILE->xmax) { terminate = 1; break; }
            double f = gsl_interp_eval(interp_PROFILE, PHASE, profile_ptr, x, acc_p);
            if (f > 0.0 || perform_correction) flux[(size_t)p * N_P + k] += cellArea[i * n_azi + j] * f;
          }
        }
      }
    }
    gsl_interp_free(interp_alt);
  }
  for (int p = 0; p < N_E; p++) for (int k = 0; k < N_P; k++) flux[(size_t)p * N_P + k] /= (energies[p] * KEV);
  gsl_interp_free(interp_alpha); gsl_interp_free(interp_lag); gsl_interp_free(interp_PROFILE);
  gsl_interp_accel_free(acc_a); gsl_interp_accel_free(acc_alt); gsl_interp_accel_free(acc_l); gsl_interp_accel_free(acc_p);
  free(PHASE); free(PROFILE); free(cos_deflection); free(cos_alpha_alt);
  return terminate ? ORACLE_ERROR : ORACLE_OK;
}

/* ------------------------------------------------- tools/energy_integrator.pyx:27-114 */
int oracle_energy_integrator(const double *signal, int N_E, int N_P, const double *energies,
                             const double *edges, int n_in, int phase_interp, double *out /*[n_in][N_P]*/) {
  gsl_interp *it = gsl_interp_alloc(phase_interpolant(phase_interp), N_E);
  gsl_interp_accel *acc = gsl_interp_accel_alloc();
  double *cpy = malloc(sizeof(double) * N_E);
  const double max_energy = energies[N_E - 1];
  memset(out, 0, sizeof(double) * (size_t)n_in * N_P);
  for (int i = 0; i < N_P; i++) {
    for (int j = 0; j < N_E; j++) cpy[j] = pow(10.0, energies[j]) * signal[(size_t)j * N_P + i] * log(10.0);
    gsl_interp_accel_reset(acc);
    gsl_interp_init(it, energies, cpy, N_E);
    for (int j = 0; j < n_in; j++) {
      double upper = edges[j + 1] > max_energy ? max_energy : edges[j + 1];
      out[(size_t)j * N_P + i] = gsl_interp_eval_integ(it, energies, cpy, edges[j], upper, acc);
      if (edges[j + 1] > max_energy) break;
    }
  }
  gsl_interp_free(it); gsl_interp_accel_free(acc); free(cpy);
  return ORACLE_OK;
}

/* ------------------------------------------------- Instrument.__call__, Instrument.py:192-197 */
int oracle_fold(const double *matrix, int n_chan, int n_in, const double *signal, int N_P, double *out) {
  for (int c = 0; c < n_chan; c++)
    for (int p = 0; p < N_P; p++) {
      double s = 0.0;
      for (int k = 0; k < n_in; k++) s += matrix[(size_t)c * n_in + k] * signal[(size_t)k * N_P + p];
      out[(size_t)c * N_P + p] = s;
    }
  return ORACLE_OK;
}

/* ------------------------- default_background_marginalisation.pyx:38-68 (precomputation) */
int oracle_precomputation(const int *data, int n_chan, int n_bins, double *precomp) {
  for (int i = 0; i < n_chan; i++) {
    double s = 0.0;
    for (int j = 0; j < n_bins; j++) s += gsl_sf_lnfact((unsigned int)data[(size_t)i * n_bins + j]);
    precomp[i] = -1.0 * s;
  }
  return ORACLE_OK;
}

/* --------------------------------- compute_expected_counts.pyx:66-197 (single channel) */
static void expected_star_single_channel(const double *phases, int n_bins, int n_comp,
                                         const double *const *components, const double *comp_phases,
                                         int N_P, const double *shifts, gsl_interp **interp,
                                         gsl_interp_accel **acc, double *STAR, int channel,
                                         int allow_negative) {
  for (int p = 0; p < n_comp; p++) {
    const double *pulse_ptr = components[p] + (size_t)channel * N_P;
    gsl_interp_init(interp[p], comp_phases, pulse_ptr, N_P);
    for (int j = 0; j < n_bins; j++) {
      double pa = phases[j] + shifts[p], pb = phases[j + 1] + shifts[p], v;
      if (are_equal(pb - pa, 1.0)) { pa = 0.0; pb = 1.0; }
      else { pa -= floor(pa); pb -= floor(pb); }
      if (pa < pb) {
        v = gsl_interp_eval_integ(interp[p], comp_phases, pulse_ptr, pa, pb, acc[p]);
        if (v > 0.0 || allow_negative) STAR[j] += v;
      } else {
        v = gsl_interp_eval_integ(interp[p], comp_phases, pulse_ptr, pa, 1.0, acc[p]);
        if (v > 0.0 || allow_negative) STAR[j] += v;
        v = gsl_interp_eval_integ(interp[p], comp_phases, pulse_ptr, 0.0, pb, acc[p]);
        if (v > 0.0 || allow_negative) STAR[j] += v;
      }
    }
  }
  for (int j = 0; j < n_bins; j++) if (STAR[j] < 0.0) STAR[j] = 0.0;
}

/* ------------------ default_background_marginalisation.pyx:84-173 (integrand, delta) */
typedef struct {
  size_t n; double SCALE; const double *data; double *star; double T_exp; double std; double A;
} marg_args;

static double marginal_integrand(double B, void *params) {
  marg_args *a = (marg_args *)params;
  double x = 0.0;
  for (size_t j = 0; j < a->n; j++) {
    double c = a->SCALE * (a->star[j] + B);
    if (c > 0.0) x += a->data[j] * log(c) - c;
    else if (are_equal(c, 0.0) && are_equal(a->data[j], 0.0)) { }
    else return 0.0;
  }
  return exp(x - a->A);
}

static double delta(double B, marg_args *a) {
  double x = 0.0, y = 0.0;
  for (size_t j = 0; j < a->n; j++) {
    y += a->data[j] / (a->star[j] + B);
    x += a->data[j] / pow(a->star[j] + B, 2.0);
  }
  y = 2.0 * a->T_exp - 2.0 * y;
  x *= 2.0;
  a->std = sqrt(2.0 / x);
  return -1.0 * y / x;
}

/* -------------- default_background_marginalisation.pyx:176-447 and :450-734
 * returns 0 ok, 11 slim early exit, 12 non-positive marginal integral (the two paths on
 * which the reference returns a random near-llzero number). */
int oracle_eval_marginal_likelihood(
    double exposure_time, const double *phases, int n_bins, const double *counts, int n_chan,
    const double *const *components, int n_comp, const double *comp_phases, int N_P,
    const double *shifts, const double *precomp, const double *support, int workspace_intervals,
    double epsabs, double epsrel, double epsilon, double sigmas, double llzero, int allow_negative,
    double slim, const double *background, int phase_interp, double *lnL, double *STAR /*[n_chan][n_bins]*/,
    double *MCL_BACKGROUND, double *MCL_BACKGROUND_GIVEN_SUPPORT) {
  const double n = (double)n_bins, SCALE = exposure_time / n;
  double LOGLIKE = 0.0;
  int rc = 0;
  gsl_interp **interp = malloc(sizeof(gsl_interp *) * n_comp);
  gsl_interp_accel **acc = malloc(sizeof(gsl_interp_accel *) * n_comp);
  gsl_integration_cquad_workspace *w = gsl_integration_cquad_workspace_alloc(workspace_intervals);
  for (int p = 0; p < n_comp; p++) {
    interp[p] = gsl_interp_alloc(phase_interpolant(phase_interp), N_P);
    acc[p] = gsl_interp_accel_alloc();
  }
  memset(STAR, 0, sizeof(double) * (size_t)n_chan * n_bins);
  for (int i = 0; i < n_chan; i++) {
    double *star = STAR + (size_t)i * n_bins;
    const double *data = counts + (size_t)i * n_bins;
    const double *sup = support + 2 * (size_t)i;
    expected_star_single_channel(phases, n_bins, n_comp, components, comp_phases, N_P, shifts,
                                 interp, acc, star, i, allow_negative);
    double av_DATA = 0.0, av_STAR = 0.0;
    for (int j = 0; j < n_bins; j++) {
      if (background) star[j] += background[(size_t)i * n_bins + j];
      star[j] *= n;
      av_STAR += star[j];
      av_DATA += data[j];
    }
    if (slim >= 0.0) {                                                           /* :677-684 */
      double limit = av_STAR * SCALE - slim * sqrt(av_STAR * SCALE) - av_DATA;
      if (limit > 0.0) { rc = 11; break; }
    }
    av_STAR /= n; av_DATA /= exposure_time;
    marg_args a = {(size_t)n_bins, SCALE, data, star, exposure_time, 0.0, 0.0};
    double B, lower, upper;
    if (are_equal(av_DATA, 0.0) && are_equal(av_STAR, 0.0)) {                    /* :287-301 */
      lower = 0.0; if (lower < sup[0]) lower = sup[0];
      upper = 10.0 / exposure_time; if (upper > sup[1] && sup[1] > 0.0) upper = sup[1];
      B = 0.0;
      LOGLIKE += log((exp(-1.0 * lower * exposure_time) - exp(-1.0 * upper * exposure_time)) / exposure_time);
    } else {
      double B_min = 0.0, dB, std_est = 0.0, B_for_integrand, result, abserr;
      size_t nevals;
      int counter = 0;
      B = av_DATA - av_STAR;
      if (B <= B_min) {                                                          /* :309-332 */
        double min_counts = -1.0;
        for (int j = 0; j < n_bins; j++) if (are_equal(star[j], 0.0)) { min_counts = -2.0; break; }
        if (are_equal(min_counts, -2.0))
          for (int j = 0; j < n_bins; j++)
            if ((are_equal(min_counts, -2.0) && data[j] > 0.0) || (0.0 < data[j] && data[j] < min_counts))
              min_counts = data[j];
        if (!are_equal(min_counts, -1.0)) {
          if (are_equal(min_counts, -2.0)) { B = 0.0; B_min = 0.0; }
          else { B = 0.01 * min_counts / SCALE; B_min = 0.1 * B; }
        } else B = B_min;
      }
      dB = delta(B, &a);                                                         /* :336-352 */
      while (fabs(dB) > epsilon * a.std && counter < 2) {
        B += dB;
        if (B < B_min) { if (B_min > 0.0) counter += 1; else counter = 2; B = B_min; }
        dB = delta(B, &a);
      }
      for (int j = 0; j < n_bins; j++) std_est += data[j] / pow(star[j] + B, 2.0);
      std_est = std_est > 0.0 ? sqrt(1.0 / std_est) : 1e90;
      lower = B - sigmas * std_est; upper = B + sigmas * std_est;
      if (lower < B_min) lower = B_min;
      B_for_integrand = B;                                                       /* :366-393 */
      if (lower < sup[0]) {
        lower = sup[0];
        if (upper < sup[0] && sup[1] > 0.0) { upper = sup[1]; B_for_integrand = sup[0]; }
        else if (upper < sup[0]) { upper = sup[0] + sigmas * std_est; B_for_integrand = sup[0]; }
      }
      if (upper > sup[1] && sup[1] > 0.0) {
        upper = sup[1];
        if (lower > sup[1]) { lower = sup[0]; B_for_integrand = sup[1]; }
      }
      if (B_for_integrand < lower) B_for_integrand = lower;
      else if (B_for_integrand > upper) B_for_integrand = upper;
      a.A = 0.0;                                                                 /* :397-408 */
      for (int j = 0; j < n_bins; j++) {
        double c = SCALE * (star[j] + B_for_integrand);
        if (c > 0.0) a.A += data[j] * log(c) - c;
        else if (are_equal(c, 0.0) && are_equal(data[j], 0.0)) { }
        else a.A += llzero;
      }
      gsl_function f = {&marginal_integrand, &a};
      gsl_integration_cquad(&f, lower, upper, epsabs, epsrel, w, &result, &abserr, &nevals);
      if (result > 0.0) LOGLIKE += log(result) + a.A + precomp[i];
      else { rc = 12; break; }
    }
    MCL_BACKGROUND[i] = B * exposure_time;                                       /* :431-447 */
    if (B < sup[0]) B = sup[0]; else if (B > sup[1] && sup[1] > 0.0) B = sup[1];
    MCL_BACKGROUND_GIVEN_SUPPORT[i] = B * exposure_time;
    for (int j = 0; j < n_bins; j++) star[j] = SCALE * (star[j] + B);
  }
  for (int p = 0; p < n_comp; p++) { gsl_interp_free(interp[p]); gsl_interp_accel_free(acc[p]); }
  free(interp); free(acc);
  gsl_integration_cquad_workspace_free(w);
  *lnL = LOGLIKE;
  return rc;
}

/* ------------------------------------------------- tools/phase_integrator.pyx:23-121 */
int oracle_phase_integrator(double exposure_time, const double *phases, int n_bins, const double *signal,
                            int n_rows, const double *signal_phases, int N_P, double phase_shift,
                            int allow_negative, int phase_interp, double *out /*[n_rows][n_bins]*/) {
  gsl_interp *it = gsl_interp_alloc(phase_interpolant(phase_interp), N_P);
  gsl_interp_accel *acc = gsl_interp_accel_alloc();
  memset(out, 0, sizeof(double) * (size_t)n_rows * n_bins);
  for (int i = 0; i < n_rows; i++) {
    const double *sp = signal + (size_t)i * N_P;
    gsl_interp_accel_reset(acc);
    gsl_interp_init(it, signal_phases, sp, N_P);
    for (int j = 0; j < n_bins; j++) {
      double a = phases[j] + phase_shift, b = phases[j + 1] + phase_shift, v;
      double *o = out + (size_t)i * n_bins + j;
      if (b - a == 1.0) { a = 0.0; b = 1.0; } else { a -= floor(a); b -= floor(b); }
      if (a < b) {
        v = gsl_interp_eval_integ(it, signal_phases, sp, a, b, acc);
        if (v > 0.0 || allow_negative) *o = v;
      } else {
        v = gsl_interp_eval_integ(it, signal_phases, sp, a, 1.0, acc);
        if (v > 0.0 || allow_negative) *o = v;
        v = gsl_interp_eval_integ(it, signal_phases, sp, 0.0, b, acc);
        if (v > 0.0 || allow_negative) *o += v;
      }
      *o *= exposure_time;
    }
  }
  gsl_interp_free(it); gsl_interp_accel_free(acc);
  return ORACLE_OK;
}

/* ------------------------------------------------- tools/phase_interpolator.pyx:25-98 */
int oracle_phase_interpolator(const double *new_phases, int n_new, const double *phases, int N_P,
                              const double *signal, int n_rows, double phase_shift, int allow_negative,
                              int phase_interp, double *out /*[n_rows][n_new]*/) {
  gsl_interp *it = gsl_interp_alloc(phase_interpolant(phase_interp), N_P);
  gsl_interp_accel *acc = gsl_interp_accel_alloc();
  memset(out, 0, sizeof(double) * (size_t)n_rows * n_new);
  for (int i = 0; i < n_rows; i++) {
    const double *sp = signal + (size_t)i * N_P;
    gsl_interp_accel_reset(acc);
    gsl_interp_init(it, phases, sp, N_P);
    for (int j = 0; j < n_new; j++) {
      double PHASE = new_phases[j] + phase_shift;
      PHASE -= floor(PHASE);
      double v = gsl_interp_eval(it, phases, sp, PHASE, acc);
      if (v > 0.0 || allow_negative) out[(size_t)i * n_new + j] = v;
    }
  }
  gsl_interp_free(it); gsl_interp_accel_free(acc);
  return ORACLE_OK;
}

/* ------------------------------------------------- tools/energy_interpolator.pyx:27-125 */
int oracle_energy_interpolator(const double *signal, int N_E, int N_P, const double *energies,
                               const double *new_energies, int n_new, int energy_interp,
                               double *out /*[n_new][N_P]*/) {
  const gsl_interp_type *T = energy_interp == 0 ? gsl_interp_akima
                           : (energy_interp == 2 ? gsl_interp_cspline : gsl_interp_steffen);  /* core.pyx:101-116 */
  gsl_interp *it = gsl_interp_alloc(T, N_E);
  gsl_interp_accel *acc = gsl_interp_accel_alloc();
  double *cpy = malloc(sizeof(double) * N_E);
  const double max_energy = energies[N_E - 1];
  memset(out, 0, sizeof(double) * (size_t)n_new * N_P);
  for (int i = 0; i < N_P; i++) {
    int mode = 1;
    for (int j = 0; j < N_E; j++) if (signal[(size_t)j * N_P + i] <= 0.0) mode = 0;
    for (int j = 0; j < N_E; j++) cpy[j] = mode ? log10(signal[(size_t)j * N_P + i]) : signal[(size_t)j * N_P + i];
    gsl_interp_accel_reset(acc);
    gsl_interp_init(it, energies, cpy, N_E);
    for (int j = 0; j < n_new; j++) {
      if (new_energies[j] > max_energy) continue;
      double v = gsl_interp_eval(it, energies, cpy, new_energies[j], acc);
      out[(size_t)j * N_P + i] = mode ? pow(10.0, v) : v;
    }
  }
  gsl_interp_free(it); gsl_interp_accel_free(acc); free(cpy);
  return ORACLE_OK;
}

/* -------------------------- elsewhere_wrapper.pyx:23-81 + integrator_for_time_invariance.pyx:59-338 */
int oracle_integrate_tinv(
    double omega, double inclination, int n, double cellArea, const double *radial, const double *r_s_over_r,
    const double *theta, const double *phi, const double *srcCellParams, int n_params, int N_R,
    const double *deflection, const double *cos_alpha, const double *maxDeflection,
    const double *cos_gammaArray, int N_E, const double *energies, int atm_ext, const double *logT, int nT,
    const double *logg, int ng, const double *mu_ax, int nmu, const double *logE, int nE, const double *buf,
    int image_order_limit, double *flux) {
  atm_table tab = {{logT, logg, mu_ax, logE}, {nT, ng, nmu, nE}, buf};
  const double sin_i = sin(inclination), cos_i = cos(inclination);
  int terminate = 0;
  double *cos_deflection = malloc(sizeof(double) * N_R), *cos_alpha_alt = malloc(sizeof(double) * N_R);
  gsl_interp_accel *acc_a = gsl_interp_accel_alloc(), *acc_alt = gsl_interp_accel_alloc();
  gsl_interp *interp_alpha = gsl_interp_alloc(gsl_interp_steffen, N_R);
  memset(flux, 0, sizeof(double) * N_E);
  for (int i = 0; i < n && !terminate; i++) {
    const double *defl = deflection + (size_t)i * N_R, *calpha = cos_alpha + (size_t)i * N_R;
    for (int j = 0; j < N_R; j++) { cos_deflection[j] = cos(defl[N_R - j - 1]); cos_alpha_alt[j] = calpha[N_R - j - 1]; }
    int jh = 0;
    while (jh < N_R - 1 && defl[jh] <= M_PI / 2.0) jh++;
    const double *defl_alt_ptr = cos_deflection + (N_R - jh - 1), *alpha_alt_ptr = cos_alpha_alt + (N_R - jh - 1);
    gsl_interp *interp_alt = gsl_interp_alloc(gsl_interp_steffen, jh + 1);
    gsl_interp_init(interp_alt, defl_alt_ptr, alpha_alt_ptr, jh + 1);
    gsl_interp_accel_reset(acc_alt); gsl_interp_accel_reset(acc_a);
    gsl_interp_init(interp_alpha, defl, calpha, N_R);
    const double Grav_z = sqrt(1.0 - r_s_over_r[i]);
    const double cos_gamma = cos_gammaArray[i], sin_gamma = sqrt(1.0 - cos_gamma * cos_gamma);
    const double cos_theta_i = cos(theta[(size_t)i * n]), sin_theta_i = sin(theta[(size_t)i * n]);
    const double theta_i_over_pi = theta[(size_t)i * n] / M_PI;
    const double beta = radial[i] * omega * sin_theta_i / (C_LIGHT * Grav_z);
    const double Lorentz = sqrt(1.0 - beta * beta);
    for (int j = 0; j < n && !terminate; j++) {
      const double ph = phi[(size_t)i * n + j];
      const double _cos_psi = cos_i * cos_theta_i + sin_i * sin_theta_i * cos(ph);
      const double _psi = acos(_cos_psi);
      const int _IO = image_order_limit > 0 ? image_order_limit : (int)ceil(maxDeflection[i] / M_PI);
      for (int I = 0; I < _IO; I++) {
        double cos_psi = _cos_psi, psi = eval_image_deflection(I, _psi), sin_psi = sin(psi), _cos_alpha, deriv;
        if (!are_equal(psi, 0.0) && are_equal(sin_psi, 0.0)) {
          double _i = cos_i >= 0.0 ? inclination + inclination * 1.0e-6 : inclination - inclination * 1.0e-6;
          cos_psi = cos(_i) * cos_theta_i + sin(_i) * sin_theta_i * cos(ph);
          psi = eval_image_deflection(I, acos(cos_psi));
          sin_psi = sin(psi);
        }
        if (psi > maxDeflection[i]) break;
        if (psi < interp_alpha->xmin || psi > interp_alpha->xmax) { terminate = 1; break; }
        const int use_alt = (psi <= M_PI / 2.0 && cos_psi >= interp_alt->xmin);
        if (use_alt) _cos_alpha = gsl_interp_eval(interp_alt, defl_alt_ptr, alpha_alt_ptr, cos_psi, acc_alt);
        else _cos_alpha = gsl_interp_eval(interp_alpha, defl, calpha, psi, acc_a);
        const double sin_alpha = sqrt(1.0 - _cos_alpha * _cos_alpha);
        double mu = _cos_alpha * cos_gamma;
        if (!are_equal(psi, 0.0)) {
          double cos_delta = (cos_i - cos_theta_i * cos_psi) / (sin_theta_i * sin_psi);
          if (theta_i_over_pi < 0.5) mu = mu + sin_alpha * sin_gamma * cos_delta;
          else mu = mu - sin_alpha * sin_gamma * cos_delta;
        }
        if (mu > 0.0) {
          double eta;
          if (!are_equal(sin_psi, 0.0)) eta = Lorentz / (1.0 + beta * (sin_alpha * sin_i * sin(ph) / sin_psi));
          else eta = Lorentz;
          if (use_alt) deriv = gsl_interp_eval_deriv(interp_alt, defl_alt_ptr, alpha_alt_ptr, cos_psi, acc_alt);
          else {
            deriv = gsl_interp_eval_deriv(interp_alpha, defl, calpha, psi, acc_a);
            deriv = exp(log(fabs(deriv)) - log(fabs(sin_psi)));
          }
          const double _Z = eta * Grav_z, _ABB = mu * eta, _GEOM = mu * fabs(deriv) * Grav_z * eta * eta * eta;
          const double *VEC = srcCellParams + ((size_t)i * n + j) * n_params;
          for (int e = 0; e < N_E; e++) flux[e] += eval_hot(atm_ext, &tab, energies[e] / _Z, _ABB, VEC) * _GEOM;
        }
      }
    }
    gsl_interp_free(interp_alt);
  }
  for (int e = 0; e < N_E; e++) flux[e] *= cellArea * eval_hot_norm(atm_ext) / (energies[e] * KEV);
  gsl_interp_free(interp_alpha); gsl_interp_accel_free(acc_a); gsl_interp_accel_free(acc_alt);
  free(cos_deflection); free(cos_alpha_alt);
  return terminate ? ORACLE_ERROR : ORACLE_OK;
}

/* ------------------------------------------------------------- cellmesh/integrator.pyx:48-667
 * The integrator without azimuthal invariance: per (ring, image) the leaf quantities GEOM, Z, ABB
 * are splined against the lagged leaf phase (GEOM with the phase interpolant, Z and ABB with Steffen,
 * :537-542) and the atmosphere is evaluated per (cell, phase, energy) with the CELL's parameters
 * (:544-592).  Argument list as oracle_integrate_azinv minus R_in (the general integrator ignores it). */
int oracle_integrate_general(
    double omega, double inclination, int n_rings, int n_azi, const double *cellArea,
    const double *radial, const double *r_s_over_r, const double *theta, const double *phi,
    const double *srcCellParams, int n_params, const int *CELL_RADIATES, int N_R,
    const double *deflection, const double *cos_alpha, const double *lag, const double *maxDeflection,
    const double *cos_gammaArray, int N_E, const double *energies, int N_L, const double *leaves,
    int N_P, const double *phases, int hot_atm_ext, const double *logT, int nT, const double *logg,
    int ng, const double *mu_ax, int nmu, const double *logE, int nE, const double *buf,
    int image_order_limit, int phase_interp, double *flux,
    const double *correction, int else_atm_ext, const double *c_logT, int c_nT, const double *c_logg,
    int c_ng, const double *c_mu, int c_nmu, const double *c_logE, int c_nE, const double *c_buf,
    int beam_opt) {
  atm_table tab = {{logT, logg, mu_ax, logE}, {nT, ng, nmu, nE}, buf};
  atm_table ctab = {{c_logT, c_logg, c_mu, c_logE}, {c_nT, c_ng, c_nmu, c_nE}, c_buf};
  const int perform_correction = correction != NULL;
  const double sin_i = sin(inclination), cos_i = cos(inclination);
  const size_t NL = (size_t)N_L;
  const size_t leaf_lim = (N_L % 2 == 0) ? NL / 2 : (NL + 1) / 2;                 /* :214-217 */
  int terminate = 0;
  double *PHASE = malloc(sizeof(double) * NL), *GEOM = calloc(NL, sizeof(double)), *Z = calloc(NL, sizeof(double)),
         *ABB = calloc(NL, sizeof(double));
  double *cos_deflection = malloc(sizeof(double) * N_R), *cos_alpha_alt = malloc(sizeof(double) * N_R);
  gsl_interp_accel *acc_a = gsl_interp_accel_alloc(), *acc_alt = gsl_interp_accel_alloc(),
                   *acc_l = gsl_interp_accel_alloc(), *acc_G = gsl_interp_accel_alloc(),
                   *acc_Z = gsl_interp_accel_alloc(), *acc_A = gsl_interp_accel_alloc();
  gsl_interp *interp_alpha = gsl_interp_alloc(gsl_interp_steffen, N_R);
  gsl_interp *interp_lag = gsl_interp_alloc(gsl_interp_steffen, N_R);
  gsl_interp *interp_Z = gsl_interp_alloc(gsl_interp_steffen, NL);                /* :174-179 */
  gsl_interp *interp_ABB = gsl_interp_alloc(gsl_interp_steffen, NL);
  gsl_interp *interp_GEOM = gsl_interp_alloc(phase_interpolant(phase_interp), NL);
  memset(flux, 0, sizeof(double) * (size_t)N_E * N_P);

  for (int i = 0; i < n_rings && !terminate; i++) {
    int J = -1;
    for (int j = 0; j < n_azi; j++) if (CELL_RADIATES[i * n_azi + j] == 1) { J = j; break; }   /* :264-274 */
    if (J < 0) continue;
    const double *defl = deflection + (size_t)i * N_R, *calpha = cos_alpha + (size_t)i * N_R,
                 *lagr = lag + (size_t)i * N_R;
    for (int j = 0; j < N_R; j++) {                                               /* :192-208 */
      cos_deflection[j] = cos(defl[N_R - j - 1]);
      cos_alpha_alt[j] = calpha[N_R - j - 1];
    }
    int jh = 0;
    while (jh < N_R - 1 && defl[jh] <= M_PI / 2.0) jh++;                          /* :279-281 */
    const double *defl_alt_ptr = cos_deflection + (N_R - jh - 1), *alpha_alt_ptr = cos_alpha_alt + (N_R - jh - 1);
    gsl_interp *interp_alt = gsl_interp_alloc(gsl_interp_steffen, jh + 1);
    gsl_interp_init(interp_alt, defl_alt_ptr, alpha_alt_ptr, jh + 1);
    gsl_interp_accel_reset(acc_alt); gsl_interp_accel_reset(acc_a); gsl_interp_accel_reset(acc_l);
    gsl_interp_init(interp_alpha, defl, calpha, N_R);
    gsl_interp_init(interp_lag, defl, lagr, N_R);

    const double radius = radial[i];                                              /* :294-306 */
    const double Grav_z = sqrt(1.0 - r_s_over_r[i]);
    const double cos_gamma = cos_gammaArray[i];
    const double sin_gamma = sqrt(1.0 - cos_gamma * cos_gamma);
    const double cos_theta_i = cos(theta[i * n_azi]), sin_theta_i = sin(theta[i * n_azi]);
    const double theta_i_over_pi = theta[i * n_azi] / M_PI;
    const double beta = radius * omega * sin_theta_i / (C_LIGHT * Grav_z);
    const double Lorentz = sqrt(1.0 - beta * beta);
    double _cos_alpha = -1.0, deriv = -1.0;
    const int _IO = image_order_limit > 0 ? image_order_limit : (int)ceil(maxDeflection[i] / M_PI);

    for (int I = 0; I < _IO && !terminate; I++) {
      int InvisFlag = 2;
      size_t _InvisPhase = 0;
      for (size_t k = 0; k < NL; k++) {                                           /* :316, all leaves */
        double cos_psi = cos_i * cos_theta_i + sin_i * sin_theta_i * cos(leaves[k]);
        double psi = eval_image_deflection(I, acos(cos_psi));
        double sin_psi = sin(psi), sin_alpha = 0.0, mu = 0.0;
        int lit = 0;
        if (!are_equal(psi, 0.0) && are_equal(sin_psi, 0.0)) {                    /* :321-334 */
          double _i = cos_i >= 0.0 ? inclination + inclination * 1.0e-6 : inclination - inclination * 1.0e-6;
          cos_psi = cos(_i) * cos_theta_i + sin(_i) * sin_theta_i * cos(leaves[k]);
          psi = eval_image_deflection(I, acos(cos_psi));
          sin_psi = sin(psi);
        }
        const int use_alt = (psi <= M_PI / 2.0 && cos_psi >= interp_alt->xmin);
        if (psi <= maxDeflection[i]) {
          if (psi < interp_alpha->xmin || psi > interp_alpha->xmax) { terminate = 1; break; }
          if (use_alt) _cos_alpha = gsl_interp_eval(interp_alt, defl_alt_ptr, alpha_alt_ptr, cos_psi, acc_alt);
          else _cos_alpha = gsl_interp_eval(interp_alpha, defl, calpha, psi, acc_a);
          sin_alpha = sqrt(1.0 - _cos_alpha * _cos_alpha);
          mu = _cos_alpha * cos_gamma;
          if (!are_equal(psi, 0.0)) {
            double cos_delta = (cos_i - cos_theta_i * cos_psi) / (sin_theta_i * sin_psi);
            if (theta_i_over_pi < 0.5) mu = mu + sin_alpha * sin_gamma * cos_delta;
            else mu = mu - sin_alpha * sin_gamma * cos_delta;
          }
          lit = mu > 0.0;
        }
        if (lit) {
          if (use_alt) deriv = gsl_interp_eval_deriv(interp_alt, defl_alt_ptr, alpha_alt_ptr, cos_psi, acc_alt);
          else {
            deriv = gsl_interp_eval_deriv(interp_alpha, defl, calpha, psi, acc_a);
            deriv = exp(log(fabs(deriv)) - log(fabs(sin_psi)));
          }
          if (psi < interp_lag->xmin || psi > interp_lag->xmax) { terminate = 1; break; }
          double _phase_lag = gsl_interp_eval(interp_lag, defl, lagr, psi, acc_l);
          for (int ks = 0; ks < 2; ks++) {                                        /* :376-398 */
            if ((0 < k && k < leaf_lim - 1) || (k == 0 && ks == 0) ||
                (k == leaf_lim - 1 && N_L % 2 == 1 && ks == 0) || (k == leaf_lim - 1 && N_L % 2 == 0)) {
              size_t _kdx = ks == 0 ? k : NL - 1 - k;
              double superlum, eta;
              if (!are_equal(psi, 0.0)) {
                double cos_xi = sin_alpha * sin_i * sin(leaves[_kdx]) / sin_psi;
                superlum = 1.0 + beta * cos_xi;
                eta = Lorentz / superlum;
              } else { superlum = 1.0; eta = Lorentz; }
              Z[_kdx] = eta * Grav_z;
              ABB[_kdx] = mu * eta;
              GEOM[_kdx] = mu * fabs(deriv) * Grav_z * eta * eta * eta / superlum;
              PHASE[_kdx] = leaves[_kdx] + _phase_lag;
            }
          }
          if (k == 0) {                                                           /* :400-405 */
            PHASE[NL - 1] = PHASE[0] + 2.0 * M_PI;
            Z[NL - 1] = Z[0]; ABB[NL - 1] = ABB[0]; GEOM[NL - 1] = GEOM[0];
          } else if (InvisFlag == 2) {                                            /* :406-437 */
            double step = leaves[k] / (double)k;
            double Zs = (Z[k] - Z[NL - k - 1]) / (2.0 * (double)k);
            double As = (ABB[k] - ABB[NL - k - 1]) / (2.0 * (double)k);
            for (size_t m = NL - k; m < NL; m++) {
              PHASE[m] = PHASE[m - 1] + step; Z[m] = Z[m - 1] + Zs; ABB[m] = ABB[m - 1] + As; GEOM[m] = 0.0;
            }
            PHASE[0] = PHASE[NL - 1] - 2.0 * M_PI;
            Z[0] = Z[NL - 1]; ABB[0] = ABB[NL - 1]; GEOM[0] = GEOM[NL - 1];
            for (size_t m = 1; m < k; m++) {
              PHASE[m] = PHASE[m - 1] + step; Z[m] = Z[m - 1] + Zs; ABB[m] = ABB[m - 1] + As; GEOM[m] = 0.0;
            }
          } else if (InvisFlag == 1) {                                            /* :438-470 */
            double den = (double)(k - _InvisPhase + 1);
            double step = (PHASE[k] - PHASE[_InvisPhase - 1]) / den;
            double Zs = (Z[k] - Z[_InvisPhase - 1]) / den, As = (ABB[k] - ABB[_InvisPhase - 1]) / den;
            for (size_t m = _InvisPhase; m < k; m++) {
              PHASE[m] = PHASE[m - 1] + step; Z[m] = Z[m - 1] + Zs; ABB[m] = ABB[m - 1] + As;
            }
            step = (PHASE[NL - _InvisPhase] - PHASE[NL - 1 - k]) / den;
            Zs = (Z[NL - _InvisPhase] - Z[NL - 1 - k]) / den;
            As = (ABB[NL - _InvisPhase] - ABB[NL - 1 - k]) / den;
            for (size_t m = NL - k; m < NL - _InvisPhase; m++) {
              PHASE[m] = PHASE[m - 1] + step; Z[m] = Z[m - 1] + Zs; ABB[m] = ABB[m - 1] + As;
            }
          }
          InvisFlag = 0;
        } else if (InvisFlag == 0) {                                              /* :475-499 / :500-520 */
          /* size_t arithmetic as in the reference: for k past the half-way leaf the range is empty */
          double den = (double)(size_t)(NL - 2 * k + 1);
          double step = (PHASE[NL - k] - PHASE[k - 1]) / den;
          double Zs = (Z[NL - k] - Z[k - 1]) / den, As = (ABB[NL - k] - ABB[k - 1]) / den;
          for (size_t m = k; m < NL - k; m++) {
            PHASE[m] = PHASE[m - 1] + step; Z[m] = Z[m - 1] + Zs; ABB[m] = ABB[m - 1] + As; GEOM[m] = 0.0;
          }
          InvisFlag = 1; _InvisPhase = k;
        }
      }
      if (terminate) break;
      if (InvisFlag == 2) break;                                                  /* :524-525 */
      for (size_t m = 1; m < NL; m++) if (PHASE[m] <= PHASE[m - 1]) { terminate = 1; break; }
      if (terminate) break;
      gsl_interp_accel_reset(acc_Z); gsl_interp_init(interp_Z, PHASE, Z, NL);     /* :537-542 */
      gsl_interp_accel_reset(acc_A); gsl_interp_init(interp_ABB, PHASE, ABB, NL);
      gsl_interp_accel_reset(acc_G); gsl_interp_init(interp_GEOM, PHASE, GEOM, NL);
      for (int j = 0; j < n_azi && !terminate; j++) {                             /* :544-592 */
        if (CELL_RADIATES[i * n_azi + j] != 1) continue;
        const double phi_shift = phi[i * n_azi + j];
        const double *VEC = srcCellParams + ((size_t)i * n_azi + j) * n_params;
        for (int k = 0; k < N_P; k++) {
          double x = phases[k] + phi_shift;
          if (x > PHASE[NL - 1]) { while (x > PHASE[NL - 1]) x -= 2.0 * M_PI; }
          else if (x < PHASE[0]) { while (x < PHASE[0]) x += 2.0 * M_PI; }
          if (x < interp_GEOM->xmin || x > interp_GEOM->xmax) { terminate = 1; break; }
          double g = gsl_interp_eval(interp_GEOM, PHASE, GEOM, x, acc_G);
          if (g > 0.0) {
            double z = gsl_interp_eval(interp_Z, PHASE, Z, x, acc_Z);
            double abb = gsl_interp_eval(interp_ABB, PHASE, ABB, x, acc_A);
            for (int p = 0; p < N_E; p++) {
              double E_prime = energies[p] / z;
              double I_E = eval_hot_I(beam_opt, hot_atm_ext, &tab, E_prime, abb, VEC);
              I_E *= eval_hot_norm(hot_atm_ext);
              double c = 0.0;
              if (perform_correction)
                c = eval_hot(else_atm_ext, &ctab, E_prime, abb, correction + ((size_t)i * n_azi + j) * n_params) *
                    eval_hot_norm(else_atm_ext);
              flux[(size_t)p * N_P + k] += cellArea[i * n_azi + j] * (I_E - c) * g;
            }
          }
        }
      }
    }
    gsl_interp_free(interp_alt);
  }
  for (int p = 0; p < N_E; p++) for (int k = 0; k < N_P; k++) flux[(size_t)p * N_P + k] /= (energies[p] * KEV);
  gsl_interp_free(interp_alpha); gsl_interp_free(interp_lag); gsl_interp_free(interp_Z); gsl_interp_free(interp_ABB);
  gsl_interp_free(interp_GEOM);
  gsl_interp_accel_free(acc_a); gsl_interp_accel_free(acc_alt); gsl_interp_accel_free(acc_l);
  gsl_interp_accel_free(acc_G); gsl_interp_accel_free(acc_Z); gsl_interp_accel_free(acc_A);
  free(PHASE); free(GEOM); free(Z); free(ABB); free(cos_deflection); free(cos_alpha_alt);
  return terminate ? ORACLE_ERROR : ORACLE_OK;
}

/* -------------------------------------------- surface_radiation_field/core.pyx:125-308 (intensity)
 * region 0 = hot (hot_wrapper.pyx:110-199 incl. beaming options 1-3), 1 = elsewhere
 * (elsewhere_wrapper.pyx:50-68, beam_opt ignored). */
int oracle_intensity(int n, const double *energies, const double *mu, const double *vars, int n_vars,
                     int region, int atm_ext, const double *logT, int nT, const double *logg, int ng,
                     const double *mu_ax, int nmu, const double *logE, int nE, const double *buf, int beam_opt,
                     double *out) {
  atm_table tab = {{logT, logg, mu_ax, logE}, {nT, ng, nmu, nE}, buf};
  for (int i = 0; i < n; i++) {
    const double *VEC = vars + (size_t)i * n_vars;
    double E = energies[i], m = mu[i];
    double I = region == 0 ? eval_hot_I(beam_opt, atm_ext, &tab, E, m, VEC) : eval_hot(atm_ext, &tab, E, m, VEC);
    out[i] = I * eval_hot_norm(atm_ext) / (E * KEV);
  }
  return ORACLE_OK;
}
