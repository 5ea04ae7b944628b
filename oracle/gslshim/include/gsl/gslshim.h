/*
 * gslshim.h -- from-scratch declarations for the subset of the GNU Scientific
 * Library API that the X-PSI likelihood hot path calls (see
 * /root/reference/xpsi/include/GSL.pxd for the call surface).
 *
 * TEST INFRASTRUCTURE ONLY.  This is NOT GSL and contains no GSL code: every
 * routine in gslshim.c is re-derived from the published algorithm (Steffen
 * 1990; Akima 1970; natural/periodic cubic splines; QUADPACK QAG with
 * Gauss-Kronrod pairs; an adaptive quadrature standing in for CQUAD).  It
 * exists so that (a) the reference's own unmodified .pyx sources can be
 * compiled in a container that has no GSL (oracle/build_ref.py -> oracle/_ref)
 * and (b) the C restatement in oracle/xpsi_oracle.c has the same numerical
 * primitives.  Nothing under xpsi_b200/ may include or link this.
 */
#ifndef XPSI_ORACLE_GSLSHIM_H
#define XPSI_ORACLE_GSLSHIM_H

#include <stddef.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- errno ------------------------------------------------------------ */
enum {
  GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1,
  GSL_ERANGE = 2, GSL_EFAULT = 3, GSL_EINVAL = 4, GSL_EFAILED = 5,
  GSL_EFACTOR = 6, GSL_ESANITY = 7, GSL_ENOMEM = 8, GSL_EBADFUNC = 9,
  GSL_ERUNAWAY = 10, GSL_EMAXITER = 11, GSL_EZERODIV = 12, GSL_EBADTOL = 13,
  GSL_ETOL = 14, GSL_EUNDRFLW = 15, GSL_EOVRFLW = 16, GSL_ELOSS = 17,
  GSL_EROUND = 18, GSL_EBADLEN = 19, GSL_ENOTSQR = 20, GSL_ESING = 21,
  GSL_EDIVERGE = 22, GSL_EUNSUP = 23, GSL_EUNIMPL = 24, GSL_ECACHE = 25,
  GSL_ETABLE = 26, GSL_ENOPROG = 27, GSL_ENOPROGJ = 28, GSL_ETOLF = 29,
  GSL_ETOLX = 30, GSL_ETOLG = 31, GSL_EOF = 32
};

typedef void gsl_error_handler_t(const char *reason, const char *file,
                                 int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler_off(void);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *new_handler);
const char *gsl_strerror(const int gsl_errno);

/* ---- math ------------------------------------------------------------- */
#ifndef M_SQRT3
#define M_SQRT3 1.73205080756887729352744634151
#endif
#ifndef M_SQRTPI
#define M_SQRTPI 1.77245385090551602729816748334
#endif
#ifndef M_LNPI
#define M_LNPI 1.14472988584940017414342735135
#endif
#ifndef M_EULER
#define M_EULER 0.57721566490153286060651209008
#endif
#define GSL_NAN (NAN)
#define GSL_POSINF (INFINITY)
#define GSL_NEGINF (-INFINITY)
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_DBL_MIN 2.2250738585072014e-308

int gsl_isnan(const double x);
int gsl_isinf(const double x);
int gsl_finite(const double x);
double gsl_log1p(const double x);
double gsl_expm1(const double x);
double gsl_hypot(const double x, const double y);
double gsl_acosh(const double x);
double gsl_asinh(const double x);
double gsl_atanh(const double x);
double gsl_ldexp(const double x, const int e);
double gsl_frexp(const double x, int *e);
double gsl_pow_int(double x, int n);
double gsl_pow_2(const double x);
double gsl_pow_3(const double x);
double gsl_pow_4(const double x);
double gsl_pow_5(const double x);
double gsl_pow_6(const double x);
double gsl_pow_7(const double x);
double gsl_pow_8(const double x);
double gsl_pow_9(const double x);

typedef struct {
  double (*function)(double x, void *params);
  void *params;
} gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)

typedef struct {
  double (*f)(double x, void *params);
  double (*df)(double x, void *params);
  void (*fdf)(double x, void *params, double *f, double *df);
  void *params;
} gsl_function_fdf;

/* ---- interpolation ---------------------------------------------------- */
typedef struct {
  size_t cache;
  size_t miss_count;
  size_t hit_count;
} gsl_interp_accel;

typedef struct gsl_interp_type_s {
  const char *name;
  unsigned int min_size;
  int kind;
} gsl_interp_type;

typedef struct {
  const gsl_interp_type *type;
  double xmin;
  double xmax;
  size_t size;
  void *state;
} gsl_interp;

extern const gsl_interp_type *gsl_interp_linear;
extern const gsl_interp_type *gsl_interp_polynomial;
extern const gsl_interp_type *gsl_interp_cspline;
extern const gsl_interp_type *gsl_interp_cspline_periodic;
extern const gsl_interp_type *gsl_interp_akima;
extern const gsl_interp_type *gsl_interp_akima_periodic;
extern const gsl_interp_type *gsl_interp_steffen;

gsl_interp_accel *gsl_interp_accel_alloc(void);
size_t gsl_interp_accel_find(gsl_interp_accel *a, const double x_array[],
                             size_t size, double x);
int gsl_interp_accel_reset(gsl_interp_accel *a);
void gsl_interp_accel_free(gsl_interp_accel *a);

gsl_interp *gsl_interp_alloc(const gsl_interp_type *T, size_t n);
int gsl_interp_init(gsl_interp *obj, const double xa[], const double ya[],
                    size_t size);
const char *gsl_interp_name(const gsl_interp *interp);
unsigned int gsl_interp_min_size(const gsl_interp *interp);
int gsl_interp_eval_e(const gsl_interp *obj, const double xa[],
                      const double ya[], double x, gsl_interp_accel *a,
                      double *y);
double gsl_interp_eval(const gsl_interp *obj, const double xa[],
                       const double ya[], double x, gsl_interp_accel *a);
int gsl_interp_eval_deriv_e(const gsl_interp *obj, const double xa[],
                            const double ya[], double x, gsl_interp_accel *a,
                            double *d);
double gsl_interp_eval_deriv(const gsl_interp *obj, const double xa[],
                             const double ya[], double x, gsl_interp_accel *a);
int gsl_interp_eval_deriv2_e(const gsl_interp *obj, const double xa[],
                             const double ya[], double x, gsl_interp_accel *a,
                             double *d2);
double gsl_interp_eval_deriv2(const gsl_interp *obj, const double xa[],
                              const double ya[], double x,
                              gsl_interp_accel *a);
int gsl_interp_eval_integ_e(const gsl_interp *obj, const double xa[],
                            const double ya[], double a, double b,
                            gsl_interp_accel *acc, double *result);
double gsl_interp_eval_integ(const gsl_interp *obj, const double xa[],
                             const double ya[], double a, double b,
                             gsl_interp_accel *acc);
void gsl_interp_free(gsl_interp *interp);
size_t gsl_interp_bsearch(const double x_array[], double x, size_t index_lo,
                          size_t index_hi);

/* gsl_spline: interp + private copies of the data */
typedef struct {
  gsl_interp *interp;
  double *x;
  double *y;
  size_t size;
} gsl_spline;

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *spline, const double xa[], const double ya[],
                    size_t size);
const char *gsl_spline_name(const gsl_spline *spline);
unsigned int gsl_spline_min_size(const gsl_spline *spline);
int gsl_spline_eval_e(const gsl_spline *spline, double x, gsl_interp_accel *a,
                      double *y);
double gsl_spline_eval(const gsl_spline *spline, double x,
                       gsl_interp_accel *a);
int gsl_spline_eval_deriv_e(const gsl_spline *spline, double x,
                            gsl_interp_accel *a, double *y);
double gsl_spline_eval_deriv(const gsl_spline *spline, double x,
                             gsl_interp_accel *a);
int gsl_spline_eval_deriv2_e(const gsl_spline *spline, double x,
                             gsl_interp_accel *a, double *y);
double gsl_spline_eval_deriv2(const gsl_spline *spline, double x,
                              gsl_interp_accel *a);
int gsl_spline_eval_integ_e(const gsl_spline *spline, double a, double b,
                            gsl_interp_accel *acc, double *y);
double gsl_spline_eval_integ(const gsl_spline *spline, double a, double b,
                             gsl_interp_accel *acc);
void gsl_spline_free(gsl_spline *spline);

/* ---- integration ------------------------------------------------------ */
enum {
  GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21 = 2, GSL_INTEG_GAUSS31 = 3,
  GSL_INTEG_GAUSS41 = 4, GSL_INTEG_GAUSS51 = 5, GSL_INTEG_GAUSS61 = 6
};

typedef struct {
  size_t limit;
  size_t size;
  double *alist;
  double *blist;
  double *rlist;
  double *elist;
} gsl_integration_workspace;

gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qag(const gsl_function *f, double a, double b,
                        double epsabs, double epsrel, size_t limit, int key,
                        gsl_integration_workspace *workspace, double *result,
                        double *abserr);

typedef struct {
  double a, b, igral, err;
} gsl_integration_cquad_ival;

typedef struct {
  size_t size;
  gsl_integration_cquad_ival *ivals;
} gsl_integration_cquad_workspace;

gsl_integration_cquad_workspace *
gsl_integration_cquad_workspace_alloc(const size_t n);
void gsl_integration_cquad_workspace_free(gsl_integration_cquad_workspace *w);
int gsl_integration_cquad(const gsl_function *f, double a, double b,
                          double epsabs, double epsrel,
                          gsl_integration_cquad_workspace *ws, double *result,
                          double *abserr, size_t *nevals);

/* ---- special functions ------------------------------------------------ */
double gsl_sf_lnfact(const unsigned int n);

/* ---- rng (data synthesis only; NOT the GSL generators) ---------------- */
typedef struct gsl_rng_type_s {
  const char *name;
} gsl_rng_type;
typedef struct {
  const gsl_rng_type *type;
  unsigned long long s[4];
} gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
extern unsigned long int gsl_rng_default_seed;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(const gsl_rng *r, unsigned long int s);
void gsl_rng_free(gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
unsigned int gsl_ran_poisson(const gsl_rng *r, double mu);
void gsl_ran_poisson_array(const gsl_rng *r, size_t n, unsigned int array[],
                           double mu);
double gsl_ran_poisson_pdf(const unsigned int k, const double mu);

#ifdef __cplusplus
}
#endif
#endif
