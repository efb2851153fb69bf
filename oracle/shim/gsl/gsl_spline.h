/* oracle/shim/gsl/gsl_spline.h -- TEST INFRASTRUCTURE (see gsl_rng.h).
 * Linear interpolation only (calc_power.cc:72-100), GSL's formula
 * y = y_lo + (x - x_lo) / (x_hi - x_lo) * (y_hi - y_lo). */
#ifndef BARCODE_ORACLE_SHIM_GSL_SPLINE_H
#define BARCODE_ORACLE_SHIM_GSL_SPLINE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { const char *name; } gsl_interp_type;
typedef struct { size_t size; } gsl_interp;
typedef struct { size_t cache; } gsl_interp_accel;
extern const gsl_interp_type *gsl_interp_linear;
gsl_interp *gsl_interp_alloc(const gsl_interp_type *T, size_t n);
int gsl_interp_init(gsl_interp *obj, const double xa[], const double ya[], size_t size);
void gsl_interp_free(gsl_interp *interp);
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
double gsl_interp_eval(const gsl_interp *obj, const double xa[], const double ya[], double x,
                       gsl_interp_accel *a);
#ifdef __cplusplus
}
#endif
#endif
