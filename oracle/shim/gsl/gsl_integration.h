/* oracle/shim/gsl/gsl_integration.h -- TEST INFRASTRUCTURE (see gsl_rng.h).
 * qagiu/qng here are NOT GSL's QUADPACK ports: they are a fixed high-order
 * composite Gauss-Legendre rule (cosmo.cc:115,153-154 integrate smooth
 * functions; results enter the hot path only through the scalars D1, D2, which
 * the GPU ABI takes as inputs). */
#ifndef BARCODE_ORACLE_SHIM_GSL_INTEGRATION_H
#define BARCODE_ORACLE_SHIM_GSL_INTEGRATION_H
#include <stddef.h>
#include "gsl_math.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t limit; } gsl_integration_workspace;
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qagiu(gsl_function *f, double a, double epsabs, double epsrel, size_t limit,
                          gsl_integration_workspace *w, double *result, double *abserr);
int gsl_integration_qng(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        double *result, double *abserr, size_t *neval);
#ifdef __cplusplus
}
#endif
#endif
