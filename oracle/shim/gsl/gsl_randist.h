/* oracle/shim/gsl/gsl_randist.h -- TEST INFRASTRUCTURE (see gsl_rng.h).
 * gsl_ran_gaussian: GSL's polar Box-Muller (x discarded, y returned). */
#ifndef BARCODE_ORACLE_SHIM_GSL_RANDIST_H
#define BARCODE_ORACLE_SHIM_GSL_RANDIST_H
#include "gsl_rng.h"
#ifdef __cplusplus
extern "C" {
#endif
double gsl_ran_gaussian(gsl_rng *r, double sigma);
double gsl_ran_ugaussian(gsl_rng *r);
double gsl_ran_gaussian_ziggurat(gsl_rng *r, double sigma);     /* aborts: unused (GR_METHOD 0 only) */
double gsl_ran_gaussian_ratio_method(gsl_rng *r, double sigma); /* aborts: unused */
unsigned int gsl_ran_poisson(gsl_rng *r, double mu);
#ifdef __cplusplus
}
#endif
#endif
