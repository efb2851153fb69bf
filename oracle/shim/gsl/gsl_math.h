/* oracle/shim/gsl/gsl_math.h -- TEST INFRASTRUCTURE (see gsl_rng.h). */
#ifndef BARCODE_ORACLE_SHIM_GSL_MATH_H
#define BARCODE_ORACLE_SHIM_GSL_MATH_H
#include <math.h>
static inline double gsl_pow_2(double x) { return x * x; }
static inline double gsl_pow_3(double x) { return x * x * x; }
static inline double gsl_pow_5(double x) { double x2 = x * x; return x2 * x2 * x; }
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#endif
