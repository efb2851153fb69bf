/*
 * oracle/shim/gsl/gsl_rng.h -- TEST INFRASTRUCTURE, not product code.
 * Stand-in for the slice of GSL's RNG API the Barcode reference uses
 * (main.cc:107,143; HMC.cc:260-261,480; random.hpp:24-25).  GSL is not
 * installed in this image.  Implements GSL's published algorithms:
 * gsl_rng_mt19937 = MT19937 with the 2002 seeding (seed 0 -> 4357),
 * gsl_rng_uniform = genrand_int32 / 4294967296.0.
 */
#ifndef BARCODE_ORACLE_SHIM_GSL_RNG_H
#define BARCODE_ORACLE_SHIM_GSL_RNG_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { const char *name; } gsl_rng_type;
typedef struct {
  unsigned long mt[624];
  int mti;
  /* shim-only: queued values returned by gsl_rng_uniform before the stream
     (lets the oracle harness pin Neps / epsilon, HMC.cc:260-261) */
  double forced[8];
  int n_forced;
} gsl_rng;

extern const gsl_rng_type *gsl_rng_mt19937;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
void gsl_rng_free(gsl_rng *r);
unsigned long gsl_rng_get(gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
double gsl_rng_uniform_pos(gsl_rng *r);
void shim_gsl_rng_force_uniform(gsl_rng *r, double u);

#ifdef __cplusplus
}
#endif
#endif
