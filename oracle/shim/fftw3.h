/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal stand-in for the FFTW 3 API surface the Barcode reference uses
 * (grep of /root/reference/barlib: fftw_malloc/free, fftw_plan_dft_{r2c,c2r}_3d,
 * fftw_plan_dft_3d, fftw_execute, fftw_destroy_plan and the threads calls).
 * FFTW itself is not installed in this image and there is no network, so the
 * oracle build (oracle/Makefile) compiles the UNMODIFIED reference sources
 * against this header and links oracle/shim/fftw_shim.cc, a from-scratch
 * OpenMP power-of-two pencil FFT.  The DFT is mathematically defined, so any
 * correct FP64 FFT agrees with FFTW to ~1e-15 relative; the shim is checked
 * against numpy.fft in tests/test_oracle_ref.py.
 */
#ifndef BARCODE_ORACLE_SHIM_FFTW3_H
#define BARCODE_ORACLE_SHIM_FFTW3_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct shim_fftw_plan_s *fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

void *fftw_malloc(size_t n);
void fftw_free(void *p);

fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in,
                               fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex *in,
                               double *out, unsigned flags);
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in,
                           fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
void fftw_cleanup_threads(void);

/* single-precision API (the reference's SINGLE_PREC build, fftwrapper.cc:32-36,62-66): each execute widens the
 * input to double, runs the double transform and rounds the result to float once -- a correctly rounded
 * single-precision DFT, at least as accurate as any float FFT */
typedef float fftwf_complex[2];
typedef struct shim_fftwf_plan_s *fftwf_plan;
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_r2c_3d(int n0, int n1, int n2, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in, float *out, unsigned flags);
fftwf_plan fftwf_plan_dft_3d(int n0, int n1, int n2, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_init_threads(void);
void fftwf_plan_with_nthreads(int nthreads);
void fftwf_cleanup_threads(void);

/* shim-only: which backend is linked (reported by bench.py's cpu_baseline) */
const char *shim_fftw_backend(void);

#ifdef __cplusplus
}
#endif

#endif
