/*
 * oracle/shim/fftw_shim.cc -- TEST INFRASTRUCTURE, not product code.
 *
 * From-scratch implementation of the slice of the FFTW 3 API declared in
 * oracle/shim/fftw3.h, so that the unmodified Barcode reference sources can be
 * compiled and run in an image without FFTW (see oracle/Makefile).
 *
 * Algorithm: 3-D transforms as three passes of batched 1-D Stockham autosort
 * FFTs (radix 4 with one radix-2 step when log2(n) is odd), eight pencils at a
 * time in split re/im layout so the compiler vectorises over the batch, OpenMP
 * over tiles.  Power-of-two sizes only (every BASELINE.json grid and the
 * reference's own test/run config, Nx=8, are powers of two); anything else
 * aborts loudly.  Conventions are FFTW's: forward sign -1, both directions
 * unnormalised, r2c output / c2r input is the n0 x n1 x (n2/2+1) half array,
 * c2r may destroy its input.
 */
#include "fftw3.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

enum Kind { R2C, C2R, C2C };

constexpr int B = 8;  // pencils per batch

struct Twiddle {
  int n = 0;
  std::vector<double> c, s;  // cos(2 pi k/n), sin(2 pi k/n), k < n
  void init(int n_) {
    n = n_;
    c.resize(n);
    s.resize(n);
    for (int k = 0; k < n; ++k) {
      // exact symmetries first, libm elsewhere
      const long double a = 2.0L * 3.141592653589793238462643383279502884L * k / n;
      c[k] = static_cast<double>(cosl(a));
      s[k] = static_cast<double>(sinl(a));
    }
  }
};

bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

/*
 * Batched Stockham FFT of length n on `w` interleaved pencils (w = batch
 * width, the fastest index).  Data: xr/xi[n*w]; scratch yr/yi[n*w].  sign = -1
 * forward, +1 backward.  Result ends in xr/xi.
 */
#if defined(__GNUC__) && defined(__x86_64__) && !defined(__clang__)
__attribute__((target_clones("default", "avx2", "avx512f")))
#endif
void stockham(int n, int w, int sign, const Twiddle &tw, double *xr, double *xi,
              double *yr, double *yi) {
  double *ar = xr, *ai = xi, *br = yr, *bi = yi;
  int len = n;  // current sub-transform length
  int s = w;    // current stride (in doubles)
  const double sg = (sign < 0) ? -1.0 : 1.0;
  while (len > 1) {
    if (len % 4 == 0) {
      const int n1 = len / 4;
      const int tstep = tw.n / len;
      for (int p = 0; p < n1; ++p) {
        const double w1r = tw.c[p * tstep], w1i = sg * tw.s[p * tstep];
        const double w2r = tw.c[2 * p * tstep], w2i = sg * tw.s[2 * p * tstep];
        const double w3r = tw.c[3 * p * tstep], w3i = sg * tw.s[3 * p * tstep];
        const double *a_r = ar + (size_t)s * p, *a_i = ai + (size_t)s * p;
        const double *b_r = a_r + (size_t)s * n1, *b_i = a_i + (size_t)s * n1;
        const double *c_r = b_r + (size_t)s * n1, *c_i = b_i + (size_t)s * n1;
        const double *d_r = c_r + (size_t)s * n1, *d_i = c_i + (size_t)s * n1;
        double *o0r = br + (size_t)s * (4 * p), *o0i = bi + (size_t)s * (4 * p);
        double *o1r = o0r + s, *o1i = o0i + s;
        double *o2r = o1r + s, *o2i = o1i + s;
        double *o3r = o2r + s, *o3i = o2i + s;
        for (int q = 0; q < s; ++q) {
          const double apcr = a_r[q] + c_r[q], apci = a_i[q] + c_i[q];
          const double amcr = a_r[q] - c_r[q], amci = a_i[q] - c_i[q];
          const double bpdr = b_r[q] + d_r[q], bpdi = b_i[q] + d_i[q];
          const double bmdr = b_r[q] - d_r[q], bmdi = b_i[q] - d_i[q];
          // sg*i*(b-d): forward (sg=-1) -> -i(b-d) = (bmdi, -bmdr)
          const double jr = -sg * bmdi, ji = sg * bmdr;
          o0r[q] = apcr + bpdr;
          o0i[q] = apci + bpdi;
          const double t1r = amcr + jr, t1i = amci + ji;
          o1r[q] = t1r * w1r - t1i * w1i;
          o1i[q] = t1r * w1i + t1i * w1r;
          const double t2r = apcr - bpdr, t2i = apci - bpdi;
          o2r[q] = t2r * w2r - t2i * w2i;
          o2i[q] = t2r * w2i + t2i * w2r;
          const double t3r = amcr - jr, t3i = amci - ji;
          o3r[q] = t3r * w3r - t3i * w3i;
          o3i[q] = t3r * w3i + t3i * w3r;
        }
      }
      len /= 4;
      s *= 4;
    } else {
      const int m = len / 2;
      const int tstep = tw.n / len;
      for (int p = 0; p < m; ++p) {
        const double wr = tw.c[p * tstep], wi = sg * tw.s[p * tstep];
        const double *a_r = ar + (size_t)s * p, *a_i = ai + (size_t)s * p;
        const double *b_r = a_r + (size_t)s * m, *b_i = a_i + (size_t)s * m;
        double *o0r = br + (size_t)s * (2 * p), *o0i = bi + (size_t)s * (2 * p);
        double *o1r = o0r + s, *o1i = o0i + s;
        for (int q = 0; q < s; ++q) {
          const double tr = a_r[q] - b_r[q], ti = a_i[q] - b_i[q];
          o0r[q] = a_r[q] + b_r[q];
          o0i[q] = a_i[q] + b_i[q];
          o1r[q] = tr * wr - ti * wi;
          o1i[q] = tr * wi + ti * wr;
        }
      }
      len /= 2;
      s *= 2;
    }
    double *t;
    t = ar; ar = br; br = t;
    t = ai; ai = bi; bi = t;
  }
  if (ar != xr) {
    std::memcpy(xr, ar, sizeof(double) * (size_t)n * w);
    std::memcpy(xi, ai, sizeof(double) * (size_t)n * w);
  }
}

struct Scratch {
  std::vector<double> xr, xi, yr, yi;
  void ensure(size_t m) {
    if (xr.size() < m) { xr.resize(m); xi.resize(m); yr.resize(m); yi.resize(m); }
  }
};

/*
 * In-place strided complex pass along one axis of a complex array laid out
 * [n0][n1][nc]: axis 0 (stride n1*nc) or axis 1 (stride nc).  Pencils are
 * batched along the contiguous index.
 */
void strided_pass(double (*a)[2], int n0, int n1, int nc, int axis, int sign,
                  const Twiddle &tw) {
  const int n = (axis == 0) ? n0 : n1;
  const size_t stride = (axis == 0) ? (size_t)n1 * nc : (size_t)nc;
  const int outer = (axis == 0) ? n1 : n0;  // the remaining non-contiguous axis
  const size_t ostride = (axis == 0) ? (size_t)nc : (size_t)n1 * nc;
  const int nchunks = (nc + B - 1) / B;
  const long ntiles = (long)outer * nchunks;
#pragma omp parallel
  {
    Scratch sc;
    sc.ensure((size_t)n * B);
#pragma omp for schedule(static)
    for (long t = 0; t < ntiles; ++t) {
      const int o = (int)(t / nchunks);
      const int c0 = (int)(t % nchunks) * B;
      const int w = (nc - c0 < B) ? (nc - c0) : B;
      double (*base)[2] = a + (size_t)o * ostride + c0;
      for (int i = 0; i < n; ++i) {
        const double (*src)[2] = base + (size_t)i * stride;
        for (int q = 0; q < w; ++q) {
          sc.xr[(size_t)i * w + q] = src[q][0];
          sc.xi[(size_t)i * w + q] = src[q][1];
        }
      }
      stockham(n, w, sign, tw, sc.xr.data(), sc.xi.data(), sc.yr.data(), sc.yi.data());
      for (int i = 0; i < n; ++i) {
        double (*dst)[2] = base + (size_t)i * stride;
        for (int q = 0; q < w; ++q) {
          dst[q][0] = sc.xr[(size_t)i * w + q];
          dst[q][1] = sc.xi[(size_t)i * w + q];
        }
      }
    }
  }
}

/* contiguous-axis complex pass (c2c only): rows of length n2 */
void contiguous_pass(double (*a)[2], long nrows, int n2, int sign, const Twiddle &tw) {
  const long ntiles = (nrows + B - 1) / B;
#pragma omp parallel
  {
    Scratch sc;
    sc.ensure((size_t)n2 * B);
#pragma omp for schedule(static)
    for (long t = 0; t < ntiles; ++t) {
      const long r0 = t * B;
      const int w = (int)((nrows - r0 < B) ? (nrows - r0) : B);
      for (int q = 0; q < w; ++q) {
        const double (*src)[2] = a + (size_t)(r0 + q) * n2;
        for (int k = 0; k < n2; ++k) {
          sc.xr[(size_t)k * w + q] = src[k][0];
          sc.xi[(size_t)k * w + q] = src[k][1];
        }
      }
      stockham(n2, w, sign, tw, sc.xr.data(), sc.xi.data(), sc.yr.data(), sc.yi.data());
      for (int q = 0; q < w; ++q) {
        double (*dst)[2] = a + (size_t)(r0 + q) * n2;
        for (int k = 0; k < n2; ++k) {
          dst[k][0] = sc.xr[(size_t)k * w + q];
          dst[k][1] = sc.xi[(size_t)k * w + q];
        }
      }
    }
  }
}

/*
 * Real rows -> half-complex rows.  Two real rows are transformed as one
 * complex sequence z = r1 + i r2 and separated with
 *   X1[k] = (Z[k] + conj Z[n-k]) / 2,  X2[k] = (Z[k] - conj Z[n-k]) / (2i).
 */
void r2c_rows(const double *in, double (*out)[2], long nrows, int n2, const Twiddle &tw) {
  const int nc = n2 / 2 + 1;
  const long npairs = (nrows + 1) / 2;
  const long ntiles = (npairs + B - 1) / B;
#pragma omp parallel
  {
    Scratch sc;
    sc.ensure((size_t)n2 * B);
#pragma omp for schedule(static)
    for (long t = 0; t < ntiles; ++t) {
      const long p0 = t * B;
      const int w = (int)((npairs - p0 < B) ? (npairs - p0) : B);
      for (int q = 0; q < w; ++q) {
        const long r1 = 2 * (p0 + q), r2 = r1 + 1;
        const double *s1 = in + (size_t)r1 * n2;
        const double *s2 = (r2 < nrows) ? in + (size_t)r2 * n2 : nullptr;
        for (int k = 0; k < n2; ++k) {
          sc.xr[(size_t)k * w + q] = s1[k];
          sc.xi[(size_t)k * w + q] = s2 ? s2[k] : 0.0;
        }
      }
      stockham(n2, w, -1, tw, sc.xr.data(), sc.xi.data(), sc.yr.data(), sc.yi.data());
      for (int q = 0; q < w; ++q) {
        const long r1 = 2 * (p0 + q), r2 = r1 + 1;
        double (*d1)[2] = out + (size_t)r1 * nc;
        double (*d2)[2] = (r2 < nrows) ? out + (size_t)r2 * nc : nullptr;
        for (int k = 0; k < nc; ++k) {
          const int km = (n2 - k) & (n2 - 1);
          const double zr = sc.xr[(size_t)k * w + q], zi = sc.xi[(size_t)k * w + q];
          const double mr = sc.xr[(size_t)km * w + q], mi = sc.xi[(size_t)km * w + q];
          d1[k][0] = 0.5 * (zr + mr);
          d1[k][1] = 0.5 * (zi - mi);
          if (d2) {
            d2[k][0] = 0.5 * (zi + mi);
            d2[k][1] = 0.5 * (mr - zr);
          }
        }
      }
    }
  }
}

/* Half-complex rows -> real rows (unnormalised), pairs packed as Z = X1 + i X2 */
void c2r_rows(const double (*in)[2], double *out, long nrows, int n2, const Twiddle &tw) {
  const int nc = n2 / 2 + 1;
  const long npairs = (nrows + 1) / 2;
  const long ntiles = (npairs + B - 1) / B;
#pragma omp parallel
  {
    Scratch sc;
    sc.ensure((size_t)n2 * B);
#pragma omp for schedule(static)
    for (long t = 0; t < ntiles; ++t) {
      const long p0 = t * B;
      const int w = (int)((npairs - p0 < B) ? (npairs - p0) : B);
      for (int q = 0; q < w; ++q) {
        const long r1 = 2 * (p0 + q), r2 = r1 + 1;
        const double (*s1)[2] = in + (size_t)r1 * nc;
        const double (*s2)[2] = (r2 < nrows) ? in + (size_t)r2 * nc : nullptr;
        for (int k = 0; k < nc; ++k) {
          double ar = s1[k][0], ai = s1[k][1];
          // FFTW's c2r ignores the imaginary parts of the self-conjugate bins
          if (k == 0 || k == n2 / 2) ai = 0.0;
          double br = s2 ? s2[k][0] : 0.0, bi = s2 ? s2[k][1] : 0.0;
          if (k == 0 || k == n2 / 2) bi = 0.0;
          // Z[k] = A + iB ; Z[n-k] = conj(A) + i conj(B)
          sc.xr[(size_t)k * w + q] = ar - bi;
          sc.xi[(size_t)k * w + q] = ai + br;
          if (k > 0 && k < n2 / 2) {
            sc.xr[(size_t)(n2 - k) * w + q] = ar + bi;
            sc.xi[(size_t)(n2 - k) * w + q] = br - ai;
          }
        }
      }
      stockham(n2, w, +1, tw, sc.xr.data(), sc.xi.data(), sc.yr.data(), sc.yi.data());
      for (int q = 0; q < w; ++q) {
        const long r1 = 2 * (p0 + q), r2 = r1 + 1;
        double *d1 = out + (size_t)r1 * n2;
        double *d2 = (r2 < nrows) ? out + (size_t)r2 * n2 : nullptr;
        for (int k = 0; k < n2; ++k) {
          d1[k] = sc.xr[(size_t)k * w + q];
          if (d2) d2[k] = sc.xi[(size_t)k * w + q];
        }
      }
    }
  }
}

}  // namespace

struct shim_fftw_plan_s {
  Kind kind;
  int n0, n1, n2, sign;
  void *in, *out;
  Twiddle t0, t1, t2;
};

static fftw_plan make_plan(Kind kind, int n0, int n1, int n2, void *in, void *out, int sign) {
  if (!is_pow2(n0) || !is_pow2(n1) || !is_pow2(n2)) {
    std::fprintf(stderr, "fftw shim: only power-of-two sizes are supported (%d x %d x %d)\n", n0, n1, n2);
    std::abort();
  }
  auto *p = new shim_fftw_plan_s;
  p->kind = kind;
  p->n0 = n0; p->n1 = n1; p->n2 = n2;
  p->sign = sign;
  p->in = in; p->out = out;
  p->t0.init(n0); p->t1.init(n1); p->t2.init(n2);
  return p;
}

extern "C" {

void *fftw_malloc(size_t n) {
  void *p = nullptr;
  if (posix_memalign(&p, 64, n ? n : 64) != 0) return nullptr;
  return p;
}
void fftw_free(void *p) { std::free(p); }

fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in, fftw_complex *out, unsigned) {
  return make_plan(R2C, n0, n1, n2, in, out, -1);
}
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex *in, double *out, unsigned) {
  return make_plan(C2R, n0, n1, n2, in, out, +1);
}
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign, unsigned) {
  return make_plan(C2C, n0, n1, n2, in, out, sign);
}

static double g_exec_seconds = 0.0;
static long g_exec_calls = 0;
static double now_seconds() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
/* wall time spent inside fftw_execute since the last reset (called from one host thread, as the reference does) */
void shim_fftw_stats(double *seconds, long *calls, int reset) {
  if (seconds) *seconds = g_exec_seconds;
  if (calls) *calls = g_exec_calls;
  if (reset) { g_exec_seconds = 0.0; g_exec_calls = 0; }
}

void fftw_execute(const fftw_plan p) {
  const double t_begin = now_seconds();
  const int n0 = p->n0, n1 = p->n1, n2 = p->n2;
  const int nc = n2 / 2 + 1;
  switch (p->kind) {
    case R2C: {
      auto *out = static_cast<double (*)[2]>(p->out);
      r2c_rows(static_cast<const double *>(p->in), out, (long)n0 * n1, n2, p->t2);
      strided_pass(out, n0, n1, nc, 1, -1, p->t1);
      strided_pass(out, n0, n1, nc, 0, -1, p->t0);
      break;
    }
    case C2R: {
      auto *in = static_cast<double (*)[2]>(p->in);
      strided_pass(in, n0, n1, nc, 0, +1, p->t0);
      strided_pass(in, n0, n1, nc, 1, +1, p->t1);
      c2r_rows(in, static_cast<double *>(p->out), (long)n0 * n1, n2, p->t2);
      break;
    }
    case C2C: {
      auto *in = static_cast<double (*)[2]>(p->in);
      auto *out = static_cast<double (*)[2]>(p->out);
      const size_t n = (size_t)n0 * n1 * n2;
      if (in != out) std::memcpy(out, in, n * sizeof(double[2]));
      contiguous_pass(out, (long)n0 * n1, n2, p->sign, p->t2);
      strided_pass(out, n0, n1, n2, 1, p->sign, p->t1);
      strided_pass(out, n0, n1, n2, 0, p->sign, p->t0);
      break;
    }
  }
  g_exec_seconds += now_seconds() - t_begin;
  ++g_exec_calls;
}

void fftw_destroy_plan(fftw_plan p) { delete p; }

int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
void fftw_cleanup_threads(void) {}

/* ---- single precision: widen, transform in double, round once ---- */
struct shim_fftwf_plan_s {
  Kind kind;
  int n0, n1, n2, sign;
  void *in, *out;
  fftw_plan dplan;
  double *dreal;          /* n0*n1*n2 (r2c / c2r) */
  double (*dcplx)[2];     /* n0*n1*(n2/2+1) (r2c / c2r) or n0*n1*n2 (c2c) */
};

static fftwf_plan make_plan_f(Kind kind, int n0, int n1, int n2, void *in, void *out, int sign) {
  auto *p = new shim_fftwf_plan_s;
  p->kind = kind;
  p->n0 = n0; p->n1 = n1; p->n2 = n2; p->sign = sign;
  p->in = in; p->out = out;
  const size_t nr = (size_t)n0 * n1 * n2, nc = (size_t)n0 * n1 * (n2 / 2 + 1);
  p->dreal = nullptr;
  if (kind == C2C) {
    p->dcplx = static_cast<double (*)[2]>(fftw_malloc(nr * sizeof(double[2])));
    p->dplan = make_plan(C2C, n0, n1, n2, p->dcplx, p->dcplx, sign);
  } else {
    p->dreal = static_cast<double *>(fftw_malloc(nr * sizeof(double)));
    p->dcplx = static_cast<double (*)[2]>(fftw_malloc(nc * sizeof(double[2])));
    p->dplan = kind == R2C ? make_plan(R2C, n0, n1, n2, p->dreal, p->dcplx, -1)
                           : make_plan(C2R, n0, n1, n2, p->dcplx, p->dreal, +1);
  }
  return p;
}

void *fftwf_malloc(size_t n) { return fftw_malloc(n); }
void fftwf_free(void *p) { std::free(p); }
fftwf_plan fftwf_plan_dft_r2c_3d(int n0, int n1, int n2, float *in, fftwf_complex *out, unsigned) {
  return make_plan_f(R2C, n0, n1, n2, in, out, -1);
}
fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in, float *out, unsigned) {
  return make_plan_f(C2R, n0, n1, n2, in, out, +1);
}
fftwf_plan fftwf_plan_dft_3d(int n0, int n1, int n2, fftwf_complex *in, fftwf_complex *out, int sign, unsigned) {
  return make_plan_f(C2C, n0, n1, n2, in, out, sign);
}
void fftwf_execute(const fftwf_plan p) {
  const size_t nr = (size_t)p->n0 * p->n1 * p->n2, nc = (size_t)p->n0 * p->n1 * (p->n2 / 2 + 1);
  const float *in = static_cast<const float *>(p->in);
  float *out = static_cast<float *>(p->out);
  double *dc = &p->dcplx[0][0];
  if (p->kind == R2C) {
#pragma omp parallel for
    for (long i = 0; i < (long)nr; ++i) p->dreal[i] = (double)in[i];
    fftw_execute(p->dplan);
#pragma omp parallel for
    for (long i = 0; i < (long)(2 * nc); ++i) out[i] = (float)dc[i];
  } else if (p->kind == C2R) {
#pragma omp parallel for
    for (long i = 0; i < (long)(2 * nc); ++i) dc[i] = (double)in[i];
    fftw_execute(p->dplan);
#pragma omp parallel for
    for (long i = 0; i < (long)nr; ++i) out[i] = (float)p->dreal[i];
  } else {
#pragma omp parallel for
    for (long i = 0; i < (long)(2 * nr); ++i) dc[i] = (double)in[i];
    fftw_execute(p->dplan);
#pragma omp parallel for
    for (long i = 0; i < (long)(2 * nr); ++i) out[i] = (float)dc[i];
  }
}
void fftwf_destroy_plan(fftwf_plan p) {
  fftw_destroy_plan(p->dplan);
  std::free(p->dreal);
  std::free(p->dcplx);
  delete p;
}
int fftwf_init_threads(void) { return 1; }
void fftwf_plan_with_nthreads(int) {}
void fftwf_cleanup_threads(void) {}

const char *shim_fftw_backend(void) {
  return "oracle/shim/fftw_shim.cc (own OpenMP Stockham radix-4 pencil FFT, not FFTW)";
}

}  // extern "C"
