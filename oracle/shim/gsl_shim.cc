/*
 * oracle/shim/gsl_shim.cc -- TEST INFRASTRUCTURE, not product code.
 *
 * From-scratch implementation of the GSL calls the Barcode reference makes
 * (see the headers in oracle/shim/gsl/).  GSL is absent from this image.  The
 * random-number parts follow GSL's published algorithms exactly so that seeds
 * reproduce the reference's streams:
 *   - gsl_rng_mt19937: MT19937, 2002 seeding, seed 0 -> 4357
 *   - gsl_rng_uniform: u32 / 4294967296.0 ; _uniform_pos rejects 0
 *   - gsl_ran_gaussian: polar Box-Muller, returns sigma*y*sqrt(-2 ln r2 / r2)
 * The raw stream is cross-checked against numpy.random.RandomState in
 * tests/test_oracle_ref.py.  Quadrature and Poisson deviates are NOT GSL's
 * algorithms (noted in the headers); they are off the hot path.
 */
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "gsl/gsl_integration.h"
#include "gsl/gsl_randist.h"
#include "gsl/gsl_rng.h"
#include "gsl/gsl_spline.h"
#include "ncurses.h"

extern "C" {

/* ---------------- RNG ---------------- */
static const gsl_rng_type mt_type = {"mt19937"};
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *) {
  auto *r = static_cast<gsl_rng *>(std::calloc(1, sizeof(gsl_rng)));
  gsl_rng_set(r, 0);
  return r;
}

void gsl_rng_set(gsl_rng *r, unsigned long s) {
  if (s == 0) s = 4357;
  r->mt[0] = s & 0xffffffffUL;
  for (int i = 1; i < 624; ++i)
    r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
  r->mti = 624;
  r->n_forced = 0;
}

void gsl_rng_free(gsl_rng *r) { std::free(r); }

unsigned long gsl_rng_get(gsl_rng *r) {
  const unsigned long UPPER = 0x80000000UL, LOWER = 0x7fffffffUL;
  unsigned long *mt = r->mt;
  if (r->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; ++kk) {
      unsigned long y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    for (; kk < 623; ++kk) {
      unsigned long y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    unsigned long y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    r->mti = 0;
  }
  unsigned long k = mt[r->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  return k & 0xffffffffUL;
}

double gsl_rng_uniform(gsl_rng *r) {
  if (r->n_forced > 0) {
    double u = r->forced[0];
    for (int i = 1; i < r->n_forced; ++i) r->forced[i - 1] = r->forced[i];
    r->n_forced--;
    return u;
  }
  return gsl_rng_get(r) / 4294967296.0;
}

double gsl_rng_uniform_pos(gsl_rng *r) {
  double x;
  do {
    x = gsl_rng_get(r) / 4294967296.0;
  } while (x == 0);
  return x;
}

void shim_gsl_rng_force_uniform(gsl_rng *r, double u) {
  if (r->n_forced < 8) r->forced[r->n_forced++] = u;
}

double gsl_ran_gaussian(gsl_rng *r, double sigma) {
  double x, y, r2;
  do {
    x = -1 + 2 * gsl_rng_uniform_pos(r);
    y = -1 + 2 * gsl_rng_uniform_pos(r);
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0);
  return sigma * y * std::sqrt(-2.0 * std::log(r2) / r2);
}

double gsl_ran_ugaussian(gsl_rng *r) { return gsl_ran_gaussian(r, 1.0); }

double gsl_ran_gaussian_ziggurat(gsl_rng *, double) {
  std::fprintf(stderr, "gsl shim: gsl_ran_gaussian_ziggurat is not implemented\n");
  std::abort();
}
double gsl_ran_gaussian_ratio_method(gsl_rng *, double) {
  std::fprintf(stderr, "gsl shim: gsl_ran_gaussian_ratio_method is not implemented\n");
  std::abort();
}

unsigned int gsl_ran_poisson(gsl_rng *r, double mu) {
  /* mock-data generation only (barcoderunner.cc:131); product-of-uniforms for
     small mu as in GSL, Gaussian approximation above (not GSL's gamma/binomial
     recursion) */
  if (mu > 10) {
    double v = mu + std::sqrt(mu) * gsl_ran_gaussian(r, 1.0) + 0.5;
    return v < 0 ? 0u : static_cast<unsigned int>(v);
  }
  const double emu = std::exp(-mu);
  double prod = 1.0;
  unsigned int k = 0;
  do {
    prod *= gsl_rng_uniform(r);
    k++;
  } while (prod > emu);
  return k - 1;
}

/* ---------------- quadrature ---------------- */
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n) {
  auto *w = static_cast<gsl_integration_workspace *>(std::malloc(sizeof(gsl_integration_workspace)));
  w->limit = n;
  return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) { std::free(w); }

static double gl_panel(const gsl_function *f, double a, double b) {
  /* 8-point Gauss-Legendre */
  static const double x[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
  static const double w[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
  const double c = 0.5 * (a + b), h = 0.5 * (b - a);
  double s = 0;
  for (int i = 0; i < 4; ++i) s += w[i] * (f->function(c - h * x[i], f->params) + f->function(c + h * x[i], f->params));
  return s * h;
}

static double composite(const gsl_function *f, double a, double b, int panels) {
  double s = 0;
  const double h = (b - a) / panels;
  for (int i = 0; i < panels; ++i) s += gl_panel(f, a + i * h, a + (i + 1) * h);
  return s;
}

struct iu_params { const gsl_function *f; double a; };
static double iu_transform(double t, void *p) {
  /* x = a + (1-t)/t, dx = dt / t^2 */
  auto *q = static_cast<iu_params *>(p);
  if (t <= 0) return 0;
  const double x = q->a + (1 - t) / t;
  return q->f->function(x, q->f->params) / (t * t);
}

int gsl_integration_qagiu(gsl_function *f, double a, double, double, size_t, gsl_integration_workspace *,
                          double *result, double *abserr) {
  iu_params q = {f, a};
  gsl_function g = {&iu_transform, &q};
  const double r1 = composite(&g, 0.0, 1.0, 2048);
  const double r2 = composite(&g, 0.0, 1.0, 4096);
  *result = r2;
  if (abserr) *abserr = std::fabs(r2 - r1);
  return 0;
}

int gsl_integration_qng(const gsl_function *f, double a, double b, double, double, double *result, double *abserr,
                        size_t *neval) {
  const double r1 = composite(f, a, b, 64);
  const double r2 = composite(f, a, b, 128);
  *result = r2;
  if (abserr) *abserr = std::fabs(r2 - r1);
  if (neval) *neval = 128 * 8;
  return 0;
}

/* ---------------- linear interpolation ---------------- */
static const gsl_interp_type lin_type = {"linear"};
const gsl_interp_type *gsl_interp_linear = &lin_type;

gsl_interp *gsl_interp_alloc(const gsl_interp_type *, size_t n) {
  auto *p = static_cast<gsl_interp *>(std::malloc(sizeof(gsl_interp)));
  p->size = n;
  return p;
}
int gsl_interp_init(gsl_interp *obj, const double[], const double[], size_t size) {
  obj->size = size;
  return 0;
}
void gsl_interp_free(gsl_interp *p) { std::free(p); }
gsl_interp_accel *gsl_interp_accel_alloc(void) {
  return static_cast<gsl_interp_accel *>(std::calloc(1, sizeof(gsl_interp_accel)));
}
void gsl_interp_accel_free(gsl_interp_accel *a) { std::free(a); }

double gsl_interp_eval(const gsl_interp *obj, const double xa[], const double ya[], double x, gsl_interp_accel *) {
  const size_t n = obj->size;
  if (x < xa[0] || x > xa[n - 1]) {
    std::fprintf(stderr, "gsl shim: interpolation point %g outside table [%g, %g]\n", x, xa[0], xa[n - 1]);
    std::abort(); /* GSL's default error handler aborts on GSL_EDOM too */
  }
  size_t lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    const size_t mid = (lo + hi) / 2;
    if (xa[mid] > x) hi = mid; else lo = mid;
  }
  const double dx = xa[lo + 1] - xa[lo];
  if (dx > 0.0) return ya[lo] + (x - xa[lo]) / dx * (ya[lo + 1] - ya[lo]);
  return 0.0;
}

/* ---------------- ncurses no-ops ---------------- */
static WINDOW the_window;
WINDOW *stdscr = &the_window;
WINDOW *initscr(void) { return &the_window; }
int endwin(void) { return 0; }
int isendwin(void) { return 1; }
int refresh(void) { return 0; }
int start_color(void) { return 0; }
int cbreak(void) { return 0; }
int noecho(void) { return 0; }
int intrflush(WINDOW *, int) { return 0; }
int keypad(WINDOW *, int) { return 0; }
int init_pair(short, short, short) { return 0; }
WINDOW *newwin(int, int, int, int) { return &the_window; }
int wattron(WINDOW *, int) { return 0; }
int wbkgd(WINDOW *, unsigned) { return 0; }
int scrollok(WINDOW *, int) { return 0; }
int wprintw(WINDOW *, const char *, ...) { return 0; }
int wrefresh(WINDOW *) { return 0; }

}  // extern "C"
