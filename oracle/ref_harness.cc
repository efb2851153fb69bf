/*
 * oracle/ref_harness.cc -- TEST INFRASTRUCTURE, not product code.
 *
 * C-callable driver around the UNMODIFIED Barcode reference sources
 * (/root/reference/barlib/src, compiled where they lie by oracle/Makefile into
 * oracle/_ref/libbarcode_ref.so).  It builds `DATA` / `HAMIL_DATA` the way
 * barcode/main.cc:106-154, init_par.cc:41-416 and call_hamil.cc:38-44 do, but
 * from a parameter struct instead of ./input.par, and exposes the hot-path
 * entry points of HMC.cc (seams S1-S6) so that tests and
 * bench.py's cpu_baseline leg can run the reference itself on the same inputs
 * as the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 */
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <omp.h>

#include "struct_main.h"
#include "struct_hamil.h"
#include "fftw_array.h"
#include "fftwrapper.h"
#include "init_par.h"
#include "calc_power.h"
#include "HMC.h"
#include "convolution.hpp"
#include "HMC_momenta.h"
#include "HMC_mass.h"
#include "HMC_help.h"
#include "Lag2Eul.h"
#include "SPH_kernel.hpp"
#include "random.hpp"
#include "cosmo.h"
#include "curses_funcs.h"
#include "massFunctions.h"
#include "gradient.hpp"
#include "field_statistics.h"

#include "gsl/gsl_rng.h"

/* non-static but undeclared in HMC.h (HMC.cc:64,124,146,209,251) */
real_prec kinetic_term(struct HAMIL_DATA *hd, real_prec *momenta, struct DATA *data);
real_prec psi(struct HAMIL_DATA *hd, real_prec *signal, struct DATA *data);
void gradient_psi(struct HAMIL_DATA *hd, real_prec *signal, struct DATA *data);
real_prec delta_Hamiltonian(struct HAMIL_DATA *hd, real_prec *signali, real_prec *momentai, real_prec *signalf,
                            real_prec *momentaf, struct DATA *data);
void Hamiltonian_EoM(struct HAMIL_DATA *hd, real_prec *signali, real_prec *momentai, real_prec *signalf,
                     real_prec *momentaf, gsl_rng *seed, struct DATA *data);

/* arrays are real_prec (double in the default build, float under SINGLE_PREC -- oracle/Makefile target
 * libbarcode_ref_sp.so); scalars are double in both */
typedef real_prec rp_t;

extern "C" {

struct ref_params {
  int N1;
  double L1;
  double xllc, yllc, zllc;
  double xobs, yobs, zobs;
  int planepar, periodic;
  int masskernel;    /* 0 NGP, 1 CIC, 2 TSC, 3 SPH */
  int likelihood;    /* 0 Poisson, 1 Gaussian, 2 lognormal, 3 GRF */
  int sfmodel;       /* 1 Zel'dovich */
  int rsd_model;
  int calc_h;        /* 0..3 */
  int mass_type;     /* 0, 1, 4 on the path */
  double z;
  double deltaQ_factor;
  int correct_delta;
  double particle_kernel_h_rel;
  double slength;
  double N_eps_fac, eps_fac;
  double mass_factor;
  int div_dH_by_N;
  double sigma_min, sigma_fac, delta_min;
  int N_bin;
};

struct ref_handle {
  DATA *data = nullptr;
  HAMIL_DATA *hd = nullptr;
  gsl_rng *rng = nullptr;
  CURSES_STRUCT *curses = nullptr;
  std::vector<fftw_array<real_prec> *> real_arrays;
  fftw_array<complex_prec> *in_c2r = nullptr, *out_r2c = nullptr;
  std::string err;
};

static thread_local std::string g_err;

const char *ref_last_error(void) { return g_err.c_str(); }

#define REF_TRY try {
#define REF_CATCH                          \
  }                                        \
  catch (const std::exception &e) {        \
    g_err = e.what();                      \
    return 1;                              \
  }                                        \
  return 0;

static fftw_array<real_prec> *new_real(ref_handle *h, ULONG N) {
  auto *a = new fftw_array<real_prec>(N);
  std::memset(a->data, 0, N * sizeof(real_prec));
  h->real_arrays.push_back(a);
  return a;
}

void *ref_create(const ref_params *p) {
  auto *h = new ref_handle;
  try {
    /* silence the reference's chatter on stdout */
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());

    DATA *data = new DATA;
    h->data = data;
    NUMERICAL *n = data->numerical;
    n->codename = "BARCODE";
    n->rejections = 0;
    n->INV_SUCCESS = 0;
    n->random_test = true;
    n->random_test_rsd = false;
    n->window_type = 1;
    n->data_model = 0;
    n->negative_obs = false;
    n->likelihood = p->likelihood;
    n->prior = 0;
    n->sfmodel = p->sfmodel;
    n->rsd_model = p->rsd_model != 0;
    n->N_eps_fac = p->N_eps_fac;
    n->eps_fac = p->eps_fac;
    n->eps_fac_target = p->eps_fac;
    n->eps_fac_initial = p->eps_fac;
    n->eps_fac_power = 2;
    n->eps_fac_update_type = 0;
    n->N_a_eps_update = 100;
    n->acc_min = 0.6;
    n->acc_max = 0.7;
    n->eps_down_smooth = 5;
    n->eps_up_fac = 1;
    n->acc_flag_N_a = std::vector<bool>(n->N_a_eps_update);
    n->epsilon_N_a = std::vector<real_prec>(n->N_a_eps_update, n->eps_fac);
    n->acc_recent = std::vector<short>(n->N_a_eps_update);
    n->s_eps_total = 1;
    n->initial_guess = 0;
    n->mass_type = p->mass_type;
    n->massnum_init = 1;
    n->massnum_burn = 1;
    n->outnum = 10;
    n->outnum_ps = 10;
    n->start_at = 0;
    n->mass_factor = p->mass_factor;
    n->calc_h = p->calc_h;
    n->seed = 1;
    n->N_bin = p->N_bin > 0 ? p->N_bin : 200;
    n->inputmode = 0;
    n->dir = "./";
    n->mk = p->masskernel;
    n->N1 = n->N2 = n->N3 = (unsigned)p->N1;
    n->N = (ULONG)p->N1 * p->N1 * p->N1;
    n->Nhalf = (ULONG)p->N1 * p->N1 * (p->N1 / 2 + 1);
    n->readPS = true;
    n->slength = p->slength;
    n->iGibbs = 1;
    n->N_Gibbs = 1;
    n->total_steps_lim = ULONG_MAX;
    n->count_attempts = 0;
    n->xllc = p->xllc; n->yllc = p->yllc; n->zllc = p->zllc;
    n->xobs = p->xobs; n->yobs = p->yobs; n->zobs = p->zobs;
    n->planepar = p->planepar != 0;
    n->periodic = p->periodic != 0;
    n->L1 = n->L2 = n->L3 = p->L1;
    n->vol = p->L1 * p->L1 * p->L1;
    n->d1 = n->d2 = n->d3 = p->L1 / real_prec(p->N1);
    n->grad_psi_prior_factor = 1.;
    n->grad_psi_likeli_factor = 1.;
    n->grad_psi_prior_conjugate = false;
    n->grad_psi_likeli_conjugate = false;
    n->grad_psi_prior_times_i = false;
    n->grad_psi_likeli_times_i = false;
    n->div_dH_by_N = p->div_dH_by_N != 0;
    n->deltaQ_factor = p->deltaQ_factor;
    n->particle_kernel = 0;
    n->particle_kernel_h_rel = p->particle_kernel_h_rel;
    n->particle_kernel_h = p->particle_kernel_h_rel * (n->d1 + n->d2 + n->d3) / 3.;
    n->N_cells = SPH_kernel_3D_cells_count(0, n->particle_kernel_h, n->d1, n->d2, n->d3);
    SPH_kernel_3D_cells(0, n->particle_kernel_h, n->d1, n->d2, n->d3, n->kernel_cells_i, n->kernel_cells_j,
                        n->kernel_cells_k);
    n->correct_delta = p->correct_delta != 0;

    data->observational->sigma_fac = p->sigma_fac;
    data->observational->sigma_min = p->sigma_min;
    data->observational->delta_min = p->delta_min;

    data->cosmology->z = p->z;
    data->cosmology->ascale = 1. / (1. + p->z);
    INIT_COSMOLOGY(data->cosmology, n->codename);

    /* main.cc:117-119 */
    auto *in_r2c = new_real(h, n->N);
    auto *out_c2r = new_real(h, n->N);
    h->in_c2r = new fftw_array<complex_prec>(n->Nhalf);
    h->out_r2c = new fftw_array<complex_prec>(n->Nhalf);
    INIT_FFTW(data, *in_r2c, *out_c2r, *h->in_c2r, *h->out_r2c);

    h->rng = gsl_rng_alloc(gsl_rng_mt19937);
    gsl_rng_set(h->rng, 1);

    /* main.cc:150-154: POWER, SIGNAL, SIGNALX, window, NOISE_SF, NOBS, CORRF */
    auto *A = new_real(h, n->N), *B = new_real(h, n->N), *C = new_real(h, n->N), *D = new_real(h, n->N),
         *E = new_real(h, n->N), *F = new_real(h, n->N), *G = new_real(h, n->N);
    INIT_OBSERVATIONAL(data, *A, *B, *C, *D, *E, *F, *G);

    h->curses = new CURSES_STRUCT("BARCODE", "oracle");
    data->curses = h->curses;

    /* call_hamil.cc:38-44: gradpsi, mass_f, mass_r, posx, posy, posz */
    auto *HA = new_real(h, n->N), *HBf = new_real(h, n->N), *HBr = new_real(h, n->N), *HC = new_real(h, n->N),
         *HD = new_real(h, n->N), *HE = new_real(h, n->N);
    h->hd = new HAMIL_DATA(data, *HA, *HBf, *HBr, *HC, *HD, *HE, n->kernel_cells_i, n->kernel_cells_j,
                           n->kernel_cells_k, n->N_cells);
    std::cout.rdbuf(old);
  } catch (const std::exception &e) {
    g_err = e.what();
    delete h;
    return nullptr;
  }
  return h;
}

void ref_destroy(void *hv) {
  auto *h = static_cast<ref_handle *>(hv);
  if (!h) return;
  delete h->hd;
  if (h->data) {
    delete h->data->numerical->R2Cplan;
    delete h->data->numerical->C2Rplan;
  }
  delete h->curses;
  for (auto *a : h->real_arrays) delete a;
  delete h->in_c2r;
  delete h->out_r2c;
  if (h->rng) gsl_rng_free(h->rng);
  delete h->data;
  delete h;
}

/* pointers to the reference's own arrays (length N doubles) */
rp_t *ref_array(void *hv, const char *name) {
  auto *h = static_cast<ref_handle *>(hv);
  OBSERVATIONAL *o = h->data->observational;
  std::string s(name);
  if (s == "Power") return o->Power;
  if (s == "signal") return o->signal;
  if (s == "signalX" || s == "deltaX") return o->signalX;
  if (s == "window") return o->window;
  if (s == "noise") return o->noise_sf;
  if (s == "nobs") return o->nobs;
  if (s == "corrf") return o->corrf;
  if (s == "gradpsi") return h->hd->gradpsi;
  if (s == "mass_f") return h->hd->mass_f;
  if (s == "mass_r") return h->hd->mass_r;
  if (s == "posx") return h->hd->posx;
  if (s == "posy") return h->hd->posy;
  if (s == "posz") return h->hd->posz;
  return nullptr;
}

double ref_scalar(void *hv, const char *name) {
  auto *h = static_cast<ref_handle *>(hv);
  std::string s(name);
  HAMIL_NUMERICAL *n = h->hd->numerical;
  if (s == "D1") return h->hd->D1;
  if (s == "D2") return h->hd->D2;
  if (s == "OM") return h->hd->OM;
  if (s == "OL") return h->hd->OL;
  if (s == "ascale") return h->hd->ascale;
  if (s == "particle_kernel_h") return n->particle_kernel_h;
  if (s == "fgrow") return fgrow(h->hd->ascale, h->hd->OM, h->hd->OL, 1);
  if (s == "c_pecvel") return c_pecvel(h->hd->ascale, h->hd->OM, h->hd->OL, 1);
  if (s == "psi_prior") return n->psi_prior;
  if (s == "psi_likeli") return n->psi_likeli;
  if (s == "epsilon") return n->epsilon;
  if (s == "Neps") return (double)n->Neps;
  if (s == "dH") return n->dH;
  if (s == "dK") return n->dK;
  if (s == "dE") return n->dE;
  if (s == "dprior") return n->dprior;
  if (s == "dlikeli") return n->dlikeli;
  if (s == "H_kin_i") return n->H_kin_i;
  if (s == "H_kin_f") return n->H_kin_f;
  if (s == "psi_prior_i") return n->psi_prior_i;
  if (s == "psi_prior_f") return n->psi_prior_f;
  if (s == "psi_likeli_i") return n->psi_likeli_i;
  if (s == "psi_likeli_f") return n->psi_likeli_f;
  return 0.0 / 0.0;
}

/* calc_power.cc:31-108: P(k) table -> Power[N] (this container only: needs the CAMB file) */
int ref_readtab(void *hv, const char *fname) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  std::streambuf *old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  h->data->numerical->fnamePS = fname;
  readtab(h->data);
  std::cout.rdbuf(old);
  REF_CATCH
}

/* S1: HMC.cc:146-206 */
int ref_gradient_psi(void *hv, const rp_t *signal, rp_t *out) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  gradient_psi(h->hd, const_cast<rp_t *>(signal), h->data);
  std::memcpy(out, h->hd->gradpsi, h->hd->numerical->N * sizeof(rp_t));
  REF_CATCH
}

/* the likelihood half alone: HMC_models.cc:377-471 */
int ref_grad_log_like(void *hv, const rp_t *signal, rp_t *out) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  likelihood_grad_log_like(h->hd, const_cast<rp_t *>(signal), out);
  REF_CATCH
}

/* the prior half alone: hmc/prior/gaussian.cpp:15-18 */
int ref_grad_log_prior(void *hv, const rp_t *signal, rp_t *out) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  h->hd->grad_log_prior(h->hd, const_cast<rp_t *>(signal), out);
  REF_CATCH
}

/* S2: HMC.cc:124-143 (deltaX side effect readable through ref_array("deltaX")) */
int ref_psi(void *hv, const rp_t *signal, double *psi_prior, double *psi_like) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  psi(h->hd, const_cast<rp_t *>(signal), h->data);
  *psi_prior = h->hd->numerical->psi_prior;
  *psi_like = h->hd->numerical->psi_likeli;
  REF_CATCH
}

/* S3: HMC.cc:64-121 */
int ref_kinetic(void *hv, const rp_t *momenta, double *K) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  *K = kinetic_term(h->hd, const_cast<rp_t *>(momenta), h->data);
  REF_CATCH
}

/* S4: HMC.cc:251-369.  u_Neps, u_eps are the two uniforms the reference draws
 * first (Neps = floor(N_eps_fac*u)+1, eps = eps_fac*u). */
int ref_EoM(void *hv, const rp_t *si, const rp_t *pi, rp_t *sf, rp_t *pf, double u_Neps, double u_eps) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  shim_gsl_rng_force_uniform(h->rng, u_Neps);
  shim_gsl_rng_force_uniform(h->rng, u_eps);
  Hamiltonian_EoM(h->hd, const_cast<rp_t *>(si), const_cast<rp_t *>(pi), sf, pf, h->rng, h->data);
  REF_CATCH
}

/* HMC.cc:209-248; scalars readable through ref_scalar */
int ref_delta_hamiltonian(void *hv, const rp_t *si, const rp_t *pi, const rp_t *sf, const rp_t *pf,
                          double *dH) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  *dH = delta_Hamiltonian(h->hd, const_cast<rp_t *>(si), const_cast<rp_t *>(pi), const_cast<rp_t *>(sf),
                          const_cast<rp_t *>(pf), h->data);
  REF_CATCH
}

/* S5: HMC_momenta.cc:42-74 with a fresh mt19937 stream of the given seed */
int ref_draw_momenta(void *hv, unsigned long seed, rp_t *momenta) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  gsl_rng_set(h->rng, seed);
  draw_momenta(h->hd, h->rng, momenta, h->data);
  REF_CATCH
}

/* random.cpp:48-511 */
int ref_create_garfield(void *hv, unsigned long seed, const rp_t *power, rp_t *out) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  HAMIL_NUMERICAL *n = h->hd->numerical;
  gsl_rng_set(h->rng, seed);
  create_GARFIELD(n->N1, n->N2, n->N3, n->L1, n->L2, n->L3, out, power, h->rng);
  REF_CATCH
}

/* random.hpp:36-120, full grid: 2*N1^3 doubles (re, im interleaved) */
int ref_white_noise(int N1, unsigned long seed, rp_t *out) {
  REF_TRY
  gsl_rng *r = gsl_rng_alloc(gsl_rng_mt19937);
  gsl_rng_set(r, seed);
  std::vector<std::complex<real_prec> > g = resolution_independent_random_grid_FS<real_prec>((unsigned)N1, r, false);
  std::memcpy(out, g.data(), g.size() * sizeof(std::complex<real_prec>));
  gsl_rng_free(r);
  REF_CATCH
}

/* raw stream checks for the GSL shim */
int ref_rng_stream(unsigned long seed, int n_raw, unsigned long *raw, int n_gauss, double *gauss) {
  gsl_rng *r = gsl_rng_alloc(gsl_rng_mt19937);
  gsl_rng_set(r, seed);
  for (int i = 0; i < n_raw; ++i) raw[i] = gsl_rng_get(r);
  gsl_rng_set(r, seed);
  for (int i = 0; i < n_gauss; ++i) gauss[i] = gsl_ran_ugaussian(r);
  gsl_rng_free(r);
  return 0;
}

/* S6: HMC_mass.cc:315-368 */
int ref_hamiltonian_mass(void *hv) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  Hamiltonian_mass(h->hd, h->hd->x, h->data);
  REF_CATCH
}

/* forward model as the gradient / likelihood call it (HMC_models.cc:389-406) */
int ref_forward(void *hv, const rp_t *signal, rp_t *deltaX, rp_t *posx, rp_t *posy, rp_t *posz) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  HAMIL_DATA *hd = h->hd;
  HAMIL_NUMERICAL *n = hd->numerical;
  fftw_array<real_prec> in(n->N);
  for (ULONG i = 0; i < n->N; ++i) in[i] = n->deltaQ_factor * signal[i];
  real_prec kernel_scale = SPH_kernel_scale(hd);
  if (hd->rsd_model)
    Lag2Eul_rsd_zeldovich(in, hd->deltaX, hd->posx, hd->posy, hd->posz, n->N1, n->N2, n->N3, n->L1, n->L2, n->L3,
                          n->d1, n->d2, n->d3, n->min1, n->min2, n->min3, hd->D1, hd->ascale, hd->OM, hd->OL, n->mk, 1,
                          true, nullptr, kernel_scale, n->xobs, n->yobs, n->zobs, n->planepar, n->periodic,
                          n->R2Cplan, n->C2Rplan);
  else
    Lag2Eul(in, hd->deltaX, hd->posx, hd->posy, hd->posz, n->N1, n->N2, n->N3, n->L1, n->L2, n->L3, n->d1, n->d2,
            n->d3, n->min1, n->min2, n->min3, hd->D1, hd->D2, hd->ascale, hd->OM, hd->OL, hd->sfmodel, n->mk, n->kth, 1,
            true, nullptr, "", kernel_scale, n->R2Cplan, n->C2Rplan);
  std::memcpy(deltaX, hd->deltaX, n->N * sizeof(rp_t));
  if (posx) std::memcpy(posx, hd->posx, n->N * sizeof(rp_t));
  if (posy) std::memcpy(posy, hd->posy, n->N * sizeof(rp_t));
  if (posz) std::memcpy(posz, hd->posz, n->N * sizeof(rp_t));
  REF_CATCH
}

/* kernelcomp as barcoderunner calls it before sampling (barcoderunner.cc:371-374): writes
 * <dir>auxkernelr<int(slength)>.dat, which convcomp re-reads on every call (convolution.cpp:354-359) */
int ref_kernelcomp(void *hv) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  NUMERICAL *n = h->data->numerical;
  /* the HMC call sites read the kernel with dir = "" (HMC_models.cc:405), i.e. from the working directory
   * as "auxkernelr<r>.dat"; write it there (with dir = "./" add_extension_if_missing sees the dot of "./"
   * and drops the ".dat", IOfunctionsGen.cc:185-191) */
  const std::string saved = n->dir;
  n->dir = "";
  kernelcomp(n->L1, n->L2, n->L3, n->N1, n->N2, n->N3, n->slength, 1, h->data);
  n->dir = saved;
  REF_CATCH
}

/* mass assignment alone on given positions (massFunctions.cc:49-495), no overdens */
int ref_density(void *hv, const rp_t *x, const rp_t *y, const rp_t *z, rp_t *rho) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  HAMIL_NUMERICAL *n = h->hd->numerical;
  fftw_array<real_prec> ones(n->N);
  for (ULONG i = 0; i < n->N; ++i) ones[i] = 1.;
  switch (n->mk) {
    case 0:
      getDensity_NGP(n->N1, n->N2, n->N3, n->L1, n->L2, n->L3, n->d1, n->d2, n->d3, n->min1, n->min2, n->min3, x, y,
                     z, ones, n->N, rho);
      break;
    case 1:
      getDensity_CIC(n->N1, n->N2, n->N3, n->L1, n->L2, n->L3, n->d1, n->d2, n->d3, n->min1, n->min2, n->min3, x, y,
                     z, ones, n->N, rho, true);
      break;
    case 2:
      getDensity_TSC(n->N1, n->N2, n->N3, n->L1, n->L2, n->L3, n->d1, n->d2, n->d3, n->min1, n->min2, n->min3, x, y,
                     z, ones, n->N, rho);
      break;
    case 3:
      getDensity_SPH(n->N1, n->N2, n->N3, n->L1, n->L2, n->L3, n->d1, n->d2, n->d3, n->min1, n->min2, n->min3, x, y,
                     z, ones, n->N, rho, true, SPH_kernel_scale(h->hd));
      break;
    default:
      throw std::runtime_error("ref_density: bad masskernel");
  }
  REF_CATCH
}

/* residual alone: hd->partial_f_delta_x_log_like (gaussian_independent.cpp:24-42 etc.) */
int ref_partial_f(void *hv, const rp_t *deltaX, rp_t *out) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  h->hd->partial_f_delta_x_log_like(h->hd, const_cast<rp_t *>(deltaX), out);
  REF_CATCH
}

/* A5: HMC_help.cc:16-64 */
int ref_convolve_inv_corr(void *hv, const rp_t *signal, const rp_t *corr, rp_t *out) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  convolveInvCorrFuncWithSignal(h->hd, const_cast<rp_t *>(signal), out, corr);
  REF_CATCH
}

/* fftwrapper.cc:26-84 through the shim: r2c then c2r */
int ref_fft_r2c(int N1, const rp_t *in, rp_t *out_complex) {
  REF_TRY
  ULONG N = (ULONG)N1 * N1 * N1;
  fftw_array<real_prec> tmp(N);
  std::memcpy(tmp.data, in, N * sizeof(rp_t));
  fftR2C(N1, N1, N1, tmp, reinterpret_cast<complex_prec *>(out_complex));
  REF_CATCH
}
int ref_fft_c2r(int N1, const rp_t *in_complex, rp_t *out) {
  REF_TRY
  ULONG Nh = (ULONG)N1 * N1 * (N1 / 2 + 1);
  fftw_array<complex_prec> tmp(Nh);
  std::memcpy(tmp.data, in_complex, Nh * sizeof(complex_prec));
  fftC2R(N1, N1, N1, tmp, out);
  REF_CATCH
}

/* spectral / finite-difference gradient components (gradient.cpp:22-153) */
int ref_gradfft(int N1, double L1, const rp_t *in, rp_t *out, unsigned dim) {
  REF_TRY
  gradfft(N1, N1, N1, L1, L1, L1, const_cast<rp_t *>(in), out, dim);
  REF_CATCH
}
int ref_gradfindif(int N1, double L1, const rp_t *in, rp_t *out, unsigned dim) {
  REF_TRY
  gradfindif(N1, L1, in, out, dim);
  REF_CATCH
}

/* measure_spectrum, field_statistics.cpp:20-90 (per-sample diagnostics) */
int ref_measure_spectrum(int N1, double L1, const rp_t *signal, unsigned long N_bin, rp_t *kmode, rp_t *power) {
  REF_TRY
  measure_spectrum(N1, N1, N1, L1, L1, L1, const_cast<rp_t *>(signal), kmode, power, N_bin);
  REF_CATCH
}

/* CPU baseline timing: seconds per gradient_psi call, 1 warm-up + reps timed (omp_get_wtime) */
int ref_time_gradient_psi(void *hv, const rp_t *signal, int reps, double *seconds_per_call) {
  auto *h = static_cast<ref_handle *>(hv);
  REF_TRY
  gradient_psi(h->hd, const_cast<rp_t *>(signal), h->data);
  const double t0 = omp_get_wtime();
  for (int r = 0; r < reps; ++r) gradient_psi(h->hd, const_cast<rp_t *>(signal), h->data);
  *seconds_per_call = (omp_get_wtime() - t0) / reps;
  REF_CATCH
}

int ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }
const char *ref_fft_backend(void) { return shim_fftw_backend(); }

}  // extern "C"
