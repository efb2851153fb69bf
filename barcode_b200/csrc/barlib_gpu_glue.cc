// barcode_b200/csrc/barlib_gpu_glue.cc
//
// Reference-side binding of the GPU path: this one file is compiled INTO the
// Barcode host program in place of barlib/src/HMC.cc and barlib/src/HMC_momenta.cc
// (see INTEGRATION.md).  It defines the same free functions with the same
// signatures -- HamiltonianMC (HMC.h:15), Hamiltonian_EoM, gradient_psi, psi,
// kinetic_term, delta_Hamiltonian (HMC.cc:64-369), draw_momenta (HMC_momenta.h)
// -- so main.cc, barcoderunner.cc, sample_maker.cc and call_hamil.cc, input.par
// and every output file stay exactly as they are.  The arithmetic on N-cell
// arrays happens behind include/barcode_gpu.h; what stays here is what the
// reference keeps on the host thread anyway: the GSL random stream (Neps,
// epsilon, white noise, the Metropolis uniform), the adaptive step-size tables
// (hmc/leapfrog/time_step.cpp, unchanged), the performance log, the ncurses
// status table and the auxmass_{f,r}.dat files.
//
// Errors: the C ABI returns codes; they are rethrown here as
// std::runtime_error, the reference's only error channel (main.cc:195-197).
//
// Build: needs the reference's headers (struct_main.h, struct_hamil.h, ...),
// so it is compiled by oracle/Makefile in the build container, never on the
// GPU box.

#include "struct_main.h"
#include "struct_hamil.h"

#include <cassert>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <complex>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "fftw_array.h"
#include "IOfunctionsGen.h"
#include "IOfunctions.h"
#include "HMC.h"
#include "HMC_mass.h"
#include "HMC_momenta.h"
#include "convenience.h"
#include "random.hpp"
#include "curses_funcs.h"
#include "hmc/leapfrog/time_step.hpp"

#include "barcode_gpu.h"

#ifdef MASKING
#error "the GPU path does not implement the MASKING build option (HMC.cc:33-37, 70-80, 299-307)"
#endif
#ifndef DOUBLE_PREC
#error "the GPU glue is written for the reference's default DOUBLE_PREC build"
#endif

namespace {

void check(int rc, const char *where) {
  if (rc != 0) throw std::runtime_error(std::string(where) + ": " + bgpu_last_error());
}

// BARCODE_GPU_TIMING=1: wall time spent inside every C-ABI call site, printed to stderr when the process ends
struct CallTimes {
  struct Row { const char *name; unsigned long calls; double ms; };
  std::vector<Row> rows;
  bool on = std::getenv("BARCODE_GPU_TIMING") != nullptr;
  void add(const char *name, double ms) {
    for (Row &r : rows)
      if (r.name == name) { r.calls++; r.ms += ms; return; }
    rows.push_back(Row{name, 1, ms});
  }
  ~CallTimes() {
    if (!on) return;
    for (const Row &r : rows) std::fprintf(stderr, "[barcode_gpu] %-28s %8lu calls %12.3f ms\n", r.name, r.calls, r.ms);
  }
};
CallTimes g_times;
struct CallTimer {
  const char *name;
  std::chrono::steady_clock::time_point t0;
  explicit CallTimer(const char *n) : name(n), t0(std::chrono::steady_clock::now()) {}
  ~CallTimer() {
    if (g_times.on)
      g_times.add(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};
// BGPU_CALL(bgpu_x(...), "bgpu_x") with the call timed
#define BGPU_CALL(expr, where) \
  do {                         \
    CallTimer timer__(where);  \
    check((expr), where);      \
  } while (0)

// One device-resident chain per process, rebuilt only if the run parameters change.  The
// reference allocates HAMIL_DATA afresh for every sample (call_hamil.cc:38-44); the device
// state is a pure function of the host arrays, so it can outlive a sample.
struct Session {
  bgpu_handle *h = nullptr;
  bgpu_params p{};
  bool valid = false;
  ~Session() {
    if (h) bgpu_destroy(h);
  }
};
Session g_session;

bool same_params(const bgpu_params &a, const bgpu_params &b) {
  return a.N1 == b.N1 && a.L1 == b.L1 && a.xllc == b.xllc && a.yllc == b.yllc && a.zllc == b.zllc &&
         a.planepar == b.planepar && a.periodic == b.periodic && a.masskernel == b.masskernel &&
         a.likelihood == b.likelihood && a.sfmodel == b.sfmodel && a.rsd_model == b.rsd_model &&
         a.calc_h == b.calc_h && a.mass_type == b.mass_type && a.D1 == b.D1 && a.D2 == b.D2 &&
         a.slength == b.slength && a.particle_kernel_h_rel == b.particle_kernel_h_rel && a.ascale == b.ascale &&
         a.OM == b.OM && a.OL == b.OL && a.rho_c == b.rho_c && a.biasP == b.biasP && a.biasE == b.biasE &&
         a.deltaQ_factor == b.deltaQ_factor && a.correct_delta == b.correct_delta &&
         a.mass_factor == b.mass_factor && a.delta_min == b.delta_min && a.N_bin == b.N_bin;
}

bgpu_params params_from(struct HAMIL_DATA *hd, struct DATA *data) {
  const HAMIL_NUMERICAL *n = hd->numerical;
  if (n->grad_psi_prior_factor != 1. || n->grad_psi_likeli_factor != 1. || n->grad_psi_prior_conjugate ||
      n->grad_psi_likeli_conjugate || n->grad_psi_prior_times_i || n->grad_psi_likeli_times_i)
    throw std::runtime_error("GPU path: the grad_psi_* test knobs (HMC.cc:170-201) must be left at their defaults");
  bgpu_params p;
  bgpu_default_params(&p);
  p.N1 = (int)n->N1; p.N2 = (int)n->N2; p.N3 = (int)n->N3;
  p.L1 = n->L1; p.L2 = n->L2; p.L3 = n->L3;
  p.xllc = n->min1; p.yllc = n->min2; p.zllc = n->min3;
  p.xobs = n->xobs; p.yobs = n->yobs; p.zobs = n->zobs;
  p.planepar = n->planepar; p.periodic = n->periodic;
  p.masskernel = n->mk;
  p.likelihood = data->numerical->likelihood;
  p.sfmodel = hd->sfmodel;
  p.rsd_model = hd->rsd_model;
  p.calc_h = n->calc_h;
  p.mass_type = n->mass_type;
  p.D1 = hd->D1; p.D2 = hd->D2; p.ascale = hd->ascale; p.OM = hd->OM; p.OL = hd->OL;
  p.slength = n->kth;
  p.particle_kernel_h_rel = n->particle_kernel_h / ((n->d1 + n->d2 + n->d3) / 3.);
  p.rho_c = hd->rho_c; p.biasP = hd->biasP; p.biasE = hd->biasE;
  p.deltaQ_factor = n->deltaQ_factor;
  p.correct_delta = n->correct_delta;
  p.mass_factor = n->mass_factor;
  p.div_dH_by_N = n->div_dH_by_N;
  p.delta_min = hd->delta_min;
  p.N_bin = (int)n->N_bin;
  const char *dev = std::getenv("BARCODE_GPU_DEVICE");
  p.device = dev ? std::atoi(dev) : 0;
  return p;
}

bgpu_handle *session(struct HAMIL_DATA *hd, struct DATA *data) {
  bgpu_params p = params_from(hd, data);
  if (!g_session.valid || !same_params(p, g_session.p)) {
    if (g_session.h) bgpu_destroy(g_session.h);
    g_session.h = nullptr;
    g_session.valid = false;
    BGPU_CALL(bgpu_create(&p, &g_session.h), "bgpu_create");
    g_session.p = p;
    g_session.valid = true;
  }
  return g_session.h;
}

// HMC.cc:40-60
void write_to_performance_log(struct DATA *data, struct HAMIL_DATA *hd) {
  const HAMIL_NUMERICAL *n = hd->numerical;
  std::ofstream &plog = data->numerical->performance_log;
  assert(plog.is_open());
  const char tab = '\t';
  plog << n->accepted << tab << n->epsilon << tab << n->Neps << tab << n->dH << tab << n->dK << tab << n->dE << tab
       << n->dprior << tab << n->dlikeli << tab << n->psi_prior_i << tab << n->psi_prior_f << tab << n->psi_likeli_i
       << tab << n->psi_likeli_f << tab << n->H_kin_i << tab << n->H_kin_f << std::endl;
}

}  // namespace

// ---------------------------------------------------------------------------
// S3  kinetic_term, HMC.cc:64-121
// ---------------------------------------------------------------------------
real_prec kinetic_term(struct HAMIL_DATA *hd, real_prec *momenta, struct DATA *data) {
  double K = 0.;
  BGPU_CALL(bgpu_kinetic(session(hd, data), momenta, &K), "bgpu_kinetic");
  wprintw(data->curses->table, "%5.0e ", K);
  return K;
}

// ---------------------------------------------------------------------------
// S2  psi, HMC.cc:124-143 (hd->deltaX is refreshed as the reference's log_like does)
// ---------------------------------------------------------------------------
real_prec psi(struct HAMIL_DATA *hd, real_prec *signal, struct DATA *data) {
  HAMIL_NUMERICAL *n = hd->numerical;
  double prior = 0., like = 0.;
  BGPU_CALL(bgpu_psi(session(hd, data), signal, &prior, &like, hd->deltaX), "bgpu_psi");
  wprintw(data->curses->table, "%5.0e ", prior);
  wprintw(data->curses->table, "%5.0e ", like);
  n->psi_prior = prior;
  n->psi_likeli = like;
  return prior + like;
}

// ---------------------------------------------------------------------------
// S1  gradient_psi, HMC.cc:146-206
// ---------------------------------------------------------------------------
void gradient_psi(struct HAMIL_DATA *hd, real_prec *signal, struct DATA *data) {
  BGPU_CALL(bgpu_gradient_psi(session(hd, data), signal, hd->gradpsi), "bgpu_gradient_psi");
}

// ---------------------------------------------------------------------------
// delta_Hamiltonian, HMC.cc:209-248
// ---------------------------------------------------------------------------
real_prec delta_Hamiltonian(struct HAMIL_DATA *hd, real_prec *signali, real_prec *momentai, real_prec *signalf,
                            real_prec *momentaf, struct DATA *data) {
  HAMIL_NUMERICAL *n = hd->numerical;
  const real_prec Hkini = kinetic_term(hd, momentai, data);
  const real_prec Hpsii = psi(hd, signali, data);
  n->psi_prior_i = n->psi_prior;
  n->psi_likeli_i = n->psi_likeli;
  n->H_kin_i = Hkini;
  const real_prec Hami = Hkini + Hpsii;

  const real_prec Hkinf = kinetic_term(hd, momentaf, data);
  const real_prec Hpsif = psi(hd, signalf, data);
  n->dprior = n->psi_prior - n->psi_prior_i;
  n->dlikeli = n->psi_likeli - n->psi_likeli_i;
  const real_prec Hamf = Hkinf + Hpsif;

  real_prec dHam = Hamf - Hami;
  if (n->div_dH_by_N) dHam /= static_cast<real_prec>(n->N);
  n->dH = dHam;
  n->dK = Hkinf - Hkini;
  n->dE = Hpsif - Hpsii;
  n->psi_prior_f = n->psi_prior;
  n->psi_likeli_f = n->psi_likeli;
  n->H_kin_f = Hkinf;
  return dHam;
}

// ---------------------------------------------------------------------------
// S4  Hamiltonian_EoM, HMC.cc:251-369: the two RNG draws stay on the host stream, the
// whole Neps-step trajectory runs on the device (momentum run-away guard included).
// ---------------------------------------------------------------------------
void Hamiltonian_EoM(struct HAMIL_DATA *hd, real_prec *signali, real_prec *momentai, real_prec *signalf,
                     real_prec *momentaf, gsl_rng *seed, struct DATA *data) {
  HAMIL_NUMERICAL *n = hd->numerical;
  n->Neps = static_cast<ULONG>(n->N_eps_fac * (gsl_rng_uniform(seed))) + 1;
  n->epsilon = static_cast<real_prec>(n->eps_fac * gsl_rng_uniform(seed));
  if (n->epsilon > 2.) n->epsilon = 2.;

  wprintw(data->curses->table, "%5.0e ", n->epsilon);
  wprintw(data->curses->table, "%4lu ", n->Neps);
  wrefresh(data->curses->table);
  wprintw(data->curses->status, "\nLeap-frogging on the GPU ... %lu steps", n->Neps);
  wrefresh(data->curses->status);

  BGPU_CALL(bgpu_leapfrog(session(hd, data), signali, momentai, n->Neps, n->epsilon, signalf, momentaf),
        "bgpu_leapfrog");
  if (std::abs(momentaf[0]) > 1e50) {
    wprintw(data->curses->message, "\nLeap-frogging ... stopped, momentum too high (momenta[0] = %e)", momentaf[0]);
    wrefresh(data->curses->message);
  }
  data->numerical->count_attempts++;
}

// ---------------------------------------------------------------------------
// S5  draw_momenta, HMC_momenta.cc:42-92: GSL stream on the host in the reference's order
// (2N ugaussians in shell order, then N gaussians if mass_rs), colouring + C2R on the device
// ---------------------------------------------------------------------------
//
// BARCODE_GPU_DEVICE_RNG=1 (opt-in, NOT seed-compatible with the CPU code): the draw itself happens on the device
// (bgpu_draw_momenta_device: Philox4x32-10, same distribution).  The host stream then only gives one 32-bit word
// per candidate as the draw's key, so it stays in step for Neps, epsilon and the Metropolis uniform.
static bool device_rng_enabled() {
  static const bool on = [] {
    const char *e = std::getenv("BARCODE_GPU_DEVICE_RNG");
    return e && e[0] == '1';
  }();
  return on;
}
static uint64_t g_draw_index = 0;  // device generator: one counter per momentum draw of the process

void draw_momenta(struct HAMIL_DATA *hd, gsl_rng *seed, real_prec *momenta, struct DATA *data) {
  HAMIL_NUMERICAL *n = hd->numerical;
  if (device_rng_enabled()) {
    const uint64_t key = gsl_rng_get(seed);
    BGPU_CALL(bgpu_draw_momenta_device(session(hd, data), key, g_draw_index++, momenta), "bgpu_draw_momenta_device");
    return;
  }
  std::vector<std::complex<real_prec> > white;
  if (n->mass_fs) white = resolution_independent_random_grid_FS<real_prec>(n->N1, seed, false);
  std::vector<real_prec> gauss;
  if (n->mass_rs) {
    gauss.resize(n->N);
    for (ULONG i = 0; i < n->N; ++i) gauss[i] = static_cast<real_prec>(GR_NUM(seed, 1., 0));
  }
  BGPU_CALL(bgpu_color_momenta(session(hd, data), n->mass_fs ? reinterpret_cast<const double *>(white.data()) : nullptr,
                           n->mass_rs ? gauss.data() : nullptr, momenta),
        "bgpu_color_momenta");
}

// ---------------------------------------------------------------------------
// A1  HamiltonianMC, HMC.cc:372-548: candidate loop, Metropolis step, logs
// ---------------------------------------------------------------------------
void HamiltonianMC(struct HAMIL_DATA *hd, gsl_rng *seed, struct DATA *data) {
  CallTimer whole__("HamiltonianMC (whole)");
  HAMIL_NUMERICAL *n = hd->numerical;
  NUMERICAL *dn = data->numerical;
  bgpu_handle *h = session(hd, data);

  // static inputs of this sample (main.cc:150-154; the mock data are made before the loop)
  BGPU_CALL(bgpu_set_static(h, hd->signal_PS, hd->nobs, hd->noise, hd->window), "bgpu_set_static");

  // Hamiltonian masses: recomputed every massnum samples, otherwise re-read from disk (HMC.cc:386-423)
  const ULONG massnum = (n->iGibbs > n->massnum_burn) ? n->massnum_burn : n->massnum_init;
  const std::string name_r = dn->dir + std::string("auxmass_r"), name_f = dn->dir + std::string("auxmass_f");
  if (0 == n->iGibbs % massnum || n->iGibbs == 1) {
    if (n->mass_type == 2 || n->mass_type == 3) {
      // the likelihood-force masses (HMC_mass.cc:39-160) need likelihood_grad_log_like + measure_spectrum: on the
      // device, with the forcespec.dat dump the reference writes (likeli_force_power, :48-50)
      std::vector<real_prec> kmode(n->N_bin), fpower(n->N_bin);
      BGPU_CALL(bgpu_likeli_force_power(h, hd->x, kmode.data(), fpower.data()), "bgpu_likeli_force_power");
      dump_measured_spec(kmode.data(), fpower.data(), dn->dir + std::string("forcespec.dat"), n->N_bin);
      BGPU_CALL(bgpu_hamiltonian_mass_x(h, hd->x, hd->mass_f, nullptr), "bgpu_hamiltonian_mass_x");
    } else {
      Hamiltonian_mass(hd, hd->x, data);   // host, unchanged (types 0/1/4 are one pass over Power)
    }
    if (n->mass_rs) {
      if (contains_nan(hd->mass_r, n->N)) throw std::runtime_error("auxmass_r contains a NaN! aborting.");
      write_array(name_r, hd->mass_r, n->N1, n->N2, n->N3);
    }
    if (n->mass_fs) write_array(name_f, hd->mass_f, n->N1, n->N2, n->N3);
  } else {
    if (n->mass_rs) read_array(name_r, hd->mass_r, n->N1, n->N2, n->N3);
    if (n->mass_fs) read_array(name_f, hd->mass_f, n->N1, n->N2, n->N3);
  }
  BGPU_CALL(bgpu_set_mass(h, n->mass_fs ? hd->mass_f : nullptr, n->mass_rs ? hd->mass_r : nullptr), "bgpu_set_mass");

  wprintw(data->curses->status, "starting Hamiltonian sampling (GPU path)");
  wrefresh(data->curses->status);

  // With the device generator (BARCODE_GPU_DEVICE_RNG=1) nothing of a candidate needs the host: the signal is
  // uploaded once per sample and every candidate -- momenta, trajectory, the four energies -- stays on the device
  // (bgpu_candidate); only the accepted field comes back.  The host stream is consumed exactly as in the separate
  // calls: one word as the draw's key, then Neps and epsilon, then the Metropolis uniform.
  static const bool fused_off = [] {   // BARCODE_GPU_FUSED=0: device generator, but the separate host-array calls
    const char *e = std::getenv("BARCODE_GPU_FUSED");
    return e && e[0] == '0';
  }();
  const bool fused = device_rng_enabled() && !fused_off;
  fftw_array<real_prec> momentai(fused ? 1 : n->N), momentaf(fused ? 1 : n->N), signali(fused ? 1 : n->N),
      signalf(fused ? 1 : n->N);
  if (fused) BGPU_CALL(bgpu_set_signal(h, hd->x), "bgpu_set_signal");
  bool accepted = false;
  for (ULONG iter = 1; iter <= n->itmax && !accepted; ++iter) {
    wprintw(data->curses->table, "%6lu ", n->iGibbs);
    wprintw(data->curses->table, "%4lu ", iter);
    wrefresh(data->curses->table);

    real_prec dH;
    if (fused) {
      const uint64_t key = gsl_rng_get(seed);                                  // draw_momenta's key
      update_eps_fac(hd, data);                                                // :453
      n->Neps = static_cast<ULONG>(n->N_eps_fac * (gsl_rng_uniform(seed))) + 1;  // :260-263
      n->epsilon = static_cast<real_prec>(n->eps_fac * gsl_rng_uniform(seed));
      if (n->epsilon > 2.) n->epsilon = 2.;
      wprintw(data->curses->table, "%5.0e ", n->epsilon);
      wprintw(data->curses->table, "%4lu ", n->Neps);
      wrefresh(data->curses->table);
      double E[6], pf0 = 0.;
      BGPU_CALL(bgpu_candidate(h, key, g_draw_index++, n->Neps, n->epsilon, E, &pf0), "bgpu_candidate");
      if (std::abs(pf0) > 1e50) {
        wprintw(data->curses->message, "\nLeap-frogging ... stopped, momentum too high (momenta[0] = %e)", pf0);
        wrefresh(data->curses->message);
      }
      data->numerical->count_attempts++;
      // delta_Hamiltonian, :209-248
      n->H_kin_i = E[0]; n->psi_prior_i = E[1]; n->psi_likeli_i = E[2];
      n->H_kin_f = E[3]; n->psi_prior_f = E[4]; n->psi_likeli_f = E[5];
      n->psi_prior = E[4]; n->psi_likeli = E[5];
      n->dprior = E[4] - E[1];
      n->dlikeli = E[5] - E[2];
      const real_prec Hami = E[0] + (E[1] + E[2]), Hamf = E[3] + (E[4] + E[5]);
      dH = Hamf - Hami;
      if (n->div_dH_by_N) dH /= static_cast<real_prec>(n->N);
      n->dH = dH;
      n->dK = E[3] - E[0];
      n->dE = (E[4] + E[5]) - (E[1] + E[2]);
    } else {
      copyArray(hd->x, signali, n->N);                                           // :445
      draw_momenta(hd, seed, momentai, data);                                    // :449
      update_eps_fac(hd, data);                                                  // :453
      Hamiltonian_EoM(hd, signali, momentai, signalf, momentaf, seed, data);     // :455
      dH = delta_Hamiltonian(hd, signali, momentai, signalf, momentaf, data);
    }

    // acceptance probability, :462-467
    real_prec p_acceptance = 1.;
    if (!(dH < 0.0) && std::exp(-dH) < 1.0) p_acceptance = std::exp(-dH);
    wprintw(data->curses->table, "%6.0e ", dH);
    wprintw(data->curses->table, "%5.0e ", p_acceptance);
    wrefresh(data->curses->table);

    // the uniform is drawn only when it is needed, :477-486
    if (p_acceptance >= 1.0) {
      accepted = true;
    } else {
      const auto u = static_cast<real_prec>(gsl_rng_uniform(seed));
      accepted = (u < p_acceptance);
    }
    wprintw(data->curses->table, accepted ? "y " : "n ");
    wrefresh(data->curses->table);

    if (accepted) {
      if (fused) BGPU_CALL(bgpu_accept(h, hd->x, hd->deltaX), "bgpu_accept");
      else copyArray(signalf, hd->x, n->N);
    } else {
      n->rejections++;
    }
    n->accepted = accepted;

    write_to_performance_log(data, hd);
    update_epsilon_acc_rate_tables(data, hd);

    const ULONG total_steps = n->iGibbs + n->rejections + dn->rejections + (accepted ? 1 : 0);

    // recent acceptance rate, :517-531
    const ULONG ix_acc = (dn->count_attempts - 1) % dn->N_a_eps_update;
    dn->acc_recent[ix_acc] = accepted ? 1 : 0;
    real_prec acc_N_a = 0;
    for (unsigned i = 0; i < dn->N_a_eps_update; ++i) acc_N_a += static_cast<real_prec>(dn->acc_recent[i]);
    acc_N_a /= static_cast<real_prec>(dn->N_a_eps_update);
    wprintw(data->curses->table, "%4.2f", acc_N_a);
    wrefresh(data->curses->table);

    if (total_steps >= n->total_steps_lim) throw std::runtime_error("ABORTING: total steps exceeds total_steps_lim");
    wprintw(data->curses->table, "\n");
    wrefresh(data->curses->table);
  }
  n->INV_SUCCESS = accepted ? 1 : 0;
}
