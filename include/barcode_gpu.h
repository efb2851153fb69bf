/*
 * include/barcode_gpu.h -- C ABI of the B200-native HMC gradient-and-leapfrog
 * path for Barcode (libbarcode_b200.so).
 *
 * The reference (egpbos/barcode, /root/reference) has no plugin or FFI layer;
 * the seam this library replaces is the set of free functions in
 * barlib/src/HMC.cc, HMC_momenta.cc and HMC_mass.cc that the unchanged host
 * driver (main.cc -> barcoderunner -> sample_maker -> call_hamil ->
 * HamiltonianMC) calls once per candidate / leapfrog step.  Every entry point
 * below cites the reference function it stands in for; INTEGRATION.md shows
 * the glue a maintainer adds on the reference side
 * (barcode_b200/csrc/barlib_gpu_glue.cc).
 *
 * Conventions
 *  - plain C: pointers, sizes, PODs; no exceptions cross the boundary.  Every
 *    call returns 0 on success, non-zero on failure; bgpu_last_error() returns
 *    the message (the glue rethrows it as std::runtime_error, the reference's
 *    only error channel, main.cc:195-197).
 *  - arrays are caller-owned, double[N1*N2*N3], row-major with z fastest
 *    (idx = k + N3*(j + N2*i), disp_part.cc:60), exactly the reference's
 *    fftw_malloc'ed arrays.  Host-pointer calls copy in and out; the *_dev
 *    calls take device pointers and run on the handle's stream.
 *  - one host thread per handle (the reference is single-threaded and not
 *    re-entrant either); use one handle per GPU / chain.
 *  - there is no CPU fallback: creation fails loudly without a CUDA device.
 */
#ifndef BARCODE_GPU_H
#define BARCODE_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGPU_ABI_VERSION 2

/* calc_h value for the exact mass-assignment adjoint (new; the reference only
 * has exact adjoints for its SPH kernel, HMC_models.cc:200-372) */
#define BGPU_CALC_H_EXACT 4

typedef struct bgpu_handle bgpu_handle;

/* The fields of DATA / HAMIL_DATA the path reads (struct_main.h:23-258,
 * struct_hamil.h:51-212); names follow the reference / input.par. */
typedef struct bgpu_params {
  int N1, N2, N3;            /* cells per axis; cubic, power of two in [8, 1024] (init_par.cc:116-122) */
  double L1, L2, L3;         /* box size [Mpc/h] */
  double xllc, yllc, zllc;   /* origin -> min1..3 */
  double xobs, yobs, zobs;   /* observer (only plane-parallel RSD is supported, as in rsd.cc:60-62) */
  int planepar, periodic;
  int masskernel;            /* mk: 0 NGP, 1 CIC, 2 TSC, 3 SPH spline */
  int likelihood;            /* 0 Poisson, 1 Gaussian, 2 log-normal, 3 Gaussian random field (init_par.cc:534-559) */
  int sfmodel;               /* 1 Zel'dovich; else Lag2Eul_non_zeldovich (2LPT + spherical collapse, ALPT);
                              * with rsd_model the reference runs Zel'dovich for any value */
  int rsd_model;
  int calc_h;                /* 0, 1, 2 (SPH adjoint) and 3 (its Fourier / TSC variant, HMC_models_testing.cpp:54-188; both
                              * need masskernel 3) as the reference; BGPU_CALC_H_EXACT = exact adjoint of NGP / CIC / TSC
                              * under the Zel'dovich or the 2LPT/ALPT model */
  int mass_type;             /* 0 ones (R), 1 1/P (FS), 4 P (FS); 2 / 3: 1/P + likelihood-force spectrum / its mean (FS,
                              * HMC_mass.cc:39-160; bgpu_hamiltonian_mass_x) */
  double D1, D2, ascale, OM, OL;
  double particle_kernel_h_rel; /* SPH scale length in cells (input.par particle_kernel_h_rel); masskernel 3 */
  double slength;            /* ALPT smoothing radius [Mpc/h] (input.par slength -> n->kth, struct_hamil.h:259); sfmodel != 1 */
  double rho_c, biasP, biasE;
  double deltaQ_factor;
  int correct_delta;
  double mass_factor;
  int div_dH_by_N;
  int device;                /* CUDA device ordinal */
  double delta_min;          /* log-normal likelihood: density floor (input.par delta_min, default -0.999) */
  int N_bin;                 /* bins of measure_spectrum (input.par N_bin, default 200); mass types 2 / 3 */
  int reserved[5];
} bgpu_params;

/* the reference's defaults (data/input.par, init_par.cc:574-578) */
void bgpu_default_params(bgpu_params *p);
int bgpu_abi_version(void);
const char *bgpu_last_error(void);

/* call_hamil.cc:38-44 (per-sample HAMIL_DATA) + INIT_FFTW (init_par.cc:418-428) */
int bgpu_create(const bgpu_params *p, bgpu_handle **out);
void bgpu_destroy(bgpu_handle *h);

/* One chain across G GPUs of a box, x-slab decomposition: rank r owns planes
 * i in [r*N1/G, (r+1)*N1/G) of EVERY array argument below, i.e. the contiguous sub-array
 * double[N1/G][N2][N3] of the reference's layout (and of its output files).  One process per GPU;
 * rank 0 calls bgpu_nccl_unique_id and the host program hands the 128 bytes to the other ranks
 * (MPI_Bcast, a file, torch.distributed ...).  The distributed FFT transposes with NCCL
 * all-to-all, the mass-assignment halo goes to the two x neighbours, scalars are all-reduced; every
 * rank must make the same sequence of calls.  Supported: N1 in {128, 256, 512, 1024}; NGP / CIC / TSC / SPH;
 * Zel'dovich and 2LPT/ALPT forward models, RSD; all four likelihoods; calc_h 0, 1, 2 (SPH adjoint) and
 * BGPU_CALC_H_EXACT (Zel'dovich); mass types 0 - 4; bgpu_measure_spectrum; bgpu_draw_momenta_device;
 * bgpu_color_momenta (every rank passes the FULL white-noise grid, real_gauss and momenta are slabs).
 * Not on slabs: calc_h 3 and the exact adjoint of the 2LPT/ALPT model. */
int bgpu_nccl_unique_id(void *out128);
int bgpu_slab_create(const bgpu_params *p, int rank, int nranks, const void *nccl_id128, bgpu_handle **out);
int bgpu_slab_info(const bgpu_handle *h, int *rank, int *nranks, int *x0, int *nx_local);
/* The same slab-decomposed chain with all its ranks inside ONE process on ONE device: rank r is driven by its own host
 * thread (every rank still makes the same sequence of calls, each from its thread), the collectives are host
 * rendezvous + device copies.  NCCL refuses two ranks on one GPU; this is how a single-GPU box runs -- and tests --
 * the slab code path (layouts, halos, fused transposes).  Not a performance configuration. */
typedef struct bgpu_local_group bgpu_local_group;
int bgpu_local_group_create(int nranks, bgpu_local_group **out);
void bgpu_local_group_destroy(bgpu_local_group *g);
int bgpu_slab_create_local(const bgpu_params *p, int rank, int nranks, bgpu_local_group *g, bgpu_handle **out);

/* static inputs: data->observational->{Power, nobs, noise_sf, window} (main.cc:150-154).
 * NULL leaves an array unchanged. */
int bgpu_set_static(bgpu_handle *h, const double *Power, const double *nobs, const double *noise,
                    const double *window);
/* hd->mass_f / hd->mass_r as read back from auxmass_{f,r}.dat (HMC.cc:414-423) */
int bgpu_set_mass(bgpu_handle *h, const double *mass_f, const double *mass_r);
/* S6 Hamiltonian_mass (HMC_mass.cc:315-368), types 0/1/4; outputs may be NULL */
int bgpu_hamiltonian_mass(bgpu_handle *h, double *mass_f_out, double *mass_r_out);
/* the same for every supported mass type; types 2 / 3 measure the spectrum of the likelihood force at `signal`
 * (hd->x; likeli_force_power, HMC_mass.cc:39-51) */
int bgpu_hamiltonian_mass_x(bgpu_handle *h, const double *signal, double *mass_f_out, double *mass_r_out);
/* likeli_force_power (HMC_mass.cc:39-51): binned spectrum of likelihood_grad_log_like(signal), N_bin entries each
 * (what the reference dumps as forcespec.dat); needs a handle created with mass_type 2 or 3 */
int bgpu_likeli_force_power(bgpu_handle *h, const double *signal, double *kmode, double *power);

/* S1 gradient_psi (HMC.cc:146-206): writes hd->gradpsi */
int bgpu_gradient_psi(bgpu_handle *h, const double *signal, double *gradpsi);
/* S2 psi (HMC.cc:124-143): n->psi_prior, n->psi_likeli and the hd->deltaX side effect (deltaX_out may be NULL) */
int bgpu_psi(bgpu_handle *h, const double *signal, double *psi_prior, double *psi_likeli, double *deltaX_out);
/* S3 kinetic_term (HMC.cc:64-121) */
int bgpu_kinetic(bgpu_handle *h, const double *momenta, double *K);
/* S4 Hamiltonian_EoM (HMC.cc:251-369) after its two RNG draws (Neps, epsilon stay with the host RNG): the whole
 * Neps-step trajectory on the device, stopped where the reference stops it (|momenta[0]| > 1e50, :360-364).  With a
 * Fourier-space mass and the Zel'dovich model the trajectory runs in k-space (s^ and p^ updated on the half grid, s
 * and p transformed once at each end); otherwise in real space with the kicks merged into the gradient's last store.
 * Either form agrees with the reference's step-by-step form to rounding (tests/test_leapfrog_forms_gpu.py). */
int bgpu_leapfrog(bgpu_handle *h, const double *s_i, const double *p_i, uint64_t Neps, double epsilon,
                  double *s_f, double *p_f);
/* S5 draw_momenta (HMC_momenta.cc:42-92) after the RNG: white = the 2*N doubles of
 * resolution_independent_random_grid_FS<double>(N1, rng, false) (random.hpp:36-120);
 * real_gauss = the N gsl_ran_gaussian draws of draw_real_space_momenta, or NULL when !mass_rs */
int bgpu_color_momenta(bgpu_handle *h, const double *white_complex_fullgrid, const double *real_gauss,
                       double *momenta);
/* S5 with the generator on the device: the same distribution as draw_momenta -- Gaussian
 * momenta with covariance M, DC mode zero -- from the counter-based Philox4x32-10 generator + Box-Muller,
 * keyed by (seed, draw_index).  NOT seed-compatible with the reference's GSL mt19937 stream (which is serial
 * and costs 2N host Gaussians per candidate); use bgpu_color_momenta when runs must reproduce the CPU code's.
 * bgpu_device_normals exposes the raw generator (elements [first, first + n) of a draw) for tests. */
int bgpu_draw_momenta_device(bgpu_handle *h, uint64_t seed, uint64_t draw_index, double *momenta);
int bgpu_device_normals(bgpu_handle *h, uint64_t seed, uint64_t draw_index, unsigned stream, size_t first, size_t n,
                        double *out);
/* One HMC candidate with the signal and the momenta resident on the device -- the body of HamiltonianMC's loop
 * (HMC.cc:436-506) between the host's RNG draws and its Metropolis decision: momenta from the device generator
 * (seed, draw_index), the Neps-step trajectory from the signal given to bgpu_set_signal, and the six energies of
 * delta_Hamiltonian: {H_kin_i, psi_prior_i, psi_likeli_i, H_kin_f, psi_prior_f, psi_likeli_f}.  p_f0 (optional)
 * receives momenta_f[0] for the reference's run-away message.  bgpu_accept makes the candidate the current signal
 * and copies it (and its deltaX, as psi() leaves it in hd->deltaX) to the host.  Per candidate only scalars cross
 * PCIe; with host arrays (bgpu_leapfrog, bgpu_psi, ...) it is ~10 array transfers from pageable memory. */
int bgpu_set_signal(bgpu_handle *h, const double *x);
int bgpu_candidate(bgpu_handle *h, uint64_t seed, uint64_t draw_index, uint64_t Neps, double epsilon, double *energies6,
                   double *p_f0);
int bgpu_accept(bgpu_handle *h, double *x_out, double *deltaX_out);
/* setup_random_test (barcoderunner.cc:42-205) and make_initial_guess (:207-247) with the generator on the device
 * (SURVEY 8f F4): delta_lag ~ GRF(Power), delta_eul = Lag2Eul(delta_lag) with the handle's own forward model, the
 * window by window_type (1 ones, 10 lower half masked, 23 as written there), nobs / noise_sf by data model and
 * likelihood -- Poisson counts, Gaussian with sigma = sigma_min + sigma_fac * Lambda clamped at 0 unless
 * negative_obs, the GRF test, or the log-normal data model.  nobs / noise / window become the handle's static
 * inputs; any output pointer may be null.  Counter-based Philox streams keyed by `seed`: the same distribution as
 * the reference's draw, NOT its GSL mt19937 stream (serial: 2 N^3 host Gaussians per field).  Single-GPU handles.
 * bgpu_initial_guess: initial_guess 0 (zeros), 2 (GRF), 3 (GRF smoothed with the Gaussian filter of width
 * smoothing_scale), 4 (0.1 * white noise); 1 (a file) stays with the host. */
typedef struct bgpu_mock_params {
  int window_type;   /* input.par window_type */
  int data_model;    /* 0 linear, 1 log-normal */
  double sigma_min, sigma_fac;
  int negative_obs;
} bgpu_mock_params;
int bgpu_mock_data(bgpu_handle *h, uint64_t seed, const bgpu_mock_params *mp, double *delta_lag, double *delta_eul,
                   double *nobs, double *noise, double *window);
int bgpu_initial_guess(bgpu_handle *h, uint64_t seed, int initial_guess, double smoothing_scale, double *signal);

/* Lag2Eul / Lag2Eul_rsd_zeldovich as likelihood_grad_log_like calls them (HMC_models.cc:383-406);
 * pos* may be NULL */
int bgpu_forward(bgpu_handle *h, const double *signal, double *deltaX, double *posx, double *posy, double *posz);
/* measure_spectrum (field_statistics.cpp:20-90): spherically binned power spectrum of a real field, N_bin bins
 * up to |k| of the cube's corner mode; the per-sample diagnostic behind powSpecit<it>.dat */
int bgpu_measure_spectrum(bgpu_handle *h, const double *signal, uint64_t N_bin, double *kmode, double *power);

/* building blocks exposed for parity tests */
int bgpu_assign_density(bgpu_handle *h, const double *x, const double *y, const double *z, double *rho);
int bgpu_cell_indices(bgpu_handle *h, const double *x, const double *y, const double *z, size_t n, int *ci,
                      int *cj, int *ck);
int bgpu_fft_r2c(bgpu_handle *h, const double *in, double *out_complex);       /* fftR2C, fftwrapper.cc:56-84 */
int bgpu_fft_c2r(bgpu_handle *h, const double *in_complex, double *out);       /* fftC2R, fftwrapper.cc:26-53 */
int bgpu_convolve_inv_corr(bgpu_handle *h, const double *signal, const double *corr, double *out); /* HMC_help.cc:16-64 */

/* device-pointer variants (fused on-device trajectory, PyTorch harness) */
int bgpu_set_stream(bgpu_handle *h, void *cuda_stream);
int bgpu_synchronize(bgpu_handle *h);
int bgpu_gradient_psi_dev(bgpu_handle *h, const double *d_signal, double *d_gradpsi);
int bgpu_psi_dev(bgpu_handle *h, const double *d_signal, double *psi_prior, double *psi_likeli, double *d_deltaX);
int bgpu_kinetic_dev(bgpu_handle *h, const double *d_momenta, double *K);
int bgpu_leapfrog_dev(bgpu_handle *h, double *d_signal, double *d_momenta, uint64_t Neps, double epsilon);
int bgpu_draw_momenta_device_dev(bgpu_handle *h, uint64_t seed, uint64_t draw_index, double *d_momenta);

/* ---------------------------------------------------------------------------
 * Single-precision mode.  The reference chooses its arithmetic at build time: SINGLE_PREC makes
 * real_prec = float and routes the transforms through fftwf (barlib/include/define_opt.h:50-59,
 * cmake/Modules/Options.cmake:65-66, barlib/src/fftwrapper.cc:32-36,62-66).  A SINGLE_PREC build binds
 * the entry points below where a DOUBLE_PREC build binds their bgpu_* namesakes (same seams S1-S4,
 * same argument meaning, real_prec arrays as float; energies stay double).
 * Built for the north-star path: Zel'dovich model (sfmodel 1 or rsd_model), CIC, plane-parallel RSD,
 * Poisson / Gaussian likelihood, calc_h 0 / 1 / 4, mass_type 0 / 1 / 4, N1 = 32 ... 512, one GPU, box at
 * the origin; bgpu_f32_create refuses anything else with a message.  Arrays and transforms are float,
 * reductions accumulate in double.  Tolerance against the FP64 path: 1e-5 relative (BASELINE.json).
 * --------------------------------------------------------------------------- */
typedef struct bgpu_f32_handle bgpu_f32_handle;
int bgpu_f32_create(const bgpu_params *p, bgpu_f32_handle **out);
void bgpu_f32_destroy(bgpu_f32_handle *h);
int bgpu_f32_set_static(bgpu_f32_handle *h, const float *Power, const float *nobs, const float *noise,
                        const float *window);
int bgpu_f32_set_mass(bgpu_f32_handle *h, const float *mass_f, const float *mass_r);
int bgpu_f32_hamiltonian_mass(bgpu_f32_handle *h, float *mass_f_out, float *mass_r_out);  /* HMC_mass.cc:315-368, types 0/1/4 */
int bgpu_f32_gradient_psi(bgpu_f32_handle *h, const float *signal, float *gradpsi);         /* S1: HMC.cc:146-206 */
int bgpu_f32_psi(bgpu_f32_handle *h, const float *signal, double *psi_prior, double *psi_likeli,
                 float *deltaX_out);                                                         /* S2: HMC.cc:124-143 */
int bgpu_f32_kinetic(bgpu_f32_handle *h, const float *momenta, double *K);                   /* S3: HMC.cc:64-121 */
int bgpu_f32_leapfrog(bgpu_f32_handle *h, const float *s_i, const float *p_i, uint64_t Neps, double epsilon,
                      float *s_f, float *p_f);                                               /* S4: HMC.cc:251-369 */
int bgpu_f32_set_stream(bgpu_f32_handle *h, void *cuda_stream);
int bgpu_f32_synchronize(bgpu_f32_handle *h);
int bgpu_f32_gradient_psi_dev(bgpu_f32_handle *h, const float *d_signal, float *d_gradpsi);
int bgpu_f32_psi_dev(bgpu_f32_handle *h, const float *d_signal, double *psi_prior, double *psi_likeli, float *d_deltaX);
int bgpu_f32_leapfrog_dev(bgpu_f32_handle *h, float *d_signal, float *d_momenta, uint64_t Neps, double epsilon);

/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t bgpu_kernel_launches(void);

/* per-kernel-class device timing (bench.py's roofline leg): CUDA events on the launching
 * stream around every launch between begin and end; a launch of the reduce / residual class
 * is a pair of kernels.  nkinds <= BGPU_PROFILE_KINDS. */
#define BGPU_PROFILE_KINDS 15
int bgpu_profile_begin(void);
int bgpu_profile_end(double *ms_per_kind, uint64_t *launches_per_kind, int nkinds);
const char *bgpu_profile_kind_name(int kind);

#ifdef __cplusplus
}
#endif

#endif /* BARCODE_GPU_H */
