// barcode_b200/csrc/kernels.h -- launchers of the non-FFT hot-path kernels
// (particles, likelihood, reductions, leapfrog, momentum colouring).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace bgpu {

struct GridGeom {
  int N;           // cells per axis (cubic)
  double L;        // box length
  double d;        // cell size L/N
  double min1, min2, min3;
  int masskernel;  // 0 NGP, 1 CIC, 2 TSC
  int rsd;         // plane-parallel redshift-space shift of z
  double cpecvel;  // f*100*E*a        (cosmo.cc:220-235)
  double v_norm;   // 1/Hub/a          (rsd.cc:27,39)
  double fgrow;    // f                (cosmo.cc:182-217)
  // x-slab decomposition; a full cube has x0 = 0, Ns = N, H = 0.
  int x0;          // global x index of the first Lagrangian plane this rank owns
  int Ns;          // planes owned
  int H;           // halo planes each side of the density tile: rho is [(Ns + 2H)][N][N], plane 0 = global x0 - H
  int *flag;       // set to 1 if a particle left the halo (device int), may be null when H == 0 and Ns == N
  int cellbound;   // displacements are cell-boundary averaged on read (cellboundcomp; non-Zel'dovich model)
  const double *cb_lo;  // slab: plane x0-1 of Psi_x, Psi_y, Psi_z ([3][N][N], from the lower neighbour); null on a cube
  int sweep;       // x-sweep scatter / gather (particles_sweep.cu): 0 off, 1 on, > 1 = planes per sweep segment
  int lean;        // lean per-particle arithmetic in the sweep kernels where it applies (BGPU_LEAN=0 turns it off)
  double sph_h;    // SPH scale length particle_kernel_h = h_rel * d (init_par.cc:379), masskernel 3
  const struct SphColumns *sph_cols;  // HOST pointer: the static column lists of particles_sph.cu (null: general SPH kernels)
};

// masskernel 3: the columns (i1, i2, half-range K) the scatter / the adjoint gather walk (particles_sph.cu)
struct SphColumns {
  int n_scatter = 0, n_gather = 0;
  void *dev = nullptr;  // int4[n_scatter + n_gather]
  bool z5 = false;      // every half-range K <= 2: the kernels with the unrolled z loop run
};
SphColumns *sph_columns_create(const GridGeom &g, const int *kmax_host, int R);  // null if the hull is too wide for them
void sph_columns_destroy(SphColumns *c);
void launch_scatter_sph_cols(const GridGeom &g, const SphColumns *c, const double *psix, const double *psiy,
                             const double *psiz, double *rho, double *posx, double *posy, double *posz, cudaStream_t st);
void launch_scatter_sph_cols_positions(const GridGeom &g, const SphColumns *c, const double *x, const double *y,
                                       const double *z, double *rho, cudaStream_t st);
void launch_gather_sph_cols(const GridGeom &g, const SphColumns *c, double *ax, double *ay, double *az,
                            const double *resid, double normalize, cudaStream_t st);

struct LikeParams {
  int likelihood;  // 0 Poisson, 1 Gaussian, 2 log-normal (3, Gaussian random field, never reaches the residual kernel)
  double rho_c, biasP, biasE;
  double delta_min;  // log-normal density floor
  int exact_sign;  // Poisson residual in the Gaussian sign convention (exact adjoint only)
};

// number of partial sums the two-stage reductions use (fixed: deterministic order)
constexpr int kReduceBlocks = 1184;  // 148 SMs x 8 resident CTAs
constexpr int kReduceThreads = 256;

// half-grid multiplier normFS / C(k) (0 where C <= 0), from a full real-indexed spectrum
void launch_inverse_spectrum(const double *full, double *half, int N, double normFS, cudaStream_t st);

// log-normal likelihood: f(delta_x) = log(rho_c (1 + max(delta_x, delta_min))) (lognormal_independent.cpp:57-79)
void launch_lognormal_f(const double *delta, double *out, size_t n, double rho_c, double delta_min, cudaStream_t st);
// Gaussian-random-field likelihood (gaussian_random_field.cpp:25-52): grad += (s - nobs)/sigma^2 where w > 0;
// -lnL = sum 1/2 ((s - nobs)/sigma)^2 where w > 0
void launch_grf_grad_add(double *grad, const double *s, const double *nobs, const double *noise, const double *window,
                         size_t n, cudaStream_t st);
void launch_grf_nll(const double *s, const double *nobs, const double *noise, const double *window, size_t n,
                    double *scratch, double *out, cudaStream_t st);

// counter-based Gaussians (Philox4x32-10 + Box-Muller): out[0..n) = elements [first, first + n) of (seed, draw, stream)
void launch_philox_normals(double *out, size_t n, size_t first, uint64_t seed, uint64_t draw, unsigned stream,
                           cudaStream_t st);
// W (half grid, transform of a real white field) *= sqrt(c2 * spec) at the folded index, DC = 0
void launch_colour_white(double2 *W, const double *spec_full, int N, double c2, cudaStream_t st);
// W *= 1/sqrt(inv) with inv = the padded half-grid multiplier (V/N)/M of the kinetic term, 0 where inv <= 0
void launch_colour_white_rows(double2 *W, const double *inv_half, int N, size_t n_half, bool owns_dc, cudaStream_t st);

// measure_spectrum (field_statistics.cpp:20-90) of a half-complex transform F; acc = [power | kmode | nmode] (3 nbin, device)
// (two steps: bin this rank's modes of the [x][Ns][N/2+1] layout, y = y0 + y_local; then -- after a slab has
// all-reduced acc -- normalise)
void launch_measure_spectrum_bin(const double2 *F, int N, int Ns, int y0, double L, int nbin, double *acc, cudaStream_t st);
void launch_measure_spectrum_finish(double *acc, int N, double L, int nbin, cudaStream_t st);

// mass types 2 / 3: factor (2/P + sqrt(F/P)), F = force_spec[bin(|k|)] (type 2) or `mean` (type 3)
void launch_force_mass(const double *power, const double *force_spec, double *mass_f, int N, int Ns, int x0, double L,
                       int nbin, int type, double mean, double factor, cudaStream_t st);

// particle scatter: Psi -> rho (zeroed here); optional positions out
void launch_scatter(const GridGeom &g, const double *psix, const double *psiy, const double *psiz, double *rho,
                    double *posx, double *posy, double *posz, cudaStream_t st);
// SPH spline mass assignment (masskernel 3, getDensity_SPH) and its exact adjoint (calc_h = 2,
// likelihood_calc_V_SPH): kmax[(i+R)(2R+1) + (j+R)] is the hull's k half-range of column (i, j), -1 = none
void launch_scatter_sph(const GridGeom &g, const double *psix, const double *psiy, const double *psiz, double *rho,
                        double *posx, double *posy, double *posz, cudaStream_t st);
void launch_gather_sph(const GridGeom &g, double *psix_Vx, double *psiy_Vy, double *psiz_Vz, const double *resid,
                       const int *kmax, int R, double normalize, cudaStream_t st);
// mass assignment on explicit positions (tests, bgpu_assign_density)
void launch_scatter_positions(const GridGeom &g, const double *x, const double *y, const double *z, double *rho,
                              cudaStream_t st);
// integer cell indices (bit-exact parity tests): lower CIC cell or NGP/TSC centre cell per axis
void launch_cell_indices(const GridGeom &g, const double *x, const double *y, const double *z, int *ci, int *cj,
                         int *ck, size_t n, cudaStream_t st);

// max |a[i]| -> *out (device scalar); scratch: kReduceBlocks doubles
void launch_max_abs(const double *a, size_t n, double *scratch, double *out, cudaStream_t st);
// dst[i] += src[i]
void launch_add(double *dst, const double *src, size_t n, cudaStream_t st);
// slab variant of launch_inverse_spectrum, stage 1: the multiplier of my x planes as complex numbers in
// the packed all-to-all layout [peer][x_l][y_l][z <= N/2]; stage 2 (after the all-to-all): real parts
// into the transposed half grid [x][y_l][N/2+2]
void launch_inverse_spectrum_pack(const double *full_xslab, double2 *packed, int N, int Ns, double normFS,
                                  cudaStream_t st);
void launch_inverse_spectrum_unpack(const double2 *transposed, double *half, int N, int Ns, cudaStream_t st);

// out = mult * s + norm * h on a half grid (mult has the padded row pitch N/2+2)
void launch_kfinal_combine(const double2 *s, const double *mult, const double2 *h, double2 *out, double norm, int N,
                           size_t n_half, cudaStream_t st);

// Lag2Eul_non_zeldovich's real-space pieces (Lag2Eul.cc:138-268): 2LPT source D1 dQ s - D2 delta2(phi),
// spherical-collapse divergence, and the ALPT combination K D2^ + (1 - K) D4^ (in place over d2)
// Ns / xoff: phi holds planes [-xoff, Ns + xoff) of this rank's slab (xoff = 4 halo planes from the x neighbours,
// the reach of the twice-applied stencil); a cube passes Ns = N, xoff = 0 and wraps in x
void launch_lpt2_source(const double *phi, const double *s, double *out, int N, int Ns, int xoff, double L, double dQ,
                        double D1, double D2, cudaStream_t st);
void launch_sc_divergence(const double *s, double *out, size_t n, double dQ, double D1, cudaStream_t st);
// exact adjoint of the 2LPT/ALPT model (cube): transpose of the cell-boundary average; P_ab = dm2v/dL_ab * u
// (six arrays); G = sum_ab FD_a FD_b P_ab; out = dQ (D1 u_lpt - D2 q + theta_SC' u_sc)
void launch_cellbound_transpose(const double *v, double *w, int N, cudaStream_t st);
void launch_lpt2_adjoint_coef(const double *phi, const double *u, double *const out[6], int N, double L, cudaStream_t st);
void launch_lpt2_adjoint_div(double *const in[6], double *out, int N, double L, cudaStream_t st);
void launch_alpt_adjoint_combine(double *out, const double *u_lpt, const double *q, const double *u_sc, const double *s,
                                 size_t n, double dQ, double D1, double D2, cudaStream_t st);
// k-space layout [x][Ns][N/2+1] with y = y0 + y_local (a cube: Ns = N, y0 = 0)
void launch_alpt_combine(double2 *d2, const double2 *d4, int N, int Ns, int y0, double kfac, double rS, cudaStream_t st);

// deterministic sum of an array -> *out (device scalar); scratch: kReduceBlocks doubles
void launch_sum(const double *a, size_t n, double *scratch, double *out, cudaStream_t st);
// deterministic 0.5 * sum a*b
void launch_half_dot(const double *a, const double *b, size_t n, double *scratch, double *out, cudaStream_t st);
// kinetic real-space part: 0.5 * sum p * (c + p/mass_r) with mass_r <= 0 -> 0; c may be null
void launch_kinetic(const double *p, const double *conv, const double *mass_r, size_t n, double *scratch,
                    double *out, cudaStream_t st);

// delta = rho / mean - 1 (in place), residual r, and -lnL partial sum -> *nll (device scalar).
// mean is read from the device scalar `sum_rho` (/N).  resid may be null (value only); nll may be null (residual
// only: a gradient evaluation), and then keep_delta = false leaves rho_delta untouched as well.
void launch_overdens_residual(const LikeParams &lp, double *rho_delta, const double *sum_rho, const double *nobs,
                              const double *noise, const double *window, double *resid, size_t n, double ncells_global,
                              double *scratch, double *nll, cudaStream_t st, bool keep_delta = true);

// x-sweep variants (particles_sweep.cu): full cube, Lagrangian lattice without cell-boundary averaging, N >= 32.
// launch_scatter / launch_gather_adjoint dispatch to them where they apply (BGPU_SWEEP=0 keeps the first-generation kernels).
bool scatter_sweep_applicable(const GridGeom &g, const double *posx);
void launch_scatter_sweep(const GridGeom &g, const double *psix, const double *psiy, const double *psiz, double *rho,
                          cudaStream_t st);
bool gather_sweep_applicable(const GridGeom &g);
void launch_gather_sweep(const GridGeom &g, double *ax, double *ay, double *az, const double *resid, cudaStream_t st);

// exact adjoint of the mass assignment: V_c(p) = sum_cells r_c dW_c/dx_c, in place over Psi
void launch_gather_adjoint(const GridGeom &g, double *psix_Vx, double *psiy_Vy, double *psiz_Vz,
                           const double *resid, cudaStream_t st);
// the same out of place (required under g.cellbound, where a particle's position also reads its neighbour's Psi)
void launch_gather_adjoint_to(const GridGeom &g, const double *psix, const double *psiy, const double *psiz, double *vx,
                              double *vy, double *vz, const double *resid, cudaStream_t st);

// out = resid * d_c(delta) with the 4th-order finite difference of gradient.cpp:81-153
void launch_findif_product(const double *delta, const double *resid, double *out, int N, int Ns, int xo, double L,
                           int comp, cudaStream_t st);

// y += a * x ; y = a * x ; y += a * x / m (m <= 0 -> 0)
void launch_axpy(double *y, const double *x, double a, size_t n, cudaStream_t st, const int *skip = nullptr);
// the run-away test of Hamiltonian_EoM (HMC.cc:360-364) on the device, see kernels.cu: scal2 = {stopped at step,
// momenta[0] after the previous kick}; flag (may be null) is raised with the step where the trajectory stopped
// leapfrog in k-space (kernels.cu): drift s^ += eps (V/N)/M p^ ; kick p^ += a ((V/N)/P s^ + norm h^), which also returns
// momenta[0] = (1/N) sum_k p^_k in *p0_out
void launch_kspace_drift(double2 *shat, const double2 *phat, const double *inv_mass, double eps, int N, size_t nh,
                         cudaStream_t st, const int *skip);
void launch_kspace_kick(double2 *phat, const double2 *shat, const double2 *hhat, const double *prior, double a,
                        double norm, int N, size_t nh, double ncells, double *scratch, double *p0_out, cudaStream_t st,
                        const int *skip);
void launch_runaway_guard(const double *p, double *scal2, int *flag, int step, int mode, cudaStream_t st);
void launch_runaway_apply(const double *scal2, int *flag, cudaStream_t st);
void launch_scale(double *y, const double *x, double a, size_t n, cudaStream_t st);
void launch_axpy_div(double *y, const double *x, const double *m, double a, size_t n, cudaStream_t st,
                     const int *skip = nullptr);
void launch_fill(double *y, double v, size_t n, cudaStream_t st);

// Hamiltonian mass types 0 / 1 / 4 (HMC_mass.cc:315-368)
void launch_mass(const double *power, double *mass_f, double *mass_r, int mass_type, double mass_factor, size_t n,
                 cudaStream_t st);

// create_GARFIELD colouring + Hermitian symmetrisation into the half array (random.cpp:102-507)
void launch_colour_momenta(const double2 *white_full, const double *spec_full, double2 *half, int N, double amp,
                           cudaStream_t st);
// the same on the rows [x][y0 + y_local][z <= N/2] of a slab-decomposed chain, sigma from the kinetic term's multiplier
void launch_colour_momenta_rows(const double2 *white_full, const double *inv_half, double2 *half, int N, int Ns, int y0,
                                double ncells, cudaStream_t st);
// p += sqrt(mass_r) * gauss (HMC_momenta.cc:76-92)
void launch_add_real_momenta(double *p, const double *mass_r, const double *gauss, size_t n, cudaStream_t st);

// calc_h = 3 (likelihood_calc_V_SPH_fourier_TSC, HMC_models_testing.cpp:54-188)
//   out = i k_c hW(k) r^(k) on the cube's half grid, hW = h * SPH_kernel_F with the multipliers' padded pitch
void launch_sph_fourier_comp(const double2 *rhat, const double *hW_half, double2 *out, int N, double kfac, int comp,
                             cudaStream_t st);
//   out[p] = v + fz * v, v = interpolate_TSC(field, x_p) at the particle positions recomputed from Psi
//   (interpolate_grid.cpp:134-202; fz = the growth rate for the z component under RSD, else 0)
void launch_interp_tsc(const GridGeom &g, const double *psix, const double *psiy, const double *psiz,
                       const double *field, double *out, double fz, cudaStream_t st);

// setup_random_test's window / nobs / noise (barcoderunner.cc:94-195) drawn on the device (Philox, not GSL's stream)
struct MockObs {
  int likelihood, data_model, window_type, negative_obs;
  double rho_c, delta_min, sigma_min, sigma_fac;
};
void launch_mock_obs(const MockObs &mp, const double *delta_eul, const double *delta_lag, double *window, double *nobs,
                     double *noise, size_t n, size_t first, size_t n_global, uint64_t seed, cudaStream_t st);

// out = 1/2 sum_x a (C^-1 a) from a^ on the half grid (Parseval; see kernels.cu HalfQuadF)
void launch_half_quadratic(const double2 *vhat, const double *mult_half, int N, size_t n_half, double ncells,
                           double *scratch, double *out, cudaStream_t st);

}  // namespace bgpu
