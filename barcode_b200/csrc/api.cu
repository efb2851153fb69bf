// barcode_b200/csrc/api.cu -- implementation of the C ABI in include/barcode_gpu.h:
// the device-resident state of one HMC chain (the GPU counterpart of the
// reference's HAMIL_DATA, struct_hamil.h:146-400) and the launch sequences for
// gradient_psi / psi / kinetic_term / Hamiltonian_EoM / draw_momenta /
// Hamiltonian_mass (HMC.cc, HMC_momenta.cc, HMC_mass.cc).
#include "barcode_gpu.h"

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <stdexcept>
#include <string>
#include <thread>

#include "fft3d.h"
#include "host_math.h"
#include "kernels.h"
#include "nccl_comm.h"
#include "util.h"

namespace bgpu {
std::atomic<uint64_t> g_kernel_launches{0};
Profiler g_prof;

void Profiler::record_start(int k, cudaStream_t st) {
  if (2 * used + 2 > ev.size()) {
    const size_t old = ev.size();
    ev.resize(old + 1024);
    for (size_t i = old; i < ev.size(); ++i) BGPU_CUDA(cudaEventCreate(&ev[i]));
  }
  if (kind.size() <= used) kind.resize(used + 512);
  kind[used] = k;
  BGPU_CUDA(cudaEventRecord(ev[2 * used], st));
}
void Profiler::record_stop(cudaStream_t st) {
  cudaEventRecord(ev[2 * used + 1], st);
  ++used;
}
}  // namespace bgpu

using namespace bgpu;

static thread_local std::string g_last_error;
namespace bgpu {
void set_last_error(const std::string &msg) { g_last_error = msg; }  // f32_path.cu reports through the same string
}

struct bgpu_handle {
  bgpu_params p{};
  int N = 0;
  size_t n = 0;    // N^3
  size_t nh = 0;   // N^2 (N/2+1)
  size_t nhp = 0;  // N^2 (N/2+2): padded row pitch of the real half-grid multipliers
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Fft3d fft;
  GridGeom geom{};
  LikeParams like{};
  double kfac = 0.0, normFS = 0.0;
  bool mass_fs = false, mass_rs = false;
  bool have_power = false, have_obs = false, have_mass = false;

  // static inputs
  double *power = nullptr, *nobs = nullptr, *noise = nullptr, *window = nullptr;
  double *inv_power = nullptr;               // half grid, (V/N)/P
  double *zero_half = nullptr;               // half grid of zeros: the prior switched off (mass types 2 / 3)
  bool like_only = false;                    // gradient_device leaves the prior out (likelihood force)
  double *mass_f = nullptr, *mass_r = nullptr;
  double *inv_mass = nullptr;                // half grid, (V/N)/M_f
  // state / scratch (real)
  double *sig = nullptr, *mom = nullptr, *grad = nullptr;
  double *psi[3] = {nullptr, nullptr, nullptr};
  double *delta = nullptr, *resid = nullptr, *tmp = nullptr;
  // scratch (half-complex)
  double2 *shat = nullptr, *dhat = nullptr, *work = nullptr, *acc = nullptr;
  double2 *ubuf = nullptr;  // BGPU_SHARE_X=1: the x-passed field / the y-passed sum that a component triple shares
  // reductions
  double *partials = nullptr, *dscal = nullptr;
  double *hscal = nullptr;  // pinned

  // x-slab decomposition.  A cube handle has G = 1, Ns = N, x0 = 0 and no halo;
  // n / nh / nhp are always the LOCAL element counts.
  int G = 1, rank = 0, Ns = 0, x0 = 0;
  double ncells = 0.0;          // N^3, the global cell count (FFT normalisation, mean density)
  ChainComm *comm = nullptr;      // NCCL (one process per GPU) or the in-process communicator (ranks = threads, tests)
  int Hmax = 0;                 // halo planes allocated each side of rho_ext
  double *rho_ext = nullptr;    // [(Ns + 2 Hmax)][N][N]; delta points at the owned planes inside it
  double *halo_recv = nullptr;  // 2 * Hmax * N^2
  double *cand_s = nullptr, *cand_p = nullptr, *cur_s = nullptr;  // device-resident HMC candidate and current signal (bgpu_candidate)
  bool have_signal = false;
  // psi() of the current signal (prior, -lnL): valid until the signal or the static inputs change.  A candidate's
  // initial energies are the final ones of the candidate that was accepted last (or its own predecessor's initial
  // ones after a rejection); the reference recomputes them every time (HMC.cc:214-215)
  bool cur_psi_valid = false;
  double cur_psi[2] = {0., 0.}, cand_psi[2] = {0., 0.};
  bool cache_psi = true;         // BGPU_CANDIDATE_CACHE=0: recompute, as the reference does
  // The last kick of a k-space trajectory evaluates the forward model at s_f: with the Gaussian likelihood (whose
  // psi() and gradient share deltaQ_factor and the RSD switch) it can return psi(s_f) -- prior by Parseval from s^,
  // -lnL from the residual kernel's value variant, deltaX left in h->delta -- and bgpu_candidate skips psi(s_f).
  bool psi_at_end = false;       // request (bgpu_candidate -> leapfrog_kspace)
  bool want_psi = false;         // gradient_device: this evaluation also fills dscal[S_PRIOR], dscal[S_NLL], delta
  bool psi_from_kick = false;    // answer: dscal holds psi(s_f)
  bool constructed = false;
  // fused leapfrog (HMC.cc:251-369): the kick p += kick_a * gradpsi rides on the store of the gradient's last z pass
  // (bulk f64 reduce-add), the drift on the store of M^-1 p's; a device flag stops a run-away trajectory
  bool kick_on = false;          // gradient_device: apply kick_a * gradpsi to d_out (+=) instead of storing gradpsi
  double kick_a = 0.0;
  // leapfrog in k-space (leapfrog_device): s^ lives in shat, p^ in phat for the whole trajectory; gradient_device then
  // neither transforms s nor inverse-transforms gradpsi: its last sum goes straight into p^ (kernels.cu launch_kspace_kick)
  bool kspace_lf = true;         // BGPU_LEAPFROG_KSPACE=0: the fused real-space form
  bool kspace_on = false;        // gradient_device: shat is current, finish with the k-space kick
  double2 *phat = nullptr;
  int *stopflag = nullptr;       // device: the step after which |momenta[0]| > 1e50 stopped the trajectory, 0 = running
  int *hflag2 = nullptr;         // pinned copy
  bool parseval = true;          // BGPU_PARSEVAL=0: kinetic energy and prior through the inverse transform, as the reference
  bool fused_leapfrog = true;    // BGPU_LEAPFROG_FUSED=0: the step-by-step form with separate kicks and a host test per step     // create_impl ran to completion (the destructor's collectives are safe)
  double *phi1 = nullptr, *xa = nullptr, *xb = nullptr, *xc = nullptr;  // exact 2LPT/ALPT adjoint: phi^(1) + 3 scratch arrays
  double *fext = nullptr;       // log-normal + calc_h 0 on a slab: f(delta_x) with 2 halo planes each side
  double *resid_ext = nullptr;  // exact adjoint on a slab: the residual with H halo planes each side
  int H_cur = 0;                // halo width of the evaluation in flight
  double2 *sendbuf = nullptr, *recvbuf = nullptr;
  int *dflag = nullptr;         // device flag: a particle left the halo
  int *hflag = nullptr;         // pinned copy
  void *peer_base[8] = {};      // other ranks' receive buffers, opened through CUDA IPC

  // host-pointer API: input rows stream in under the first z pass, result rows stream out under the last
  cudaStream_t copy_stream = nullptr;
  static constexpr int kChunks = 8;
  cudaEvent_t ev_up[kChunks] = {}, ev_dn[kChunks] = {};
  const ChunkHooks *in_hooks = nullptr, *out_hooks = nullptr;

  // SPH kernel hull (SPH_kernel_3D_cells_hull_1): k half-range per (i, j) column, on the device
  int *sph_kmax = nullptr;
  int sph_R = 0;
  double *sph_hW = nullptr;  // calc_h = 3: h * SPH_kernel_F on the padded half grid (HMC_models_testing.cpp:88-105)
};

namespace {

enum Scal { S_SUMRHO = 0, S_NLL = 1, S_PRIOR = 2, S_KIN = 3, S_P0 = 4, S_MAXPSI = 5, S_STOP = 6, S_P0PREV = 7, S_COUNT = 8 };

template <class T>
void dalloc(T *&ptr, size_t count) {
  BGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&ptr), count * sizeof(T)));
}

void h2d(bgpu_handle *h, double *dst, const double *src, size_t count) {
  BGPU_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
}
void d2h(bgpu_handle *h, double *dst, const double *src, size_t count) {
  BGPU_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
}
void sync(bgpu_handle *h) {
  if (h->dflag)
    BGPU_CUDA(cudaMemcpyAsync(h->hflag, h->dflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  BGPU_CUDA(cudaStreamSynchronize(h->stream));
  if (h->dflag && *h->hflag) {
    *h->hflag = 0;
    cudaMemsetAsync(h->dflag, 0, sizeof(int), h->stream);
    throw std::runtime_error("bgpu: a particle moved beyond the mass-assignment halo of its slab (displacement > " +
                             std::to_string(h->Hmax) + " cells along x); use fewer ranks for this field");
  }
}

// sum a device scalar over the ranks of a slab-decomposed chain (no-op on a cube)
void allreduce_scalar(bgpu_handle *h, int which) {
  if (h->G > 1) h->comm->all_reduce_sum(h->dscal + which, 1, h->stream);
}

void require(bool ok, const char *msg) {
  if (!ok) throw std::runtime_error(msg);
}

// E_Hubble_a, fgrow, c_pecvel: cosmo.cc:26-31,182-235 (host_math.h, shared with the single-precision mode)
double E_Hubble_a(double a, double OM, double OL) { return host_E_Hubble_a(a, OM, OL); }
double fgrow1(double a, double OM, double OL) { return host_fgrow(a, OM, OL); }

void validate(const bgpu_params &p) {
  require(p.N1 == p.N2 && p.N2 == p.N3, "bgpu: only cubic grids are supported (the reference sets N2=N3=N1, init_par.cc:116-122)");
  require(Fft3d::supported(p.N1), "bgpu: N1 must be a power of two in [8, 1024]");
  require(p.L1 == p.L2 && p.L2 == p.L3 && p.L1 > 0, "bgpu: only cubic boxes are supported");
  require(p.masskernel >= 0 && p.masskernel <= 3, "bgpu: masskernel must be 0 (NGP), 1 (CIC), 2 (TSC) or 3 (SPH)");
  if (p.masskernel == 3)
    require(p.particle_kernel_h_rel > 0. && p.particle_kernel_h_rel <= p.N1 / 4.,
            "bgpu: particle_kernel_h_rel must be in (0, N1/4] for the SPH kernel (init_par.cc:373-375)");
  require(p.likelihood >= 0 && p.likelihood <= 3,
          "bgpu: likelihood must be 0 (Poisson), 1 (Gaussian), 2 (log-normal) or 3 (Gaussian random field)");
  require(p.calc_h == 0 || p.calc_h == 1 || p.calc_h == 2 || p.calc_h == 3 || p.calc_h == BGPU_CALC_H_EXACT,
          "bgpu: calc_h must be 0, 1, 2 (SPH adjoint), 3 (its Fourier / TSC variant) or 4 (NGP/CIC/TSC adjoint)");
  require(p.calc_h == 0 || p.calc_h == 1 || ((p.calc_h == 2 || p.calc_h == 3) && p.masskernel == 3) ||
              (p.calc_h == BGPU_CALC_H_EXACT && p.masskernel != 3),
          "Must use SPH mass kernel (masskernel = 3) when using likelihood_calc_h_SPH (calc_h = 2 or 3)!  (The exact "
          "adjoint of NGP/CIC/TSC is calc_h = 4.)");
  require(p.mass_type >= 0 && p.mass_type <= 4,
          "bgpu: mass_type must be 0 ... 4 on the GPU path (5/6/60, the first-order likelihood-force expansion, are O(N) FFTs of cold set-up)");
  if (p.mass_type == 2 || p.mass_type == 3)
    require(p.N_bin >= 1 && p.N_bin <= 2048, "bgpu: N_bin must be in [1, 2048] for mass_type 2 / 3");
  if (p.sfmodel != 1 && !p.rsd_model) {
    // Lag2Eul_non_zeldovich (2LPT + spherical collapse split at slength).  The reference has no adjoint
    // for it (HMC_models.cc:458): its gradients are calc_h 0 / 1 on the forward density.
    // calc_h 0 / 1 are the reference's gradients on the forward density; calc_h = 4 is the exact adjoint (new)
    require(p.calc_h == 0 || p.calc_h == 1 || (p.calc_h == BGPU_CALC_H_EXACT && p.masskernel != 3),
            "bgpu: sfmodel != 1 supports calc_h 0, 1 and 4 (NGP / CIC / TSC); the reference's SPH adjoints (calc_h 2, 3) "
            "are Zel'dovich-only (HMC_models.cc:442-458)");
    if (p.calc_h == BGPU_CALC_H_EXACT)
      require(p.correct_delta != 0, "bgpu: the exact adjoint of the 2LPT/ALPT model needs correct_delta = true");
    require(p.slength > 0., "bgpu: sfmodel != 1 needs slength > 0 (the ALPT smoothing radius, input.par slength)");
  }
  if (p.rsd_model) {
    require(p.planepar != 0, "Non-plane-parallel RSD model is not yet implemented in calc_V! Use planepar = true.");
    require(p.periodic != 0, "bgpu: RSD needs periodic boundary conditions (rsd.cc:59-64)");
  }
  require(p.periodic != 0, "bgpu: only periodic boundary conditions are supported (disp_part.cc:28)");
}

// ---------------------------------------------------------------------------
// device pipelines
// ---------------------------------------------------------------------------

void r2c_plain(bgpu_handle *h, const double *in, double2 *out);

// Lag2Eul_zeldovich / _rsd_zeldovich / _non_zeldovich from s (d_s) and s^ (already in h->shat): Psi -> rho -> sum(rho).
// dQ / rsd are arguments because the Poisson log-likelihood ignores both (poissonian.cpp:54-56).
void forward_from_shat(bgpu_handle *h, const double *d_s, double dQ, bool rsd, double *px, double *py, double *pz) {
  const double inv_n = 1.0 / h->ncells;
  const bool zeldovich = (h->p.sfmodel == 1 || h->p.rsd_model);  // HMC_models.cc:395-406
  ROp scale_n;
  scale_n.kind = R_SCALE;
  scale_n.a = inv_n;
  const double2 *disp_src = h->shat;
  double disp_a = -h->p.D1 * dQ;  // in = dQ * s ; phi = -D1 * in (Lag2Eul.cc:88)
  if (!zeldovich) {
    // Lag2Eul_non_zeldovich (Lag2Eul.cc:138-268) in 7 transforms instead of the reference's 22:
    //   phi1 = IFFT[-s^/k^2]; d2 = D1 in - D2 delta2(phi1); sc = spherical collapse(in);
    //   C^ = K FFT[d2] + (1 - K) FFT[sc]; Psi_c = IFFT[-i k_c/k^2 C^] -- the projection and the smoothing
    //   are linear, so K o Psi^LPT + Psi^SC - K o Psi^SC is formed once in k-space
    KOp pois;
    pois.kind = K_NEGINVK2;
    pois.a = dQ;
    pois.kfac = h->kfac;
    const size_t plane = (size_t)h->N * h->N;
    if (h->G == 1) {
      double *phi = h->phi1 ? h->phi1 : h->psi[0];  // the exact adjoint needs phi^(1) again
      h->fft.c2r(h->shat, h->work, phi, pois, scale_n);
      launch_lpt2_source(phi, d_s, h->psi[1], h->N, h->N, 0, h->p.L1, dQ, h->p.D1, h->p.D2, h->stream);
    } else {
      // slab: the twice-applied 4th-order stencil reaches 4 planes across the slab boundary.  phi^(1) goes into
      // the (still unused) density tile with 4 halo planes each side, filled from the x neighbours.
      constexpr int XH = 4;
      double *ext = h->rho_ext, *phi = ext + XH * plane;
      h->fft.c2r(h->shat, h->work, phi, pois, scale_n);
      const int lo = (h->rank + h->G - 1) % h->G, hi = (h->rank + 1) % h->G;
      {
        ProfScope prof(KK_HALO, h->stream);
        // my first planes are rank-1's upper halo, my last planes rank+1's lower halo
        h->comm->exchange2(phi, lo, phi + (size_t)(h->Ns - XH) * plane, hi, ext + (size_t)(h->Ns + XH) * plane, hi, ext, lo,
                           XH * plane, h->stream);
      }
      launch_lpt2_source(ext, d_s, h->psi[1], h->N, h->Ns, XH, h->p.L1, dQ, h->p.D1, h->p.D2, h->stream);
    }
    r2c_plain(h, h->psi[1], h->dhat);
    launch_sc_divergence(d_s, h->psi[2], h->n, dQ, h->p.D1, h->stream);
    r2c_plain(h, h->psi[2], h->acc);
    launch_alpt_combine(h->dhat, h->acc, h->N, h->Ns, h->G > 1 ? h->rank * h->Ns : 0, h->kfac, h->p.slength, h->stream);
    disp_src = h->dhat;
    disp_a = 1.0;  // theta2velcomp on +theta (Lag2Eul.cc:238): the sign differs from the Zel'dovich branch
  }
  // Psi^_c = a (k_c/k^2)(Im f^, -Re f^)
  if (h->ubuf) {
    // shared x pass: k_y and k_z are constant along an x pencil, so B = IFFT_x[a/k^2 (Im f^, -Re f^)] (same zeroing
    // rules) serves both: Psi_y = IFFT_zy[k_y B], Psi_z = IFFT_zy[k_z B] -- 2 x passes instead of 3
    KOp lop;
    lop.kind = K_DISP;
    lop.comp = K_COMP_UNIT;
    lop.a = disp_a;
    lop.kfac = h->kfac;
    h->fft.xpass(disp_src, h->ubuf, +1, lop, KOp{});
    for (int c = 2; c >= 1; --c) {
      KOp yl;
      yl.kind = K_MULK;
      yl.comp = c;
      yl.kfac = h->kfac;
      h->fft.c2r_yz(h->ubuf, h->work, h->psi[c], yl, scale_n);
    }
    lop.comp = 0;
    h->fft.c2r(disp_src, h->work, h->psi[0], lop, scale_n);
  } else if (h->fft.can_share_x_slab()) {
    // the same on a slab chain: ONE transposing x pass whose result stays in the receive buffer, read by the y passes
    // of Psi_z and Psi_y -- 2 transposes for the three displacement components instead of 3
    KOp lop;
    lop.kind = K_DISP;
    lop.comp = K_COMP_UNIT;
    lop.a = disp_a;
    lop.kfac = h->kfac;
    h->fft.xpass_shared_inverse(disp_src, lop);
    for (int c = 2; c >= 1; --c) {
      KOp yl;
      yl.kind = K_MULK;
      yl.comp = c;
      yl.kfac = h->kfac;
      h->fft.c2r_yz_shared(h->work, h->psi[c], yl, scale_n);
    }
    lop.comp = 0;
    h->fft.c2r(disp_src, h->work, h->psi[0], lop, scale_n);
  } else
  for (int c = 2; c >= 0; --c) {
    KOp lop;
    lop.kind = K_DISP;
    lop.comp = c;
    lop.a = disp_a;
    lop.kfac = h->kfac;
    h->fft.c2r(disp_src, h->work, h->psi[c], lop, scale_n);
  }
  GridGeom g = h->geom;
  g.rsd = rsd ? 1 : 0;
  g.cellbound = zeldovich ? 0 : 1;
  g.cb_lo = nullptr;
  const size_t plane = (size_t)h->N * h->N;
  if (g.cellbound && h->G > 1) {
    // cellboundcomp averages with the (i-1, j-1, k-1) neighbour: plane x0-1 of the three displacement
    // components comes from the lower neighbour (its last plane), mine goes to the upper one
    const int lo = (h->rank + h->G - 1) % h->G, hi = (h->rank + 1) % h->G;
    ProfScope prof(KK_HALO, h->stream);
    for (int c = 0; c < 3; ++c)
      h->comm->shift(h->psi[c] + (size_t)(h->Ns - 1) * plane, hi, h->halo_recv + (size_t)c * plane, lo, plane, h->stream);
    g.cb_lo = h->halo_recv;
  }
  if (h->G > 1) {
    // halo width from the largest x displacement on any rank (+1 plane for the upper CIC / TSC
    // neighbour, +1 for the lower TSC neighbour and rounding)
    launch_max_abs(h->psi[0], h->n, h->partials, h->dscal + S_MAXPSI, h->stream);
    h->comm->all_reduce_max(h->dscal + S_MAXPSI, 1, h->stream);
    BGPU_CUDA(cudaMemcpyAsync(h->hscal + S_MAXPSI, h->dscal + S_MAXPSI, sizeof(double), cudaMemcpyDeviceToHost,
                              h->stream));
    BGPU_CUDA(cudaStreamSynchronize(h->stream));
    int H = (int)std::ceil(h->hscal[S_MAXPSI] / g.d) + 2;
    // the SPH spline reaches sph_R cells around the particle's own cell (massFunctions.cc:420, SPH_kernel.cpp:62-139)
    if (g.masskernel == 3) H = (int)std::ceil(h->hscal[S_MAXPSI] / g.d) + 1 + h->sph_R;
    if (!(h->hscal[S_MAXPSI] == h->hscal[S_MAXPSI]) || H > h->Hmax)
      throw std::runtime_error("bgpu: displacement of " + std::to_string(h->hscal[S_MAXPSI] / g.d) +
                               " cells along x" + (g.masskernel == 3 ? " plus the SPH kernel's reach" : "") +
                               " exceeds the slab halo (" + std::to_string(h->Hmax) + " planes)");
    g.H = H;
    h->H_cur = H;
    h->delta = h->rho_ext + (size_t)H * plane;
  }
  if (g.masskernel == 3)
    launch_scatter_sph(g, h->psi[0], h->psi[1], h->psi[2], h->rho_ext, px, py, pz, h->stream);
  else
    launch_scatter(g, h->psi[0], h->psi[1], h->psi[2], h->rho_ext, px, py, pz, h->stream);
  if (h->G > 1) {
    // my lower halo belongs to rank-1's last planes, my upper halo to rank+1's first planes
    const int H = g.H, lo = (h->rank + h->G - 1) % h->G, hi = (h->rank + 1) % h->G;
    const size_t cnt = (size_t)H * plane;
    double *ext = h->rho_ext;
    {
      ProfScope prof(KK_HALO, h->stream);
      h->comm->exchange2(ext, lo, ext + (size_t)(h->Ns + H) * plane, hi, h->halo_recv, hi, h->halo_recv + cnt, lo, cnt,
                         h->stream);
    }
    launch_add(ext + (size_t)h->Ns * plane, h->halo_recv, cnt, h->stream);        // from rank+1 -> my last H planes
    launch_add(ext + (size_t)H * plane, h->halo_recv + cnt, cnt, h->stream);      // from rank-1 -> my first H planes
  }
  launch_sum(h->delta, h->n, h->partials, h->dscal + S_SUMRHO, h->stream);
  allreduce_scalar(h, S_SUMRHO);
}

void r2c_plain(bgpu_handle *h, const double *in, double2 *out) {
  ROp lop;
  lop.kind = R_LOAD;
  h->fft.r2c(in, out, nullptr, lop, KOp{});
}

// h->acc = sum_c (k_c/k^2)(Im V^_c, -Re V^_c) (grad_inv_lap_FS + add_to_array, gradient.cpp:157-211): three forward
// transforms whose x passes accumulate -- or, with the shared x pass, (k_x V^_x + FFT_x[k_y V~_y + k_z V~_z]) / k^2
// with the sum of the y and z components formed after their y passes: 2 x passes instead of 3
void backproject(bgpu_handle *h, double *const V[3]) {
  ROp lop2;
  lop2.kind = R_LOAD;
  KOp sop2;
  sop2.kfac = h->kfac;
  if (h->ubuf) {
    sop2.kind = K_INVLAP_SET;
    sop2.comp = 0;
    h->fft.r2c(V[0], h->work, h->acc, lop2, sop2);
    for (int c = 1; c < 3; ++c) {
      KOp ys;
      ys.kind = (c == 1) ? K_MULK_SET : K_MULK_ADD;
      ys.comp = c;
      ys.kfac = h->kfac;
      h->fft.r2c_zy(V[c], h->work, h->ubuf, lop2, ys);
    }
    sop2.kind = K_INVLAP_ADD;
    sop2.comp = K_COMP_UNIT;
    h->fft.xpass(h->ubuf, h->acc, -1, KOp{}, sop2);
    return;
  }
  for (int c = 0; c < 3; ++c) {
    sop2.kind = (c == 0) ? K_INVLAP_SET : K_INVLAP_ADD;
    sop2.comp = c;
    h->fft.r2c(V[c], h->work, h->acc, lop2, sop2);
  }
}

// likelihood_grad_log_like + prior + sum (HMC.cc:146-206, HMC_models.cc:377-471):
// d_out (op) a_out * gradpsi(d_s), with op = set (R_SCALE) or accumulate (R_AXPY).
void gradient_device(bgpu_handle *h, const double *d_s, double *d_out) {
  require(h->have_power && h->have_obs, "bgpu: bgpu_set_static (Power, nobs, noise, window) must be called first");
  const bgpu_params &p = h->p;
  const double inv_n = 1.0 / h->ncells;
  if (h->kick_on) {
    // only the evaluations that end in the prior + likelihood x pass can fuse the kick into their last store
    const bool alpt_exact = p.calc_h == BGPU_CALC_H_EXACT && h->phi1;
    if (p.likelihood == 3 || p.calc_h == 1 || alpt_exact) {
      const double a = h->kick_a;
      h->kick_on = false;
      gradient_device(h, d_s, h->grad);
      launch_axpy(d_out, h->grad, a, h->n, h->stream, h->stopflag);
      return;
    }
  }
  // the prior's multiplier (V/N)/P -- or zeros when only the likelihood force is wanted (mass types 2 / 3)
  const double *prior_mult = h->like_only ? h->zero_half : h->inv_power;
  if (!h->kspace_on) {
    h->fft.hooks = h->in_hooks;   // rows of the signal may still be arriving from the host
    r2c_plain(h, d_s, h->shat);
    h->fft.hooks = nullptr;
  }
  if (h->want_psi)  // as psi_device: 1/2 s . S^-1 s by Parseval
    launch_half_quadratic(h->shat, h->inv_power, h->N, h->nh, h->ncells, h->partials, h->dscal + S_PRIOR, h->stream);
  if (p.likelihood == 3) {
    // Gaussian random field (HMC.cc:159-160, gaussian_random_field.cpp:25-38): no structure formation at all,
    // gradpsi = IFFT[(V/N)/P s^] + (s - nobs)/sigma^2
    KOp lop;
    lop.kind = K_MULREAL;
    lop.real0 = prior_mult;
    ROp sop;
    sop.kind = R_SCALE;
    sop.a = inv_n;
    h->fft.c2r(h->shat, h->work, d_out, lop, sop);
    launch_grf_grad_add(d_out, d_s, h->nobs, h->noise, h->window, h->n, h->stream);
    if (h->out_hooks && h->out_hooks->after)
      for (int c = 0; c < h->out_hooks->chunks; ++c) h->out_hooks->after(h->out_hooks->ctx, c);
    return;
  }
  forward_from_shat(h, d_s, p.deltaQ_factor, p.rsd_model != 0, nullptr, nullptr, nullptr);
  LikeParams lp = h->like;
  lp.exact_sign = (p.calc_h == BGPU_CALC_H_EXACT) ? 1 : 0;
  // a gradient evaluation uses the residual only (-lnL belongs to psi()); delta_x itself is read again by
  // likelihood_calc_h alone (HMC_models_testing.cpp:25-50: gradfft / gradfindif of delta_x)
  launch_overdens_residual(lp, h->delta, h->dscal + S_SUMRHO, h->nobs, h->noise, h->window, h->resid, h->n,
                           h->ncells, h->partials, h->want_psi ? h->dscal + S_NLL : nullptr, h->stream,
                           /*keep_delta=*/p.calc_h == 0 || h->want_psi);

  // norm = -1 * deltaQ_factor * (D1 if correct_delta)   (HMC_models.cc:460-469)
  double norm = -1.0;
  norm *= p.deltaQ_factor;
  if (p.correct_delta) norm *= p.D1;

  if (p.calc_h == 1) {
    // h = r (HMC_models.cc:413-415): gradpsi = IFFT[(V/N)/P s^] + norm * r
    KOp lop;
    lop.kind = K_MULREAL;
    lop.real0 = prior_mult;
    ROp sop;
    sop.kind = R_SCALE;
    sop.a = inv_n;
    h->fft.c2r(h->shat, h->work, d_out, lop, sop);
    launch_axpy(d_out, h->resid, norm, h->n, h->stream);
    if (h->out_hooks && h->out_hooks->after)
      for (int c = 0; c < h->out_hooks->chunks; ++c) h->out_hooks->after(h->out_hooks->ctx, c);
    return;
  }

  if (p.calc_h == 0) {
    // likelihood_calc_h (HMC_models_testing.cpp:25-50): g_c = r * d_c(delta)
    if (p.likelihood == 1) r2c_plain(h, h->delta, h->dhat);  // gradfft shares one forward transform
    // gradfindif of delta_x (Poisson, poissonian.cpp:37-42) or of f(delta_x) (log-normal,
    // lognormal_independent.cpp:81-91).  On a slab the 4th-order stencil reaches 2 planes into the x neighbours:
    // their delta goes into the halo planes of the density tile (free once the deposits there have been added).
    const double *fd_field = h->delta;
    int fd_xo = 0;
    if (p.likelihood != 1) {
      const size_t plane = (size_t)h->N * h->N;
      if (h->G > 1) {
        constexpr int XH = 2;
        const int lo = (h->rank + h->G - 1) % h->G, hi = (h->rank + 1) % h->G;
        double *own = h->delta;
        ProfScope prof(KK_HALO, h->stream);
        h->comm->exchange2(own, lo, own + (size_t)(h->Ns - XH) * plane, hi, own + (size_t)h->Ns * plane, hi,
                           own - (size_t)XH * plane, lo, XH * plane, h->stream);
        fd_xo = XH;
      }
      if (p.likelihood == 2) {  // f goes to Psi_x (free once the density exists) / to its own halo-extended buffer
        if (h->G == 1) {
          launch_lognormal_f(h->delta, h->psi[0], h->n, p.rho_c, p.delta_min, h->stream);
          fd_field = h->psi[0];
        } else {
          launch_lognormal_f(h->delta - (size_t)fd_xo * plane, h->fext, h->n + (size_t)2 * fd_xo * plane, p.rho_c,
                             p.delta_min, h->stream);
          fd_field = h->fext + (size_t)fd_xo * plane;
        }
      }
    }
    if (p.likelihood == 1 && h->ubuf && h->fft.can_zround()) {
      // As below (shared x passes on both sides), and the two z passes that meet in real space around the product
      // r * d_c(delta) run as ONE pass on the row in shared memory (fft_tma.cuh fft_zround_tma): per component
      //   x (c = 0 only) -> y -> [c2r z, * r / N, r2c z] -> y -> x (accumulating; shared by c = 1, 2)
      KOp lop;
      lop.kind = K_GRAD;
      lop.comp = 0;
      lop.kfac = h->kfac;
      ROp mul;
      mul.kind = R_SCALE_MUL;
      mul.a = inv_n;
      mul.aux = h->resid;
      KOp none;
      h->fft.xpass(h->dhat, h->work, +1, lop, none);
      h->fft.ypass(h->work, h->work, +1, none, none);
      h->fft.zround(h->work, mul);
      h->fft.ypass(h->work, h->work, -1, none, none);
      KOp sop2;
      sop2.kind = K_INVLAP_SET;
      sop2.comp = 0;
      sop2.kfac = h->kfac;
      h->fft.xpass(h->work, h->acc, -1, none, sop2);
      lop.comp = K_COMP_UNIT;
      h->fft.xpass(h->dhat, h->dhat, +1, lop, none);   // in place: delta^ is not needed again
      for (int c = 1; c < 3; ++c) {
        KOp yl;
        yl.kind = K_MULK;
        yl.comp = c;
        yl.kfac = h->kfac;
        h->fft.ypass(h->dhat, h->work, +1, yl, none);
        h->fft.zround(h->work, mul);
        KOp ys;
        ys.kind = (c == 1) ? K_MULK_SET : K_MULK_ADD;
        ys.comp = c;
        ys.kfac = h->kfac;
        h->fft.ypass(h->work, h->ubuf, -1, none, ys);
      }
      sop2.kind = K_INVLAP_ADD;
      sop2.comp = K_COMP_UNIT;
      h->fft.xpass(h->ubuf, h->acc, -1, none, sop2);
    } else     if (p.likelihood == 1 && h->ubuf) {
      // shared x passes on both sides: d_c(delta) = IFFT_zy[k_c IFFT_x[-(Im, -Re) delta^]] for c = y, z, and the
      // products r d_y(delta), r d_z(delta) are summed after their forward y passes (see backproject)
      KOp lop;
      lop.kind = K_GRAD;
      lop.comp = 0;
      lop.kfac = h->kfac;
      ROp sop;
      sop.kind = R_SCALE_MUL;
      sop.a = inv_n;
      sop.aux = h->resid;
      h->fft.c2r(h->dhat, h->work, h->tmp, lop, sop);
      ROp lop2;
      lop2.kind = R_LOAD;
      KOp sop2;
      sop2.kind = K_INVLAP_SET;
      sop2.comp = 0;
      sop2.kfac = h->kfac;
      h->fft.r2c(h->tmp, h->work, h->acc, lop2, sop2);
      lop.comp = K_COMP_UNIT;
      h->fft.xpass(h->dhat, h->dhat, +1, lop, KOp{});   // in place: delta^ is not needed again
      for (int c = 1; c < 3; ++c) {
        KOp yl;
        yl.kind = K_MULK;
        yl.comp = c;
        yl.kfac = h->kfac;
        h->fft.c2r_yz(h->dhat, h->work, h->tmp, yl, sop);
        KOp ys;
        ys.kind = (c == 1) ? K_MULK_SET : K_MULK_ADD;
        ys.comp = c;
        ys.kfac = h->kfac;
        h->fft.r2c_zy(h->tmp, h->work, h->ubuf, lop2, ys);
      }
      sop2.kind = K_INVLAP_ADD;
      sop2.comp = K_COMP_UNIT;
      h->fft.xpass(h->ubuf, h->acc, -1, KOp{}, sop2);
    } else if (p.likelihood == 1 && h->fft.can_share_x_slab()) {
      // slab chain: d_y(delta) and d_z(delta) share one transposing x pass (as the displacement components do);
      // both products are formed before their forward transforms, so that the receive buffer the shared pass
      // left its result in is read twice before any transpose writes it again.  Psi_y / Psi_z are free by now.
      KOp lop;
      lop.kind = K_GRAD;
      lop.comp = 0;
      lop.kfac = h->kfac;
      ROp sop;
      sop.kind = R_SCALE_MUL;
      sop.a = inv_n;
      sop.aux = h->resid;
      h->fft.c2r(h->dhat, h->work, h->tmp, lop, sop);
      ROp lop2;
      lop2.kind = R_LOAD;
      KOp sop2;
      sop2.kfac = h->kfac;
      sop2.kind = K_INVLAP_SET;
      sop2.comp = 0;
      h->fft.r2c(h->tmp, h->work, h->acc, lop2, sop2);
      lop.comp = K_COMP_UNIT;
      h->fft.xpass_shared_inverse(h->dhat, lop);
      for (int c = 1; c < 3; ++c) {
        KOp yl;
        yl.kind = K_MULK;
        yl.comp = c;
        yl.kfac = h->kfac;
        h->fft.c2r_yz_shared(h->work, h->psi[c], yl, sop);
      }
      for (int c = 1; c < 3; ++c) {
        sop2.kind = K_INVLAP_ADD;
        sop2.comp = c;
        h->fft.r2c(h->psi[c], h->work, h->acc, lop2, sop2);
      }
    } else
    for (int c = 0; c < 3; ++c) {
      if (p.likelihood == 1) {
        KOp lop;
        lop.kind = K_GRAD;
        lop.comp = c;
        lop.kfac = h->kfac;
        ROp sop;
        sop.kind = R_SCALE_MUL;
        sop.a = inv_n;
        sop.aux = h->resid;
        h->fft.c2r(h->dhat, h->work, h->tmp, lop, sop);
      } else {
        launch_findif_product(fd_field, h->resid, h->tmp, h->N, h->Ns, fd_xo, p.L1, c, h->stream);
      }
      ROp lop2;
      lop2.kind = R_LOAD;
      KOp sop2;
      sop2.kfac = h->kfac;
      if (h->ubuf && c > 0) {
        // shared x pass of the back-projection (see backproject): the y and z products meet after their y passes
        sop2.kind = (c == 1) ? K_MULK_SET : K_MULK_ADD;
        sop2.comp = c;
        h->fft.r2c_zy(h->tmp, h->work, h->ubuf, lop2, sop2);
        if (c == 2) {
          sop2.kind = K_INVLAP_ADD;
          sop2.comp = K_COMP_UNIT;
          h->fft.xpass(h->ubuf, h->acc, -1, KOp{}, sop2);
        }
        continue;
      }
      sop2.kind = (c == 0) ? K_INVLAP_SET : K_INVLAP_ADD;
      sop2.comp = c;
      h->fft.r2c(h->tmp, h->work, h->acc, lop2, sop2);
    }
  } else if (p.calc_h == BGPU_CALC_H_EXACT && h->phi1) {
    // exact adjoint of Lag2Eul_non_zeldovich (new; oracle: non_zeldovich_adjoint, FD-validated):
    //   W_c = CB^T V_c;  A^ = sum_c T_c W^_c;  u_lpt = IFFT[K A^], u_sc = IFFT[(1 - K) A^];
    //   grad = dQ (D1 u_lpt - D2 Poisson[sum_ab FD_a FD_b (dm2v/dL_ab u_lpt)] + theta_SC' u_sc)
    GridGeom g = h->geom;
    g.cellbound = 1;
    launch_gather_adjoint_to(g, h->psi[0], h->psi[1], h->psi[2], h->xa, h->xb, h->tmp, h->resid, h->stream);
    double *V[3] = {h->xa, h->xb, h->tmp};
    for (int c = 0; c < 3; ++c) launch_cellbound_transpose(V[c], h->psi[c], h->N, h->stream);
    backproject(h, h->psi);
    ROp unit;
    unit.kind = R_SCALE;
    unit.a = inv_n;
    KOp kg;
    kg.kfac = h->kfac;
    kg.a = p.slength;
    kg.kind = K_GAUSS;
    h->fft.c2r(h->acc, h->work, h->xa, kg, unit);            // u_lpt
    kg.kind = K_ONE_MINUS_GAUSS;
    h->fft.c2r(h->acc, h->work, h->xb, kg, unit);            // u_sc
    double *P6[6] = {h->psi[0], h->psi[1], h->psi[2], h->resid, h->tmp, h->xc};
    launch_lpt2_adjoint_coef(h->phi1, h->xa, P6, h->N, p.L1, h->stream);
    launch_lpt2_adjoint_div(P6, h->phi1, h->N, p.L1, h->stream);   // G over phi^(1), which is done
    r2c_plain(h, h->phi1, h->dhat);
    KOp pois;
    pois.kind = K_NEGINVK2;
    pois.a = 1.0;
    pois.kfac = h->kfac;
    h->fft.c2r(h->dhat, h->work, h->phi1, pois, unit);          // q = Poisson[G]
    launch_alpt_adjoint_combine(h->tmp, h->xa, h->phi1, h->xb, d_s, h->n, p.deltaQ_factor, p.D1, p.D2, h->stream);
    // gradpsi = IFFT[(V/N)/P s^] + that
    KOp lop;
    lop.kind = K_MULREAL;
    lop.real0 = prior_mult;
    h->fft.c2r(h->shat, h->work, d_out, lop, unit);
    launch_axpy(d_out, h->tmp, 1.0, h->n, h->stream);
    if (h->out_hooks && h->out_hooks->after)
      for (int c = 0; c < h->out_hooks->chunks; ++c) h->out_hooks->after(h->out_hooks->ctx, c);
    return;
  } else if (p.calc_h == 3) {
    // likelihood_calc_V_SPH_fourier_TSC (HMC_models_testing.cpp:54-188): V_c(p) = TSC-interpolation at the particle of
    // IFFT[i k_c h W^(k) r^(k)], z times (1 + f) under RSD; then the back-projection of calc_h = 2.  The residual and
    // the density are free once r^ exists and take V_x, V_y; V_z goes over Psi_z last (positions come from Psi).
    r2c_plain(h, h->resid, h->dhat);
    ROp unit;
    unit.kind = R_SCALE;
    unit.a = inv_n;
    double *V[3] = {h->resid, h->delta, h->psi[2]};
    for (int c = 0; c < 3; ++c) {
      launch_sph_fourier_comp(h->dhat, h->sph_hW, h->acc, h->N, h->kfac, c, h->stream);
      h->fft.c2r(h->acc, h->acc, h->tmp, KOp{}, unit);
      const double fz = (c == 2 && p.rsd_model) ? h->geom.fgrow : 0.0;   // out_z += f1 * out_z (:176-186)
      launch_interp_tsc(h->geom, h->psi[0], h->psi[1], h->psi[2], h->tmp, V[c], fz, h->stream);
    }
    backproject(h, V);
  } else {
    // exact adjoint: V = gather(r) in place over Psi, then the same back-projection
    if (p.calc_h == 2 && h->G == 1)
      launch_gather_sph(h->geom, h->psi[0], h->psi[1], h->psi[2], h->resid, h->sph_kmax, h->sph_R,
                        p.rho_c * (p.L1 * p.L2 * p.L3) / h->ncells, h->stream);
    else if (h->G == 1)
      launch_gather_adjoint(h->geom, h->psi[0], h->psi[1], h->psi[2], h->resid, h->stream);
    else {
      // slab: a particle of mine may sit in cells of the neighbouring slabs (the same halo the scatter used),
      // so the gather reads the residual from a tile extended by H planes of each x neighbour
      const int H = h->H_cur, lo = (h->rank + h->G - 1) % h->G, hi = (h->rank + 1) % h->G;
      const size_t plane = (size_t)h->N * h->N, cnt = (size_t)H * plane;
      double *own = h->resid_ext + cnt;
      BGPU_CUDA(cudaMemcpyAsync(own, h->resid, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      {
        ProfScope prof(KK_HALO, h->stream);
        // my first planes are rank-1's upper halo, my last planes rank+1's lower halo
        h->comm->exchange2(own, lo, own + (size_t)(h->Ns - H) * plane, hi, own + (size_t)h->Ns * plane, hi, h->resid_ext, lo,
                           cnt, h->stream);
      }
      GridGeom g = h->geom;
      g.H = H;
      if (p.calc_h == 2)
        launch_gather_sph(g, h->psi[0], h->psi[1], h->psi[2], h->resid_ext, h->sph_kmax, h->sph_R,
                          p.rho_c * (p.L1 * p.L2 * p.L3) / h->ncells, h->stream);
      else
        launch_gather_adjoint(g, h->psi[0], h->psi[1], h->psi[2], h->resid_ext, h->stream);
    }
    backproject(h, h->psi);
  }
  if (h->kspace_on) {  // p^ += kick_a ((V/N)/P s^ + norm h^): the kick without gradpsi's inverse transform
    launch_kspace_kick(h->phat, h->shat, h->acc, prior_mult, h->kick_a, norm, h->N, h->nh, h->ncells, h->partials,
                       h->dscal + S_P0, h->stream, h->stopflag);
    allreduce_scalar(h, S_P0);   // a slab holds its modes' share of momenta[0]; every rank then tests the same number
    return;
  }
  // gradpsi = IFFT[(V/N)/P s^ + norm * h^]
  ROp sop;
  sop.kind = R_SCALE;
  sop.a = inv_n;
  if (h->kick_on) {  // d_out += kick_a * gradpsi, in the store of the last z pass
    sop.kind = R_AXPY;
    sop.a = h->kick_a * inv_n;
    sop.skip = h->stopflag;
    h->kick_on = false;
  }
  if (h->G > 1 && (h->N >= 512 || h->fft.force_generic)) {
    // the TMA-staged x pass cannot hold both operand tiles at this size: combine in a pass of its own
    launch_kfinal_combine(h->shat, prior_mult, h->acc, h->acc, norm, h->N, h->nh, h->stream);
    h->fft.hooks = h->out_hooks;
    h->fft.c2r(h->acc, h->work, d_out, KOp{}, sop);
    h->fft.hooks = nullptr;
    return;
  }
  KOp lop;
  lop.kind = K_FINAL;
  lop.a = norm;
  lop.real0 = prior_mult;
  lop.cplx0 = h->acc;
  h->fft.hooks = h->out_hooks;  // rows of the gradient leave for the host as the last z pass produces them
  h->fft.c2r(h->shat, h->work, d_out, lop, sop);
  h->fft.hooks = nullptr;
}

// psi (HMC.cc:124-143): prior 1/2 s.S^-1 s (gaussian.cpp:20-35) and -lnL
// (gaussian_independent.cpp:51-92 / poissonian.cpp:44-74); leaves deltaX in h->delta
void psi_device(bgpu_handle *h, const double *d_s) {
  require(h->have_power && h->have_obs, "bgpu: bgpu_set_static (Power, nobs, noise, window) must be called first");
  const bgpu_params &p = h->p;
  const double inv_n = 1.0 / h->ncells;
  r2c_plain(h, d_s, h->shat);
  if (h->parseval) {
    // 1/2 s . S^-1 s = (1/2N) sum_k w_k (V/N)/P_k |s^_k|^2: no inverse transform (kernels.cu HalfQuadF)
    launch_half_quadratic(h->shat, h->inv_power, h->N, h->nh, h->ncells, h->partials, h->dscal + S_PRIOR, h->stream);
  } else {
    KOp lop;
    lop.kind = K_MULREAL;
    lop.real0 = h->inv_power;
    ROp sop;
    sop.kind = R_SCALE;
    sop.a = inv_n;
    h->fft.c2r(h->shat, h->work, h->tmp, lop, sop);
    launch_half_dot(d_s, h->tmp, h->n, h->partials, h->dscal + S_PRIOR, h->stream);
  }
  allreduce_scalar(h, S_PRIOR);

  if (p.likelihood == 3) {  // gaussian_random_field.cpp:40-52: -lnL on the Lagrangian field itself
    launch_grf_nll(d_s, h->nobs, h->noise, h->window, h->n, h->partials, h->dscal + S_NLL, h->stream);
    allreduce_scalar(h, S_NLL);
    return;
  }
  const bool gauss = p.likelihood == 1;
  forward_from_shat(h, d_s, gauss ? p.deltaQ_factor : 1.0, gauss ? (p.rsd_model != 0) : false, nullptr, nullptr, nullptr);
  LikeParams lp = h->like;
  lp.exact_sign = 0;
  launch_overdens_residual(lp, h->delta, h->dscal + S_SUMRHO, h->nobs, h->noise, h->window, nullptr, h->n,
                           h->ncells, h->partials, h->dscal + S_NLL, h->stream);
  allreduce_scalar(h, S_NLL);
}

// M^-1 p (HMC.cc:298-327, :69-99) -> h->tmp ; returns false if there is no Fourier part
void apply_inv_mass_fs(bgpu_handle *h, const double *d_p) {
  require(h->have_mass, "bgpu: bgpu_set_mass or bgpu_hamiltonian_mass must be called first");
  const double inv_n = 1.0 / h->ncells;
  r2c_plain(h, d_p, h->work);
  KOp lop;
  lop.kind = K_MULREAL;
  lop.real0 = h->inv_mass;
  ROp sop;
  sop.kind = R_SCALE;
  sop.a = inv_n;
  h->fft.c2r(h->work, h->work, h->tmp, lop, sop);
}

void kinetic_device(bgpu_handle *h, const double *d_p) {
  require(h->have_mass, "bgpu: bgpu_set_mass or bgpu_hamiltonian_mass must be called first");
  if (h->mass_fs && !h->mass_rs && h->parseval) {
    // 1/2 p . M^-1 p in k-space (Parseval): one forward transform and a half-grid reduction
    r2c_plain(h, d_p, h->work);
    launch_half_quadratic(h->work, h->inv_mass, h->N, h->nh, h->ncells, h->partials, h->dscal + S_KIN, h->stream);
    allreduce_scalar(h, S_KIN);
    return;
  }
  if (h->mass_fs) apply_inv_mass_fs(h, d_p);
  launch_kinetic(d_p, h->mass_fs ? h->tmp : nullptr, h->mass_rs ? h->mass_r : nullptr, h->n, h->partials,
                 h->dscal + S_KIN, h->stream);
  allreduce_scalar(h, S_KIN);
}

// p += a * gradpsi(s): the kick of the leapfrog, fused into the gradient's last store where the path allows
void kick_device(bgpu_handle *h, const double *d_s, double *d_p, double a) {
  h->kick_on = true;
  h->kick_a = a;
  try {
    gradient_device(h, d_s, d_p);
  } catch (...) {
    h->kick_on = false;
    throw;
  }
  h->kick_on = false;
}

// Hamiltonian_EoM (HMC.cc:251-369) after the RNG draws, in place on device.
//
// Fused form (default): the two half kicks that meet between steps are one kick p -= eps * gradpsi (half kicks only
// at the two ends, :293-294 / :351-352), applied by the gradient's last z pass (bulk f64 reduce-add); the drift
// s += eps * M^-1 p (:338-339) is the store of M^-1 p's last z pass.  No host round trip inside the trajectory: the
// run-away test (:360-364) runs on the device (kernels.cu runaway_guard_kernel: the value the reference tests is the
// midpoint of momenta[0] before and after a merged kick) and raises a flag that every later update of the trajectory
// honours; a slab chain shares rank 0's verdict with a one-element all-reduce per step.  The host reads the flag
// once, after the trajectory.  The merged kick rounds p - eps g once where the reference rounds two half kicks: a
// relative 1e-16 per step.
// The trajectory in k-space: possible when both updates are diagonal there (Fourier-space mass only) and the
// evaluation reads s^ alone and ends in the k-space sum (Zel'dovich model; not the GRF likelihood, calc_h = 1 or the
// likelihood-only force) -- BASELINE.json's ZA + CIC configurations.  Per step it saves the forward transform of s,
// the inverse transform of gradpsi and the transform pair around M^-1 p: 12 of 36 passes at calc_h = 0.  s and p
// are transformed once at each end of the trajectory (a round trip through the FFT: ~1e-16 relative).
static bool kspace_leapfrog_applies(const bgpu_handle *h) {
  const bgpu_params &p = h->p;
  const bool zeldovich = (p.sfmodel == 1 || p.rsd_model);
  // (a slab chain as well: the updates are local to a rank's k-space slab, momenta[0] is one all-reduced scalar)
  return h->kspace_lf && h->fused_leapfrog && h->mass_fs && !h->mass_rs && zeldovich && !h->like_only &&
         !(p.likelihood == 3 || p.calc_h == 1);
}

static void kspace_kick(bgpu_handle *h, double a) {
  h->kspace_on = true;
  h->kick_a = a;
  try {
    gradient_device(h, nullptr, nullptr);
  } catch (...) {
    h->kspace_on = false;
    throw;
  }
  h->kspace_on = false;
}

static void leapfrog_kspace(bgpu_handle *h, double *d_s, double *d_p, uint64_t Neps, double eps) {
  require(h->have_mass, "bgpu: bgpu_set_mass or bgpu_hamiltonian_mass must be called first");
  if (!h->phat) dalloc(h->phat, h->nh);
  double *guard = h->dscal + S_STOP;
  // kernels.cu runaway_guard_kernel on momenta[0] = the sum the kick returns in dscal[S_P0]
  auto test = [&](int step, int mode) { launch_runaway_guard(h->dscal + S_P0, guard, h->stopflag, step, mode, h->stream); };
  BGPU_CUDA(cudaMemsetAsync(h->stopflag, 0, sizeof(int), h->stream));
  BGPU_CUDA(cudaMemsetAsync(guard, 0, 2 * sizeof(double), h->stream));
  const bool psi_wanted = h->psi_at_end && Neps > 0 && h->p.likelihood == 1 && h->parseval && h->G == 1;
  h->psi_at_end = false;
  h->psi_from_kick = false;
  r2c_plain(h, d_s, h->shat);
  r2c_plain(h, d_p, h->phat);
  kspace_kick(h, -(0.5 * eps));                                           // HMC.cc:293-294
  test(0, 0);
  for (uint64_t jj = 0; jj < Neps; ++jj) {
    launch_kspace_drift(h->shat, h->phat, h->inv_mass, eps, h->N, h->nh, h->stream, h->stopflag);   // :298-339
    const bool last = jj + 1 == Neps;
    h->want_psi = last && psi_wanted;
    kspace_kick(h, last ? -(0.5 * eps) : -eps);                           // :343-352 (+ :293-294 of the next step)
    h->want_psi = false;
    test((int)(jj + 1 < 0x7fffffff ? jj + 1 : 0x7fffffff), last ? 2 : 1);
  }
  // the one host round trip of the trajectory (see the fused real-space form below for the undo)
  BGPU_CUDA(cudaMemcpyAsync(h->hflag2, h->stopflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  const int stopped = *h->hflag2;
  if (stopped > 0 && (uint64_t)stopped < Neps) {
    BGPU_CUDA(cudaMemsetAsync(h->stopflag, 0, sizeof(int), h->stream));
    kspace_kick(h, +(0.5 * eps));
  }
  h->psi_from_kick = psi_wanted && stopped == 0;   // the last kick's evaluation was at the state being returned
  ROp back;
  back.kind = R_SCALE;
  back.a = 1.0 / h->ncells;
  h->fft.c2r(h->shat, h->work, d_s, KOp{}, back);
  h->fft.c2r(h->phat, h->work, d_p, KOp{}, back);
}

void leapfrog_device(bgpu_handle *h, double *d_s, double *d_p, uint64_t Neps, double eps) {
  if (kspace_leapfrog_applies(h)) {
    leapfrog_kspace(h, d_s, d_p, Neps, eps);
    return;
  }
  if (h->fused_leapfrog) {
    const int *skip = h->stopflag;
    static_assert(S_P0PREV == S_STOP + 1, "the guard's two scalars are adjacent");
    double *guard = h->dscal + S_STOP;
    // the run-away test after the kick that ends step `step`: rank 0 owns momenta[0], a slab shares the verdict
    auto test = [&](int step, int mode) {
      if (h->rank == 0) launch_runaway_guard(d_p, guard, h->G == 1 ? h->stopflag : nullptr, step, mode, h->stream);
      if (h->G > 1 && mode != 0) {
        if (h->rank != 0) BGPU_CUDA(cudaMemsetAsync(guard, 0, sizeof(double), h->stream));
        h->comm->all_reduce_max(guard, 1, h->stream);
        launch_runaway_apply(guard, h->stopflag, h->stream);
      }
    };
    BGPU_CUDA(cudaMemsetAsync(h->stopflag, 0, sizeof(int), h->stream));
    BGPU_CUDA(cudaMemsetAsync(guard, 0, 2 * sizeof(double), h->stream));
    kick_device(h, d_s, d_p, -(0.5 * eps));
    test(0, 0);
    for (uint64_t jj = 0; jj < Neps; ++jj) {
      if (h->mass_fs) {
        require(h->have_mass, "bgpu: bgpu_set_mass or bgpu_hamiltonian_mass must be called first");
        r2c_plain(h, d_p, h->work);
        KOp lop;
        lop.kind = K_MULREAL;
        lop.real0 = h->inv_mass;
        ROp sop;
        sop.kind = R_AXPY;
        sop.a = eps / h->ncells;
        sop.skip = skip;
        h->fft.c2r(h->work, h->work, d_s, lop, sop);
      }
      if (h->mass_rs) launch_axpy_div(d_s, d_p, h->mass_r, eps, h->n, h->stream, skip);
      const bool last = jj + 1 == Neps;
      kick_device(h, d_s, d_p, last ? -(0.5 * eps) : -eps);
      test((int)(jj + 1 < 0x7fffffff ? jj + 1 : 0x7fffffff), last ? 2 : 1);
    }
    // the one host round trip of the trajectory.  Stopped before the last step, the state is frozen at the
    // reference's (s, p) plus the half kick the merged kick had already added for the step that never ran:
    // take it back with one more gradient at the same s.
    BGPU_CUDA(cudaMemcpyAsync(h->hflag2, h->stopflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    sync(h);
    const int stopped = *h->hflag2;
    if (stopped > 0 && (uint64_t)stopped < Neps) {
      BGPU_CUDA(cudaMemsetAsync(h->stopflag, 0, sizeof(int), h->stream));
      kick_device(h, d_s, d_p, +(0.5 * eps));
    }
    return;
  }
  gradient_device(h, d_s, h->grad);
  for (uint64_t jj = 0; jj < Neps; ++jj) {
    launch_axpy(d_p, h->grad, -(0.5 * eps), h->n, h->stream);   // :293-294
    if (h->mass_fs) {
      apply_inv_mass_fs(h, d_p);                                 // :310
      launch_axpy(d_s, h->tmp, eps, h->n, h->stream);            // :338-339
    }
    if (h->mass_rs) launch_axpy_div(d_s, d_p, h->mass_r, eps, h->n, h->stream);  // :317-327
    gradient_device(h, d_s, h->grad);                            // :343-344
    launch_axpy(d_p, h->grad, -(0.5 * eps), h->n, h->stream);   // :351-352
    // :360-364 -- stop a trajectory whose momentum has run away
    if (h->G > 1) {  // every rank must take the same decision: rank 0 owns momenta[0]
      launch_max_abs(d_p, 1, h->partials, h->dscal + S_P0, h->stream);
      if (h->rank != 0) BGPU_CUDA(cudaMemsetAsync(h->dscal + S_P0, 0, sizeof(double), h->stream));
      h->comm->all_reduce_max(h->dscal + S_P0, 1, h->stream);
      BGPU_CUDA(cudaMemcpyAsync(h->hscal + S_P0, h->dscal + S_P0, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    } else {
      BGPU_CUDA(cudaMemcpyAsync(h->hscal + S_P0, d_p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    sync(h);
    if (std::fabs(h->hscal[S_P0]) > 1e50) break;
  }
}

void update_inverse(bgpu_handle *h, const double *full, double *half) {
  if (h->G == 1) {
    launch_inverse_spectrum(full, half, h->N, h->normFS, h->stream);
    return;
  }
  // the multiplier is needed at (all x, my y rows): transpose it like a k-space array
  launch_inverse_spectrum_pack(full, h->sendbuf, h->N, h->Ns, h->normFS, h->stream);
  h->comm->all_to_all(h->sendbuf, h->recvbuf, (size_t)h->Ns * h->Ns * (h->N / 2 + 1) * 2, h->stream);
  launch_inverse_spectrum_unpack(h->recvbuf, half, h->N, h->Ns, h->stream);
  h->comm->barrier(h->stream);  // the receive buffer is free again before any peer's next transform writes it
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
#define BGPU_TRY try {
#define BGPU_CATCH                      \
  }                                     \
  catch (const std::exception &e) {     \
    g_last_error = e.what();            \
    return 1;                           \
  }                                     \
  catch (...) {                         \
    g_last_error = "bgpu: unknown error"; \
    return 1;                           \
  }                                     \
  return 0;

extern "C" {

int bgpu_abi_version(void) { return BGPU_ABI_VERSION; }
const char *bgpu_last_error(void) { return g_last_error.c_str(); }
uint64_t bgpu_kernel_launches(void) { return g_kernel_launches.load(); }

int bgpu_profile_begin(void) {
  g_prof.used = 0;
  g_prof.on = true;
  return 0;
}

int bgpu_profile_end(double *ms_per_kind, uint64_t *launches_per_kind, int nkinds) {
  BGPU_TRY
  g_prof.on = false;
  for (int k = 0; k < nkinds; ++k) {
    ms_per_kind[k] = 0.0;
    launches_per_kind[k] = 0;
  }
  BGPU_CUDA(cudaDeviceSynchronize());
  for (size_t i = 0; i < g_prof.used; ++i) {
    float ms = 0.f;
    BGPU_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
    const int k = g_prof.kind[i];
    if (k < nkinds) {
      ms_per_kind[k] += ms;
      launches_per_kind[k] += 1;
    }
  }
  g_prof.used = 0;
  BGPU_CATCH
}

const char *bgpu_profile_kind_name(int kind) {
  static const char *names[KK_COUNT] = {"fft_strided_pass_y", "fft_r2c_zpass", "fft_c2r_zpass", "scatter",
                                        "gather_adjoint", "overdens_residual", "reduce", "stream", "colour_momenta",
                                        "fft_strided_pass_x", "all_to_all", "halo_exchange", "fft_zy_fused_r2c",
                                        "fft_zy_fused_c2r", "fft_z_roundtrip"};
  return (kind >= 0 && kind < KK_COUNT) ? names[kind] : "?";
}

void bgpu_default_params(bgpu_params *p) {
  std::memset(p, 0, sizeof(*p));
  p->N1 = p->N2 = p->N3 = 64;               // data/input.par:117-125
  p->L1 = p->L2 = p->L3 = 200.0;
  p->xobs = p->yobs = p->zobs = 90.0;
  p->planepar = 1;
  p->periodic = 1;
  p->masskernel = 1;
  p->likelihood = 1;
  p->sfmodel = 1;
  p->rsd_model = 0;
  p->calc_h = 0;
  p->mass_type = 1;
  p->D1 = 1.0;
  p->D2 = -3. / 7. * std::pow(0.272, -1. / 143.);  // init_par.cc:526-528 at z = 0
  p->slength = 4.0;                          // data/input.par:121
  p->particle_kernel_h_rel = 1.0;            // data/input.par:134
  p->ascale = 1.0;
  p->OM = 0.272;                            // init_par.cc:38,480-483 (cmbcosm = 3)
  p->OL = 0.728;
  p->rho_c = p->biasP = p->biasE = 1.0;     // init_par.cc:574-578
  p->deltaQ_factor = 1.0;
  p->correct_delta = 1;
  p->mass_factor = 1.0;
  p->div_dH_by_N = 0;
  p->device = 0;
  p->delta_min = -0.999;  // data/input.par:51
  p->N_bin = 200;         // data/input.par:129
}

static int create_impl(const bgpu_params *p, int rank, int nranks, const void *nccl_id, bgpu_handle **out,
                       LocalGroup *local = nullptr) {
  bgpu_handle *h = nullptr;
  BGPU_TRY
  require(p && out, "bgpu_create: null argument");
  *out = nullptr;
  validate(*p);
  if (nranks > 1) {
    // the slab path runs on the TMA-staged FFT passes and covers what needs no particle
    // data across slabs beyond the density halo
    require(p->N1 == 128 || p->N1 == 256 || p->N1 == 512 || p->N1 == 1024,
            "bgpu_slab_create: the slab-decomposed transform supports N = 128, 256, 512, 1024");
    require(rank >= 0 && rank < nranks && p->N1 % nranks == 0 && p->N1 / nranks >= 8,
            "bgpu_slab_create: N1 must be a multiple of the number of ranks, at least 8 planes per rank");
    require(!(p->calc_h == BGPU_CALC_H_EXACT && p->sfmodel != 1 && !p->rsd_model),
            "bgpu_slab_create: the exact adjoint of the 2LPT/ALPT model is not built for slabs yet");
    require(p->calc_h == 0 || p->calc_h == 1 || p->calc_h == 2 || p->calc_h == BGPU_CALC_H_EXACT,
            "bgpu_slab_create: calc_h must be 0, 1, 2 or 4 (the Fourier / TSC variant of the SPH adjoint, calc_h 3, is "
            "not built for slabs)");
    require(p->sfmodel == 1 || p->rsd_model || p->N1 / nranks >= 8,
            "bgpu_slab_create: the 2LPT/ALPT model needs at least 8 planes per rank (4-plane stencil halo)");
    require(nccl_id != nullptr || local != nullptr, "bgpu_slab_create: a NCCL unique id is required");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw std::runtime_error(std::string("bgpu: no CUDA device available (") + cudaGetErrorString(e) +
                             "); this path has no CPU fallback");
  require(p->device >= 0 && p->device < ndev, "bgpu: device ordinal out of range");
  BGPU_CUDA(cudaSetDevice(p->device));
  h = new bgpu_handle;
  h->p = *p;
  h->N = p->N1;
  h->G = nranks;
  h->rank = rank;
  h->Ns = h->N / nranks;
  h->x0 = rank * h->Ns;
  h->ncells = (double)h->N * h->N * h->N;
  h->n = (size_t)h->Ns * h->N * h->N;
  h->nh = (size_t)h->Ns * h->N * (h->N / 2 + 1);
  h->nhp = (size_t)h->Ns * h->N * (h->N / 2 + 2);
  BGPU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  BGPU_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (int c = 0; c < bgpu_handle::kChunks; ++c) {
    BGPU_CUDA(cudaEventCreateWithFlags(&h->ev_up[c], cudaEventDisableTiming));
    BGPU_CUDA(cudaEventCreateWithFlags(&h->ev_dn[c], cudaEventDisableTiming));
  }
  if (nranks > 1) {
    if (local) h->comm = new LocalComm(local, rank, nranks);
    else h->comm = new NcclComm(nccl_id, rank, nranks);
    dalloc(h->sendbuf, h->nh);
    dalloc(h->recvbuf, 2 * h->nh);   // two receive buffers, alternating between transforms
    h->Hmax = h->Ns < 24 ? h->Ns : 24;
    dalloc(h->halo_recv, (size_t)2 * h->Hmax * h->N * h->N);
    if (p->calc_h == BGPU_CALC_H_EXACT || p->calc_h == 2)
      dalloc(h->resid_ext, (size_t)(h->Ns + 2 * h->Hmax) * h->N * h->N);
    if (p->calc_h == 0 && p->likelihood == 2) dalloc(h->fext, (size_t)(h->Ns + 4) * h->N * h->N);
    BGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&h->dflag), sizeof(int)));
    BGPU_CUDA(cudaMemsetAsync(h->dflag, 0, sizeof(int), h->stream));
    BGPU_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h->hflag), sizeof(int)));
    *h->hflag = 0;
    h->fft.G = nranks;
    h->fft.rank = rank;
    h->fft.Ns = h->Ns;
    h->fft.comm = h->comm;
    h->fft.sendbuf = h->sendbuf;
    h->fft.recvbuf = h->recvbuf;
    const char *e2 = std::getenv("BGPU_SLAB_P2P");
    if (!(e2 && e2[0] == '0') && local) {
      // ranks of one process on one device: every rank's receive buffer is directly addressable
      require(nranks <= 8, "bgpu_slab_create: at most 8 ranks");
      void *all[8] = {};
      local_group_exchange_ptr(local, rank, h->recvbuf, all);
      for (int r = 0; r < nranks; ++r) {
        h->fft.peer_recv[0][r] = static_cast<double2 *>(all[r]);
        h->fft.peer_recv[1][r] = static_cast<double2 *>(all[r]) + h->nh;
      }
      h->fft.p2p = true;
    } else if (!(e2 && e2[0] == '0')) {
      // fused transpose: map every rank's receive buffers into this process (CUDA IPC over NVLink peer
      // access); the 64-byte handles travel through the NCCL communicator itself
      static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
      require(nranks <= 8, "bgpu_slab_create: at most 8 ranks (one box)");
      cudaIpcMemHandle_t mine;
      BGPU_CUDA(cudaIpcGetMemHandle(&mine, h->recvbuf));
      std::vector<cudaIpcMemHandle_t> out_h(nranks, mine), in_h(nranks);
      double *d_x = nullptr;
      dalloc(d_x, (size_t)2 * nranks * 8);
      BGPU_CUDA(cudaMemcpyAsync(d_x, out_h.data(), (size_t)nranks * 64, cudaMemcpyHostToDevice, h->stream));
      h->comm->all_to_all(d_x, d_x + (size_t)nranks * 8, 8, h->stream);
      BGPU_CUDA(cudaMemcpyAsync(in_h.data(), d_x + (size_t)nranks * 8, (size_t)nranks * 64, cudaMemcpyDeviceToHost,
                                h->stream));
      BGPU_CUDA(cudaStreamSynchronize(h->stream));
      cudaFree(d_x);
      for (int r = 0; r < nranks; ++r) {
        void *base = h->recvbuf;
        if (r != rank) {
          BGPU_CUDA(cudaIpcOpenMemHandle(&base, in_h[r], cudaIpcMemLazyEnablePeerAccess));
          h->peer_base[r] = base;
        }
        h->fft.peer_recv[0][r] = static_cast<double2 *>(base);
        h->fft.peer_recv[1][r] = static_cast<double2 *>(base) + h->nh;
      }
      h->fft.p2p = true;
    }
  }
  h->fft.init(h->N, h->stream);
  h->kfac = 2. * M_PI / p->L1;                                      // scale_space.cpp:42
  h->normFS = (p->L1 * p->L2 * p->L3) / h->ncells;                  // HMC_help.cc:26
  h->mass_fs = (p->mass_type >= 1 && p->mass_type <= 4);            // struct_hamil.h:276-296
  h->mass_rs = (p->mass_type == 0);

  GridGeom &g = h->geom;
  g.N = h->N;
  g.L = p->L1;
  g.d = p->L1 / (double)h->N;                                       // init_par.cc:245
  g.min1 = p->xllc; g.min2 = p->yllc; g.min3 = p->zllc;
  g.masskernel = p->masskernel;
  g.rsd = p->rsd_model ? 1 : 0;
  g.fgrow = fgrow1(p->ascale, p->OM, p->OL);
  g.cpecvel = g.fgrow * 100. * E_Hubble_a(p->ascale, p->OM, p->OL) * p->ascale;  // cosmo.cc:232
  {
    const double OC = 1. - p->OM - p->OL;                            // rsd.cc:27-28
    const double Hub = 100. * std::sqrt(p->OM / p->ascale / p->ascale / p->ascale + p->OL + OC / p->ascale / p->ascale);
    g.v_norm = 1. / Hub / p->ascale;                                 // rsd.cc:39
  }
  g.x0 = h->x0;
  g.Ns = h->Ns;
  g.H = 0;
  g.flag = h->dflag;
  g.cellbound = 0;
  g.cb_lo = nullptr;
  g.sph_h = p->particle_kernel_h_rel * g.d;
  {
    const char *sw = std::getenv("BGPU_SWEEP");  // 0 = first-generation particle-per-thread kernels; n > 1 = segment length
    g.sweep = sw ? std::atoi(sw) : 1;
    const char *ln = std::getenv("BGPU_LEAN");
    g.lean = ln ? std::atoi(ln) : 1;   // 21: the gather with 2 CTAs per SM and no read-ahead (particles_sweep.cu)
  }
  if (p->masskernel == 3) {
    // SPH_kernel_3D_cells + _hull_1 (SPH_kernel.cpp:62-139)
    const double d = g.d, reach_len = 2. * g.sph_h;
    const int R = (int)(reach_len / d) + 1, W = 2 * R + 1;
    std::vector<int> kmax((size_t)W * W, -1);
    for (int i1 = -R; i1 <= R; ++i1)
      for (int i2 = -R; i2 <= R; ++i2)
        for (int i3 = -R; i3 <= R; ++i3) {
          const double dx = (std::abs((double)i1) - 0.5) * d, dy = (std::abs((double)i2) - 0.5) * d,
                       dz = (std::abs((double)i3) - 0.5) * d;
          if (dx * dx + dy * dy + dz * dz <= reach_len * reach_len) {
            int &km = kmax[(size_t)(i1 + R) * W + (i2 + R)];
            if (std::abs(i3) > km) km = std::abs(i3);
          }
        }
    h->sph_R = R;
    BGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&h->sph_kmax), kmax.size() * sizeof(int)));
    BGPU_CUDA(cudaMemcpy(h->sph_kmax, kmax.data(), kmax.size() * sizeof(int), cudaMemcpyHostToDevice));
    // the static column lists both SPH kernels walk (particles_sph.cu); BGPU_SPH_COLS=0: the general kernels
    const char *sc = std::getenv("BGPU_SPH_COLS");
    if (!(sc && sc[0] == '0')) g.sph_cols = sph_columns_create(g, kmax.data(), R);
  }
  if (p->calc_h == 3) {
    // h * SPH_kernel_F (HMC_models_testing.cpp:62-105), evaluated on the HOST, operation by operation, with the C
    // library the reference itself calls: the numerator 3 + cos 2k - k sin k + cos k (k sin k - 4) cancels to
    // ~k^6 / 240, so at the grid's small k its value is rounding noise that depends on libm's last bits -- a
    // table from the same libm is the only way to agree with the reference there.  Static: built once per handle.
    const int N = h->N, nzh = N / 2 + 1;
    const double hh = g.sph_h;
    const double norm = (24. / (hh * hh * hh)) * (p->rho_c * p->L1 * p->L2 * p->L3 / h->ncells);
    std::vector<double> tab((size_t)N * N * (nzh + 1), 0.0);
    auto kv = [&](int m) { return (m <= N / 2) ? h->kfac * (double)m : -h->kfac * (double)(N - m); };
    auto fill_planes = [&](int i0, int i1) {
    for (int i = i0; i < i1; ++i)
      for (int j = 0; j < N; ++j)
        for (int k = 0; k < nzh; ++k) {
          const double kx = kv(i), ky = kv(j), kz = kv(k);
          const double k_sq = kx * kx + ky * ky + kz * kz;
          double F;
          if (k_sq == 0.) {
            F = 1. / (hh * hh * hh);
          } else {
            const double kk = std::sqrt(k_sq);
            const double ksink = kk * std::sin(kk);
            F = norm * (3 + std::cos(2 * kk) - ksink + std::cos(kk) * (ksink - 4)) / (k_sq * k_sq * k_sq);
          }
          tab[((size_t)i * N + j) * (nzh + 1) + k] = hh * F;
        }
    };
    {
      unsigned nt = std::thread::hardware_concurrency();
      if (nt < 1) nt = 1;
      if (nt > 16) nt = 16;
      if ((int)nt > N) nt = (unsigned)N;
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < nt; ++t) pool.emplace_back(fill_planes, (int)((size_t)N * t / nt), (int)((size_t)N * (t + 1) / nt));
      for (auto &th : pool) th.join();
    }
    dalloc(h->sph_hW, tab.size());
    BGPU_CUDA(cudaMemcpy(h->sph_hW, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  h->like.likelihood = p->likelihood;
  h->like.rho_c = p->rho_c;
  h->like.biasP = p->biasP;
  h->like.biasE = p->biasE;
  h->like.delta_min = p->delta_min;
  h->like.exact_sign = 0;

  dalloc(h->power, h->n); dalloc(h->nobs, h->n); dalloc(h->noise, h->n); dalloc(h->window, h->n);
  dalloc(h->inv_power, h->nhp);
  if (p->mass_type == 2 || p->mass_type == 3) {
    dalloc(h->zero_half, h->nhp);
    BGPU_CUDA(cudaMemsetAsync(h->zero_half, 0, h->nhp * sizeof(double), h->stream));
  }
  dalloc(h->mass_f, h->n); dalloc(h->mass_r, h->n); dalloc(h->inv_mass, h->nhp);
  dalloc(h->sig, h->n); dalloc(h->mom, h->n); dalloc(h->grad, h->n);
  for (int c = 0; c < 3; ++c) dalloc(h->psi[c], h->n);
  dalloc(h->rho_ext, (size_t)(h->Ns + 2 * h->Hmax) * h->N * h->N);
  h->delta = h->rho_ext;   // a cube has no halo; a slab moves delta to its owned planes per evaluation
  dalloc(h->resid, h->n); dalloc(h->tmp, h->n);
  if (p->sfmodel != 1 && !p->rsd_model && p->calc_h == BGPU_CALC_H_EXACT) {
    dalloc(h->phi1, h->n); dalloc(h->xa, h->n); dalloc(h->xb, h->n); dalloc(h->xc, h->n);
  }
  dalloc(h->shat, h->nh); dalloc(h->dhat, h->nh); dalloc(h->work, h->nh); dalloc(h->acc, h->nh);
  if (h->fft.can_share_x()) dalloc(h->ubuf, h->nh);
  dalloc(h->partials, (size_t)kReduceBlocks); dalloc(h->dscal, (size_t)S_COUNT);
  BGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&h->stopflag), sizeof(int)));
  BGPU_CUDA(cudaMemsetAsync(h->stopflag, 0, sizeof(int), h->stream));
  BGPU_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h->hflag2), sizeof(int)));
  *h->hflag2 = 0;
  {
    const char *lf = std::getenv("BGPU_LEAPFROG_FUSED");
    h->fused_leapfrog = !(lf && lf[0] == '0');
    const char *cc = std::getenv("BGPU_CANDIDATE_CACHE");
    h->cache_psi = !(cc && cc[0] == '0');
    const char *kl = std::getenv("BGPU_LEAPFROG_KSPACE");
    h->kspace_lf = !(kl && kl[0] == '0');
    const char *pv = std::getenv("BGPU_PARSEVAL");
    h->parseval = !(pv && pv[0] == '0');
  }
  BGPU_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h->hscal), S_COUNT * sizeof(double)));
  BGPU_CUDA(cudaMemsetAsync(h->dscal, 0, S_COUNT * sizeof(double), h->stream));
  sync(h);
  h->constructed = true;
  *out = h;
  h = nullptr;
  }
  catch (const std::exception &e) {
    g_last_error = e.what();
    if (h) {
      h->constructed = false;  // no collective in the destructor: the peers may not be there
      bgpu_destroy(h);
    }
    return 1;
  }
  catch (...) {
    g_last_error = "bgpu: unknown error while creating the handle";
    if (h) {
      h->constructed = false;
      bgpu_destroy(h);
    }
    return 1;
  }
  return 0;
}

int bgpu_create(const bgpu_params *p, bgpu_handle **out) { return create_impl(p, 0, 1, nullptr, out); }

int bgpu_nccl_unique_id(void *out128) {
  BGPU_TRY
  require(out128 != nullptr, "bgpu_nccl_unique_id: null argument");
  NcclComm::unique_id(out128);
  BGPU_CATCH
}

int bgpu_slab_create(const bgpu_params *p, int rank, int nranks, const void *nccl_id128, bgpu_handle **out) {
  return create_impl(p, rank, nranks, nccl_id128, out);
}

int bgpu_local_group_create(int nranks, bgpu_local_group **out) {
  BGPU_TRY
  require(out != nullptr, "bgpu_local_group_create: null argument");
  *out = reinterpret_cast<bgpu_local_group *>(local_group_create(nranks));
  BGPU_CATCH
}

void bgpu_local_group_destroy(bgpu_local_group *g) { local_group_destroy(reinterpret_cast<LocalGroup *>(g)); }

int bgpu_slab_create_local(const bgpu_params *p, int rank, int nranks, bgpu_local_group *g, bgpu_handle **out) {
  if (!g) {
    g_last_error = "bgpu_slab_create_local: null group";
    return 1;
  }
  return create_impl(p, rank, nranks, nullptr, out, reinterpret_cast<LocalGroup *>(g));
}

int bgpu_slab_info(const bgpu_handle *h, int *rank, int *nranks, int *x0, int *nx_local) {
  if (!h) return 1;
  if (rank) *rank = h->rank;
  if (nranks) *nranks = h->G;
  if (x0) *x0 = h->x0;
  if (nx_local) *nx_local = h->Ns;
  return 0;
}

void bgpu_destroy(bgpu_handle *h) {
  if (!h) return;
  cudaSetDevice(h->p.device);
  // (a handle whose creation failed half way skips the collectives: its peers may never arrive)
  if (h->comm && h->stream && h->constructed) {
    // collective: nobody unmaps or frees a receive buffer a peer may still be writing to
    try {
      h->comm->barrier(h->stream);
    } catch (...) {
    }
  }
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void *q : h->peer_base)
    if (q) cudaIpcCloseMemHandle(q);
  if (h->comm && h->stream && h->constructed) {
    try {
      h->comm->barrier(h->stream);
      cudaStreamSynchronize(h->stream);
    } catch (...) {
    }
  }
  double *reals[] = {h->power, h->nobs, h->noise, h->window, h->inv_power, h->zero_half, h->mass_f, h->mass_r, h->inv_mass,
                     h->sig, h->mom, h->grad, h->psi[0], h->psi[1], h->psi[2], h->rho_ext, h->resid, h->resid_ext, h->fext, h->tmp, h->cand_s, h->cand_p, h->cur_s, h->phi1, h->xa, h->xb, h->xc,
                     h->partials, h->dscal, h->halo_recv};
  for (double *q : reals)
    if (q) cudaFree(q);
  double2 *cplx[] = {h->shat, h->dhat, h->work, h->acc, h->ubuf, h->sendbuf, h->recvbuf, h->phat};
  for (double2 *q : cplx)
    if (q) cudaFree(q);
  if (h->copy_stream) {
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamDestroy(h->copy_stream);
  }
  for (int c = 0; c < bgpu_handle::kChunks; ++c) {
    if (h->ev_up[c]) cudaEventDestroy(h->ev_up[c]);
    if (h->ev_dn[c]) cudaEventDestroy(h->ev_dn[c]);
  }
  if (h->sph_kmax) cudaFree(h->sph_kmax);
  sph_columns_destroy(const_cast<SphColumns *>(h->geom.sph_cols));
  if (h->sph_hW) cudaFree(h->sph_hW);
  if (h->dflag) cudaFree(h->dflag);
  if (h->stopflag) cudaFree(h->stopflag);
  if (h->hflag2) cudaFreeHost(h->hflag2);
  if (h->hflag) cudaFreeHost(h->hflag);
  if (h->hscal) cudaFreeHost(h->hscal);
  h->fft.destroy();
  delete h->comm;
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int bgpu_set_stream(bgpu_handle *h, void *cuda_stream) {
  BGPU_TRY
  sync(h);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->own_stream = false;
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  h->fft.stream = h->stream;
  BGPU_CATCH
}

int bgpu_synchronize(bgpu_handle *h) {
  BGPU_TRY
  sync(h);
  BGPU_CATCH
}

int bgpu_set_static(bgpu_handle *h, const double *Power, const double *nobs, const double *noise,
                    const double *window) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  h->cur_psi_valid = false;  // psi() depends on every one of these
  if (Power) {
    h2d(h, h->power, Power, h->n);
    update_inverse(h, h->power, h->inv_power);
    h->have_power = true;
  }
  if (nobs) h2d(h, h->nobs, nobs, h->n);
  if (noise) h2d(h, h->noise, noise, h->n);
  if (window) h2d(h, h->window, window, h->n);
  if (nobs && noise && window) h->have_obs = true;
  sync(h);
  BGPU_CATCH
}

int bgpu_set_mass(bgpu_handle *h, const double *mass_f, const double *mass_r) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  if (h->mass_fs) {
    require(mass_f != nullptr, "bgpu_set_mass: mass_f is required for this mass_type");
    h2d(h, h->mass_f, mass_f, h->n);
    update_inverse(h, h->mass_f, h->inv_mass);
  }
  if (h->mass_rs) {
    require(mass_r != nullptr, "bgpu_set_mass: mass_r is required for this mass_type");
    h2d(h, h->mass_r, mass_r, h->n);
  }
  h->have_mass = true;
  sync(h);
  BGPU_CATCH
}

int bgpu_hamiltonian_mass(bgpu_handle *h, double *mass_f_out, double *mass_r_out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(h->have_power || h->p.mass_type == 0, "bgpu_hamiltonian_mass: Power must be set first");
  require(h->p.mass_type != 2 && h->p.mass_type != 3,
          "bgpu_hamiltonian_mass: mass types 2 / 3 depend on the signal; call bgpu_hamiltonian_mass_x");
  launch_mass(h->power, h->mass_f, h->mass_r, h->p.mass_type, h->p.mass_factor, h->n, h->stream);
  if (h->mass_fs) update_inverse(h, h->mass_f, h->inv_mass);
  h->have_mass = true;
  if (mass_f_out && h->mass_fs) d2h(h, mass_f_out, h->mass_f, h->n);
  if (mass_r_out && h->mass_rs) d2h(h, mass_r_out, h->mass_r, h->n);
  sync(h);
  BGPU_CATCH
}

// Hamiltonian_mass for every supported type (HMC_mass.cc:315-368); types 2 / 3 measure the spectrum of the
// likelihood force at `signal` (likeli_force_power, :39-51)
// likeli_force_power (HMC_mass.cc:39-51): spectrum of the likelihood gradient at `signal` -> h->tmp as
// [power | kmode | nmode], N_bin entries each
static void likeli_force_power_device(bgpu_handle *h, const double *signal) {
  const bgpu_params &p = h->p;
  require(signal != nullptr && h->have_power && h->have_obs,
          "bgpu: signal, Power and the observations are needed for the likelihood-force spectrum");
  require(h->zero_half != nullptr, "bgpu: the likelihood-force spectrum needs a handle created with mass_type 2 or 3");
  require(3 * (size_t)p.N_bin <= h->n, "bgpu: N_bin is too large for this grid (3 * N_bin doubles of bin scratch must fit one local array)");
  h2d(h, h->sig, signal, h->n);
  h->like_only = true;
  try {
    gradient_device(h, h->sig, h->grad);   // likelihood_grad_log_like alone
  } catch (...) {
    h->like_only = false;
    throw;
  }
  h->like_only = false;
  r2c_plain(h, h->grad, h->work);
  const int nb = p.N_bin;
  double *acc = h->tmp;
  launch_measure_spectrum_bin(h->work, h->N, h->Ns, h->G > 1 ? h->rank * h->Ns : 0, p.L1, nb, acc, h->stream);
  if (h->G > 1) h->comm->all_reduce_sum(acc, 3 * (size_t)nb, h->stream);
  launch_measure_spectrum_finish(acc, h->N, p.L1, nb, h->stream);
}

int bgpu_likeli_force_power(bgpu_handle *h, const double *signal, double *kmode, double *power) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  likeli_force_power_device(h, signal);
  d2h(h, power, h->tmp, (size_t)h->p.N_bin);
  d2h(h, kmode, h->tmp + h->p.N_bin, (size_t)h->p.N_bin);
  sync(h);
  BGPU_CATCH
}

int bgpu_hamiltonian_mass_x(bgpu_handle *h, const double *signal, double *mass_f_out, double *mass_r_out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  const bgpu_params &p = h->p;
  if (p.mass_type != 2 && p.mass_type != 3) return bgpu_hamiltonian_mass(h, mass_f_out, mass_r_out);
  likeli_force_power_device(h, signal);
  const int nb = p.N_bin;
  double *acc = h->tmp;                     // [power | kmode | nmode]
  double mean = 0.0;
  if (p.mass_type == 3) {                   // Hamiltonian_mass_mean_likeli_force, HMC_mass.cc:86-114
    std::vector<double> spec(2 * (size_t)nb);
    d2h(h, spec.data(), acc, 2 * (size_t)nb);
    sync(h);
    const double kfac = 2. * M_PI / p.L1, kny = kfac * (double)(h->N / 2);
    const double dk = std::sqrt((kny * kny + kny * kny) + kny * kny) / (double)nb;
    double fm = 0.0, kv = 0.0;
    for (int i = 0; i < nb; ++i) fm += 4. * M_PI * spec[nb + i] * spec[nb + i] * dk * spec[i];
    for (int i = 0; i < nb; ++i) kv += 4. * M_PI * spec[nb + i] * spec[nb + i] * dk;
    mean = fm / kv;
  }
  launch_force_mass(h->power, acc, h->mass_f, h->N, h->Ns, h->x0, p.L1, nb, p.mass_type, mean, p.mass_factor, h->stream);
  update_inverse(h, h->mass_f, h->inv_mass);
  h->have_mass = true;
  if (mass_f_out) d2h(h, mass_f_out, h->mass_f, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_gradient_psi_dev(bgpu_handle *h, const double *d_signal, double *d_gradpsi) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  gradient_device(h, d_signal, d_gradpsi);
  BGPU_CATCH
}

namespace {
struct StreamIo {
  bgpu_handle *h;
  double *host_out;
};
void wait_upload(void *ctx, int c) {
  auto *io = static_cast<StreamIo *>(ctx);
  cudaStreamWaitEvent(io->h->stream, io->h->ev_up[c], 0);
}
void start_download(void *ctx, int c) {
  auto *io = static_cast<StreamIo *>(ctx);
  bgpu_handle *h = io->h;
  const size_t per = h->n / bgpu_handle::kChunks;
  cudaEventRecord(h->ev_dn[c], h->stream);
  cudaStreamWaitEvent(h->copy_stream, h->ev_dn[c], 0);
  cudaMemcpyAsync(io->host_out + c * per, h->grad + c * per, per * sizeof(double), cudaMemcpyDeviceToHost,
                  h->copy_stream);
}
}  // namespace

int bgpu_gradient_psi(bgpu_handle *h, const double *signal, double *gradpsi) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  // The signal goes up in kChunks slabs of x planes on a copy stream; the first z pass starts on a slab
  // as soon as it has landed, and the last z pass hands each slab of the gradient to the copy stream
  // while it transforms the next -- the two PCIe transfers hide under the transforms at either end.
  constexpr int C = bgpu_handle::kChunks;
  const size_t per = h->n / C;
  cudaStreamSynchronize(h->copy_stream);  // nothing of a previous call is in flight
  cudaEventRecord(h->ev_dn[0], h->stream);
  cudaStreamWaitEvent(h->copy_stream, h->ev_dn[0], 0);  // h->sig is free once earlier work on the chain is done
  for (int c = 0; c < C; ++c) {
    BGPU_CUDA(cudaMemcpyAsync(h->sig + c * per, signal + c * per, per * sizeof(double), cudaMemcpyHostToDevice,
                              h->copy_stream));
    BGPU_CUDA(cudaEventRecord(h->ev_up[c], h->copy_stream));
  }
  StreamIo io{h, gradpsi};
  ChunkHooks in_h, out_h;
  in_h.chunks = out_h.chunks = C;
  in_h.before = wait_upload;
  out_h.after = start_download;
  in_h.ctx = out_h.ctx = &io;
  h->in_hooks = &in_h;
  h->out_hooks = &out_h;
  try {
    gradient_device(h, h->sig, h->grad);
  } catch (...) {
    h->in_hooks = h->out_hooks = nullptr;
    h->fft.hooks = nullptr;
    cudaStreamSynchronize(h->copy_stream);
    throw;
  }
  h->in_hooks = h->out_hooks = nullptr;
  BGPU_CUDA(cudaStreamSynchronize(h->copy_stream));
  sync(h);
  BGPU_CATCH
}

int bgpu_psi_dev(bgpu_handle *h, const double *d_signal, double *psi_prior, double *psi_likeli, double *d_deltaX) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  psi_device(h, d_signal);
  d2h(h, h->hscal, h->dscal, S_COUNT);
  if (d_deltaX)
    BGPU_CUDA(cudaMemcpyAsync(d_deltaX, h->delta, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  sync(h);
  *psi_prior = h->hscal[S_PRIOR];
  *psi_likeli = h->hscal[S_NLL];
  BGPU_CATCH
}

int bgpu_psi(bgpu_handle *h, const double *signal, double *psi_prior, double *psi_likeli, double *deltaX_out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  h2d(h, h->sig, signal, h->n);
  psi_device(h, h->sig);
  d2h(h, h->hscal, h->dscal, S_COUNT);
  if (deltaX_out) d2h(h, deltaX_out, h->delta, h->n);
  sync(h);
  *psi_prior = h->hscal[S_PRIOR];
  *psi_likeli = h->hscal[S_NLL];
  BGPU_CATCH
}

int bgpu_kinetic_dev(bgpu_handle *h, const double *d_momenta, double *K) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  kinetic_device(h, d_momenta);
  d2h(h, h->hscal, h->dscal, S_COUNT);
  sync(h);
  *K = h->hscal[S_KIN];
  BGPU_CATCH
}

int bgpu_kinetic(bgpu_handle *h, const double *momenta, double *K) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  h2d(h, h->mom, momenta, h->n);
  kinetic_device(h, h->mom);
  d2h(h, h->hscal, h->dscal, S_COUNT);
  sync(h);
  *K = h->hscal[S_KIN];
  BGPU_CATCH
}

int bgpu_leapfrog_dev(bgpu_handle *h, double *d_signal, double *d_momenta, uint64_t Neps, double epsilon) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  leapfrog_device(h, d_signal, d_momenta, Neps, epsilon);
  BGPU_CATCH
}

int bgpu_leapfrog(bgpu_handle *h, const double *s_i, const double *p_i, uint64_t Neps, double epsilon, double *s_f,
                  double *p_f) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  h2d(h, h->sig, s_i, h->n);
  h2d(h, h->mom, p_i, h->n);
  leapfrog_device(h, h->sig, h->mom, Neps, epsilon);
  d2h(h, s_f, h->sig, h->n);
  d2h(h, p_f, h->mom, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_color_momenta(bgpu_handle *h, const double *white, const double *real_gauss, double *momenta) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(h->have_mass, "bgpu: bgpu_set_mass or bgpu_hamiltonian_mass must be called first");
  if (h->mass_fs) {
    require(white != nullptr, "bgpu_color_momenta: white noise is required for a Fourier-space mass");
    // the full complex white-noise grid (2 N^3 doubles) does not fit any resident scratch array: own temporary.
    // A slab rank takes the FULL grid too (a mode's source entry or its mirror partner's can sit anywhere in it:
    // the host stream is serial anyway, random.cpp:57-60) and colours its own k-space rows.
    double2 *d_white = nullptr;
    const size_t n_white = (size_t)h->N * h->N * h->N;
    dalloc(d_white, n_white);
    try {
      BGPU_CUDA(cudaMemcpyAsync(d_white, white, n_white * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
      const double amp = h->ncells * h->ncells / (h->p.L1 * h->p.L2 * h->p.L3);  // random.cpp:81-83
      if (h->G == 1) launch_colour_momenta(d_white, h->mass_f, h->work, h->N, amp, h->stream);
      else launch_colour_momenta_rows(d_white, h->inv_mass, h->work, h->N, h->Ns, h->rank * h->Ns, h->ncells, h->stream);
      ROp sop;
      sop.kind = R_SCALE;
      sop.a = 1.0 / h->ncells;
      h->fft.c2r(h->work, h->work, h->mom, KOp{}, sop);
      sync(h);
    } catch (...) {
      cudaFree(d_white);
      throw;
    }
    cudaFree(d_white);
  } else {
    launch_fill(h->mom, 0.0, h->n, h->stream);                      // HMC_momenta.cc:60
  }
  if (h->mass_rs) {
    require(real_gauss != nullptr, "bgpu_color_momenta: real_gauss is required for a real-space mass");
    h2d(h, h->tmp, real_gauss, h->n);
    launch_add_real_momenta(h->mom, h->mass_r, h->tmp, h->n, h->stream);
  }
  d2h(h, momenta, h->mom, h->n);
  sync(h);
  BGPU_CATCH
}

// S5 with the generator on the device.  Same distribution as draw_momenta
// (HMC_momenta.cc:42-92), different stream: the GSL mt19937 stream is serial by construction and costs
// 2N host Gaussians per candidate (0.7 s at 256^3, ten times a whole trajectory on the GPU), so production
// runs that do not need seed-for-seed agreement with the CPU code draw here instead.
static void draw_momenta_device(bgpu_handle *h, uint64_t seed, uint64_t draw, double *d_out) {
  require(h->have_mass, "bgpu: bgpu_set_mass or bgpu_hamiltonian_mass must be called first");
  // element e of the global grid always gets the same number, whatever the decomposition
  const size_t first = (size_t)h->x0 * h->N * h->N;
  if (h->mass_fs) {
    launch_philox_normals(h->tmp, h->n, first, seed, draw, 0, h->stream);
    r2c_plain(h, h->tmp, h->work);
    // colour with the multiplier the kinetic term uses, (N/V) M = 1/inv_mass, on the cube [x][y][.] and on the
    // transposed k-space slab [x][y_local][.] alike: the draw is bitwise independent of the decomposition
    launch_colour_white_rows(h->work, h->inv_mass, h->N, h->nh, h->rank == 0, h->stream);
    ROp sop;
    sop.kind = R_SCALE;
    sop.a = 1.0 / h->ncells;
    h->fft.c2r(h->work, h->work, d_out, KOp{}, sop);
  } else {
    launch_fill(d_out, 0.0, h->n, h->stream);
  }
  if (h->mass_rs) {
    launch_philox_normals(h->tmp, h->n, first, seed, draw, 1, h->stream);
    launch_add_real_momenta(d_out, h->mass_r, h->tmp, h->n, h->stream);
  }
}

int bgpu_draw_momenta_device_dev(bgpu_handle *h, uint64_t seed, uint64_t draw_index, double *d_momenta) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  draw_momenta_device(h, seed, draw_index, d_momenta);
  BGPU_CATCH
}

int bgpu_draw_momenta_device(bgpu_handle *h, uint64_t seed, uint64_t draw_index, double *momenta) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  draw_momenta_device(h, seed, draw_index, h->mom);
  d2h(h, momenta, h->mom, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_device_normals(bgpu_handle *h, uint64_t seed, uint64_t draw_index, unsigned stream, size_t first, size_t n,
                        double *out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(n <= h->n && n % 2 == 0 && first % 2 == 0, "bgpu_device_normals: n must be even, at most N^3; first even");
  launch_philox_normals(h->tmp, n, first, seed, draw_index, stream, h->stream);
  d2h(h, out, h->tmp, n);
  sync(h);
  BGPU_CATCH
}

// measure_spectrum (field_statistics.cpp:20-90), the per-sample P(k) diagnostic
int bgpu_measure_spectrum(bgpu_handle *h, const double *signal, uint64_t N_bin, double *kmode, double *power) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(N_bin >= 1 && N_bin <= 2048 && 3 * N_bin <= h->n, "bgpu_measure_spectrum: N_bin must be in [1, 2048]");
  h2d(h, h->tmp, signal, h->n);
  r2c_plain(h, h->tmp, h->work);
  double *acc = h->grad;  // scratch: 3 * N_bin doubles
  launch_measure_spectrum_bin(h->work, h->N, h->Ns, h->G > 1 ? h->rank * h->Ns : 0, h->p.L1, (int)N_bin, acc, h->stream);
  if (h->G > 1) h->comm->all_reduce_sum(acc, 3 * (size_t)N_bin, h->stream);
  launch_measure_spectrum_finish(acc, h->N, h->p.L1, (int)N_bin, h->stream);
  d2h(h, power, acc, N_bin);
  d2h(h, kmode, acc + N_bin, N_bin);
  sync(h);
  BGPU_CATCH
}

// ---------------------------------------------------------------------------
// One HMC candidate without the signal or the momenta ever leaving the device: the body of HamiltonianMC's loop
// (HMC.cc:436-506) between the host's RNG draws and its Metropolis decision -- S5 (device generator), S4, S3, S2.
// ---------------------------------------------------------------------------
int bgpu_set_signal(bgpu_handle *h, const double *x) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(x != nullptr, "bgpu_set_signal: null signal");
  if (!h->cand_s) {
    dalloc(h->cand_s, h->n);
    dalloc(h->cand_p, h->n);
    dalloc(h->cur_s, h->n);   // the chain's current signal has its own buffer: every other entry point may overwrite h->sig
  }
  h2d(h, h->cur_s, x, h->n);
  h->have_signal = true;
  h->cur_psi_valid = false;
  sync(h);
  BGPU_CATCH
}

int bgpu_candidate(bgpu_handle *h, uint64_t seed, uint64_t draw_index, uint64_t Neps, double epsilon, double *energies6,
                   double *p_f0) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(h->have_signal, "bgpu_candidate: bgpu_set_signal must be called first");
  require(energies6 != nullptr, "bgpu_candidate: null output");
  draw_momenta_device(h, seed, draw_index, h->cand_p);                       // HMC.cc:449
  // delta_Hamiltonian's initial energies (HMC.cc:214-215): the ends of the trajectory do not change them
  kinetic_device(h, h->cand_p);
  const bool cached = h->cache_psi && h->cur_psi_valid;
  if (!cached) psi_device(h, h->cur_s);
  BGPU_CUDA(cudaMemcpyAsync(h->hscal, h->dscal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  BGPU_CUDA(cudaMemcpyAsync(h->cand_s, h->cur_s, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  sync(h);
  if (!cached) {
    h->cur_psi[0] = h->hscal[S_PRIOR];
    h->cur_psi[1] = h->hscal[S_NLL];
    h->cur_psi_valid = true;
  }
  energies6[0] = h->hscal[S_KIN];
  energies6[1] = h->cur_psi[0];
  energies6[2] = h->cur_psi[1];
  h->psi_from_kick = false;
  h->psi_at_end = h->cache_psi;
  leapfrog_device(h, h->cand_s, h->cand_p, Neps, epsilon);                   // HMC.cc:455
  h->psi_at_end = false;
  kinetic_device(h, h->cand_p);                                              // :224-225; deltaX is left at s_f's
  if (!h->psi_from_kick) psi_device(h, h->cand_s);                           // else the last kick has left psi(s_f) in dscal
  h->psi_from_kick = false;
  BGPU_CUDA(cudaMemcpyAsync(h->hscal, h->dscal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (p_f0) BGPU_CUDA(cudaMemcpyAsync(p_f0, h->cand_p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  energies6[3] = h->hscal[S_KIN];
  energies6[4] = h->cand_psi[0] = h->hscal[S_PRIOR];
  energies6[5] = h->cand_psi[1] = h->hscal[S_NLL];
  BGPU_CATCH
}

int bgpu_accept(bgpu_handle *h, double *x_out, double *deltaX_out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(h->have_signal && h->cand_s, "bgpu_accept: no candidate");
  BGPU_CUDA(cudaMemcpyAsync(h->cur_s, h->cand_s, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->cur_psi[0] = h->cand_psi[0];   // psi of the new current signal = the accepted candidate's final energies
  h->cur_psi[1] = h->cand_psi[1];
  if (x_out) d2h(h, x_out, h->cur_s, h->n);
  if (deltaX_out) d2h(h, deltaX_out, h->delta, h->n);
  sync(h);
  BGPU_CATCH
}

// ---------------------------------------------------------------------------
// F4: the mock-data generator and the initial guess on the device (barcoderunner.cc:42-247) -- a production
// run never draws 2 N^3 serial GSL Gaussians for them.  Counter-based Philox streams, not GSL's.
// ---------------------------------------------------------------------------
namespace {
// a Gaussian random field with spectrum h->power (create_GARFIELD's normalisation, random.cpp:81-83), optionally
// smoothed with the Gaussian filter of kernelcomp (filtertype 1, convolution.cpp:224-322): d_out, real space
void grf_device(bgpu_handle *h, uint64_t seed, uint64_t draw, unsigned stream, double smoothing, double *d_out) {
  require(h->have_power, "bgpu: Power must be set first (bgpu_set_static)");
  require(h->G == 1, "bgpu: the device field generator is single-GPU (draw on a cube handle and distribute the slabs)");
  launch_philox_normals(h->tmp, h->n, 0, seed, draw, stream, h->stream);
  r2c_plain(h, h->tmp, h->work);
  launch_colour_white(h->work, h->power, h->N, h->ncells / (h->p.L1 * h->p.L2 * h->p.L3), h->stream);
  KOp lop;
  if (smoothing > 0.) {
    lop.kind = K_GAUSS;
    lop.a = smoothing;
    lop.kfac = h->kfac;
  }
  ROp sop;
  sop.kind = R_SCALE;
  sop.a = 1.0 / h->ncells;
  h->fft.c2r(h->work, h->work, d_out, lop, sop);
}
}  // namespace

int bgpu_mock_data(bgpu_handle *h, uint64_t seed, const bgpu_mock_params *mp, double *delta_lag, double *delta_eul,
                   double *nobs, double *noise, double *window) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(mp != nullptr, "bgpu_mock_data: null parameters");
  const bgpu_params &p = h->p;
  require(mp->window_type == 1 || mp->window_type == 10 || mp->window_type == 23,
          "in barcoderunner: window_type is not a valid choice! (1, 10 or 23)");
  require(mp->data_model == 0 || mp->data_model == 1, "in barcoderunner: data_model is not a valid choice! (0 or 1)");
  if (mp->data_model == 0)
    require(p.likelihood == 0 || p.likelihood == 1 || p.likelihood == 3,
            "in barcoderunner: linear data model was chosen (additive error), but incompatible likelihood!");
  // the truth: delta_lag ~ GRF(Power) -> h->sig ; delta_eul = Lag2Eul(delta_lag) with the handle's own model -> h->delta
  grf_device(h, seed, 0, 2, 0.0, h->sig);
  if (p.likelihood == 3) {
    // no structure formation in the Gaussian-random-field test: the "Eulerian" field is not used by its data model
    BGPU_CUDA(cudaMemcpyAsync(h->delta, h->sig, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  } else {
    r2c_plain(h, h->sig, h->shat);
    forward_from_shat(h, h->sig, 1.0, p.rsd_model != 0, nullptr, nullptr, nullptr);
    LikeParams lp = h->like;
    lp.exact_sign = 0;
    launch_fill(h->mom, 1.0, h->n, h->stream);
    launch_overdens_residual(lp, h->delta, h->dscal + S_SUMRHO, h->mom, h->mom, h->mom, nullptr, h->n, h->ncells,
                             h->partials, h->dscal + S_NLL, h->stream);
  }
  MockObs mo;
  mo.likelihood = p.likelihood;
  mo.data_model = mp->data_model;
  mo.window_type = mp->window_type;
  mo.negative_obs = mp->negative_obs;
  mo.rho_c = p.rho_c;
  mo.delta_min = p.delta_min;
  mo.sigma_min = mp->sigma_min;
  mo.sigma_fac = mp->sigma_fac;
  launch_fill(h->noise, 0.0, h->n, h->stream);
  launch_mock_obs(mo, h->delta, h->sig, h->window, h->nobs, h->noise, h->n, 0, h->n, seed, h->stream);
  h->have_obs = true;
  h->cur_psi_valid = false;
  if (delta_lag) d2h(h, delta_lag, h->sig, h->n);
  if (delta_eul) d2h(h, delta_eul, h->delta, h->n);
  if (nobs) d2h(h, nobs, h->nobs, h->n);
  if (noise) d2h(h, noise, h->noise, h->n);
  if (window) d2h(h, window, h->window, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_initial_guess(bgpu_handle *h, uint64_t seed, int initial_guess, double smoothing_scale, double *signal) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(signal != nullptr, "bgpu_initial_guess: null output");
  switch (initial_guess) {
    case 0:  // zero initial guess
      launch_fill(h->sig, 0.0, h->n, h->stream);
      break;
    case 2:  // GRF initial guess
      grf_device(h, seed, 0, 3, 0.0, h->sig);
      break;
    case 3:  // smoothed GRF initial guess (Gaussian filter)
      require(smoothing_scale > 0., "bgpu_initial_guess: initial_guess 3 needs a smoothing scale > 0");
      grf_device(h, seed, 0, 3, smoothing_scale, h->sig);
      break;
    case 4:  // zero plus some random noise: sigma = 0.1
      launch_philox_normals(h->tmp, h->n, (size_t)h->x0 * h->N * h->N, seed, 0, 4, h->stream);
      launch_scale(h->sig, h->tmp, 1.e-1, h->n, h->stream);
      break;
    default:
      throw std::runtime_error("In barcoderunner: invalid choice of initial_guess (" + std::to_string(initial_guess) +
                               ")! (1, a file, is the host's job)");
  }
  d2h(h, signal, h->sig, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_forward(bgpu_handle *h, const double *signal, double *deltaX, double *posx, double *posy, double *posz) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  h2d(h, h->sig, signal, h->n);
  r2c_plain(h, h->sig, h->shat);
  const bool want_pos = posx && posy && posz;
  // positions are staged in grad / resid / tmp (free during the forward model)
  forward_from_shat(h, h->sig, h->p.deltaQ_factor, h->p.rsd_model != 0, want_pos ? h->grad : nullptr,
                    want_pos ? h->resid : nullptr, want_pos ? h->tmp : nullptr);
  if (want_pos) {
    d2h(h, posx, h->grad, h->n);
    d2h(h, posy, h->resid, h->n);
    d2h(h, posz, h->tmp, h->n);
  }
  // overdens only (massFunctions.cc:30-47): reuse the residual kernel in value-only mode with a unit window
  LikeParams lp = h->like;
  lp.exact_sign = 0;
  launch_fill(h->mom, 1.0, h->n, h->stream);
  launch_overdens_residual(lp, h->delta, h->dscal + S_SUMRHO, h->mom, h->mom, h->mom, nullptr, h->n, h->ncells,
                           h->partials, h->dscal + S_NLL, h->stream);
  d2h(h, deltaX, h->delta, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_assign_density(bgpu_handle *h, const double *x, const double *y, const double *z, double *rho) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(h->G == 1, "bgpu_assign_density: arbitrary particle lists are not available on a slab-decomposed chain "
                     "(its density tile holds the local planes and their halo only); use a cube handle");
  h2d(h, h->psi[0], x, h->n);
  h2d(h, h->psi[1], y, h->n);
  h2d(h, h->psi[2], z, h->n);
  launch_scatter_positions(h->geom, h->psi[0], h->psi[1], h->psi[2], h->delta, h->stream);
  d2h(h, rho, h->delta, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_cell_indices(bgpu_handle *h, const double *x, const double *y, const double *z, size_t n, int *ci, int *cj,
                      int *ck) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(n <= h->n, "bgpu_cell_indices: at most as many positions per call as the handle has local cells");
  h2d(h, h->psi[0], x, n);
  h2d(h, h->psi[1], y, n);
  h2d(h, h->psi[2], z, n);
  int *di = reinterpret_cast<int *>(h->delta), *dj = reinterpret_cast<int *>(h->resid),
      *dk = reinterpret_cast<int *>(h->tmp);
  launch_cell_indices(h->geom, h->psi[0], h->psi[1], h->psi[2], di, dj, dk, n, h->stream);
  BGPU_CUDA(cudaMemcpyAsync(ci, di, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  BGPU_CUDA(cudaMemcpyAsync(cj, dj, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  BGPU_CUDA(cudaMemcpyAsync(ck, dk, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  BGPU_CATCH
}

int bgpu_fft_r2c(bgpu_handle *h, const double *in, double *out_complex) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  h2d(h, h->sig, in, h->n);
  r2c_plain(h, h->sig, h->work);
  d2h(h, out_complex, reinterpret_cast<double *>(h->work), 2 * h->nh);
  sync(h);
  BGPU_CATCH
}

int bgpu_fft_c2r(bgpu_handle *h, const double *in_complex, double *out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  h2d(h, reinterpret_cast<double *>(h->work), in_complex, 2 * h->nh);
  ROp sop;
  sop.kind = R_SCALE;
  sop.a = 1.0 / h->ncells;   // fftwrapper.cc:43-45
  h->fft.c2r(h->work, h->work, h->tmp, KOp{}, sop);
  d2h(h, out, h->tmp, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_convolve_inv_corr(bgpu_handle *h, const double *signal, const double *corr, double *out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  h2d(h, h->sig, signal, h->n);
  h2d(h, h->tmp, corr, h->n);
  // the half-grid multiplier goes to the first half of h->acc's storage
  double *half = reinterpret_cast<double *>(h->acc);
  update_inverse(h, h->tmp, half);
  r2c_plain(h, h->sig, h->work);
  KOp lop;
  lop.kind = K_MULREAL;
  lop.real0 = half;
  ROp sop;
  sop.kind = R_SCALE;
  sop.a = 1.0 / h->ncells;
  h->fft.c2r(h->work, h->work, h->tmp, lop, sop);
  d2h(h, out, h->tmp, h->n);
  sync(h);
  BGPU_CATCH
}

}  // extern "C"
