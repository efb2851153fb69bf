// barcode_b200/csrc/kernels.cu -- the non-FFT kernels of the HMC hot path.
//
// Each kernel names the reference loop it replaces (paths under
// /root/reference/barlib/src).  Integer cell indices and interpolation weights
// follow the reference's arithmetic operation by operation (explicit
// round-to-nearest intrinsics where nvcc would otherwise contract a*b+c into
// an FMA, which the reference's x86-64 build never does), so positions, cell
// indices and weights are bit-identical for identical displacement input.
#include "kernels.h"

#include <math.h>

#include <cstdlib>

#include "particle_math.cuh"
#include "util.h"

namespace bgpu {

__device__ __forceinline__ void red_add(double *addr, double v) { atomicAdd(addr, v); }

// deposit one particle per lane, getDensity_NGP / _CIC / _TSC (massFunctions.cc:49-364), unit mass.
//
// The reference issues one `#pragma omp atomic` per cell and particle (8 for CIC, 27 for TSC).  On
// the GPU the L2's atomic units are the bound (measured: ~270 G red.f64/s = one per L2 slice per
// clock), so the warp aggregates first.  Lanes hold particles that are neighbours along z on the
// Lagrangian lattice, and neighbours mostly land in neighbouring cells: the upper z-cell of lane
// l is the lower z-cell of lane l+1.  Contributions are therefore exchanged with shuffles and
// merged whenever the ADDRESSES agree (so any displacement field is handled exactly; without a
// match the lane just issues its own atomic): CIC goes from 8 to ~4.1 atomics per particle, TSC
// from 27 to ~9.6.  All lanes of the warp must call this (invalid lanes deposit nothing).
// A shared-memory-tile variant with a TMA reduce-add flush was built and measured slower: f64
// shared atomics are CAS loops (ATOMS.CAST.SPIN) and neighbouring lanes collide on them.
__device__ __forceinline__ void deposit(const GridGeom &g, bool valid, double x, double y, double z,
                                        double *__restrict__ rho) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  valid = valid && in_domain(g, x, y, z);
  const int N = g.N;
  // plane of the (possibly halo-extended, slab-local) density tile that global cell plane c maps to
  const int lo = g.x0 - g.H, np = g.Ns + 2 * g.H;
  auto plane = [&](int c) -> unsigned {
    int l = c - lo;
    if (l < 0) l += N;
    if (l >= N) l -= N;
    if (l >= np) {  // beyond the halo: drop the deposit and raise the flag (the host turns it into an error)
      if (g.flag) *g.flag = 1;
      l = 0;
      valid = false;
    }
    return (unsigned)l;
  };
  if (g.masskernel == 1) {
    int ci[2] = {0, 0}, cj[2] = {0, 0}, ck[2] = {0, 0};
    double wi[2] = {0, 0}, wj[2] = {0, 0}, wk[2] = {0, 0};
    if (valid) {
      cic_axis(x, g.d, g.L, N, ci[0], ci[1], wi[0], wi[1]);
      cic_axis(y, g.d, g.L, N, cj[0], cj[1], wj[0], wj[1]);
      cic_axis(z, g.d, g.L, N, ck[0], ck[1], wk[0], wk[1]);
      ci[0] = (int)plane(ci[0]);
      ci[1] = (int)plane(ci[1]);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const unsigned row = ((unsigned)ci[a] * N + cj[b]) * N;
        const unsigned alo = valid ? row + ck[0] : 0xffffffffu, ahi = valid ? row + ck[1] : 0xfffffffeu;
        // mass*w_x*w_y*w_z evaluated left to right, massFunctions.cc:129-157
        const double wab = __dmul_rn(wi[a], wj[b]);
        double lo = __dmul_rn(wab, wk[0]);
        const double hi = __dmul_rn(wab, wk[1]);
        const unsigned p_ahi = __shfl_up_sync(FULL, ahi, 1), n_alo = __shfl_down_sync(FULL, alo, 1);
        const double p_hi = __shfl_up_sync(FULL, hi, 1);
        if (lane > 0 && p_ahi == alo) lo += p_hi;            // the lane below hands over its upper cell
        const bool handed_up = lane < 31 && n_alo == ahi;    // ... and the lane above takes mine
        if (valid) {
          red_add(rho + alo, lo);
          if (!handed_up) red_add(rho + ahi, hi);
        }
      }
  } else if (g.masskernel == 2) {
    int ci[3] = {0, 0, 0}, cj[3] = {0, 0, 0}, ck[3] = {0, 0, 0};
    double wi[3] = {0, 0, 0}, wj[3] = {0, 0, 0}, wk[3] = {0, 0, 0}, u;
    if (valid) {
      tsc_axis(x, g.min1, g.d, N, ci, wi, u);
      tsc_axis(y, g.min2, g.d, N, cj, wj, u);
      tsc_axis(z, g.min3, g.d, N, ck, wk, u);
#pragma unroll
      for (int a = 0; a < 3; ++a) ci[a] = (int)plane(ci[a]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const unsigned row = ((unsigned)ci[a] * N + cj[b]) * N;
        const unsigned am = valid ? row + ck[0] : 0xffffffffu, a0 = valid ? row + ck[1] : 0xfffffffeu,
                       ap = valid ? row + ck[2] : 0xfffffffdu;
        const double wab = __dmul_rn(wi[a], wj[b]);
        const double vm = __dmul_rn(wab, wk[0]), vp = __dmul_rn(wab, wk[2]);
        double v0 = __dmul_rn(wab, wk[1]);
        // lane l-1 sits one cell below: its centre is my lower cell and its upper cell is my centre
        const unsigned p_a0 = __shfl_up_sync(FULL, a0, 1), p_ap = __shfl_up_sync(FULL, ap, 1);
        const unsigned n_a0 = __shfl_down_sync(FULL, a0, 1), n_am = __shfl_down_sync(FULL, am, 1);
        const double p_vp = __shfl_up_sync(FULL, vp, 1), n_vm = __shfl_down_sync(FULL, vm, 1);
        const bool left = lane > 0 && p_a0 == am && p_ap == a0;
        const bool right = lane < 31 && n_a0 == ap && n_am == a0;
        if (left) v0 += p_vp;
        if (right) v0 += n_vm;
        if (valid) {
          red_add(rho + a0, v0);
          if (!left) red_add(rho + am, vm);
          if (!right) red_add(rho + ap, vp);
        }
      }
  } else if (valid) {
    const int i = (int)plane(ngp_axis(x, g.min1, g.d, N)), j = ngp_axis(y, g.min2, g.d, N),
              k = ngp_axis(z, g.min3, g.d, N);
    if (valid) red_add(rho + ((size_t)i * N + j) * N + k, 1.0);
  }
}

// ---------------------------------------------------------------------------
// K7: scatter
// ---------------------------------------------------------------------------
__global__ void scatter_kernel(GridGeom g, const double *__restrict__ psix, const double *__restrict__ psiy,
                               const double *__restrict__ psiz, double *__restrict__ rho, double *__restrict__ posx,
                               double *__restrict__ posy, double *__restrict__ posz) {
  const size_t n = (size_t)g.Ns * g.N * g.N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = idx < n;  // no early return: deposit() shuffles across the whole warp
  double x = 0., y = 0., z = 0.;
  if (valid) {
    // N is a power of two (the FFT's constraint): shifts, not three 64-bit divisions by a run-time N
    const int sh = 31 - __clz(g.N);
    const int k = (int)(idx & (size_t)(g.N - 1));
    const int j = (int)((idx >> sh) & (size_t)(g.N - 1));
    const int il = (int)(idx >> (2 * sh));
    const int i = g.x0 + il;
    double px = psix[idx], py = psiy[idx], pz = psiz[idx];
    if (g.cellbound) {
      // cellboundcomp (massFunctions.cc:588-660): every displacement averaged with its (i-1, j-1, k-1)
      // neighbour, periodically -- folded into the read instead of a pass over three arrays
      const int N = g.N;
      const size_t jk = (size_t)(j == 0 ? N - 1 : j - 1) * N + (k == 0 ? N - 1 : k - 1);
      if (il == 0 && g.cb_lo) {  // slab: the plane below my first one came from the lower neighbour
        const size_t pl = (size_t)N * N;
        px = 0.5 * (g.cb_lo[jk] + px);
        py = 0.5 * (g.cb_lo[pl + jk] + py);
        pz = 0.5 * (g.cb_lo[2 * pl + jk] + pz);
      } else {
        const size_t m = (size_t)(il == 0 ? N - 1 : il - 1) * N * N + jk;
        px = 0.5 * (psix[m] + px);
        py = 0.5 * (psiy[m] + py);
        pz = 0.5 * (psiz[m] + pz);
      }
    }
    particle_position(g, i, j, k, px, py, pz, x, y, z);
    if (posx) {
      posx[idx] = x;
      posy[idx] = y;
      posz[idx] = z;
    }
  }
  deposit(g, valid, x, y, z, rho);
}

__global__ void scatter_positions_kernel(GridGeom g, const double *__restrict__ x, const double *__restrict__ y,
                                         const double *__restrict__ z, double *__restrict__ rho) {
  const size_t n = (size_t)g.N * g.N * g.N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = idx < n;
  deposit(g, valid, valid ? x[idx] : 0., valid ? y[idx] : 0., valid ? z[idx] : 0., rho);
}

__global__ void cell_indices_kernel(GridGeom g, const double *__restrict__ x, const double *__restrict__ y,
                                    const double *__restrict__ z, int *ci, int *cj, int *ck, size_t n) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  if (g.masskernel == 1 && g.lean) {
    // the arithmetic of the lean sweep kernels (particles_sweep.cu), so that the bit-exact index tests pin it
    const LeanConst lc = lean_const(g);
    const double pos[3] = {x[idx], y[idx], z[idx]};
    int *const out[3] = {ci, cj, ck};
    for (int c = 0; c < 3; ++c) {
      unsigned c0, c1;
      double w0, w1;
      bool slow = !(pos[c] >= 0. && pos[c] < g.L);
      lean_axis(pos[c], lc, c0, c1, w0, w1, slow);
      if (slow) {
        int a, b;
        cic_axis(pos[c], g.d, g.L, g.N, a, b, w0, w1);
        c0 = (unsigned)a;
      }
      out[c][idx] = (int)c0;
    }
  } else if (g.masskernel == 1) {
    int a, b;
    double t, dx;
    cic_axis(x[idx], g.d, g.L, g.N, a, b, t, dx);
    ci[idx] = a;
    cic_axis(y[idx], g.d, g.L, g.N, a, b, t, dx);
    cj[idx] = a;
    cic_axis(z[idx], g.d, g.L, g.N, a, b, t, dx);
    ck[idx] = a;
  } else {
    ci[idx] = ngp_axis(x[idx], g.min1, g.d, g.N);
    cj[idx] = ngp_axis(y[idx], g.min2, g.d, g.N);
    ck[idx] = ngp_axis(z[idx], g.min3, g.d, g.N);
  }
}

static inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

void launch_scatter(const GridGeom &g, const double *psix, const double *psiy, const double *psiz, double *rho,
                    double *posx, double *posy, double *posz, cudaStream_t st) {
  if (scatter_sweep_applicable(g, posx)) {
    launch_scatter_sweep(g, psix, psiy, psiz, rho, st);
    return;
  }
  ProfScope prof(KK_SCATTER, st);
  const size_t n = (size_t)g.Ns * g.N * g.N;
  BGPU_CUDA(cudaMemsetAsync(rho, 0, (size_t)(g.Ns + 2 * g.H) * g.N * g.N * sizeof(double), st));
  scatter_kernel<<<blocks_for(n, 256), 256, 0, st>>>(g, psix, psiy, psiz, rho, posx, posy, posz);
  BGPU_LAUNCHED(1);
}

__global__ void scatter_sph_positions_kernel(GridGeom g, const double *__restrict__ x, const double *__restrict__ y,
                                             const double *__restrict__ z, double *__restrict__ rho);

void launch_scatter_positions(const GridGeom &g, const double *x, const double *y, const double *z, double *rho,
                              cudaStream_t st) {
  ProfScope prof(KK_SCATTER, st);
  const size_t n = (size_t)g.N * g.N * g.N;
  BGPU_CUDA(cudaMemsetAsync(rho, 0, n * sizeof(double), st));
  if (g.masskernel == 3 && g.sph_cols) {
    launch_scatter_sph_cols_positions(g, g.sph_cols, x, y, z, rho, st);
    return;
  } else if (g.masskernel == 3)
    scatter_sph_positions_kernel<<<blocks_for(n, 128), 128, 0, st>>>(g, x, y, z, rho);
  else
    scatter_positions_kernel<<<blocks_for(n, 256), 256, 0, st>>>(g, x, y, z, rho);
  BGPU_LAUNCHED(1);
}

void launch_cell_indices(const GridGeom &g, const double *x, const double *y, const double *z, int *ci, int *cj,
                         int *ck, size_t n, cudaStream_t st) {
  cell_indices_kernel<<<blocks_for(n, 256), 256, 0, st>>>(g, x, y, z, ci, cj, ck, n);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// deterministic two-stage reductions (fixed grid, fixed tree)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double warp_part[kReduceThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kReduceThreads / 32; ++w) r += warp_part[w];
  }
  __syncthreads();
  return r;  // valid in thread 0
}

__global__ void __launch_bounds__(kReduceThreads) final_sum_kernel(const double *__restrict__ part, int nparts,
                                                                    double *__restrict__ out) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nparts; i += kReduceThreads) v += part[i];
  const double r = block_sum(v);
  if (threadIdx.x == 0) *out = r;
}

template <class F>
__global__ void __launch_bounds__(kReduceThreads) partial_sum_kernel(F f, size_t n, double *__restrict__ part) {
  double v = 0.0;
  const size_t stride = (size_t)gridDim.x * kReduceThreads;
  size_t i = (size_t)blockIdx.x * kReduceThreads + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const double a = f(i), b = f(i + stride), c = f(i + 2 * stride), d = f(i + 3 * stride);
    v += a;
    v += b;
    v += c;
    v += d;
  }
  for (; i < n; i += stride) v += f(i);
  const double r = block_sum(v);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

struct SumF {
  const double *a;
  __device__ double operator()(size_t i) const { return a[i]; }
};
struct HalfDotF {
  const double *a, *b;
  __device__ double operator()(size_t i) const { return 0.5 * a[i] * b[i]; }
};
struct KineticF {  // HMC.cc:88-110
  const double *p, *conv, *mass_r;
  __device__ double operator()(size_t i) const {
    double dummy = conv ? conv[i] : 0.0;
    if (mass_r) {
      const double m = mass_r[i];
      const double invM = m > 0.0 ? 1.0 / m : 0.0;
      dummy += invM * p[i];
    }
    return 0.5 * p[i] * dummy;
  }
};

// 1/2 sum_x a(x) (C^-1 a)(x) by Parseval on the half grid: (1/2N) sum_k w_k mult_k |a^_k|^2 with w = 1 on the planes
// k_z = 0 and N/2 (their mirrors are stored) and 2 elsewhere; mult = the padded half-grid multiplier (V/N)/C of
// convolveInvCorrFuncWithSignal (HMC_help.cc:41-58).  Saves the inverse transform of kinetic_term (HMC.cc:82-110)
// and of the prior (gaussian.cpp:20-35); any row layout (cube, transposed slab): only z matters.
struct HalfQuadF {
  const double2 *v;
  const double *mult;
  int nzh;
  double inv_2n;
  __device__ double operator()(size_t i) const {
    const size_t row = i / (size_t)nzh;
    const int z = (int)(i - row * (size_t)nzh);
    const double2 a = v[i];
    const double w = (z == 0 || z == nzh - 1) ? 1.0 : 2.0;
    return inv_2n * w * mult[row * (size_t)(nzh + 1) + z] * (a.x * a.x + a.y * a.y);
  }
};

template <class F>
static void reduce(F f, size_t n, double *scratch, double *out, cudaStream_t st, int kind = KK_REDUCE) {
  ProfScope prof(kind, st);
  const int blocks = (int)((n + kReduceThreads - 1) / kReduceThreads < (size_t)kReduceBlocks
                               ? (n + kReduceThreads - 1) / kReduceThreads
                               : (size_t)kReduceBlocks);
  partial_sum_kernel<F><<<blocks, kReduceThreads, 0, st>>>(f, n, scratch);
  final_sum_kernel<<<1, kReduceThreads, 0, st>>>(scratch, blocks, out);
  BGPU_LAUNCHED(2);
}

// ---------------------------------------------------------------------------
// Leapfrog in k-space (api.cu leapfrog_device): with a Fourier-space mass and a forward model that reads s^ only,
// both updates of Hamiltonian_EoM (HMC.cc:293-352) are diagonal on the half grid --
//   drift  s^ += eps (V/N)/M p^                       (the transform pair of :298-339 without the transforms)
//   kick   p^ += a ((V/N)/P s^ + norm h^)             (gradpsi's last sum, HMC.cc:205, before ITS inverse transform)
// -- so a step needs neither the forward transform of s, nor the inverse transform of gradpsi, nor the pair around
// M^-1 p.  The kick also returns momenta[0] = (1/N) sum_k w_k Re p^_k (w = 1 on the planes k_z = 0 and N/2, whose
// mirrors are stored, 2 elsewhere) for the run-away test of :360-364, in the fixed order of the two-stage sum.
// Multipliers use the padded row pitch N/2 + 2 (launch_inverse_spectrum).
// ---------------------------------------------------------------------------
__global__ void kspace_drift_kernel(double2 *__restrict__ shat, const double2 *__restrict__ phat,
                                    const double *__restrict__ inv_mass, double eps, int nzh, size_t nh,
                                    const int *__restrict__ skip) {
  if (skip && *skip) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const size_t row = i / (size_t)nzh;
  const double f = eps * inv_mass[row * (size_t)(nzh + 1) + (i - row * (size_t)nzh)];
  const double2 pk = phat[i];
  double2 sk = shat[i];
  sk.x = fma(f, pk.x, sk.x);
  sk.y = fma(f, pk.y, sk.y);
  shat[i] = sk;
}

struct KspaceKickF {
  double2 *phat;
  const double2 *shat, *hhat;
  const double *prior;
  double a, norm, inv_n;
  int nzh;
  const int *skip;
  __device__ double operator()(size_t i) const {
    const size_t row = i / (size_t)nzh;
    const int z = (int)(i - row * (size_t)nzh);
    double2 pk = phat[i];
    if (!(skip && *skip)) {
      const double f = prior[row * (size_t)(nzh + 1) + z];
      const double2 sk = shat[i], hk = hhat[i];
      pk.x = fma(a, fma(sk.x, f, norm * hk.x), pk.x);
      pk.y = fma(a, fma(sk.y, f, norm * hk.y), pk.y);
      phat[i] = pk;
    }
    return ((z == 0 || z == nzh - 1) ? inv_n : 2.0 * inv_n) * pk.x;
  }
};

void launch_kspace_drift(double2 *shat, const double2 *phat, const double *inv_mass, double eps, int N, size_t nh,
                         cudaStream_t st, const int *skip) {
  ProfScope prof(KK_STREAM, st);
  kspace_drift_kernel<<<blocks_for(nh, 256), 256, 0, st>>>(shat, phat, inv_mass, eps, N / 2 + 1, nh, skip);
  BGPU_LAUNCHED(1);
}
void launch_kspace_kick(double2 *phat, const double2 *shat, const double2 *hhat, const double *prior, double a,
                        double norm, int N, size_t nh, double ncells, double *scratch, double *p0_out, cudaStream_t st,
                        const int *skip) {
  reduce(KspaceKickF{phat, shat, hhat, prior, a, norm, 1.0 / ncells, N / 2 + 1, skip}, nh, scratch, p0_out, st, KK_STREAM);
}

void launch_half_quadratic(const double2 *vhat, const double *mult_half, int N, size_t n_half, double ncells,
                           double *scratch, double *out, cudaStream_t st) {
  reduce(HalfQuadF{vhat, mult_half, N / 2 + 1, 0.5 / ncells}, n_half, scratch, out, st);
}
void launch_sum(const double *a, size_t n, double *scratch, double *out, cudaStream_t st) {
  reduce(SumF{a}, n, scratch, out, st);
}
void launch_half_dot(const double *a, const double *b, size_t n, double *scratch, double *out, cudaStream_t st) {
  reduce(HalfDotF{a, b}, n, scratch, out, st);
}
void launch_kinetic(const double *p, const double *conv, const double *mass_r, size_t n, double *scratch,
                    double *out, cudaStream_t st) {
  reduce(KineticF{p, conv, mass_r}, n, scratch, out, st);
}

// ---------------------------------------------------------------------------
// K8: overdensity + likelihood residual + -lnL
//   overdens, massFunctions.cc:30-47
//   Gaussian: gaussian_independent.cpp:24-42 (residual), :80-91 (value)
//   Poisson:  poissonian.cpp:19-34 (residual), :60-73 (value)
// ---------------------------------------------------------------------------
template <bool UNIT>
__device__ __forceinline__ double pow_bias(double x, double e) {
  if constexpr (UNIT) return x;  // biasE == 1 (the reference fixes it, init_par.cc:574-578): no pow() in the kernel
  else return pow(x, e);
}

// VALUE = false: the residual only (gradient evaluations do not use -lnL; its divisions and logs are compiled out)
template <bool UNIT, bool VALUE = true>
struct ResidualEval {
  LikeParams lp;
  double nmean;
  // returns the -lnL term; writes delta and r
  __device__ __forceinline__ double operator()(double rho, double n, double sg, double w, double &delta,
                                               double &r) const {
    delta = __dsub_rn(__ddiv_rn(rho, nmean), 1.0);
    r = 0.0;
    double val = 0.0;
    if (lp.likelihood == 1) {
      const double base = __dadd_rn(1.0, __dmul_rn(lp.biasP, delta));
      const double Lambda = __dmul_rn(__dmul_rn(w, lp.rho_c), pow_bias<UNIT>(base, lp.biasE));
      if (w > 0. && Lambda > 0.0) {
        r = __ddiv_rn(__dsub_rn(n, Lambda), __dmul_rn(sg, sg));
        if constexpr (VALUE) {
          const double q = __ddiv_rn(__dsub_rn(Lambda, n), sg);
          val = __dmul_rn(0.5, __dmul_rn(q, q));
        }
      }
    } else if (lp.likelihood == 2) {
      // log-normal (lognormal_independent.cpp:41-55, 57-62, 112-122): the residual takes the log of the
      // unclamped density (NaN / inf where 1 + delta <= 0, as in the reference), the value clamps at delta_min
      if (w > 0.) {
        const double base = __dadd_rn(1.0, __dmul_rn(lp.biasP, delta));
        const double Lr = log(__dmul_rn(lp.rho_c, pow_bias<UNIT>(base, lp.biasE)));
        const double dl = delta < lp.delta_min ? lp.delta_min : delta;
        const double Lv = log(__dmul_rn(lp.rho_c, __dadd_rn(1.0, dl)));
        r = __ddiv_rn(__dsub_rn(n, Lr), __dmul_rn(sg, sg));
        if (lp.exact_sign)  // exact adjoint: -d(-lnL)/d delta of the clamped value, chain factor 1/(1 + delta) included
          r = delta < lp.delta_min ? 0.0 : __ddiv_rn(__ddiv_rn(__dsub_rn(n, Lv), __dmul_rn(sg, sg)), __dadd_rn(1.0, delta));
        if constexpr (VALUE) {
          const double q = __dsub_rn(Lv, n);
          val = __ddiv_rn(__dmul_rn(__dmul_rn(0.5, q), q), __dmul_rn(sg, sg));
        }
      }
    } else {
      const double dens = __dadd_rn(1.0, __dmul_rn(lp.biasP, delta));
      const double Lambda = __dmul_rn(__dmul_rn(w, lp.rho_c), pow_bias<UNIT>(dens, lp.biasE));
      if (w > 0.0 && dens > 0.0) {
        r = (1 - n / Lambda) * lp.rho_c * lp.biasE * lp.biasP * (UNIT ? 1.0 : pow_bias<false>(dens, lp.biasE - 1));
        if (lp.exact_sign) r = -r;
      }
      if constexpr (VALUE) {
        if (w > 0. && Lambda > 0.0) val = __dsub_rn(Lambda, __dmul_rn(n, log(Lambda)));
      }
    }
    return val;
  }
};

// two elements per thread per trip (16-byte loads), two trips in flight.  VALUE = false (gradient evaluations):
// no -lnL partial sums; `keep_delta` = false additionally leaves the density array alone (only calc_h = 0 reads
// delta_x again after the residual), one array less to write.
template <bool UNIT, bool VALUE>
__global__ void __launch_bounds__(kReduceThreads, 4)
    overdens_residual_kernel(LikeParams lp, double2 *__restrict__ rho_delta, const double *__restrict__ sum_rho,
                             const double2 *__restrict__ nobs, const double2 *__restrict__ noise,
                             const double2 *__restrict__ window, double2 *__restrict__ resid, size_t n2, double count,
                             double *__restrict__ part, int keep_delta) {
  ResidualEval<UNIT, VALUE> ev{lp, __ddiv_rn(*sum_rho, count)};
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * kReduceThreads;
  size_t i = (size_t)blockIdx.x * kReduceThreads + threadIdx.x;
  for (; i + stride < n2; i += 2 * stride) {
    const size_t j = i + stride;
    const double2 ra = rho_delta[i], rb = rho_delta[j];
    const double2 na = nobs[i], nb = nobs[j];
    const double2 sa = noise[i], sb = noise[j];
    const double2 wa = window[i], wb = window[j];
    double2 da, db, qa, qb;
    acc += ev(ra.x, na.x, sa.x, wa.x, da.x, qa.x);
    acc += ev(ra.y, na.y, sa.y, wa.y, da.y, qa.y);
    acc += ev(rb.x, nb.x, sb.x, wb.x, db.x, qb.x);
    acc += ev(rb.y, nb.y, sb.y, wb.y, db.y, qb.y);
    if (keep_delta) {
      rho_delta[i] = da;
      rho_delta[j] = db;
    }
    if (resid) {
      resid[i] = qa;
      resid[j] = qb;
    }
  }
  for (; i < n2; i += stride) {
    const double2 ra = rho_delta[i], na = nobs[i], sa = noise[i], wa = window[i];
    double2 da, qa;
    acc += ev(ra.x, na.x, sa.x, wa.x, da.x, qa.x);
    acc += ev(ra.y, na.y, sa.y, wa.y, da.y, qa.y);
    if (keep_delta) rho_delta[i] = da;
    if (resid) resid[i] = qa;
  }
  if constexpr (VALUE) {
    const double r = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = r;
  }
}

void launch_overdens_residual(const LikeParams &lp, double *rho_delta, const double *sum_rho, const double *nobs,
                              const double *noise, const double *window, double *resid, size_t n,
                              double ncells_global, double *scratch, double *nll, cudaStream_t st, bool keep_delta) {
  ProfScope prof(KK_RESIDUAL, st);
  const size_t n2 = n / 2;  // n = N^3 with N a power of two >= 8
  const int blocks = (int)((n2 + kReduceThreads - 1) / kReduceThreads < (size_t)kReduceBlocks
                               ? (n2 + kReduceThreads - 1) / kReduceThreads
                               : (size_t)kReduceBlocks);
  const bool unit = lp.biasE == 1.0;
  auto kern = nll ? (unit ? overdens_residual_kernel<true, true> : overdens_residual_kernel<false, true>)
                  : (unit ? overdens_residual_kernel<true, false> : overdens_residual_kernel<false, false>);
  kern<<<blocks, kReduceThreads, 0, st>>>(
      lp, reinterpret_cast<double2 *>(rho_delta), sum_rho, reinterpret_cast<const double2 *>(nobs),
      reinterpret_cast<const double2 *>(noise), reinterpret_cast<const double2 *>(window),
      reinterpret_cast<double2 *>(resid), n2, ncells_global, scratch, keep_delta ? 1 : 0);
  if (nll) final_sum_kernel<<<1, kReduceThreads, 0, st>>>(scratch, blocks, nll);
  BGPU_LAUNCHED(nll ? 2 : 1);
}

__global__ void lognormal_f_kernel(const double *__restrict__ delta, double *__restrict__ out, size_t n, double rho_c,
                                   double delta_min) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = delta[i];
  if (d < delta_min) d = delta_min;
  out[i] = log(__dmul_rn(rho_c, __dadd_rn(1.0, d)));
}

void launch_lognormal_f(const double *delta, double *out, size_t n, double rho_c, double delta_min, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  lognormal_f_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(delta, out, n, rho_c, delta_min);
  BGPU_LAUNCHED(1);
}

__global__ void grf_grad_add_kernel(double *__restrict__ grad, const double *__restrict__ s,
                                    const double *__restrict__ nobs, const double *__restrict__ noise,
                                    const double *__restrict__ window, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (window[i] > 0.) grad[i] = __dadd_rn(grad[i], __ddiv_rn(__dsub_rn(s[i], nobs[i]), __dmul_rn(noise[i], noise[i])));
}

void launch_grf_grad_add(double *grad, const double *s, const double *nobs, const double *noise, const double *window,
                         size_t n, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  grf_grad_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(grad, s, nobs, noise, window, n);
  BGPU_LAUNCHED(1);
}

struct GrfNll {
  const double *s, *nobs, *noise, *window;
  __device__ __forceinline__ double operator()(size_t i) const {
    if (!(window[i] > 0.)) return 0.0;
    const double q = __ddiv_rn(__dsub_rn(s[i], nobs[i]), noise[i]);
    return __dmul_rn(0.5, __dmul_rn(q, q));
  }
};

void launch_grf_nll(const double *s, const double *nobs, const double *noise, const double *window, size_t n,
                    double *scratch, double *out, cudaStream_t st) {
  reduce(GrfNll{s, nobs, noise, window}, n, scratch, out, st);
}

// ---------------------------------------------------------------------------
// K9: exact adjoint of the mass assignment (new).  One thread per
// particle, pure gather: V_c = sum_cells r_cell * dW_cell/dx_c.  Deterministic.
// The particle's own displacement is read and its V written in place.
// ---------------------------------------------------------------------------
// Under the cell-boundary averaging of the 2LPT model (g.cellbound) the positions come from Psi averaged with
// its (i-1, j-1, k-1) neighbour, so V must not overwrite Psi: (ox, oy, oz) are then separate arrays.
__global__ void gather_adjoint_kernel(GridGeom g, const double *ax, const double *ay, const double *az,
                                      double *ox, double *oy, double *oz, const double *__restrict__ resid) {
  const int N = g.N;
  const size_t n = (size_t)g.Ns * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);  // N is a power of two
  const int k = (int)(idx & (size_t)(N - 1));
  const int j = (int)((idx >> sh) & (size_t)(N - 1));
  const int i = g.x0 + (int)(idx >> (2 * sh));
  // plane of the (halo-extended, slab-local) residual tile that global cell plane c maps to (a cube: c itself);
  // the scatter of the same positions has already checked that every cell lies inside the halo
  const int lo_plane = g.x0 - g.H;
  auto plane = [&](int c) {
    int l = c - lo_plane;
    if (l < 0) l += N;
    if (l >= N) l -= N;
    return l;
  };
  double x, y, z;
  double px = ax[idx], py = ay[idx], pz = az[idx];
  if (g.cellbound) {  // cube only (the slab path keeps to the Zel'dovich model for the exact adjoint)
    const int il = (int)(idx >> (2 * sh));
    const size_t m = ((size_t)(il == 0 ? N - 1 : il - 1) * N + (j == 0 ? N - 1 : j - 1)) * N + (k == 0 ? N - 1 : k - 1);
    px = 0.5 * (ax[m] + px);
    py = 0.5 * (ay[m] + py);
    pz = 0.5 * (az[m] + pz);
  }
  particle_position(g, i, j, k, px, py, pz, x, y, z);
  double vx = 0.0, vy = 0.0, vz = 0.0;
  if (in_domain(g, x, y, z)) {
    const double inv_d = 1.0 / g.d;
    if (g.masskernel == 1) {
      int ci[2], cj[2], ck[2];
      double wi[2], wj[2], wk[2];
      cic_axis(x, g.d, g.L, N, ci[0], ci[1], wi[0], wi[1]);
      cic_axis(y, g.d, g.L, N, cj[0], cj[1], wj[0], wj[1]);
      cic_axis(z, g.d, g.L, N, ck[0], ck[1], wk[0], wk[1]);
      ci[0] = plane(ci[0]);
      ci[1] = plane(ci[1]);
      const double gsgn[2] = {-inv_d, inv_d};
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const size_t row = ((size_t)ci[a] * N + cj[b]) * N;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const double rc = __ldg(resid + row + ck[c]);
            vx += rc * gsgn[a] * wj[b] * wk[c];
            vy += rc * wi[a] * gsgn[b] * wk[c];
            vz += rc * wi[a] * wj[b] * gsgn[c];
          }
        }
    } else if (g.masskernel == 2) {
      int ci[3], cj[3], ck[3];
      double wi[3], wj[3], wk[3], dx, dy, dz;
      tsc_axis(x, g.min1, g.d, N, ci, wi, dx);
      tsc_axis(y, g.min2, g.d, N, cj, wj, dy);
      tsc_axis(z, g.min3, g.d, N, ck, wk, dz);
#pragma unroll
      for (int a = 0; a < 3; ++a) ci[a] = plane(ci[a]);
      // d/dx of (1/2 (1/2 - D)^2, 3/4 - D^2, 1/2 (1/2 + D)^2), D = x/d - (i + 1/2)
      const double gi[3] = {-(0.5 - dx) * inv_d, -2.0 * dx * inv_d, (0.5 + dx) * inv_d};
      const double gj[3] = {-(0.5 - dy) * inv_d, -2.0 * dy * inv_d, (0.5 + dy) * inv_d};
      const double gk[3] = {-(0.5 - dz) * inv_d, -2.0 * dz * inv_d, (0.5 + dz) * inv_d};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const size_t row = ((size_t)ci[a] * N + cj[b]) * N;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const double rc = __ldg(resid + row + ck[c]);
            vx += rc * gi[a] * wj[b] * wk[c];
            vy += rc * wi[a] * gj[b] * wk[c];
            vz += rc * wi[a] * wj[b] * gk[c];
          }
        }
    }  // NGP: piecewise-constant weights, derivative identically zero
  }
  if (g.rsd) vz += g.fgrow * vz;  // d z_s / d Psi_z = 1 + f (cf. HMC_models.cc:295-301)
  ox[idx] = vx;
  oy[idx] = vy;
  oz[idx] = vz;
}

void launch_gather_adjoint(const GridGeom &g, double *ax, double *ay, double *az, const double *resid,
                           cudaStream_t st) {
  if (gather_sweep_applicable(g)) {
    launch_gather_sweep(g, ax, ay, az, resid, st);
    return;
  }
  ProfScope prof(KK_GATHER, st);
  const size_t n = (size_t)g.Ns * g.N * g.N;
  gather_adjoint_kernel<<<blocks_for(n, 256), 256, 0, st>>>(g, ax, ay, az, ax, ay, az, resid);
  BGPU_LAUNCHED(1);
}

void launch_gather_adjoint_to(const GridGeom &g, const double *psix, const double *psiy, const double *psiz, double *vx,
                              double *vy, double *vz, const double *resid, cudaStream_t st) {
  ProfScope prof(KK_GATHER, st);
  const size_t n = (size_t)g.Ns * g.N * g.N;
  gather_adjoint_kernel<<<blocks_for(n, 256), 256, 0, st>>>(g, psix, psiy, psiz, vx, vy, vz, resid);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// K5: r * d_c(delta), 4th-order central difference, gradient.cpp:81-153
// ---------------------------------------------------------------------------
// Slab: `in` points at the first OWNED plane of an array that carries xo >= 2 valid planes of the x neighbours on
// each side (no wrap along x); a cube passes Ns = N, xo = 0 and wraps.
__global__ void findif_product_kernel(const double *__restrict__ in, const double *__restrict__ resid,
                                      double *__restrict__ out, int N, int Ns, int xo, double fac, int comp) {
  const size_t n = (size_t)Ns * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);  // N is a power of two
  int c[3] = {(int)(idx >> (2 * sh)), (int)((idx >> sh) & (size_t)(N - 1)), (int)(idx & (size_t)(N - 1))};
  const ptrdiff_t stride = comp == 0 ? (ptrdiff_t)N * N : (comp == 1 ? (ptrdiff_t)N : 1);
  const int ii = c[comp];
  const ptrdiff_t base = (ptrdiff_t)idx - (ptrdiff_t)ii * stride;
  int r = ii + 1, rr = ii + 2, l = ii - 1, ll = ii - 2;
  if (!(comp == 0 && xo)) {  // periodic along this axis
    r = r >= N ? r - N : r;
    rr = rr >= N ? rr - N : rr;
    l = l < 0 ? l + N : l;
    ll = ll < 0 ? ll + N : ll;
  }
  const double g = -(fac * ((4.0 / 3) * (in[base + l * stride] - in[base + r * stride]) -
                            (1.0 / 6) * (in[base + ll * stride] - in[base + rr * stride])));
  out[idx] = resid[idx] * g;
}

void launch_findif_product(const double *delta, const double *resid, double *out, int N, int Ns, int xo, double L,
                           int comp, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const size_t n = (size_t)Ns * N * N;
  const double fac = (double)N / (2. * L);
  findif_product_kernel<<<blocks_for(n, 256), 256, 0, st>>>(delta, resid, out, N, Ns, xo, fac, comp);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// counter-based Gaussian generator: Philox4x32-10 (Salmon et al. 2011; constants and
// known-answer vectors of Random123) + Box-Muller.  Element pair (2i, 2i+1) of draw `draw`, stream
// `stream` comes from counter {i_lo, i_hi, draw_lo, draw_hi ^ stream << 24} under key {seed_lo, seed_hi}:
// any element of any draw can be regenerated independently, on any number of GPUs.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// 53-bit uniform in (0, 1) from two 32-bit words
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void philox_normal_kernel(double2 *__restrict__ out, size_t npairs, size_t pair0, uint64_t seed, uint64_t draw,
                                     uint32_t stream) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npairs) return;
  const uint64_t g = (uint64_t)(pair0 + i);
  uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)draw, (uint32_t)(draw >> 32) ^ (stream << 24)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u1 = u53(c[0], c[1]), u2 = u53(c[2], c[3]);
  const double r = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  out[i] = make_double2(r * cs, r * sn);
}

void launch_philox_normals(double *out, size_t n, size_t first, uint64_t seed, uint64_t draw, unsigned stream,
                           cudaStream_t st) {
  ProfScope prof(KK_COLOUR, st);
  const size_t npairs = n / 2;  // n and first are even
  philox_normal_kernel<<<blocks_for(npairs, 256), 256, 0, st>>>(reinterpret_cast<double2 *>(out), npairs, first / 2, seed,
                                                               draw, stream);
  BGPU_LAUNCHED(1);
}

// colour the transform of a real white field w (<|w^|^2> = N): A = w^ sqrt(N/V M) has <|A|^2> = N^2/V M, the
// variance create_GARFIELD gives its modes (random.cpp:81-83); sigma is read at the folded index and the DC
// mode is zeroed as there (:102-135, :347-351).
__global__ void colour_white_kernel(double2 *__restrict__ W, const double *__restrict__ spec, int N, double c2) {
  const int nzh = N / 2 + 1;
  const size_t n = (size_t)N * N * nzh;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int k = (int)(idx % nzh);
  const int j = (int)((idx / nzh) % N);
  const int i = (int)(idx / ((size_t)nzh * N));
  const int fi = i <= N / 2 ? i : N - i, fj = j <= N / 2 ? j : N - j;
  double a = sqrt(c2 * spec[((size_t)fi * N + fj) * N + k]);
  if ((i | j | k) == 0 || !(a == a)) a = 0.0;
  const double2 w = W[idx];
  W[idx] = make_double2(a * w.x, a * w.y);
}

// The same colouring from the half-grid multiplier the kinetic term uses, inv = (V/N)/M (0 where M <= 0):
// (N/V) M = 1/inv, so A = w^ / sqrt(inv).  Works on any row layout with the multiplier's padded pitch
// (cube [x][y][.] and the transposed slab [x][y_local][.] alike) and keeps the draw and K consistent.
__global__ void colour_white_rows_kernel(double2 *__restrict__ W, const double *__restrict__ inv, int nzh, size_t n,
                                         int zero_dc) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const size_t row = idx / nzh;
  const int z = (int)(idx - row * nzh);
  const double m = inv[row * (nzh + 1) + z];
  // 1 / sqrt (correctly rounded operations, not rsqrt): the cube and every slab decomposition round alike.
  // The DC mode carries no momentum (random.cpp:347-351), whatever M(0) is: element 0 of the rank that owns y = 0.
  double a = m > 0.0 ? __ddiv_rn(1.0, __dsqrt_rn(m)) : 0.0;
  if (zero_dc && idx == 0) a = 0.0;
  const double2 w = W[idx];
  W[idx] = make_double2(a * w.x, a * w.y);
}

void launch_colour_white_rows(double2 *W, const double *inv_half, int N, size_t n_half, bool owns_dc, cudaStream_t st) {
  ProfScope prof(KK_COLOUR, st);
  colour_white_rows_kernel<<<blocks_for(n_half, 256), 256, 0, st>>>(W, inv_half, N / 2 + 1, n_half, owns_dc ? 1 : 0);
  BGPU_LAUNCHED(1);
}

void launch_colour_white(double2 *W, const double *spec_full, int N, double c2, cudaStream_t st) {
  ProfScope prof(KK_COLOUR, st);
  const size_t n = (size_t)N * N * (N / 2 + 1);
  colour_white_kernel<<<blocks_for(n, 256), 256, 0, st>>>(W, spec_full, N, c2);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// F3: measure_spectrum (field_statistics.cpp:20-90): spherically binned power of a field's transform.
// The reference loops over the FULL complex grid; here one thread takes one half-grid mode and counts it
// twice where its mirror (N-i, N-j, N-k) is not in the half array (0 < k < N/2): same |k|, same |F|^2.
// acc = [power | kmode | nmode], 3 * nbin doubles, zeroed by the launcher.
// ---------------------------------------------------------------------------
__global__ void spectrum_bin_kernel(const double2 *__restrict__ F, int N, int Ns, int y0, double kfac, double dk, int nbin,
                                    double *__restrict__ acc) {
  extern __shared__ double sh_acc[];  // 3 * nbin
  for (int b = threadIdx.x; b < 3 * nbin; b += blockDim.x) sh_acc[b] = 0.0;
  __syncthreads();
  const int nzh = N / 2 + 1;
  const size_t n = (size_t)N * Ns * nzh;   // a cube: Ns = N, y0 = 0; a slab: the transposed layout [x][y_local][z]
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % nzh);
    const int j = y0 + (int)((idx / nzh) % Ns);
    const int i = (int)(idx / ((size_t)nzh * Ns));
    auto kv = [&](int m) { return (m <= N / 2) ? kfac * (double)m : -kfac * (double)(N - m); };  // scale_space.cpp:41-51
    const double kx = kv(i), ky = kv(j), kz = kv(k);
    // k_squared (scale_space.cpp:16-39) without FMA contraction: modes that sit exactly on a bin edge
    // (|k|/dk an integer up to rounding) must fall on the reference's side of it
    const double ktot = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)), __dmul_rn(kz, kz)));
    const unsigned long long b = (unsigned long long)__ddiv_rn(ktot, dk);
    if (b < (unsigned long long)nbin) {
      const double w = (k > 0 && k < N / 2) ? 2.0 : 1.0;
      const double2 f = F[idx];
      atomicAdd(sh_acc + b, w * (f.x * f.x + f.y * f.y));
      atomicAdd(sh_acc + nbin + b, w * ktot);
      atomicAdd(sh_acc + 2 * nbin + b, w);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 3 * nbin; b += blockDim.x)
    if (sh_acc[b] != 0.0) atomicAdd(acc + b, sh_acc[b]);
}

__global__ void spectrum_finish_kernel(double *__restrict__ acc, int nbin, double norm) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbin) return;
  const double nm = acc[2 * nbin + b];
  if (nm > 0.0) {
    acc[nbin + b] = acc[nbin + b] / nm;
    acc[b] = acc[b] / nm * norm;
  }
}

// bin this rank's modes into acc (zeroed here); a slab all-reduces acc over the ranks before the finish
void launch_measure_spectrum_bin(const double2 *F, int N, int Ns, int y0, double L, int nbin, double *acc,
                                 cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const double kfac = 2.0 * M_PI / L;
  const double kny = kfac * (double)(N / 2);
  const double dk = std::sqrt((kny * kny + kny * kny) + kny * kny) / (double)nbin;   // kmax = |k| of (N/2, N/2, N/2)
  BGPU_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 3 * nbin, st));
  const size_t n = (size_t)N * Ns * (N / 2 + 1);
  const int blocks = (int)(blocks_for(n, 256) < 1184u ? blocks_for(n, 256) : 1184u);
  spectrum_bin_kernel<<<blocks, 256, sizeof(double) * 3 * nbin, st>>>(F, N, Ns, y0, kfac, dk, nbin, acc);
  BGPU_LAUNCHED(1);
}

void launch_measure_spectrum_finish(double *acc, int N, double L, int nbin, cudaStream_t st) {
  const double V = L * L * L, nn = (double)N * N * N;
  spectrum_finish_kernel<<<(nbin + 255) / 256, 256, 0, st>>>(acc, nbin, V / nn / nn);  // FOURIER_DEF_2 norm
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// streaming helpers
// ---------------------------------------------------------------------------
__global__ void inverse_spectrum_kernel(const double *__restrict__ full, double *__restrict__ half, int N,
                                        double normFS) {
  // rows of the half-grid multiplier are padded to N/2+2 doubles so that the row pitch is a
  // multiple of 16 bytes (a TMA tensor-map requirement, fft_tma.cuh); the pad element is 0
  const int nzp = N / 2 + 2;
  const size_t n = (size_t)N * N * nzp;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int k = (int)(idx % nzp);
  const size_t ij = idx / nzp;
  if (k > N / 2) {
    half[idx] = 0.;
    return;
  }
  const double c = full[ij * N + k];  // HMC_help.cc:44: corrFunc[k + N3*(j + N2*i)]
  half[idx] = c > 0.0 ? normFS / c : 0.;
}

void launch_inverse_spectrum(const double *full, double *half, int N, double normFS, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const size_t n = (size_t)N * N * (N / 2 + 2);
  inverse_spectrum_kernel<<<blocks_for(n, 256), 256, 0, st>>>(full, half, N, normFS);
  BGPU_LAUNCHED(1);
}

__global__ void axpy_kernel(double *__restrict__ y, const double *__restrict__ x, double a, size_t n,
                            const int *__restrict__ skip) {
  if (skip && *skip) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));  // HMC.cc:294,339,352 (two roundings)
}
// HMC.cc:360-364: a trajectory whose momentum has run away is stopped -- on the device by raising a flag that
// every later update of the trajectory honours (ROp::skip, launch_axpy's skip).  The reference tests momenta[0]
// after the second half kick of every step.  With merged kicks that value is never stored, but it is the midpoint
// of momenta[0] before and after the merged kick (p - eps/2 g between p and p - eps g), which this kernel keeps.
//   mode 0: remember momenta[0] (after the trajectory's first half kick; the reference does not test there)
//   mode 1: after a merged kick (both half kicks of a step boundary): test the midpoint
//   mode 2: after a plain half kick (the last step): test momenta[0] itself
// scal[0] = the step at which the trajectory stopped (0 = running), scal[1] = momenta[0] after the previous kick.
__global__ void runaway_guard_kernel(const double *__restrict__ p, double *__restrict__ scal, int *__restrict__ flag,
                                     int step, int mode) {
  const double now = p[0];
  if (mode != 0 && scal[0] == 0.0) {
    const double seen = mode == 1 ? 0.5 * (scal[1] + now) : now;
    if (fabs(seen) > 1e50) {
      scal[0] = (double)step;
      if (flag) *flag = step;
    }
  }
  scal[1] = now;
}
// slab chains: rank 0 owns momenta[0]; its verdict reaches the others by an all-reduce of scal[0]
__global__ void runaway_apply_kernel(const double *__restrict__ scal, int *__restrict__ flag) {
  if (*flag == 0 && scal[0] > 0.0) *flag = (int)scal[0];
}
__global__ void scale_kernel(double *__restrict__ y, const double *__restrict__ x, double a, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i];
}
__global__ void axpy_div_kernel(double *__restrict__ y, const double *__restrict__ x, const double *__restrict__ m,
                                double a, size_t n, const int *__restrict__ skip) {
  if (skip && *skip) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double mm = m[i];
    const double invM = mm > 0.0 ? 1.0 / mm : 0.0;  // HMC.cc:321-326
    y[i] += a * (x[i] * invM);
  }
}
__global__ void fill_kernel(double *__restrict__ y, double v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = v;
}

void launch_axpy(double *y, const double *x, double a, size_t n, cudaStream_t st, const int *skip) {
  ProfScope prof(KK_STREAM, st);
  axpy_kernel<<<blocks_for(n, 256), 256, 0, st>>>(y, x, a, n, skip);
  BGPU_LAUNCHED(1);
}
void launch_runaway_guard(const double *p, double *scal2, int *flag, int step, int mode, cudaStream_t st) {
  runaway_guard_kernel<<<1, 1, 0, st>>>(p, scal2, flag, step, mode);
  BGPU_LAUNCHED(1);
}
void launch_runaway_apply(const double *scal2, int *flag, cudaStream_t st) {
  runaway_apply_kernel<<<1, 1, 0, st>>>(scal2, flag);
  BGPU_LAUNCHED(1);
}
void launch_scale(double *y, const double *x, double a, size_t n, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  scale_kernel<<<blocks_for(n, 256), 256, 0, st>>>(y, x, a, n);
  BGPU_LAUNCHED(1);
}
void launch_axpy_div(double *y, const double *x, const double *m, double a, size_t n, cudaStream_t st, const int *skip) {
  ProfScope prof(KK_STREAM, st);
  axpy_div_kernel<<<blocks_for(n, 256), 256, 0, st>>>(y, x, m, a, n, skip);
  BGPU_LAUNCHED(1);
}
void launch_fill(double *y, double v, size_t n, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  fill_kernel<<<blocks_for(n, 256), 256, 0, st>>>(y, v, n);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// A17: Hamiltonian mass types 0 / 1 / 4, HMC_mass.cc:117-124,163-172,315-368
// ---------------------------------------------------------------------------
__global__ void mass_kernel(const double *__restrict__ power, double *__restrict__ mass_f,
                            double *__restrict__ mass_r, int type, double factor, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (type == 0) {
    mass_r[i] = 1.0;
  } else if (type == 1) {
    const double P = power[i];
    const double invP = P > 0.0 ? 1. / P : 0.;
    mass_f[i] = factor * (1.0 * invP);
  } else if (type == 4) {
    mass_f[i] = factor * power[i];
  }
}

// mass types 2 / 3 (HMC_mass.cc:52-160): mass_f = factor (2/P + sqrt(F/P)) with F = the likelihood-force spectrum
// at the cell's |k| bin (type 2; 0 at k = 0, and at the one corner mode whose bin index equals N_bin, where the
// reference reads past its array) or its mean over k-space shells (type 3, `mean`)
__global__ void force_mass_kernel(const double *__restrict__ power, const double *__restrict__ force_spec,
                                  double *__restrict__ mass_f, int N, int Ns, int x0, double kfac, double dk, int nbin,
                                  int type, double mean, double factor) {
  const size_t n = (size_t)Ns * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const double P = power[idx];
  const double invP = P > 0.0 ? 1. / P : 0.;
  double F = mean;
  if (type == 2) {
    const int sh = 31 - __clz(N);
    const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)), i = x0 + (int)(idx >> (2 * sh));
    auto kv = [&](int m) { return (m <= N / 2) ? kfac * (double)m : -kfac * (double)(N - m); };
    const double kx = kv(i), ky = kv(j), kz = kv(k);
    const double kr = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)), __dmul_rn(kz, kz)));
    const unsigned long long b = (unsigned long long)__ddiv_rn(kr, dk);
    F = (kr > 0. && b < (unsigned long long)nbin) ? force_spec[b] : 0.0;
  }
  mass_f[idx] = factor * (1.0 * (2 * invP + sqrt(invP * F)));
}

void launch_force_mass(const double *power, const double *force_spec, double *mass_f, int N, int Ns, int x0, double L,
                       int nbin, int type, double mean, double factor, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const double kfac = 2.0 * M_PI / L;
  const double kny = kfac * (double)(N / 2);
  const double dk = std::sqrt((kny * kny + kny * kny) + kny * kny) / (double)nbin;
  const size_t n = (size_t)Ns * N * N;
  force_mass_kernel<<<blocks_for(n, 256), 256, 0, st>>>(power, force_spec, mass_f, N, Ns, x0, kfac, dk, nbin, type, mean,
                                                        factor);
  BGPU_LAUNCHED(1);
}

void launch_mass(const double *power, double *mass_f, double *mass_r, int mass_type, double mass_factor, size_t n,
                 cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  mass_kernel<<<blocks_for(n, 256), 256, 0, st>>>(power, mass_f, mass_r, mass_type, mass_factor, n);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// K13: create_GARFIELD's colouring and Hermitian symmetrisation, random.cpp:
// 102-507.  The reference walks the folded octant (i,j,k <= N/2) serially
// through 27 cases, but every half-array element (c <= N/2) is written exactly
// once, so the loop is data parallel: one thread per half-array element picks
// its source entry of the white-noise grid, conjugates it if it is the mirror
// partner, and scales by sigma = sqrt(N^2/V * spec/2) read at the folded index.
// ---------------------------------------------------------------------------
__global__ void colour_momenta_kernel(const double2 *__restrict__ W, const double *__restrict__ spec,
                                      double2 *__restrict__ half, int N, double amp) {
  const int h = N / 2, nzh = h + 1;
  const size_t n = (size_t)N * N * nzh;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int c = (int)(idx % nzh);
  const int b = (int)((idx / nzh) % N);
  const int a = (int)(idx / ((size_t)nzh * N));
  const int fa = a > h ? N - a : a, fb = b > h ? N - b : b;
  const double sigma = sqrt(amp * spec[((size_t)fa * N + fb) * N + c] / 2.0);
  const bool a_fix = (a == 0 || a == h), b_fix = (b == 0 || b == h), c_fix = (c == 0 || c == h);
  const int ma = (N - a) % N, mb = (N - b) % N;
  double2 out;
  if (a_fix && b_fix && c_fix) {
    if (a == 0 && b == 0 && c == 0) {
      out = make_double2(0.0, 0.0);                                 // random.cpp:347-351
    } else {
      const double2 w = W[((size_t)a * N + b) * N + c];
      out = make_double2(w.x * (sqrt(2.0) * sigma), 0.0);           // :243-247, 439-483
    }
  } else {
    bool mirror;
    int sa = a, sb = b, sc = c;
    if (!c_fix) {
      mirror = (!a_fix && !b_fix && a > h && b > h);                // :136-140
      if (mirror) { sa = ma; sb = mb; sc = N - c; }
    } else {
      mirror = (!b_fix && b > h) || (b_fix && !a_fix && a > h);     // :144-162, 251-268, 308-345, 355-409
      if (mirror) { sa = ma; sb = mb; }
    }
    const double2 w = W[((size_t)sa * N + sb) * N + sc];
    out = make_double2(w.x * sigma, mirror ? -(w.y * sigma) : w.y * sigma);
  }
  half[idx] = out;
}

void launch_colour_momenta(const double2 *white_full, const double *spec_full, double2 *half, int N, double amp,
                           cudaStream_t st) {
  ProfScope prof(KK_COLOUR, st);
  const size_t n = (size_t)N * N * (N / 2 + 1);
  colour_momenta_kernel<<<blocks_for(n, 256), 256, 0, st>>>(white_full, spec_full, half, N, amp);
  BGPU_LAUNCHED(1);
}

// The same draw on a slab-decomposed chain: this rank colours its rows of the transposed k-space layout
// [x][y_local][z <= N/2] (y = y0 + y_local) from the FULL white-noise grid -- a mode's source entry, or its mirror
// partner's, can sit anywhere in it -- with sigma^2 = (N^2/V) M / 2 = N / (2 inv), inv = (V/N)/M being the padded
// half-grid multiplier of the kinetic term (the full spectrum M is x-slab decomposed and not addressable here).
__global__ void colour_momenta_rows_kernel(const double2 *__restrict__ W, const double *__restrict__ inv,
                                           double2 *__restrict__ half, int N, int Ns, int y0, double ncells) {
  const int h = N / 2, nzh = h + 1;
  const size_t n = (size_t)N * Ns * nzh;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const size_t row = idx / nzh;
  const int c = (int)(idx - row * nzh);
  const int b = y0 + (int)(row % Ns);
  const int a = (int)(row / Ns);
  const double m = inv[row * (nzh + 1) + c];
  const double sigma = m > 0.0 ? sqrt(ncells / (2.0 * m)) : 0.0;
  const bool a_fix = (a == 0 || a == h), b_fix = (b == 0 || b == h), c_fix = (c == 0 || c == h);
  const int ma = (N - a) % N, mb = (N - b) % N;
  double2 out;
  if (a_fix && b_fix && c_fix) {
    if (a == 0 && b == 0 && c == 0) {
      out = make_double2(0.0, 0.0);                                 // random.cpp:347-351
    } else {
      const double2 w = W[((size_t)a * N + b) * N + c];
      out = make_double2(w.x * (sqrt(2.0) * sigma), 0.0);           // :243-247, 439-483
    }
  } else {
    bool mirror;
    int sa = a, sb = b, sc = c;
    if (!c_fix) {
      mirror = (!a_fix && !b_fix && a > h && b > h);                // :136-140
      if (mirror) { sa = ma; sb = mb; sc = N - c; }
    } else {
      mirror = (!b_fix && b > h) || (b_fix && !a_fix && a > h);     // :144-162, 251-268, 308-345, 355-409
      if (mirror) { sa = ma; sb = mb; }
    }
    const double2 w = W[((size_t)sa * N + sb) * N + sc];
    out = make_double2(w.x * sigma, mirror ? -(w.y * sigma) : w.y * sigma);
  }
  half[idx] = out;
}

void launch_colour_momenta_rows(const double2 *white_full, const double *inv_half, double2 *half, int N, int Ns, int y0,
                                double ncells, cudaStream_t st) {
  ProfScope prof(KK_COLOUR, st);
  const size_t n = (size_t)N * Ns * (N / 2 + 1);
  colour_momenta_rows_kernel<<<blocks_for(n, 256), 256, 0, st>>>(white_full, inv_half, half, N, Ns, y0, ncells);
  BGPU_LAUNCHED(1);
}

__global__ void add_real_momenta_kernel(double *__restrict__ p, const double *__restrict__ mass_r,
                                        const double *__restrict__ gauss, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += sqrt(mass_r[i]) * gauss[i];
}

void launch_add_real_momenta(double *p, const double *mass_r, const double *gauss, size_t n, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  add_real_momenta_kernel<<<blocks_for(n, 256), 256, 0, st>>>(p, mass_r, gauss, n);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// slab helpers: halo sizing (max |Psi_x|), halo accumulation, transposed inverse spectrum
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kReduceThreads) partial_max_abs_kernel(const double *__restrict__ a, size_t n,
                                                                         double *__restrict__ part) {
  __shared__ double wp[kReduceThreads / 32];
  double v = 0.0;
  for (size_t i = (size_t)blockIdx.x * kReduceThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kReduceThreads)
    v = fmax(v, fabs(a[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) wp[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kReduceThreads / 32; ++w) v = fmax(v, wp[w]);
    part[blockIdx.x] = v;
  }
}
__global__ void final_max_kernel(const double *__restrict__ part, int nparts, double *__restrict__ out) {
  double v = 0.0;
  for (int i = 0; i < nparts; ++i) v = fmax(v, part[i]);
  *out = v;
}

void launch_max_abs(const double *a, size_t n, double *scratch, double *out, cudaStream_t st) {
  ProfScope prof(KK_REDUCE, st);
  const int blocks = (int)((n + kReduceThreads - 1) / kReduceThreads < (size_t)kReduceBlocks
                               ? (n + kReduceThreads - 1) / kReduceThreads
                               : (size_t)kReduceBlocks);
  partial_max_abs_kernel<<<blocks, kReduceThreads, 0, st>>>(a, n, scratch);
  final_max_kernel<<<1, 1, 0, st>>>(scratch, blocks, out);
  BGPU_LAUNCHED(2);
}

__global__ void add_kernel(double *__restrict__ dst, const double *__restrict__ src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
void launch_add(double *dst, const double *src, size_t n, cudaStream_t st) {
  ProfScope prof(KK_HALO, st);
  add_kernel<<<blocks_for(n, 256), 256, 0, st>>>(dst, src, n);
  BGPU_LAUNCHED(1);
}

__global__ void inverse_spectrum_pack_kernel(const double *__restrict__ full, double2 *__restrict__ packed, int N,
                                             int Ns, double normFS) {
  const int nzh = N / 2 + 1;
  const size_t n = (size_t)Ns * N * nzh;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int z = (int)(idx % nzh);
  const int y = (int)((idx / nzh) % N);
  const int xl = (int)(idx / ((size_t)nzh * N));
  const double c = full[((size_t)xl * N + y) * N + z];  // HMC_help.cc:44
  const int peer = y / Ns, yl = y % Ns;
  packed[(((size_t)peer * Ns + xl) * Ns + yl) * nzh + z] = make_double2(c > 0.0 ? normFS / c : 0., 0.);
}
__global__ void inverse_spectrum_unpack_kernel(const double2 *__restrict__ tr, double *__restrict__ half, int N,
                                               int Ns) {
  const int nzh = N / 2 + 1, nzp = N / 2 + 2;
  const size_t n = (size_t)N * Ns * nzp;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int z = (int)(idx % nzp);
  const size_t row = idx / nzp;  // x * Ns + y_l
  half[idx] = z < nzh ? tr[row * nzh + z].x : 0.;
}
void launch_inverse_spectrum_pack(const double *full_xslab, double2 *packed, int N, int Ns, double normFS,
                                  cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const size_t n = (size_t)Ns * N * (N / 2 + 1);
  inverse_spectrum_pack_kernel<<<blocks_for(n, 256), 256, 0, st>>>(full_xslab, packed, N, Ns, normFS);
  BGPU_LAUNCHED(1);
}
void launch_inverse_spectrum_unpack(const double2 *transposed, double *half, int N, int Ns, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const size_t n = (size_t)N * Ns * (N / 2 + 2);
  inverse_spectrum_unpack_kernel<<<blocks_for(n, 256), 256, 0, st>>>(transposed, half, N, Ns);
  BGPU_LAUNCHED(1);
}

// out = a_real * s + norm * h on the (transposed) half grid: the K_FINAL combination as its own pass,
// for the configurations whose strided pass cannot stage both operand tiles (N = 512 on a slab)
// (hh and out may be the same array: no __restrict__ on them)
__global__ void kfinal_combine_kernel(const double2 *__restrict__ s, const double *__restrict__ mult,
                                      const double2 *hh, double2 *out, double norm, int nzh,
                                      size_t n) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const size_t row = idx / nzh;
  const int z = (int)(idx - row * nzh);
  const double f = mult[row * (nzh + 1) + z];
  const double2 a = s[idx], b = hh[idx];
  out[idx] = make_double2(a.x * f + norm * b.x, a.y * f + norm * b.y);
}
void launch_kfinal_combine(const double2 *s, const double *mult, const double2 *h, double2 *out, double norm, int N,
                           size_t n_half, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  kfinal_combine_kernel<<<blocks_for(n_half, 256), 256, 0, st>>>(s, mult, h, out, norm, N / 2 + 1, n_half);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// A9: Lag2Eul_non_zeldovich (Lag2Eul.cc:138-312), the pieces that are not FFTs
// ---------------------------------------------------------------------------
// 4th-order central difference of gradient.cpp:81-153 along one axis at cell (c0, c1, c2), evaluated
// with periodic wrap: -fac * ((4/3)(f[-1] - f[+1]) - (1/6)(f[-2] - f[+2]))
__device__ __forceinline__ int wrap_idx(int i, int N) { return i < 0 ? i + N : (i >= N ? i - N : i); }

// `xo` = planes of x halo in front of the local slab (0 on a cube, which wraps in x instead)
template <int AX>
__device__ __forceinline__ double findif_at(const double *__restrict__ f, int N, int xo, int i, int j, int k,
                                            double fac) {
  auto at = [&](int o) {
    const int ii = AX == 0 ? (xo ? i + o : wrap_idx(i + o, N)) : i, jj = AX == 1 ? wrap_idx(j + o, N) : j,
              kk = AX == 2 ? wrap_idx(k + o, N) : k;
    return f[((size_t)(ii + xo) * N + jj) * N + kk];
  };
  return -(fac * ((4.0 / 3) * (at(-1) - at(1)) - (1.0 / 6) * (at(-2) - at(2))));
}

// second derivative d_B d_A phi at (i, j, k): the same stencil applied twice, as calc_m2v_mem does with
// GFINDIFF (EqSolvers.cc:396-407)
template <int A, int B>
__device__ __forceinline__ double findif2_at(const double *__restrict__ f, int N, int xo, int i, int j, int k,
                                             double fac) {
  auto g = [&](int o) {  // d_A phi at the point shifted by o along B
    const int ii = B == 0 ? (xo ? i + o : wrap_idx(i + o, N)) : i, jj = B == 1 ? wrap_idx(j + o, N) : j,
              kk = B == 2 ? wrap_idx(k + o, N) : k;
    return findif_at<A>(f, N, xo, ii, jj, kk, fac);
  };
  return -(fac * ((4.0 / 3) * (g(-1) - g(1)) - (1.0 / 6) * (g(-2) - g(2))));
}

// out = D1 * (dQ s) - D2 * delta2, delta2 = sum of the 2x2 minors of the Hessian of phi
// (calc_m2v_mem, EqSolvers.cc:373-422; Lag2Eul.cc:196-198)
__global__ void lpt2_source_kernel(const double *__restrict__ phi, const double *__restrict__ s, double *__restrict__ out,
                                   int N, int Ns, int xo, double fac, double dQ, double D1, double D2) {
  const size_t n = (size_t)Ns * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);  // N is a power of two
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)), i = (int)(idx >> (2 * sh));
  const double xx = findif2_at<0, 0>(phi, N, xo, i, j, k, fac), yy = findif2_at<1, 1>(phi, N, xo, i, j, k, fac),
               zz = findif2_at<2, 2>(phi, N, xo, i, j, k, fac);
  const double xy = findif2_at<0, 1>(phi, N, xo, i, j, k, fac), xz = findif2_at<0, 2>(phi, N, xo, i, j, k, fac),
               yz = findif2_at<1, 2>(phi, N, xo, i, j, k, fac);
  const double m2v = xx * yy - xy * xy + xx * zz - xz * xz + yy * zz - yz * yz;
  out[idx] = D1 * (dQ * s[idx]) - D2 * m2v;
}

// spherical-collapse divergence (Lag2Eul.cc:206-223): -3 (sqrt(1 + 2/3 psilin) - 1), psilin = -D1 in,
// and +3 where the root's argument is not positive
__global__ void sc_divergence_kernel(const double *__restrict__ s, double *__restrict__ out, size_t n, double dQ,
                                     double D1) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const double psilin = -D1 * (dQ * s[idx]);
  const double arg = 1. + 2. / 3. * psilin;
  double psisc = arg > 0. ? 3. * (sqrt(arg) - 1.) : -3.;
  out[idx] = -psisc;
}

// C^ = K D2^ + (1 - K) D4^ on the half grid, K = exp(-k^2 rS^2 / 2) (kernelcomp filtertype 1,
// convolution.cpp:224-322; its normalisation, the k = 0 value, is 1): Psi_c = K o Psi^LPT_c +
// Psi^SC_c - K o Psi^SC_c (Lag2Eul.cc:238-268) before the -i k_c / k^2 projection, which is linear
__global__ void alpt_combine_kernel(double2 *__restrict__ d2, const double2 *__restrict__ d4, int N, int Ns, int y0,
                                    double kfac, double rS) {
  const int nzh = N / 2 + 1;
  const size_t n = (size_t)N * Ns * nzh;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int z = (int)(idx % nzh), y = y0 + (int)((idx / nzh) % Ns), x = (int)(idx / ((size_t)nzh * Ns));
  auto kv = [&](int i) { return (i <= N / 2) ? kfac * (double)i : -kfac * (double)(N - i); };
  const double kx = kv(x), ky = kv(y), kz = kv(z);
  const double K = exp(-(kx * kx + ky * ky + kz * kz) * (rS * rS) / 2.);
  const double2 a = d2[idx], b = d4[idx];
  d2[idx] = make_double2(K * a.x + (b.x - K * b.x), K * a.y + (b.y - K * b.y));
}

void launch_lpt2_source(const double *phi, const double *s, double *out, int N, int Ns, int xoff, double L, double dQ,
                        double D1, double D2, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const size_t n = (size_t)Ns * N * N;
  lpt2_source_kernel<<<blocks_for(n, 256), 256, 0, st>>>(phi, s, out, N, Ns, xoff, (double)N / (2. * L), dQ, D1, D2);
  BGPU_LAUNCHED(1);
}
void launch_sc_divergence(const double *s, double *out, size_t n, double dQ, double D1, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  sc_divergence_kernel<<<blocks_for(n, 256), 256, 0, st>>>(s, out, n, dQ, D1);
  BGPU_LAUNCHED(1);
}
void launch_alpt_combine(double2 *d2, const double2 *d4, int N, int Ns, int y0, double kfac, double rS,
                         cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const size_t n = (size_t)N * Ns * (N / 2 + 1);
  alpt_combine_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d2, d4, N, Ns, y0, kfac, rS);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// Exact adjoint of Lag2Eul_non_zeldovich (new: the reference has none, HMC_models.cc:458; derivation in
// oracle/barcode_oracle.py non_zeldovich_adjoint, validated there by finite differences of psi()).
// ---------------------------------------------------------------------------
// transpose of cellboundcomp: W(x) = 1/2 (V(x) + V(x + (1, 1, 1))), periodic
__global__ void cellbound_transpose_kernel(const double *__restrict__ v, double *__restrict__ w, int N) {
  const size_t n = (size_t)N * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)), i = (int)(idx >> (2 * sh));
  const size_t m = ((size_t)wrap_idx(i + 1, N) * N + wrap_idx(j + 1, N)) * N + wrap_idx(k + 1, N);
  w[idx] = 0.5 * (v[m] + v[idx]);
}

// P_ab = (d m2v / d L_ab) u: (Lyy + Lzz) u, (Lxx + Lzz) u, (Lxx + Lyy) u, -2 Lxy u, -2 Lxz u, -2 Lyz u
struct Six {
  double *p[6];
};
__global__ void lpt2_adjoint_coef_kernel(const double *__restrict__ phi, const double *__restrict__ u, Six out, int N,
                                         double fac) {
  const size_t n = (size_t)N * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)), i = (int)(idx >> (2 * sh));
  const double xx = findif2_at<0, 0>(phi, N, 0, i, j, k, fac), yy = findif2_at<1, 1>(phi, N, 0, i, j, k, fac),
               zz = findif2_at<2, 2>(phi, N, 0, i, j, k, fac);
  const double xy = findif2_at<0, 1>(phi, N, 0, i, j, k, fac), xz = findif2_at<0, 2>(phi, N, 0, i, j, k, fac),
               yz = findif2_at<1, 2>(phi, N, 0, i, j, k, fac);
  const double uu = u[idx];
  out.p[0][idx] = (yy + zz) * uu;
  out.p[1][idx] = (xx + zz) * uu;
  out.p[2][idx] = (xx + yy) * uu;
  out.p[3][idx] = -2.0 * xy * uu;
  out.p[4][idx] = -2.0 * xz * uu;
  out.p[5][idx] = -2.0 * yz * uu;
}

// G = sum_ab FD_a FD_b P_ab (the stencil is antisymmetric: the transpose of FD_b FD_a is FD_a FD_b)
__global__ void lpt2_adjoint_div_kernel(Six in, double *__restrict__ out, int N, double fac) {
  const size_t n = (size_t)N * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)), i = (int)(idx >> (2 * sh));
  out[idx] = findif2_at<0, 0>(in.p[0], N, 0, i, j, k, fac) + findif2_at<1, 1>(in.p[1], N, 0, i, j, k, fac) +
             findif2_at<2, 2>(in.p[2], N, 0, i, j, k, fac) + findif2_at<0, 1>(in.p[3], N, 0, i, j, k, fac) +
             findif2_at<0, 2>(in.p[4], N, 0, i, j, k, fac) + findif2_at<1, 2>(in.p[5], N, 0, i, j, k, fac);
}

// out = dQ (D1 u_lpt - D2 q + theta_SC'(dQ s) u_sc), theta_SC' = D1 / sqrt(1 - 2/3 D1 dQ s) where positive, else 0
__global__ void alpt_adjoint_combine_kernel(double *__restrict__ out, const double *__restrict__ u_lpt,
                                            const double *__restrict__ q, const double *__restrict__ u_sc,
                                            const double *__restrict__ s, size_t n, double dQ, double D1, double D2) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const double arg = 1. - 2. / 3. * D1 * (dQ * s[idx]);
  const double dsc = arg > 0. ? D1 / sqrt(arg) : 0.;
  out[idx] = dQ * (D1 * u_lpt[idx] - D2 * q[idx] + dsc * u_sc[idx]);
}

void launch_cellbound_transpose(const double *v, double *w, int N, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const size_t n = (size_t)N * N * N;
  cellbound_transpose_kernel<<<blocks_for(n, 256), 256, 0, st>>>(v, w, N);
  BGPU_LAUNCHED(1);
}
void launch_lpt2_adjoint_coef(const double *phi, const double *u, double *const out[6], int N, double L, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  Six o;
  for (int a = 0; a < 6; ++a) o.p[a] = out[a];
  const size_t n = (size_t)N * N * N;
  lpt2_adjoint_coef_kernel<<<blocks_for(n, 256), 256, 0, st>>>(phi, u, o, N, (double)N / (2. * L));
  BGPU_LAUNCHED(1);
}
void launch_lpt2_adjoint_div(double *const in[6], double *out, int N, double L, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  Six o;
  for (int a = 0; a < 6; ++a) o.p[a] = in[a];
  const size_t n = (size_t)N * N * N;
  lpt2_adjoint_div_kernel<<<blocks_for(n, 256), 256, 0, st>>>(o, out, N, (double)N / (2. * L));
  BGPU_LAUNCHED(1);
}
void launch_alpt_adjoint_combine(double *out, const double *u_lpt, const double *q, const double *u_sc, const double *s,
                                 size_t n, double dQ, double D1, double D2, cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  alpt_adjoint_combine_kernel<<<blocks_for(n, 256), 256, 0, st>>>(out, u_lpt, q, u_sc, s, n, dQ, D1, D2);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// SPH spline mass assignment and its exact adjoint (the reference's shipped default,
// data/input.par:11-13,134): getDensity_SPH (massFunctions.cc:392-495) and
// likelihood_calc_V_SPH (HMC_models.cc:200-303, SPH_kernel.cpp:62-208)
// ---------------------------------------------------------------------------
// Monaghan W_4 spline, SPH_kernel_3D (massFunctions.cc:366-384)
__device__ __forceinline__ double sph_kernel(double r, double h) {
  const double q = r / h;
  const double a = 1. / M_PI / (h * h * h);
  if (q <= 1.) return a * (1 - 3. / 2 * q * q + 3. / 4 * q * q * q);
  if (q <= 2.) {
    const double t = 2. - q;
    return a * (1. / 4 * (t * t * t));
  }
  return 0.;
}

__device__ void deposit_sph(const GridGeom &g, double x, double y, double z, double *__restrict__ rho);

// one thread per particle; every cell within `reach` of the particle's own cell whose centre is
// within 2h receives W(r, h) (unit mass, no normalisation: overdens divides by the mean afterwards)
__global__ void scatter_sph_kernel(GridGeom g, const double *__restrict__ psix, const double *__restrict__ psiy,
                                   const double *__restrict__ psiz, double *__restrict__ rho, double *__restrict__ posx,
                                   double *__restrict__ posy, double *__restrict__ posz) {
  const int N = g.N;
  const size_t n = (size_t)g.Ns * N * N;  // the Lagrangian planes this rank owns (a cube: all of them)
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);  // N is a power of two
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)),
            i = g.x0 + (int)(idx >> (2 * sh));
  double x, y, z;
  particle_position(g, i, j, k, psix[idx], psiy[idx], psiz[idx], x, y, z);
  if (posx) {
    posx[idx] = x;
    posy[idx] = y;
    posz[idx] = z;
  }
  deposit_sph(g, x, y, z, rho);
}

__device__ void deposit_sph(const GridGeom &g, double x, double y, double z, double *__restrict__ rho) {
  const int N = g.N;
  if (!in_domain(g, x, y, z)) return;
  const double d = g.d, h = g.sph_h;
  const int reach = (int)(2 * h / d) + 1;
  const int ix = (int)(unsigned long long)(x / d), iy = (int)(unsigned long long)(y / d),
            iz = (int)(unsigned long long)(z / d);
  const double ccx = ((double)ix + 0.5) * d, ccy = ((double)iy + 0.5) * d, ccz = ((double)iz + 0.5) * d;
  // The reference tests all (2 reach + 1)^3 cells with a square root and two divisions each
  // (massFunctions.cc:423-476, :366-384).  Here:
  //  * a cell whose distance along x alone, or in the (x, y) plane alone, already exceeds 2h cannot pass r/h <= 2:
  //    those planes and rows are skipped on the squared distance (margin far wider than rounding) -- the set of
  //    cells that receive mass is unchanged;
  //  * q = r/h is formed as r^2 * rsqrt(r^2) / h with the reciprocal of h hoisted: one special-function call per
  //    cell instead of a square root and two divisions.  q differs from the reference's by an ulp or two, W by as
  //    much; a cell within that of the edge q = 2 carries (2 - q)^3 / 4 ~ 1e-45 either way.
  const double lim = 4. * h * h * (1. + 1.e-9);
  const double h_inv = 1. / h;
  const double a = 1. / M_PI / (h * h * h);
  // plane of the (halo-extended, slab-local) density tile that global cell plane c maps to (a cube: c itself)
  const int lo = g.x0 - g.H, np = g.Ns + 2 * g.H;
  for (int i1 = -reach; i1 <= reach; ++i1) {
    const double dx = x - (ccx + (double)i1 * d);
    if (dx * dx > lim) continue;
    int kx = (N + i1 + ix) % N - lo;
    if (kx < 0) kx += N;
    if (kx >= N) kx -= N;
    if (kx >= np) {  // beyond the halo: drop the deposit and raise the flag (the host turns it into an error)
      if (g.flag) *g.flag = 1;
      continue;
    }
    for (int i2 = -reach; i2 <= reach; ++i2) {
      const double dy = y - (ccy + (double)i2 * d);
      const double rxy = dx * dx + dy * dy;
      if (rxy > lim) continue;
      const int ky = (N + i2 + iy) % N;
      double *row = rho + ((size_t)kx * N + ky) * N;
      for (int i3 = -reach; i3 <= reach; ++i3) {
        const double dz = z - (ccz + (double)i3 * d);
        const double r2 = rxy + dz * dz;
        if (r2 > lim) continue;
        const double q = r2 > 0. ? r2 * rsqrt(r2) * h_inv : 0.;
        double w;
        if (q <= 1.) {
          w = a * (1 - 3. / 2 * q * q + 3. / 4 * q * q * q);
        } else if (q <= 2.) {
          const double t = 2. - q;
          w = a * (1. / 4 * (t * t * t));
        } else {
          continue;
        }
        red_add(row + (N + i3 + iz) % N, w);
      }
    }
  }
}

__global__ void scatter_sph_positions_kernel(GridGeom g, const double *__restrict__ x, const double *__restrict__ y,
                                             const double *__restrict__ z, double *__restrict__ rho) {
  const size_t n = (size_t)g.N * g.N * g.N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) deposit_sph(g, x[idx], y[idx], z[idx], rho);
}

void launch_scatter_sph(const GridGeom &g, const double *psix, const double *psiy, const double *psiz, double *rho,
                        double *posx, double *posy, double *posz, cudaStream_t st) {
  if (g.sph_cols) {  // static column lists (particles_sph.cu); this file's kernels are the general path
    launch_scatter_sph_cols(g, g.sph_cols, psix, psiy, psiz, rho, posx, posy, posz, st);
    return;
  }
  ProfScope prof(KK_SCATTER, st);
  const size_t n = (size_t)g.Ns * g.N * g.N;
  BGPU_CUDA(cudaMemsetAsync(rho, 0, (size_t)(g.Ns + 2 * g.H) * g.N * g.N * sizeof(double), st));
  scatter_sph_kernel<<<blocks_for(n, 128), 128, 0, st>>>(g, psix, psiy, psiz, rho, posx, posy, posz);
  BGPU_LAUNCHED(1);
}

// V_p = (rho_c V/N) sum_{hull cells} r_c gradW((x_p - x_c)/h); the hull is every (i, j) column with the
// inclusive k range that can reach the central cell (kmax[(i+R)(2R+1) + (j+R)], -1 = column not in the hull);
// gradW = partial(q) * (x_p - x_c)/h / (pi h^4).  In place over Psi.  Deterministic (pure gather).
__global__ void gather_sph_kernel(GridGeom g, double *__restrict__ ax, double *__restrict__ ay, double *__restrict__ az,
                                  const double *__restrict__ resid, const int *__restrict__ kmax, int R,
                                  double normalize) {
  const int N = g.N;
  const size_t n = (size_t)g.Ns * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);  // N is a power of two
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)),
            i = g.x0 + (int)(idx >> (2 * sh));
  // the residual is [(Ns + 2H)][N][N] with plane 0 = global x0 - H (a cube: the whole grid)
  const int lo = g.x0 - g.H, np = g.Ns + 2 * g.H;
  double px, py, pz;
  particle_position(g, i, j, k, ax[idx], ay[idx], az[idx], px, py, pz);
  const double d = g.d, h = g.sph_h, h_inv = 1. / h, d_h = d * h_inv;
  const double norm = 1. / (M_PI * (h * h) * (h * h));
  const int ix = (int)(px / d), iy = (int)(py / d), iz = (int)(pz / d);
  const double dpcx = px * h_inv - ((double)ix + 0.5) * d_h, dpcy = py * h_inv - ((double)iy + 0.5) * d_h,
               dpcz = pz * h_inv - ((double)iz + 0.5) * d_h;
  double vx = 0., vy = 0., vz = 0.;
  for (int i1 = -R; i1 <= R; ++i1) {
    const double dxh = dpcx - (double)i1 * d_h;
    int kx = (ix + i1 + N) % N - lo;
    if (kx < 0) kx += N;
    if (kx >= N) kx -= N;
    if (kx >= np) continue;  // beyond the halo: the scatter of the same evaluation has raised the flag already
    for (int i2 = -R; i2 <= R; ++i2) {
      const int K = kmax[(i1 + R) * (2 * R + 1) + (i2 + R)];
      if (K < 0) continue;
      const double dyh = dpcy - (double)i2 * d_h;
      const double qxy = dxh * dxh + dyh * dyh;
      const double *row = resid + ((size_t)kx * N + (iy + i2 + N) % N) * N;
      for (int i3 = -K; i3 <= K; ++i3) {
        const double dzh = dpcz - (double)i3 * d_h;
        const double q_sq = qxy + dzh * dzh;
        if (q_sq > 4.) continue;
        // q and 1/q from one reciprocal square root (the reference takes a square root and divides,
        // SPH_kernel.cpp:148-208; an ulp of difference in q)
        const double q_inv = q_sq > 0. ? rsqrt(q_sq) : 0.;
        const double q = q_sq * q_inv;
        double partial;
        if (q_sq > 1.) {
          const double qm = q - 2.;
          partial = -0.75 * qm * qm * norm * q_inv;
        } else {
          partial = (2.25 * q - 3.) * norm;
        }
        const double c = __ldg(row + (iz + i3 + N) % N) * partial;
        vx += c * dxh;
        vy += c * dyh;
        vz += c * dzh;
      }
    }
  }
  vx *= normalize;
  vy *= normalize;
  vz *= normalize;
  if (g.rsd) vz += g.fgrow * vz;  // HMC_models.cc:295-301
  ax[idx] = vx;
  ay[idx] = vy;
  az[idx] = vz;
}

void launch_gather_sph(const GridGeom &g, double *ax, double *ay, double *az, const double *resid, const int *kmax,
                       int R, double normalize, cudaStream_t st) {
  if (g.sph_cols) {
    launch_gather_sph_cols(g, g.sph_cols, ax, ay, az, resid, normalize, st);
    return;
  }
  ProfScope prof(KK_GATHER, st);
  const size_t n = (size_t)g.Ns * g.N * g.N;
  gather_sph_kernel<<<blocks_for(n, 128), 128, 0, st>>>(g, ax, ay, az, resid, kmax, R, normalize);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// calc_h = 3: likelihood_calc_V_SPH_fourier_TSC (HMC_models_testing.cpp:54-188) -- the SPH adjoint with the gather
// replaced by a k-space convolution and a TSC interpolation to the particles.
// ---------------------------------------------------------------------------
// out = i k_c * hW(k) * r^(k), hW = h * SPH_kernel_F on the padded half grid (built on the host with the C library the
// reference uses, api.cu); no Nyquist zeroing (:111-128).  Cube layout [x][y][z <= N/2].
__global__ void sph_fourier_comp_kernel(const double2 *__restrict__ rhat, const double *__restrict__ hW,
                                        double2 *__restrict__ out, int N, double kfac, int comp) {
  const int nzh = N / 2 + 1;
  const size_t n = (size_t)N * N * nzh;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const size_t row = idx / nzh;
  const int z = (int)(idx - row * nzh);
  const int y = (int)(row % N), x = (int)(row / N);
  const int m = comp == 0 ? x : (comp == 1 ? y : z);
  const double kc = (m <= N / 2) ? kfac * (double)m : -kfac * (double)(N - m);  // scale_space.cpp:41-51
  const double f = kc * hW[row * (nzh + 1) + z];
  const double2 v = rhat[idx];
  out[idx] = make_double2(f * -v.y, f * v.x);
}

void launch_sph_fourier_comp(const double2 *rhat, const double *hW_half, double2 *out, int N, double kfac, int comp,
                             cudaStream_t st) {
  ProfScope prof(KK_STREAM, st);
  const size_t n = (size_t)N * N * (N / 2 + 1);
  sph_fourier_comp_kernel<<<blocks_for(n, 256), 256, 0, st>>>(rhat, hW_half, out, N, kfac, comp);
  BGPU_LAUNCHED(1);
}

// interpolate_TSC (interpolate_grid.cpp:134-202) at the particle positions (recomputed from Psi), bug-compatible:
// the upper weights of x and y are formed from dz (:166-167).  Safe in place over one of the Psi arrays: a thread
// reads only its own element of them.
__global__ void interp_tsc_kernel(GridGeom g, const double *psix, const double *psiy, const double *psiz,
                                  const double *__restrict__ field, double *out, double fz) {
  const int N = g.N;
  const size_t n = (size_t)N * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)), i = (int)(idx >> (2 * sh));
  double px, py, pz;
  particle_position(g, i, j, k, psix[idx], psiy[idx], psiz[idx], px, py, pz);
  const double xk = __ddiv_rn(px, g.d), yk = __ddiv_rn(py, g.d), zk = __ddiv_rn(pz, g.d);
  unsigned ix = (unsigned)xk, iy = (unsigned)yk, iz = (unsigned)zk;
  const double dx = __dsub_rn(xk, (double)ix + 0.5), dy = __dsub_rn(yk, (double)iy + 0.5),
               dz = __dsub_rn(zk, (double)iz + 0.5);
  // (a position that rounds to exactly L would index one past the grid in the reference: fold it)
  if (ix >= (unsigned)N) ix -= (unsigned)N;
  if (iy >= (unsigned)N) iy -= (unsigned)N;
  if (iz >= (unsigned)N) iz -= (unsigned)N;
  auto sq = [](double a) { return __dmul_rn(a, a); };
  const double up = __dmul_rn(0.5, sq(__dsub_rn(1.5, fabs(__dsub_rn(dz, 1.0)))));  // shared by x, y and z: the slip
  const double wx[3] = {__dmul_rn(0.5, sq(__dsub_rn(1.5, fabs(__dadd_rn(dx, 1.0))))), __dsub_rn(0.75, sq(dx)), up};
  const double wy[3] = {__dmul_rn(0.5, sq(__dsub_rn(1.5, fabs(__dadd_rn(dy, 1.0))))), __dsub_rn(0.75, sq(dy)), up};
  const double wz[3] = {__dmul_rn(0.5, sq(__dsub_rn(1.5, fabs(__dadd_rn(dz, 1.0))))), __dsub_rn(0.75, sq(dz)), up};
  const unsigned cx[3] = {(ix + N - 1) % N, ix, (ix + 1) % N}, cy[3] = {(iy + N - 1) % N, iy, (iy + 1) % N},
                 cz[3] = {(iz + N - 1) % N, iz, (iz + 1) % N};
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double *row = field + ((size_t)cx[a] * N + cy[b]) * N;
      const double wab = __dmul_rn(wx[a], wy[b]);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(wab, wz[c]), __ldg(row + cz[c])));
    }
  out[idx] = fz != 0.0 ? acc + fz * acc : acc;
}

void launch_interp_tsc(const GridGeom &g, const double *psix, const double *psiy, const double *psiz,
                       const double *field, double *out, double fz, cudaStream_t st) {
  ProfScope prof(KK_GATHER, st);
  const size_t n = (size_t)g.N * g.N * g.N;
  interp_tsc_kernel<<<blocks_for(n, 256), 256, 0, st>>>(g, psix, psiy, psiz, field, out, fz);
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// F4: setup_random_test's observations (barcoderunner.cc:94-195) with the generator on the device.  One
// thread per cell: window (window_type 1 / 10 / 23), then nobs and noise by data model and likelihood.
// Counter-based stream: cell g of (seed, draw 1) always gets the same numbers, whatever the launch shape;
// a Poisson rejection loop takes its k-th attempt from counter word 3 = k.  NOT GSL's stream (which is serial).
// ---------------------------------------------------------------------------
struct MockRng {
  uint64_t g, seed;
  uint32_t attempt;
  __device__ __forceinline__ void next(double &u1, double &u2) {
    uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), 1u, (3u << 24) ^ attempt};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    ++attempt;
    u1 = u53(c[0], c[1]);
    u2 = u53(c[2], c[3]);
  }
  __device__ __forceinline__ double normal() {  // Box-Muller, as philox_normal_kernel
    double u1, u2;
    next(u1, u2);
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    return sqrt(-2.0 * log(u1)) * cs;
  }
  // Poisson(lam): CDF inversion below 30, Hoermann's transformed rejection (PTRS, 1993) above
  __device__ double poisson(double lam) {
    if (!(lam > 0.0)) return 0.0;  // gsl_ran_poisson returns 0 for mu <= 0
    double u1, u2;
    if (lam < 30.0) {
      next(u1, u2);
      double p = exp(-lam), F = p;
      int k = 0;
      while (u1 > F && k < 2000) {
        ++k;
        p *= lam / (double)k;
        F += p;
      }
      return (double)k;
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < 1000; ++it) {
      next(u1, u2);
      const double U = u1 - 0.5, V = u2, us = 0.5 - fabs(U);
      const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
      if (us >= 0.07 && V <= vr) return k;
      if (k < 0.0 || (us < 0.013 && V > us)) continue;
      if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lam);
  }
};

__global__ void mock_obs_kernel(MockObs mp, const double *__restrict__ delta_eul, const double *__restrict__ delta_lag,
                                double *__restrict__ window, double *__restrict__ nobs, double *__restrict__ noise,
                                size_t n, size_t first, size_t n_global, uint64_t seed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t g = first + i;
  const double de = delta_eul[i];
  double w = 1.0;                                             // window_type 1
  if (mp.window_type == 10) w = g < n_global / 2 ? 0.0 : 1.0;  // :105-108
  else if (mp.window_type == 23) w = de > 3. ? 1.0 : 0.0;      // :109-117, as written there
  window[i] = w;
  MockRng rng{(uint64_t)g, seed, 0u};
  if (mp.data_model == 0) {  // linear data model, :127-166
    const double Lambda = mp.rho_c * (1.0 + de);
    if (w > 0.) {
      if (mp.likelihood == 0) {
        nobs[i] = rng.poisson(Lambda);
      } else if (mp.likelihood == 1) {
        const double sigma = mp.sigma_min + mp.sigma_fac * Lambda;
        noise[i] = sigma;
        double v = Lambda + sigma * rng.normal();
        if (!mp.negative_obs && v < 0) v = 0;
        nobs[i] = v;
      } else {  // 3: Gaussian random field, sigma quadratic in the Lagrangian field
        const double dl = delta_lag[i];
        const double sigma = mp.sigma_min + mp.sigma_fac * (dl * dl);
        noise[i] = sigma;
        nobs[i] = dl + sigma * rng.normal();
      }
    } else {
      nobs[i] = 0.0;
    }
  } else {  // log-normal data model, :167-188
    const double d = de < mp.delta_min ? mp.delta_min : de;
    const double Lambda = log(mp.rho_c * (1.0 + d));
    if (w > 0.) {
      noise[i] = mp.sigma_fac;
      nobs[i] = Lambda + mp.sigma_fac * rng.normal();
    } else {
      const double q = mp.rho_c * (1 + mp.delta_min);
      nobs[i] = log(q * q);  // "huh, why squared?" (:186)
    }
  }
}

void launch_mock_obs(const MockObs &mp, const double *delta_eul, const double *delta_lag, double *window, double *nobs,
                     double *noise, size_t n, size_t first, size_t n_global, uint64_t seed, cudaStream_t st) {
  ProfScope prof(KK_COLOUR, st);
  mock_obs_kernel<<<blocks_for(n, 256), 256, 0, st>>>(mp, delta_eul, delta_lag, window, nobs, noise, n, first, n_global, seed);
  BGPU_LAUNCHED(1);
}

}  // namespace bgpu
