// barcode_b200/csrc/f32_path.cu -- the single-precision mode of the HMC gradient-and-leapfrog path
// (bgpu_f32_* in include/barcode_gpu.h).
//
// The reference selects its arithmetic at build time: SINGLE_PREC makes real_prec = float and routes the transforms
// through fftwf (define_opt.h:50-59, cmake/Modules/Options.cmake:65-66, fftwrapper.cc:32-36).  A SINGLE_PREC build of
// the reference binds these entry points where a DOUBLE_PREC build binds bgpu_*; one library serves both.
//
// Scope: the path BASELINE.json's north star names -- Zel'dovich displacements (sfmodel 1, or any sfmodel under
// rsd_model, HMC_models.cc:395-406), CIC mass assignment, plane-parallel RSD, Poisson or Gaussian likelihood,
// calc_h = 0 (the reference's gradient, HMC_models_testing.cpp:25-50), 1, or 4 (exact CIC adjoint), Gaussian prior,
// psi, kinetic_term, Hamiltonian_EoM; one GPU, N = 32 ... 512.  Everything else stays FP64-only and is refused here.
//
// Arithmetic: arrays, transforms and per-cell / per-particle work in float; reductions (sum rho, -lnL, prior,
// kinetic energy) accumulate in double with the fixed two-stage tree of the FP64 path.  A particle's cell and CIC
// weights come from its displacement in cell units (cell = i + floor(Psi/d), weight = frac(Psi/d)) instead of from
// the rounded absolute position d (i + 1/2) + Psi: mathematically the same cell and weights as getCICcells /
// getCICweights (interpolate_grid.cpp:27-79), without losing log2(N) bits of the weight to the position's magnitude.
#include "barcode_gpu.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "f32_fft.cuh"
#include "host_math.h"
#include "util.h"

namespace bgpu {
void set_last_error(const std::string &msg);  // api.cu: what bgpu_last_error() returns on this thread

namespace f32 {

constexpr int kThreads = 256;
constexpr int kBlocks = 1184;  // 148 SMs x 8 resident CTAs: the fixed grid of the two-stage reductions

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double s[kThreads / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < kThreads / 32 ? s[threadIdx.x] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

template <class F>
__global__ void __launch_bounds__(kThreads) partial_sum_kernel(F f, size_t n, double *__restrict__ part) {
  double v = 0.0;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) v += f(i);
  const double r = block_sum(v);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}
__global__ void __launch_bounds__(kThreads) final_sum_kernel(const double *__restrict__ part, int n, double *out) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) v += part[i];
  const double r = block_sum(v);
  if (threadIdx.x == 0) *out = r;
}
static int grid_for(size_t n) {
  const size_t b = (n + kThreads - 1) / kThreads;
  return (int)(b < (size_t)kBlocks ? b : (size_t)kBlocks);
}
template <class F>
static void reduce(F f, size_t n, double *scratch, double *out, cudaStream_t st) {
  ProfScope prof(KK_REDUCE, st);
  const int blocks = grid_for(n);
  partial_sum_kernel<F><<<blocks, kThreads, 0, st>>>(f, n, scratch);
  final_sum_kernel<<<1, kThreads, 0, st>>>(scratch, blocks, out);
  BGPU_LAUNCHED(2);
}

struct SumF {  // four cells per index (16-byte loads)
  const float4 *a;
  __device__ double operator()(size_t i) const {
    const float4 v = a[i];
    return ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
  }
};
// 1/2 sum_x a (C^-1 a) by Parseval on the half grid (kernels.cu HalfQuadF; gaussian.cpp:20-35, HMC.cc:82-110)
struct HalfQuadF {
  const float2 *v;
  const float *mult;
  int nzh;
  double inv_2n;
  __device__ double operator()(size_t i) const {
    const int z = (int)(i % (size_t)nzh);
    const float2 a = v[i];
    const double w = (z == 0 || z == nzh - 1) ? 1.0 : 2.0;
    return inv_2n * w * (double)mult[i] * ((double)a.x * a.x + (double)a.y * a.y);
  }
};
struct KineticRealF {  // mass_type 0: 1/2 p M_r^-1 p (HMC.cc:88-110)
  const float *p, *mass_r;
  __device__ double operator()(size_t i) const {
    const double m = mass_r[i];
    return m > 0.0 ? 0.5 * (double)p[i] * (double)p[i] / m : 0.0;
  }
};

// ---------------------------------------------------------------------------
// particles
// ---------------------------------------------------------------------------
struct GeomF {
  int N, sh;
  float inv_d;   // 1 / cell size
  float zfac;    // 1 + c_pecvel * v_norm under RSD (Lag2Eul.cc:378-381, rsd.cc:52-63), else 1
  float fgrow;   // f under RSD, else 0: d z_s / d Psi_z = 1 + f
};

// cell and CIC weights along one axis of the particle on lattice site i displaced by u cells
__device__ __forceinline__ void cic_axis_f(int i, float u, int N, unsigned &c0, unsigned &c1, float &w0, float &w1) {
  const float fl = floorf(u);
  w1 = u - fl;
  w0 = 1.f - w1;
  c0 = (unsigned)(i + (int)fl) & (unsigned)(N - 1);
  c1 = (c0 + 1u) & (unsigned)(N - 1);
}

// getDensity_CIC (massFunctions.cc:100-164) of the displaced lattice (disp_part.cc:55-126, calc_pos_rsd rsd.cc:30-64).
// One thread per particle, z fastest: lanes are z neighbours, so a lane's upper z cell usually is the next lane's
// lower one; it is handed over by shuffle and four of the eight reductions disappear.
__global__ void __launch_bounds__(kThreads) scatter_cic_kernel(GeomF g, const float *__restrict__ psix,
                                                               const float *__restrict__ psiy,
                                                               const float *__restrict__ psiz, float *__restrict__ rho) {
  const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;  // the grid covers N^3 exactly (N^3 % 256 == 0)
  const int N = g.N, sh = g.sh;
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)), i = (int)(idx >> (2 * sh));
  const int lane = threadIdx.x & 31;
  unsigned ci[2], cj[2], ck0, ck1;
  float wi[2], wj[2], wk0, wk1;
  cic_axis_f(i, psix[idx] * g.inv_d, N, ci[0], ci[1], wi[0], wi[1]);
  cic_axis_f(j, psiy[idx] * g.inv_d, N, cj[0], cj[1], wj[0], wj[1]);
  cic_axis_f(k, psiz[idx] * g.inv_d * g.zfac, N, ck0, ck1, wk0, wk1);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const unsigned row = ((ci[a] << sh) + cj[b]) << sh;
      const unsigned lo = row + ck0, hi = row + ck1;
      const float wab = wi[a] * wj[b];
      float vlo = wab * wk0;
      const float vhi = wab * wk1;
      const unsigned nlo = __shfl_down_sync(0xffffffffu, lo, 1);
      const bool give = lane < 31 && nlo == hi;             // the next lane adds my upper cell to its lower one
      const float got = __shfl_up_sync(0xffffffffu, vhi, 1);
      const unsigned givers = __ballot_sync(0xffffffffu, give);
      if (lane > 0 && ((givers >> (lane - 1)) & 1u)) vlo += got;
      atomicAdd(rho + lo, vlo);
      if (!give) atomicAdd(rho + hi, vhi);
    }
}

// exact adjoint of the CIC deposit (kernels.cu gather_adjoint_kernel): V_c(p) = sum_cells r(cell) dW/dx_c, written in
// place over Psi; z carries d z_s / d Psi_z = 1 + f under RSD (cf. HMC_models.cc:295-301)
__global__ void __launch_bounds__(kThreads) gather_cic_kernel(GeomF g, float *ax, float *ay, float *az,
                                                              const float *__restrict__ resid) {
  const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const int N = g.N, sh = g.sh;
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)), i = (int)(idx >> (2 * sh));
  unsigned ci[2], cj[2], ck[2];
  float wi[2], wj[2], wk[2];
  cic_axis_f(i, ax[idx] * g.inv_d, N, ci[0], ci[1], wi[0], wi[1]);
  cic_axis_f(j, ay[idx] * g.inv_d, N, cj[0], cj[1], wj[0], wj[1]);
  cic_axis_f(k, az[idx] * g.inv_d * g.zfac, N, ck[0], ck[1], wk[0], wk[1]);
  const float gs[2] = {-g.inv_d, g.inv_d};
  float vx = 0.f, vy = 0.f, vz = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const unsigned row = ((ci[a] << sh) + cj[b]) << sh;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float rc = __ldg(resid + row + ck[c]);
        vx += rc * gs[a] * wj[b] * wk[c];
        vy += rc * wi[a] * gs[b] * wk[c];
        vz += rc * wi[a] * wj[b] * gs[c];
      }
    }
  vz += g.fgrow * vz;
  ax[idx] = vx;
  ay[idx] = vy;
  az[idx] = vz;
}

// ---------------------------------------------------------------------------
// overdensity + likelihood residual + -lnL (kernels.cu ResidualEval; overdens massFunctions.cc:30-47,
// gaussian_independent.cpp:24-42,80-91, poissonian.cpp:19-34,60-73)
// ---------------------------------------------------------------------------
struct LikeF {
  int likelihood;  // 0 Poisson, 1 Gaussian
  float rho_c, biasP, biasE;
  int exact_sign;
};

__device__ __forceinline__ float residual_one(const LikeF &lp, bool unit, float rho, float inv_mean, float nn, float sg,
                                              float w, float &delta, float &r) {
  delta = rho * inv_mean - 1.f;
  const float dens = 1.f + lp.biasP * delta;
  const float Lambda = w * lp.rho_c * (unit ? dens : powf(dens, lp.biasE));
  float val = 0.f;
  r = 0.f;
  if (lp.likelihood == 1) {
    if (w > 0.f && Lambda > 0.f) {
      r = (nn - Lambda) / (sg * sg);
      const float q = (Lambda - nn) / sg;
      val = 0.5f * q * q;
    }
  } else {
    if (w > 0.f && dens > 0.f) {
      r = (1.f - nn / Lambda) * lp.rho_c * lp.biasE * lp.biasP * (unit ? 1.f : powf(dens, lp.biasE - 1.f));
      if (lp.exact_sign) r = -r;
    }
    if (w > 0.f && Lambda > 0.f) val = Lambda - nn * logf(Lambda);
  }
  return val;
}

// four cells per thread per trip (16-byte loads); n4 = n / 4 (N^3 is a multiple of 4)
__global__ void __launch_bounds__(kThreads)
    residual_kernel(LikeF lp, float4 *__restrict__ rho_delta, const double *__restrict__ sum_rho, double count,
                    const float4 *__restrict__ nobs, const float4 *__restrict__ noise, const float4 *__restrict__ window,
                    float4 *__restrict__ resid, size_t n4, double *__restrict__ part) {
  const float inv_mean = (float)(count / *sum_rho);
  const bool unit = lp.biasE == 1.f;
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 ro = rho_delta[i], nn = nobs[i], sg = noise[i], w = window[i];
    float4 d, r;
    float v = residual_one(lp, unit, ro.x, inv_mean, nn.x, sg.x, w.x, d.x, r.x);
    v += residual_one(lp, unit, ro.y, inv_mean, nn.y, sg.y, w.y, d.y, r.y);
    v += residual_one(lp, unit, ro.z, inv_mean, nn.z, sg.z, w.z, d.z, r.z);
    v += residual_one(lp, unit, ro.w, inv_mean, nn.w, sg.w, w.w, d.w, r.w);
    rho_delta[i] = d;
    if (resid) resid[i] = r;
    acc += (double)v;
  }
  const double r = block_sum(acc);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// r * d_c(delta), 4th-order central difference (gradfindif, gradient.cpp:81-153; Poisson calc_h = 0)
__global__ void __launch_bounds__(kThreads) findif_product_kernel(const float *__restrict__ in,
                                                                  const float *__restrict__ resid,
                                                                  float *__restrict__ out, int N, int sh, float fac,
                                                                  int comp) {
  const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const int c[3] = {(int)(idx >> (2 * sh)), (int)((idx >> sh) & (size_t)(N - 1)), (int)(idx & (size_t)(N - 1))};
  const ptrdiff_t stride = comp == 0 ? (ptrdiff_t)N * N : (comp == 1 ? (ptrdiff_t)N : 1);
  const int ii = c[comp];
  const ptrdiff_t base = (ptrdiff_t)idx - (ptrdiff_t)ii * stride;
  const int r = (ii + 1) & (N - 1), rr = (ii + 2) & (N - 1), l = (ii - 1) & (N - 1), ll = (ii - 2) & (N - 1);
  const float g = -(fac * ((4.f / 3) * (in[base + l * stride] - in[base + r * stride]) -
                           (1.f / 6) * (in[base + ll * stride] - in[base + rr * stride])));
  out[idx] = resid[idx] * g;
}

// half-grid multiplier normFS / C(k) from a full real-indexed spectrum (HMC_help.cc:41-58; 0 where C <= 0)
__global__ void inverse_spectrum_kernel(const float *__restrict__ full, float *__restrict__ half, int N, float normFS) {
  const int nzh = N / 2 + 1;
  const size_t n = (size_t)N * N * nzh;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int k = (int)(idx % nzh);
  const size_t ij = idx / nzh;
  const float c = full[ij * N + k];
  half[idx] = c > 0.f ? normFS / c : 0.f;
}

// Hamiltonian mass types 0 / 1 / 4 (HMC_mass.cc:117-124,163-172,315-368)
__global__ void mass_kernel(const float *__restrict__ power, float *__restrict__ mass_f, float *__restrict__ mass_r,
                            int type, float factor, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (type == 0) mass_r[i] = 1.f;
  else if (type == 1) {
    const float P = power[i];
    mass_f[i] = factor * (P > 0.f ? 1.f / P : 0.f);
  } else mass_f[i] = factor * power[i];
}

__global__ void axpy_kernel(float *__restrict__ y, const float *__restrict__ x, float a, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}
// s += eps p / M_r (HMC.cc:317-327)
__global__ void axpy_div_kernel(float *__restrict__ y, const float *__restrict__ x, const float *__restrict__ m, float a,
                                size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float mm = m[i];
    y[i] += a * (mm > 0.f ? x[i] / mm : 0.f);
  }
}

// leapfrog in k-space (kernels.cu kspace_drift_kernel / KspaceKickF): s^ += eps (V/N)/M p^ ;
// p^ += a ((V/N)/P s^ + norm h^) -- both diagonal on the half grid
__global__ void kspace_drift_kernel(float2 *__restrict__ shat, const float2 *__restrict__ phat,
                                    const float *__restrict__ inv_mass, float eps, size_t nh) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const float f = eps * inv_mass[i];
  const float2 pk = phat[i];
  float2 sk = shat[i];
  sk.x = fmaf(f, pk.x, sk.x);
  sk.y = fmaf(f, pk.y, sk.y);
  shat[i] = sk;
}
__global__ void kspace_kick_kernel(float2 *__restrict__ phat, const float2 *__restrict__ shat,
                                   const float2 *__restrict__ hhat, const float *__restrict__ prior, float a, float norm,
                                   size_t nh) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const float f = prior[i];
  const float2 sk = shat[i], hk = hhat[i];
  float2 pk = phat[i];
  pk.x = fmaf(a, fmaf(sk.x, f, norm * hk.x), pk.x);
  pk.y = fmaf(a, fmaf(sk.y, f, norm * hk.y), pk.y);
  phat[i] = pk;
}

static unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ---------------------------------------------------------------------------
// the transform: plan (twiddles, launch attributes) and the pass sequences
// ---------------------------------------------------------------------------
struct Fft {
  int N = 0;
  cudaStream_t stream = nullptr;
  float2 *twN = nullptr, *twM = nullptr;
  int grid_strided = 0;

  template <int N_>
  struct Cfg {
    static constexpr int T = 16;                                  // pencils per tile: 128-byte rows
    static constexpr int TR = (N_ / 16 >= 256) ? 1 : 256 / (N_ / 16);  // rows per z-pass CTA: 256 threads
    static constexpr size_t smem_strided = (size_t)3 * N_ * T * sizeof(float2);
    static constexpr size_t smem_z = (size_t)TR * (N_ / 2 + N_ / 16 + 1) * sizeof(float2);
  };

  // tiles in flight ahead of the one being transformed: 1.  Two (three tile buffers, BGPU_F32_PD=2) measured slower on
  // B200 -- y pass 43.9 against 40.3 us at 256^3, 328 against 311 us at 512^3: the pass waits on its CTA barriers and
  // on shared memory, not on the copies
  int pd = 1;
  int grid_strided_pd1 = 0;

  template <int N_, int T_, int PD>
  int init_strided() {
    constexpr size_t smem = (size_t)(1 + PD) * N_ * T_ * sizeof(float2);
    BGPU_CUDA(cudaFuncSetAttribute(strided_pass<N_, T_, -1, 0, PD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BGPU_CUDA(cudaFuncSetAttribute(strided_pass<N_, T_, -1, 1, PD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BGPU_CUDA(cudaFuncSetAttribute(strided_pass<N_, T_, +1, 0, PD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BGPU_CUDA(cudaFuncSetAttribute(strided_pass<N_, T_, +1, 1, PD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, dev = 0, sms = 0;
    BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, strided_pass<N_, T_, +1, 0, PD>, T_ * N_ / 8, smem));
    BGPU_CUDA(cudaGetDevice(&dev));
    BGPU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (occ < 1) throw std::runtime_error("bgpu_f32: the strided pass does not fit an SM at this size");
    constexpr int tiles = N_ * ((N_ / 2) / T_) + N_ / T_;
    return occ * sms < tiles ? occ * sms : tiles;
  }

  template <int N_>
  void init_n() {
    using C = Cfg<N_>;
    grid_strided = init_strided<N_, C::T, 2>();
    grid_strided_pd1 = init_strided<N_, C::T, 1>();
    const char *e = std::getenv("BGPU_F32_PD");
    pd = (e && std::atoi(e) == 2) ? 2 : 1;
    BGPU_CUDA(cudaFuncSetAttribute(r2c_zpass<N_, C::TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_z));
    BGPU_CUDA(cudaFuncSetAttribute(c2r_zpass<N_, C::TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_z));
  }

  void init(int n, cudaStream_t st) {
    N = n;
    stream = st;
    std::vector<float2> t((size_t)n);
    for (int k = 0; k < n; ++k) {
      const long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)n;
      t[k] = make_float2((float)cosl(a), (float)sinl(a));
    }
    BGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&twN), sizeof(float2) * n));
    BGPU_CUDA(cudaMemcpy(twN, t.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    const int m = n / 2;
    for (int k = 0; k < m; ++k) {
      const long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)m;
      t[k] = make_float2((float)cosl(a), (float)sinl(a));
    }
    BGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&twM), sizeof(float2) * m));
    BGPU_CUDA(cudaMemcpy(twM, t.data(), sizeof(float2) * m, cudaMemcpyHostToDevice));
    switch (n) {
      case 32: init_n<32>(); break;
      case 64: init_n<64>(); break;
      case 128: init_n<128>(); break;
      case 256: init_n<256>(); break;
      case 512: init_n<512>(); break;
      default: throw std::runtime_error("bgpu_f32: N1 must be 32, 64, 128, 256 or 512");
    }
  }
  void destroy() {
    cudaFree(twN);
    cudaFree(twM);
    twN = twM = nullptr;
  }

  template <int N_, int DIR, int AXIS>
  void strided_n(const float2 *in, float2 *out, const KOpF &lop, const KOpF &sop) {
    using C = Cfg<N_>;
    ProfScope prof(AXIS == 0 ? KK_FFT_STRIDED_X : KK_FFT_STRIDED, stream);
    if (pd == 1)
      strided_pass<N_, C::T, DIR, AXIS, 1><<<grid_strided_pd1, C::T * N_ / 8, (size_t)2 * N_ * C::T * sizeof(float2), stream>>>(in, out, twN, lop, sop);
    else
      strided_pass<N_, C::T, DIR, AXIS, 2><<<grid_strided, C::T * N_ / 8, C::smem_strided, stream>>>(in, out, twN, lop, sop);
    BGPU_LAUNCHED(1);
  }
  template <int N_>
  void r2c_n(const float *in, float2 *work, float2 *out, const ROpF &lop, const KOpF &sop) {
    using C = Cfg<N_>;
    {
      ProfScope prof(KK_FFT_R2C_Z, stream);
      r2c_zpass<N_, C::TR><<<(unsigned)((size_t)N_ * N_ / C::TR), C::TR * N_ / 16, C::smem_z, stream>>>(in, work, twN, twM, lop);
      BGPU_LAUNCHED(1);
    }
    strided_n<N_, -1, 1>(work, work, KOpF{}, KOpF{});
    strided_n<N_, -1, 0>(work, out, KOpF{}, sop);
  }
  template <int N_>
  void c2r_n(const float2 *in, float2 *work, float *out, const KOpF &lop, const ROpF &sop) {
    using C = Cfg<N_>;
    strided_n<N_, +1, 0>(in, work, lop, KOpF{});
    strided_n<N_, +1, 1>(work, work, KOpF{}, KOpF{});
    ProfScope prof(KK_FFT_C2R_Z, stream);
    c2r_zpass<N_, C::TR><<<(unsigned)((size_t)N_ * N_ / C::TR), C::TR * N_ / 16, C::smem_z, stream>>>(work, out, twN, twM, sop);
    BGPU_LAUNCHED(1);
  }
#define BGPU_F32_DISPATCH(CALL)                 \
  switch (N) {                                  \
    case 32: CALL(32); break;                   \
    case 64: CALL(64); break;                   \
    case 128: CALL(128); break;                 \
    case 256: CALL(256); break;                 \
    default: CALL(512); break;                  \
  }
  // fftR2C (fftwrapper.cc:56-84): z pass in -> work, y pass in place, x pass work -> out with the store functor
  void r2c(const float *in, float2 *work, float2 *out, const ROpF &lop, const KOpF &sop) {
#define CALL(n) r2c_n<n>(in, work, out, lop, sop)
    BGPU_F32_DISPATCH(CALL)
#undef CALL
  }
  // fftC2R (fftwrapper.cc:26-53): x pass in -> work with the load functor, y pass in place, z pass work -> out
  void c2r(const float2 *in, float2 *work, float *out, const KOpF &lop, const ROpF &sop) {
#define CALL(n) c2r_n<n>(in, work, out, lop, sop)
    BGPU_F32_DISPATCH(CALL)
#undef CALL
  }
};

}  // namespace f32
}  // namespace bgpu

using namespace bgpu;
using namespace bgpu::f32;

enum { S_SUMRHO = 0, S_NLL = 1, S_PRIOR = 2, S_KIN = 3, S_COUNT = 8 };

struct bgpu_f32_handle {
  bgpu_params p{};
  int N = 0;
  size_t n = 0, nh = 0;
  double ncells = 0.0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Fft fft;
  GeomF geom{};
  LikeF like{};
  float kfac = 0.f, normFS = 0.f;
  bool mass_fs = false, mass_rs = false, have_power = false, have_obs = false, have_mass = false;
  float *power = nullptr, *nobs = nullptr, *noise = nullptr, *window = nullptr, *inv_power = nullptr;
  float *mass_f = nullptr, *mass_r = nullptr, *inv_mass = nullptr;
  float *sig = nullptr, *mom = nullptr, *grad = nullptr;
  float *psi[3] = {nullptr, nullptr, nullptr};
  float *delta = nullptr, *resid = nullptr, *tmp = nullptr;
  float2 *shat = nullptr, *dhat = nullptr, *work = nullptr, *acc = nullptr;
  // leapfrog in k-space (as the FP64 path, api.cu leapfrog_kspace): s^ in shat and p^ in phat for the whole trajectory
  float2 *phat = nullptr;
  bool kspace_on = false;   // gradient_device: shat is current, finish with the k-space kick
  bool kspace_lf = true;    // BGPU_LEAPFROG_KSPACE=0: the real-space form
  double *partials = nullptr, *dscal = nullptr, *hscal = nullptr;
  bool kick_on = false;
  float kick_a = 0.f;
};

namespace {

void require(bool ok, const char *msg) {
  if (!ok) throw std::runtime_error(msg);
}
template <class T>
void dalloc(T *&ptr, size_t count) {
  BGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&ptr), count * sizeof(T)));
}
void h2d(bgpu_f32_handle *h, float *dst, const float *src, size_t count) {
  BGPU_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyHostToDevice, h->stream));
}
void d2h(bgpu_f32_handle *h, float *dst, const float *src, size_t count) {
  BGPU_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
}
void sync(bgpu_f32_handle *h) { BGPU_CUDA(cudaStreamSynchronize(h->stream)); }

void r2c_plain(bgpu_f32_handle *h, const float *in, float2 *out) {
  ROpF lop;
  lop.kind = R_LOAD;
  h->fft.r2c(in, h->work == out ? out : h->work, out, lop, KOpF{});
}

// Lag2Eul_zeldovich / _rsd_zeldovich (Lag2Eul.cc:60-136) from s^ in h->shat: Psi -> rho -> sum(rho)
void forward_from_shat(bgpu_f32_handle *h, float dQ, bool rsd) {
  ROpF scale_n;
  scale_n.kind = R_SCALE;
  scale_n.a = (float)(1.0 / h->ncells);
  for (int c = 2; c >= 0; --c) {
    KOpF lop;
    lop.kind = K_DISP;
    lop.comp = c;
    lop.a = (float)(-h->p.D1 * dQ);  // in = dQ * s; phi = -D1 * in (Lag2Eul.cc:88)
    lop.kfac = h->kfac;
    h->fft.c2r(h->shat, h->work, h->psi[c], lop, scale_n);
  }
  GeomF g = h->geom;
  if (!rsd) {
    g.zfac = 1.f;
    g.fgrow = 0.f;
  }
  {
    ProfScope prof(KK_SCATTER, h->stream);
    BGPU_CUDA(cudaMemsetAsync(h->delta, 0, h->n * sizeof(float), h->stream));
    scatter_cic_kernel<<<blocks_for(h->n, kThreads), kThreads, 0, h->stream>>>(g, h->psi[0], h->psi[1], h->psi[2], h->delta);
    BGPU_LAUNCHED(1);
  }
  reduce(SumF{reinterpret_cast<const float4 *>(h->delta)}, h->n / 4, h->partials, h->dscal + S_SUMRHO, h->stream);
}

void residual(bgpu_f32_handle *h, bool exact_sign, float *resid) {
  ProfScope prof(KK_RESIDUAL, h->stream);
  LikeF lp = h->like;
  lp.exact_sign = exact_sign ? 1 : 0;
  const int blocks = grid_for(h->n / 4);
  residual_kernel<<<blocks, kThreads, 0, h->stream>>>(
      lp, reinterpret_cast<float4 *>(h->delta), h->dscal + S_SUMRHO, h->ncells, reinterpret_cast<const float4 *>(h->nobs),
      reinterpret_cast<const float4 *>(h->noise), reinterpret_cast<const float4 *>(h->window),
      reinterpret_cast<float4 *>(resid), h->n / 4, h->partials);
  final_sum_kernel<<<1, kThreads, 0, h->stream>>>(h->partials, blocks, h->dscal + S_NLL);
  BGPU_LAUNCHED(2);
}

// likelihood_grad_log_like + prior + sum (HMC.cc:146-206, HMC_models.cc:377-471): d_out = gradpsi(d_s), or
// d_out += kick_a * gradpsi(d_s) when a leapfrog kick rides on the last store
void gradient_device(bgpu_f32_handle *h, const float *d_s, float *d_out) {
  require(h->have_power && h->have_obs, "bgpu_f32: bgpu_f32_set_static (Power, nobs, noise, window) must be called first");
  const bgpu_params &p = h->p;
  const float inv_n = (float)(1.0 / h->ncells);
  if (!h->kspace_on) r2c_plain(h, d_s, h->shat);
  forward_from_shat(h, (float)p.deltaQ_factor, p.rsd_model != 0);
  residual(h, p.calc_h == BGPU_CALC_H_EXACT, h->resid);
  double norm = -1.0 * p.deltaQ_factor;  // HMC_models.cc:460-469
  if (p.correct_delta) norm *= p.D1;

  ROpF sop;
  sop.kind = h->kick_on ? R_AXPY : R_SCALE;
  sop.a = h->kick_on ? h->kick_a * inv_n : inv_n;

  if (p.calc_h == 1) {
    // h = r (HMC_models.cc:413-415): gradpsi = IFFT[(V/N)/P s^] + norm * r
    KOpF lop;
    lop.kind = K_MULREAL;
    lop.real0 = h->inv_power;
    h->fft.c2r(h->shat, h->work, d_out, lop, sop);
    axpy_kernel<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(d_out, h->resid, (float)norm * (h->kick_on ? h->kick_a : 1.f), h->n);
    BGPU_LAUNCHED(1);
    return;
  }
  KOpF inv;
  inv.kfac = h->kfac;
  if (p.calc_h == 0) {
    // likelihood_calc_h (HMC_models_testing.cpp:25-50): g_c = r * d_c(delta), gradfft for the Gaussian likelihood
    // (gaussian_independent.cpp:44-49), gradfindif for the Poissonian (poissonian.cpp:37-42); then
    // h^ = sum_c (k_c/k^2)(Im g^_c, -Re g^_c) (grad_inv_lap_FS + add_to_array, gradient.cpp:157-211)
    if (p.likelihood == 1) r2c_plain(h, h->delta, h->dhat);
    for (int c = 0; c < 3; ++c) {
      if (p.likelihood == 1) {
        KOpF lop;
        lop.kind = K_GRAD;
        lop.comp = c;
        lop.kfac = h->kfac;
        ROpF mul;
        mul.kind = R_SCALE_MUL;
        mul.a = inv_n;
        mul.aux = h->resid;
        h->fft.c2r(h->dhat, h->work, h->tmp, lop, mul);
      } else {
        ProfScope prof(KK_STREAM, h->stream);
        findif_product_kernel<<<blocks_for(h->n, kThreads), kThreads, 0, h->stream>>>(
            h->delta, h->resid, h->tmp, h->N, h->geom.sh, (float)((double)h->N / (2. * p.L1)), c);
        BGPU_LAUNCHED(1);
      }
      ROpF ld;
      ld.kind = R_LOAD;
      inv.kind = (c == 0) ? K_INVLAP_SET : K_INVLAP_ADD;
      inv.comp = c;
      h->fft.r2c(h->tmp, h->work, h->acc, ld, inv);
    }
  } else {
    // exact adjoint: V = gather(r) in place over Psi, then the same back-projection
    GeomF g = h->geom;
    if (!p.rsd_model) {
      g.zfac = 1.f;
      g.fgrow = 0.f;
    }
    {
      ProfScope prof(KK_GATHER, h->stream);
      gather_cic_kernel<<<blocks_for(h->n, kThreads), kThreads, 0, h->stream>>>(g, h->psi[0], h->psi[1], h->psi[2], h->resid);
      BGPU_LAUNCHED(1);
    }
    for (int c = 0; c < 3; ++c) {
      ROpF ld;
      ld.kind = R_LOAD;
      inv.kind = (c == 0) ? K_INVLAP_SET : K_INVLAP_ADD;
      inv.comp = c;
      h->fft.r2c(h->psi[c], h->work, h->acc, ld, inv);
    }
  }
  if (h->kspace_on) {  // p^ += kick_a ((V/N)/P s^ + norm h^): the kick without gradpsi's inverse transform
    ProfScope prof(KK_STREAM, h->stream);
    kspace_kick_kernel<<<blocks_for(h->nh, 256), 256, 0, h->stream>>>(h->phat, h->shat, h->acc, h->inv_power, h->kick_a,
                                                                     (float)norm, h->nh);
    BGPU_LAUNCHED(1);
    return;
  }
  // gradpsi = IFFT[(V/N)/P s^ + norm * h^]
  KOpF lop;
  lop.kind = K_FINAL;
  lop.a = (float)norm;
  lop.real0 = h->inv_power;
  lop.cplx0 = h->acc;
  h->fft.c2r(h->shat, h->work, d_out, lop, sop);
}

// psi (HMC.cc:124-143): prior 1/2 s.S^-1 s (gaussian.cpp:20-35, by Parseval) and -lnL; leaves deltaX in h->delta
void psi_device(bgpu_f32_handle *h, const float *d_s) {
  require(h->have_power && h->have_obs, "bgpu_f32: bgpu_f32_set_static (Power, nobs, noise, window) must be called first");
  const bgpu_params &p = h->p;
  r2c_plain(h, d_s, h->shat);
  reduce(HalfQuadF{h->shat, h->inv_power, h->N / 2 + 1, 0.5 / h->ncells}, h->nh, h->partials, h->dscal + S_PRIOR, h->stream);
  const bool gauss = p.likelihood == 1;  // the Poisson log-likelihood ignores deltaQ and RSD (poissonian.cpp:54-56)
  forward_from_shat(h, gauss ? (float)p.deltaQ_factor : 1.f, gauss ? (p.rsd_model != 0) : false);
  residual(h, false, nullptr);
}

void kinetic_device(bgpu_f32_handle *h, const float *d_p) {
  require(h->have_mass, "bgpu_f32: bgpu_f32_set_mass or bgpu_f32_hamiltonian_mass must be called first");
  if (h->mass_fs) {
    r2c_plain(h, d_p, h->work);
    reduce(HalfQuadF{h->work, h->inv_mass, h->N / 2 + 1, 0.5 / h->ncells}, h->nh, h->partials, h->dscal + S_KIN, h->stream);
  } else {
    reduce(KineticRealF{d_p, h->mass_r}, h->n, h->partials, h->dscal + S_KIN, h->stream);
  }
}

void kick_device(bgpu_f32_handle *h, const float *d_s, float *d_p, float a) {
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

// Hamiltonian_EoM (HMC.cc:251-369) after the RNG draws, in place on the device, in the fused form of the FP64 path:
// one kick p -= eps * gradpsi between steps (half kicks at the two ends, :293-294 / :351-352) applied by the
// gradient's last z pass, the drift s += eps * M^-1 p (:338-339) by the store of M^-1 p's last z pass.  The
// reference's run-away test |momenta[0]| > 1e50 (:360-364) cannot fire on a finite float; a trajectory that
// overflows ends in inf / nan, which the caller's Metropolis step rejects.
void kspace_kick(bgpu_f32_handle *h, float a) {
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

void leapfrog_device(bgpu_f32_handle *h, float *d_s, float *d_p, uint64_t Neps, float eps) {
  require(h->have_mass, "bgpu_f32: bgpu_f32_set_mass or bgpu_f32_hamiltonian_mass must be called first");
  if (h->kspace_lf && h->mass_fs && h->p.calc_h != 1) {
    // in k-space, as the FP64 path (api.cu leapfrog_kspace): the mode's forward model is Zel'dovich and reads s^ only;
    // a step drops the transform of s, the inverse transform of gradpsi and the pair around M^-1 p
    if (!h->phat) BGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&h->phat), h->nh * sizeof(float2)));
    r2c_plain(h, d_s, h->shat);
    r2c_plain(h, d_p, h->phat);
    kspace_kick(h, -(0.5f * eps));
    for (uint64_t jj = 0; jj < Neps; ++jj) {
      {
        ProfScope prof(KK_STREAM, h->stream);
        kspace_drift_kernel<<<blocks_for(h->nh, 256), 256, 0, h->stream>>>(h->shat, h->phat, h->inv_mass, eps, h->nh);
        BGPU_LAUNCHED(1);
      }
      kspace_kick(h, jj + 1 == Neps ? -(0.5f * eps) : -eps);
    }
    ROpF back;
    back.kind = R_SCALE;
    back.a = (float)(1.0 / h->ncells);
    h->fft.c2r(h->shat, h->work, d_s, KOpF{}, back);
    h->fft.c2r(h->phat, h->work, d_p, KOpF{}, back);
    return;
  }
  kick_device(h, d_s, d_p, -(0.5f * eps));
  for (uint64_t jj = 0; jj < Neps; ++jj) {
    if (h->mass_fs) {
      r2c_plain(h, d_p, h->work);
      KOpF lop;
      lop.kind = K_MULREAL;
      lop.real0 = h->inv_mass;
      ROpF sop;
      sop.kind = R_AXPY;
      sop.a = (float)((double)eps / h->ncells);
      h->fft.c2r(h->work, h->work, d_s, lop, sop);
    } else {
      axpy_div_kernel<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(d_s, d_p, h->mass_r, eps, h->n);
      BGPU_LAUNCHED(1);
    }
    kick_device(h, d_s, d_p, jj + 1 == Neps ? -(0.5f * eps) : -eps);
  }
}

void update_inverse(bgpu_f32_handle *h, const float *full, float *half) {
  inverse_spectrum_kernel<<<blocks_for(h->nh, 256), 256, 0, h->stream>>>(full, half, h->N, h->normFS);
  BGPU_LAUNCHED(1);
}

}  // namespace

#define BGPU_TRY try {
#define BGPU_CATCH                                 \
  }                                                \
  catch (const std::exception &e) {                \
    bgpu::set_last_error(e.what());                \
    return 1;                                      \
  }                                                \
  catch (...) {                                    \
    bgpu::set_last_error("bgpu_f32: unknown error"); \
    return 1;                                      \
  }                                                \
  return 0;

extern "C" {

int bgpu_f32_create(const bgpu_params *p, bgpu_f32_handle **out) {
  bgpu_f32_handle *h = nullptr;
  BGPU_TRY
  require(p && out, "bgpu_f32_create: null argument");
  *out = nullptr;
  require(p->N1 == p->N2 && p->N2 == p->N3, "bgpu_f32: only cubic grids are supported");
  require(p->N1 == 32 || p->N1 == 64 || p->N1 == 128 || p->N1 == 256 || p->N1 == 512,
          "bgpu_f32: N1 must be 32, 64, 128, 256 or 512 in the single-precision mode");
  require(p->L1 == p->L2 && p->L2 == p->L3 && p->L1 > 0, "bgpu_f32: only cubic boxes are supported");
  require(p->xllc == 0. && p->yllc == 0. && p->zllc == 0., "bgpu_f32: the box must start at the origin");
  require(p->periodic != 0, "bgpu_f32: only periodic boundary conditions are supported (disp_part.cc:28)");
  require(p->masskernel == 1, "bgpu_f32: the single-precision mode is built for the CIC mass kernel (masskernel = 1)");
  require(p->likelihood == 0 || p->likelihood == 1,
          "bgpu_f32: the single-precision mode is built for the Poisson (0) and Gaussian (1) likelihoods");
  require(p->sfmodel == 1 || p->rsd_model,
          "bgpu_f32: the single-precision mode is built for the Zel'dovich model (sfmodel = 1, or rsd_model)");
  require(p->calc_h == 0 || p->calc_h == 1 || p->calc_h == BGPU_CALC_H_EXACT,
          "bgpu_f32: calc_h must be 0, 1 or 4 in the single-precision mode");
  require(p->mass_type == 0 || p->mass_type == 1 || p->mass_type == 4,
          "bgpu_f32: mass_type must be 0, 1 or 4 in the single-precision mode");
  if (p->rsd_model) require(p->planepar != 0, "Non-plane-parallel RSD model is not yet implemented in calc_V! Use planepar = true.");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw std::runtime_error(std::string("bgpu: no CUDA device available (") + cudaGetErrorString(e) +
                             "); this path has no CPU fallback");
  require(p->device >= 0 && p->device < ndev, "bgpu_f32: device ordinal out of range");
  BGPU_CUDA(cudaSetDevice(p->device));
  h = new bgpu_f32_handle;
  h->p = *p;
  h->N = p->N1;
  h->ncells = (double)h->N * h->N * h->N;
  h->n = (size_t)h->N * h->N * h->N;
  h->nh = (size_t)h->N * h->N * (h->N / 2 + 1);
  BGPU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  h->fft.init(h->N, h->stream);
  h->kfac = (float)(2. * M_PI / p->L1);                                  // scale_space.cpp:42
  h->normFS = (float)((p->L1 * p->L2 * p->L3) / h->ncells);              // HMC_help.cc:26
  h->mass_fs = p->mass_type != 0;                                        // struct_hamil.h:276-296
  h->mass_rs = p->mass_type == 0;
  {
    const char *kl = std::getenv("BGPU_LEAPFROG_KSPACE");
    h->kspace_lf = !(kl && kl[0] == '0');
  }
  GeomF &g = h->geom;
  g.N = h->N;
  g.sh = 0;
  while ((1 << g.sh) < h->N) ++g.sh;
  g.inv_d = (float)((double)h->N / p->L1);                               // init_par.cc:245
  {
    const double f = host_fgrow(p->ascale, p->OM, p->OL);                // cosmo.cc:182-217
    const double cpecvel = f * 100. * host_E_Hubble_a(p->ascale, p->OM, p->OL) * p->ascale;  // cosmo.cc:232
    const double OC = 1. - p->OM - p->OL;                                // rsd.cc:27-28
    const double Hub = 100. * std::sqrt(p->OM / p->ascale / p->ascale / p->ascale + p->OL + OC / p->ascale / p->ascale);
    const double v_norm = 1. / Hub / p->ascale;                          // rsd.cc:39
    g.zfac = p->rsd_model ? (float)(1. + cpecvel * v_norm) : 1.f;
    g.fgrow = p->rsd_model ? (float)f : 0.f;
  }
  h->like.likelihood = p->likelihood;
  h->like.rho_c = (float)p->rho_c;
  h->like.biasP = (float)p->biasP;
  h->like.biasE = (float)p->biasE;
  h->like.exact_sign = 0;
  for (float **a : {&h->power, &h->nobs, &h->noise, &h->window, &h->sig, &h->mom, &h->grad, &h->psi[0], &h->psi[1],
                    &h->psi[2], &h->delta, &h->resid, &h->tmp})
    dalloc(*a, h->n);
  if (h->mass_fs) {
    dalloc(h->mass_f, h->n);
    dalloc(h->inv_mass, h->nh);
  } else {
    dalloc(h->mass_r, h->n);
  }
  dalloc(h->inv_power, h->nh);
  for (float2 **a : {&h->shat, &h->dhat, &h->work, &h->acc}) dalloc(*a, h->nh);
  dalloc(h->partials, (size_t)kBlocks);
  dalloc(h->dscal, (size_t)S_COUNT);
  BGPU_CUDA(cudaMemsetAsync(h->dscal, 0, S_COUNT * sizeof(double), h->stream));
  BGPU_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h->hscal), S_COUNT * sizeof(double)));
  BGPU_CUDA(cudaStreamSynchronize(h->stream));
  *out = h;
  h = nullptr;
  }
  catch (const std::exception &e) {
    bgpu::set_last_error(e.what());
    if (h) bgpu_f32_destroy(h);
    return 1;
  }
  catch (...) {
    bgpu::set_last_error("bgpu_f32: unknown error");
    if (h) bgpu_f32_destroy(h);
    return 1;
  }
  return 0;
}

void bgpu_f32_destroy(bgpu_f32_handle *h) {
  if (!h) return;
  cudaSetDevice(h->p.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (float *a : {h->power, h->nobs, h->noise, h->window, h->inv_power, h->mass_f, h->mass_r, h->inv_mass, h->sig, h->mom,
                   h->grad, h->psi[0], h->psi[1], h->psi[2], h->delta, h->resid, h->tmp})
    cudaFree(a);
  for (float2 *a : {h->shat, h->dhat, h->work, h->acc, h->phat}) cudaFree(a);
  cudaFree(h->partials);
  cudaFree(h->dscal);
  if (h->hscal) cudaFreeHost(h->hscal);
  h->fft.destroy();
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int bgpu_f32_set_stream(bgpu_f32_handle *h, void *cuda_stream) {
  BGPU_TRY
  require(h != nullptr, "bgpu_f32_set_stream: null handle");
  BGPU_CUDA(cudaStreamSynchronize(h->stream));
  if (h->own_stream) BGPU_CUDA(cudaStreamDestroy(h->stream));
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  h->own_stream = false;
  h->fft.stream = h->stream;
  BGPU_CATCH
}

int bgpu_f32_synchronize(bgpu_f32_handle *h) {
  BGPU_TRY
  sync(h);
  BGPU_CATCH
}

int bgpu_f32_set_static(bgpu_f32_handle *h, const float *Power, const float *nobs, const float *noise,
                        const float *window) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
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

int bgpu_f32_set_mass(bgpu_f32_handle *h, const float *mass_f, const float *mass_r) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  if (h->mass_fs) {
    require(mass_f != nullptr, "bgpu_f32_set_mass: mass_f is required for this mass_type");
    h2d(h, h->mass_f, mass_f, h->n);
    update_inverse(h, h->mass_f, h->inv_mass);
  } else {
    require(mass_r != nullptr, "bgpu_f32_set_mass: mass_r is required for this mass_type");
    h2d(h, h->mass_r, mass_r, h->n);
  }
  h->have_mass = true;
  sync(h);
  BGPU_CATCH
}

int bgpu_f32_hamiltonian_mass(bgpu_f32_handle *h, float *mass_f_out, float *mass_r_out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(h->have_power || h->p.mass_type == 0, "bgpu_f32_hamiltonian_mass: Power must be set first");
  mass_kernel<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->power, h->mass_f, h->mass_r, h->p.mass_type,
                                                            (float)h->p.mass_factor, h->n);
  BGPU_LAUNCHED(1);
  if (h->mass_fs) update_inverse(h, h->mass_f, h->inv_mass);
  h->have_mass = true;
  if (mass_f_out && h->mass_fs) d2h(h, mass_f_out, h->mass_f, h->n);
  if (mass_r_out && h->mass_rs) d2h(h, mass_r_out, h->mass_r, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_f32_gradient_psi_dev(bgpu_f32_handle *h, const float *d_signal, float *d_gradpsi) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  gradient_device(h, d_signal, d_gradpsi);
  BGPU_CATCH
}

int bgpu_f32_gradient_psi(bgpu_f32_handle *h, const float *signal, float *gradpsi) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(signal && gradpsi, "bgpu_f32_gradient_psi: null argument");
  h2d(h, h->sig, signal, h->n);
  gradient_device(h, h->sig, h->grad);
  d2h(h, gradpsi, h->grad, h->n);
  sync(h);
  BGPU_CATCH
}

int bgpu_f32_psi_dev(bgpu_f32_handle *h, const float *d_signal, double *psi_prior, double *psi_likeli, float *d_deltaX) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  psi_device(h, d_signal);
  if (d_deltaX) BGPU_CUDA(cudaMemcpyAsync(d_deltaX, h->delta, h->n * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  BGPU_CUDA(cudaMemcpyAsync(h->hscal, h->dscal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  if (psi_prior) *psi_prior = h->hscal[S_PRIOR];
  if (psi_likeli) *psi_likeli = h->hscal[S_NLL];
  BGPU_CATCH
}

int bgpu_f32_psi(bgpu_f32_handle *h, const float *signal, double *psi_prior, double *psi_likeli, float *deltaX_out) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(signal != nullptr, "bgpu_f32_psi: null argument");
  h2d(h, h->sig, signal, h->n);
  psi_device(h, h->sig);
  if (deltaX_out) d2h(h, deltaX_out, h->delta, h->n);
  BGPU_CUDA(cudaMemcpyAsync(h->hscal, h->dscal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  if (psi_prior) *psi_prior = h->hscal[S_PRIOR];
  if (psi_likeli) *psi_likeli = h->hscal[S_NLL];
  BGPU_CATCH
}

int bgpu_f32_kinetic(bgpu_f32_handle *h, const float *momenta, double *K) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(momenta && K, "bgpu_f32_kinetic: null argument");
  h2d(h, h->mom, momenta, h->n);
  kinetic_device(h, h->mom);
  BGPU_CUDA(cudaMemcpyAsync(h->hscal, h->dscal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  *K = h->hscal[S_KIN];
  BGPU_CATCH
}

int bgpu_f32_leapfrog_dev(bgpu_f32_handle *h, float *d_signal, float *d_momenta, uint64_t Neps, double epsilon) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  leapfrog_device(h, d_signal, d_momenta, Neps, (float)epsilon);
  BGPU_CATCH
}

int bgpu_f32_leapfrog(bgpu_f32_handle *h, const float *s_i, const float *p_i, uint64_t Neps, double epsilon, float *s_f,
                      float *p_f) {
  BGPU_TRY
  BGPU_CUDA(cudaSetDevice(h->p.device));
  require(s_i && p_i && s_f && p_f, "bgpu_f32_leapfrog: null argument");
  h2d(h, h->sig, s_i, h->n);
  h2d(h, h->mom, p_i, h->n);
  leapfrog_device(h, h->sig, h->mom, Neps, (float)epsilon);
  d2h(h, s_f, h->sig, h->n);
  d2h(h, p_f, h->mom, h->n);
  sync(h);
  BGPU_CATCH
}

}  // extern "C"
