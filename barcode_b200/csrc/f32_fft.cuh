// barcode_b200/csrc/f32_fft.cuh -- the 3-D real FFT of the single-precision mode (reference build option
// SINGLE_PREC: real_prec = float, fftwf_* behind fftR2C / fftC2R, define_opt.h:50-59, fftwrapper.cc:32-36,62-66).
//
// Same layout and conventions as the FP64 transform (fft.cuh): real arrays float[N][N][N] with z fastest,
// half-complex arrays float2[N][N][N/2+1], FOURIER_DEF_2 (forward unnormalised, the 1/N of the inverse folded into a
// store functor), three pencil passes per transform (z, y, x / x, y, z), Stockham autosort with eight elements per
// thread in registers and shared memory only for the exchanges between stages, k-space functors on the x pass and
// real-space functors on the z pass.
//
// What single precision changes on B200: a half-complex array of 256^3 is 68 MB and fits the 126 MB L2, so a pass
// mostly reads what the previous pass left there; an element is 8 bytes, so a tile of 16 z-adjacent pencils makes
// the 128-byte rows the memory system wants (the FP64 passes use 8).  The strided pass is persistent and
// software-pipelined with cp.async (8-byte copies into thread-private landing slots, no barrier on the pipeline).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_ops.h"

namespace bgpu {
namespace f32 {

// functor descriptors: the kinds of fft_ops.h with single-precision operands
struct KOpF {
  int kind = K_NONE;
  int comp = 0;
  float a = 1.f;
  float kfac = 0.f;
  const float *real0 = nullptr;   // [N][N][N/2+1]
  const float2 *cplx0 = nullptr;  // [N][N][N/2+1]
};
struct ROpF {
  int kind = R_SCALE;
  float a = 1.f;
  const float *aux = nullptr;
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
template <int DIR>
__device__ __forceinline__ float2 mul_j(float2 a) {
  return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// tw[k] = exp(-2 pi i k / n), rounded from a double table; the inverse conjugates
template <int DIR>
__device__ __forceinline__ float2 twiddle(const float2 *__restrict__ tw, int idx) {
  float2 w = __ldg(tw + idx);
  if (DIR > 0) w.y = -w.y;
  return w;
}

template <int DIR>
__device__ __forceinline__ void bf2(float2 &a, float2 &b) {
  const float2 t = csub(a, b);
  a = cadd(a, b);
  b = t;
}
template <int DIR>
__device__ __forceinline__ void bf4(float2 &c0, float2 &c1, float2 &c2, float2 &c3) {
  const float2 s02 = cadd(c0, c2), d02 = csub(c0, c2);
  const float2 s13 = cadd(c1, c3), d13 = mul_j<DIR>(csub(c1, c3));
  c0 = cadd(s02, s13);
  c2 = csub(s02, s13);
  c1 = cadd(d02, d13);
  c3 = csub(d02, d13);
}
template <int DIR>
__device__ __forceinline__ void bf8(float2 (&v)[8]) {
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
  float2 a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
  float2 a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
  float2 a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
  if (DIR < 0) {
    b1 = make_float2(h * (b1.x + b1.y), h * (b1.y - b1.x));
    b3 = make_float2(h * (b3.y - b3.x), -h * (b3.x + b3.y));
  } else {
    b1 = make_float2(h * (b1.x - b1.y), h * (b1.y + b1.x));
    b3 = make_float2(-h * (b3.x + b3.y), h * (b3.x - b3.y));
  }
  b2 = mul_j<DIR>(b2);
  bf4<DIR>(a0, a1, a2, a3);
  bf4<DIR>(b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

template <int N, int S>
struct StageRadix {
  static constexpr int rem = N / S;
  static constexpr int value = rem >= 8 ? 8 : rem;
};

// One Stockham stage on the eight register elements of thread t (v[m] = element t + m N/8 on entry and on exit).
// WARP: the N/8 threads of a transform sit in one warp (the z passes: a row is M/8 <= 32 lanes), so the exchange
// needs __syncwarp only and a CTA never waits on a barrier; otherwise __syncthreads.
template <bool WARP>
__device__ __forceinline__ void stage_sync() {
  if constexpr (WARP) __syncwarp();
  else __syncthreads();
}

template <int N, int S, int DIR, bool WARP, class Sm>
__device__ __forceinline__ void fft_stages(float2 (&v)[8], int t, const float2 *__restrict__ tw, Sm sm) {
  constexpr int R = StageRadix<N, S>::value;
  constexpr int NB = 8 / R;
  constexpr bool last = (S * R == N);
  static_assert(R == 2 || R == 4 || R == 8, "bad radix");
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if constexpr (R == 8) bf8<DIR>(v);
    else if constexpr (R == 4) bf4<DIR>(v[j], v[j + NB], v[j + 2 * NB], v[j + 3 * NB]);
    else bf2<DIR>(v[j], v[j + NB]);
  }
  if constexpr (!last) {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int b = t + j * (N / 8);
      const int q = b & (S - 1);
      const int base = b - q;
      // (powers w^2 ... w^7 formed by multiplication instead of six more table loads were measured: no faster, and the
      // gradient's distance from FP64 doubled, 1.1e-6 -> 2.4e-6 at 256^3)
#pragma unroll
      for (int k = 1; k < R; ++k) v[j + k * NB] = cmul(v[j + k * NB], twiddle<DIR>(tw, base * k));
#pragma unroll
      for (int k = 0; k < R; ++k) sm.store(q + R * base + k * S, v[j + k * NB]);
    }
    stage_sync<WARP>();
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = sm.load(t + m * (N / 8));
    stage_sync<WARP>();
    fft_stages<N, S * R, DIR, WARP, Sm>(v, t, tw, sm);
  }
}

struct SmStrided {  // [element][pencil]
  float2 *base;
  int T, p;
  __device__ __forceinline__ void store(int e, float2 x) const { base[e * T + p] = x; }
  __device__ __forceinline__ float2 load(int e) const { return base[e * T + p]; }
};
struct SmRow {  // one contiguous row, one pad element every 8
  float2 *row;
  __device__ __forceinline__ void store(int e, float2 x) const { row[e + (e >> 3)] = x; }
  __device__ __forceinline__ float2 load(int e) const { return row[e + (e >> 3)]; }
};

// scale_space.cpp:41-51
__device__ __forceinline__ float kval(int i, int N, float kfac) {
  return (i <= N / 2) ? kfac * (float)i : -kfac * (float)(N - i);
}

// load functors that need no operand array (fft.cuh kop_load; EqSolvers.cc:208-268, gradient.cpp:38-74)
template <int N>
__device__ __forceinline__ float2 kop_load(const KOpF &op, float2 v, int ix, int iy, int iz) {
  switch (op.kind) {
    case K_DISP: {
      if (ix == N / 2 || iy == N / 2 || iz == N / 2) return make_float2(0.f, 0.f);
      const float kx = kval(ix, N, op.kfac), ky = kval(iy, N, op.kfac), kz = kval(iz, N, op.kfac);
      const float ksq = kx * kx + ky * ky + kz * kz;
      if (!(ksq > 1.e-14f)) return make_float2(0.f, 0.f);
      const float kc = op.comp == 0 ? kx : (op.comp == 1 ? ky : kz);
      const float f = op.a * (kc / ksq);
      return make_float2(f * v.y, f * -v.x);
    }
    case K_GRAD: {
      if (ix == N / 2 || iy == N / 2 || iz == N / 2) return make_float2(0.f, 0.f);
      const float kc = kval(op.comp == 0 ? ix : (op.comp == 1 ? iy : iz), N, op.kfac);
      return make_float2(-kc * v.y, kc * v.x);
    }
    default:
      return v;
  }
}

// store functors (fft.cuh kop_store; grad_inv_lap_FS + add_to_array, gradient.cpp:167-210)
template <int N>
__device__ __forceinline__ void kop_store(const KOpF &op, float2 *out, float2 v, size_t off, int ix, int iy, int iz) {
  if (op.kind == K_INVLAP_SET || op.kind == K_INVLAP_ADD) {
    float2 r = make_float2(0.f, 0.f);
    if (!(ix == N / 2 || iy == N / 2 || iz == N / 2)) {
      const float kx = kval(ix, N, op.kfac), ky = kval(iy, N, op.kfac), kz = kval(iz, N, op.kfac);
      const float ksq = kx * kx + ky * ky + kz * kz;
      if (ksq > 0.f) {
        const float kc = op.comp == 0 ? kx : (op.comp == 1 ? ky : kz);
        const float f = kc / ksq;
        r = make_float2(f * v.y, -f * v.x);
      }
    }
    if (op.kind == K_INVLAP_ADD) {
      const float2 o = out[off];
      r.x += o.x;
      r.y += o.y;
    }
    out[off] = r;
    return;
  }
  out[off] = v;
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// tile -> (index along the other strided axis, z index of pencil p, offset of the pencil's first element).  The
// Nyquist plane z = N/2 does not fit the z tiling (N/2+1 is odd): extra tiles take T pencils of it that are adjacent
// along the other axis.
template <int N, int T, int AXIS>
__device__ __forceinline__ void strided_tile_coords(int tile, int p, int &other, int &iz, size_t &base) {
  constexpr int NZH = N / 2 + 1;
  constexpr int NTZ = (N / 2) / T;
  constexpr int NMAIN = N * NTZ;
  if (tile < NMAIN) {
    other = tile / NTZ;
    iz = (tile % NTZ) * T + p;
  } else {
    other = (tile - NMAIN) * T + p;
    iz = N / 2;
  }
  base = ((AXIS == 0) ? (size_t)other * NZH : (size_t)other * N * NZH) + iz;
}

// ---------------------------------------------------------------------------
// strided pass (AXIS 0 = x, 1 = y), persistent and software-pipelined: before a CTA transforms tile i every thread
// issues the cp.async copies of the eight elements IT will own in tile i + gridDim.x.  in == out is allowed (a tile
// is read completely before it is written, and tiles are disjoint).
// ---------------------------------------------------------------------------
template <int N, int T, int DIR, int AXIS, int PD = 1>
__global__ void __launch_bounds__(T *N / 8, (T * N / 8 >= 1024) ? 1 : ((1024 / (T * N / 8) > 8) ? 8 : 1024 / (T * N / 8)))
    strided_pass(const float2 *in, float2 *out, const float2 *__restrict__ tw, KOpF lop, KOpF sop) {
  // PD = tiles in flight ahead of the one being transformed (1 or 2): shared memory holds the exchange tile and PD
  // prefetch tiles.  A thread reads back only the elements it copied itself, so cp.async.wait_group is all the
  // ordering the ring needs.
  extern __shared__ float2 smem_f32[];
  constexpr int NZH = N / 2 + 1;
  constexpr int NTILES = N * ((N / 2) / T) + N / T;
  float2 *xch = smem_f32;
  float2 *pre = smem_f32 + N * T;
  const int p = threadIdx.x % T;
  const int t = threadIdx.x / T;
  constexpr size_t stride = (AXIS == 0) ? (size_t)N * NZH : (size_t)NZH;

  auto prefetch = [&](int tl, float2 *buf) {
    if (tl < NTILES) {
      int o2, z2;
      size_t b2;
      strided_tile_coords<N, T, AXIS>(tl, p, o2, z2, b2);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = t + m * (N / 8);
        cp_async8(buf + r * T + p, in + b2 + (size_t)r * stride);
      }
    }
    cp_async_commit();
  };

  int tile = blockIdx.x;
  int other = 0, iz = 0, slot = 0;
  size_t base = 0;
  if (tile < NTILES) strided_tile_coords<N, T, AXIS>(tile, p, other, iz, base);
#pragma unroll
  for (int a = 0; a < PD; ++a) prefetch(tile + a * (int)gridDim.x, pre + a * N * T);

  while (tile < NTILES) {
    if constexpr (PD == 1) cp_async_wait_all();
    else asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    float2 *cur = pre + slot * N * T;
    float2 v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = cur[(t + m * (N / 8)) * T + p];
    const int next = tile + gridDim.x;
    prefetch(tile + PD * (int)gridDim.x, cur);
    if constexpr (PD == 2) slot ^= 1;

    if (lop.kind == K_MULREAL || lop.kind == K_FINAL) {
      // v * real0 (HMC_help.cc:41-58, the multiplier precomputed) [+ a * cplx0: prior + norm * h, HMC.cc:205]
      float f[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) f[m] = __ldg(lop.real0 + base + (size_t)(t + m * (N / 8)) * stride);
      if (lop.kind == K_FINAL) {
        float2 hh[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) hh[m] = __ldg(lop.cplx0 + base + (size_t)(t + m * (N / 8)) * stride);
#pragma unroll
        for (int m = 0; m < 8; ++m)
          v[m] = make_float2(v[m].x * f[m] + lop.a * hh[m].x, v[m].y * f[m] + lop.a * hh[m].y);
      } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = make_float2(v[m].x * f[m], v[m].y * f[m]);
      }
    } else if (lop.kind != K_NONE) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = t + m * (N / 8);
        v[m] = kop_load<N>(lop, v[m], (AXIS == 0) ? r : other, (AXIS == 0) ? other : r, iz);
      }
    }

    SmStrided sm{xch, T, p};
    fft_stages<N, 1, DIR, false, SmStrided>(v, t, tw, sm);

#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int r = t + m * (N / 8);
      const size_t off = base + (size_t)r * stride;
      if (sop.kind != K_NONE) kop_store<N>(sop, out, v[m], off, (AXIS == 0) ? r : other, (AXIS == 0) ? other : r, iz);
      else out[off] = v[m];
    }
    tile = next;
    if (tile < NTILES) strided_tile_coords<N, T, AXIS>(tile, p, other, iz, base);
  }
}

// Measured and dropped (B200, 256^3, round 2): a warp-synchronous form of the strided pass -- the CTA copies a tile in
// with coalesced cp.async rows, each warp transforms its own pencil column in place with __syncwarp only, the CTA
// stores the tile out with coalesced rows -- took 47.7 us per y pass against 40.1 us for the kernel above (x: 65 against
// 53 us): three CTA barriers per tile and nine trips through shared memory per element cost more than the barriers
// inside the exchanges it removes.

// ---------------------------------------------------------------------------
// z pass, real -> half-complex: N reals as M = N/2 complex z[j] = x[2j] + i x[2j+1], then
// X[k] = E[k] + w_N^k O[k], E = (Z[k] + conj Z[M-k])/2, O = (Z[k] - conj Z[M-k])/(2i).
// A CTA owns TR consecutive rows; threads TR * M/8.  A row's M/8 <= 32 threads are lanes of ONE warp (M/8 divides
// 32), so every exchange of the pass is warp-synchronous: no CTA barrier anywhere in the z passes.
// ---------------------------------------------------------------------------
template <int N, int TR>
__global__ void __launch_bounds__(TR *N / 16)
    r2c_zpass(const float *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ twN,
              const float2 *__restrict__ twM, ROpF lop) {
  extern __shared__ float2 smem_f32[];
  constexpr int M = N / 2;
  constexpr int ROWP = M + M / 8 + 1;
  const int t = threadIdx.x % (M / 8);
  const int rl = threadIdx.x / (M / 8);
  const size_t row = (size_t)blockIdx.x * TR + rl;
  const float2 *src = reinterpret_cast<const float2 *>(in + row * N);
  float2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    float2 x = src[t + m * (M / 8)];
    if (lop.kind == R_LOAD_SCALE) {
      x.x *= lop.a;
      x.y *= lop.a;
    }
    v[m] = x;
  }
  SmRow sm{smem_f32 + (size_t)rl * ROWP};
  fft_stages<M, 1, -1, true, SmRow>(v, t, twM, sm);
#pragma unroll
  for (int m = 0; m < 8; ++m) sm.store(t + m * (M / 8), v[m]);
  __syncwarp();
  float2 *dst = out + row * (M + 1);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = t + m * (M / 8);
    const float2 zk = v[m];
    const float2 zm = sm.load((M - k) & (M - 1));
    const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    const float2 o = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
    const float2 w = __ldg(twN + k);
    dst[k] = cadd(e, cmul(w, o));
    if (k == 0) dst[M] = make_float2(e.x - o.x, 0.f);
  }
}

// ---------------------------------------------------------------------------
// z pass, half-complex -> real (FFTW c2r semantics times the store functor's scale): Z[k] = E' + i O',
// E' = X[k] + conj X[M-k], O' = (X[k] - conj X[M-k]) conj(w_N^k); inverse M-point FFT.
// `skip`: a stopped leapfrog trajectory stores nothing (HMC.cc:360-364).
// ---------------------------------------------------------------------------
template <int N, int TR>
__global__ void __launch_bounds__(TR *N / 16)
    c2r_zpass(const float2 *__restrict__ in, float *out, const float2 *__restrict__ twN,
              const float2 *__restrict__ twM, ROpF sop) {
  extern __shared__ float2 smem_f32[];
  constexpr int M = N / 2;
  constexpr int ROWP = M + M / 8 + 1;
  const int t = threadIdx.x % (M / 8);
  const int rl = threadIdx.x / (M / 8);
  const size_t row = (size_t)blockIdx.x * TR + rl;
  const float2 *src = in + row * (M + 1);
  SmRow sm{smem_f32 + (size_t)rl * ROWP};
#pragma unroll
  for (int m = 0; m < 8; ++m) sm.store(t + m * (M / 8), src[t + m * (M / 8)]);
  if (t == 0) sm.store(M, src[M]);
  __syncwarp();
  float2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = t + m * (M / 8);
    float2 xk = sm.load(k);
    float2 xm = sm.load(M - k);
    if (k == 0) {
      xk.y = 0.f;
      xm.y = 0.f;
    }
    const float2 e = make_float2(xk.x + xm.x, xk.y - xm.y);
    const float2 d = make_float2(xk.x - xm.x, xk.y + xm.y);
    const float2 o = cmul(d, cconj(__ldg(twN + k)));
    v[m] = make_float2(e.x - o.y, e.y + o.x);
  }
  __syncwarp();
  fft_stages<M, 1, +1, true, SmRow>(v, t, twM, sm);
  float2 *dst = reinterpret_cast<float2 *>(out + row * N);
  const float2 *aux = sop.aux ? reinterpret_cast<const float2 *>(sop.aux + row * N) : nullptr;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int j = t + m * (M / 8);
    float2 x = make_float2(sop.a * v[m].x, sop.a * v[m].y);
    if (sop.kind == R_SCALE_MUL) {
      const float2 y = aux[j];
      x.x *= y.x;
      x.y *= y.y;
    } else if (sop.kind == R_AXPY) {
      const float2 o = dst[j];
      x.x += o.x;
      x.y += o.y;
    }
    dst[j] = x;
  }
}

}  // namespace f32
}  // namespace bgpu
