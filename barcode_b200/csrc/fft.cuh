// barcode_b200/csrc/fft.cuh
//
// Hand-written FP64 3-D real FFT for sm_100a, replacing the FFTW calls behind
// the reference's fftR2C / fftC2R / fftR2Cplanned / fftC2Rplanned
// (/root/reference/barlib/src/fftwrapper.cc:26-125) on the HMC hot path.
//
// Layout follows the reference: real arrays double[N1][N2][N3] with z fastest,
// half-complex arrays double2[N1][N2][N3/2+1] (HMC_help.cc:45).  Convention
// FOURIER_DEF_2: forward unnormalised, inverse carries 1/N (fftwrapper.cc:
// 100-101) -- the 1/N is folded into the store functor of the last pass.
//
// A 3-D transform is three pencil passes (z: real<->half-complex along the
// contiguous axis, then y, then x; the inverse runs x, y, z).  Every pass is a
// Stockham autosort FFT with eight elements per thread held in registers
// (radix 8, with one radix-4/2 stage when log2 n is not a multiple of 3).  In
// the Stockham formulation thread t owns elements t + m*n/8 (m = 0..7) both
// before the first and after the last stage, so the first stage reads global
// memory straight into registers and the last stage stores straight from
// registers -- coalesced in both directions -- and shared memory is touched
// only for the exchanges between stages.  k-space work (displacement kernel,
// spectral gradient, inverse power spectrum, -ik/k^2 back-projection) rides on
// the x pass as load/store functors, real-space work (1/N, residual product,
// leapfrog axpy) on the z pass, so those never cost a pass over HBM of their
// own.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_ops.h"

namespace bgpu {

// ---------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by exp(DIR * i*pi/2): forward (DIR=-1) is -i, backward is +i
template <int DIR>
__device__ __forceinline__ double2 mul_j(double2 a) {
  return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}

// twiddle table: tw[k] = exp(-2 pi i k / n); the inverse uses the conjugate
template <int DIR>
__device__ __forceinline__ double2 twiddle(const double2 *__restrict__ tw, int idx) {
  double2 w = __ldg(tw + idx);
  if (DIR > 0) w.y = -w.y;
  return w;
}

// ---------------------------------------------------------------------------
// radix butterflies (decimation in frequency, outputs in natural order)
// ---------------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ void bf2(double2 &a, double2 &b) {
  double2 t = csub(a, b);
  a = cadd(a, b);
  b = t;
}

template <int DIR>
__device__ __forceinline__ void bf4(double2 &c0, double2 &c1, double2 &c2, double2 &c3) {
  double2 s02 = cadd(c0, c2), d02 = csub(c0, c2);
  double2 s13 = cadd(c1, c3), d13 = mul_j<DIR>(csub(c1, c3));
  c0 = cadd(s02, s13);
  c2 = csub(s02, s13);
  c1 = cadd(d02, d13);
  c3 = csub(d02, d13);
}

template <int DIR>
__device__ __forceinline__ void bf8(double2 (&v)[8]) {
  const double h = 0.70710678118654752440;
  double2 a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
  double2 a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
  double2 a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
  double2 a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
  // b_k *= w8^k, w8 = exp(DIR*i*pi/4)
  if (DIR < 0) {
    b1 = make_double2(h * (b1.x + b1.y), h * (b1.y - b1.x));
    b3 = make_double2(h * (b3.y - b3.x), -h * (b3.x + b3.y));
  } else {
    b1 = make_double2(h * (b1.x - b1.y), h * (b1.y + b1.x));
    b3 = make_double2(-h * (b3.x + b3.y), h * (b3.x - b3.y));
  }
  b2 = mul_j<DIR>(b2);
  bf4<DIR>(a0, a1, a2, a3);  // even outputs 0,2,4,6
  bf4<DIR>(b0, b1, b2, b3);  // odd outputs 1,3,5,7
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// ---------------------------------------------------------------------------
// one Stockham stage on the eight register elements of thread t.
//   N: transform length, S: product of the radices already applied.
//   v[m] holds element t + m*N/8 on entry; the stage's results are written to
//   shared memory through `Sm` (unless it is the last stage, which leaves
//   v[m] = element t + m*N/8 of the output).
// ---------------------------------------------------------------------------
template <int N, int S>
struct StageRadix {
  static constexpr int rem = N / S;
  static constexpr int value = rem >= 8 ? 8 : rem;
};

template <int N, int S, int DIR, class Sm>
__device__ __forceinline__ void fft_stages(double2 (&v)[8], int t, const double2 *__restrict__ tw, Sm sm) {
  constexpr int R = StageRadix<N, S>::value;
  constexpr int NB = 8 / R;          // butterflies per thread in this stage
  constexpr bool last = (S * R == N);
  static_assert(R == 2 || R == 4 || R == 8, "bad radix");

  // butterflies: j-th butterfly owns register elements m = j + k*NB, k < R
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if constexpr (R == 8) {
      bf8<DIR>(v);
    } else if constexpr (R == 4) {
      bf4<DIR>(v[j], v[j + NB], v[j + 2 * NB], v[j + 3 * NB]);
    } else {
      bf2<DIR>(v[j], v[j + NB]);
    }
  }
  if constexpr (!last) {
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int b = t + j * (N / 8);
    const int q = b & (S - 1);
    const int base = b - q;  // S * p
#pragma unroll
    for (int k = 1; k < R; ++k) {
      v[j + k * NB] = cmul(v[j + k * NB], twiddle<DIR>(tw, base * k));
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      sm.store(q + R * base + k * S, v[j + k * NB]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = sm.load(t + m * (N / 8));
  __syncthreads();
  fft_stages<N, S * R, DIR, Sm>(v, t, tw, sm);
  }  // last stage: twiddles are all one (base == 0) and v[m] already is output element t + m*N/8
}

// shared-memory accessors
struct SmStrided {  // layout [element][pencil], T pencils
  double2 *base;
  int T, p;
  __device__ __forceinline__ void store(int e, double2 x) const { base[e * T + p] = x; }
  __device__ __forceinline__ double2 load(int e) const { return base[e * T + p]; }
};

struct SmRow {  // one contiguous row, one pad element every 8 (kills the stride-8 bank conflict of stage 1)
  double2 *row;
  __device__ __forceinline__ void store(int e, double2 x) const { row[e + (e >> 3)] = x; }
  __device__ __forceinline__ double2 load(int e) const { return row[e + (e >> 3)]; }
};

// ---------------------------------------------------------------------------
// k-space functor evaluation
// ---------------------------------------------------------------------------
__device__ __forceinline__ double kval(int i, int N, double kfac) {
  // scale_space.cpp:41-51
  return (i <= N / 2) ? kfac * (double)i : -kfac * (double)(N - i);
}

template <int N>
__device__ __forceinline__ double2 kop_load(const KOp &op, double2 v, size_t off, int ix, int iy, int iz) {
  switch (op.kind) {
    case K_NONE:
      return v;
    // K_MULREAL / K_FINAL read operand arrays and are handled by the pass kernels themselves
    case K_DISP: {
      if (ix == N / 2 || iy == N / 2 || iz == N / 2) return make_double2(0.0, 0.0);
      const double kx = kval(ix, N, op.kfac), ky = kval(iy, N, op.kfac), kz = kval(iz, N, op.kfac);
      const double ksq = kx * kx + ky * ky + kz * kz;
      if (!(ksq > 1.e-14)) return make_double2(0.0, 0.0);
      const double kc = op.comp == 0 ? kx : (op.comp == 1 ? ky : kz);
      const double f = op.a * (__drcp_rn(ksq) * kc);
      return make_double2(f * v.y, f * -v.x);
    }
    case K_GRAD: {
      if (ix == N / 2 || iy == N / 2 || iz == N / 2) return make_double2(0.0, 0.0);
      const double kc = kval(op.comp == 0 ? ix : (op.comp == 1 ? iy : iz), N, op.kfac);
      return make_double2(-kc * v.y, kc * v.x);
    }
    case K_NEGINVK2: {
      const double kx = kval(ix, N, op.kfac), ky = kval(iy, N, op.kfac), kz = kval(iz, N, op.kfac);
      const double ksq = kx * kx + ky * ky + kz * kz;
      const double f = ksq > 0.0 ? -op.a * __drcp_rn(ksq) : 0.0;
      return make_double2(f * v.x, f * v.y);
    }
    case K_GAUSS:
    case K_ONE_MINUS_GAUSS: {
      const double kx = kval(ix, N, op.kfac), ky = kval(iy, N, op.kfac), kz = kval(iz, N, op.kfac);
      const double K = exp(-(kx * kx + ky * ky + kz * kz) * (op.a * op.a) / 2.);
      const double f = op.kind == K_GAUSS ? K : 1.0 - K;
      return make_double2(f * v.x, f * v.y);
    }
    default:
      return v;
  }
}

template <int N>
__device__ __forceinline__ void kop_store(const KOp &op, double2 *__restrict__ out, double2 v, size_t off, int ix,
                                          int iy, int iz) {
  switch (op.kind) {
    case K_INVLAP_SET:
    case K_INVLAP_ADD: {
      double2 r = make_double2(0.0, 0.0);
      if (!(ix == N / 2 || iy == N / 2 || iz == N / 2)) {
        const double kx = kval(ix, N, op.kfac), ky = kval(iy, N, op.kfac), kz = kval(iz, N, op.kfac);
        const double ksq = kx * kx + ky * ky + kz * kz;
        if (ksq > 0.0) {
          const double kc = op.comp == 0 ? kx : (op.comp == 1 ? ky : kz);
          const double f = kc * __drcp_rn(ksq);
          r = make_double2(f * v.y, -f * v.x);
        }
      }
      if (op.kind == K_INVLAP_ADD) {
        const double2 o = out[off];
        r.x += o.x;
        r.y += o.y;
      }
      out[off] = r;
      return;
    }
    default:
      out[off] = v;
  }
}

// ---------------------------------------------------------------------------
// strided pass (axis 0 = x, axis 1 = y) over a half-complex array [N][N][N/2+1]
//
// A CTA transforms T pencils that are adjacent along z (T*16 B contiguous per
// axis index).  The Nyquist plane z = N/2 does not fit that tiling (N/2+1 is
// odd), so it is covered by extra CTAs whose T pencils are adjacent along the
// other non-transformed axis instead; those touch 1/(N/2+1) of the data.
// Threads: T * N/8, thread = p + T*t (pencil fastest).
// ---------------------------------------------------------------------------
template <int N, int T, int DIR, int AXIS>
__global__ void __launch_bounds__(T *N / 8)
    fft_strided_pass(const double2 *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ tw,
                     KOp lop, KOp sop) {
  extern __shared__ double2 smem[];
  constexpr int NZH = N / 2 + 1;
  constexpr int NTZ = (N / 2) / T;     // z tiles per (other) index
  constexpr int NMAIN = N * NTZ;
  const int p = threadIdx.x % T;
  const int t = threadIdx.x / T;
  const int tile = blockIdx.x;

  int other, iz;
  if (tile < NMAIN) {
    other = tile / NTZ;
    iz = (tile % NTZ) * T + p;
  } else {
    other = (tile - NMAIN) * T + p;
    iz = N / 2;
  }
  // element (r along AXIS, other, iz)
  const size_t stride = (AXIS == 0) ? (size_t)N * NZH : (size_t)NZH;
  const size_t base = ((AXIS == 0) ? (size_t)other * NZH : (size_t)other * N * NZH) + iz;

  // all eight loads are issued before anything consumes them (one round trip, not eight)
  double2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = in[base + (size_t)(t + m * (N / 8)) * stride];
  if (lop.kind == K_MULREAL || lop.kind == K_FINAL) {
    double f[8];
    // the real multiplier array has row pitch N/2+2 (16-byte rows, see RealPitch)
    const size_t rstride = (AXIS == 0) ? (size_t)N * (NZH + 1) : (size_t)(NZH + 1);
    const size_t rbase = ((AXIS == 0) ? (size_t)other * (NZH + 1) : (size_t)other * N * (NZH + 1)) + iz;
#pragma unroll
    for (int m = 0; m < 8; ++m) f[m] = __ldg(lop.real0 + rbase + (size_t)(t + m * (N / 8)) * rstride);
    if (lop.kind == K_FINAL) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double2 hh[4];
#pragma unroll
        for (int m = 0; m < 4; ++m)
          hh[m] = __ldg(lop.cplx0 + base + (size_t)(t + (4 * half + m) * (N / 8)) * stride);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int mm = 4 * half + m;
          v[mm] = make_double2(v[mm].x * f[mm] + lop.a * hh[m].x, v[mm].y * f[mm] + lop.a * hh[m].y);
        }
      }
    } else {
#pragma unroll
      for (int m = 0; m < 8; ++m) v[m] = make_double2(v[m].x * f[m], v[m].y * f[m]);
    }
  } else if (lop.kind != K_NONE) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int r = t + m * (N / 8);
      const int ix = (AXIS == 0) ? r : other;
      const int iy = (AXIS == 0) ? other : r;
      v[m] = kop_load<N>(lop, v[m], 0, ix, iy, iz);
    }
  }

  SmStrided sm{smem, T, p};
  fft_stages<N, 1, DIR, SmStrided>(v, t, tw, sm);

#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int r = t + m * (N / 8);
    const size_t off = base + (size_t)r * stride;
    if (sop.kind != K_NONE) {
      const int ix = (AXIS == 0) ? r : other;
      const int iy = (AXIS == 0) ? other : r;
      kop_store<N>(sop, out, v[m], off, ix, iy, iz);
    } else {
      out[off] = v[m];
    }
  }
}

// ---------------------------------------------------------------------------
// persistent, software-pipelined variant of the strided pass (used for N >= 128).
//
// The plain kernel above serialises load -> stages -> store inside a CTA, and
// with the register file limiting it to ~32 warps per SM the memory pipe idles
// while warps compute (ncu, round 1: long_scoreboard 63 % of stall samples, 36 %
// of DRAM peak).  Here a CTA walks tiles blockIdx.x, +gridDim.x, ... and, before
// it transforms tile i, every thread issues cp.async (LDGSTS) copies of the
// eight elements IT will own in tile i+1 into a second shared-memory buffer.
// The slots are thread-private, so no barrier guards them: a thread waits on
// its own copy group, lifts its eight values into registers and immediately
// re-arms the buffer with the tile after.  HBM reads therefore overlap the
// whole butterfly / exchange / store phase of the previous tile.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

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

template <int N, int T, int DIR, int AXIS>
__global__ void __launch_bounds__(T *N / 8, (T * N / 8 <= 256) ? 3 : 1)
    fft_strided_pass_pipelined(const double2 *__restrict__ in, double2 *__restrict__ out,
                               const double2 *__restrict__ tw, KOp lop, KOp sop) {
  extern __shared__ double2 smem[];
  constexpr int NZH = N / 2 + 1;
  constexpr int NTILES = N * ((N / 2) / T) + N / T;
  double2 *xch = smem;           // stage exchanges of the current tile
  double2 *pre = smem + N * T;   // thread-private landing slots of the next tile
  const int p = threadIdx.x % T;
  const int t = threadIdx.x / T;
  constexpr size_t stride = (AXIS == 0) ? (size_t)N * NZH : (size_t)NZH;

  int tile = blockIdx.x;
  int other, iz;
  size_t base;
  if (tile < NTILES) {
    strided_tile_coords<N, T, AXIS>(tile, p, other, iz, base);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int r = t + m * (N / 8);
      cp_async16(pre + r * T + p, in + base + (size_t)r * stride);
    }
  }
  cp_async_commit();

  while (tile < NTILES) {
    cp_async_wait_all();
    double2 v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = pre[(t + m * (N / 8)) * T + p];

    // re-arm the landing slots with the tile after this one
    const int next = tile + gridDim.x;
    if (next < NTILES) {
      int o2, z2;
      size_t b2;
      strided_tile_coords<N, T, AXIS>(next, p, o2, z2, b2);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = t + m * (N / 8);
        cp_async16(pre + r * T + p, in + b2 + (size_t)r * stride);
      }
    }
    cp_async_commit();

    if (lop.kind == K_MULREAL || lop.kind == K_FINAL) {
      double f[8];
      const size_t rstride = (AXIS == 0) ? (size_t)N * (NZH + 1) : (size_t)(NZH + 1);
      const size_t rbase = ((AXIS == 0) ? (size_t)other * (NZH + 1) : (size_t)other * N * (NZH + 1)) + iz;
#pragma unroll
      for (int m = 0; m < 8; ++m) f[m] = __ldg(lop.real0 + rbase + (size_t)(t + m * (N / 8)) * rstride);
      if (lop.kind == K_FINAL) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          double2 hh[4];
#pragma unroll
          for (int m = 0; m < 4; ++m)
            hh[m] = __ldg(lop.cplx0 + base + (size_t)(t + (4 * half + m) * (N / 8)) * stride);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int mm = 4 * half + m;
            v[mm] = make_double2(v[mm].x * f[mm] + lop.a * hh[m].x, v[mm].y * f[mm] + lop.a * hh[m].y);
          }
        }
      } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = make_double2(v[m].x * f[m], v[m].y * f[m]);
      }
    } else if (lop.kind != K_NONE) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = t + m * (N / 8);
        const int ix = (AXIS == 0) ? r : other;
        const int iy = (AXIS == 0) ? other : r;
        v[m] = kop_load<N>(lop, v[m], 0, ix, iy, iz);
      }
    }

    SmStrided sm{xch, T, p};
    fft_stages<N, 1, DIR, SmStrided>(v, t, tw, sm);

#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int r = t + m * (N / 8);
      const size_t off = base + (size_t)r * stride;
      if (sop.kind != K_NONE) {
        const int ix = (AXIS == 0) ? r : other;
        const int iy = (AXIS == 0) ? other : r;
        kop_store<N>(sop, out, v[m], off, ix, iy, iz);
      } else {
        out[off] = v[m];
      }
    }
    tile = next;
    if (tile < NTILES) strided_tile_coords<N, T, AXIS>(tile, p, other, iz, base);
  }
}

// ---------------------------------------------------------------------------
// z pass, real -> half-complex.  N reals are transformed as M = N/2 complex
// z[j] = x[2j] + i x[2j+1]; then X[k] = E[k] + w_N^k O[k] with
// E = (Z[k] + conj Z[M-k])/2, O = (Z[k] - conj Z[M-k])/(2i).
// A CTA owns TR consecutive rows; threads: TR * M/8, thread = t + (M/8)*row.
// twN is the length-N table (post-processing), twM the length-M one (stages).
// ---------------------------------------------------------------------------
template <int N, int TR>
__global__ void __launch_bounds__(TR *N / 16)
    fft_r2c_zpass(const double *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ twN,
                  const double2 *__restrict__ twM, ROp lop, size_t nrows) {
  extern __shared__ double2 smem[];
  constexpr int M = N / 2;
  constexpr int ROWP = M + M / 8 + 1;  // padded row length
  const int t = threadIdx.x % (M / 8);
  const int rl = threadIdx.x / (M / 8);
  const size_t row = (size_t)blockIdx.x * TR + rl;  // host guarantees nrows % TR == 0
  const double2 *src = reinterpret_cast<const double2 *>(in + row * N);
  const double2 *aux = lop.aux ? reinterpret_cast<const double2 *>(lop.aux + row * N) : nullptr;

  double2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    double2 x = src[t + m * (M / 8)];
    if (lop.kind == R_LOAD_SCALE) {
      x.x *= lop.a;
      x.y *= lop.a;
    } else if (lop.kind == R_SCALE_MUL) {
      const double2 y = aux[t + m * (M / 8)];
      x.x *= lop.a * y.x;
      x.y *= lop.a * y.y;
    }
    v[m] = x;
  }
  SmRow sm{smem + (size_t)rl * ROWP};
  fft_stages<M, 1, -1, SmRow>(v, t, twM, sm);
#pragma unroll
  for (int m = 0; m < 8; ++m) sm.store(t + m * (M / 8), v[m]);
  __syncthreads();

  double2 *dst = out + row * (M + 1);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = t + m * (M / 8);
    const double2 zk = v[m];
    const double2 zm = sm.load((M - k) & (M - 1));
    // E = (zk + conj zm)/2 ; O = (zk - conj zm)/(2i) = (-i/2)(zk - conj zm)
    const double2 e = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
    const double2 o = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
    const double2 w = __ldg(twN + k);
    dst[k] = cadd(e, cmul(w, o));
    if (k == 0) dst[M] = make_double2(e.x - o.x, 0.0);  // w_N^M = -1, E[0], O[0] real
  }
}

// ---------------------------------------------------------------------------
// z pass, half-complex -> real (unnormalised FFTW c2r semantics times the
// store functor's scale): Z[k] = E' + i O', E' = X[k] + conj X[M-k],
// O' = (X[k] - conj X[M-k]) conj(w_N^k); inverse M-point FFT; x[2j] = Re z[j],
// x[2j+1] = Im z[j].
// ---------------------------------------------------------------------------
template <int N, int TR>
__global__ void __launch_bounds__(TR *N / 16)
    fft_c2r_zpass(const double2 *__restrict__ in, double *__restrict__ out, const double2 *__restrict__ twN,
                  const double2 *__restrict__ twM, ROp sop, size_t nrows) {
  extern __shared__ double2 smem[];
  if (sop.skip && *sop.skip) return;
  constexpr int M = N / 2;
  constexpr int ROWP = M + M / 8 + 1;
  const int t = threadIdx.x % (M / 8);
  const int rl = threadIdx.x / (M / 8);
  const size_t row = (size_t)blockIdx.x * TR + rl;  // host guarantees nrows % TR == 0
  const double2 *src = in + row * (M + 1);
  SmRow sm{smem + (size_t)rl * ROWP};

  // stage the M+1 inputs in shared memory (coalesced), then pair k with M-k
#pragma unroll
  for (int m = 0; m < 8; ++m) sm.store(t + m * (M / 8), src[t + m * (M / 8)]);
  if (t == 0) sm.store(M, src[M]);
  __syncthreads();

  double2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = t + m * (M / 8);
    double2 xk = sm.load(k);
    double2 xm = sm.load(M - k);
    if (k == 0) {  // FFTW's c2r ignores the imaginary parts of the self-conjugate bins
      xk.y = 0.0;
      xm.y = 0.0;
    }
    const double2 e = make_double2(xk.x + xm.x, xk.y - xm.y);
    const double2 d = make_double2(xk.x - xm.x, xk.y + xm.y);
    const double2 w = cconj(__ldg(twN + k));
    const double2 o = cmul(d, w);
    // Z = E' + i O'
    v[m] = make_double2(e.x - o.y, e.y + o.x);
  }
  __syncthreads();
  fft_stages<M, 1, +1, SmRow>(v, t, twM, sm);

  double2 *dst = reinterpret_cast<double2 *>(out + row * N);
  const double2 *aux = sop.aux ? reinterpret_cast<const double2 *>(sop.aux + row * N) : nullptr;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int j = t + m * (M / 8);
    double2 x = make_double2(sop.a * v[m].x, sop.a * v[m].y);
    if (sop.kind == R_SCALE_MUL) {
      const double2 y = aux[j];
      x.x *= y.x;
      x.y *= y.y;
    } else if (sop.kind == R_AXPY) {
      const double2 o = dst[j];
      x.x += o.x;
      x.y += o.y;
    }
    dst[j] = x;
  }
}

}  // namespace bgpu
