// barcode_b200/csrc/fft_tma.cuh
//
// TMA-staged strided FFT pass for sm_100a (the y and x pencil passes of the 3-D
// transform behind fftR2C / fftC2R, /root/reference/barlib/src/fftwrapper.cc:
// 26-125, with the k-space loops of EqSolvers.cc:208-268, gradient.cpp:38-74,
// 167-210 and HMC_help.cc:41-58 fused in as functors).
//
// Data movement is done by the Tensor Memory Accelerator, not by the threads:
//   * a CTA walks tiles of T = 8 pencils adjacent in z: N rows (the transformed
//     axis) x 128 contiguous bytes.  One thread issues
//     cp.async.bulk.tensor.3d (UTMALDG) for the tile two iterations ahead into
//     a three-stage shared-memory ring; completion is signalled on an mbarrier.
//     N/2+1 is odd, so the last z tile hangs over the edge of the array: TMA
//     zero-fills the out-of-bounds columns on load and clips them on store.
//   * the tile lands with the hardware 128-byte swizzle (16-byte chunk index
//     XOR row mod 8), which makes "lane t reads row t of pencil p" free of bank
//     conflicts without padding.
//   * every pencil is transformed by the lanes of ONE warp (N/E lanes, E
//     elements each), in place in its own column of the tile, so the Stockham
//     exchanges between radix-8 stages need __syncwarp() only -- one
//     __syncthreads() per tile remains, ahead of the TMA store.  The exchange
//     after the first stage permutes rows (row ^= (row >> 3) & 7) so that its
//     stride-8 scatter is conflict-free as well.
//   * twiddles are tile-invariant per thread and live in registers.
//   * results go back with cp.async.bulk.tensor (UTMASTG); the accumulating
//     back-projection (K_INVLAP_ADD) uses the TMA's f64 reduce-add instead of a
//     read-modify-write.  Real / complex multiplier arrays (K_MULREAL, K_FINAL)
//     ride along as extra TMA tiles on the same mbarrier.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft.cuh"

namespace bgpu {

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, int c0, int c1, int c2,
                                                  const void *smem_src) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int K>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(K) : "memory");
}
template <int K>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(K) : "memory");
}

// ---------------------------------------------------------------------------
// warp-private Stockham stages.  Thread t (of LP = N/E lanes) holds elements
// t + m*LP in v[m].  A stage of radix R runs NB = E/R butterflies per thread;
// butterfly j owns registers j + k*NB (k < R) = elements b + k*N/R, b = t + j*LP.
// Results are exchanged through the pencil's own column of the tile.
// ---------------------------------------------------------------------------
template <int N>
struct StageCount {
  static constexpr int value = (N <= 8) ? 1 : 1 + StageCount<(N + 7) / 8>::value;
};
template <>
struct StageCount<1> {
  static constexpr int value = 0;
};

// shared-memory address of (row, pencil p) in a 128-byte-swizzled tile
__device__ __forceinline__ uint32_t tile_addr(uint32_t tile, int row, int p) {
  return tile + (uint32_t)row * 128u + (uint32_t)((p ^ (row & 7)) << 4);
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

template <int R, int DIR>
__device__ __forceinline__ void butterfly(double2 &a0, double2 &a1, double2 &a2, double2 &a3, double2 &a4, double2 &a5,
                                          double2 &a6, double2 &a7) {
  if constexpr (R == 8) {
    double2 w[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
    bf8<DIR>(w);
    a0 = w[0]; a1 = w[1]; a2 = w[2]; a3 = w[3]; a4 = w[4]; a5 = w[5]; a6 = w[6]; a7 = w[7];
  } else if constexpr (R == 4) {
    bf4<DIR>(a0, a1, a2, a3);
  } else {
    bf2<DIR>(a0, a1);
  }
}

// load the per-thread twiddles of every non-final stage: twr[stage][m] multiplies register m
template <int N, int E, int S, int DIR, int STG>
__device__ __forceinline__ void wp_load_twiddles(double2 (*twr)[E], int t, const double2 *__restrict__ tw) {
  constexpr int LP = N / E;
  constexpr int R = StageRadix<N, S>::value;
  constexpr int NB = E / R;
  if constexpr (S * R < N) {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int b = t + j * LP;
      const int base = b - (b & (S - 1));
#pragma unroll
      for (int k = 0; k < R; ++k) twr[STG][j + k * NB] = twiddle<DIR>(tw, base * k);
    }
    wp_load_twiddles<N, E, S * R, DIR, STG + 1>(twr, t, tw);
  }
}

template <int N, int E, int S, int DIR, int STG>
__device__ __forceinline__ void wp_stages(double2 (&v)[E], int t, int p, uint32_t tile, const double2 (*twr)[E]) {
  constexpr int LP = N / E;
  constexpr int R = StageRadix<N, S>::value;
  constexpr int NB = E / R;
  constexpr bool last = (S * R == N);
  static_assert(E % R == 0, "elements per thread must be a multiple of the radix");
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if constexpr (R == 8)
      butterfly<8, DIR>(v[j], v[j + NB], v[j + 2 * NB], v[j + 3 * NB], v[j + 4 * NB], v[j + 5 * NB], v[j + 6 * NB],
                        v[j + 7 * NB]);
    else if constexpr (R == 4)
      bf4<DIR>(v[j], v[j + NB], v[j + 2 * NB], v[j + 3 * NB]);
    else
      bf2<DIR>(v[j], v[j + NB]);
  }
  if constexpr (!last) {
    __syncwarp();  // every lane has lifted its inputs out of the column
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int b = t + j * LP;
      const int q = b & (S - 1);
      const int base = b - q;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        double2 x = v[j + k * NB];
        if (k > 0) x = cmul(x, twr[STG][j + k * NB]);
        int e = q + R * base + k * S;
        if constexpr (S == 1) e ^= (e >> 3) & 7;  // conflict-free stride-8 scatter
        sts128(tile_addr(tile, e, p), x);
      }
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < E; ++m) {
      int e = t + m * LP;
      if constexpr (S == 1) e ^= (e >> 3) & 7;
      v[m] = lds128(tile_addr(tile, e, p));
    }
    wp_stages<N, E, S * R, DIR, STG + 1>(v, t, p, tile, twr);
  }
}

// ---------------------------------------------------------------------------
// the kernel.  AUX: 0 none, 1 real multiplier tile (K_MULREAL), 2 real + complex (K_FINAL)
// ---------------------------------------------------------------------------
template <int N, int AUX>
struct TmaTile {
  static constexpr int T = 8;
  static constexpr int main_bytes = N * T * 16;
  static constexpr int auxr_bytes = AUX >= 1 ? N * T * 8 : 0;
  static constexpr int auxc_bytes = AUX >= 2 ? N * T * 16 : 0;
  static constexpr int stage_bytes = main_bytes + auxr_bytes + auxc_bytes;
};

struct TmaMaps {
  CUtensorMap in, out, auxr, auxc;
};

// shared-memory address of (row, pencil p) in a 64-byte-swizzled tile of doubles (T = 8 per row)
__device__ __forceinline__ uint32_t tile_addr_r(uint32_t tile, int row, int p) {
  return tile + (uint32_t)row * 64u + (uint32_t)((((p >> 1) ^ ((row >> 1) & 3)) << 4) | ((p & 1) << 3));
}

// ---------------------------------------------------------------------------
// k-space functors of the strided pass.  K_DISP, K_GRAD and K_INVLAP_* all have the form
// out = f * (Im v, -Re v) with a real f(k); everything that depends only on the tile (the two
// wave numbers that are constant along the pencil, the Nyquist mask) is hoisted out of the
// per-element work.
//   K_DISP    f = a k_c / k^2, 0 if k^2 <= 1e-14 or on a Nyquist plane   (EqSolvers.cc:208-268)
//   K_GRAD    f = -k_c,        0 on a Nyquist plane                       (gradient.cpp:38-74)
//   K_INVLAP  f = k_c / k^2,   0 if k^2 == 0 or on a Nyquist plane        (gradient.cpp:167-210)
// ---------------------------------------------------------------------------
template <int N, int AXIS>
struct RotCtx {
  double k_oth, k_z, c2, kfac, a;
  int sel;       // which wave number is k_c: 0 = along the pencil, 1 = the other strided axis, 2 = z
  int kind;
  bool masked;   // the whole tile column sits on a Nyquist plane
  __device__ __forceinline__ void setup(const KOp &op, int other, int iz) {
    kind = op.kind;
    kfac = op.kfac;
    a = op.a;
    k_oth = kval(other, N, op.kfac);
    k_z = kval(iz, N, op.kfac);
    c2 = k_oth * k_oth + k_z * k_z;
    masked = (other == N / 2) || (iz == N / 2);
    // comp: 0 = x, 1 = y, 2 = z ; the pencil runs along x for AXIS == 0, along y for AXIS == 1
    sel = (op.comp == 2) ? 2 : ((op.comp == (AXIS == 0 ? 0 : 1)) ? 0 : 1);
  }
  __device__ __forceinline__ double2 apply(double2 v, int r) const {
    const double kr = kval(r, N, kfac);
    const double kc = sel == 0 ? kr : (sel == 1 ? k_oth : k_z);
    double f;
    if (kind == K_GRAD) {
      f = -kc;
    } else {
      const double ksq = kr * kr + c2;
      f = kc * __drcp_rn(ksq);
      if (kind == K_DISP) {
        f *= a;
        if (!(ksq > 1.e-14)) f = 0.0;
      } else if (!(ksq > 0.0)) {
        f = 0.0;
      }
    }
    if (masked || r == N / 2) f = 0.0;
    return make_double2(f * v.y, -(f * v.x));
  }
};

template <int N, int E, int NSTAGE, int DIR, int AXIS, int AUX, int MINB>
__global__ void __launch_bounds__(8 * (N / E), MINB)
    fft_strided_tma(const __grid_constant__ TmaMaps maps, const double2 *__restrict__ tw, KOp lop, KOp sop) {
  constexpr int T = 8;
  constexpr int LP = N / E;
  constexpr int NZH = N / 2 + 1;
  constexpr int ZT = (NZH + T - 1) / T;
  constexpr int NTILES = N * ZT;
  constexpr int NSTG = StageCount<N>::value;
  constexpr int ROWS_PER_BOX = N > 256 ? 256 : N;
  using Tile = TmaTile<N, AUX>;
  static_assert(LP <= 32 && LP >= 8, "a pencil must live inside one warp");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_al = smem_raw + (smem0 - smem_u32(smem_raw));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_al + NSTAGE * Tile::stage_bytes);

  const int tid = threadIdx.x;
  const int p = tid / LP;
  const int t = tid % LP;

  const int my_count = (NTILES - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto issue_load = [&](int i) {
    const int tile = blockIdx.x + i * gridDim.x;
    const int s = i % NSTAGE;
    const int other = tile / ZT, zt = tile % ZT;
    uint8_t *dst = smem_al + s * Tile::stage_bytes;
    mbar_expect_tx(&full[s], Tile::stage_bytes);
#pragma unroll
    for (int r0 = 0; r0 < N; r0 += ROWS_PER_BOX) {
      const int c1 = AXIS == 0 ? other : r0, c2 = AXIS == 0 ? r0 : other;
      tma_load_3d(dst + r0 * 128, &maps.in, zt * 2 * T, c1, c2, &full[s]);
      if constexpr (AUX >= 1) tma_load_3d(dst + Tile::main_bytes + r0 * 64, &maps.auxr, zt * T, c1, c2, &full[s]);
      if constexpr (AUX >= 2)
        tma_load_3d(dst + Tile::main_bytes + Tile::auxr_bytes + r0 * 128, &maps.auxc, zt * 2 * T, c1, c2, &full[s]);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NSTAGE - 1; ++i)
      if (i < my_count) issue_load(i);
  }

  // tile-invariant twiddles
  double2 twr[NSTG > 1 ? NSTG - 1 : 1][E];
  wp_load_twiddles<N, E, 1, DIR, 0>(twr, t, tw);

  for (int i = 0; i < my_count; ++i) {
    const int s = i % NSTAGE;
    const int tile = blockIdx.x + i * gridDim.x;
    const int other = tile / ZT, zt = tile % ZT;
    const int iz = zt * T + p;
    const uint32_t tbase = smem0 + s * Tile::stage_bytes;
    mbar_wait(&full[s], (i / NSTAGE) & 1);

    double2 v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = lds128(tile_addr(tbase, t + m * LP, p));

    if constexpr (AUX >= 1) {
      // K_MULREAL: v * real0 ; K_FINAL: v * real0 + a * cplx0
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const int r = t + m * LP;
        double f;
        asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(f) : "r"(tile_addr_r(tbase + Tile::main_bytes, r, p)));
        if constexpr (AUX >= 2) {
          const double2 h = lds128(tile_addr(tbase + Tile::main_bytes + Tile::auxr_bytes, r, p));
          v[m] = make_double2(v[m].x * f + lop.a * h.x, v[m].y * f + lop.a * h.y);
        } else {
          v[m] = make_double2(v[m].x * f, v[m].y * f);
        }
      }
    } else if (lop.kind != K_NONE) {
      RotCtx<N, AXIS> rc;
      rc.setup(lop, other, iz);
#pragma unroll
      for (int m = 0; m < E; ++m) v[m] = rc.apply(v[m], t + m * LP);
    }

    wp_stages<N, E, 1, DIR, 0>(v, t, p, tbase, twr);

    if (sop.kind == K_INVLAP_SET || sop.kind == K_INVLAP_ADD) {
      RotCtx<N, AXIS> rc;
      rc.setup(sop, other, iz);
#pragma unroll
      for (int m = 0; m < E; ++m) v[m] = rc.apply(v[m], t + m * LP);
    }
    __syncwarp();  // the last exchange's reads are done
#pragma unroll
    for (int m = 0; m < E; ++m) sts128(tile_addr(tbase, t + m * LP, p), v[m]);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int r0 = 0; r0 < N; r0 += ROWS_PER_BOX) {
        const int c1 = AXIS == 0 ? other : r0, c2 = AXIS == 0 ? r0 : other;
        const void *src = smem_al + s * Tile::stage_bytes + r0 * 128;
        if (sop.kind == K_INVLAP_ADD)
          tma_reduce_add_3d(&maps.out, zt * 2 * T, c1, c2, src);
        else
          tma_store_3d(&maps.out, zt * 2 * T, c1, c2, src);
      }
      bulk_commit();
      const int j = i + NSTAGE - 1;
      if (j < my_count) {
        bulk_wait_read<1>();  // the store issued one iteration ago has drained its stage
        issue_load(j);
      }
    }
  }
  if (tid == 0) bulk_wait<0>();
}

}  // namespace bgpu
