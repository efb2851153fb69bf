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

#include <type_traits>

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
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                             const void *smem_src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// Programmatic dependent launch: a pass launched with the programmatic-stream-serialization attribute may become
// resident while the previous kernel of the stream drains; everything up to pdl_wait() (barrier set-up, twiddle
// loads -- nothing the previous kernel writes) overlaps that tail, everything after it sees the previous kernel's
// results.  pdl_trigger() lets the NEXT kernel of the stream do the same with this one.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

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

// the lanes of one pencil meet: a warp, or -- a pencil spread over two warps -- named barrier 1 + p
// (barrier 0 is __syncthreads'; a CTA has 16)
template <int LANES>
__device__ __forceinline__ void pencil_sync(int p) {
  if constexpr (LANES <= 32) {
    __syncwarp();
  } else {
    static_assert(LANES % 32 == 0, "a named barrier counts whole warps");
    asm volatile("bar.sync %0, %1;\n" ::"r"(p + 1), "n"(LANES) : "memory");
  }
}

// where a pencil's elements live in shared memory, and how its lanes synchronise
struct ColAccess {  // column p of a 128-byte-swizzled tile (strided pass), pencil inside one warp
  uint32_t tile;
  int p;
  __device__ __forceinline__ uint32_t at(int e) const { return tile_addr(tile, e, p); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};
template <int LANES>
struct ColAccessWide {  // the same column, pencil spread over LANES / 32 warps
  uint32_t tile;
  int p;
  __device__ __forceinline__ uint32_t at(int e) const { return tile_addr(tile, e, p); }
  __device__ __forceinline__ void sync() const { pencil_sync<LANES>(p); }
};
struct RowAccess {  // a contiguous row of double2 (z pass)
  uint32_t row;
  __device__ __forceinline__ uint32_t at(int e) const { return row + (uint32_t)e * 16u; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

// CONJ: the register twiddles were loaded for the opposite direction (a kernel that transforms both ways keeps one set)
template <int N, int E, int S, int DIR, int STG, class Acc, bool CONJ = false>
__device__ __forceinline__ void wp_stages(double2 (&v)[E], int t, const Acc &acc, const double2 (*twr)[E]) {
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
    acc.sync();  // every lane has lifted its inputs out of the pencil's storage
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int b = t + j * LP;
      const int q = b & (S - 1);
      const int base = b - q;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        double2 x = v[j + k * NB];
        if (k > 0) x = cmul(x, CONJ ? cconj(twr[STG][j + k * NB]) : twr[STG][j + k * NB]);
        int e = q + R * base + k * S;
        if constexpr (S == 1) e ^= (e >> 3) & 7;  // conflict-free stride-8 scatter
        sts128(acc.at(e), x);
      }
    }
    acc.sync();
#pragma unroll
    for (int m = 0; m < E; ++m) {
      int e = t + m * LP;
      if constexpr (S == 1) e ^= (e >> 3) & 7;
      v[m] = lds128(acc.at(e));
    }
    wp_stages<N, E, S * R, DIR, STG + 1, Acc, CONJ>(v, t, acc, twr);
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

constexpr int kMaxPeers = 8;  // GPUs of one box

struct TmaMaps {
  CUtensorMap in, out, auxr, auxc;
  CUtensorMap peer[kMaxPeers];  // receive buffers of the slab ranks (peer-mapped over NVLink), rank-4
};

// Shape of one strided pass.  A full cube has n_other = N pencils' worth of "other" index
// (x planes for the y pass, y rows for the x pass) and no packing.  In the slab-decomposed
// transform (fft_slab.cu) a rank holds n_other = N/G of them, `other0` is the global index of the
// first (the k-space functors need global wave numbers), and the y pass reads / writes the
// all-to-all buffers directly: `*_packed` = rows per peer block of the rank-4 tensor map
// [z][y_local][x_local][peer], 0 for the plain rank-3 map.
struct PassGeom {
  int n_other;
  int other0;
  int in_packed;
  int out_packed;
  // fused transpose: > 0 = rows per peer (Ns); the tile's rows [h*Ns, (h+1)*Ns) are stored straight into
  // block `my_rank` of rank h's receive buffer [src][x_l][y_l][z] over NVLink (maps.peer[h]), for both
  // pass axes -- the all-to-all of the distributed transform happens inside the pass kernel
  int out_peers;
  int my_rank;
  // walk the tiles from the last to the first: consecutive passes that alternate direction start on the part of
  // their input the previous pass wrote last, which is still in L2 (Fft3d::pingpong)
  int reverse;
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
// K_NEGINVK2 (PoissonSolver, EqSolvers.cc:29-64) is the scalar -a / k^2 (0 at k = 0) on both parts.
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
    if (kind == K_NEGINVK2) {  // a scalar multiplier, not a rotation: PoissonSolver's -1/k^2
      const double ksq = kr * kr + c2;
      const double g = ksq > 0.0 ? -a * __drcp_rn(ksq) : 0.0;
      return make_double2(g * v.x, g * v.y);
    }
    if (kind == K_GAUSS || kind == K_ONE_MINUS_GAUSS) {  // the ALPT smoothing kernel and its complement
      const double K = exp(-(kr * kr + c2) * (a * a) / 2.);
      const double g = kind == K_GAUSS ? K : 1.0 - K;
      return make_double2(g * v.x, g * v.y);
    }
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

// RotCtx plus the kinds of the shared x pass (fft_ops.h): comp = K_COMP_UNIT and the real k_c multipliers.
// A separate struct so that the kernels of the default path compile to the same code as before.
template <int N, int AXIS>
struct RotCtxX {
  double k_oth, k_z, c2, kfac, a;
  int sel;       // 0 = k along the pencil, 1 = the other strided axis, 2 = z, 3 = unit
  int kind;
  bool masked;
  __device__ __forceinline__ void setup(const KOp &op, int other, int iz) {
    kind = op.kind;
    kfac = op.kfac;
    a = op.a;
    k_oth = kval(other, N, op.kfac);
    k_z = kval(iz, N, op.kfac);
    c2 = k_oth * k_oth + k_z * k_z;
    masked = (other == N / 2) || (iz == N / 2);
    sel = (op.comp == K_COMP_UNIT) ? 3 : ((op.comp == 2) ? 2 : ((op.comp == (AXIS == 0 ? 0 : 1)) ? 0 : 1));
  }
  __device__ __forceinline__ double2 apply(double2 v, int r) const {
    const double kr = kval(r, N, kfac);
    const double kc = sel == 3 ? 1.0 : (sel == 0 ? kr : (sel == 1 ? k_oth : k_z));
    if (kind == K_MULK || kind == K_MULK_SET || kind == K_MULK_ADD) return make_double2(kc * v.x, kc * v.y);
    double f;
    if (kind == K_GRAD) {
      f = -kc;
    } else {  // K_DISP, K_INVLAP_SET, K_INVLAP_ADD
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

// AUX: 0 none, 1 real multiplier tile (K_MULREAL), 2 real + complex (K_FINAL), -1 none + the functors of RotCtxX
template <int N, int E, int NSTAGE, int DIR, int AXIS, int AUX, int MINB>
__global__ void __launch_bounds__(8 * (N / E), MINB)
    fft_strided_tma(const __grid_constant__ TmaMaps maps, const double2 *__restrict__ tw, KOp lop, KOp sop,
                    PassGeom geo) {
  constexpr int T = 8;
  constexpr int LP = N / E;
  constexpr int NZH = N / 2 + 1;
  constexpr int ZT = (NZH + T - 1) / T;
  const int NTILES = geo.n_other * ZT;
  constexpr int NSTG = StageCount<N>::value;
  constexpr int ROWS_PER_BOX = N > 256 ? 256 : N;
  using Tile = TmaTile<N, AUX>;
  static_assert(LP >= 8 && (LP <= 32 || LP == 64), "a pencil lives inside one warp, or in two (named barrier per pencil)");
  using Col = typename std::conditional<(LP <= 32), ColAccess, ColAccessWide<LP>>::type;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_al = smem_raw + (smem0 - smem_u32(smem_raw));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_al + NSTAGE * Tile::stage_bytes);

  const int tid = threadIdx.x;
  const int p = tid / LP;
  const int t = tid % LP;

  const int my_count = (NTILES - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto tile_of = [&](int i) {
    const int tl = blockIdx.x + i * gridDim.x;
    return geo.reverse ? NTILES - 1 - tl : tl;
  };
  auto issue_load = [&](int i) {
    const int tile = tile_of(i);
    const int s = i % NSTAGE;
    const int other = tile / ZT, zt = tile % ZT;
    uint8_t *dst = smem_al + s * Tile::stage_bytes;
    mbar_expect_tx(&full[s], Tile::stage_bytes);
    if (AXIS == 1 && geo.in_packed) {  // rows arrive grouped by the peer that sent them
      const int rb = geo.in_packed < 256 ? geo.in_packed : 256;
      for (int r0 = 0; r0 < N; r0 += rb)
        tma_load_4d(dst + r0 * 128, &maps.in, zt * 2 * T, r0 % geo.in_packed, other, r0 / geo.in_packed, &full[s]);
      return;
    }
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
  // With three stages the tile two iterations ahead is requested after this iteration's store
  // (its stage drained an iteration ago).  With two stages (the variants whose operand tiles eat
  // the shared memory) the next tile is requested at the TOP of the iteration instead, as soon
  // as the store issued a moment ago has finished reading its stage, so the load still overlaps
  // the whole transform.
  constexpr bool EARLY = (NSTAGE == 2);
  // tile-invariant twiddles (a constant table: safe ahead of pdl_wait)
  double2 twr[NSTG > 1 ? NSTG - 1 : 1][E];
  wp_load_twiddles<N, E, 1, DIR, 0>(twr, t, tw);
  pdl_trigger();
  pdl_wait();
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < (EARLY ? 1 : NSTAGE - 1); ++i)
      if (i < my_count) issue_load(i);
  }

  for (int i = 0; i < my_count; ++i) {
    const int s = i % NSTAGE;
    const int tile = tile_of(i);
    const int other = tile / ZT, zt = tile % ZT;
    const int iz = zt * T + p;
    const uint32_t tbase = smem0 + s * Tile::stage_bytes;
    if constexpr (EARLY) {
      if (tid == 0 && i + 1 < my_count) {
        bulk_wait_read<0>();
        issue_load(i + 1);
      }
    }
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
      typename std::conditional<AUX == -1, RotCtxX<N, AXIS>, RotCtx<N, AXIS>>::type rc;
      rc.setup(lop, other + geo.other0, iz);
#pragma unroll
      for (int m = 0; m < E; ++m) v[m] = rc.apply(v[m], t + m * LP);
    }

    wp_stages<N, E, 1, DIR, 0>(v, t, Col{tbase, p}, twr);

    if (AUX == -1 ? sop.kind != K_NONE : (sop.kind == K_INVLAP_SET || sop.kind == K_INVLAP_ADD)) {
      typename std::conditional<AUX == -1, RotCtxX<N, AXIS>, RotCtx<N, AXIS>>::type rc;
      rc.setup(sop, other + geo.other0, iz);
#pragma unroll
      for (int m = 0; m < E; ++m) v[m] = rc.apply(v[m], t + m * LP);
    }
    pencil_sync<LP>(p);  // the last exchange's reads are done
#pragma unroll
    for (int m = 0; m < E; ++m) sts128(tile_addr(tbase, t + m * LP, p), v[m]);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      if (geo.out_peers) {
        const int ns = geo.out_peers, rb = ns < 256 ? ns : 256;
        for (int r0 = 0; r0 < N; r0 += rb) {
          const int h = r0 / ns, rl = r0 % ns;
          // [z][y_l][x_l][src]: the y pass (AXIS 1) scatters its y rows, the x pass its x rows
          tma_store_4d(&maps.peer[h], zt * 2 * T, AXIS == 1 ? rl : other, AXIS == 1 ? other : rl, geo.my_rank,
                       smem_al + s * Tile::stage_bytes + r0 * 128);
        }
      } else if (AXIS == 1 && geo.out_packed) {  // rows leave grouped by the peer they go to
        const int rb = geo.out_packed < 256 ? geo.out_packed : 256;
        for (int r0 = 0; r0 < N; r0 += rb)
          tma_store_4d(&maps.out, zt * 2 * T, r0 % geo.out_packed, other, r0 / geo.out_packed,
                       smem_al + s * Tile::stage_bytes + r0 * 128);
      } else
#pragma unroll
      for (int r0 = 0; r0 < N; r0 += ROWS_PER_BOX) {
        const int c1 = AXIS == 0 ? other : r0, c2 = AXIS == 0 ? r0 : other;
        const void *src = smem_al + s * Tile::stage_bytes + r0 * 128;
        if (sop.kind == K_INVLAP_ADD || (AUX == -1 && sop.kind == K_MULK_ADD))
          tma_reduce_add_3d(&maps.out, zt * 2 * T, c1, c2, src);
        else
          tma_store_3d(&maps.out, zt * 2 * T, c1, c2, src);
      }
      bulk_commit();
      if constexpr (!EARLY) {
        const int j = i + NSTAGE - 1;
        if (j < my_count) {
          bulk_wait_read<1>();  // the store issued one iteration ago has drained its stage
          issue_load(j);
        }
      }
    }
  }
  if (tid == 0) bulk_wait<0>();
}


// ---------------------------------------------------------------------------
// z pass (contiguous axis) with bulk-copy staging.
//
// Rows are contiguous in global memory on both sides (N doubles / N/2+1 double2), so a tile of
// TR rows moves with one cp.async.bulk (UBLKCP) per row -- issued by the lanes of warp 0, each
// lane owning "its" row of every stage, store and reload included (bulk groups are per thread).
// A row is transformed in place by the N/(2E) lanes of one warp as an N/2-point complex FFT of
// z[j] = x[2j] + i x[2j+1] plus the Hermitian split / merge; the row pitch in shared memory is
// N*8 + 32 bytes so the N/2+1 outputs fit over the N/2 inputs.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f64_1d(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;\n" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

// exp(-2 pi i m / 16), m = 0..7: with LP = M/E lanes per row and E = 8, w_N^(t + m LP) = w_N^t * this
__device__ __forceinline__ double2 root16(int m) {
  constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
  switch (m & 7) {
    case 0: return make_double2(1.0, 0.0);
    case 1: return make_double2(c1, -s1);
    case 2: return make_double2(h, -h);
    case 3: return make_double2(s1, -c1);
    case 4: return make_double2(0.0, -1.0);
    case 5: return make_double2(-s1, -c1);
    case 6: return make_double2(-h, -h);
    default: return make_double2(-c1, -s1);
  }
}

// One row of the z pass, in place in shared memory (rbase = the row, N/2 + 1 double2 of room): forward =
// N/2-point complex FFT of z[j] = x[2j] + i x[2j+1] + Hermitian split; inverse = Hermitian merge + FFT.
// Lane t of the LP = N/(2E) lanes that own the row; `auxrow` (AUX) = the row of the real multiplier.
template <int N, int E, bool C2R, bool AUX, bool TWCONJ = false>
__device__ __forceinline__ void zrow_transform(uint32_t rbase, uint32_t auxrow, int t,
                                               const double2 (*twr)[E], double2 wt,
                                               const double2 *__restrict__ twN, const ROp &op) {
  constexpr int M = N / 2;
  constexpr int LP = M / E;
  constexpr int DIR = C2R ? +1 : -1;
  const RowAccess acc{rbase};
  auto split_twiddle = [&](int m) -> double2 {
    if constexpr (E == 8) return m == 0 ? wt : cmul(wt, root16(m));
    else return __ldg(twN + t + m * LP);
  };
  double2 v[E];
  if constexpr (!C2R) {
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = lds128(acc.at(t + m * LP));
    if (op.kind == R_LOAD_SCALE) {
#pragma unroll
      for (int m = 0; m < E; ++m) v[m] = make_double2(v[m].x * op.a, v[m].y * op.a);
    }
    wp_stages<M, E, 1, DIR, 0, RowAccess, TWCONJ>(v, t, acc, twr);
    // Hermitian split: X[k] = E + w^k O, E = (Z[k] + conj Z[M-k])/2, O = (Z[k] - conj Z[M-k])/(2i)
    __syncwarp();
#pragma unroll
    for (int m = 0; m < E; ++m) sts128(acc.at(t + m * LP), v[m]);
    __syncwarp();
    double2 zm[E];
#pragma unroll
    for (int m = 0; m < E; ++m) zm[m] = lds128(acc.at((M - (t + m * LP)) & (M - 1)));
    __syncwarp();
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int k = t + m * LP;
      const double2 zk = v[m];
      const double2 e = make_double2(0.5 * (zk.x + zm[m].x), 0.5 * (zk.y - zm[m].y));
      const double2 o = make_double2(0.5 * (zk.y + zm[m].y), -0.5 * (zk.x - zm[m].x));
      const double2 w = split_twiddle(m);
      sts128(acc.at(k), cadd(e, cmul(w, o)));
      if (k == 0) sts128(acc.at(M), make_double2(e.x - o.x, 0.0));  // w_N^M = -1; E[0], O[0] real
    }
  } else {
    // Hermitian merge: Z[k] = E' + i O', E' = X[k] + conj X[M-k], O' = (X[k] - conj X[M-k]) conj(w^k)
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int k = t + m * LP;
      double2 xk = lds128(acc.at(k));
      double2 xm = lds128(acc.at(M - k));
      if (k == 0) {  // FFTW's c2r ignores the imaginary parts of the self-conjugate bins
        xk.y = 0.0;
        xm.y = 0.0;
      }
      const double2 e = make_double2(xk.x + xm.x, xk.y - xm.y);
      const double2 d = make_double2(xk.x - xm.x, xk.y + xm.y);
      const double2 o = cmul(d, cconj(split_twiddle(m)));
      v[m] = make_double2(e.x - o.y, e.y + o.x);
    }
    wp_stages<M, E, 1, DIR, 0, RowAccess, TWCONJ>(v, t, acc, twr);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int j = t + m * LP;
      double2 x = make_double2(op.a * v[m].x, op.a * v[m].y);
      if constexpr (AUX) {
        const double2 y = lds128(auxrow + j * 16);
        x.x *= y.x;
        x.y *= y.y;
      }
      sts128(acc.at(j), x);
    }
  }
}

template <int N, int TR, int NSTAGE, bool AUX>
struct ZTile {
  static constexpr int pitch = N * 8 + 32;
  static constexpr int main_bytes = TR * pitch;
  static constexpr int aux_bytes = AUX ? TR * N * 8 : 0;
  static constexpr int stage_bytes = main_bytes + aux_bytes;
  static constexpr int smem_bytes = NSTAGE * stage_bytes + 1024 + 64;
};

// C2R = false: real rows -> half-complex rows (forward);  C2R = true: half-complex -> real (inverse).
// AUX (C2R only): multiply the real result by a second real array (R_SCALE_MUL).
template <int N, int E, int TR, int NSTAGE, bool C2R, bool AUX, int MINB>
__global__ void __launch_bounds__(TR *(N / 2 / E), MINB)
    fft_zpass_tma(const void *__restrict__ in, void *__restrict__ out, const double2 *__restrict__ twN,
                  const double2 *__restrict__ twM, ROp op, int ntiles, int rev) {
  constexpr int M = N / 2;
  constexpr int LP = M / E;
  constexpr int DIR = C2R ? +1 : -1;
  constexpr int NSTG = StageCount<M>::value;
  using Z = ZTile<N, TR, NSTAGE, AUX>;
  constexpr uint32_t in_row_bytes = C2R ? (M + 1) * 16 : N * 8;
  constexpr uint32_t out_row_bytes = C2R ? N * 8 : (M + 1) * 16;
  static_assert(LP >= 8 && LP <= 32 && TR <= 32, "row must live inside one warp; warp 0 issues one copy per row");
  if constexpr (C2R) {
    if (op.skip) {
      pdl_wait();  // the flag is an earlier kernel's result
      if (*op.skip) return;  // uniform over the grid; nothing has been issued yet
    }
  }

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t *smem_al = smem_raw + (smem0 - smem_u32(smem_raw));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_al + NSTAGE * Z::stage_bytes);

  const int tid = threadIdx.x;
  const int row = tid / LP;
  const int t = tid % LP;
  const int my_count = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int i) {  // rev: last tile first (see PassGeom::reverse)
    const int tl = (int)blockIdx.x + i * (int)gridDim.x;
    return rev ? ntiles - 1 - tl : tl;
  };
  const char *gin = static_cast<const char *>(in);
  char *gout = static_cast<char *>(out);

  // warp 0, lane r: load row r of my i-th tile
  auto issue_load = [&](int i) {
    const int s = i % NSTAGE;
    const size_t grow = (size_t)tile_of(i) * TR + tid;
    if (tid == 0) mbar_expect_tx(&full[s], TR * (in_row_bytes + (AUX ? N * 8 : 0)));
    __syncwarp();
    if (tid < TR) {
      bulk_load_1d(smem_al + s * Z::stage_bytes + tid * Z::pitch, gin + grow * in_row_bytes, in_row_bytes, &full[s]);
      if constexpr (AUX)
        bulk_load_1d(smem_al + s * Z::stage_bytes + Z::main_bytes + tid * (N * 8),
                     reinterpret_cast<const char *>(op.aux) + grow * (N * 8), N * 8, &full[s]);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  double2 twr[NSTG > 1 ? NSTG - 1 : 1][E];
  wp_load_twiddles<M, E, 1, DIR, 0>(twr, t, twM);
  // split / merge twiddle w_N^k, k = t + m LP: for E == 8, LP = N/16 and w_N^(m LP) is a 16th root of unity
  const double2 wt = __ldg(twN + t);
  pdl_trigger();
  pdl_wait();
  if (tid < 32) {
#pragma unroll
    for (int i = 0; i < NSTAGE - 1; ++i)
      if (i < my_count) issue_load(i);
  }

  for (int i = 0; i < my_count; ++i) {
    const int s = i % NSTAGE;
    const uint32_t rbase = smem0 + s * Z::stage_bytes + row * Z::pitch;
    mbar_wait(&full[s], (i / NSTAGE) & 1);

    zrow_transform<N, E, C2R, AUX>(rbase, smem0 + s * Z::stage_bytes + Z::main_bytes + row * (N * 8), t, twr, wt, twN, op);
    fence_proxy_async();
    __syncthreads();
    if (tid < 32) {
      if (tid < TR) {
        const size_t grow = (size_t)tile_of(i) * TR + tid;
        const void *src = smem_al + s * Z::stage_bytes + tid * Z::pitch;
        if (C2R && op.kind == R_AXPY)
          bulk_reduce_add_f64_1d(gout + grow * out_row_bytes, src, out_row_bytes);
        else
          bulk_store_1d(gout + grow * out_row_bytes, src, out_row_bytes);
      }
      bulk_commit();
      const int j = i + NSTAGE - 1;
      if (j < my_count) {
        bulk_wait_read<1>();  // my row of the stage used one iteration ago has drained
        issue_load(j);
      }
    }
  }
  if (tid < 32) bulk_wait<0>();
}

// ---------------------------------------------------------------------------
// z round trip: half-complex row -> real row, times a real array, -> half-complex row, in ONE pass.
//
// likelihood_calc_h (HMC_models_testing.cpp:25-50) forms g_c = r * d_c(delta) in real space between an inverse and a
// forward transform (gradfft, gradient.cpp:38-74, then grad_inv_lap_FS, :157-211).  Only the z passes of the two
// transforms meet real space, and the product is local, so the pair "c2r z pass (x r) -> r2c z pass" runs on
// the row while it sits in shared memory: 24 B per cell cross HBM instead of 40, and three launches fewer per
// evaluation.  Same row arithmetic as fft_zpass_tma (zrow_transform), one twiddle set for both directions.
// ---------------------------------------------------------------------------
template <int N, int E, int TR, int NSTAGE, int MINB>
__global__ void __launch_bounds__(TR *(N / 2 / E), MINB)
    fft_zround_tma(const double2 *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ twN,
                   const double2 *__restrict__ twM, ROp op, int ntiles, int rev) {
  constexpr int M = N / 2;
  constexpr int LP = M / E;
  constexpr int NSTG = StageCount<M>::value;
  using Z = ZTile<N, TR, NSTAGE, true>;
  constexpr uint32_t row_bytes = (M + 1) * 16;
  static_assert(LP >= 8 && LP <= 32 && TR <= 32, "row must live inside one warp; warp 0 issues one copy per row");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t *smem_al = smem_raw + (smem0 - smem_u32(smem_raw));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_al + NSTAGE * Z::stage_bytes);

  const int tid = threadIdx.x;
  const int row = tid / LP;
  const int t = tid % LP;
  const int my_count = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int i) {
    const int tl = (int)blockIdx.x + i * (int)gridDim.x;
    return rev ? ntiles - 1 - tl : tl;
  };
  const char *gin = reinterpret_cast<const char *>(in);
  char *gout = reinterpret_cast<char *>(out);

  auto issue_load = [&](int i) {
    const int s = i % NSTAGE;
    const size_t grow = (size_t)tile_of(i) * TR + tid;
    if (tid == 0) mbar_expect_tx(&full[s], TR * (row_bytes + N * 8));
    __syncwarp();
    if (tid < TR) {
      bulk_load_1d(smem_al + s * Z::stage_bytes + tid * Z::pitch, gin + grow * row_bytes, row_bytes, &full[s]);
      bulk_load_1d(smem_al + s * Z::stage_bytes + Z::main_bytes + tid * (N * 8),
                   reinterpret_cast<const char *>(op.aux) + grow * (N * 8), N * 8, &full[s]);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  double2 twr[NSTG > 1 ? NSTG - 1 : 1][E];
  wp_load_twiddles<M, E, 1, +1, 0>(twr, t, twM);  // inverse-direction twiddles; the forward transform conjugates them
  const double2 wt = __ldg(twN + t);
  pdl_trigger();
  pdl_wait();
  if (tid < 32) {
#pragma unroll
    for (int i = 0; i < NSTAGE - 1; ++i)
      if (i < my_count) issue_load(i);
  }
  ROp fwd;
  fwd.kind = R_LOAD;

  for (int i = 0; i < my_count; ++i) {
    const int s = i % NSTAGE;
    const uint32_t rbase = smem0 + s * Z::stage_bytes + row * Z::pitch;
    const uint32_t auxrow = smem0 + s * Z::stage_bytes + Z::main_bytes + row * (N * 8);
    mbar_wait(&full[s], (i / NSTAGE) & 1);

    zrow_transform<N, E, true, true>(rbase, auxrow, t, twr, wt, twN, op);      // -> a * x(z) * aux(z), N reals in place
    __syncwarp();
    zrow_transform<N, E, false, false, true>(rbase, auxrow, t, twr, wt, twN, fwd);  // -> N/2 + 1 complex in place
    fence_proxy_async();
    __syncthreads();
    if (tid < 32) {
      if (tid < TR) {
        const size_t grow = (size_t)tile_of(i) * TR + tid;
        bulk_store_1d(gout + grow * row_bytes, smem_al + s * Z::stage_bytes + tid * Z::pitch, row_bytes);
      }
      bulk_commit();
      const int j = i + NSTAGE - 1;
      if (j < my_count) {
        bulk_wait_read<1>();
        issue_load(j);
      }
    }
  }
  if (tid < 32) bulk_wait<0>();
}

}  // namespace bgpu
