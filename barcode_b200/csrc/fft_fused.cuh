// barcode_b200/csrc/fft_fused.cuh
//
// The z and y passes of the 3-D transform in ONE persistent, warp-specialised kernel, with the
// intermediate array kept in L2 instead of making a round trip through HBM.
//
// Both passes work inside one x plane ([N][N/2+1] complex numbers): the z pass turns the N real
// rows of a plane into half-complex rows, the y pass transforms the N/2+1 columns of the same
// plane.  Run as two kernels (fft_tma.cuh) the array crosses the HBM interface four times (z: real
// in, complex out; y: complex in, complex out) because N^2 (N/2+1) complex numbers (135 MB at 256^3)
// do not fit the 126 MB L2.  Here every CTA (one per SM) carries two roles of 8 warps each:
//   * the PRODUCER role (z pass for r2c, y pass for c2r) walks its tiles plane by plane and writes
//     its results with bulk stores, then bumps a per-plane completion counter in global memory
//     (red.release.gpu) once the store group has completed;
//   * the CONSUMER role (y pass for r2c, z pass for c2r) walks the same planes a little later: before
//     it requests a tile it waits (ld.acquire.gpu) until every producer tile of that plane is done,
//     anywhere on the GPU, and then loads the tile by TMA -- from L2, where the producer left it;
//   * the producer is throttled to at most `lead` planes ahead of the consumer of its own CTA, so the
//     live intermediate stays a few tens of MB.
// Each role keeps its own shared-memory ring, mbarriers, named barrier and register-resident
// twiddles; the transforms themselves are the ones of fft_tma.cuh (zrow_transform, wp_stages).
// Per element the pair of passes now moves 8 B (real) + 16 B (complex) across HBM instead of
// 8 + 16 + 16 + 16.
#pragma once

#include "fft_tma.cuh"

namespace bgpu {

struct ZyCtl {
  unsigned long long *ready;   // [N] per-plane count of completed producer tiles, monotone over launches
  unsigned long long target;   // value ready[x] reaches when plane x is complete in THIS launch
  int lead;                    // planes the producer may run ahead of its CTA's consumer
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void role_barrier(int id) { asm volatile("bar.sync %0, 256;\n" ::"r"(id) : "memory"); }

template <int N, int TR, int NSZ, int NSY, bool AUX>
struct ZySmem {
  using Z = ZTile<N, TR, NSZ, AUX>;
  static constexpr int ystage = N * 8 * 16;                             // N rows x 128 B, 1024-aligned
  static constexpr int zstage = (Z::stage_bytes + 127) & ~127;
  static constexpr int yring = 0;
  static constexpr int zring = NSY * ystage;
  static constexpr int bars = zring + NSZ * zstage;                     // fullY[NSY], fullZ[NSZ], progress
  static constexpr int bytes = bars + (NSY + NSZ) * 8 + 16 + 1024;      // + alignment slack
};

// C2R = false: real rows (zsrc) -> half-complex rows (cplx) -> y pass in place (forward).
// C2R = true : y pass in place on cplx, then half-complex rows -> real rows (zdst) (inverse).
// AUX (C2R only): the real result is multiplied by op.aux (R_SCALE_MUL).
template <int N, int EZ, int TR, int EY, int NSZ, int NSY, bool C2R, bool AUX>
__global__ void __launch_bounds__(512, 1)
    fft_zy_fused(const __grid_constant__ CUtensorMap ymap, const void *__restrict__ zsrc, void *__restrict__ zdst,
                 const double2 *__restrict__ twN, const double2 *__restrict__ twM, ROp op, ZyCtl ctl) {
  constexpr int M = N / 2;
  constexpr int NZH = M + 1;
  constexpr int ZT = (NZH + 7) / 8;        // y tiles per plane
  constexpr int ZPP = N / TR;              // z tiles per plane
  constexpr int ROWS_PER_BOX = N > 256 ? 256 : N;
  using L = ZySmem<N, TR, NSZ, NSY, AUX>;
  using Z = typename L::Z;
  static_assert(TR * (M / EZ) == 256 && 8 * (N / EY) == 256, "both roles are 256 threads wide");
  static_assert(L::bytes <= 227 * 1024, "rings do not fit");
  constexpr bool Z_PRODUCES = !C2R;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_al = smem_raw + (smem0 - smem_u32(smem_raw));
  uint64_t *fullY = reinterpret_cast<uint64_t *>(smem_al + L::bars);
  uint64_t *fullZ = fullY + NSY;
  volatile int *cons_plane = reinterpret_cast<volatile int *>(fullZ + NSZ);

  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSY; ++s) mbar_init(&fullY[s], 1);
#pragma unroll
    for (int s = 0; s < NSZ; ++s) mbar_init(&fullZ[s], 1);
    *cons_plane = 0;
    fence_barrier_init();
  }
  __syncthreads();

  // producer side: hold back while more than `lead` planes ahead of this CTA's consumer
  auto throttle = [&](int plane) {
    while (plane > *cons_plane + ctl.lead) __nanosleep(64);
  };
  // consumer side: publish how far this CTA's consumer has got, then wait for the plane to be complete
  // (`seen` = the counter as read one iteration earlier, so the usual case costs no L2 round trip here)
  auto await_plane = [&](int plane, unsigned long long seen) {
    *cons_plane = plane;
    while (seen < ctl.target) {
      __nanosleep(32);
      seen = ld_acquire_u64(ctl.ready + plane);
    }
  };

  if (tid < 256) {
    // ------------------------------------------------------------------ z role
    constexpr int LP = M / EZ;
    constexpr int DIR = C2R ? +1 : -1;
    constexpr int NSTG = StageCount<M>::value;
    constexpr uint32_t in_row_bytes = C2R ? (M + 1) * 16 : N * 8;
    constexpr uint32_t out_row_bytes = C2R ? N * 8 : (M + 1) * 16;
    const int row = tid / LP, t = tid % LP;
    const int ntiles = N * ZPP;
    const int my_count = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const char *gin = static_cast<const char *>(zsrc);
    char *gout = static_cast<char *>(zdst);
    uint8_t *ring = smem_al + L::zring;
    const uint32_t ring0 = smem0 + L::zring;

    auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
    unsigned long long seen = 0;
    // warp 0 of the role, lane r: load row r of my i-th tile
    auto issue_load = [&](int i) {
      const int s = i % NSZ;
      const int tile = tile_of(i);
      if (tid == 0) {
        if (Z_PRODUCES) throttle(tile / ZPP);
        else await_plane(tile / ZPP, seen);
        mbar_expect_tx(&fullZ[s], TR * (in_row_bytes + (AUX ? N * 8 : 0)));
      }
      __syncwarp();
      if (tid < TR) {
        if (!Z_PRODUCES) fence_proxy_async_all();
        const size_t grow = (size_t)tile * TR + tid;
        bulk_load_1d(ring + s * L::zstage + tid * Z::pitch, gin + grow * in_row_bytes, in_row_bytes, &fullZ[s]);
        if constexpr (AUX)
          bulk_load_1d(ring + s * L::zstage + Z::main_bytes + tid * (N * 8),
                       reinterpret_cast<const char *>(op.aux) + grow * (N * 8), N * 8, &fullZ[s]);
      }
      if (!Z_PRODUCES && tid == 0 && i + 1 < my_count) seen = ld_relaxed_u64(ctl.ready + tile_of(i + 1) / ZPP);
    };
    // the stores of tile i have completed (every lane waited for its own rows): publish it
    auto signal = [&](int i) {
      if (!Z_PRODUCES) return;
      fence_proxy_async_all();
      __syncwarp();
      if (tid == 0) red_release_add_u64(ctl.ready + tile_of(i) / ZPP, 1ull);
    };

    if (tid < 32) {
#pragma unroll
      for (int i = 0; i < NSZ - 1; ++i)
        if (i < my_count) issue_load(i);
    }
    double2 twr[NSTG > 1 ? NSTG - 1 : 1][EZ];
    wp_load_twiddles<M, EZ, 1, DIR, 0>(twr, t, twM);
    const double2 wt = __ldg(twN + t);

    for (int i = 0; i < my_count; ++i) {
      const int s = i % NSZ;
      mbar_wait(&fullZ[s], (i / NSZ) & 1);
      zrow_transform<N, EZ, C2R, AUX>(ring0 + s * L::zstage + row * Z::pitch,
                                      ring0 + s * L::zstage + Z::main_bytes + row * (N * 8), t, twr, wt, twN, op);
      fence_proxy_async();
      role_barrier(1);
      if (tid < 32) {
        if (tid < TR) {
          const size_t grow = (size_t)tile_of(i) * TR + tid;
          const void *src = ring + s * L::zstage + tid * Z::pitch;
          if (C2R && op.kind == R_AXPY)
            bulk_reduce_add_f64_1d(gout + grow * out_row_bytes, src, out_row_bytes);
          else
            bulk_store_1d(gout + grow * out_row_bytes, src, out_row_bytes);
        }
        bulk_commit();
        if (Z_PRODUCES && i > 1) {
          bulk_wait<2>();  // tile i-2 is in global memory; publish it BEFORE a load that may have to wait
          signal(i - 2);
        }
        bulk_wait_read<1>();  // the stage of tile i-1 can be refilled
        const int j = i + NSZ - 1;
        if (j < my_count) issue_load(j);
      }
    }
    if (tid < 32) {
      bulk_wait<0>();
      if (my_count > 1) signal(my_count - 2);
      if (my_count > 0) signal(my_count - 1);
    }
  } else {
    // ------------------------------------------------------------------ y role
    constexpr int LP = N / EY;
    constexpr int DIR = C2R ? +1 : -1;
    constexpr int NSTG = StageCount<N>::value;
    const int yt = tid - 256;
    const int p = yt / LP, t = yt % LP;
    const int ntiles = N * ZT;
    const int my_count = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    uint8_t *ring = smem_al + L::yring;
    const uint32_t ring0 = smem0 + L::yring;

    auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
    unsigned long long seen = 0;
    auto issue_load = [&](int i) {  // role thread 0 only
      const int s = i % NSY;
      const int tile = tile_of(i);
      const int plane = tile / ZT, zt = tile % ZT;
      if (Z_PRODUCES) {
        await_plane(plane, seen);
        fence_proxy_async_all();
      } else {
        throttle(plane);
      }
      mbar_expect_tx(&fullY[s], L::ystage);
#pragma unroll
      for (int r0 = 0; r0 < N; r0 += ROWS_PER_BOX)
        tma_load_3d(ring + s * L::ystage + r0 * 128, &ymap, zt * 16, r0, plane, &fullY[s]);
      if (Z_PRODUCES && i + 1 < my_count) seen = ld_relaxed_u64(ctl.ready + tile_of(i + 1) / ZT);
    };
    auto signal = [&](int i) {  // role thread 0 only; its store group has completed
      if (Z_PRODUCES) return;
      fence_proxy_async_all();
      red_release_add_u64(ctl.ready + tile_of(i) / ZT, 1ull);
    };

    if (yt == 0) {
#pragma unroll
      for (int i = 0; i < NSY - 1; ++i)
        if (i < my_count) issue_load(i);
    }
    double2 twr[NSTG > 1 ? NSTG - 1 : 1][EY];
    wp_load_twiddles<N, EY, 1, DIR, 0>(twr, t, twN);

    for (int i = 0; i < my_count; ++i) {
      const int s = i % NSY;
      const uint32_t tbase = ring0 + s * L::ystage;
      mbar_wait(&fullY[s], (i / NSY) & 1);
      double2 v[EY];
#pragma unroll
      for (int m = 0; m < EY; ++m) v[m] = lds128(tile_addr(tbase, t + m * LP, p));
      wp_stages<N, EY, 1, DIR, 0>(v, t, ColAccess{tbase, p}, twr);
      __syncwarp();
#pragma unroll
      for (int m = 0; m < EY; ++m) sts128(tile_addr(tbase, t + m * LP, p), v[m]);
      fence_proxy_async();
      role_barrier(2);
      if (yt == 0) {
        const int tile = tile_of(i);
        const int plane = tile / ZT, zt = tile % ZT;
#pragma unroll
        for (int r0 = 0; r0 < N; r0 += ROWS_PER_BOX)
          tma_store_3d(&ymap, zt * 16, r0, plane, ring + s * L::ystage + r0 * 128);
        bulk_commit();
        if (!Z_PRODUCES && i > 1) {
          bulk_wait<2>();
          signal(i - 2);
        }
        bulk_wait_read<1>();
        const int j = i + NSY - 1;
        if (j < my_count) issue_load(j);
      }
    }
    if (yt == 0) {
      bulk_wait<0>();
      if (my_count > 1) signal(my_count - 2);
      if (my_count > 0) signal(my_count - 1);
      if (Z_PRODUCES) *cons_plane = 0x3fffffff;  // consumer done: release this CTA's producer for good
    }
  }
  // a z-role consumer that has finished likewise stops throttling the y-role producer
  if (!Z_PRODUCES && tid == 0) *cons_plane = 0x3fffffff;
}

}  // namespace bgpu
