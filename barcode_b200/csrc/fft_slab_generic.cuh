// barcode_b200/csrc/fft_slab_generic.cuh
//
// Strided pass of the slab-decomposed transform for sizes the TMA-staged kernel (fft_tma.cuh) does not
// cover: N = 1024, where a pencil no longer fits the lanes of one warp.  Same structure as
// fft_strided_pass_pipelined (fft.cuh): persistent CTAs, T pencils adjacent in z per tile, eight
// elements per thread, radix-8 Stockham stages through shared memory, the next tile prefetched with
// cp.async into thread-private slots -- but every address goes through the slab geometry:
//   AXIS 1 (y pass): this rank's x planes, [Ns][N][N/2+1]; the rows may instead be grouped by the peer
//                    that sends / receives them, [peer][x_l][y_l][z] (the all-to-all buffers);
//   AXIS 0 (x pass): this rank's y rows of every x plane, [N][Ns][N/2+1] (the transposed layout).
// BASELINE.json configs[4] (1024^3 over 8 GPUs) runs through this kernel; it is also instantiated
// for N = 128 so that the geometry can be tested against the single-GPU chain (BGPU_FFT_TMA=0).
#pragma once

#include "fft.cuh"

namespace bgpu {

struct SlabGeom {
  int n_other;     // "other" indices this rank holds (Ns)
  int other0;      // global index of other == 0 (k-space functors need global wave numbers)
  int in_packed;   // AXIS 1 only: rows per peer block of the input (Ns), 0 = plain layout
  int out_packed;  // AXIS 1 only: same for the output
};

template <int N, int AXIS>
__device__ __forceinline__ size_t slab_off(int r, int other, int iz, int n_other, int packed, int pitch) {
  if (AXIS == 0) return ((size_t)r * n_other + other) * pitch + iz;
  if (packed) {
    const int peer = r / packed, rl = r - peer * packed;
    return (((size_t)peer * packed + other) * packed + rl) * pitch + iz;
  }
  return ((size_t)other * N + r) * pitch + iz;
}

template <int N, int T>
__device__ __forceinline__ void slab_tile_coords(int tile, int p, int n_other, int &other, int &iz) {
  constexpr int NTZ = (N / 2) / T;
  const int nmain = n_other * NTZ;
  if (tile < nmain) {
    other = tile / NTZ;
    iz = (tile % NTZ) * T + p;
  } else {  // the Nyquist plane z = N/2: T pencils adjacent along the other axis
    other = (tile - nmain) * T + p;
    iz = N / 2;
  }
}

template <int N, int T, int DIR, int AXIS>
__global__ void __launch_bounds__(T *N / 8, (T * N / 8 <= 256) ? 3 : 1)
    fft_strided_pass_slab(const double2 *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ tw,
                          KOp lop, KOp sop, SlabGeom geo) {
  extern __shared__ double2 smem[];
  constexpr int NZH = N / 2 + 1;
  const int ntiles = geo.n_other * ((N / 2) / T) + geo.n_other / T;
  double2 *xch = smem;          // stage exchanges of the current tile
  double2 *pre = smem + N * T;  // thread-private landing slots of the next tile
  const int p = threadIdx.x % T;
  const int t = threadIdx.x / T;

  int tile = blockIdx.x;
  int other = 0, iz = 0;
  if (tile < ntiles) {
    slab_tile_coords<N, T>(tile, p, geo.n_other, other, iz);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int r = t + m * (N / 8);
      cp_async16(pre + r * T + p, in + slab_off<N, AXIS>(r, other, iz, geo.n_other, geo.in_packed, NZH));
    }
  }
  cp_async_commit();

  while (tile < ntiles) {
    cp_async_wait_all();
    double2 v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = pre[(t + m * (N / 8)) * T + p];

    const int next = tile + gridDim.x;
    if (next < ntiles) {
      int o2, z2;
      slab_tile_coords<N, T>(next, p, geo.n_other, o2, z2);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = t + m * (N / 8);
        cp_async16(pre + r * T + p, in + slab_off<N, AXIS>(r, o2, z2, geo.n_other, geo.in_packed, NZH));
      }
    }
    cp_async_commit();

    const int og = other + geo.other0;
    if (lop.kind == K_MULREAL) {
      // the real multiplier lives in the same layout as the (plain) input, row pitch N/2+2
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = t + m * (N / 8);
        const double f = __ldg(lop.real0 + slab_off<N, AXIS>(r, other, iz, geo.n_other, 0, NZH + 1));
        v[m] = make_double2(v[m].x * f, v[m].y * f);
      }
    } else if (lop.kind != K_NONE) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = t + m * (N / 8);
        v[m] = kop_load<N>(lop, v[m], 0, AXIS == 0 ? r : og, AXIS == 0 ? og : r, iz);
      }
    }

    SmStrided sm{xch, T, p};
    fft_stages<N, 1, DIR, SmStrided>(v, t, tw, sm);

#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int r = t + m * (N / 8);
      const size_t off = slab_off<N, AXIS>(r, other, iz, geo.n_other, geo.out_packed, NZH);
      if (sop.kind != K_NONE)
        kop_store<N>(sop, out, v[m], off, AXIS == 0 ? r : og, AXIS == 0 ? og : r, iz);
      else
        out[off] = v[m];
    }
    tile = next;
    if (tile < ntiles) slab_tile_coords<N, T>(tile, p, geo.n_other, other, iz);
  }
}

}  // namespace bgpu
