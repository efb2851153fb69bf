// barcode_b200/csrc/fft3d.h -- host interface of the hand-written 3-D real FFT
// (kernels in fft.cuh, launch sequences in fft_plan.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>

#include "fft_ops.h"

namespace bgpu {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();

// what the slab transform needs from the communication layer (nccl_comm.cu)
struct SlabComm {
  virtual ~SlabComm() {}
  // peer h receives doubles [h*count, (h+1)*count) of `send` into block `my rank` of its `recv`
  virtual void all_to_all(const void *send, void *recv, size_t count_doubles, cudaStream_t st) = 0;
  // every rank's work enqueued before this call is complete before any rank's work enqueued after it starts
  virtual void barrier(cudaStream_t st) = 0;
};

// Streaming hooks for the host-pointer API: the first transform of an evaluation can start on the rows
// that have already arrived from the host, the last one can hand rows back while later rows are still
// being transformed.  `before(ctx, c)` is called ahead of the z pass over row chunk c of an r2c,
// `after(ctx, c)` behind the z pass over chunk c of a c2r; chunk c = rows [c, c+1) * Ns*N/chunks.
struct ChunkHooks {
  int chunks = 1;
  void (*before)(void *ctx, int chunk) = nullptr;
  void (*after)(void *ctx, int chunk) = nullptr;
  void *ctx = nullptr;
};

struct Fft3d {
  int N = 0;
  double2 *twN = nullptr;  // exp(-2 pi i k / N)
  double2 *twM = nullptr;  // exp(-2 pi i k / (N/2))
  cudaStream_t stream = nullptr;
  int strided_blocks = 0;  // persistent grid of the pipelined strided pass: SMs x resident CTAs
  int sm_count = 0;
  int device = 0;          // CUDA device this plan lives on (kernel attributes and occupancy are per device)
  bool use_tma = true;     // TMA-staged strided pass (fft_tma.cuh); BGPU_FFT_TMA=0 selects the cp.async one
  mutable const ChunkHooks *hooks = nullptr;  // set around ONE transform by the caller, consumed by its z pass
  // fused z+y kernel (fft_fused.cuh): per-plane completion counters, their running target, producer lead
  bool use_fused = false;  // BGPU_FFT_FUSED=1
  bool two_warp = false;   // BGPU_FFT_2WARP=1: 512-point strided pencils over two warps (fft_tma.cuh, ColAccessWide)
  bool use_pdl = true;     // (BGPU_PDL=0 turns it off) programmatic dependent launch of the TMA-staged passes
  bool z_round = true;     // (BGPU_ZROUND=0 turns it off) calc_h = 0: the c2r / r2c z passes around r * d_c(delta) fused
  bool share_x = true;     // (BGPU_SHARE_X=0 turns it off) the y and z components of a K_DISP / K_GRAD / K_INVLAP triple share one x pass
  // (BGPU_PINGPONG=0 turns it off; single GPU, N <= 256) consecutive passes walk their tiles in alternating directions,
  // so a pass starts on the end of the array the previous pass wrote last (still in the 126 MB L2 at 256^3)
  bool pingpong = false;
  mutable int rev_state = 0;
  int next_rev() const {
    if (!pingpong || G > 1) return 0;
    rev_state ^= 1;
    return rev_state;
  }
  bool force_generic = false;  // BGPU_FFT_SLAB_GENERIC=1: slab passes through fft_slab_generic.cuh even where TMA fits
  unsigned long long *zy_ready = nullptr;
  mutable unsigned long long zy_epoch = 0;
  int zy_lead = 0;

  // tensor maps of the half-grid arrays the TMA pass has touched (keyed by base pointer and layout)
  struct MapEntry {
    const void *base;
    int axis;
    bool cplx;
    int layout, n_slow, n_mid;
    CUtensorMap map;
  };
  mutable std::vector<MapEntry> maps_;
  const CUtensorMap &tensor_map(const void *base, int axis, bool cplx, int layout, int n_slow, int n_mid) const;

  // x-slab decomposition over G ranks: this rank holds x planes [rank*Ns, (rank+1)*Ns) of
  // every real array, [Ns][N][N], and y rows [rank*Ns, ...) of every k-space array, [N][Ns][N/2+1]
  // ("transposed" layout).  Set before init(); G == 1 is the plain cube.
  int G = 1, rank = 0, Ns = 0;
  SlabComm *comm = nullptr;                  // all-to-all provider (NCCL), owned by the caller
  double2 *sendbuf = nullptr, *recvbuf = nullptr;  // packed [peer][Ns][Ns][N/2+1], owned by the caller
  // fused transpose over peer memory: peer_recv[b][h] = receive buffer b of rank h as mapped into this
  // process (own rank: the local pointer); two buffers alternate (`parity`)
  bool p2p = false;
  double2 *peer_recv[2][8] = {};
  mutable int parity = 0;
  void barrier() const;  // cross-rank, stream-ordered

  void init(int n, cudaStream_t st);
  void destroy();
  static bool supported(int n);

  // real [N][N][N] -> half-complex [N][N][N/2+1], unnormalised (fftR2C).
  //   lop : real-space load functor of the z pass
  //   sop : k-space store functor of the last (x) pass; it writes to `xout`
  //         when that is non-null (accumulating back-projection), else to `out`.
  void r2c(const double *in, double2 *out, double2 *xout, ROp lop, KOp sop) const;

  // half-complex -> real (fftC2R).  `in` is preserved when work != in.
  //   lop : k-space load functor of the first (x) pass
  //   sop : real-space store functor of the z pass (carries the 1/N)
  void c2r(const double2 *in, double2 *work, double *out, KOp lop, ROp sop) const;

  // Single passes for transforms that share their x pass (fft_ops.h K_MULK*, K_COMP_UNIT; single GPU, TMA sizes).
  //   xpass   one x pass in -> out, dir = -1 forward / +1 inverse, load and store functors of RotCtxX
  //   c2r_yz  the rest of an inverse transform: y pass in -> work with load functor ylop (in is preserved), z pass -> out
  //   r2c_zy  the start of a forward transform: z pass in -> work, y pass work -> yout with store functor ysop
  //   ypass   one y pass in -> out with the extended functors (K_MULK* included)
  //   zround  z round trip in place on a y-passed array: c2r z pass, times sop.a * sop.aux (real array), r2c z pass
  //           (fft_tma.cuh fft_zround_tma) -- the real-space product of likelihood_calc_h without its HBM round trip
  bool can_share_x() const;
  bool can_zround() const;
  void ypass(const double2 *in, double2 *out, int dir, KOp lop, KOp sop) const;
  void zround(double2 *work, ROp sop) const;
  void xpass(const double2 *in, double2 *out, int dir, KOp lop, KOp sop) const;
  void c2r_yz(const double2 *in, double2 *work, double *out, KOp ylop, ROp sop) const;
  void r2c_zy(const double *in, double2 *work, double2 *yout, ROp lop, KOp ysop) const;

  // The same sharing on a slab-decomposed chain, inverse direction (a triple of inverse transforms whose y and z
  // components differ only by k_y / k_z): ONE transposing x pass (k_c := 1) whose result stays in the receive
  // buffer, then per component a y pass that reads that buffer with K_MULK and the z pass -- 2 transposes per triple
  // instead of 3.  (BGPU_SHARE_X_SLAB=0 turns it off; the TMA-staged sizes.)
  bool share_x_slab = true;
  bool can_share_x_slab() const;
  void xpass_shared_inverse(const double2 *in, KOp lop) const;
  void c2r_yz_shared(double2 *work, double *out, KOp ylop, ROp sop) const;
  mutable const double2 *shared_recv = nullptr;  // where xpass_shared_inverse left its result (packed [src][x_l][y_l][z])
};

}  // namespace bgpu
