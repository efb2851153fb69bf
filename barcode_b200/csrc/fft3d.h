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

struct Fft3d {
  int N = 0;
  double2 *twN = nullptr;  // exp(-2 pi i k / N)
  double2 *twM = nullptr;  // exp(-2 pi i k / (N/2))
  cudaStream_t stream = nullptr;
  int strided_blocks = 0;  // persistent grid of the pipelined strided pass: SMs x resident CTAs
  int sm_count = 0;
  bool use_tma = true;     // TMA-staged strided pass (fft_tma.cuh); BGPU_FFT_TMA=0 selects the cp.async one

  // tensor maps of the half-grid arrays the TMA pass has touched (keyed by base pointer)
  struct MapEntry {
    const void *base;
    int axis;
    bool cplx;
    CUtensorMap map;
  };
  mutable std::vector<MapEntry> maps_;
  const CUtensorMap &tensor_map(const void *base, int axis, bool cplx) const;

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
};

}  // namespace bgpu
