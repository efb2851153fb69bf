// barcode_b200/csrc/util.h -- error plumbing shared by the CUDA translation units.
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#define BGPU_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess)                                                                   \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " \
                               + __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");   \
  } while (0)

#include <atomic>
#include <cstdint>
namespace bgpu {
extern std::atomic<uint64_t> g_kernel_launches;
}
// after a <<<>>> launch (or n of them): surface launch errors and count it
#define BGPU_LAUNCHED(n)                                       \
  do {                                                         \
    BGPU_CUDA(cudaGetLastError());                             \
    ::bgpu::g_kernel_launches.fetch_add((n), std::memory_order_relaxed); \
  } while (0)

// ---------------------------------------------------------------------------
// optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on
// the launching stream around every launch of a kernel class.  Off by default;
// costs nothing when off.
// ---------------------------------------------------------------------------
#include <vector>
namespace bgpu {
enum KernelKind : int {
  KK_FFT_STRIDED = 0,
  KK_FFT_R2C_Z = 1,
  KK_FFT_C2R_Z = 2,
  KK_SCATTER = 3,
  KK_GATHER = 4,
  KK_RESIDUAL = 5,
  KK_REDUCE = 6,
  KK_STREAM = 7,
  KK_COLOUR = 8,
  KK_FFT_STRIDED_X = 9,
  KK_ALLTOALL = 10,
  KK_HALO = 11,
  KK_FFT_ZY_R2C = 12,  // fused z+y passes (fft_fused.cuh)
  KK_FFT_ZY_C2R = 13,
  KK_FFT_ZROUND = 14,  // c2r z pass x real array -> r2c z pass on the row in shared memory (fft_tma.cuh)
  KK_COUNT = 15
};

struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // start/stop pairs
  std::vector<int> kind;
  size_t used = 0;              // pairs in use
  void record_start(int k, cudaStream_t st);
  void record_stop(cudaStream_t st);
};
extern Profiler g_prof;

struct ProfScope {
  cudaStream_t st;
  bool active;
  ProfScope(int kind, cudaStream_t s) : st(s), active(g_prof.on) {
    if (active) g_prof.record_start(kind, st);
  }
  ~ProfScope() {
    if (active) g_prof.record_stop(st);
  }
};
}  // namespace bgpu
