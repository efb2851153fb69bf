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
