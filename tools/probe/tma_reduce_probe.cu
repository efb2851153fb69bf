// Probe: which box shapes does cp.reduce.async.bulk.tensor.3d (f64 add, SWIZZLE_NONE) accept on sm_100a?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_reduce_probe tma_reduce_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../barcode_b200/csrc/fft3d.h"
#include "../../barcode_b200/csrc/fft_tma.cuh"
using namespace bgpu;

__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, int cells, int c0, int c1, int c2, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 127u) & ~127u;
  double *tile = reinterpret_cast<double *>(smem_raw + (s0 - smem_u32(smem_raw)));
  for (int c = threadIdx.x; c < cells; c += blockDim.x) tile[c] = 1.0;
  if (mode & 4) atomicAdd(&tile[threadIdx.x % cells], 1.0);
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mode & 1) tma_reduce_add_3d(&map, c0, c1, c2, tile);
    else tma_store_3d(&map, c0, c1, c2, tile);
    bulk_commit();
    bulk_wait<0>();
  }
}

int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  const int N = 128;
  double *rho;
  cudaMalloc(&rho, sizeof(double) * N * N * N);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<EncodeTiledFn>(p);
  struct Case { int lz, ly, lx, c0, c1, c2, mode; };
  std::vector<Case> cases = {{40, 15, 23, 100, 120, 120, 1}, {40, 15, 23, -4, 0, 0, 1}, {40, 15, 23, 0, -3, 0, 1},
                             {40, 15, 23, 0, 0, -3, 1}, {40, 15, 23, -4, -3, -3, 0}, {40, 15, 23, 0, 0, 0, 5},
                             {40, 15, 23, 124, 0, 0, 1}, {40, 15, 23, 0, 125, 0, 1}, {40, 15, 23, 0, 0, 125, 1},
                             {40, 16, 24, 0, -4, 0, 1}, {40, 16, 24, -2, 0, 0, 1}};
  for (size_t ic = 0; ic < cases.size(); ++ic) {
    if (only >= 0 && (int)ic != only) continue;
    auto &c = cases[ic];
    cudaMemset(rho, 0, sizeof(double) * N * N * N);
    CUtensorMap m;
    const cuuint64_t dims[3] = {N, N, N};
    const cuuint64_t strides[2] = {N * 8, (cuuint64_t)N * N * 8};
    const cuuint32_t box[3] = {(cuuint32_t)c.lz, (cuuint32_t)c.ly, (cuuint32_t)c.lx};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, rho, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int cells = c.lz * c.ly * c.lx;
    const int smem = cells * 8 + 128;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 256, smem>>>(m, cells, c.c0, c.c1, c.c2, c.mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> h((size_t)N * N * N);
    double sum = 0;
    if (e == cudaSuccess) {
      cudaMemcpy(h.data(), rho, sizeof(double) * h.size(), cudaMemcpyDeviceToHost);
      for (double v : h) sum += v;
    }
    printf("box %dx%dx%d (%d B) at (%d,%d,%d) mode %d: encode %d, run: %s, sum %.1f (cells %d)\n", c.lz, c.ly, c.lx,
           cells * 8, c.c0, c.c1, c.c2, c.mode, (int)r, cudaGetErrorString(e), sum, cells);
    if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
  }
  return 0;
}
