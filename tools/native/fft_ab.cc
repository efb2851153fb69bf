// tools/native/fft_ab.cc -- A/B of an env-selected FFT variant through the C ABI, without Python:
// the same real field goes through bgpu_fft_r2c / bgpu_fft_c2r on a default handle and on a handle
// created with VAR=1; results are compared bit for bit and the per-kernel-class device times of one
// r2c + c2r are printed for both (bgpu_profile_*).  Starts in a second, so it fits the tail of a GPU budget.
//   g++ -O2 -fopenmp -I include tools/native/fft_ab.cc -L barcode_b200 -lbarcode_b200 -Wl,-rpath,'$ORIGIN/../../barcode_b200' -o tools/native/fft_ab
//   tools/native/fft_ab BGPU_FFT_2WARP 512
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "barcode_gpu.h"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define CHECK(call)                                                      \
  do {                                                                   \
    if ((call) != 0) {                                                   \
      std::printf("FAILED %s: %s\n", #call, bgpu_last_error());          \
      std::fflush(stdout);                                               \
      return 1;                                                          \
    }                                                                    \
  } while (0)

static void report(const char *tag) {
  double ms[BGPU_PROFILE_KINDS];
  uint64_t cnt[BGPU_PROFILE_KINDS];
  if (bgpu_profile_end(ms, cnt, BGPU_PROFILE_KINDS) != 0) {
    std::printf("%s: profile_end failed: %s\n", tag, bgpu_last_error());
    return;
  }
  std::printf("%s:", tag);
  for (int k = 0; k < BGPU_PROFILE_KINDS; ++k)
    if (cnt[k]) std::printf("  %s %.4f ms / %llu", bgpu_profile_kind_name(k), ms[k], (unsigned long long)cnt[k]);
  std::printf("\n");
  std::fflush(stdout);
}

int main(int argc, char **argv) {
  const char *var = argc > 1 ? argv[1] : "BGPU_FFT_2WARP";
  const int N = argc > 2 ? std::atoi(argv[2]) : 512;
  const double t0 = now();
  const size_t n = (size_t)N * N * N, nh = (size_t)N * N * (N / 2 + 1);
  bgpu_params p;
  bgpu_default_params(&p);
  p.N1 = p.N2 = p.N3 = N;
  p.L1 = p.L2 = p.L3 = N * (200.0 / 64.0);
  bgpu_handle *A = nullptr, *B = nullptr;
  unsetenv(var);
  CHECK(bgpu_create(&p, &A));
  setenv(var, "1", 1);
  CHECK(bgpu_create(&p, &B));
  unsetenv(var);
  std::printf("%s A/B at %d^3: handles up at %.2f s\n", var, N, now() - t0);
  std::fflush(stdout);

  std::vector<double> in(n), backA(n), backB(n);
  std::vector<double> outA(2 * nh), outB(2 * nh);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) {
    uint64_t x = (uint64_t)i * 6364136223846793005ull + 1442695040888963407ull;
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    in[i] = (double)(int64_t)(x >> 11) * (1.0 / 4503599627370496.0) - 1.0;  // [-1, 1)
  }
  std::printf("input filled at %.2f s\n", now() - t0);
  std::fflush(stdout);

  CHECK(bgpu_fft_r2c(A, in.data(), outA.data()));
  CHECK(bgpu_fft_r2c(B, in.data(), outB.data()));
  size_t bad = 0;
  for (size_t i = 0; i < 2 * nh; ++i) bad += std::memcmp(&outA[i], &outB[i], 8) != 0;
  double s2 = 0;
  for (size_t i = 0; i < 2 * nh; i += 4097) s2 += outA[i] * outA[i];
  std::printf("r2c: %zu of %zu doubles differ between default and %s=1 (sampled |A|^2 %.6e) at %.2f s\n", bad, 2 * nh, var,
              s2, now() - t0);
  std::fflush(stdout);

  CHECK(bgpu_profile_begin());
  CHECK(bgpu_fft_r2c(B, in.data(), outB.data()));
  CHECK(bgpu_fft_c2r(B, outB.data(), backB.data()));
  report("variant r2c+c2r");
  CHECK(bgpu_profile_begin());
  CHECK(bgpu_fft_r2c(A, in.data(), outA.data()));
  CHECK(bgpu_fft_c2r(A, outA.data(), backA.data()));
  report("default r2c+c2r");
  bad = 0;
  double err = 0;
  for (size_t i = 0; i < n; ++i) {
    bad += std::memcmp(&backA[i], &backB[i], 8) != 0;
    const double e = backA[i] - in[i];
    err = e * e > err ? e * e : err;
  }
  std::printf("c2r: %zu of %zu doubles differ; max |c2r(r2c(in)) - in| (default) %.3e at %.2f s\n", bad, n, std::sqrt(err),
              now() - t0);
  std::fflush(stdout);

  // the operand pass (AUX = 1, K_MULREAL) and a second round of timings
  {
    std::vector<double> &corr = backA;  // reuse: positive multipliers
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) corr[i] = 1.0 + 0.5 * in[i];
    std::vector<double> &cA = outA, &cB = outB;  // n <= 2 nh doubles
    CHECK(bgpu_profile_begin());
    CHECK(bgpu_convolve_inv_corr(B, in.data(), corr.data(), cB.data()));
    report("variant convolve");
    CHECK(bgpu_profile_begin());
    CHECK(bgpu_convolve_inv_corr(A, in.data(), corr.data(), cA.data()));
    report("default convolve");
    bad = 0;
    for (size_t i = 0; i < n; ++i) bad += std::memcmp(&cA[i], &cB[i], 8) != 0;
    std::printf("convolve_inv_corr: %zu of %zu doubles differ at %.2f s\n", bad, n, now() - t0);
    std::fflush(stdout);
  }
  bgpu_destroy(A);
  bgpu_destroy(B);
  std::printf("done at %.2f s\n", now() - t0);
  return 0;
}
