// tools/native/grad_ab.cc -- A/B of an env-selected variant on the whole gradient, through the C ABI, without Python:
// the same synthetic problem (smooth positive spectrum, unit noise / window, random data and signal) goes through
// bgpu_gradient_psi and bgpu_psi on a default handle and on a handle created with VAR=1; prints the relative L2
// difference of the gradients, the two energies, and the per-kernel-class device times of one evaluation each.
//   g++ -O2 -fopenmp -I include tools/native/grad_ab.cc -L barcode_b200 -lbarcode_b200 -Wl,-rpath,'$ORIGIN/../../barcode_b200' -o tools/native/grad_ab
//   tools/native/grad_ab BGPU_SHARE_X[=value] 256 [calc_h [sfmodel [rsd [likelihood [masskernel]]]]]
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "barcode_gpu.h"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define CHECK(call)                                             \
  do {                                                          \
    if ((call) != 0) {                                          \
      std::printf("FAILED %s: %s\n", #call, bgpu_last_error()); \
      std::fflush(stdout);                                      \
      return 1;                                                 \
    }                                                           \
  } while (0)

static double unit_noise(uint64_t i, uint64_t salt) {  // [-1, 1), hash of the index
  uint64_t x = (i + salt * 0x9e3779b97f4a7c15ull) * 6364136223846793005ull + 1442695040888963407ull;
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  return (double)(int64_t)(x >> 11) * (1.0 / 4503599627370496.0) - 1.0;
}

static void report(const char *tag) {
  double ms[BGPU_PROFILE_KINDS], total = 0;
  uint64_t cnt[BGPU_PROFILE_KINDS];
  if (bgpu_profile_end(ms, cnt, BGPU_PROFILE_KINDS) != 0) {
    std::printf("%s: profile_end failed: %s\n", tag, bgpu_last_error());
    return;
  }
  std::printf("%s:", tag);
  for (int k = 0; k < BGPU_PROFILE_KINDS; ++k)
    if (cnt[k]) {
      std::printf("  %s %.4f ms / %llu", bgpu_profile_kind_name(k), ms[k], (unsigned long long)cnt[k]);
      total += ms[k];
    }
  std::printf("  | kernels total %.4f ms\n", total);
  std::fflush(stdout);
}

int main(int argc, char **argv) {
  // NAME (variant = NAME=1 against NAME unset) or NAME=VALUE (variant = that assignment against NAME unset)
  std::string var_s = argc > 1 ? argv[1] : "BGPU_SHARE_X", val_s = "1";
  if (var_s.find('=') != std::string::npos) {
    val_s = var_s.substr(var_s.find('=') + 1);
    var_s = var_s.substr(0, var_s.find('='));
  }
  const char *var = var_s.c_str();
  const int N = argc > 2 ? std::atoi(argv[2]) : 256;
  const int calc_h = argc > 3 ? std::atoi(argv[3]) : 0;
  const int sfmodel = argc > 4 ? std::atoi(argv[4]) : 1;
  const int rsd = argc > 5 ? std::atoi(argv[5]) : (sfmodel == 1 ? 1 : 0);
  const int likelihood = argc > 6 ? std::atoi(argv[6]) : 1;
  const int masskernel = argc > 7 ? std::atoi(argv[7]) : 1;
  const double t0 = now();
  const size_t n = (size_t)N * N * N;
  bgpu_params p;
  bgpu_default_params(&p);
  p.N1 = p.N2 = p.N3 = N;
  p.L1 = p.L2 = p.L3 = N * (200.0 / 64.0);
  p.calc_h = calc_h;
  p.sfmodel = sfmodel;
  p.rsd_model = rsd;
  p.likelihood = likelihood;
  p.masskernel = masskernel;
  p.correct_delta = 1;
  p.deltaQ_factor = 1.0;
  bgpu_handle *H[2] = {nullptr, nullptr};
  unsetenv(var);
  CHECK(bgpu_create(&p, &H[0]));
  setenv(var, val_s.c_str(), 1);
  CHECK(bgpu_create(&p, &H[1]));
  unsetenv(var);
  std::printf("%s A/B at %d^3, calc_h %d, sfmodel %d, rsd %d, likelihood %d: handles up at %.2f s\n", var, N, calc_h,
              sfmodel, rsd, likelihood, now() - t0);

  std::vector<double> power(n), nobs(n), ones(n, 1.0), sig(n), grad[2];
  const double kf = 2.0 * M_PI / p.L1;
#pragma omp parallel for schedule(static)
  for (long idx = 0; idx < (long)n; ++idx) {
    const int k = (int)(idx % N), j = (int)((idx / N) % N), i = (int)(idx / ((size_t)N * N));
    auto kv = [&](int a) { return a <= N / 2 ? kf * a : -kf * (N - a); };
    const double kk = std::sqrt(kv(i) * kv(i) + kv(j) * kv(j) + kv(k) * kv(k));
    power[idx] = idx == 0 ? 0.0 : 2.2e4 * (kk / 0.02) / (1.0 + std::pow(kk / 0.02, 2.5));  // CDM-like, P(0) = 0
    nobs[idx] = 1.0 + 0.3 * unit_noise(idx, 1);
    sig[idx] = std::sqrt(3.0) * unit_noise(idx, 2);  // unit-variance white noise, coloured below
  }
  {
    // the signal: half a Gaussian-ish random field with that spectrum (create_GARFIELD's normalisation,
    // random.cpp:81-83: |s^|^2 = P N^2 / V), i.e. displacements of a few cells that vary smoothly -- what the
    // chain sees in production and what bench.py's synthetic problem is; white noise would not be
    const size_t nh = (size_t)N * N * (N / 2 + 1);
    std::vector<double> hat(2 * nh);
    CHECK(bgpu_fft_r2c(H[0], sig.data(), hat.data()));
    const double vol = p.L1 * p.L1 * p.L1;
#pragma omp parallel for schedule(static)
    for (long m = 0; m < (long)nh; ++m) {
      const int k = (int)(m % (N / 2 + 1)), j = (int)((m / (N / 2 + 1)) % N), i = (int)(m / ((size_t)(N / 2 + 1) * N));
      const double a = 0.5 * std::sqrt(power[((size_t)i * N + j) * N + k] * (double)n / vol);
      hat[2 * m] *= a;
      hat[2 * m + 1] *= a;
    }
    CHECK(bgpu_fft_c2r(H[0], hat.data(), sig.data()));
  }
  double e[2][2];
  for (int v = 0; v < 2; ++v) {
    grad[v].resize(n);
    CHECK(bgpu_set_static(H[v], power.data(), nobs.data(), ones.data(), ones.data()));
    CHECK(bgpu_gradient_psi(H[v], sig.data(), grad[v].data()));  // warm-up + the compared result
    CHECK(bgpu_psi(H[v], sig.data(), &e[v][0], &e[v][1], nullptr));
    CHECK(bgpu_profile_begin());
    CHECK(bgpu_gradient_psi(H[v], sig.data(), grad[v].data()));
    report(v ? "variant gradient_psi" : "default gradient_psi");
  }
  double num = 0, den = 0, amax = 0;
  for (size_t i = 0; i < n; ++i) {
    const double d = grad[1][i] - grad[0][i];
    num += d * d;
    den += grad[0][i] * grad[0][i];
    amax = std::fabs(grad[0][i]) > amax ? std::fabs(grad[0][i]) : amax;
  }
  std::printf("gradient: relative L2 difference %.3e (|grad|_2 %.6e, max %.3e)\n", std::sqrt(num / den), std::sqrt(den), amax);
  std::printf("psi_prior %.15e vs %.15e   psi_likeli %.15e vs %.15e\n", e[0][0], e[1][0], e[0][1], e[1][1]);
  bgpu_destroy(H[0]);
  bgpu_destroy(H[1]);
  std::printf("done at %.2f s\n", now() - t0);
  return 0;
}
