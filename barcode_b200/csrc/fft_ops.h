// barcode_b200/csrc/fft_ops.h -- functor descriptors fused into the FFT passes.
#pragma once
#include <cuda_runtime.h>

namespace bgpu {

// ---------------------------------------------------------------------------
// functor descriptors (plain structs: passed by value as kernel arguments)
// ---------------------------------------------------------------------------
enum KKind : int {
  K_NONE = 0,
  K_DISP = 1,        // a * (k_c/k^2) * (Im v, -Re v); 0 if k^2 <= 1e-14 or on a Nyquist plane (EqSolvers.cc:208-268)
  K_GRAD = 2,        // k_c * (-Im v, Re v); 0 on a Nyquist plane (gradient.cpp:38-74)
  K_MULREAL = 3,     // v * real0[off]  (HMC_help.cc:41-58 with the multiplier precomputed)
  K_FINAL = 4,       // v * real0[off] + a * cplx0[off]   (prior + norm * h, HMC.cc:205)
  K_INVLAP_SET = 5,  // store: out[off]  = (k_c/k^2)(Im v, -Re v); k^2 == 0 -> 0, Nyquist -> 0 (gradient.cpp:167-210)
  K_INVLAP_ADD = 6,  // store: out[off] += same
  K_NEGINVK2 = 7,    // a * v * (-1/k^2), 0 at k^2 == 0, no Nyquist zeroing (PoissonSolver, EqSolvers.cc:29-64)
  K_GAUSS = 8,       // v * K,       K = exp(-k^2 a^2 / 2), a = smoothing radius (kernelcomp filtertype 1, convolution.cpp:224-322)
  K_ONE_MINUS_GAUSS = 9, // v * (1 - K)
  // Shared x pass (BGPU_SHARE_X=1; understood by the extended strided kernels only, fft_tma.cuh RotCtxX).  The three
  // components of a K_DISP / K_GRAD / K_INVLAP triple differ by the REAL factor k_c, and k_y, k_z are constant along
  // an x pencil, so the y and z components can share one x pass run with comp = K_COMP_UNIT (k_c := 1, same
  // zeroing rules) and pick up k_y / k_z in their y passes:
  K_MULK = 10,       // load:  v * k_c                  (comp 1 or 2)
  K_MULK_SET = 11,   // store: out[off]  = v * k_c
  K_MULK_ADD = 12    // store: out[off] += v * k_c
};
constexpr int K_COMP_UNIT = 3;  // KOp::comp value: k_c := 1

struct KOp {
  int kind = K_NONE;
  int comp = 0;                  // 0,1,2 -> k_x, k_y, k_z
  double a = 1.0;
  double kfac = 0.0;             // 2 pi / L
  const double *real0 = nullptr;
  const double2 *cplx0 = nullptr;
};

enum RKind : int {
  R_SCALE = 0,      // out = a * v
  R_SCALE_MUL = 1,  // out = a * v * aux[idx]
  R_AXPY = 2,       // out += a * v
  R_LOAD = 10,      // r2c load: in[idx]
  R_LOAD_SCALE = 11 // r2c load: a * in[idx]
};

struct ROp {
  int kind = R_SCALE;
  double a = 1.0;
  const double *aux = nullptr;
  // c2r store only: when non-null and *skip != 0 the z pass stores nothing at all (a leapfrog trajectory that has
  // been stopped by its run-away guard leaves the state where it was, HMC.cc:360-364)
  const int *skip = nullptr;
};

}  // namespace bgpu
