// barcode_b200/csrc/particle_math.cuh -- per-particle arithmetic shared by the scatter / gather kernels.
//
// Integer cell indices and interpolation weights follow the reference operation by operation
// (explicit round-to-nearest intrinsics where nvcc would otherwise contract a*b+c into an FMA,
// which the reference's x86-64 build never does), so positions, cell indices and weights are
// bit-identical for identical displacement input.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"

namespace bgpu {

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
// pacman_coordinate, pacman.cpp:20-28
__device__ __forceinline__ double pacman(double x, double L) {
  if (x < -L || x >= 2.0 * L) {  // several box lengths away: the reference's own sequence
    if (x < 0.) {
      x = fmod(x, L);
      x = __dadd_rn(x, L);
    }
    if (x >= L) x = fmod(x, L);
    return x;
  }
  // inside [-L, 2L) the same sequence without branches (seven calls per particle; the scatter is bound by
  // instruction issue and a fifth of it was branch bookkeeping): fmod(x, L) is x itself for -L <= x < 0, and
  // x - L, exactly (Sterbenz), for L <= x < 2L
  x = x < 0. ? __dadd_rn(x, L) : x;
  x = x >= L ? __dsub_rn(x, L) : x;
  return x;
}

// Lagrangian position + displacement (+ plane-parallel RSD), disp_part.cc:55-126, rsd.cc:30-64
__device__ __forceinline__ void particle_position(const GridGeom &g, int i, int j, int k, double px, double py,
                                                  double pz, double &x, double &y, double &z) {
  const double r = __dmul_rn(0.5, g.d);
  x = __dadd_rn(__dadd_rn(__dmul_rn(g.d, (double)i), r), px);
  y = __dadd_rn(__dadd_rn(__dmul_rn(g.d, (double)j), r), py);
  z = __dadd_rn(__dadd_rn(__dmul_rn(g.d, (double)k), r), pz);
  x = pacman(x, g.L);
  y = pacman(y, g.L);
  z = pacman(z, g.L);
  if (g.rsd) {
    const double vez = __dmul_rn(g.cpecvel, pz);       // Lag2Eul.cc:378-381
    const double ruxv = __dmul_rn(vez, g.v_norm);      // rsd.cc:52
    z = pacman(__dadd_rn(z, ruxv), g.L);               // rsd.cc:55,63
  }
}

// getCICcells + getCICweights for one coordinate, interpolate_grid.cpp:27-79.
// The reference computes the cell as (ULONG)(xpos/d), then (i + N) % N and (i + 1) % N in 64-bit
// integers; xpos is in [0, L) after the wrap, so the quotient is in [0, N] and the same values
// follow from a 32-bit conversion and conditional subtractions (a 64-bit `%` costs ~100 SASS
// instructions per call and dominated the first scatter kernel).
__device__ __forceinline__ void cic_axis(double x, double d, double L, int N, int &i0, int &i1, double &t,
                                         double &dx) {
  double xpos = __dsub_rn(x, __dmul_rn(0.5, d));
  xpos = pacman(xpos, L);
  const double q = __ddiv_rn(xpos, d);
  unsigned c = (unsigned)q;             // truncation, as the (ULONG) cast
  if (c >= (unsigned)N) c -= (unsigned)N;
  dx = __dsub_rn(q, (double)c);         // interpolate_grid.cpp:66-72 subtracts the WRAPPED cell index
  unsigned c1 = c + 1u;
  if (c1 >= (unsigned)N) c1 -= (unsigned)N;
  i0 = (int)c;
  i1 = (int)c1;
  t = __dsub_rn(1.0, dx);
}

// NGP / TSC centre cell, massFunctions.cc:72-79,198-204: floor((x - xmin)/d) folded into [0, N)
// with fmod on doubles in the reference; x is inside the domain, so the quotient is in [0, N]
__device__ __forceinline__ int ngp_axis(double x, double xmin, double d, int N) {
  unsigned c = (unsigned)floor(__ddiv_rn(__dsub_rn(x, xmin), d));
  if (c >= (unsigned)N) c -= (unsigned)N;
  return (int)c;
}

// TSC cells and weights for one coordinate, massFunctions.cc:198-235
__device__ __forceinline__ void tsc_axis(double x, double xmin, double d, int N, int (&c)[3], double (&w)[3],
                                         double &dx) {
  const unsigned i = (unsigned)ngp_axis(x, xmin, d, N);
  c[1] = (int)i;
  c[2] = (int)(i + 1u >= (unsigned)N ? i + 1u - (unsigned)N : i + 1u);
  c[0] = (int)(i == 0u ? (unsigned)N - 1u : i - 1u);
  const double xc = (double)i + 0.5;
  dx = __dsub_rn(__ddiv_rn(__dsub_rn(x, xmin), d), xc);
  w[1] = __dsub_rn(0.75, __dmul_rn(dx, dx));
  const double a = __dadd_rn(0.5, dx), b = __dsub_rn(0.5, dx);
  w[2] = __dmul_rn(__dmul_rn(0.5, a), a);
  w[0] = __dmul_rn(__dmul_rn(0.5, b), b);
}

// ---------------------------------------------------------------------------
// lean variants for the sweep kernels (particles_sweep.cu, which explains them): same bits as pacman / cic_axis for
// a coordinate within 3/4 of a box length of the box; anything else raises `slow` and the caller takes the general path
// ---------------------------------------------------------------------------
struct LeanConst {
  double d, L, half, rd, Lmid, far;
  unsigned N;
};

// pacman() for a coordinate in [-L, 2L): identical to its branch-free part; `slow` if the coordinate is not safely there
__device__ __forceinline__ double lean_wrap(double x, const LeanConst &c, bool &slow) {
  slow |= fabs(__dsub_rn(x, c.Lmid)) > c.far;
  if (x < 0.) x = __dadd_rn(x, c.L);
  if (x >= c.L) x = __dsub_rn(x, c.L);
  return x;
}

// cic_axis() (particle_math.cuh) for x in [0, L): cells c0, c1 and weights w0 = 1 - dx, w1 = dx
__device__ __forceinline__ void lean_axis(double x, const LeanConst &c, unsigned &c0, unsigned &c1, double &w0,
                                          double &w1, bool &slow) {
  double xpos = __dsub_rn(x, c.half);          // in [-d/2, L): pacman's in-range part
  if (xpos < 0.) xpos = __dadd_rn(xpos, c.L);
  if (xpos >= c.L) xpos = __dsub_rn(xpos, c.L);
  const double q0 = __dmul_rn(xpos, c.rd);     // RN(xpos / d), see above
  const double r = __fma_rn(-q0, c.d, xpos);
  const double q = __fma_rn(r, c.rd, q0);
  const double t = __dadd_rz(q, 4503599627370496.0);  // 2^52 + trunc(q): the cell index sits in the low word
  c0 = (unsigned)__double2loint(t);
  slow |= c0 >= c.N;                            // q rounded up to N: the reference wraps the cell and keeps q (dx = N)
  w1 = __dsub_rn(q, __dsub_rn(t, 4503599627370496.0));
  w0 = __dsub_rn(1.0, w1);
  c1 = (c0 + 1u) & (c.N - 1u);
}

__device__ __forceinline__ LeanConst lean_const(const GridGeom &g) {
  LeanConst lc;
  lc.d = g.d;
  lc.L = g.L;
  lc.half = __dmul_rn(0.5, g.d);
  lc.rd = __ddiv_rn(1.0, g.d);
  lc.Lmid = 0.5 * g.L;
  lc.far = 1.25 * g.L;
  lc.N = (unsigned)g.N;
  return lc;
}

__device__ __forceinline__ bool in_domain(const GridGeom &g, double x, double y, double z) {
  if (g.masskernel == 2)  // massFunctions.cc:195 (closed upper bound)
    return (x >= g.min1 && x <= g.min1 + g.L) && (y >= g.min2 && y <= g.min2 + g.L) &&
           (z >= g.min3 && z <= g.min3 + g.L);
  return (x >= g.min1 && x < g.min1 + g.L) && (y >= g.min2 && y < g.min2 + g.L) &&
         (z >= g.min3 && z < g.min3 + g.L);
}

}  // namespace bgpu
