// barcode_b200/csrc/particles_sph.cu -- SPH spline mass assignment and its exact adjoint over a STATIC hull.
//
// getDensity_SPH (/root/reference/barlib/src/massFunctions.cc:392-495) tests all (2 reach + 1)^3 cells around a
// particle (343 for h = one cell) with a square root and two divisions each; likelihood_calc_V_SPH
// (HMC_models.cc:200-303, SPH_kernel.cpp:62-208) walks the hull of cells the central cell can reach (81).  The
// first GPU version (kernels.cu) kept the per-particle pruning of those loops: every lane prunes differently, so a
// warp ran the union of its lanes' iterations with most lanes idle, and each surviving iteration paid a library
// rsqrt (slow-path branches) and two integer `%`.
//
// Here both directions walk ONE list of columns (i1, i2, half-range K) built on the host from the geometry alone --
// the cells whose centre can lie within 2h of SOME point of the particle's own cell -- so all 32 lanes execute
// the same iterations (no divergence; only the reduction / the accumulation is predicated), with
//   * q = q^2 * rsqrt(q^2) from MUFU.RSQ64H and one third-order correction (5 FP64 operations, relative error
//     < 1e-16 where the hardware seed is good to 2^-22; no special-case branches: q^2 = 0 is selected away),
//   * offsets in units of h formed once per particle, wraps by compare-and-add.
// The set of cells that receive mass is the reference's (the test q <= 2 is still made per cell); W differs by an
// ulp or two in q.  The sum a cell receives is order-free in the reference as well (OpenMP atomics).
#include <cmath>
#include <cstdlib>
#include <vector>

#include "kernels.h"
#include "particle_math.cuh"
#include "util.h"

namespace bgpu {
namespace {

constexpr int kMaxCols = 128;  // (2R + 1)^2 <= 121: h up to 2.5 cells; beyond that the general kernels run

// 1 / sqrt(x) for x > 0 (normal range): hardware seed (relative error 2^-22) and one Halley step,
// y (1 + e/2 + 3 e^2/8) with e = 1 - x y^2: error ~ e^3 = 2^-67, i.e. correctly rounded up to an ulp
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

struct Cols {
  int n;
  const int4 *tab;  // (i1, i2, K, unused)
};

__device__ __forceinline__ void load_cols(int4 *s, const Cols &c) {
  for (int t = threadIdx.x + threadIdx.y * blockDim.x; t < c.n; t += blockDim.x * blockDim.y) s[t] = c.tab[t];
  __syncthreads();
}

__device__ __forceinline__ int wrap(int c, int N) {
  c += c < 0 ? N : 0;
  c -= c >= N ? N : 0;
  return c;
}

// plane of the (halo-extended, slab-local) tile that global cell plane c maps to (a cube: c itself); >= np = outside
__device__ __forceinline__ int tile_plane(int c, int N, int lo) {
  int kx = wrap(c, N) - lo;
  kx += kx < 0 ? N : 0;
  kx -= kx >= N ? N : 0;
  return kx;
}

__device__ __forceinline__ void deposit_cols(const GridGeom &g, const int4 *cols, int ncol, double x, double y, double z,
                                             double *__restrict__ rho) {
  const int N = g.N;
  if (!in_domain(g, x, y, z)) return;
  const double d = g.d, h = g.sph_h, h_inv = 1. / h, d_h = d * h_inv;
  const double a = 1. / M_PI / (h * h * h);
  const int ix = (int)(unsigned long long)(x / d), iy = (int)(unsigned long long)(y / d),
            iz = (int)(unsigned long long)(z / d);
  // particle minus the centre of its own cell, in units of h
  const double ax = (x - ((double)ix + 0.5) * d) * h_inv, ay = (y - ((double)iy + 0.5) * d) * h_inv,
               az = (z - ((double)iz + 0.5) * d) * h_inv;
  const int lo = g.x0 - g.H, np = g.Ns + 2 * g.H;
  for (int c = 0; c < ncol; ++c) {
    const int4 col = cols[c];
    const int kx = tile_plane(ix + col.x, N, lo);
    if (kx >= np) {  // beyond the halo: drop the deposit and raise the flag (the host turns it into an error)
      if (g.flag) *g.flag = 1;
      continue;
    }
    const double dx = ax - (double)col.x * d_h, dy = ay - (double)col.y * d_h;
    const double qxy = fma(dy, dy, dx * dx);
    double *row = rho + ((size_t)kx * N + wrap(iy + col.y, N)) * N;
    const int K = col.z;
    double dz = fma((double)K, d_h, az);  // i3 = -K first; stepping by d_h instead of converting i3 every cell
    int kz = wrap(iz - K, N);
    for (int i3 = -K; i3 <= K; ++i3, dz -= d_h, kz = (kz + 1 == N) ? 0 : kz + 1) {
      const double q2 = fma(dz, dz, qxy);
      const double q = q2 > 0. ? q2 * rsqrt_fast(q2) : 0.;
      // Monaghan W_4 spline, SPH_kernel_3D (massFunctions.cc:366-384)
      const double t = 2. - q;
      const double w_in = fma(q2, fma(0.75, q, -1.5), 1.0);
      const double w_out = 0.25 * t * (t * t);
      const double w = a * (q <= 1. ? w_in : w_out);
      if (q <= 2.) atomicAdd(row + kz, w);
    }
  }
}

__global__ void __launch_bounds__(128) scatter_sph_cols_kernel(GridGeom g, const double *__restrict__ psix,
                                                               const double *__restrict__ psiy,
                                                               const double *__restrict__ psiz, double *__restrict__ rho,
                                                               double *__restrict__ posx, double *__restrict__ posy,
                                                               double *__restrict__ posz, Cols cols) {
  __shared__ int4 s_cols[kMaxCols];
  load_cols(s_cols, cols);
  const int N = g.N;
  const size_t n = (size_t)g.Ns * N * N;  // the Lagrangian planes this rank owns (a cube: all of them)
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);  // N is a power of two
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)),
            i = g.x0 + (int)(idx >> (2 * sh));
  double x, y, z;
  particle_position(g, i, j, k, psix[idx], psiy[idx], psiz[idx], x, y, z);
  if (posx) {
    posx[idx] = x;
    posy[idx] = y;
    posz[idx] = z;
  }
  deposit_cols(g, s_cols, cols.n, x, y, z, rho);
}

__global__ void __launch_bounds__(128) scatter_sph_cols_positions_kernel(GridGeom g, const double *__restrict__ x,
                                                                         const double *__restrict__ y,
                                                                         const double *__restrict__ z,
                                                                         double *__restrict__ rho, Cols cols) {
  __shared__ int4 s_cols[kMaxCols];
  load_cols(s_cols, cols);
  const size_t n = (size_t)g.N * g.N * g.N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) deposit_cols(g, s_cols, cols.n, x[idx], y[idx], z[idx], rho);
}

// V_p = (rho_c V/N) sum_{hull cells} r_c gradW((x_p - x_c)/h), gradW = partial(q) (x_p - x_c)/h / (pi h^4)
// (SPH_kernel.cpp:148-208).  In place over Psi.  Deterministic (pure gather).  A CTA is 32 k x 8 j neighbouring
// particles, so that the ~12 residual rows x 5 planes its hulls cover are read from HBM / L2 once and from L1 after.
__global__ void __launch_bounds__(256) gather_sph_cols_kernel(GridGeom g, double *__restrict__ ax_, double *__restrict__ ay_,
                                                              double *__restrict__ az_, const double *__restrict__ resid,
                                                              Cols cols, double normalize) {
  __shared__ int4 s_cols[kMaxCols];
  load_cols(s_cols, cols);
  const int N = g.N;
  const int k = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y, il = blockIdx.z;
  if (k >= N || j >= N) return;
  const int i = g.x0 + il;
  const size_t idx = ((size_t)il * N + j) * N + k;
  // the residual is [(Ns + 2H)][N][N] with plane 0 = global x0 - H (a cube: the whole grid)
  const int lo = g.x0 - g.H, np = g.Ns + 2 * g.H;
  double px, py, pz;
  particle_position(g, i, j, k, ax_[idx], ay_[idx], az_[idx], px, py, pz);
  const double d = g.d, h = g.sph_h, h_inv = 1. / h, d_h = d * h_inv;
  const double norm = 1. / (M_PI * (h * h) * (h * h));
  const int ix = (int)(px / d), iy = (int)(py / d), iz = (int)(pz / d);
  const double dpcx = px * h_inv - ((double)ix + 0.5) * d_h, dpcy = py * h_inv - ((double)iy + 0.5) * d_h,
               dpcz = pz * h_inv - ((double)iz + 0.5) * d_h;
  double vx = 0., vy = 0., vz = 0.;
  const int ncol = cols.n;
  for (int c = 0; c < ncol; ++c) {
    const int4 col = s_cols[c];
    const int kx = tile_plane(ix + col.x, N, lo);
    if (kx >= np) continue;  // beyond the halo: the scatter of the same evaluation has raised the flag already
    const double dxh = dpcx - (double)col.x * d_h, dyh = dpcy - (double)col.y * d_h;
    const double qxy = fma(dyh, dyh, dxh * dxh);
    const double *row = resid + ((size_t)kx * N + wrap(iy + col.y, N)) * N;
    const int K = col.z;
    double sx = 0., sz = 0.;  // sum of c over the column, and of c * dzh: dxh, dyh are constant along it
    double dzh = fma((double)K, d_h, dpcz);  // i3 = -K first
    int kz = wrap(iz - K, N);
    for (int i3 = -K; i3 <= K; ++i3, dzh -= d_h, kz = (kz + 1 == N) ? 0 : kz + 1) {
      const double r = __ldg(row + kz);
      const double q2 = fma(dzh, dzh, qxy);
      const double y = rsqrt_fast(q2);  // q2 = 0: inf / NaN, selected away below (the inner branch has no 1/q)
      const double q = q2 > 0. ? q2 * y : 0.;
      const double qm = q - 2.;
      const double p_out = (-0.75 * y) * (qm * qm);   // -3/4 (q - 2)^2 / q
      const double p_in = fma(2.25, q, -3.);          // (9/4 q - 3)
      double partial = q2 > 1. ? p_out : p_in;
      partial = q2 > 4. ? 0. : partial;
      const double cc = r * partial;
      sx += cc;
      sz = fma(cc, dzh, sz);
    }
    vx = fma(sx, dxh, vx);
    vy = fma(sx, dyh, vy);
    vz += sz;
  }
  const double f = normalize * norm;
  vx *= f;
  vy *= f;
  vz *= f;
  if (g.rsd) vz += g.fgrow * vz;  // HMC_models.cc:295-301
  ax_[idx] = vx;
  ay_[idx] = vy;
  az_[idx] = vz;
}


// ---------------------------------------------------------------------------
// Half-ranges K <= 2 (h up to 1.25 cells; the shipped default h = d has 21 columns, 81 cells): the z loop unrolled.
//
// ncu of the list kernels above (profiles/ncu_full_r02_sph_256.txt): instruction issue 75-83 % busy, FP64 pipe
// 53-57 %; of 4684 warp instructions per 32 particles, 110 go to every COLUMN (two-stage wraps, 64-bit row
// addresses, int -> double offsets) and 35 to every cell (of which the z step, its wrap and the loop are 8).  Here
//   * everything along z is formed once per particle: five wrapped cell indices, five offsets, five squares;
//     the column's cells are five unrolled bodies behind uniform tests of K, no stepping;
//   * the column list in shared memory carries the offsets as doubles; the x wrap and dx^2 are redone only when
//     i1 changes (the list is sorted by i1; a uniform branch); row offsets are 32-bit (N^3 <= 2^30 per tile);
//   * q^2 >= 1 and q^2 >= 4 are tested on the high word of q^2 -- they differ from q > 1, q > 2 only AT q = 1 and
//     q = 2, where the spline's two branches agree in value and slope / the kernel is zero -- and q^2 = 0 is kept
//     finite by raising the high word fed to the rsqrt seed to the smallest normal (q = 0 * y = 0).
// (A fully unrolled 5 x 5 x 5 variant was measured: scatter 2.59 ms, gather 4.43 ms against 2.86 / 2.84 for the
// list kernels at 256^3 -- 104 registers and ~4000 instructions of straight-line code; dropped.)
// ---------------------------------------------------------------------------
struct ColD {
  double ox, oy;  // i1 d / h, i2 d / h
  int i1, i2, K, pad;
};

__device__ __forceinline__ void load_cols5(ColD *s, const Cols &c, double d_h) {
  for (int t = threadIdx.x + threadIdx.y * blockDim.x; t < c.n; t += blockDim.x * blockDim.y) {
    const int4 v = c.tab[t];
    s[t] = ColD{(double)v.x * d_h, (double)v.y * d_h, v.x, v.y, v.z, 0};
  }
  __syncthreads();
}

__device__ __forceinline__ double rsqrt_seeded(double x, int hi) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(max(hi, 0x00100000), 0)));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

struct ZAxis {
  unsigned kz[5];
  double dz[5], dz2[5];
};

__device__ __forceinline__ void z_axis(int iz, double az, double d_h, int N, ZAxis &zz) {
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    zz.kz[c] = (unsigned)wrap(iz + c - 2, N);
    zz.dz[c] = az - (double)(c - 2) * d_h;
    zz.dz2[c] = zz.dz[c] * zz.dz[c];
  }
}

__device__ __forceinline__ void deposit_cols5(const GridGeom &g, const ColD *cols, int ncol, double x, double y, double z,
                                              double *__restrict__ rho) {
  const int N = g.N;
  if (!in_domain(g, x, y, z)) return;
  const double d = g.d, h = g.sph_h, h_inv = 1. / h, d_h = d * h_inv;
  const double a0 = 1. / M_PI / (h * h * h);
  const int ix = (int)(unsigned long long)(x / d), iy = (int)(unsigned long long)(y / d),
            iz = (int)(unsigned long long)(z / d);
  const double ax = (x - ((double)ix + 0.5) * d) * h_inv, ay = (y - ((double)iy + 0.5) * d) * h_inv,
               az = (z - ((double)iz + 0.5) * d) * h_inv;
  const int lo = g.x0 - g.H, np = g.Ns + 2 * g.H;
  ZAxis zz;
  z_axis(iz, az, d_h, N, zz);
  int cur = 0x7fffffff, kx = 0;
  double dx2 = 0.;
  for (int c = 0; c < ncol; ++c) {
    const ColD col = cols[c];
    if (col.i1 != cur) {  // uniform: all lanes walk the same list
      cur = col.i1;
      kx = tile_plane(ix + cur, N, lo);
      const double dx = ax - col.ox;
      dx2 = dx * dx;
    }
    if (kx >= np) {  // beyond the halo: drop the deposit and raise the flag (the host turns it into an error)
      if (g.flag) *g.flag = 1;
      continue;
    }
    const double dy = ay - col.oy;
    const double qxy = fma(dy, dy, dx2);
    double *row = rho + ((unsigned)kx * (unsigned)N + (unsigned)wrap(iy + col.i2, N)) * (unsigned)N;
#pragma unroll
    for (int cz = 0; cz < 5; ++cz) {
      if (col.K < (cz < 2 ? 2 - cz : cz - 2)) continue;
      const double q2 = qxy + zz.dz2[cz];
      const int hi = __double2hiint(q2);
      const double q = q2 * rsqrt_seeded(q2, hi);
      // Monaghan W_4 spline, SPH_kernel_3D (massFunctions.cc:366-384): a (1 - 3/2 q^2 + 3/4 q^3) inside q = 1,
      // a (2 - q)^3 / 4 outside
      const double t = 2. - q;
      double w = fma(q2, fma(0.75 * a0, q, -1.5 * a0), a0);
      if (hi >= 0x3ff00000) w = (0.25 * a0 * t) * (t * t);
      if (hi < 0x40100000) atomicAdd(row + zz.kz[cz], w);
    }
  }
}

__global__ void __launch_bounds__(128) scatter_sph_cols5_kernel(GridGeom g, const double *__restrict__ psix,
                                                                const double *__restrict__ psiy,
                                                                const double *__restrict__ psiz, double *__restrict__ rho,
                                                                double *__restrict__ posx, double *__restrict__ posy,
                                                                double *__restrict__ posz, Cols cols) {
  __shared__ ColD s_cols[25];
  load_cols5(s_cols, cols, g.d / g.sph_h);
  const int N = g.N;
  const size_t n = (size_t)g.Ns * N * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int sh = 31 - __clz(N);  // N is a power of two
  const int k = (int)(idx & (size_t)(N - 1)), j = (int)((idx >> sh) & (size_t)(N - 1)),
            i = g.x0 + (int)(idx >> (2 * sh));
  double x, y, z;
  particle_position(g, i, j, k, psix[idx], psiy[idx], psiz[idx], x, y, z);
  if (posx) {
    posx[idx] = x;
    posy[idx] = y;
    posz[idx] = z;
  }
  deposit_cols5(g, s_cols, cols.n, x, y, z, rho);
}

__global__ void __launch_bounds__(128) scatter_sph_cols5_positions_kernel(GridGeom g, const double *__restrict__ x,
                                                                          const double *__restrict__ y,
                                                                          const double *__restrict__ z,
                                                                          double *__restrict__ rho, Cols cols) {
  __shared__ ColD s_cols[25];
  load_cols5(s_cols, cols, g.d / g.sph_h);
  const size_t n = (size_t)g.N * g.N * g.N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) deposit_cols5(g, s_cols, cols.n, x[idx], y[idx], z[idx], rho);
}

__global__ void __launch_bounds__(256) gather_sph_cols5_kernel(GridGeom g, double *__restrict__ ax_, double *__restrict__ ay_,
                                                               double *__restrict__ az_, const double *__restrict__ resid,
                                                               Cols cols, double normalize) {
  __shared__ ColD s_cols[25];
  const double d = g.d, h = g.sph_h, h_inv = 1. / h, d_h = d * h_inv;
  load_cols5(s_cols, cols, d_h);
  const int N = g.N;
  const int k = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y, il = blockIdx.z;
  if (k >= N || j >= N) return;
  const size_t idx = ((size_t)il * N + j) * N + k;
  const int lo = g.x0 - g.H, np = g.Ns + 2 * g.H;
  double px, py, pz;
  particle_position(g, g.x0 + il, j, k, ax_[idx], ay_[idx], az_[idx], px, py, pz);
  const int ix = (int)(px / d), iy = (int)(py / d), iz = (int)(pz / d);
  const double dpcx = px * h_inv - ((double)ix + 0.5) * d_h, dpcy = py * h_inv - ((double)iy + 0.5) * d_h,
               dpcz = pz * h_inv - ((double)iz + 0.5) * d_h;
  ZAxis zz;
  z_axis(iz, dpcz, d_h, N, zz);
  double vx = 0., vy = 0., vz = 0.;
  int cur = 0x7fffffff, kx = 0;
  double dxh = 0., dx2 = 0.;
  const int ncol = cols.n;
  for (int c = 0; c < ncol; ++c) {
    const ColD col = s_cols[c];
    if (col.i1 != cur) {  // uniform
      cur = col.i1;
      kx = tile_plane(ix + cur, N, lo);
      dxh = dpcx - col.ox;
      dx2 = dxh * dxh;
    }
    if (kx >= np) continue;  // beyond the halo: the scatter of the same evaluation has raised the flag already
    const double dyh = dpcy - col.oy;
    const double qxy = fma(dyh, dyh, dx2);
    const double *row = resid + ((unsigned)kx * (unsigned)N + (unsigned)wrap(iy + col.i2, N)) * (unsigned)N;
    double sx = 0., sz = 0.;  // sum of r partial over the column, and of r partial dz: dx, dy are constant along it
#pragma unroll
    for (int cz = 0; cz < 5; ++cz) {
      if (col.K < (cz < 2 ? 2 - cz : cz - 2)) continue;
      const double r = __ldg(row + zz.kz[cz]);
      const double q2 = qxy + zz.dz2[cz];
      const int hi = __double2hiint(q2);
      const double y = rsqrt_seeded(q2, hi);
      const double q = q2 * y;
      // partial(q) of SPH_kernel.cpp:148-208 without its 1 / (pi h^4): 9/4 q - 3 inside q = 1, -3/4 (q - 2)^2 / q outside
      const double qm = q - 2.;
      double partial = fma(2.25, q, -3.);
      if (hi >= 0x3ff00000) partial = (-0.75 * y) * (qm * qm);
      const double cc = hi < 0x40100000 ? r * partial : 0.;
      sx += cc;
      sz = fma(cc, zz.dz[cz], sz);
    }
    vx = fma(sx, dxh, vx);
    vy = fma(sx, dyh, vy);
    vz += sz;
  }
  const double f = normalize / (M_PI * (h * h) * (h * h));
  vx *= f;
  vy *= f;
  vz *= f;
  if (g.rsd) vz += g.fgrow * vz;  // HMC_models.cc:295-301
  ax_[idx] = vx;
  ay_[idx] = vy;
  az_[idx] = vz;
}

}  // namespace

// ---------------------------------------------------------------------------
// host side: the column lists
// ---------------------------------------------------------------------------
// Scatter: every (i1, i2, i3) whose cell centre can be within 2h of a point of the central cell -- the distance
// along an axis between a point of cell 0 and the centre of cell i is at least max(0, |i| - 1/2) d.  (A superset of
// the cells a given particle reaches; the kernel still tests q <= 2.)  Gather: the reference's own hull table.
SphColumns *sph_columns_create(const GridGeom &g, const int *kmax_host, int R) {
  const int W = 2 * R + 1;
  std::vector<int4> sc, ga;
  const double lim = 4. * g.sph_h * g.sph_h * (1. + 1e-9);
  for (int i1 = -R; i1 <= R; ++i1)
    for (int i2 = -R; i2 <= R; ++i2) {
      int K = -1;
      for (int i3 = 0; i3 <= R; ++i3) {
        const double dx = std::fmax(0., std::abs((double)i1) - 0.5) * g.d, dy = std::fmax(0., std::abs((double)i2) - 0.5) * g.d,
                     dz = std::fmax(0., std::abs((double)i3) - 0.5) * g.d;
        if (dx * dx + dy * dy + dz * dz <= lim) K = i3;
      }
      if (K >= 0) sc.push_back(make_int4(i1, i2, K, 0));
      const int Kg = kmax_host[(size_t)(i1 + R) * W + (i2 + R)];
      if (Kg >= 0) ga.push_back(make_int4(i1, i2, Kg, 0));
    }
  if ((int)sc.size() > kMaxCols || (int)ga.size() > kMaxCols) return nullptr;  // the general kernels take over
  auto *out = new SphColumns;
  // every half-range <= 2 and at most 25 columns: the kernels with the unrolled z loop (BGPU_SPH_Z5=0: the list kernels)
  {
    bool fits = sc.size() <= 25 && ga.size() <= 25;
    for (const int4 &c : sc) fits = fits && c.z <= 2;
    for (const int4 &c : ga) fits = fits && c.z <= 2;
    const char *e = std::getenv("BGPU_SPH_Z5");
    out->z5 = fits && !(e && e[0] == '0');
  }
  out->n_scatter = (int)sc.size();
  out->n_gather = (int)ga.size();
  BGPU_CUDA(cudaMalloc(&out->dev, (sc.size() + ga.size()) * sizeof(int4)));
  BGPU_CUDA(cudaMemcpy(out->dev, sc.data(), sc.size() * sizeof(int4), cudaMemcpyHostToDevice));
  BGPU_CUDA(cudaMemcpy(static_cast<int4 *>(out->dev) + sc.size(), ga.data(), ga.size() * sizeof(int4), cudaMemcpyHostToDevice));
  return out;
}

void sph_columns_destroy(SphColumns *c) {
  if (!c) return;
  if (c->dev) cudaFree(c->dev);
  delete c;
}

void launch_scatter_sph_cols(const GridGeom &g, const SphColumns *c, const double *psix, const double *psiy,
                             const double *psiz, double *rho, double *posx, double *posy, double *posz, cudaStream_t st) {
  ProfScope prof(KK_SCATTER, st);
  const size_t n = (size_t)g.Ns * g.N * g.N;
  BGPU_CUDA(cudaMemsetAsync(rho, 0, (size_t)(g.Ns + 2 * g.H) * g.N * g.N * sizeof(double), st));
  Cols cols{c->n_scatter, static_cast<const int4 *>(c->dev)};
  if (c->z5) {
    scatter_sph_cols5_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, psix, psiy, psiz, rho, posx, posy, posz, cols);
    BGPU_LAUNCHED(1);
    return;
  }
  scatter_sph_cols_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, psix, psiy, psiz, rho, posx, posy, posz, cols);
  BGPU_LAUNCHED(1);
}

void launch_scatter_sph_cols_positions(const GridGeom &g, const SphColumns *c, const double *x, const double *y,
                                       const double *z, double *rho, cudaStream_t st) {
  const size_t n = (size_t)g.N * g.N * g.N;
  Cols cols{c->n_scatter, static_cast<const int4 *>(c->dev)};
  if (c->z5) {
    scatter_sph_cols5_positions_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, x, y, z, rho, cols);
    BGPU_LAUNCHED(1);
    return;
  }
  scatter_sph_cols_positions_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, x, y, z, rho, cols);
  BGPU_LAUNCHED(1);
}

void launch_gather_sph_cols(const GridGeom &g, const SphColumns *c, double *ax, double *ay, double *az, const double *resid,
                            double normalize, cudaStream_t st) {
  ProfScope prof(KK_GATHER, st);
  Cols cols{c->n_gather, static_cast<const int4 *>(c->dev) + c->n_scatter};
  const dim3 grid((unsigned)((g.N + 31) / 32), (unsigned)((g.N + 7) / 8), (unsigned)g.Ns), block(32, 8);
  if (c->z5) {
    gather_sph_cols5_kernel<<<grid, block, 0, st>>>(g, ax, ay, az, resid, cols, normalize);
    BGPU_LAUNCHED(1);
    return;
  }
  gather_sph_cols_kernel<<<grid, block, 0, st>>>(g, ax, ay, az, resid, cols, normalize);
  BGPU_LAUNCHED(1);
}

}  // namespace bgpu
