// barcode_b200/csrc/particles_sweep.cu -- CIC / TSC mass assignment and its exact adjoint as an x sweep.
//
// getDensity_CIC / getDensity_TSC (/root/reference/barlib/src/massFunctions.cc:100-364) deposit every particle
// into 8 / 27 cells with one `#pragma omp atomic` each.  On B200 the bound of a particle-per-thread scatter is
// the number of global reductions (RED.E.ADD.F64) the SM's load/store unit can issue -- ~1.3 cycles per lane-op --
// and the instructions spent around them, not HBM.  Both are cut here by walking the Lagrangian lattice in the
// two directions in which neighbouring particles share cells:
//   * along z the 32 lanes of a warp hold the particles (i, j, k .. k+31): the upper z cell of lane l is the
//     lower z cell of lane l+1, handed over with one shuffle (as in the first-generation kernel, kernels.cu);
//   * along x every thread SWEEPS a segment of `seg` planes i = i0 .. i0+seg-1 for its fixed (j, k): the cells of
//     the particle's upper x plane(s) are the lower x plane(s) of the next particle of the sweep, so their
//     contributions stay in registers and only the plane that no later particle of the sweep can reach is
//     reduced into global memory.
// Wherever the displacement field is not that regular (shell crossing, wrap-around) the ADDRESSES do not match
// and the carried values are flushed on their own: any field is handled exactly, only slower.
// CIC goes from 8 (reference) / 4.1 (z hand-over only) to ~2.1 reductions per particle, TSC from 27 / 9.6 to ~3.2.
// The sum a cell receives is the same set of products w_x w_y w_z (computed operation by operation as the
// reference does, particle_math.cuh) in a different order -- the reference's own OpenMP order is not defined either.
//
// The adjoint (gather) uses the same walk: residual values of the upper x plane are carried to the next particle,
// the upper z value comes from the neighbouring lane, so a CIC particle issues ~2 loads instead of 8.
#include <cstdlib>

#include "kernels.h"
#include "particle_math.cuh"
#include "util.h"

namespace bgpu {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned kNoAddr = 0xfffffff0u;  // never a cell offset (N^3 <= 2^30); +lane bits keep neighbours distinct

// reduce (alo, lo) and (ahi, hi) = a particle's two z cells of one (x, y) row; the upper one is handed to the lane
// above when that lane's lower cell is the same address.  Every lane of the warp must call (invalid: valid = false).
__device__ __forceinline__ void red_zpair(double *__restrict__ rho, unsigned alo, unsigned ahi, double lo, double hi,
                                          bool valid, int lane) {
  if (!valid) {
    alo = kNoAddr + 1u;
    ahi = kNoAddr + 2u;
  }
  const unsigned p_ahi = __shfl_up_sync(FULL, ahi, 1), n_alo = __shfl_down_sync(FULL, alo, 1);
  const double p_hi = __shfl_up_sync(FULL, hi, 1);
  if (lane > 0 && p_ahi == alo) lo += p_hi;
  const bool handed_up = lane < 31 && n_alo == ahi;
  if (valid) {
    atomicAdd(rho + alo, lo);
    if (!handed_up) atomicAdd(rho + ahi, hi);
  }
}

// the three z cells of one row of a TSC particle: centre lane l = lower cell of lane l+1 = upper cell of lane l-1
__device__ __forceinline__ void red_ztriple(double *__restrict__ rho, unsigned am, unsigned a0, unsigned ap, double vm,
                                            double v0, double vp, bool valid, int lane) {
  if (!valid) {
    am = kNoAddr + 1u;
    a0 = kNoAddr + 2u;
    ap = kNoAddr + 3u;
  }
  const unsigned p_a0 = __shfl_up_sync(FULL, a0, 1), p_ap = __shfl_up_sync(FULL, ap, 1);
  const unsigned n_a0 = __shfl_down_sync(FULL, a0, 1), n_am = __shfl_down_sync(FULL, am, 1);
  const double p_vp = __shfl_up_sync(FULL, vp, 1), n_vm = __shfl_down_sync(FULL, vm, 1);
  const bool left = lane > 0 && p_a0 == am && p_ap == a0;
  const bool right = lane < 31 && n_a0 == ap && n_am == a0;
  if (left) v0 += p_vp;
  if (right) v0 += n_vm;
  if (valid) {
    atomicAdd(rho + a0, v0);
    if (!left) atomicAdd(rho + am, vm);
    if (!right) atomicAdd(rho + ap, vp);
  }
}

struct SweepIdx {
  int lane, j, k, i_begin;
  size_t idx;  // element (i_begin, j, k)
  bool live;
};

// warp w of the grid -> (segment, j, 32 consecutive k): lanes run along z, the contiguous axis
__device__ __forceinline__ SweepIdx sweep_index(int N, int seg) {
  SweepIdx s;
  s.lane = threadIdx.x & 31;
  const int sh = 31 - __clz(N);
  const unsigned w = (unsigned)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int kb = (int)(w & (unsigned)((N >> 5) - 1));
  s.j = (int)((w >> (sh - 5)) & (unsigned)(N - 1));
  const int sg = (int)(w >> (2 * sh - 5));
  s.k = kb * 32 + s.lane;
  s.i_begin = sg * seg;
  s.live = s.i_begin < N;
  s.idx = ((size_t)s.i_begin * N + s.j) * N + s.k;
  return s;
}

// Lagrangian position + displacement (+ RSD) with the y, z lattice coordinates hoisted (particle_position)
template <bool RSD>
__device__ __forceinline__ void sweep_position(const GridGeom &g, int i, double half, double qy, double qz, double px,
                                               double py, double pz, double &x, double &y, double &z) {
  x = pacman(__dadd_rn(__dadd_rn(__dmul_rn(g.d, (double)i), half), px), g.L);
  y = pacman(__dadd_rn(qy, py), g.L);
  z = pacman(__dadd_rn(qz, pz), g.L);
  if constexpr (RSD) {
    const double vez = __dmul_rn(g.cpecvel, pz);   // Lag2Eul.cc:378-381
    const double ruxv = __dmul_rn(vez, g.v_norm);  // rsd.cc:52
    z = pacman(__dadd_rn(z, ruxv), g.L);           // rsd.cc:55,63
  }
}

// ---------------------------------------------------------------------------
// CIC scatter
// ---------------------------------------------------------------------------
template <bool RSD>
__global__ void __launch_bounds__(256) scatter_cic_sweep_kernel(GridGeom g, const double *__restrict__ psix,
                                                                const double *__restrict__ psiy,
                                                                const double *__restrict__ psiz,
                                                                double *__restrict__ rho, int seg) {
  const int N = g.N;
  const SweepIdx s = sweep_index(N, seg);
  if (!s.live) return;  // whole warps only
  const int lane = s.lane;
  const size_t pl = (size_t)N * N;
  const double half = __dmul_rn(0.5, g.d);
  const double qy = __dadd_rn(__dmul_rn(g.d, (double)s.j), half), qz = __dadd_rn(__dmul_rn(g.d, (double)s.k), half);
  size_t idx = s.idx;
  // carried: the particle's upper x plane, 2 x 2 (y, z) cells
  double c00 = 0., c01 = 0., c10 = 0., c11 = 0.;
  unsigned ca00 = kNoAddr, ca01 = kNoAddr, ca10 = kNoAddr, ca11 = kNoAddr;
  bool cvalid = false;
  double px = psix[idx], py = psiy[idx], pz = psiz[idx];
  const int i_end = s.i_begin + seg;
  for (int i = s.i_begin; i < i_end; ++i, idx += pl) {
    double nx = 0., ny = 0., nz = 0.;
    if (i + 1 < i_end) {  // the next plane's displacement is in flight while this one is deposited
      nx = psix[idx + pl];
      ny = psiy[idx + pl];
      nz = psiz[idx + pl];
    }
    double x, y, z;
    sweep_position<RSD>(g, i, half, qy, qz, px, py, pz, x, y, z);
    const bool valid = in_domain(g, x, y, z);
    int ci0 = 0, ci1 = 0, cj0 = 0, cj1 = 0, ck0 = 0, ck1 = 0;
    double wi0 = 0., wi1 = 0., wj0 = 0., wj1 = 0., wk0 = 0., wk1 = 0.;
    if (valid) {
      cic_axis(x, g.d, g.L, N, ci0, ci1, wi0, wi1);
      cic_axis(y, g.d, g.L, N, cj0, cj1, wj0, wj1);
      cic_axis(z, g.d, g.L, N, ck0, ck1, wk0, wk1);
    }
    const unsigned r00 = ((unsigned)ci0 * N + cj0) * N, r01 = ((unsigned)ci0 * N + cj1) * N;
    const unsigned r10 = ((unsigned)ci1 * N + cj0) * N, r11 = ((unsigned)ci1 * N + cj1) * N;
    // mass * w_x * w_y * w_z evaluated left to right, massFunctions.cc:129-157
    const double w00 = __dmul_rn(wi0, wj0), w01 = __dmul_rn(wi0, wj1), w10 = __dmul_rn(wi1, wj0),
                 w11 = __dmul_rn(wi1, wj1);
    double v00 = __dmul_rn(w00, wk0), v01 = __dmul_rn(w00, wk1), v10 = __dmul_rn(w01, wk0), v11 = __dmul_rn(w01, wk1);
    // my lower x plane is the carried upper plane of the previous particle when the base cells agree
    const bool match = valid && cvalid && (r00 + (unsigned)ck0 == ca00);
    if (match) {
      v00 += c00;
      v01 += c01;
      v10 += c10;
      v11 += c11;
    }
    const bool orphan = cvalid && !match;
    if (__any_sync(FULL, orphan)) {  // carried values nobody continues: reduce them where they belong
      red_zpair(rho, ca00, ca01, c00, c01, orphan, lane);
      red_zpair(rho, ca10, ca11, c10, c11, orphan, lane);
    }
    red_zpair(rho, r00 + ck0, r00 + ck1, v00, v01, valid, lane);
    red_zpair(rho, r01 + ck0, r01 + ck1, v10, v11, valid, lane);
    c00 = __dmul_rn(w10, wk0);
    c01 = __dmul_rn(w10, wk1);
    c10 = __dmul_rn(w11, wk0);
    c11 = __dmul_rn(w11, wk1);
    ca00 = r10 + ck0;
    ca01 = r10 + ck1;
    ca10 = r11 + ck0;
    ca11 = r11 + ck1;
    cvalid = valid;
    px = nx;
    py = ny;
    pz = nz;
  }
  red_zpair(rho, ca00, ca01, c00, c01, cvalid, lane);
  red_zpair(rho, ca10, ca11, c10, c11, cvalid, lane);
}

// ---------------------------------------------------------------------------
// TSC scatter: planes (centre - 1, centre, centre + 1); the sweep carries two planes of 3 x 3 cells
// ---------------------------------------------------------------------------
template <bool RSD>
__global__ void __launch_bounds__(256) scatter_tsc_sweep_kernel(GridGeom g, const double *__restrict__ psix,
                                                                const double *__restrict__ psiy,
                                                                const double *__restrict__ psiz,
                                                                double *__restrict__ rho, int seg) {
  const int N = g.N;
  const SweepIdx s = sweep_index(N, seg);
  if (!s.live) return;
  const int lane = s.lane;
  const size_t pl = (size_t)N * N;
  const double half = __dmul_rn(0.5, g.d);
  const double qy = __dadd_rn(__dmul_rn(g.d, (double)s.j), half), qz = __dadd_rn(__dmul_rn(g.d, (double)s.k), half);
  size_t idx = s.idx;
  // carried planes: A = the previous particle's centre plane (my lower one if I moved on by one cell),
  //                 B = its upper plane (my centre plane)
  double A[3][3], B[3][3];
  unsigned arow[3] = {kNoAddr, kNoAddr, kNoAddr}, brow[3] = {kNoAddr, kNoAddr, kNoAddr};  // row offsets per y cell
  unsigned ckc[3] = {0, 0, 0};                                                           // the carried z cells
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int c = 0; c < 3; ++c) A[b][c] = B[b][c] = 0.;
  bool cvalid = false;
  double px = psix[idx], py = psiy[idx], pz = psiz[idx];
  const int i_end = s.i_begin + seg;
  for (int i = s.i_begin; i < i_end; ++i, idx += pl) {
    double nx = 0., ny = 0., nz = 0.;
    if (i + 1 < i_end) {
      nx = psix[idx + pl];
      ny = psiy[idx + pl];
      nz = psiz[idx + pl];
    }
    double x, y, z;
    sweep_position<RSD>(g, i, half, qy, qz, px, py, pz, x, y, z);
    const bool valid = in_domain(g, x, y, z);
    int ci[3] = {0, 0, 0}, cj[3] = {0, 0, 0}, ck[3] = {0, 0, 0};
    double wi[3] = {0, 0, 0}, wj[3] = {0, 0, 0}, wk[3] = {0, 0, 0}, u;
    if (valid) {
      tsc_axis(x, g.min1, g.d, N, ci, wi, u);
      tsc_axis(y, g.min2, g.d, N, cj, wj, u);
      tsc_axis(z, g.min3, g.d, N, ck, wk, u);
    }
    unsigned row[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) row[a][b] = ((unsigned)ci[a] * N + cj[b]) * N;
    // continuing the sweep: my lower plane is carried plane A, my centre plane is carried plane B, same (y, z) cells
    const bool match = valid && cvalid && row[0][0] == arow[0] && row[1][0] == brow[0] && (unsigned)ck[0] == ckc[0];
    const bool orphan = cvalid && !match;
    if (__any_sync(FULL, orphan)) {
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        red_ztriple(rho, arow[b] + ckc[0], arow[b] + ckc[1], arow[b] + ckc[2], A[b][0], A[b][1], A[b][2], orphan, lane);
        red_ztriple(rho, brow[b] + ckc[0], brow[b] + ckc[1], brow[b] + ckc[2], B[b][0], B[b][1], B[b][2], orphan, lane);
      }
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      // w_x * w_y * w_z left to right, massFunctions.cc:237-360
      const double w0 = __dmul_rn(wi[0], wj[b]), w1 = __dmul_rn(wi[1], wj[b]), w2 = __dmul_rn(wi[2], wj[b]);
      double lo[3], mid[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        lo[c] = __dmul_rn(w0, wk[c]);
        mid[c] = __dmul_rn(w1, wk[c]);
        if (match) {
          lo[c] += A[b][c];
          mid[c] += B[b][c];
        }
      }
      // the lower plane is complete: no later particle of a regular sweep reaches it
      red_ztriple(rho, row[0][b] + ck[0], row[0][b] + ck[1], row[0][b] + ck[2], lo[0], lo[1], lo[2], valid, lane);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        A[b][c] = mid[c];
        B[b][c] = __dmul_rn(w2, wk[c]);
      }
      arow[b] = row[1][b];
      brow[b] = row[2][b];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) ckc[c] = (unsigned)ck[c];
    cvalid = valid;
    px = nx;
    py = ny;
    pz = nz;
  }
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    red_ztriple(rho, arow[b] + ckc[0], arow[b] + ckc[1], arow[b] + ckc[2], A[b][0], A[b][1], A[b][2], cvalid, lane);
    red_ztriple(rho, brow[b] + ckc[0], brow[b] + ckc[1], brow[b] + ckc[2], B[b][0], B[b][1], B[b][2], cvalid, lane);
  }
}

// ---------------------------------------------------------------------------
// CIC exact adjoint (gather), in place over Psi: V_c = sum_cells r_cell dW_cell/dx_c  (kernels.cu gather_adjoint_kernel)
// ---------------------------------------------------------------------------
template <bool RSD>
__global__ void __launch_bounds__(256) gather_cic_sweep_kernel(GridGeom g, double *ax, double *ay, double *az,
                                                               const double *__restrict__ resid, int seg) {
  const int N = g.N;
  const SweepIdx s = sweep_index(N, seg);
  if (!s.live) return;
  const int lane = s.lane;
  const size_t pl = (size_t)N * N;
  const double half = __dmul_rn(0.5, g.d);
  const double qy = __dadd_rn(__dmul_rn(g.d, (double)s.j), half), qz = __dadd_rn(__dmul_rn(g.d, (double)s.k), half);
  const double inv_d = 1.0 / g.d;
  size_t idx = s.idx;
  // carried: residual at the particle's upper x plane
  double c00 = 0., c01 = 0., c10 = 0., c11 = 0.;
  unsigned cbase = kNoAddr;
  double px = ax[idx], py = ay[idx], pz = az[idx];
  const int i_end = s.i_begin + seg;
  for (int i = s.i_begin; i < i_end; ++i, idx += pl) {
    double nx = 0., ny = 0., nz = 0.;
    if (i + 1 < i_end) {
      nx = ax[idx + pl];
      ny = ay[idx + pl];
      nz = az[idx + pl];
    }
    double x, y, z;
    sweep_position<RSD>(g, i, half, qy, qz, px, py, pz, x, y, z);
    const bool valid = in_domain(g, x, y, z);
    int ci0 = 0, ci1 = 0, cj0 = 0, cj1 = 0, ck0 = 0, ck1 = 0;
    double wi0 = 0., wi1 = 0., wj0 = 0., wj1 = 0., wk0 = 0., wk1 = 0.;
    if (valid) {
      cic_axis(x, g.d, g.L, N, ci0, ci1, wi0, wi1);
      cic_axis(y, g.d, g.L, N, cj0, cj1, wj0, wj1);
      cic_axis(z, g.d, g.L, N, ck0, ck1, wk0, wk1);
    }
    const unsigned r00 = ((unsigned)ci0 * N + cj0) * N, r01 = ((unsigned)ci0 * N + cj1) * N;
    const unsigned r10 = ((unsigned)ci1 * N + cj0) * N, r11 = ((unsigned)ci1 * N + cj1) * N;
    // lower x plane: carried from the previous particle of the sweep, or loaded
    double a00, a01, a10, a11;
    const bool match = valid && (r00 + (unsigned)ck0 == cbase);
    if (match) {
      a00 = c00;
      a01 = c01;
      a10 = c10;
      a11 = c11;
    } else if (valid) {
      a00 = __ldg(resid + r00 + ck0);
      a01 = __ldg(resid + r00 + ck1);
      a10 = __ldg(resid + r01 + ck0);
      a11 = __ldg(resid + r01 + ck1);
    } else {
      a00 = a01 = a10 = a11 = 0.;
    }
    // upper x plane: the lower z cell is loaded, the upper one is the neighbouring lane's lower cell where it is
    const unsigned alo0 = valid ? r10 + ck0 : kNoAddr + 1u, ahi0 = valid ? r10 + ck1 : kNoAddr + 2u;
    const unsigned alo1 = valid ? r11 + ck0 : kNoAddr + 1u, ahi1 = valid ? r11 + ck1 : kNoAddr + 2u;
    double b00 = valid ? __ldg(resid + alo0) : 0., b10 = valid ? __ldg(resid + alo1) : 0.;
    const unsigned n0 = __shfl_down_sync(FULL, alo0, 1), n1 = __shfl_down_sync(FULL, alo1, 1);
    double b01 = __shfl_down_sync(FULL, b00, 1), b11 = __shfl_down_sync(FULL, b10, 1);
    if (valid && !(lane < 31 && n0 == ahi0)) b01 = __ldg(resid + ahi0);
    if (valid && !(lane < 31 && n1 == ahi1)) b11 = __ldg(resid + ahi1);
    // V = sum r dW/dx: the cells in the order of gather_adjoint_kernel (a, b, c)
    double vx = 0., vy = 0., vz = 0.;
    auto acc = [&](double rc, double gx, double wx, double gy, double wy, double gz, double wz) {
      vx += rc * gx * wy * wz;
      vy += rc * wx * gy * wz;
      vz += rc * wx * wy * gz;
    };
    acc(a00, -inv_d, wi0, -inv_d, wj0, -inv_d, wk0);
    acc(a01, -inv_d, wi0, -inv_d, wj0, inv_d, wk1);
    acc(a10, -inv_d, wi0, inv_d, wj1, -inv_d, wk0);
    acc(a11, -inv_d, wi0, inv_d, wj1, inv_d, wk1);
    acc(b00, inv_d, wi1, -inv_d, wj0, -inv_d, wk0);
    acc(b01, inv_d, wi1, -inv_d, wj0, inv_d, wk1);
    acc(b10, inv_d, wi1, inv_d, wj1, -inv_d, wk0);
    acc(b11, inv_d, wi1, inv_d, wj1, inv_d, wk1);
    if (!valid) vx = vy = vz = 0.;
    if constexpr (RSD) vz += g.fgrow * vz;  // d z_s / d Psi_z = 1 + f (cf. HMC_models.cc:295-301)
    ax[idx] = vx;
    ay[idx] = vy;
    az[idx] = vz;
    c00 = b00;
    c01 = b01;
    c10 = b10;
    c11 = b11;
    cbase = valid ? r10 + (unsigned)ck0 : kNoAddr;
    px = nx;
    py = ny;
    pz = nz;
  }
}

// ---------------------------------------------------------------------------
// Lean CIC scatter: the same sweep with the per-particle arithmetic cut to what the common case needs.
//
// ncu's source page of the kernel above (profiles/ncu_source_r01_scatter_instruction_mix.txt) shows a scatter bound by
// instruction issue, not by its reductions: seven pacman() calls with their range tests and inlined fmod loops, three
// IEEE divisions, F2I / I2F conversions, validity predicates around every reduction, 64-bit index arithmetic.  Here
//   * positions, cells and weights come from lean_axis(): branch-free wraps valid for a coordinate within 3/4 of a box
//     length of the box, a correctly rounded quotient by Markstein's correction (q0 = x rd, r = fma(-q0, d, x),
//     q = fma(r, rd, q0) with rd = RN(1/d): three FP64 instructions, same bits as the division), truncation by the
//     2^52 trick.  Same operations in the same order as particle_math.cuh otherwise, so the same bits.  A coordinate
//     outside that range, or a quotient that rounds up to N, raises `slow`: the whole warp then recomputes the particle
//     with the general functions (never in practice: a displacement of 3/4 of the box);
//   * the box starts at the origin (min = 0; the launcher checks), so every wrapped position is inside the domain and
//     the validity predicates disappear;
//   * one address comparison decides the z hand-over of both (x, y) rows of a plane (equal lower-row base addresses
//     imply equal upper rows), and the decision travels to the lane below by ballot instead of a second shuffle;
//   * row offsets by shifts (N is a power of two), 32-bit throughout.
// ---------------------------------------------------------------------------
// the general path for one particle (any displacement), out of line: eight plain reductions, nothing carried
template <bool RSD>
__device__ __noinline__ void lean_slow_deposit(GridGeom g, int i, double half, double qy, double qz, double px, double py,
                                               double pz, double *__restrict__ rho) {
  double x, y, z;
  sweep_position<RSD>(g, i, half, qy, qz, px, py, pz, x, y, z);
  if (!in_domain(g, x, y, z)) return;
  int ci[2], cj[2], ck[2];
  double wi[2], wj[2], wk[2];
  cic_axis(x, g.d, g.L, g.N, ci[0], ci[1], wi[0], wi[1]);
  cic_axis(y, g.d, g.L, g.N, cj[0], cj[1], wj[0], wj[1]);
  cic_axis(z, g.d, g.L, g.N, ck[0], ck[1], wk[0], wk[1]);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const double wab = __dmul_rn(wi[a], wj[b]);
      double *row = rho + ((size_t)ci[a] * g.N + cj[b]) * g.N;
      atomicAdd(row + ck[0], __dmul_rn(wab, wk[0]));
      atomicAdd(row + ck[1], __dmul_rn(wab, wk[1]));
    }
}

// MINB: resident CTAs per SM the register allocation is bounded for; PF: planes of displacement in flight ahead of
// the one being deposited (the sweep is a dependent chain per thread: what keeps HBM busy is loads in flight)
template <bool RSD, int MINB, int PF>
__global__ void __launch_bounds__(256, MINB) scatter_cic_lean_kernel(GridGeom g, const double *__restrict__ psix,
                                                                  const double *__restrict__ psiy,
                                                                  const double *__restrict__ psiz,
                                                                  double *__restrict__ rho, int seg) {
  const int N = g.N;
  const SweepIdx s = sweep_index(N, seg);
  if (!s.live) return;  // whole warps only
  const int lane = s.lane;
  const int sh = 31 - __clz(N);
  const size_t pl = (size_t)N * N;
  const LeanConst lc = lean_const(g);
  const double qy = __dadd_rn(__dmul_rn(g.d, (double)s.j), lc.half), qz = __dadd_rn(__dmul_rn(g.d, (double)s.k), lc.half);
  const double rsd_a = g.cpecvel, rsd_b = g.v_norm;
  size_t idx = s.idx;
  // carried: the particle's upper x plane, 2 x 2 (y, z) cells, and the address of its (cj0, ck0) cell
  double c00 = 0., c01 = 0., c10 = 0., c11 = 0.;
  unsigned cbase = kNoAddr, cj1s_c = 0, ck1_c = 0;
  const int i_end = s.i_begin + seg;
  double fx[PF], fy[PF], fz[PF];  // displacements of planes i .. i + PF - 1
#pragma unroll
  for (int f = 0; f < PF; ++f) {
    const bool in = s.i_begin + f < i_end;
    fx[f] = in ? psix[idx + (size_t)f * pl] : 0.;
    fy[f] = in ? psiy[idx + (size_t)f * pl] : 0.;
    fz[f] = in ? psiz[idx + (size_t)f * pl] : 0.;
  }
  double fi = (double)s.i_begin;
  for (int i = s.i_begin; i < i_end; ++i, idx += pl, fi += 1.0) {
    const double px = fx[0], py = fy[0], pz = fz[0];
    double nx = 0., ny = 0., nz = 0.;
    if (i + PF < i_end) {  // the displacement PF planes ahead goes in flight while this one is deposited
      nx = psix[idx + (size_t)PF * pl];
      ny = psiy[idx + (size_t)PF * pl];
      nz = psiz[idx + (size_t)PF * pl];
    }
#pragma unroll
    for (int f = 0; f + 1 < PF; ++f) {
      fx[f] = fx[f + 1];
      fy[f] = fy[f + 1];
      fz[f] = fz[f + 1];
    }
    fx[PF - 1] = nx;
    fy[PF - 1] = ny;
    fz[PF - 1] = nz;
    unsigned ci0, ci1, cj0, cj1, ck0, ck1;
    double wi0, wi1, wj0, wj1, wk0, wk1;
    bool slow = false;
    {
      // disp_part.cc:55-126, rsd.cc:30-64 (particle_math.cuh particle_position / sweep_position)
      const double x = lean_wrap(__dadd_rn(__dadd_rn(__dmul_rn(lc.d, fi), lc.half), px), lc, slow);
      const double y = lean_wrap(__dadd_rn(qy, py), lc, slow);
      double z = lean_wrap(__dadd_rn(qz, pz), lc, slow);
      if constexpr (RSD) z = lean_wrap(__dadd_rn(z, __dmul_rn(__dmul_rn(rsd_a, pz), rsd_b)), lc, slow);
      lean_axis(x, lc, ci0, ci1, wi0, wi1, slow);
      lean_axis(y, lc, cj0, cj1, wj0, wj1, slow);
      lean_axis(z, lc, ck0, ck1, wk0, wk1, slow);
    }
    if (__any_sync(FULL, slow)) {
      // never in practice (a displacement of 3/4 of the box): the warp reduces what it carries, deposits this
      // particle the general way and carries nothing into the next plane
      const unsigned crow0 = cbase & ~(lc.N - 1u), ck0_c = cbase & (lc.N - 1u);
      const unsigned crow1 = (crow0 & ~((lc.N << sh) - lc.N)) + cj1s_c;
      const bool have = cbase != kNoAddr;
      red_zpair(rho, cbase, crow0 + ck1_c, c00, c01, have, lane);
      red_zpair(rho, crow1 + ck0_c, crow1 + ck1_c, c10, c11, have, lane);
      lean_slow_deposit<RSD>(g, i, lc.half, qy, qz, px, py, pz, rho);
      cbase = kNoAddr;
      continue;
    }
    const unsigned r0 = ci0 << (2 * sh), r1 = ci1 << (2 * sh), j0 = cj0 << sh, j1 = cj1 << sh;
    const unsigned a_lo = r0 + j0 + ck0, a_hi = r0 + j0 + ck1;
    // mass * w_x * w_y * w_z evaluated left to right, massFunctions.cc:129-157
    const double w00 = __dmul_rn(wi0, wj0), w01 = __dmul_rn(wi0, wj1), w10 = __dmul_rn(wi1, wj0),
                 w11 = __dmul_rn(wi1, wj1);
    double v00 = __dmul_rn(w00, wk0), v01 = __dmul_rn(w00, wk1), v10 = __dmul_rn(w01, wk0),
           v11 = __dmul_rn(w01, wk1);
    // my lower x plane is the carried upper plane of the previous particle when the base cells agree
    const bool match = a_lo == cbase;
    if (match) {
      v00 += c00;
      v01 += c01;
      v10 += c10;
      v11 += c11;
    }
    const bool orphan = cbase != kNoAddr && !match;
    if (__any_sync(FULL, orphan)) {  // carried values nobody continues: reduce them where they belong
      const unsigned crow0 = cbase & ~(lc.N - 1u), ck0_c = cbase & (lc.N - 1u);
      const unsigned crow1 = (crow0 & ~((lc.N << sh) - lc.N)) + cj1s_c;
      red_zpair(rho, cbase, crow0 + ck1_c, c00, c01, orphan, lane);
      red_zpair(rho, crow1 + ck0_c, crow1 + ck1_c, c10, c11, orphan, lane);
    }
    // z hand-over: the lane below hands me its upper z cells when they are my lower ones (both rows at once)
    const unsigned p_ahi = __shfl_up_sync(FULL, a_hi, 1);
    const double p_v01 = __shfl_up_sync(FULL, v01, 1), p_v11 = __shfl_up_sync(FULL, v11, 1);
    const bool take = lane > 0 && p_ahi == a_lo;
    const unsigned takers = __ballot_sync(FULL, take);
    const bool handed_up = ((takers >> 1) >> lane) & 1u;  // lane + 1 takes mine
    if (take) {
      v00 += p_v01;
      v10 += p_v11;
    }
    atomicAdd(rho + a_lo, v00);
    atomicAdd(rho + (r0 + j1 + ck0), v10);
    if (!handed_up) {
      atomicAdd(rho + a_hi, v01);
      atomicAdd(rho + (r0 + j1 + ck1), v11);
    }
    c00 = __dmul_rn(w10, wk0);
    c01 = __dmul_rn(w10, wk1);
    c10 = __dmul_rn(w11, wk0);
    c11 = __dmul_rn(w11, wk1);
    cbase = r1 + j0 + ck0;
    cj1s_c = j1;
    ck1_c = ck1;
  }
  {
    const unsigned crow0 = cbase & ~(lc.N - 1u), ck0_c = cbase & (lc.N - 1u);
    const unsigned crow1 = (crow0 & ~((lc.N << sh) - lc.N)) + cj1s_c;
    const bool have = cbase != kNoAddr;
    red_zpair(rho, cbase, crow0 + ck1_c, c00, c01, have, lane);
    red_zpair(rho, crow1 + ck0_c, crow1 + ck1_c, c10, c11, have, lane);
  }
}

// ---------------------------------------------------------------------------
// Lean CIC gather (exact adjoint): the sweep of gather_cic_sweep_kernel with the lean per-particle arithmetic above
// ---------------------------------------------------------------------------
// the general path for one particle, out of line: eight loads, nothing carried
template <bool RSD>
__device__ __noinline__ void lean_slow_gather(GridGeom g, int i, double half, double qy, double qz, double px, double py,
                                              double pz, const double *__restrict__ resid, double *vout) {
  double x, y, z;
  sweep_position<RSD>(g, i, half, qy, qz, px, py, pz, x, y, z);
  double vx = 0., vy = 0., vz = 0.;
  if (in_domain(g, x, y, z)) {
    int ci[2], cj[2], ck[2];
    double wi[2], wj[2], wk[2];
    cic_axis(x, g.d, g.L, g.N, ci[0], ci[1], wi[0], wi[1]);
    cic_axis(y, g.d, g.L, g.N, cj[0], cj[1], wj[0], wj[1]);
    cic_axis(z, g.d, g.L, g.N, ck[0], ck[1], wk[0], wk[1]);
    const double inv_d = 1.0 / g.d;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c) {
          const double rc = __ldg(resid + ((size_t)ci[a] * g.N + cj[b]) * g.N + ck[c]);
          const double gx = a ? inv_d : -inv_d, gy = b ? inv_d : -inv_d, gz = c ? inv_d : -inv_d;
          vx += rc * gx * wj[b] * wk[c];
          vy += rc * wi[a] * gy * wk[c];
          vz += rc * wi[a] * wj[b] * gz;
        }
  }
  vout[0] = vx;
  vout[1] = vy;
  vout[2] = vz;
}

template <bool RSD, int MINB, int PF>
__global__ void __launch_bounds__(256, MINB) gather_cic_lean_kernel(GridGeom g, double *ax, double *ay, double *az,
                                                                 const double *__restrict__ resid, int seg) {
  const int N = g.N;
  const SweepIdx s = sweep_index(N, seg);
  if (!s.live) return;
  const int lane = s.lane;
  const int sh = 31 - __clz(N);
  const size_t pl = (size_t)N * N;
  const LeanConst lc = lean_const(g);
  const double qy = __dadd_rn(__dmul_rn(g.d, (double)s.j), lc.half), qz = __dadd_rn(__dmul_rn(g.d, (double)s.k), lc.half);
  const double rsd_a = g.cpecvel, rsd_b = g.v_norm;
  const double inv_d = 1.0 / g.d;
  size_t idx = s.idx;
  // carried: residual at the particle's upper x plane
  double c00 = 0., c01 = 0., c10 = 0., c11 = 0.;
  unsigned cbase = kNoAddr;
  const int i_end = s.i_begin + seg;
  double fx[PF], fy[PF], fz[PF];  // displacements of planes i .. i + PF - 1 (read ahead of the in-place stores)
#pragma unroll
  for (int f = 0; f < PF; ++f) {
    const bool in = s.i_begin + f < i_end;
    fx[f] = in ? ax[idx + (size_t)f * pl] : 0.;
    fy[f] = in ? ay[idx + (size_t)f * pl] : 0.;
    fz[f] = in ? az[idx + (size_t)f * pl] : 0.;
  }
  double fi = (double)s.i_begin;
  for (int i = s.i_begin; i < i_end; ++i, idx += pl, fi += 1.0) {
    const double px = fx[0], py = fy[0], pz = fz[0];
    double nx = 0., ny = 0., nz = 0.;
    if (i + PF < i_end) {
      nx = ax[idx + (size_t)PF * pl];
      ny = ay[idx + (size_t)PF * pl];
      nz = az[idx + (size_t)PF * pl];
    }
#pragma unroll
    for (int f = 0; f + 1 < PF; ++f) {
      fx[f] = fx[f + 1];
      fy[f] = fy[f + 1];
      fz[f] = fz[f + 1];
    }
    fx[PF - 1] = nx;
    fy[PF - 1] = ny;
    fz[PF - 1] = nz;
    unsigned ci0, ci1, cj0, cj1, ck0, ck1;
    double wi0, wi1, wj0, wj1, wk0, wk1;
    bool slow = false;
    {
      const double x = lean_wrap(__dadd_rn(__dadd_rn(__dmul_rn(lc.d, fi), lc.half), px), lc, slow);
      const double y = lean_wrap(__dadd_rn(qy, py), lc, slow);
      double z = lean_wrap(__dadd_rn(qz, pz), lc, slow);
      if constexpr (RSD) z = lean_wrap(__dadd_rn(z, __dmul_rn(__dmul_rn(rsd_a, pz), rsd_b)), lc, slow);
      lean_axis(x, lc, ci0, ci1, wi0, wi1, slow);
      lean_axis(y, lc, cj0, cj1, wj0, wj1, slow);
      lean_axis(z, lc, ck0, ck1, wk0, wk1, slow);
    }
    if (__any_sync(FULL, slow)) {  // never in practice: the general functions, nothing carried
      double v[3];
      lean_slow_gather<RSD>(g, i, lc.half, qy, qz, px, py, pz, resid, v);
      if constexpr (RSD) v[2] += g.fgrow * v[2];
      ax[idx] = v[0];
      ay[idx] = v[1];
      az[idx] = v[2];
      cbase = kNoAddr;
      continue;
    }
    const unsigned r0 = ci0 << (2 * sh), r1 = ci1 << (2 * sh), j0 = cj0 << sh, j1 = cj1 << sh;
    // lower x plane: carried from the previous particle of the sweep, or loaded
    double a00, a01, a10, a11;
    if (r0 + j0 + ck0 == cbase) {
      a00 = c00;
      a01 = c01;
      a10 = c10;
      a11 = c11;
    } else {
      a00 = __ldg(resid + (r0 + j0 + ck0));
      a01 = __ldg(resid + (r0 + j0 + ck1));
      a10 = __ldg(resid + (r0 + j1 + ck0));
      a11 = __ldg(resid + (r0 + j1 + ck1));
    }
    // upper x plane: the lower z cells are loaded, the upper ones are the lane above's lower cells where they are
    const unsigned b_lo = r1 + j0 + ck0, b_hi = r1 + j0 + ck1;
    const double b00 = __ldg(resid + b_lo), b10 = __ldg(resid + (r1 + j1 + ck0));
    const unsigned n_lo = __shfl_down_sync(FULL, b_lo, 1);
    double b01 = __shfl_down_sync(FULL, b00, 1), b11 = __shfl_down_sync(FULL, b10, 1);
    if (!(lane < 31 && n_lo == b_hi)) {  // equal row bases imply the same second row (see the scatter)
      b01 = __ldg(resid + b_hi);
      b11 = __ldg(resid + (r1 + j1 + ck1));
    }
    // V = sum r dW/dx: the cells in the order of gather_adjoint_kernel (a, b, c)
    double vx = 0., vy = 0., vz = 0.;
    auto acc = [&](double rc, double gx, double wx, double gy, double wy, double gz, double wz) {
      vx += rc * gx * wy * wz;
      vy += rc * wx * gy * wz;
      vz += rc * wx * wy * gz;
    };
    acc(a00, -inv_d, wi0, -inv_d, wj0, -inv_d, wk0);
    acc(a01, -inv_d, wi0, -inv_d, wj0, inv_d, wk1);
    acc(a10, -inv_d, wi0, inv_d, wj1, -inv_d, wk0);
    acc(a11, -inv_d, wi0, inv_d, wj1, inv_d, wk1);
    acc(b00, inv_d, wi1, -inv_d, wj0, -inv_d, wk0);
    acc(b01, inv_d, wi1, -inv_d, wj0, inv_d, wk1);
    acc(b10, inv_d, wi1, inv_d, wj1, -inv_d, wk0);
    acc(b11, inv_d, wi1, inv_d, wj1, inv_d, wk1);
    if constexpr (RSD) vz += g.fgrow * vz;  // d z_s / d Psi_z = 1 + f (cf. HMC_models.cc:295-301)
    ax[idx] = vx;
    ay[idx] = vy;
    az[idx] = vz;
    c00 = b00;
    c01 = b01;
    c10 = b10;
    c11 = b11;
    cbase = b_lo;
  }
}

int sweep_seg(const GridGeom &g) {
  const int N = g.N;
  // short segments keep enough threads in flight (N^2 * N/seg of them); each one ends with a flush of the carried plane
  int seg = g.sweep > 1 ? g.sweep : (N >= 256 ? 16 : 8);
  if (seg > N) seg = N;
  while (N % seg) --seg;
  return seg;
}

// a full cube (no slab halo / plane mapping), plain Lagrangian lattice (no cell-boundary averaging)
bool sweep_geometry_ok(const GridGeom &g) {
  return g.N >= 128 && g.Ns == g.N && g.H == 0 && g.x0 == 0 && !g.cellbound && g.sweep;
}

unsigned sweep_blocks(int N, int seg) {
  const size_t warps = (size_t)(N / 32) * N * (N / seg);
  return (unsigned)((warps * 32 + 255) / 256);
}

}  // namespace

bool scatter_sweep_applicable(const GridGeom &g, const double *posx) {
  return posx == nullptr && (g.masskernel == 1 || g.masskernel == 2) && sweep_geometry_ok(g);
}

void launch_scatter_sweep(const GridGeom &g, const double *psix, const double *psiy, const double *psiz, double *rho,
                          cudaStream_t st) {
  ProfScope prof(KK_SCATTER, st);
  const int seg = sweep_seg(g);
  BGPU_CUDA(cudaMemsetAsync(rho, 0, (size_t)g.N * g.N * g.N * sizeof(double), st));
  const unsigned blocks = sweep_blocks(g.N, seg);
  if (g.masskernel == 1 && g.lean && g.min1 == 0. && g.min2 == 0. && g.min3 == 0.) {
    // measured (B200, 256^3, round 2): more resident CTAs (3 per SM) or deeper read-ahead (2-3 planes) leave the
    // scatter where it is or slow it down (0.341 -> 0.344 / 0.361 / 0.414 / 0.545 ms for (2,2) / (3,2) / (3,3) / (4,3)):
    // it is bound by the L2 reduction units, not by loads in flight
    if (g.rsd) scatter_cic_lean_kernel<true, 2, 1><<<blocks, 256, 0, st>>>(g, psix, psiy, psiz, rho, seg);
    else scatter_cic_lean_kernel<false, 2, 1><<<blocks, 256, 0, st>>>(g, psix, psiy, psiz, rho, seg);
  } else if (g.masskernel == 1) {
    if (g.rsd) scatter_cic_sweep_kernel<true><<<blocks, 256, 0, st>>>(g, psix, psiy, psiz, rho, seg);
    else scatter_cic_sweep_kernel<false><<<blocks, 256, 0, st>>>(g, psix, psiy, psiz, rho, seg);
  } else {
    if (g.rsd) scatter_tsc_sweep_kernel<true><<<blocks, 256, 0, st>>>(g, psix, psiy, psiz, rho, seg);
    else scatter_tsc_sweep_kernel<false><<<blocks, 256, 0, st>>>(g, psix, psiy, psiz, rho, seg);
  }
  BGPU_LAUNCHED(1);
}

bool gather_sweep_applicable(const GridGeom &g) { return g.masskernel == 1 && sweep_geometry_ok(g); }

void launch_gather_sweep(const GridGeom &g, double *ax, double *ay, double *az, const double *resid, cudaStream_t st) {
  ProfScope prof(KK_GATHER, st);
  const int seg = sweep_seg(g);
  const unsigned blocks = sweep_blocks(g.N, seg);
  if (g.lean && g.min1 == 0. && g.min2 == 0. && g.min3 == 0.) {
    // measured (B200, 256^3, round 2): 3 resident CTAs per SM with the displacement read 2 planes ahead:
    // 0.287 -> 0.252 ms ((2,2): 0.287, (3,3): 0.281, (4,3): 0.373 -- spills); BGPU_LEAN=21 keeps the (2,1) form
    if (g.lean == 21) {
      if (g.rsd) gather_cic_lean_kernel<true, 2, 1><<<blocks, 256, 0, st>>>(g, ax, ay, az, resid, seg);
      else gather_cic_lean_kernel<false, 2, 1><<<blocks, 256, 0, st>>>(g, ax, ay, az, resid, seg);
    } else if (g.rsd) gather_cic_lean_kernel<true, 3, 2><<<blocks, 256, 0, st>>>(g, ax, ay, az, resid, seg);
    else gather_cic_lean_kernel<false, 3, 2><<<blocks, 256, 0, st>>>(g, ax, ay, az, resid, seg);
  } else if (g.rsd) gather_cic_sweep_kernel<true><<<blocks, 256, 0, st>>>(g, ax, ay, az, resid, seg);
  else gather_cic_sweep_kernel<false><<<blocks, 256, 0, st>>>(g, ax, ay, az, resid, seg);
  BGPU_LAUNCHED(1);
}

}  // namespace bgpu
