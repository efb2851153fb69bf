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
  if (g.masskernel == 1) {
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
  if (g.rsd) gather_cic_sweep_kernel<true><<<blocks, 256, 0, st>>>(g, ax, ay, az, resid, seg);
  else gather_cic_sweep_kernel<false><<<blocks, 256, 0, st>>>(g, ax, ay, az, resid, seg);
  BGPU_LAUNCHED(1);
}

}  // namespace bgpu
