// barcode_b200/csrc/fft_plan.cu -- host side of the hand-written 3-D FFT:
// twiddle tables, per-size tile shapes and the launch sequences for r2c / c2r.
// Replaces plan_pkg / fftR2Cplanned / fftC2Rplanned
// (/root/reference/barlib/src/fftwrapper.cc:88-125,281-324).
#include "fft3d.h"

#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include <cstdlib>
#include <cstring>

#include "fft.cuh"
#include "fft_tma.cuh"
#include "fft_fused.cuh"
#include "fft_slab_generic.cuh"
#include "util.h"

namespace bgpu {

// ---------------------------------------------------------------------------
// tiny grids (N == 8): the eight-elements-per-thread z pass needs N/2 >= 8, so
// rows are transformed by direct summation, one thread per row.  Only the
// reference's smoke configuration (test/run/input.par, Nx = 8) lands here.
// ---------------------------------------------------------------------------
template <int N>
__global__ void tiny_r2c_zpass(const double *__restrict__ in, double2 *__restrict__ out,
                               const double2 *__restrict__ twN, ROp lop, size_t nrows) {
  const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  double x[N];
  for (int j = 0; j < N; ++j) {
    double v = in[row * N + j];
    if (lop.kind == R_LOAD_SCALE) v *= lop.a;
    else if (lop.kind == R_SCALE_MUL) v *= lop.a * lop.aux[row * N + j];
    x[j] = v;
  }
  for (int k = 0; k <= N / 2; ++k) {
    double2 acc = make_double2(0.0, 0.0);
    for (int j = 0; j < N; ++j) {
      const double2 w = twN[(j * k) % N];
      acc.x += x[j] * w.x;
      acc.y += x[j] * w.y;
    }
    if (k == 0 || k == N / 2) acc.y = 0.0;
    out[row * (N / 2 + 1) + k] = acc;
  }
}

template <int N>
__global__ void tiny_c2r_zpass(const double2 *__restrict__ in, double *__restrict__ out,
                               const double2 *__restrict__ twN, ROp sop, size_t nrows) {
  const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows || (sop.skip && *sop.skip)) return;
  double2 X[N / 2 + 1];
  for (int k = 0; k <= N / 2; ++k) X[k] = in[row * (N / 2 + 1) + k];
  X[0].y = 0.0;
  X[N / 2].y = 0.0;
  for (int j = 0; j < N; ++j) {
    double acc = X[0].x + ((j & 1) ? -X[N / 2].x : X[N / 2].x);
    for (int k = 1; k < N / 2; ++k) {
      const double2 w = twN[(j * k) % N];  // exp(-i a); we need Re[X exp(+i a)] * 2
      acc += 2.0 * (X[k].x * w.x + X[k].y * w.y);
    }
    double v = sop.a * acc;
    const size_t idx = row * N + j;
    if (sop.kind == R_SCALE_MUL) v *= sop.aux[idx];
    else if (sop.kind == R_AXPY) v += out[idx];
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------
// per-size tile shapes
// ---------------------------------------------------------------------------
template <int N> struct Shape;
template <> struct Shape<8>    { static constexpr int T = 4,  TR = 0;  };
template <> struct Shape<16>   { static constexpr int T = 8,  TR = 16; };
template <> struct Shape<32>   { static constexpr int T = 16, TR = 16; };
template <> struct Shape<64>   { static constexpr int T = 16, TR = 16; };
template <> struct Shape<128>  { static constexpr int T = 16, TR = 16; };
template <> struct Shape<256>  { static constexpr int T = 8,  TR = 8;  };
template <> struct Shape<512>  { static constexpr int T = 8,  TR = 8;  };
template <> struct Shape<1024> { static constexpr int T = 4,  TR = 4;  };

// ---------------------------------------------------------------------------
// TMA-staged strided pass (fft_tma.cuh): tensor maps and launch
// ---------------------------------------------------------------------------

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    BGPU_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("bgpu: cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor maps over half-grid arrays seen as doubles.
//   layout 0: [n_slow][n_mid][row]     rank 3; the box is 8 pencils wide (cplx: 16 doubles = 128 B,
//             real: 8 doubles = 64 B) and `rows` long along the transformed axis (axis 1 = mid,
//             axis 0 = slow).  Cube: n_slow = n_mid = N.  x-slab [Ns][N][.]: the y pass.
//             Transposed slab [N][Ns][.]: the x pass.
//   layout 1: [G][Ns][Ns][row]         rank 4, the packed all-to-all buffer [peer][x_l][y_l][z]; the
//             box is min(Ns, 256) rows of one (peer, x_l).
static CUtensorMap make_half_grid_map(const void *base, int N, int axis, bool cplx, int layout, int n_slow,
                                      int n_mid) {
  CUtensorMap m;
  const cuuint64_t nzh = (cuuint64_t)N / 2 + 1;
  const cuuint64_t row_doubles = cplx ? 2 * nzh : nzh;                 // valid extent
  const cuuint64_t pitch = cplx ? nzh * 16 : (nzh + 1) * 8;            // bytes, multiple of 16
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r;
  if (layout == 0) {
    const cuuint64_t dims[3] = {row_doubles, (cuuint64_t)n_mid, (cuuint64_t)n_slow};
    const cuuint64_t strides[2] = {pitch, pitch * (cuuint64_t)n_mid};
    const cuuint32_t rows = N > 256 ? 256 : N;
    const cuuint32_t box[3] = {cplx ? 16u : 8u, axis == 1 ? rows : 1u, axis == 0 ? rows : 1u};
    r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, cplx ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t ns = (cuuint64_t)n_mid, g = (cuuint64_t)n_slow;  // n_mid = Ns, n_slow = G
    const cuuint64_t dims[4] = {row_doubles, ns, ns, g};
    const cuuint64_t strides[3] = {pitch, pitch * ns, pitch * ns * ns};
    const cuuint32_t rb = (cuuint32_t)(ns < 256 ? ns : 256);
    const cuuint32_t box[4] = {16u, axis == 1 ? rb : 1u, axis == 1 ? 1u : rb, 1u};  // rows along y_l or x_l
    r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) throw std::runtime_error("bgpu: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return m;
}

const CUtensorMap &Fft3d::tensor_map(const void *base, int axis, bool cplx, int layout, int n_slow, int n_mid) const {
  for (auto &e : maps_)
    if (e.base == base && e.axis == axis && e.cplx == cplx && e.layout == layout && e.n_slow == n_slow &&
        e.n_mid == n_mid)
      return e.map;
  if (maps_.size() >= 96) maps_.clear();
  maps_.push_back(MapEntry{base, axis, cplx, layout, n_slow, n_mid,
                           make_half_grid_map(base, N, axis, cplx, layout, n_slow, n_mid)});
  return maps_.back().map;
}

// launch with (or without) the programmatic-stream-serialization attribute, see fft_tma.cuh pdl_wait()
template <class... KArgs, class... Args>
static void launch_pass(void (*kern)(KArgs...), int blocks, int threads, size_t smem, cudaStream_t st, bool pdl,
                        Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  BGPU_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}

constexpr int kMaxDevices = 16;  // GPUs one process may drive (per-device launch configuration below)

template <int N> struct TmaShape;
template <> struct TmaShape<128> { static constexpr int E = 8,  MINB = 4; };
template <> struct TmaShape<256> { static constexpr int E = 8,  MINB = 2; };
template <> struct TmaShape<512> { static constexpr int E = 16, MINB = 1; };

// ring depth: as many stages (at most 3) as fit beside the co-resident CTAs
template <int N, int AUX>
struct TmaStages {
  static constexpr int per_cta = (227 * 1024 - 2048) / (AUX <= 0 ? TmaShape<N>::MINB : 1);
  static constexpr int fit = per_cta / TmaTile<N, AUX>::stage_bytes;
  static constexpr int value = fit >= 3 ? 3 : fit;
};

template <int N> constexpr bool tma_has_size() { return N == 128 || N == 256 || N == 512; }

template <int N, int AUX>
constexpr bool tma_supported() {
  if constexpr (!tma_has_size<N>()) return false;
  else return TmaStages<N, AUX>::value >= 2;
}

// how one strided pass sees its input and output arrays (see make_half_grid_map)
struct PassIo {
  int n_other = 0;      // pencils' "other" extent; 0 = N (cube)
  int other0 = 0;       // global index of other == 0
  bool in_packed = false, out_packed = false;
  int G = 1, Ns = 0;
  double2 *const *peer_out = nullptr;  // fused transpose: receive buffer of every rank, as mapped here
  int my_rank = 0;
};

// E = elements per lane: TmaShape<N>::E puts a pencil inside one warp; half of it spreads a 512-point pencil over
// two warps (16 warps per SM instead of 8, <= 128 registers; BGPU_FFT_2WARP=1, opt-in until measured)
template <int N, int DIR, int AXIS, int AUX, int E = TmaShape<N>::E>
static void launch_strided_tma(const Fft3d &f, const double2 *in, double2 *out, KOp lop, KOp sop, const PassIo &io,
                               cudaStream_t st) {
  if constexpr (N == 512 && E == TmaShape<N>::E) {
    if (f.two_warp) {
      launch_strided_tma<N, DIR, AXIS, AUX, TmaShape<N>::E / 2>(f, in, out, lop, sop, io, st);
      return;
    }
  }
  constexpr int MINB = AUX <= 0 ? TmaShape<N>::MINB : 1;
  constexpr int NSTAGE = TmaStages<N, AUX>::value;
  constexpr int threads = 8 * (N / E);
  constexpr int smem = NSTAGE * TmaTile<N, AUX>::stage_bytes + 1024 + 64;
  auto kern = fft_strided_tma<N, E, NSTAGE, DIR, AXIS, AUX, MINB>;
  static int blocks_per_sm_dev[kMaxDevices] = {};  // per instantiation AND device: the attribute is a per-device setting
  int &blocks_per_sm = blocks_per_sm_dev[f.device % kMaxDevices];
  if (!blocks_per_sm) {
    BGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    blocks_per_sm = occ > 0 ? occ : 1;
  }
  const int n_other = io.n_other ? io.n_other : N;
  // extents of the rank-3 layout: the pass axis has N entries, the other strided axis n_other
  const int n_slow = AXIS == 0 ? N : n_other, n_mid = AXIS == 0 ? n_other : N;
  TmaMaps maps;
  maps.in = io.in_packed ? f.tensor_map(in, AXIS, true, 1, io.G, io.Ns) : f.tensor_map(in, AXIS, true, 0, n_slow, n_mid);
  maps.out = io.peer_out ? maps.in
             : io.out_packed ? f.tensor_map(out, AXIS, true, 1, io.G, io.Ns)
                             : f.tensor_map(out, AXIS, true, 0, n_slow, n_mid);
  maps.auxr = AUX >= 1 ? f.tensor_map(lop.real0, AXIS, false, 0, n_slow, n_mid) : maps.in;
  maps.auxc = AUX >= 2 ? f.tensor_map(lop.cplx0, AXIS, true, 0, n_slow, n_mid) : maps.in;
  if (io.peer_out)
    for (int h = 0; h < io.G; ++h) maps.peer[h] = f.tensor_map(io.peer_out[h], AXIS, true, 1, io.G, io.Ns);
  PassGeom geo{n_other, io.other0, io.in_packed ? io.Ns : 0, io.out_packed ? io.Ns : 0, io.peer_out ? io.Ns : 0,
               io.my_rank, f.next_rev()};
  const int tiles = n_other * ((N / 2 + 1 + 7) / 8);
  int blocks = f.sm_count * blocks_per_sm;
  if (blocks > tiles) blocks = tiles;
  ProfScope prof(AXIS == 0 ? KK_FFT_STRIDED_X : KK_FFT_STRIDED, st);
  launch_pass(kern, blocks, threads, smem, st, f.use_pdl, maps, f.twN, lop, sop, geo);
  BGPU_LAUNCHED(1);
}

// returns true if the TMA kernel took the launch
template <int N, int DIR, int AXIS>
static bool try_strided_tma(const Fft3d &f, const double2 *in, double2 *out, KOp lop, KOp sop, const PassIo &io,
                            cudaStream_t st) {
  if (!f.use_tma) return false;
  if constexpr (!tma_has_size<N>()) {
    return false;
  } else {
    const bool sop_ok = sop.kind == K_NONE || sop.kind == K_INVLAP_SET || sop.kind == K_INVLAP_ADD;
    if (!sop_ok) return false;
    if (lop.kind == K_MULREAL || lop.kind == K_FINAL) {
      if constexpr (DIR == +1 && AXIS == 0) {
        if (lop.kind == K_MULREAL) {
          if constexpr (tma_supported<N, 1>()) {
            launch_strided_tma<N, DIR, AXIS, 1>(f, in, out, lop, sop, io, st);
            return true;
          }
        } else {
          if constexpr (tma_supported<N, 2>()) {
            launch_strided_tma<N, DIR, AXIS, 2>(f, in, out, lop, sop, io, st);
            return true;
          }
        }
      }
      return false;
    }
    launch_strided_tma<N, DIR, AXIS, 0>(f, in, out, lop, sop, io, st);
    return true;
  }
}

// generic (cp.async) strided pass over slab layouts: N = 1024, and N = 128 for tests (fft_slab_generic.cuh)
template <int N> constexpr bool slab_generic_has_size() { return N == 128 || N == 1024; }

template <int N, int DIR, int AXIS>
static void launch_strided_slab(const Fft3d &f, const double2 *in, double2 *out, KOp lop, KOp sop, const PassIo &io,
                                cudaStream_t st) {
  if constexpr (slab_generic_has_size<N>()) {
    constexpr int T = Shape<N>::T;
    constexpr int threads = T * N / 8;
    constexpr int smem = 2 * N * T * (int)sizeof(double2);
    auto kern = fft_strided_pass_slab<N, T, DIR, AXIS>;
    static int blocks_per_sm_dev[kMaxDevices] = {};
    int &blocks_per_sm = blocks_per_sm_dev[f.device % kMaxDevices];
    if (!blocks_per_sm) {
      BGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int occ = 0;
      BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
      blocks_per_sm = occ > 0 ? occ : 1;
    }
    if (io.peer_out || lop.kind == K_FINAL || ((io.in_packed || io.out_packed) && AXIS != 1))
      throw std::runtime_error("bgpu: layout not supported by the generic slab pass");
    const int n_other = io.n_other ? io.n_other : N;
    if (n_other % T) throw std::runtime_error("bgpu: planes per rank must be a multiple of the tile width");
    SlabGeom geo{n_other, io.other0, io.in_packed ? io.Ns : 0, io.out_packed ? io.Ns : 0};
    const int tiles = n_other * ((N / 2) / T) + n_other / T;
    int blocks = f.sm_count * blocks_per_sm;
    if (blocks > tiles) blocks = tiles;
    ProfScope prof(AXIS == 0 ? KK_FFT_STRIDED_X : KK_FFT_STRIDED, st);
    kern<<<blocks, threads, smem, st>>>(in, out, f.twN, lop, sop, geo);
    BGPU_LAUNCHED(1);
  }
}

template <int N, int DIR, int AXIS>
static void launch_strided(const Fft3d &f, const double2 *in, double2 *out, const double2 *tw, KOp lop, KOp sop,
                           cudaStream_t st, const PassIo &io = PassIo{}) {
  constexpr int T = Shape<N>::T;
  constexpr int threads = T * N / 8;
  constexpr int tiles = N * ((N / 2) / T) + N / T;
  constexpr size_t smem = (size_t)N * T * sizeof(double2);
  const bool generic = slab_generic_has_size<N>() && (f.G > 1 || io.n_other) && (f.force_generic || !tma_has_size<N>());
  if (generic) {
    launch_strided_slab<N, DIR, AXIS>(f, in, out, lop, sop, io, st);
    return;
  }
  if (try_strided_tma<N, DIR, AXIS>(f, in, out, lop, sop, io, st)) return;
  if (f.G > 1 || io.n_other)
    throw std::runtime_error("bgpu: the slab-decomposed transform needs the TMA-staged pass (N = 128, 256 or 512)");
  ProfScope prof(AXIS == 0 ? KK_FFT_STRIDED_X : KK_FFT_STRIDED, st);
  if constexpr (N >= 128) {
    // persistent + cp.async prefetch: one wave of CTAs walks all tiles
    const int blocks = f.strided_blocks < tiles ? f.strided_blocks : tiles;
    fft_strided_pass_pipelined<N, T, DIR, AXIS><<<blocks, threads, 2 * smem, st>>>(in, out, tw, lop, sop);
  } else {
    fft_strided_pass<N, T, DIR, AXIS><<<tiles, threads, smem, st>>>(in, out, tw, lop, sop);
  }
  BGPU_LAUNCHED(1);
}

// ---------------------------------------------------------------------------
// bulk-copy-staged z pass (fft_tma.cuh)
// ---------------------------------------------------------------------------
template <int N> struct ZShape;
template <> struct ZShape<128>  { static constexpr int E = 8,  TR = 32, MINB = 2; };
template <> struct ZShape<256>  { static constexpr int E = 8,  TR = 16, MINB = 2; };
template <> struct ZShape<512>  { static constexpr int E = 8,  TR = 8,  MINB = 2; };
template <> struct ZShape<1024> { static constexpr int E = 16, TR = 4,  MINB = 1; };

template <int N> constexpr bool ztma_has_size() { return N == 128 || N == 256 || N == 512 || N == 1024; }

template <int N, bool C2R, bool AUX>
static void launch_zpass_tma(const Fft3d &f, const void *in, void *out, ROp op, cudaStream_t st, size_t row0 = 0,
                             size_t nrows = 0) {
  constexpr int E = ZShape<N>::E, TR = ZShape<N>::TR;
  constexpr int MINB = AUX ? 1 : ZShape<N>::MINB;
  constexpr int NSTAGE = 3;
  constexpr int threads = TR * (N / 2 / E);
  using Z = ZTile<N, TR, NSTAGE, AUX>;
  constexpr int smem = Z::smem_bytes;
  static_assert(smem <= 227 * 1024, "z-pass ring does not fit");
  auto kern = fft_zpass_tma<N, E, TR, NSTAGE, C2R, AUX, MINB>;
  static int blocks_per_sm_dev[kMaxDevices] = {};
  int &blocks_per_sm = blocks_per_sm_dev[f.device % kMaxDevices];
  if (!blocks_per_sm) {
    BGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    blocks_per_sm = occ > 0 ? occ : 1;
  }
  // rows [row0, row0 + nrows) of the (x, y) row set; all of them by default
  if (!nrows) nrows = (size_t)f.Ns * N;
  const size_t real_off = row0 * N * sizeof(double), cplx_off = row0 * (N / 2 + 1) * sizeof(double2);
  const char *pin = static_cast<const char *>(in) + (C2R ? cplx_off : real_off);
  char *pout = static_cast<char *>(out) + (C2R ? real_off : cplx_off);
  if (op.aux) op.aux += row0 * N;
  const int ntiles = (int)(nrows / TR);
  int blocks = f.sm_count * blocks_per_sm;
  if (blocks > ntiles) blocks = ntiles;
  ProfScope prof(C2R ? KK_FFT_C2R_Z : KK_FFT_R2C_Z, st);
  // a pass over a row range (the streaming host API) keeps the plain order
  const int rev = (row0 == 0 && nrows == (size_t)f.Ns * N) ? f.next_rev() : 0;
  launch_pass(kern, blocks, threads, smem, st, f.use_pdl, (const void *)pin, (void *)pout, f.twN, f.twM, op, ntiles, rev);
  BGPU_LAUNCHED(1);
}

template <int N>
static bool try_r2c_zpass_tma(const Fft3d &f, const double *in, double2 *out, ROp lop) {
  if constexpr (!ztma_has_size<N>()) {
    return false;
  } else {
    if (!f.use_tma || !(lop.kind == R_LOAD || lop.kind == R_LOAD_SCALE)) return false;
    const ChunkHooks *hk = f.hooks;
    if (hk && hk->before && hk->chunks > 1 && ((size_t)f.Ns * N) % ((size_t)hk->chunks * ZShape<N>::TR) == 0) {
      // the caller streams the input in: wait for each slab of rows just before its z pass
      const size_t per = (size_t)f.Ns * N / hk->chunks;
      for (int c = 0; c < hk->chunks; ++c) {
        hk->before(hk->ctx, c);
        launch_zpass_tma<N, false, false>(f, in, out, lop, f.stream, c * per, per);
      }
      return true;
    }
    if (hk && hk->before)
      for (int c = 0; c < hk->chunks; ++c) hk->before(hk->ctx, c);
    launch_zpass_tma<N, false, false>(f, in, out, lop, f.stream);
    return true;
  }
}

template <int N>
static bool try_c2r_zpass_tma(const Fft3d &f, const double2 *in, double *out, ROp sop) {
  if constexpr (!ztma_has_size<N>()) {
    return false;
  } else {
    if (!f.use_tma) return false;
    const ChunkHooks *hk = f.hooks;
    if (sop.kind == R_SCALE_MUL) {
      launch_zpass_tma<N, true, true>(f, in, out, sop, f.stream);
    } else if (sop.kind == R_SCALE || sop.kind == R_AXPY) {
      if (hk && hk->after && hk->chunks > 1 && ((size_t)f.Ns * N) % ((size_t)hk->chunks * ZShape<N>::TR) == 0) {
        // the caller streams the result out: hand over each slab of rows as soon as its z pass is queued
        const size_t per = (size_t)f.Ns * N / hk->chunks;
        for (int c = 0; c < hk->chunks; ++c) {
          launch_zpass_tma<N, true, false>(f, in, out, sop, f.stream, c * per, per);
          hk->after(hk->ctx, c);
        }
        return true;
      }
      launch_zpass_tma<N, true, false>(f, in, out, sop, f.stream);
    } else {
      return false;
    }
    if (hk && hk->after)
      for (int c = 0; c < hk->chunks; ++c) hk->after(hk->ctx, c);
    return true;
  }
}

// ---------------------------------------------------------------------------
// fused z+y passes (fft_fused.cuh): one persistent kernel, intermediate kept in L2
// ---------------------------------------------------------------------------
template <int N> struct ZyShape { static constexpr bool ok = false; static constexpr int EZ = 8, TR = 8, EY = 8; };
template <> struct ZyShape<256> { static constexpr bool ok = true; static constexpr int EZ = 8, TR = 16, EY = 8; };

template <int N>
static bool zy_fused_ok(const Fft3d &f) {
  return ZyShape<N>::ok && f.use_tma && f.use_fused && f.G == 1 && !f.hooks && f.zy_ready;
}

template <int N, bool C2R, bool AUX>
static void launch_zy_fused(const Fft3d &f, const void *zsrc, void *zdst, const double2 *cplx, ROp op) {
  if constexpr (ZyShape<N>::ok) {
    using S = ZyShape<N>;
    constexpr int NS = AUX ? 2 : 3;
    using L = ZySmem<N, S::TR, NS, NS, AUX>;
    auto kern = fft_zy_fused<N, S::EZ, S::TR, S::EY, NS, NS, C2R, AUX>;
    static bool configured_dev[kMaxDevices] = {};
    bool &configured = configured_dev[f.device % kMaxDevices];
    if (!configured) {
      BGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::bytes));
      int occ = 0;
      BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 512, L::bytes));
      if (occ < 1) throw std::runtime_error("bgpu: the fused z+y kernel does not fit on this device");
      configured = true;
    }
    const CUtensorMap &ymap = f.tensor_map(cplx, 1, true, 0, N, N);
    constexpr int per_plane = C2R ? (N / 2 + 1 + 7) / 8 : N / S::TR;  // producer tiles per plane
    f.zy_epoch += per_plane;
    ZyCtl ctl{f.zy_ready, f.zy_epoch, f.zy_lead};
    ProfScope prof(C2R ? KK_FFT_ZY_C2R : KK_FFT_ZY_R2C, f.stream);
    // every CTA must be resident (the roles wait on each other across CTAs): one per SM
    kern<<<f.sm_count, 512, L::bytes, f.stream>>>(ymap, zsrc, zdst, f.twN, f.twM, op, ctl);
    BGPU_LAUNCHED(1);
  }
}

template <int N>
static bool r2c_zy_fused(const Fft3d &f, const double *in, double2 *out, ROp lop) {
  if (!zy_fused_ok<N>(f) || !(lop.kind == R_LOAD || lop.kind == R_LOAD_SCALE)) return false;
  launch_zy_fused<N, false, false>(f, in, out, out, lop);
  return true;
}

template <int N>
static bool c2r_yz_fused(const Fft3d &f, double2 *work, double *out, ROp sop) {
  if (!zy_fused_ok<N>(f)) return false;
  if (sop.kind == R_SCALE_MUL) launch_zy_fused<N, true, true>(f, work, out, work, sop);
  else if (sop.kind == R_SCALE || sop.kind == R_AXPY) launch_zy_fused<N, true, false>(f, work, out, work, sop);
  else return false;
  return true;
}

// all-to-all of the packed buffers: peer h gets / gives block h (Ns*Ns*(N/2+1) complex numbers)
static void slab_all_to_all(const Fft3d &f, const double2 *send, double2 *recv) {
  const size_t blk = (size_t)f.Ns * f.Ns * (f.N / 2 + 1);
  ProfScope prof(KK_ALLTOALL, f.stream);
  f.comm->all_to_all(send, recv, blk * 2, f.stream);
}

template <int N>
static void r2c_impl(const Fft3d &f, const double *in, double2 *out, double2 *xout, ROp lop, KOp sop) {
  const size_t nrows = (size_t)f.Ns * N;
  if (r2c_zy_fused<N>(f, in, out, lop)) {
    launch_strided<N, -1, 0>(f, out, xout ? xout : out, f.twN, KOp{}, sop, f.stream);
    return;
  }
  if (!try_r2c_zpass_tma<N>(f, in, out, lop)) {
  if (f.G > 1) throw std::runtime_error("bgpu: the slab-decomposed transform needs the bulk-copy z pass (N >= 128)");
  if (f.hooks && f.hooks->before)
    for (int c = 0; c < f.hooks->chunks; ++c) f.hooks->before(f.hooks->ctx, c);
  ProfScope prof(KK_FFT_R2C_Z, f.stream);
  if constexpr (N == 8) {
    tiny_r2c_zpass<N><<<(unsigned)((nrows + 63) / 64), 64, 0, f.stream>>>(in, out, f.twN, lop, nrows);
  } else {
    constexpr int TR = Shape<N>::TR;
    constexpr int M = N / 2;
    constexpr size_t smem = (size_t)TR * (M + M / 8 + 1) * sizeof(double2);
    fft_r2c_zpass<N, TR><<<(unsigned)(nrows / TR), TR * M / 8, smem, f.stream>>>(in, out, f.twN, f.twM, lop, nrows);
  }
  BGPU_LAUNCHED(1);
  }
  if (f.G == 1) {
    launch_strided<N, -1, 1>(f, out, out, f.twN, KOp{}, KOp{}, f.stream);
    launch_strided<N, -1, 0>(f, out, xout ? xout : out, f.twN, KOp{}, sop, f.stream);
    return;
  }
  // slab: y pass on my x planes, stored straight into the packed send buffer; all-to-all; x pass on
  // the transposed layout [x][y_local][z] (= the receive buffer, block h holding the x planes of rank h)
  PassIo y_io;
  y_io.n_other = f.Ns;
  y_io.G = f.G;
  y_io.Ns = f.Ns;
  PassIo x_io;
  x_io.n_other = f.Ns;
  x_io.other0 = f.rank * f.Ns;
  if (f.p2p && tma_has_size<N>() && !f.force_generic) {
    // fused transpose: the y pass stores every tile straight into the peers' receive buffers (TMA over
    // NVLink); one cross-rank barrier, then the x pass reads what the peers wrote here.  Receive
    // buffers alternate between transforms, so the barrier of this transform also orders the next
    // transform's writes after every rank's reads of that buffer.
    double2 *const *peers = f.peer_recv[f.parity];
    y_io.peer_out = peers;
    y_io.my_rank = f.rank;
    launch_strided<N, -1, 1>(f, out, nullptr, f.twN, KOp{}, KOp{}, f.stream, y_io);
    f.barrier();
    launch_strided<N, -1, 0>(f, peers[f.rank], xout ? xout : out, f.twN, KOp{}, sop, f.stream, x_io);
    f.parity ^= 1;
    return;
  }
  y_io.out_packed = true;
  launch_strided<N, -1, 1>(f, out, f.sendbuf, f.twN, KOp{}, KOp{}, f.stream, y_io);
  slab_all_to_all(f, f.sendbuf, f.recvbuf);
  launch_strided<N, -1, 0>(f, f.recvbuf, xout ? xout : out, f.twN, KOp{}, sop, f.stream, x_io);
}

template <int N>
static void c2r_impl(const Fft3d &f, const double2 *in, double2 *work, double *out, KOp lop, ROp sop) {
  const size_t nrows = (size_t)f.Ns * N;
  if (f.G == 1) {
    launch_strided<N, +1, 0>(f, in, work, f.twN, lop, KOp{}, f.stream);
    if (c2r_yz_fused<N>(f, work, out, sop)) return;
    launch_strided<N, +1, 1>(f, work, work, f.twN, KOp{}, KOp{}, f.stream);
  } else {
    // slab: x pass on the transposed layout into the send buffer (block h = the x planes of rank h,
    // contiguous); all-to-all; y pass reads the packed receive buffer and writes my x planes
    PassIo x_io;
    x_io.n_other = f.Ns;
    x_io.other0 = f.rank * f.Ns;
    x_io.G = f.G;
    x_io.Ns = f.Ns;
    PassIo y_io;
    y_io.n_other = f.Ns;
    y_io.in_packed = true;
    y_io.G = f.G;
    y_io.Ns = f.Ns;
    if (f.p2p && tma_has_size<N>() && !f.force_generic) {
      double2 *const *peers = f.peer_recv[f.parity];
      x_io.peer_out = peers;
      x_io.my_rank = f.rank;
      launch_strided<N, +1, 0>(f, in, nullptr, f.twN, lop, KOp{}, f.stream, x_io);
      f.barrier();
      launch_strided<N, +1, 1>(f, peers[f.rank], work, f.twN, KOp{}, KOp{}, f.stream, y_io);
      f.parity ^= 1;
    } else {
      launch_strided<N, +1, 0>(f, in, f.sendbuf, f.twN, lop, KOp{}, f.stream, x_io);
      slab_all_to_all(f, f.sendbuf, f.recvbuf);
      launch_strided<N, +1, 1>(f, f.recvbuf, work, f.twN, KOp{}, KOp{}, f.stream, y_io);
    }
  }
  if (try_c2r_zpass_tma<N>(f, work, out, sop)) return;
  if (f.G > 1) throw std::runtime_error("bgpu: the slab-decomposed transform needs the bulk-copy z pass (N >= 128)");
  ProfScope prof(KK_FFT_C2R_Z, f.stream);
  if constexpr (N == 8) {
    tiny_c2r_zpass<N><<<(unsigned)((nrows + 63) / 64), 64, 0, f.stream>>>(work, out, f.twN, sop, nrows);
  } else {
    constexpr int TR = Shape<N>::TR;
    constexpr int M = N / 2;
    constexpr size_t smem = (size_t)TR * (M + M / 8 + 1) * sizeof(double2);
    fft_c2r_zpass<N, TR><<<(unsigned)(nrows / TR), TR * M / 8, smem, f.stream>>>(work, out, f.twN, f.twM, sop, nrows);
  }
  BGPU_LAUNCHED(1);
  if (f.hooks && f.hooks->after)
    for (int c = 0; c < f.hooks->chunks; ++c) f.hooks->after(f.hooks->ctx, c);
}

// opt in to > 48 KB dynamic shared memory; per device, so done at plan creation
template <int N>
static void configure_impl(Fft3d &f) {
  constexpr int T = Shape<N>::T;
  constexpr int smem_s = N * T * (int)sizeof(double2);
  if constexpr (N >= 128) {
    constexpr int smem_p = 2 * smem_s;
    BGPU_CUDA(cudaFuncSetAttribute(fft_strided_pass_pipelined<N, T, -1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
    BGPU_CUDA(cudaFuncSetAttribute(fft_strided_pass_pipelined<N, T, -1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
    BGPU_CUDA(cudaFuncSetAttribute(fft_strided_pass_pipelined<N, T, +1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
    BGPU_CUDA(cudaFuncSetAttribute(fft_strided_pass_pipelined<N, T, +1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
    int dev = 0, sms = 0, occ = 0, occ_min = 1 << 30;
    BGPU_CUDA(cudaGetDevice(&dev));
    BGPU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_strided_pass_pipelined<N, T, +1, 0>, T * N / 8, smem_p));
    occ_min = occ < occ_min ? occ : occ_min;
    BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_strided_pass_pipelined<N, T, -1, 0>, T * N / 8, smem_p));
    occ_min = occ < occ_min ? occ : occ_min;
    BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_strided_pass_pipelined<N, T, +1, 1>, T * N / 8, smem_p));
    occ_min = occ < occ_min ? occ : occ_min;
    f.strided_blocks = sms * (occ_min > 0 ? occ_min : 1);
  }
  BGPU_CUDA(cudaFuncSetAttribute(fft_strided_pass<N, T, -1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s));
  BGPU_CUDA(cudaFuncSetAttribute(fft_strided_pass<N, T, -1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s));
  BGPU_CUDA(cudaFuncSetAttribute(fft_strided_pass<N, T, +1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s));
  BGPU_CUDA(cudaFuncSetAttribute(fft_strided_pass<N, T, +1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s));
  if constexpr (N > 8) {
    constexpr int TR = Shape<N>::TR;
    constexpr int M = N / 2;
    constexpr int smem_z = TR * (M + M / 8 + 1) * (int)sizeof(double2);
    BGPU_CUDA(cudaFuncSetAttribute(fft_r2c_zpass<N, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_z));
    BGPU_CUDA(cudaFuncSetAttribute(fft_c2r_zpass<N, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_z));
  }
}

static std::vector<double2> make_twiddles(int n) {
  std::vector<double2> tw(n);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < n; ++k) {
    const long double a = two_pi * (long double)k / (long double)n;
    tw[k] = make_double2((double)cosl(a), (double)-sinl(a));
  }
  // exact values on the axes and diagonals
  for (int k = 0; k < n; ++k) {
    if ((4 * k) % n == 0) {
      const int q = (4 * k) / n;
      tw[k] = make_double2(q == 0 ? 1.0 : (q == 2 ? -1.0 : 0.0), q == 1 ? -1.0 : (q == 3 ? 1.0 : 0.0));
    }
  }
  return tw;
}

#define BGPU_DISPATCH_N(CALL)                     \
  switch (N) {                                    \
    case 8: CALL(8); break;                       \
    case 16: CALL(16); break;                     \
    case 32: CALL(32); break;                     \
    case 64: CALL(64); break;                     \
    case 128: CALL(128); break;                   \
    case 256: CALL(256); break;                   \
    case 512: CALL(512); break;                   \
    case 1024: CALL(1024); break;                 \
    default: throw std::runtime_error("bgpu: unsupported FFT size"); \
  }

void Fft3d::init(int n, cudaStream_t st) {
  if (!supported(n)) throw std::runtime_error("bgpu: FFT size must be a power of two in [8, 1024], got " + std::to_string(n));
  N = n;
  if (G == 1) Ns = n;
  stream = st;
  {
    int dev = 0;
    BGPU_CUDA(cudaGetDevice(&dev));
    device = dev;
    BGPU_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    const char *e = std::getenv("BGPU_FFT_TMA");
    use_tma = !(e && e[0] == '0');
    maps_.clear();
    // fused z+y kernel (fft_fused.cuh).  Opt-in: measured on B200 at 256^3 it is correct but ~25 % slower than
    // the two separate passes -- the passes are bound by shared-memory bandwidth and
    // latency on the SM, not by HBM, so halving the HBM traffic does not pay by itself.
    const char *fg = std::getenv("BGPU_FFT_SLAB_GENERIC");
    force_generic = fg && fg[0] == '1';
    const char *sx = std::getenv("BGPU_SHARE_X");
    share_x = !(sx && sx[0] == '0');  // default since round 2 (measured +7 % at 256^3, parity-green); BGPU_SHARE_X=0 = three x passes
    const char *pd = std::getenv("BGPU_PDL");
    // measured (B200, round 2): +1.3 % per evaluation at 256^3, where a pass is ~55 us and the launch ramp shows;
    // -2 % at 512^3 -- so on by default up to 256 only (BGPU_PDL=0 / 1 overrides)
    use_pdl = pd ? pd[0] != '0' : n <= 256;
    const char *pp = std::getenv("BGPU_PINGPONG");
    // measured (B200, round 2, 256^3): +1.1 % per evaluation (y passes -2 %); nothing to gain once the arrays are many
    // times the L2 (512^3: 1.07 GB against 126 MB) -- on by default up to 256 (BGPU_PINGPONG=0 / 1 overrides)
    pingpong = pp ? pp[0] != '0' : n <= 256;
    const char *sxs = std::getenv("BGPU_SHARE_X_SLAB");
    share_x_slab = !(sxs && sxs[0] == '0');
    const char *zr = std::getenv("BGPU_ZROUND");
    z_round = !(zr && zr[0] == '0');
    const char *tw2 = std::getenv("BGPU_FFT_2WARP");
    two_warp = tw2 && tw2[0] == '1';
    const char *fu = std::getenv("BGPU_FFT_FUSED");
    use_fused = fu && fu[0] == '1';
    const char *ld = std::getenv("BGPU_FFT_LEAD");
    zy_lead = ld ? std::atoi(ld) : 96;
    if (zy_lead < 48) zy_lead = 48;  // must exceed the planes spanned by a role's ring (3 tiles of stride SMs/tiles-per-plane)
    BGPU_CUDA(cudaMalloc(&zy_ready, sizeof(unsigned long long) * n));
    BGPU_CUDA(cudaMemset(zy_ready, 0, sizeof(unsigned long long) * n));
    zy_epoch = 0;
  }
  auto a = make_twiddles(n), b = make_twiddles(n / 2);
  BGPU_CUDA(cudaMalloc(&twN, sizeof(double2) * n));
  BGPU_CUDA(cudaMalloc(&twM, sizeof(double2) * (n / 2)));
  BGPU_CUDA(cudaMemcpy(twN, a.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
  BGPU_CUDA(cudaMemcpy(twM, b.data(), sizeof(double2) * (n / 2), cudaMemcpyHostToDevice));
#define CALL(n_) configure_impl<n_>(*this)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::destroy() {
  if (twN) cudaFree(twN);
  if (twM) cudaFree(twM);
  twN = twM = nullptr;
  if (zy_ready) cudaFree(zy_ready);
  zy_ready = nullptr;
}

void Fft3d::barrier() const {
  ProfScope prof(KK_ALLTOALL, stream);
  comm->barrier(stream);
}

bool Fft3d::supported(int n) { return n >= 8 && n <= 1024 && (n & (n - 1)) == 0; }

// ---------------------------------------------------------------------------
// shared x pass (BGPU_SHARE_X=1): single passes with the extended functors (fft_tma.cuh RotCtxX, AUX = -1)
// ---------------------------------------------------------------------------
bool Fft3d::can_share_x() const { return share_x && G == 1 && use_tma && !use_fused && (N == 128 || N == 256 || N == 512); }

template <int N>
static void xpass_impl(const Fft3d &f, const double2 *in, double2 *out, int dir, KOp lop, KOp sop) {
  if constexpr (tma_has_size<N>()) {
    if (dir > 0)
      launch_strided_tma<N, +1, 0, -1>(f, in, out, lop, sop, PassIo{}, f.stream);
    else
      launch_strided_tma<N, -1, 0, -1>(f, in, out, lop, sop, PassIo{}, f.stream);
  } else {
    throw std::runtime_error("bgpu: the shared x pass needs the TMA-staged strided pass (N = 128, 256 or 512)");
  }
}

template <int N>
static void c2r_yz_impl(const Fft3d &f, const double2 *in, double2 *work, double *out, KOp ylop, ROp sop) {
  if constexpr (tma_has_size<N>()) {
    launch_strided_tma<N, +1, 1, -1>(f, in, work, ylop, KOp{}, PassIo{}, f.stream);
    if (!try_c2r_zpass_tma<N>(f, work, out, sop)) throw std::runtime_error("bgpu: shared x pass: no bulk-copy z pass");
  } else {
    throw std::runtime_error("bgpu: the shared x pass needs the TMA-staged strided pass (N = 128, 256 or 512)");
  }
}

template <int N>
static void r2c_zy_impl(const Fft3d &f, const double *in, double2 *work, double2 *yout, ROp lop, KOp ysop) {
  if constexpr (tma_has_size<N>()) {
    if (!try_r2c_zpass_tma<N>(f, in, work, lop)) throw std::runtime_error("bgpu: shared x pass: no bulk-copy z pass");
    launch_strided_tma<N, -1, 1, -1>(f, work, yout, KOp{}, ysop, PassIo{}, f.stream);
  } else {
    throw std::runtime_error("bgpu: the shared x pass needs the TMA-staged strided pass (N = 128, 256 or 512)");
  }
}

template <int N>
static void ypass_impl(const Fft3d &f, const double2 *in, double2 *out, int dir, KOp lop, KOp sop) {
  if constexpr (tma_has_size<N>()) {
    if (dir > 0)
      launch_strided_tma<N, +1, 1, -1>(f, in, out, lop, sop, PassIo{}, f.stream);
    else
      launch_strided_tma<N, -1, 1, -1>(f, in, out, lop, sop, PassIo{}, f.stream);
  } else {
    throw std::runtime_error("bgpu: single y passes need the TMA-staged strided pass (N = 128, 256 or 512)");
  }
}

template <int N>
static void zround_impl(const Fft3d &f, double2 *work, ROp op) {
  if constexpr (tma_has_size<N>()) {
    constexpr int E = ZShape<N>::E, TR = ZShape<N>::TR, NSTAGE = 3;
    constexpr int threads = TR * (N / 2 / E);
    using Z = ZTile<N, TR, NSTAGE, true>;
    constexpr int smem = Z::smem_bytes;
    static_assert(smem <= 227 * 1024, "z round-trip ring does not fit");
    auto kern = fft_zround_tma<N, E, TR, NSTAGE, 1>;
    static int blocks_per_sm_dev[kMaxDevices] = {};
    int &blocks_per_sm = blocks_per_sm_dev[f.device % kMaxDevices];
    if (!blocks_per_sm) {
      BGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int occ = 0;
      BGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
      blocks_per_sm = occ > 0 ? occ : 1;
    }
    const int ntiles = (int)((size_t)f.Ns * N / TR);
    int blocks = f.sm_count * blocks_per_sm;
    if (blocks > ntiles) blocks = ntiles;
    ProfScope prof(KK_FFT_ZROUND, f.stream);
    launch_pass(kern, blocks, threads, smem, f.stream, f.use_pdl, (const double2 *)work, work, f.twN, f.twM, op, ntiles,
                f.next_rev());
    BGPU_LAUNCHED(1);
  } else {
    throw std::runtime_error("bgpu: the z round trip needs the bulk-copy z pass (N = 128, 256 or 512)");
  }
}

bool Fft3d::can_zround() const { return z_round && can_share_x(); }

// ---------------------------------------------------------------------------
// shared x pass on slabs, inverse direction (fft3d.h)
// ---------------------------------------------------------------------------
bool Fft3d::can_share_x_slab() const {
  return share_x && share_x_slab && G > 1 && use_tma && !force_generic && (N == 128 || N == 256 || N == 512);
}

template <int N>
static void xpass_shared_inverse_impl(const Fft3d &f, const double2 *in, KOp lop) {
  if constexpr (tma_has_size<N>()) {
    // as the x pass of c2r_impl: pencils along x on the transposed layout [x][y_local][z], block h of the output =
    // the x planes of rank h
    PassIo x_io;
    x_io.n_other = f.Ns;
    x_io.other0 = f.rank * f.Ns;
    x_io.G = f.G;
    x_io.Ns = f.Ns;
    if (f.p2p) {
      double2 *const *peers = f.peer_recv[f.parity];
      x_io.peer_out = peers;
      x_io.my_rank = f.rank;
      launch_strided_tma<N, +1, 0, -1>(f, in, nullptr, lop, KOp{}, x_io, f.stream);
      f.barrier();
      f.shared_recv = peers[f.rank];
      // the buffers alternate per transpose: the transform after the component passes writes the other one, and the
      // one after that writes this one again only behind a barrier every rank enters after its component passes
      f.parity ^= 1;
    } else {
      launch_strided_tma<N, +1, 0, -1>(f, in, f.sendbuf, lop, KOp{}, x_io, f.stream);
      slab_all_to_all(f, f.sendbuf, f.recvbuf);
      f.shared_recv = f.recvbuf;
    }
  } else {
    throw std::runtime_error("bgpu: the shared x pass needs the TMA-staged strided pass (N = 128, 256 or 512)");
  }
}

template <int N>
static void c2r_yz_shared_impl(const Fft3d &f, double2 *work, double *out, KOp ylop, ROp sop) {
  if constexpr (tma_has_size<N>()) {
    if (!f.shared_recv) throw std::runtime_error("bgpu: c2r_yz_shared without xpass_shared_inverse");
    PassIo y_io;
    y_io.n_other = f.Ns;
    y_io.in_packed = true;
    y_io.G = f.G;
    y_io.Ns = f.Ns;
    launch_strided_tma<N, +1, 1, -1>(f, f.shared_recv, work, ylop, KOp{}, y_io, f.stream);
    if (!try_c2r_zpass_tma<N>(f, work, out, sop)) throw std::runtime_error("bgpu: shared x pass: no bulk-copy z pass");
  } else {
    throw std::runtime_error("bgpu: the shared x pass needs the TMA-staged strided pass (N = 128, 256 or 512)");
  }
}

void Fft3d::xpass_shared_inverse(const double2 *in, KOp lop) const {
#define CALL(n) xpass_shared_inverse_impl<n>(*this, in, lop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::c2r_yz_shared(double2 *work, double *out, KOp ylop, ROp sop) const {
#define CALL(n) c2r_yz_shared_impl<n>(*this, work, out, ylop, sop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::ypass(const double2 *in, double2 *out, int dir, KOp lop, KOp sop) const {
#define CALL(n) ypass_impl<n>(*this, in, out, dir, lop, sop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::zround(double2 *work, ROp sop) const {
#define CALL(n) zround_impl<n>(*this, work, sop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::xpass(const double2 *in, double2 *out, int dir, KOp lop, KOp sop) const {
#define CALL(n) xpass_impl<n>(*this, in, out, dir, lop, sop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::c2r_yz(const double2 *in, double2 *work, double *out, KOp ylop, ROp sop) const {
#define CALL(n) c2r_yz_impl<n>(*this, in, work, out, ylop, sop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::r2c_zy(const double *in, double2 *work, double2 *yout, ROp lop, KOp ysop) const {
#define CALL(n) r2c_zy_impl<n>(*this, in, work, yout, lop, ysop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::r2c(const double *in, double2 *out, double2 *xout, ROp lop, KOp sop) const {
#define CALL(n) r2c_impl<n>(*this, in, out, xout, lop, sop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

void Fft3d::c2r(const double2 *in, double2 *work, double *out, KOp lop, ROp sop) const {
#define CALL(n) c2r_impl<n>(*this, in, work, out, lop, sop)
  BGPU_DISPATCH_N(CALL)
#undef CALL
}

}  // namespace bgpu
