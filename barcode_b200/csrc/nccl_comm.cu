// barcode_b200/csrc/nccl_comm.cu -- see nccl_comm.h
#include "nccl_comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace bgpu {

namespace {

struct Api {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

template <class F>
void bind(void *lib, F &fn, const char *name) {
  fn = reinterpret_cast<F>(dlsym(lib, name));
  if (!fn) throw std::runtime_error(std::string("bgpu: libnccl.so.2 has no symbol ") + name);
}

const Api &api() {
  static Api a;
  static bool loaded = false;
  if (!loaded) {
    // RTLD_NOLOAD first: inside a PyTorch process this picks up the NCCL torch already loaded
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw std::runtime_error(std::string("bgpu: cannot open libnccl.so.2 (") + dlerror() + ")");
    bind(lib, a.GetUniqueId, "ncclGetUniqueId");
    bind(lib, a.CommInitRank, "ncclCommInitRank");
    bind(lib, a.CommDestroy, "ncclCommDestroy");
    bind(lib, a.GroupStart, "ncclGroupStart");
    bind(lib, a.GroupEnd, "ncclGroupEnd");
    bind(lib, a.Send, "ncclSend");
    bind(lib, a.Recv, "ncclRecv");
    bind(lib, a.AllReduce, "ncclAllReduce");
    bind(lib, a.GetErrorString, "ncclGetErrorString");
    loaded = true;
  }
  return a;
}

void check(ncclResult_t r, const char *what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + api().GetErrorString(r));
}

}  // namespace

void NcclComm::unique_id(void *out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  check(api().GetUniqueId(static_cast<ncclUniqueId *>(out128)), "ncclGetUniqueId");
}

NcclComm::NcclComm(const void *id128, int rank_, int nranks_) {
  rank = rank_;
  nranks = nranks_;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  check(api().CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
  comm = c;
  cudaMalloc(reinterpret_cast<void **>(&token), sizeof(double));
  cudaMemset(token, 0, sizeof(double));
}

NcclComm::~NcclComm() {
  if (comm) api().CommDestroy(static_cast<ncclComm_t>(comm));
  if (token) cudaFree(token);
}

void NcclComm::barrier(cudaStream_t st) {
  check(api().AllReduce(token, token, 1, ncclDouble, ncclSum, static_cast<ncclComm_t>(comm), st), "ncclAllReduce");
}

void NcclComm::all_to_all(const void *send, void *recv, size_t count, cudaStream_t st) {
  const Api &a = api();
  ncclComm_t c = static_cast<ncclComm_t>(comm);
  const double *s = static_cast<const double *>(send);
  double *r = static_cast<double *>(recv);
  check(a.GroupStart(), "ncclGroupStart");
  for (int h = 0; h < nranks; ++h) {
    check(a.Send(s + (size_t)h * count, count, ncclDouble, h, c, st), "ncclSend");
    check(a.Recv(r + (size_t)h * count, count, ncclDouble, h, c, st), "ncclRecv");
  }
  check(a.GroupEnd(), "ncclGroupEnd");
}

void NcclComm::all_reduce_sum(double *buf, size_t count, cudaStream_t st) {
  check(api().AllReduce(buf, buf, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(comm), st), "ncclAllReduce");
}

void NcclComm::all_reduce_max(double *buf, size_t count, cudaStream_t st) {
  check(api().AllReduce(buf, buf, count, ncclDouble, ncclMax, static_cast<ncclComm_t>(comm), st), "ncclAllReduce");
}

void NcclComm::exchange2(const double *a_, int to_a, const double *b_, int to_b, double *ra, int from_a, double *rb,
                         int from_b, size_t count, cudaStream_t st) {
  const Api &a = api();
  ncclComm_t c = static_cast<ncclComm_t>(comm);
  check(a.GroupStart(), "ncclGroupStart");
  check(a.Send(a_, count, ncclDouble, to_a, c, st), "ncclSend");
  check(a.Send(b_, count, ncclDouble, to_b, c, st), "ncclSend");
  check(a.Recv(ra, count, ncclDouble, from_a, c, st), "ncclRecv");
  check(a.Recv(rb, count, ncclDouble, from_b, c, st), "ncclRecv");
  check(a.GroupEnd(), "ncclGroupEnd");
}

void NcclComm::shift(const double *send, int to, double *recv, int from, size_t count, cudaStream_t st) {
  const Api &a = api();
  ncclComm_t c = static_cast<ncclComm_t>(comm);
  check(a.GroupStart(), "ncclGroupStart");
  check(a.Send(send, count, ncclDouble, to, c, st), "ncclSend");
  check(a.Recv(recv, count, ncclDouble, from, c, st), "ncclRecv");
  check(a.GroupEnd(), "ncclGroupEnd");
}

// ---------------------------------------------------------------------------
// LocalComm: slab ranks as threads of one process on one device (see nccl_comm.h)
// ---------------------------------------------------------------------------
struct LocalGroup {
  int n = 0;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long generation = 0;
  struct Slot {
    const void *p[2] = {nullptr, nullptr};
    int to[2] = {-1, -1};
  };
  std::vector<Slot> slot;
  // all ranks arrive, then all leave (a classic generation barrier)
  void rendezvous() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned long gen = generation;
    if (++waiting == n) {
      waiting = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};

LocalGroup *local_group_create(int nranks) {
  if (nranks < 1 || nranks > 8) throw std::runtime_error("bgpu: a local slab group has 1 to 8 ranks");
  LocalGroup *g = new LocalGroup;
  g->n = nranks;
  g->slot.resize(nranks);
  return g;
}
void local_group_destroy(LocalGroup *g) { delete g; }

void local_group_exchange_ptr(LocalGroup *g, int rank, void *mine, void **all) {
  g->slot[rank].p[0] = mine;
  g->rendezvous();
  for (int r = 0; r < g->n; ++r) all[r] = const_cast<void *>(g->slot[r].p[0]);
  g->rendezvous();
}

namespace {
__global__ void local_reduce_kernel(double *out, const double *const *in, int n, size_t count, int op) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double v = in[0][i];
  for (int r = 1; r < n; ++r) {  // fixed rank order: every rank computes the same bits
    const double w = in[r][i];
    v = op == 0 ? v + w : (w > v ? w : v);
  }
  out[i] = v;
}
void cuda_ok(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in LocalComm ") + what + ": " + cudaGetErrorString(e));
}
}  // namespace

LocalComm::LocalComm(LocalGroup *g, int rank_, int nranks_) : group(g) {
  rank = rank_;
  nranks = nranks_;
  if (!g || g->n != nranks_) throw std::runtime_error("bgpu: local slab group of the wrong size");
}
LocalComm::~LocalComm() {
  if (stage) cudaFree(stage);
}

void LocalComm::barrier(cudaStream_t st) {
  cuda_ok(cudaStreamSynchronize(st), "barrier");
  group->rendezvous();
}

void LocalComm::all_to_all(const void *send, void *recv, size_t count, cudaStream_t st) {
  cuda_ok(cudaStreamSynchronize(st), "all_to_all");
  group->slot[rank].p[0] = send;
  group->rendezvous();
  double *r = static_cast<double *>(recv);
  for (int h = 0; h < nranks; ++h)  // block `rank` of rank h's send buffer is mine
    cuda_ok(cudaMemcpyAsync(r + (size_t)h * count, static_cast<const double *>(group->slot[h].p[0]) + (size_t)rank * count,
                            count * sizeof(double), cudaMemcpyDeviceToDevice, st), "all_to_all copy");
  cuda_ok(cudaStreamSynchronize(st), "all_to_all");
  group->rendezvous();
}

void LocalComm::reduce(double *buf, size_t count, int op, cudaStream_t st) {
  if (count > stage_count) {
    if (stage) cudaFree(stage);
    stage_count = count < 8192 ? 8192 : count;
    cuda_ok(cudaMalloc(reinterpret_cast<void **>(&stage), (stage_count + 16) * sizeof(double)), "stage");
  }
  cuda_ok(cudaStreamSynchronize(st), "all_reduce");
  group->slot[rank].p[0] = buf;
  group->rendezvous();
  const double *host_ptrs[8];
  for (int r = 0; r < nranks; ++r) host_ptrs[r] = static_cast<const double *>(group->slot[r].p[0]);
  // the pointer table rides behind the staged result (16 spare doubles = 8 pointers + slack)
  const double **d_ptrs = reinterpret_cast<const double **>(stage + stage_count);
  cuda_ok(cudaMemcpyAsync(d_ptrs, host_ptrs, nranks * sizeof(double *), cudaMemcpyHostToDevice, st), "ptrs");
  local_reduce_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(stage, d_ptrs, nranks, count, op);
  cuda_ok(cudaStreamSynchronize(st), "all_reduce");
  group->rendezvous();  // everyone has read everyone's input: overwrite mine
  cuda_ok(cudaMemcpyAsync(buf, stage, count * sizeof(double), cudaMemcpyDeviceToDevice, st), "all_reduce store");
}
void LocalComm::all_reduce_sum(double *buf, size_t count, cudaStream_t st) { reduce(buf, count, 0, st); }
void LocalComm::all_reduce_max(double *buf, size_t count, cudaStream_t st) { reduce(buf, count, 1, st); }

void LocalComm::exchange2(const double *a_, int to_a, const double *b_, int to_b, double *ra, int from_a, double *rb,
                          int from_b, size_t count, cudaStream_t st) {
  cuda_ok(cudaStreamSynchronize(st), "exchange");
  LocalGroup::Slot &mine = group->slot[rank];
  mine.p[0] = a_;
  mine.to[0] = to_a;
  mine.p[1] = b_;
  mine.to[1] = to_b;
  group->rendezvous();
  // NCCL pairs the k-th receive from a rank with that rank's k-th send to me (matters when both neighbours are
  // the same rank, G = 2)
  double *dst[2] = {ra, rb};
  const int from[2] = {from_a, from_b};
  for (int j = 0; j < 2; ++j) {
    if (!dst[j] || from[j] < 0) continue;
    int nth = 0;
    for (int e = 0; e < j; ++e)
      if (dst[e] && from[e] == from[j]) ++nth;
    const LocalGroup::Slot &src = group->slot[from[j]];
    const void *p = nullptr;
    for (int e = 0, seen = 0; e < 2; ++e)
      if (src.p[e] && src.to[e] == rank) {
        if (seen == nth) p = src.p[e];
        ++seen;
      }
    if (!p) throw std::runtime_error("bgpu: LocalComm exchange without a matching send");
    cuda_ok(cudaMemcpyAsync(dst[j], p, count * sizeof(double), cudaMemcpyDeviceToDevice, st), "exchange copy");
  }
  cuda_ok(cudaStreamSynchronize(st), "exchange");
  group->rendezvous();
  mine.p[0] = mine.p[1] = nullptr;
  mine.to[0] = mine.to[1] = -1;
}

void LocalComm::shift(const double *send, int to, double *recv, int from, size_t count, cudaStream_t st) {
  exchange2(send, to, nullptr, -1, recv, from, nullptr, -1, count, st);
}

}  // namespace bgpu
