// barcode_b200/csrc/nccl_comm.cu -- see nccl_comm.h
#include "nccl_comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <stdexcept>
#include <string>

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

NcclComm::NcclComm(const void *id128, int rank_, int nranks_) : rank(rank_), nranks(nranks_) {
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

}  // namespace bgpu
