// barcode_b200/csrc/nccl_comm.h -- the communication layer of the slab-decomposed chain
// (SURVEY 8e): all-to-all transposes of the distributed FFT, neighbour exchange of the
// mass-assignment halo, and the scalar reductions (sum rho, -lnL, prior, kinetic energy), all
// stream-ordered NCCL calls over NVLink.  libnccl.so.2 is opened at run time (dlopen), so the
// single-GPU library has no NCCL dependency.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "fft3d.h"

namespace bgpu {

// what a slab-decomposed chain needs from its communicator (api.cu); every call is collective and stream-ordered
struct ChainComm : SlabComm {
  int rank = 0, nranks = 1;
  virtual void all_reduce_sum(double *buf, size_t count, cudaStream_t st) = 0;  // in place
  virtual void all_reduce_max(double *buf, size_t count, cudaStream_t st) = 0;  // in place
  // one grouped neighbour exchange: send a -> rank `to_a`, b -> `to_b`; receive ra <- `from_a`, rb <- `from_b`
  virtual void exchange2(const double *a, int to_a, const double *b, int to_b, double *ra, int from_a, double *rb,
                         int from_b, size_t count, cudaStream_t st) = 0;
  // one-directional ring shift: send -> rank `to`, receive <- rank `from`
  virtual void shift(const double *send, int to, double *recv, int from, size_t count, cudaStream_t st) = 0;
};

struct NcclComm : ChainComm {
  void *comm = nullptr;  // ncclComm_t
  double *token = nullptr;  // device scalar the barrier reduces

  // 128-byte ncclUniqueId; call on one rank and distribute (MPI, a file, torch.distributed ...)
  static void unique_id(void *out128);
  NcclComm(const void *id128, int rank, int nranks);
  ~NcclComm() override;

  void all_to_all(const void *send, void *recv, size_t count_doubles, cudaStream_t st) override;
  void barrier(cudaStream_t st) override;  // a one-element all-reduce on `st`
  void all_reduce_sum(double *buf, size_t count, cudaStream_t st) override;
  void all_reduce_max(double *buf, size_t count, cudaStream_t st) override;
  void exchange2(const double *a, int to_a, const double *b, int to_b, double *ra, int from_a, double *rb, int from_b,
                 size_t count, cudaStream_t st) override;
  void shift(const double *send, int to, double *recv, int from, size_t count, cudaStream_t st) override;
};

// ---------------------------------------------------------------------------
// The same collectives between slab ranks that live in ONE process on ONE device, each driven by its own host
// thread and stream (bgpu_local_group_create / bgpu_slab_create_local).  NCCL refuses two ranks on one GPU, so a
// single-GPU box could never run the slab-decomposed code path; this communicator lets its tests run there (and
// lets a debugger see all ranks at once).  Not a performance path: every collective is
//   stream sync -> host rendezvous (pointers published) -> device copies / reduction on the caller's stream ->
//   stream sync -> host rendezvous (peers may reuse their buffers).
// ---------------------------------------------------------------------------
struct LocalGroup;  // shared by the ranks: rendezvous + published pointers
LocalGroup *local_group_create(int nranks);
void local_group_destroy(LocalGroup *g);
// every rank's pointer for one key, as published by that rank (collective; used to hand out receive buffers)
void local_group_exchange_ptr(LocalGroup *g, int rank, void *mine, void **all);

struct LocalComm : ChainComm {
  LocalGroup *group;
  double *stage = nullptr;  // result of a reduction until the peers have finished reading the inputs
  size_t stage_count = 0;
  LocalComm(LocalGroup *g, int rank, int nranks);
  ~LocalComm() override;
  void all_to_all(const void *send, void *recv, size_t count_doubles, cudaStream_t st) override;
  void barrier(cudaStream_t st) override;
  void all_reduce_sum(double *buf, size_t count, cudaStream_t st) override;
  void all_reduce_max(double *buf, size_t count, cudaStream_t st) override;
  void exchange2(const double *a, int to_a, const double *b, int to_b, double *ra, int from_a, double *rb, int from_b,
                 size_t count, cudaStream_t st) override;
  void shift(const double *send, int to, double *recv, int from, size_t count, cudaStream_t st) override;

 private:
  void reduce(double *buf, size_t count, int op, cudaStream_t st);
};

}  // namespace bgpu
