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

struct NcclComm : SlabComm {
  int rank = 0, nranks = 1;
  void *comm = nullptr;  // ncclComm_t
  double *token = nullptr;  // device scalar the barrier reduces

  // 128-byte ncclUniqueId; call on one rank and distribute (MPI, a file, torch.distributed ...)
  static void unique_id(void *out128);
  NcclComm(const void *id128, int rank, int nranks);
  ~NcclComm() override;

  void all_to_all(const void *send, void *recv, size_t count_doubles, cudaStream_t st) override;
  void barrier(cudaStream_t st) override;  // a one-element all-reduce on `st`
  void all_reduce_sum(double *buf, size_t count, cudaStream_t st);  // in place
  void all_reduce_max(double *buf, size_t count, cudaStream_t st);  // in place
  // one grouped neighbour exchange: send a -> rank `to_a`, b -> `to_b`; receive ra <- `from_a`, rb <- `from_b`
  void exchange2(const double *a, int to_a, const double *b, int to_b, double *ra, int from_a, double *rb, int from_b,
                 size_t count, cudaStream_t st);
  // one-directional ring shift: send -> rank `to`, receive <- rank `from`
  void shift(const double *send, int to, double *recv, int from, size_t count, cudaStream_t st);
};

}  // namespace bgpu
