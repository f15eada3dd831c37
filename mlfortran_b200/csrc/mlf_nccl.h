// NCCL without a link-time dependency: libnccl.so.2 is dlopen'ed on first use, so single-GPU callers never need it and
// one-process-per-GPU callers (C, Fortran + MPI, Python) get the per-iteration statistics all-reduce on the engine's
// stream without a host callback.  Only the handful of entry points this path needs are bound.
#pragma once
#include <cstddef>
#include <cstdint>

namespace mlfb200 {

struct NcclApi {
  struct UniqueId { char b[128]; };   // ncclUniqueId, passed by value
  // prototypes as in nccl.h (2.x): ncclResult_t = int (0 = ncclSuccess), ncclComm_t = opaque pointer,
  // ncclUniqueId = 128 bytes, ncclFloat64 = 8, ncclSum = 0
  int (*GetUniqueId)(void* id) = nullptr;
  int (*CommInitRank)(void** comm, int nranks, UniqueId id, int rank) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*CommCount)(void* comm, int* count) = nullptr;
  int (*CommUserRank)(void* comm, int* rank) = nullptr;
  int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void* stream) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

// nullptr when libnccl.so.2 cannot be loaded (message on stderr once)
const NcclApi* nccl_api();

}  // namespace mlfb200
