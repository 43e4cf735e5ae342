// Multi-GPU exchange at the C-ABI level: NCCL over NVLink, used only for the once-per-flush gathers of
// the event-sharded run (posterior histograms all-gather, proposal-counter all-reduce -- the reference's
// own end-of-run MPI_Reduce, src/cls_parallel.f90:265-268).  NCCL is loaded with dlopen, so a single-GPU
// user needs no NCCL at all; the 128-byte unique id is created by one shard and distributed by the HOST
// program (MPI_Bcast in the Fortran driver, a file or torch.distributed elsewhere).
#include <dlfcn.h>

#include <cstring>
#include <string>

#include "htm_kernels.hpp"

namespace htm {

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi& api() {
  static NcclApi a;
  return a;
}

const char* nccl_load(std::string* why) {
  NcclApi& a = api();
  if (a.lib) return nullptr;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) {
    *why = std::string("cannot load libnccl.so.2: ") + dlerror();
    return why->c_str();
  }
  a.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(a.lib, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(a.lib, "ncclCommInitRank"));
  a.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(a.lib, "ncclAllGather"));
  a.AllReduce =
      reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(a.lib, "ncclAllReduce"));
  a.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(a.lib, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(a.lib, "ncclGetErrorString"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.AllReduce || !a.CommDestroy) {
    *why = "libnccl.so.2 lacks a required symbol";
    a.lib = nullptr;
    return why->c_str();
  }
  return nullptr;
}

static std::string nccl_err(int rc) {
  NcclApi& a = api();
  return a.GetErrorString ? a.GetErrorString(rc) : ("nccl error " + std::to_string(rc));
}

bool nccl_unique_id(char id[128], std::string* why) {
  if (nccl_load(why)) return false;
  NcclId u;
  const int rc = api().GetUniqueId(&u);
  if (rc != 0) {
    *why = nccl_err(rc);
    return false;
  }
  std::memcpy(id, u.internal, 128);
  return true;
}

bool nccl_init(void** comm, const char id[128], int rank, int nranks, std::string* why) {
  if (nccl_load(why)) return false;
  NcclId u;
  std::memcpy(u.internal, id, 128);
  const int rc = api().CommInitRank(comm, nranks, u, rank);
  if (rc != 0) {
    *why = "ncclCommInitRank: " + nccl_err(rc);
    return false;
  }
  return true;
}

void nccl_destroy(void* comm) {
  if (comm && api().CommDestroy) api().CommDestroy(comm);
}

// ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0): uint32 = 3, int64 = 4, uint64 = 5, sum = 0
bool nccl_allgather_u32(void* comm, const void* send, void* recv, size_t count, cudaStream_t s, std::string* why) {
  const int rc = api().AllGather(send, recv, count, 3, comm, s);
  if (rc != 0) *why = "ncclAllGather: " + nccl_err(rc);
  return rc == 0;
}
bool nccl_allreduce_f64(void* comm, const void* send, void* recv, size_t count, cudaStream_t s, std::string* why) {
  const int rc = api().AllReduce(send, recv, count, 8, 0, comm, s);  // ncclFloat64 = 8, ncclSum = 0
  if (rc != 0) *why = "ncclAllReduce: " + nccl_err(rc);
  return rc == 0;
}
bool nccl_allreduce_u64(void* comm, const void* send, void* recv, size_t count, cudaStream_t s, std::string* why) {
  const int rc = api().AllReduce(send, recv, count, 5, 0, comm, s);
  if (rc != 0) *why = "ncclAllReduce: " + nccl_err(rc);
  return rc == 0;
}

}  // namespace htm
