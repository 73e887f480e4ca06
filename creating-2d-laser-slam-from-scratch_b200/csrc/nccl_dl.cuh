// NCCL reached through dlopen: the library has no link-time dependency on it (single-GPU hosts never load it), and a caller
// that already carries a libnccl (torch's bundled copy, the system's) keeps using its own.  Only ncclAllReduce and
// ncclGetErrorString are bound; the communicator is the caller's (ncclCommInitRank in the host application).
// Enumerator values as in nccl.h (2.x): ncclSum 0, ncclMax 2; ncclInt32 2, ncclUint32 3, ncclFloat64 8.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <mutex>

namespace b2s {

struct NcclApi {
  typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
  typedef const char *(*errstr_fn)(int);
  allreduce_fn all_reduce = nullptr;
  errstr_fn error_string = nullptr;
  bool ok = false;
};
constexpr int NCCL_SUM = 0, NCCL_MAX = 2, NCCL_INT32 = 2, NCCL_UINT32 = 3, NCCL_FLOAT64 = 8;

inline const NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);  // resolves to an already loaded copy if there is one
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.all_reduce = reinterpret_cast<NcclApi::allreduce_fn>(dlsym(h, "ncclAllReduce"));
    api.error_string = reinterpret_cast<NcclApi::errstr_fn>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.all_reduce != nullptr;
  });
  return api;
}

}  // namespace b2s
