#include "mlf_nccl.h"

#include <cstdio>
#include <dlfcn.h>
#include <mutex>

namespace mlfb200 {

const NcclApi* nccl_api() {
  static NcclApi api;
  static bool ok = false;
  static std::once_flag once;
  std::call_once(once, [] {
    // a process that already holds NCCL (torch bundles its own libnccl.so.2) gets that copy: dlopen matches the soname
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      fprintf(stderr, "mlfortran_b200: libnccl.so.2 not found (%s): use mlf_b200_set_allreduce with your own collective\n", dlerror());
      return;
    }
    auto sym = [&](const char* n) { return dlsym(h, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.CommCount = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
    api.CommUserRank = reinterpret_cast<decltype(api.CommUserRank)>(sym("ncclCommUserRank"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.CommCount && api.CommUserRank && api.AllReduce;
    if (!ok) fprintf(stderr, "mlfortran_b200: libnccl.so.2 lacks an expected symbol\n");
  });
  return ok ? &api : nullptr;
}

}  // namespace mlfb200
