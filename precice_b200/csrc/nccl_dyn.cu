#include "nccl_dyn.cuh"

#include <dlfcn.h>

#include <mutex>

namespace rbfmap {

const NcclApi &ncclApi()
{
  static NcclApi        api;
  static std::once_flag once;
  static std::string    error;
  std::call_once(once, [] {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
      h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      error = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    auto sym = [&](const char *name) {
      void *p = dlsym(h, name);
      if (!p)
        error = std::string("libnccl lacks ") + name;
      return p;
    };
    api.GetUniqueId    = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank   = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy    = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce      = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather      = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.ReduceScatter  = reinterpret_cast<decltype(api.ReduceScatter)>(sym("ncclReduceScatter"));
    api.GroupStart     = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd       = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!error.empty())
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbfmap: NCCL is not available (" + error + ")");
  return api;
}

} // namespace rbfmap
