// NCCL, bound at run time: the single-GPU path (and a machine without NCCL)
// never loads it; a multi-GPU context dlopens libnccl.so.2 — inside a process
// that already imported torch this resolves to the copy torch loaded.  Only the
// entry points the Schur step uses are bound (nccl.h 2.27: ncclGetUniqueId, ncclCommInitRank,
// ncclCommDestroy, ncclGetErrorString, ncclAllReduce, ncclBroadcast).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

struct NcclApi
{
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t)
    = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)
    = nullptr;
  std::string error;
  bool load()
  {
    if(handle)
      return true;
    for(const char *name : {"libnccl.so.2", "libnccl.so"})
      {
        handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if(handle)
          break;
      }
    if(!handle)
      {
        error = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
      }
#define SDPB_BIND(field, sym)                                                 \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));              \
  if(!field)                                                                  \
    {                                                                         \
      error = std::string("libnccl.so.2 lacks ") + sym;                       \
      return false;                                                           \
    }
    SDPB_BIND(GetUniqueId, "ncclGetUniqueId")
    SDPB_BIND(CommInitRank, "ncclCommInitRank")
    SDPB_BIND(CommDestroy, "ncclCommDestroy")
    SDPB_BIND(GetErrorString, "ncclGetErrorString")
    SDPB_BIND(AllReduce, "ncclAllReduce")
    SDPB_BIND(Broadcast, "ncclBroadcast")
#undef SDPB_BIND
    return true;
  }
};
inline NcclApi &nccl_api()
{
  static NcclApi api;
  return api;
}
