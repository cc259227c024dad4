// Partitioned vmult with the whole schedule on the C++ side: pack, NCCL
// send/recv groups on a communication stream, the three cell partitions and
// the unpack are issued by ONE C-ABI call (a few microseconds of host time),
// so that small per-GPU problems are not bound by host launch latency.
// Device side of update_ghost_values / compress(add) of
// LinearAlgebra::distributed::Vector inside CUDAWrappers::MatrixFree::cell_loop
// (benchmark_03.h:348-353).  NCCL is resolved with dlopen at run time (the
// copy torch already loaded), so libmfhn.so keeps loading on machines without it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <string>
#include <vector>

namespace mfhn
{
struct NcclApi
{
  typedef int (*GetUniqueId_t)(void *);
  struct Id
  {
    char internal[128];
  };
  typedef int (*CommInitRank_t)(void **, int, Id, int);
  typedef int (*CommDestroy_t)(void *);
  typedef int (*SendRecv_t)(void *, size_t, int, int, void *, cudaStream_t);
  typedef int (*Group_t)();
  typedef int (*AllReduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
  typedef const char *(*ErrStr_t)(int);
  GetUniqueId_t GetUniqueId = nullptr;
  CommInitRank_t CommInitRank = nullptr;
  CommDestroy_t CommDestroy = nullptr;
  SendRecv_t Send = nullptr, Recv = nullptr;
  Group_t GroupStart = nullptr, GroupEnd = nullptr;
  AllReduce_t AllReduce = nullptr;
  ErrStr_t GetErrorString = nullptr;
  bool ok = false;
  std::string error;

  static NcclApi &get()
  {
    static NcclApi api = load();
    return api;
  }
  static NcclApi load()
  {
    NcclApi a;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
      {
        a.error = std::string("cannot load libnccl.so.2: ") + dlerror();
        return a;
      }
    a.GetUniqueId    = (GetUniqueId_t)dlsym(h, "ncclGetUniqueId");
    a.CommInitRank   = (CommInitRank_t)dlsym(h, "ncclCommInitRank");
    a.CommDestroy    = (CommDestroy_t)dlsym(h, "ncclCommDestroy");
    a.Send           = (SendRecv_t)dlsym(h, "ncclSend");
    a.Recv           = (SendRecv_t)dlsym(h, "ncclRecv");
    a.GroupStart     = (Group_t)dlsym(h, "ncclGroupStart");
    a.GroupEnd       = (Group_t)dlsym(h, "ncclGroupEnd");
    a.AllReduce      = (AllReduce_t)dlsym(h, "ncclAllReduce");
    a.GetErrorString = (ErrStr_t)dlsym(h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.GroupStart && a.GroupEnd && a.GetErrorString && a.AllReduce;
    if (!a.ok) a.error = "libnccl.so.2 lacks a required symbol";
    return a;
  }
};

} // namespace mfhn
