// nccl_dyn.cuh — NCCL entry points resolved at run time (dlopen), so that librbfmap.so has no link-time dependency and
// shares the NCCL copy the host application (e.g. PyTorch) already loaded. Used only by the sharded matrix-free CG of
// rbf-global-iterative: all-gather of the direction vector slices and all-reduce of the dot products over NVLink.
#pragma once

#include <nccl.h>

#include "common.cuh"

namespace rbfmap {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *)                                                                           = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                                                    = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                                               = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)        = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)                     = nullptr;
  ncclResult_t (*GroupStart)()                                                                                          = nullptr;
  ncclResult_t (*GroupEnd)()                                                                                            = nullptr;
  const char *(*GetErrorString)(ncclResult_t)                                                                           = nullptr;
};

// loads libnccl.so.2 on first use; throws MapError(RBFMAP_ERR_UNSUPPORTED) if it cannot be found
const NcclApi &ncclApi();

#define RBF_NCCL(call)                                                                                                                    \
  do {                                                                                                                                    \
    ncclResult_t r_ = (call);                                                                                                             \
    if (r_ != ncclSuccess)                                                                                                                \
      throw ::rbfmap::MapError(RBFMAP_ERR_CUDA, std::string("NCCL failure: ") + ::rbfmap::ncclApi().GetErrorString(r_) + " (" #call ")"); \
  } while (0)

// contiguous row slice of rank `rank` out of `world` (equal slice length ceil(n / world); the last slices may be short or empty)
inline void partitionRows(int64_t n, int world, int rank, int64_t &begin, int64_t &end, int64_t &sliceLen)
{
  sliceLen = (n + world - 1) / world;
  begin    = std::min<int64_t>(n, sliceLen * rank);
  end      = std::min<int64_t>(n, begin + sliceLen);
}

} // namespace rbfmap
