// nccl_dyn.cuh — NCCL entry points resolved at run time (dlopen), so that librbfmap.so has no link-time dependency and
// shares the NCCL copy the host application (e.g. PyTorch) already loaded. Used by the mappings that shard ONE mapping
// over the GPUs of a node (rbfmap_distributed_init): rbf-pum-direct (all-gather of the uploaded mesh / value slices,
// all-reduce of weight sums and partial blends in exchange mode, the grouped agreement on errors and cluster count,
// reduce-scatter of the result for rbfmap_map_sliced) and the CG of rbf-global-iterative (all-gather of the direction
// vector slices, all-reduce of the dot products), all over NVLink.
#pragma once

#include <nccl.h>

#include "common.cuh"

#include <cstring>

namespace rbfmap {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *)                                                                           = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                                                    = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                                               = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)        = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)                     = nullptr;
  ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)    = nullptr;
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

// The group of processes (one GPU each) that share one mapping: created by rbfmap_distributed_init from the 128-byte id
// that rank 0 obtained from rbfmap_nccl_unique_id. world == 1 means "not distributed": every helper is then a no-op.
struct ShardGroup {
  ncclComm_t comm = nullptr;
  int        rank = 0, world = 1;

  ShardGroup()                              = default;
  ShardGroup(const ShardGroup &)            = delete;
  ShardGroup &operator=(const ShardGroup &) = delete;
  ~ShardGroup() { destroy(); }
  void destroy()
  {
    if (comm)
      ncclApi().CommDestroy(comm);
    comm  = nullptr;
    rank  = 0;
    world = 1;
  }
  void init(const char *uniqueId /*128 bytes*/, int rank_, int world_)
  {
    if (world_ < 1 || rank_ < 0 || rank_ >= world_)
      throw MapError(RBFMAP_ERR_INVALID, "invalid rank / world size");
    destroy();
    rank  = rank_;
    world = world_;
    // uniqueId == nullptr: a group without communicator. The ranks shard the work but cannot exchange anything, so only
    // collective-free sharding works (rbf-pum-direct, duplicate mode); every helper below is then a no-op.
    if (world > 1 && uniqueId) {
      ncclUniqueId id;
      memcpy(&id, uniqueId, sizeof(id));
      RBF_NCCL(ncclApi().CommInitRank(&comm, world, id, rank));
    }
  }
  bool active() const { return world > 1; }
  bool connected() const { return world > 1 && comm != nullptr; }
  // length of the per-rank slice and of the padded buffer that all-gathers need for `count` elements
  int64_t sliceOf(int64_t count) const { return (count + world - 1) / world; }
  int64_t padded(int64_t count) const { return connected() ? sliceOf(count) * world : count; }
  void    allReduceSum(double *v, int64_t count, cudaStream_t s) const
  {
    if (connected() && count > 0)
      RBF_NCCL(ncclApi().AllReduce(v, v, static_cast<size_t>(count), ncclDouble, ncclSum, comm, s));
  }
  void allReduceSum(int *v, int64_t count, cudaStream_t s) const
  {
    if (connected() && count > 0)
      RBF_NCCL(ncclApi().AllReduce(v, v, static_cast<size_t>(count), ncclInt32, ncclSum, comm, s));
  }
  void allReduceMin(int *v, int64_t count, cudaStream_t s) const
  {
    if (connected() && count > 0)
      RBF_NCCL(ncclApi().AllReduce(v, v, static_cast<size_t>(count), ncclInt32, ncclMin, comm, s));
  }
  // several collectives in one launch
  void groupStart() const
  {
    if (connected())
      RBF_NCCL(ncclApi().GroupStart());
  }
  void groupEnd() const
  {
    if (connected())
      RBF_NCCL(ncclApi().GroupEnd());
  }
  // v holds world * slice elements on every rank; afterwards the slice [rank * slice, (rank + 1) * slice) of v is the sum
  // over the ranks (in place), the other slices are unspecified
  void reduceScatterSumInPlace(double *v, int64_t slice, cudaStream_t s) const
  {
    if (connected() && slice > 0)
      RBF_NCCL(ncclApi().ReduceScatter(v, v + rank * slice, static_cast<size_t>(slice), ncclDouble, ncclSum, comm, s));
  }
  // every rank holds its slice [rank * slice, (rank + 1) * slice) of v (padded(count) elements); afterwards all of v
  void allGatherInPlace(double *v, int64_t slice, cudaStream_t s) const
  {
    if (connected() && slice > 0)
      RBF_NCCL(ncclApi().AllGather(v + rank * slice, v, static_cast<size_t>(slice), ncclDouble, comm, s));
  }
};

// contiguous row slice of rank `rank` out of `world` (equal slice length ceil(n / world); the last slices may be short or empty)
inline void partitionRows(int64_t n, int world, int rank, int64_t &begin, int64_t &end, int64_t &sliceLen)
{
  sliceLen = (n + world - 1) / world;
  begin    = std::min<int64_t>(n, sliceLen * rank);
  end      = std::min<int64_t>(n, begin + sliceLen);
}

} // namespace rbfmap
