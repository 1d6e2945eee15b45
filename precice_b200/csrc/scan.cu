// scan.cu — exclusive prefix sums used for CSR offsets (cell starts, cluster offsets, packed-matrix
// offsets). Replaces the Kokkos parallel_scan of device/KokkosPUMKernels_Impl.hpp:278-303.
// Two-level scan: per-block scan of kItems elements, scan of the block totals, uniform add.
#include "common.cuh"

#include <atomic>

namespace rbfmap {

thread_local int          g_launches    = 0;
thread_local cudaStream_t g_allocStream = nullptr;

void configureMemoryPool()
{
  int dev = 0;
  RBF_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  RBF_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  unsigned long long threshold = ~0ull;
  RBF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
}

int smCount()
{
  // per device, race-free: the attribute query is idempotent and the slots are atomics
  static std::atomic<int> cached[64];
  int                     dev = 0;
  RBF_CUDA(cudaGetDevice(&dev));
  const int slot = dev < 64 ? dev : 63;
  int       v    = dev < 64 ? cached[slot].load(std::memory_order_relaxed) : 0;
  if (v == 0) {
    RBF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (dev < 64)
      cached[slot].store(v, std::memory_order_relaxed);
  }
  return v;
}

namespace {
constexpr int kThreads = 256;
constexpr int kPerThr  = 8;
constexpr int kItems   = kThreads * kPerThr;

template <typename TO>
__device__ TO blockExclusive(TO v, TO *total)
{
  // exclusive scan of one value per thread across the block
  __shared__ TO warpSums[kThreads / 32];
  const int     lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  TO            inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    TO t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o)
      inc += t;
  }
  if (lane == 31)
    warpSums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    TO w  = lane < kThreads / 32 ? warpSums[lane] : TO(0);
    TO wi = w;
#pragma unroll
    for (int o = 1; o < kThreads / 32; o <<= 1) {
      TO t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o)
        wi += t;
    }
    if (lane < kThreads / 32)
      warpSums[lane] = wi - w; // exclusive warp offsets
    if (lane == kThreads / 32 - 1)
      *total = wi;
  }
  __syncthreads();
  return inc - v + warpSums[wid];
}

template <typename TI, typename TO>
__global__ void scanBlocks(const TI *__restrict__ in, TO *__restrict__ out, TO *__restrict__ blockSums, int64_t n)
{
  __shared__ TO total;
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kItems + static_cast<int64_t>(threadIdx.x) * kPerThr;
  TO            v[kPerThr];
  TO            sum = 0;
#pragma unroll
  for (int k = 0; k < kPerThr; ++k) {
    const int64_t i = base + k;
    v[k]            = i < n ? static_cast<TO>(in[i]) : TO(0);
    sum += v[k];
  }
  TO off = blockExclusive<TO>(sum, &total);
#pragma unroll
  for (int k = 0; k < kPerThr; ++k) {
    const int64_t i = base + k;
    if (i < n)
      out[i] = off;
    off += v[k];
  }
  if (threadIdx.x == 0)
    blockSums[blockIdx.x] = total;
}

template <typename TO>
__global__ void scanBlockSums(TO *__restrict__ blockSums, int nBlocks, TO *__restrict__ grandTotal)
{
  // single block, sequential over chunks of kThreads
  __shared__ TO total;
  TO            carry = 0;
  for (int base = 0; base < nBlocks; base += kThreads) {
    const int i   = base + threadIdx.x;
    const TO  v   = i < nBlocks ? blockSums[i] : TO(0);
    const TO  off = blockExclusive<TO>(v, &total);
    if (i < nBlocks)
      blockSums[i] = carry + off;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *grandTotal = carry;
}

template <typename TO>
__global__ void addBlockOffsets(TO *__restrict__ out, const TO *__restrict__ blockSums, int64_t n)
{
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kItems;
  const TO      off  = blockSums[blockIdx.x];
  for (int k = threadIdx.x; k < kItems; k += kThreads) {
    const int64_t i = base + k;
    if (i < n)
      out[i] += off;
  }
}

template <typename TI, typename TO>
void scanImpl(const TI *in, TO *out, int64_t n, cudaStream_t s)
{
  if (n <= 0) {
    RBF_CUDA(cudaMemsetAsync(out, 0, sizeof(TO), s));
    return;
  }
  const int  nBlocks = divUp(n, kItems);
  DevBuf<TO> sums;
  sums.alloc(nBlocks);
  scanBlocks<TI, TO><<<nBlocks, kThreads, 0, s>>>(in, out, sums.p, n);
  RBF_LAUNCHED();
  scanBlockSums<TO><<<1, kThreads, 0, s>>>(sums.p, nBlocks, out + n);
  RBF_LAUNCHED();
  addBlockOffsets<TO><<<nBlocks, kThreads, 0, s>>>(out, sums.p, n);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError()); // `sums` is released stream-ordered
}
} // namespace

void exclusiveScan(const int *in, int *out, int64_t n, cudaStream_t s) { scanImpl<int, int>(in, out, n, s); }
void exclusiveScan(const int *in, int64_t *out, int64_t n, cudaStream_t s) { scanImpl<int, long long>(in, reinterpret_cast<long long *>(out), n, s); }
void exclusiveScan(const int64_t *in, int64_t *out, int64_t n, cudaStream_t s)
{
  scanImpl<long long, long long>(reinterpret_cast<const long long *>(in), reinterpret_cast<long long *>(out), n, s);
}

} // namespace rbfmap
