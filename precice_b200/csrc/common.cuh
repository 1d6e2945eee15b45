// common.cuh — shared host/device plumbing of the sm_100a RBF mapping library.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rbfmap.h"

namespace rbfmap {

struct MapError : std::runtime_error {
  int code;
  MapError(int c, const std::string &msg)
      : std::runtime_error(msg), code(c) {}
};

#define RBF_CUDA(call)                                                                                                            \
  do {                                                                                                                            \
    cudaError_t e_ = (call);                                                                                                      \
    if (e_ != cudaSuccess)                                                                                                        \
      throw ::rbfmap::MapError(RBFMAP_ERR_CUDA, std::string("CUDA failure: ") + cudaGetErrorString(e_) + " (" #call ") at " __FILE__ ":" + \
                                                    std::to_string(__LINE__));                                                    \
  } while (0)

// math/differences.hpp:8
constexpr double kZeroDiff = 1.0e-14;

// Stream on which DevBuf allocates and frees (stream-ordered, from the device's default memory pool whose release
// threshold is raised so that freed blocks are reused by the next computeMapping instead of going back to the
// driver). Set by the mapping objects for the duration of each call.
extern thread_local cudaStream_t g_allocStream;
void configureMemoryPool();
struct AllocStreamScope {
  cudaStream_t prev;
  explicit AllocStreamScope(cudaStream_t s)
      : prev(g_allocStream)
  {
    g_allocStream = s;
  }
  ~AllocStreamScope() { g_allocStream = prev; }
};

// Device buffer with stream-ordered allocation; owns its memory.
template <typename T>
struct DevBuf {
  T           *p = nullptr;
  size_t       n = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf &)            = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept
      : p(o.p), n(o.n), s(o.s)
  {
    o.p = nullptr;
    o.n = 0;
  }
  DevBuf &operator=(DevBuf &&o) noexcept
  {
    if (this != &o) {
      release();
      p   = o.p;
      n   = o.n;
      s   = o.s;
      o.p = nullptr;
      o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count)
  {
    release();
    n = count;
    s = g_allocStream;
    if (count > 0)
      RBF_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&p), sizeClass(count * sizeof(T)), s));
  }
  // Requests of 1 MB and more are rounded up to m 2^k with m in 8..15 (12.5 % steps). A remeshing step asks for slightly
  // different sizes than the step before; with exact sizes the stream-ordered pool rarely finds a freed block that fits
  // and maps new memory instead (measured: the 2M-vertex surface step varied between 8.1 and 9.8 ms from run to run).
  static size_t sizeClass(size_t bytes)
  {
    if (bytes < (size_t(1) << 20))
      return bytes;
    int k = 0;
    while ((bytes >> k) > 15)
      ++k;
    const size_t m = (bytes + (size_t(1) << k) - 1) >> k; // 8..16
    return m << k;
  }
  void release()
  {
    if (p)
      cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// Grow-only scratch outside the stream-ordered pool (plain cudaMalloc): for buffers that a mapping object keeps across
// clear(). Kept in the pool they sit between the blocks that every step releases and re-allocates (the 9 GB factor
// buffer, the bins), and the pool then re-creates those blocks each step (measured: +4 ... +120 ms per step).
template <typename T>
struct ScratchBuf {
  T     *p = nullptr;
  size_t n = 0;
  ScratchBuf() = default;
  ScratchBuf(const ScratchBuf &)            = delete;
  ScratchBuf &operator=(const ScratchBuf &) = delete;
  ~ScratchBuf()
  {
    if (p)
      cudaFree(p);
  }
  void reserve(size_t count) // contents are not preserved
  {
    if (count <= n)
      return;
    if (p)
      RBF_CUDA(cudaFree(p));
    p = nullptr;
    n = 0;
    RBF_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T)));
    n = count;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// Makes `device` current for the lifetime of the object and restores the previous device (device < 0: no-op).
struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int device)
  {
    if (device < 0)
      return;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != device) {
      if (cudaSetDevice(device) != cudaSuccess)
        throw MapError(RBFMAP_ERR_CUDA, "rbfmap: cannot select the mapping's CUDA device");
      prev = cur;
    }
  }
  ~DeviceScope()
  {
    if (prev >= 0)
      cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope &)            = delete;
  DeviceScope &operator=(const DeviceScope &) = delete;
};

inline int divUp(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// number of SMs of the current device (cached per device)
int smCount();

// Exclusive prefix sums (own kernels, scan.cu): out has n+1 entries, out[n] = total.
void exclusiveScan(const int *in, int *out, int64_t n, cudaStream_t s);
void exclusiveScan(const int *in, int64_t *out, int64_t n, cudaStream_t s);
void exclusiveScan(const int64_t *in, int64_t *out, int64_t n, cudaStream_t s);

// counts kernels launched by this library on the calling thread (for bench.py's gpu_launches)
extern thread_local int g_launches;
#define RBF_LAUNCHED() (++::rbfmap::g_launches)

// ---- device helpers ----
#ifdef __CUDACC__
// order-preserving map double -> uint64 so that atomicMin/atomicMax work on doubles
__device__ __forceinline__ unsigned long long encodeOrdered(double d)
{
  unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(d));
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double decodeOrdered(unsigned long long k)
{
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
  return __longlong_as_double(static_cast<long long>(b));
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}

// RadialBasisFctSolver.hpp:102-113 with all axes active; no FMA contraction so that the value is
// bit-identical to the reference's (a-b)*(a-b) accumulated left to right.
__device__ __forceinline__ double sqDist3(double ax, double ay, double az, double bx, double by, double bz)
{
  const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
  double       s  = __dmul_rn(dx, dx);
  s               = __dadd_rn(s, __dmul_rn(dy, dy));
  s               = __dadd_rn(s, __dmul_rn(dz, dz));
  return s;
}
// with an activity mask (1.0 / 0.0 per axis): (u - v) * active
__device__ __forceinline__ double sqDist3Masked(double ax, double ay, double az, double bx, double by, double bz, double mx, double my, double mz)
{
  const double dx = __dmul_rn(__dsub_rn(ax, bx), mx), dy = __dmul_rn(__dsub_rn(ay, by), my), dz = __dmul_rn(__dsub_rn(az, bz), mz);
  double       s  = __dmul_rn(dx, dx);
  s               = __dadd_rn(s, __dmul_rn(dy, dy));
  s               = __dadd_rn(s, __dmul_rn(dz, dz));
  return s;
}
#endif

} // namespace rbfmap
