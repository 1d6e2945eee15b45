// fp64_pipe.cu — DFMA / DMMA throughput on one SM-full of warps as a function of warps per SM and independent chains
// per thread (how much instruction- and thread-level parallelism the FP64 pipe of this GPU needs).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe fp64_pipe.cu ; prints a table.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void dfma(double *out, int iters, double a, double b)
{
  double x[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k)
    x[k] = threadIdx.x * 1e-9 + k;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int k = 0; k < CH; ++k)
        x[k] = fma(x[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CH; ++k)
    s += x[k];
  if (s == 12345.678)
    out[0] = s;
}

template <int CH>
__global__ void dmma(double *out, int iters, double a, double b)
{
  double c0[CH], c1[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k)
    c0[k] = c1[k] = threadIdx.x * 1e-9 + k;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int k = 0; k < CH; ++k)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CH; ++k)
    s += c0[k] + c1[k];
  if (s == 12345.678)
    out[0] = s;
}

template <typename K>
float timeIt(K kernel, int blocks, int threads, double *out, int iters)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float t;
    cudaEventElapsedTime(&t, a, b);
    if (rep > 0 && t < best)
      best = t;
  }
  return best;
}

int main()
{
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 2048;
  printf("SMs %d, clock %.0f MHz\n", sms, khz / 1000.0);
  printf("%-6s %-6s %-8s %12s %16s %22s\n", "op", "warps", "chains", "TFLOP/s", "cyc/inst/SMSP", "cyc per dependent op");
  for (int warps : {4, 8, 12, 16, 32, 64}) {
    const int threads = warps >= 8 ? 256 : warps * 32, blocks = sms * (warps * 32 / threads);
#define RUN(OP, CH, FMAS)                                                                                                      \
  {                                                                                                                            \
    const float  ms     = timeIt(OP<CH>, blocks, threads, out, iters);                                                         \
    const double inst   = double(iters) * 16 * CH;                /* per warp */                                               \
    const double cycles = ms * 1e-3 * khz * 1e3;                                                                               \
    const double perSmsp = cycles / (inst * warps / 4.0);                                                                      \
    printf("%-6s %-6d %-8d %12.2f %16.2f %22.1f\n", #OP, warps, CH, 2.0 * FMAS * inst * warps * sms / (ms * 1e-3) / 1e12, perSmsp, \
           cycles / (double(iters) * 16));                                                                                     \
  }
    RUN(dfma, 1, 32) RUN(dfma, 2, 32) RUN(dfma, 4, 32) RUN(dfma, 8, 32)
    RUN(dmma, 1, 256) RUN(dmma, 2, 256) RUN(dmma, 4, 256)
  }
  return 0;
}
