// Micro-benchmark: FP32 throughput of scalar FFMA vs packed FFMA2 (sm_100), 8 independent
// accumulator chains per thread, full occupancy.  nvcc -arch=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__global__ void k_scalar(float* out, float a, float b, int iters)
{
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b, int iters)
{
  u64 x[8], pa, pb;
  asm("mov.b64 %0, {%1,%2};" : "=l"(pa) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
  for (int i = 0; i < 8; ++i) { float lo = threadIdx.x * 1e-3f + 2 * i, hi = lo + 1; asm("mov.b64 %0, {%1,%2};" : "=l"(x[i]) : "f"(lo), "f"(hi)); }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma2(x[i], pa, pb);
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 256, blocks = sms * 8, iters = 20000;
  float* out; cudaMalloc(&out, (size_t)threads * blocks * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0); k_scalar<<<blocks, threads>>>(out, 1.0001f, 1e-4f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 16 * iters * (double)threads * blocks;
    printf("scalar FFMA : %.3f ms  %.1f TFLOP/s  (%.1f FMA lanes/clk/SM at 1.965 GHz)\n", ms, fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / sms / 1.965e9);
    cudaEventRecord(e0); k_packed<<<blocks, threads>>>(out, 1.0001f, 1e-4f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FFMA2: %.3f ms  %.1f TFLOP/s  (%.1f FMA lanes/clk/SM)\n", ms, fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / sms / 1.965e9);
  }
  return 0;
}
