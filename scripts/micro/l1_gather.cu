// Micro-benchmark: cost of a divergent 64-byte gather per lane (one BVH node per lane) on sm_100a.
//  A: 4 x LDG.128 per lane (current traverse kernel)
//  B: 2 x LDG.256 per lane
//  C: cooperative: 4 lanes fetch each of their 4 nodes together (one 64-B segment per 4 lanes per
//     instruction), transposed through shared memory
//  D: 2 x LDG.128 per lane on 32-byte nodes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ unsigned next(unsigned x) { return x * 1664525u + 1013904223u; }

template <int MODE>
__global__ void __launch_bounds__(128, 8) gather(const float4* __restrict__ nodes, unsigned n_nodes, int iters, float* out)
{
  __shared__ float4 stage[4][32][4];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    x = next(x);
    const unsigned node = (x >> 8) % n_nodes;
    float4 a, b, c, d;
    if (MODE == 0) {
      const float4* p = nodes + (size_t)node * 4;
      a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); d = __ldg(p + 3);
    } else if (MODE == 1) {
      const float4* p = nodes + (size_t)node * 4;
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(c.x),"=f"(c.y),"=f"(c.z),"=f"(c.w),"=f"(d.x),"=f"(d.y),"=f"(d.z),"=f"(d.w) : "l"(p + 2));
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const unsigned owner = 8 * r + (lane >> 2);
        const unsigned nd = __shfl_sync(0xffffffffu, node, owner);
        stage[warp][owner][lane & 3] = __ldg(nodes + (size_t)nd * 4 + (lane & 3));
      }
      __syncwarp();
      a = stage[warp][lane][0]; b = stage[warp][lane][1]; c = stage[warp][lane][2]; d = stage[warp][lane][3];
      __syncwarp();
    } else {
      const float4* p = nodes + (size_t)node * 2;
      a = __ldg(p); b = __ldg(p + 1); c = a; d = b;
    }
    acc += a.x + b.y + c.z + d.w;
    x ^= __float_as_uint(acc) & 1u; // dependent chain like a traversal
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
  const unsigned n_nodes = 5000000; // 320 MB of 64-B nodes: misses L2 like the 10M-triangle scene
  float4* nodes; float* out;
  CHECK(cudaMalloc(&nodes, (size_t)n_nodes * 64));
  CHECK(cudaMemset(nodes, 0, (size_t)n_nodes * 64));
  const int blocks = 148 * 8, iters = 2000;
  CHECK(cudaMalloc(&out, blocks * 128 * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (unsigned nn : {5000u, 100000u, n_nodes}) {
    for (int mode = 0; mode < 4; ++mode) {
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) gather<0><<<blocks, 128>>>(nodes, nn, iters, out);
        if (mode == 1) gather<1><<<blocks, 128>>>(nodes, nn, iters, out);
        if (mode == 2) gather<2><<<blocks, 128>>>(nodes, nn, iters, out);
        if (mode == 3) gather<3><<<blocks, 128>>>(nodes, nn, iters, out);
        cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
      }
      const double visits = (double)blocks * 128 * iters;
      printf("nodes=%8u mode=%d  %.3f ms  %.2f Gvisits/s\n", nn, mode, best, visits / best * 1e-6);
    }
  }
  return 0;
}
