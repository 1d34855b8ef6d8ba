// Microbenchmark: throughput of dependent random 16-byte gathers (the trie-walk access pattern)
// from (a) a global table of a given size through L1/L2, (b) shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void gather_global(const uint4* __restrict__ tab, uint32_t mask, int iters, uint32_t* out, int ilp2) {
  uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
  uint32_t i1 = i0 * 40503u + 12345u;
  uint32_t acc = 0;
  if (ilp2) {
    for (int k = 0; k < iters; k++) {
      uint4 a = __ldg(tab + (i0 & mask));
      uint4 b = __ldg(tab + (i1 & mask));
      i0 = a.x + k; i1 = b.x + k; acc += a.y ^ b.y;
    }
  } else {
    for (int k = 0; k < iters; k++) {
      uint4 a = __ldg(tab + (i0 & mask));
      i0 = a.x + k; acc += a.y;
    }
  }
  if (acc == 0x12345678u) out[0] = acc + i0 + i1;
}

__global__ void gather_shared(const uint4* __restrict__ tab, uint32_t mask, int iters, uint32_t* out) {
  extern __shared__ uint4 s[];
  for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) s[i] = tab[i];
  __syncthreads();
  uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
  uint32_t acc = 0;
  for (int k = 0; k < iters; k++) {
    uint4 a = s[i0 & mask];
    i0 = a.x + k; acc += a.y;
  }
  if (acc == 0x12345678u) out[0] = acc + i0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SMs %d clock %d kHz\n", sms, clk);
  uint32_t* out; cudaMalloc(&out, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 2000;
  for (uint32_t logn : {8u, 11u, 13u, 15u, 17u, 18u, 20u, 22u}) {
    uint32_t n = 1u << logn;
    uint4* h = (uint4*)malloc((size_t)n * 16);
    uint32_t x = 12345;
    for (uint32_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; h[i].x = x >> 3; x = x * 1664525u + 1013904223u; h[i].y = x; h[i].z = h[i].w = 0; }
    uint4* d; cudaMalloc(&d, (size_t)n * 16); cudaMemcpy(d, h, (size_t)n * 16, cudaMemcpyHostToDevice);
    for (int ilp2 = 0; ilp2 < 2; ilp2++)
    for (int wpsm : {8, 16, 32, 64}) {
      int threads = 256, blocks = sms * wpsm * 32 / threads;
      gather_global<<<blocks, threads>>>(d, n - 1, 200, out, ilp2);
      cudaEventRecord(a);
      gather_global<<<blocks, threads>>>(d, n - 1, iters, out, ilp2);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      double lookups = (double)blocks * threads * iters * (ilp2 ? 2 : 1);
      double cyc = ms * 1e-3 * 1.965e9;
      printf("global tab=%8.1f KB ilp=%d warps/SM=%2d: %7.3f ms  %6.2f Glookup/s  %5.2f lookups/cyc/SM  lat~%6.0f cyc/iter\n", n * 16 / 1024.0, ilp2 + 1, wpsm, ms,
             lookups / ms / 1e6, lookups / cyc / sms, cyc / iters);
    }
    if (n * 16 <= 128 * 1024) {
      cudaFuncSetAttribute(gather_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 16);
      for (int tpb : {256, 512, 1024}) {
        int blocks = sms;
        gather_shared<<<blocks, tpb, n * 16>>>(d, n - 1, 200, out);
        cudaEventRecord(a);
        gather_shared<<<blocks, tpb, n * 16>>>(d, n - 1, iters, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double lookups = (double)blocks * tpb * iters;
        double cyc = ms * 1e-3 * 1.965e9;
        printf("shared tab=%8.1f KB warps/SM=%2d: %7.3f ms  %6.2f Glookup/s  %5.2f lookups/cyc/SM\n", n * 16 / 1024.0, tpb / 32, ms, lookups / ms / 1e6, lookups / cyc / sms);
      }
    }
    cudaFree(d); free(h);
  }
  return 0;
}
