// Developer probe: does pinned host memory stay registered / device-accessible on this box?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void wr(unsigned long long* p) { p[0] = 42; }
static void show(const char* tag, void* p) {
  cudaPointerAttributes a; cudaError_t e = cudaPointerGetAttributes(&a, p);
  printf("%s: err=%d type=%d dev=%p host=%p\n", tag, (int)e, (int)a.type, a.devicePointer, a.hostPointer);
}
int main() {
  void* h = nullptr; cudaError_t e = cudaHostAlloc(&h, 256, cudaHostAllocDefault); printf("alloc %d %p\n", (int)e, h);
  show("after alloc", h);
  void* d; cudaMalloc(&d, 100 << 20); show("after cudaMalloc", h);
  void* h2 = nullptr; cudaHostAlloc(&h2, 8 << 20, cudaHostAllocDefault); show("after 2nd hostalloc (small)", h); show("big", h2);
  wr<<<1, 1>>>((unsigned long long*)h); e = cudaDeviceSynchronize(); printf("kernel write: %s val=%llu\n", cudaGetErrorString(e), *(unsigned long long*)h);
  void* h3 = nullptr; e = cudaHostAlloc(&h3, 1ull << 30, cudaHostAllocDefault); printf("1GB pinned alloc: %s\n", cudaGetErrorString(e)); show("1GB", h3); show("small after 1GB", h);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  void* dd; cudaMalloc(&dd, 1ull << 30);
  for (int r = 0; r < 2; r++) { cudaEventRecord(a); cudaMemcpyAsync(dd, h3, 1ull << 30, cudaMemcpyHostToDevice); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); printf("H2D 1GB pinned: %.2f ms (%.1f GB/s)\n", ms, 1.0737 / ms * 1e3); }
  void* pg = malloc(1ull << 30); for (size_t i = 0; i < (1ull << 30); i += 4096) ((char*)pg)[i] = 1;
  { cudaEventRecord(a); cudaMemcpyAsync(dd, pg, 1ull << 30, cudaMemcpyHostToDevice); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); printf("H2D 1GB pageable: %.2f ms (%.1f GB/s)\n", ms, 1.0737 / ms * 1e3); }
  return 0;
}
