// Developer probe: do D2H / H2D copies overlap with a running kernel on this box?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long cycles, int* out) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} if (out) out[0] = 1; }
__global__ void touch(unsigned* p, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n) p[i] += 1; }
int main() {
  size_t B = 400u << 20; void *h, *d, *d2; cudaHostAlloc(&h, B, cudaHostAllocDefault); cudaMalloc(&d, B); cudaMalloc(&d2, B);
  cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
  cudaEvent_t e0, e1, e2, e3, base; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3); cudaEventCreate(&base);
  for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaDeviceSynchronize();
      cudaEventRecord(base, a);
      cudaEventRecord(e0, b);
      if (mode == 0 || mode == 2) cudaMemcpyAsync(h, d, B, cudaMemcpyDeviceToHost, b); else cudaMemcpyAsync(d, h, B, cudaMemcpyHostToDevice, b);
      cudaEventRecord(e1, b);
      cudaEventRecord(e2, a);
      if (mode < 2) spin<<<148, 256, 0, a>>>(20000000, nullptr);   // ~10 ms, all SMs busy with one CTA each
      else for (int i = 0; i < 200; i++) touch<<<(unsigned)((B / 4 + 255) / 256), 256, 0, a>>>((unsigned*)d2, B / 4);  // many short bandwidth-bound kernels
      cudaEventRecord(e3, a);
      cudaDeviceSynchronize();
      float t0, t1, t2, t3; cudaEventElapsedTime(&t0, base, e0); cudaEventElapsedTime(&t1, base, e1); cudaEventElapsedTime(&t2, base, e2); cudaEventElapsedTime(&t3, base, e3);
      if (rep) printf("mode %d (%s copy vs %s): copy %.2f-%.2f ms, kernels %.2f-%.2f ms\n", mode, (mode == 0 || mode == 2) ? "D2H" : "H2D", mode < 2 ? "spin kernel" : "200 streaming kernels", t0, t1, t2, t3);
    }
  }
  // mode 4: D2H on b, then H2D on c, then kernel on a (the order tgx_encode_batch submits them)
  cudaStream_t c; cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking);
  void *h2, *d3; cudaHostAlloc(&h2, B, cudaHostAllocDefault); cudaMalloc(&d3, B);
  cudaEvent_t f0, f1; cudaEventCreate(&f0); cudaEventCreate(&f1);
  for (int rep = 0; rep < 2; rep++) {
    cudaDeviceSynchronize();
    cudaEventRecord(base, a);
    cudaEventRecord(e0, b); cudaMemcpyAsync(h, d, B, cudaMemcpyDeviceToHost, b); cudaEventRecord(e1, b);
    cudaEventRecord(f0, c); cudaMemcpyAsync(d3, h2, B, cudaMemcpyHostToDevice, c); cudaEventRecord(f1, c);
    cudaEventRecord(e2, a); spin<<<148, 256, 0, a>>>(20000000, nullptr); cudaEventRecord(e3, a);
    cudaDeviceSynchronize();
    float t0, t1, t2, t3, g0, g1; cudaEventElapsedTime(&t0, base, e0); cudaEventElapsedTime(&t1, base, e1); cudaEventElapsedTime(&t2, base, e2); cudaEventElapsedTime(&t3, base, e3);
    cudaEventElapsedTime(&g0, base, f0); cudaEventElapsedTime(&g1, base, f1);
    if (rep) printf("mode 4: D2H %.2f-%.2f  H2D %.2f-%.2f  kernel %.2f-%.2f ms\n", t0, t1, g0, g1, t2, t3);
  }
  // mode 5: as mode 4, but stream a has itself issued a (tiny, long finished) D2H copy before:
  // its next kernel now starts only when the bulk D2H on stream b is over.
  for (int rep = 0; rep < 2; rep++) {
    cudaDeviceSynchronize();
    cudaMemcpyAsync(h2, d3, 8, cudaMemcpyDeviceToHost, a);
    cudaStreamSynchronize(a);
    cudaEventRecord(base, a);
    cudaEventRecord(e0, b); cudaMemcpyAsync(h, d, B, cudaMemcpyDeviceToHost, b); cudaEventRecord(e1, b);
    cudaEventRecord(f0, c); cudaMemcpyAsync(d3, h2, B, cudaMemcpyHostToDevice, c); cudaEventRecord(f1, c);
    cudaEventRecord(e2, a); spin<<<148, 256, 0, a>>>(20000000, nullptr); cudaEventRecord(e3, a);
    cudaDeviceSynchronize();
    float t0, t1, t2, t3, g0, g1; cudaEventElapsedTime(&t0, base, e0); cudaEventElapsedTime(&t1, base, e1); cudaEventElapsedTime(&t2, base, e2); cudaEventElapsedTime(&t3, base, e3);
    cudaEventElapsedTime(&g0, base, f0); cudaEventElapsedTime(&g1, base, f1);
    if (rep) printf("mode 5 (kernel stream issued an 8-byte D2H earlier): D2H %.2f-%.2f  H2D %.2f-%.2f  kernel %.2f-%.2f ms\n", t0, t1, g0, g1, t2, t3);
  }
  return 0;
}
