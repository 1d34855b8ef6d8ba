// Developer test: backtrack_warp_kernel vs backtrack_thread_kernel on random back-length arrays.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -I../../tokengeex_b200/csrc -o bt_test bt_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include "tgx_kernels.cuh"
using namespace tgxk;
int main() {
  std::mt19937_64 rng(1);
  const int U = 600;
  std::vector<uint64_t> ustart(U); std::vector<uint32_t> ulen(U), order(U);
  uint64_t N = 0;
  for (int i = 0; i < U; i++) { ustart[i] = N; ulen[i] = 512 + rng() % 9000; if (i % 7 == 0) ulen[i] = 1024 * (1 + rng() % 4) + (rng() % 3) - 1; N += ulen[i]; order[i] = i; }
  std::vector<uint8_t> bp(N + 64, 0);
  for (int i = 0; i < U; i++) {
    uint32_t n = ulen[i]; uint64_t s = ustart[i];
    for (uint32_t p = 1; p <= n; p++) { uint32_t l = rng() % 17; if (l > p) l = 0; bp[s + p - 1] = l; }
    if (i % 11 != 3) { uint32_t pos = n; while (pos) { uint32_t l = 1 + rng() % 16; if (l > pos) l = pos; bp[s + pos - 1] = l; pos -= l; } }
    else bp[s + n - 1] = 0;
  }
  uint8_t *d_bp, *d_m1, *d_m2; uint64_t* d_us; uint32_t *d_ul, *d_or; unsigned long long *d_n1, *d_n2; int32_t *d_s1, *d_s2;
  uint64_t NP = ((N + 4095) / 4096) * 4096 + 64;
  cudaMalloc(&d_bp, N + 64); cudaMalloc(&d_m1, NP); cudaMalloc(&d_m2, NP); cudaMalloc(&d_us, U * 8); cudaMalloc(&d_ul, U * 4); cudaMalloc(&d_or, U * 4);
  cudaMalloc(&d_n1, U * 8); cudaMalloc(&d_n2, U * 8); cudaMalloc(&d_s1, U * 4); cudaMalloc(&d_s2, U * 4);
  cudaMemcpy(d_bp, bp.data(), N + 64, cudaMemcpyHostToDevice); cudaMemset(d_m1, 0, NP); cudaMemset(d_m2, 0, NP);
  cudaMemcpy(d_us, ustart.data(), U * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_ul, ulen.data(), U * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_or, order.data(), U * 4, cudaMemcpyHostToDevice);
  BacktrackParams b; b.unit_start = d_us; b.unit_len = d_ul; b.order = d_or; b.first = 0; b.count = U; b.bp = d_bp;
  b.mark = d_m1; b.n_tokens = d_n1; b.status = d_s1;
  backtrack_thread_kernel<<<(U + 127) / 128, 128>>>(b);
  b.mark = d_m2; b.n_tokens = d_n2; b.status = d_s2;
  backtrack_warp_kernel<<<(U + BW_WARPS - 1) / BW_WARPS, BW_WARPS * 32>>>(b);
  cudaError_t e = cudaDeviceSynchronize(); printf("sync: %s\n", cudaGetErrorString(e));
  std::vector<uint8_t> m1(NP), m2(NP); std::vector<unsigned long long> n1(U), n2(U); std::vector<int32_t> s1(U), s2(U);
  cudaMemcpy(m1.data(), d_m1, NP, cudaMemcpyDeviceToHost); cudaMemcpy(m2.data(), d_m2, NP, cudaMemcpyDeviceToHost);
  cudaMemcpy(n1.data(), d_n1, U * 8, cudaMemcpyDeviceToHost); cudaMemcpy(n2.data(), d_n2, U * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(s1.data(), d_s1, U * 4, cudaMemcpyDeviceToHost); cudaMemcpy(s2.data(), d_s2, U * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < U && bad < 10; i++) {
    bool mm = false; uint64_t first = 0;
    for (uint64_t g = ustart[i]; g < ustart[i] + ulen[i]; g++) if (m1[g] != m2[g]) { mm = true; first = g - ustart[i]; break; }
    if (mm || n1[i] != n2[i] || s1[i] != s2[i]) { printf("unit %d start %llu n %u: status %d/%d ntok %llu/%llu first mark diff at %llu\n", i, (unsigned long long)ustart[i], ulen[i], s1[i], s2[i], n1[i], n2[i], (unsigned long long)first); bad++; }
  }
  printf("bad %d\n", bad);
  return 0;
}
