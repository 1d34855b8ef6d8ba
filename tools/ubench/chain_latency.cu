// Developer probe for the round-2 plan of the forward kernel ("what is left"): cycles per position of
// ONE ordered Viterbi chain for two consumer designs, on synthetic score tables in shared memory (no trie walk, no
// producers: the consumer side alone).
//
//   A  today's consumer (tgx_kernels.cuh::pair_consume): lane g owns the dp cell of positions == g (mod 16); a step
//      is  shuffle-broadcast of the final score -> DADD -> DSETP -> select  in the owner of the NEXT position, so the
//      shuffle sits on the chain.
//   B  short chain: lane 0 also keeps the chain itself.  The candidates of length 1..K — the only ones that depend on
//      the last K cells — are relaxed by lane 0 from registers (critical path: DADD -> DSETP -> select); the length
//      K+1..16 candidates are folded by the cell owners as today (off the chain: they have K-1 steps of slack) and
//      handed to lane 0 by a second shuffle.  "max, ties to the earliest start" is associative, so folding the
//      owners' partial result in late gives the reference's decision (src/model.rs:100-101) bit for bit.
//
//   C  B software-pipelined (see chain_c): B as written is SLOWER than A on B200 (measured, profiles/
//      r01_ubench_chain_latency.txt) because a warp issues in order: its shuffle -> relax -> shuffle sequence inside one
//      iteration stalls the next iteration's chain instructions behind it.
//
//   D  C with both hand-overs through shared memory and __syncwarp() (see chain_d; written after the SASS of C
//      showed ptxas sinking each shuffle to its consumer).  Not measured yet.
//
// All variants write the final score and the start of the best last token per position; the host compares them.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o chain_latency chain_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

// B and C: steps unrolled per loop trip.  Without it (first measurement, profiles/r01_ubench_chain_latency.txt) the
// score loads of a step cannot move above the back-edge and their shared-memory latency sits on the chain; the
// product's consumer prefetches its scores two steps ahead inside a 32-step unrolled tile for the same reason.
constexpr int UNROLL = 8;
constexpr int TAB = 1024;  // positions in the (repeating) score table
constexpr int ROW = 17;    // doubles per start position: column = target cell (start + len) % 16, padded

__device__ __forceinline__ double ninf() { return __longlong_as_double(0xFFF0000000000000ll); }

// ---------------------------------------------------------------------------------------------------------- A
__global__ void __launch_bounds__(32) chain_a(const double* __restrict__ g_tab, int n, double* out_best, int* out_start,
                                              long long* cycles) {
  extern __shared__ double tab[];
  for (int i = threadIdx.x; i < TAB * ROW; i += 32) tab[i] = g_tab[i];
  __syncwarp();
  const int g = threadIdx.x & 15;
  double best = (g == 0) ? 0.0 : ninf();  // dp[0] = 0.0, every other cell unreached
  int ps = 0;
  const long long t0 = clock64();
  for (int j = 0; j < n; j++) {
    const double bs = __shfl_sync(0xFFFFFFFFu, best, j & 15, 16);  // dp[j], final
    const bool own = g == (j & 15);
    if (own && threadIdx.x < 16) { out_best[j] = best; out_start[j] = ps; }
    const double cand = __dadd_rn(bs, tab[(j & (TAB - 1)) * ROW + g]);
    if (cand > best || own) { best = cand; ps = j; }  // a fresh cell (position j + 16) takes its first candidate
  }
  if (threadIdx.x == 0) cycles[0] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------------------- B
// "max, ties to the left operand" on (score, start) pairs
#define TEMAX(as, ap, bs, bp) do { if ((bs) > (as)) { (as) = (bs); (ap) = (bp); } } while (0)

template <int K>
__global__ void __launch_bounds__(32) chain_b(const double* __restrict__ g_tab, int n, double* out_best, int* out_start,
                                              long long* cycles) {
  extern __shared__ double tab[];
  for (int i = threadIdx.x; i < TAB * ROW; i += 32) tab[i] = g_tab[i];
  __syncwarp();
  const int g = threadIdx.x & 15;
  // owners: candidates of length K+1..16 only
  double cell = ninf();
  int cps = 0;
  // chain (meaningful in lane 0 of each half; every lane runs the same instructions):
  double bj = 0.0;  // dp[j]
  int bjp = 0;
  double acc[K + 1];  // acc[d], d = 1..K-1: chain candidates folded so far for target j + d (acc[0] unused)
  int accp[K + 1];
  constexpr int Q = K > 1 ? K - 1 : 1;
  double pq[Q];       // partial results of the owners for targets j+2 .. j+K (pq[0] = target j + 2)
  int pqp[Q];
#pragma unroll
  for (int d = 0; d <= K; d++) { acc[d] = ninf(); accp[d] = 0; }
#pragma unroll
  for (int d = 0; d < Q; d++) { pq[d] = ninf(); pqp[d] = 0; }
  const long long t0 = clock64();
#pragma unroll UNROLL
  for (int j = 0; j < n; j++) {
    const double* row = tab + (j & (TAB - 1)) * ROW;
    if (threadIdx.x == 0) { out_best[j] = bj; out_start[j] = bjp; }
    // ---- chain, critical part first: dp[j + 1] = temax(acc[1] (+ partial, folded earlier), dp[j] + s(j, 1))
    double c[K + 1];
#pragma unroll
    for (int d = 1; d <= K; d++) c[d] = __dadd_rn(bj, row[(j + d) & 15]);
    double nb = acc[1];
    int nbp = accp[1];
    TEMAX(nb, nbp, c[1], j);
    // ---- owners: broadcast dp[j], relax the cells whose candidate from start j is longer than K
    const double bs = __shfl_sync(0xFFFFFFFFu, bj, 0, 16);
    const int len = ((g - j) & 15) == 0 ? 16 : ((g - j) & 15);
    const double cand = __dadd_rn(bs, row[g]);
    if (len == 16 || (len > K && cand > cell)) { cell = cand; cps = j; }
    // the cell of target j + K + 1 has now seen every start up to j: hand it to the chain
    const double pv = __shfl_sync(0xFFFFFFFFu, cell, (j + K + 1) & 15, 16);
    const int pp = __shfl_sync(0xFFFFFFFFu, cps, (j + K + 1) & 15, 16);
    // ---- chain, off the critical path: fold the partial of target j + 2 (earlier starts: it wins ties), then c[2..K]
    if (K >= 2) {
      double x = pq[0];
      int xp = pqp[0];
      TEMAX(x, xp, acc[2], accp[2]);
      TEMAX(x, xp, c[2], j);
      acc[1] = x; accp[1] = xp;
#pragma unroll
      for (int d = 3; d <= K; d++) {
        double y = acc[d];
        int yp = accp[d];
        if (d == K) { y = c[K]; yp = j; } else TEMAX(y, yp, c[d], j);  // target j + K: its first chain candidate
        acc[d - 1] = y; accp[d - 1] = yp;
      }
      if (K == 2) { /* acc[2] is never read: target j + 2 only has the partial and c[2] */ }
    } else {
      acc[1] = pv; accp[1] = pp;  // K = 1: the partial of target j + 2 is complete after this step's relax
    }
#pragma unroll
    for (int d = 0; d + 1 < Q; d++) { pq[d] = pq[d + 1]; pqp[d] = pqp[d + 1]; }
    pq[Q - 1] = pv; pqp[Q - 1] = pp;  // target j + K + 1: read as pq[0] at step j + K - 1
    bj = nb; bjp = nbp;
  }
  if (threadIdx.x == 0) cycles[0] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------------------- C
// B, software-pipelined for an in-order warp: no instruction waits on a shuffle issued in the same iteration.
// The owners relax start j - 1 in iteration j (with the broadcast of dp[j - 1] issued an iteration earlier), the
// partial of target e is shuffled to the chain in iteration e - K + 1 — an iteration after it became complete, so
// the shuffle does not depend on that iteration's relax — and folded in iteration e - 2.  Needs K >= 4.
template <int K>
__global__ void __launch_bounds__(32) chain_c(const double* __restrict__ g_tab, int n, double* out_best, int* out_start,
                                              long long* cycles) {
  static_assert(K >= 4, "the partial needs one iteration in flight");
  extern __shared__ double tab[];
  for (int i = threadIdx.x; i < TAB * ROW; i += 32) tab[i] = g_tab[i];
  __syncwarp();
  const int g = threadIdx.x & 15;
  double cell = ninf();
  int cps = 0;
  double bj = 0.0, bq = ninf();
  int bjp = 0;
  double acc[K + 1];
  int accp[K + 1];
  constexpr int Q = K - 3;
  double pq[Q];
  int pqp[Q];
#pragma unroll
  for (int d = 0; d <= K; d++) { acc[d] = ninf(); accp[d] = 0; }
#pragma unroll
  for (int d = 0; d < Q; d++) { pq[d] = ninf(); pqp[d] = 0; }
  const long long t0 = clock64();
#pragma unroll UNROLL
  for (int j = 0; j < n; j++) {
    const double* row = tab + (j & (TAB - 1)) * ROW;
    const double* rowp = tab + ((j - 1) & (TAB - 1)) * ROW;
    if (threadIdx.x == 0) { out_best[j] = bj; out_start[j] = bjp; }
    // (1) partial of target j + K - 1: complete since the previous iteration, untouched by this one
    const double pvn = __shfl_sync(0xFFFFFFFFu, cell, (j + K - 1) & 15, 16);
    const int ppn = __shfl_sync(0xFFFFFFFFu, cps, (j + K - 1) & 15, 16);
    // (2) the chain
    double c[K + 1];
#pragma unroll
    for (int d = 1; d <= K; d++) c[d] = __dadd_rn(bj, row[(j + d) & 15]);
    double nb = acc[1];
    int nbp = accp[1];
    TEMAX(nb, nbp, c[1], j);
    // (3) broadcast of dp[j] for the owners' next iteration
    const double bqn = __shfl_sync(0xFFFFFFFFu, bj, 0, 16);
    // (4) owners: start j - 1
    {
      const int s = j - 1;
      const int len = ((g - s) & 15) == 0 ? 16 : ((g - s) & 15);
      const double cand = __dadd_rn(bq, rowp[g]);
      if (j >= 1 && (len == 16 || (len > K && cand > cell))) { cell = cand; cps = s; }
    }
    // (5) chain, off the critical path
    double x = pq[0];
    int xp = pqp[0];
    TEMAX(x, xp, acc[2], accp[2]);
    TEMAX(x, xp, c[2], j);
    acc[1] = x; accp[1] = xp;
#pragma unroll
    for (int d = 3; d <= K; d++) {
      double y = acc[d];
      int yp = accp[d];
      if (d == K) { y = c[K]; yp = j; } else TEMAX(y, yp, c[d], j);
      acc[d - 1] = y; accp[d - 1] = yp;
    }
#pragma unroll
    for (int d = 0; d + 1 < Q; d++) { pq[d] = pq[d + 1]; pqp[d] = pqp[d + 1]; }
    pq[Q - 1] = pvn; pqp[Q - 1] = ppn;
    bq = bqn; bj = nb; bjp = nbp;
  }
  if (threadIdx.x == 0) cycles[0] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------------------- D
// C with shared memory instead of shuffles for both hand-overs (dp[j] -> owners, partial -> chain): a store is
// fire-and-forget, the loads are issued at the top of an iteration and consumed late in it, and the __syncwarp()s
// pin the order that ptxas otherwise undoes by sinking a shuffle to its consumer.
template <int K>
__global__ void __launch_bounds__(32) chain_d(const double* __restrict__ g_tab, int n, double* out_best, int* out_start,
                                              long long* cycles) {
  static_assert(K >= 4, "the partial needs one iteration in flight");
  extern __shared__ double tab[];
  __shared__ double sm_b[2][4];      // [half][j & 3]: dp[j] for the owners
  __shared__ double sm_cell[2][16];  // [half][owner]: the owners' cells after each relax
  __shared__ int sm_cps[2][16];
  for (int i = threadIdx.x; i < TAB * ROW; i += 32) tab[i] = g_tab[i];
  const int g = threadIdx.x & 15, h = threadIdx.x >> 4;
  if (g < 4) sm_b[h][g] = ninf();
  sm_cell[h][g] = ninf();
  sm_cps[h][g] = 0;
  __syncwarp();
  double cell = ninf();
  int cps = 0;
  double bj = 0.0;
  int bjp = 0;
  double acc[K + 1];
  int accp[K + 1];
  constexpr int Q = K - 3;
  double pq[Q];
  int pqp[Q];
#pragma unroll
  for (int d = 0; d <= K; d++) { acc[d] = ninf(); accp[d] = 0; }
#pragma unroll
  for (int d = 0; d < Q; d++) { pq[d] = ninf(); pqp[d] = 0; }
  const long long t0 = clock64();
#pragma unroll UNROLL
  for (int j = 0; j < n; j++) {
    const double* row = tab + (j & (TAB - 1)) * ROW;
    const double* rowp = tab + ((j - 1) & (TAB - 1)) * ROW;
    if (threadIdx.x == 0) { out_best[j] = bj; out_start[j] = bjp; }
    // (1) loads of this iteration: partial of target j + K - 1 (complete since the previous iteration), dp[j - 1]
    const double pvn = sm_cell[h][(j + K - 1) & 15];
    const int ppn = sm_cps[h][(j + K - 1) & 15];
    const double bq = sm_b[h][(j - 1) & 3];
    // (3) dp[j] for the owners' next iteration
    if (g == 0) sm_b[h][j & 3] = bj;
    __syncwarp();
    // (2) the chain
    double c[K + 1];
#pragma unroll
    for (int d = 1; d <= K; d++) c[d] = __dadd_rn(bj, row[(j + d) & 15]);
    double nb = acc[1];
    int nbp = accp[1];
    TEMAX(nb, nbp, c[1], j);
    // (4) owners: start j - 1
    {
      const int s = j - 1;
      const int len = ((g - s) & 15) == 0 ? 16 : ((g - s) & 15);
      const double cand = __dadd_rn(bq, rowp[g]);
      if (j >= 1 && (len == 16 || (len > K && cand > cell))) { cell = cand; cps = s; }
      sm_cell[h][g] = cell;
      sm_cps[h][g] = cps;
    }
    // (5) chain, off the critical path
    double x = pq[0];
    int xp = pqp[0];
    TEMAX(x, xp, acc[2], accp[2]);
    TEMAX(x, xp, c[2], j);
    acc[1] = x; accp[1] = xp;
#pragma unroll
    for (int d = 3; d <= K; d++) {
      double y = acc[d];
      int yp = accp[d];
      if (d == K) { y = c[K]; yp = j; } else TEMAX(y, yp, c[d], j);
      acc[d - 1] = y; accp[d - 1] = yp;
    }
#pragma unroll
    for (int d = 0; d + 1 < Q; d++) { pq[d] = pq[d + 1]; pqp[d] = pqp[d + 1]; }
    pq[Q - 1] = pvn; pqp[Q - 1] = ppn;
    bj = nb; bjp = nbp;
    __syncwarp();
  }
  if (threadIdx.x == 0) cycles[0] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------------------- host
static void reference(const std::vector<double>& tab, int n, std::vector<double>& best, std::vector<int>& start) {
  const double NINF = -1.0 / 0.0;
  std::vector<double> dp(n + 17, NINF);
  std::vector<int> st(n + 17, 0);
  dp[0] = 0.0;
  for (int j = 0; j < n; j++) {
    best[j] = dp[j];
    start[j] = st[j];
    for (int len = 16; len >= 1; len--) {  // (order within one start does not matter: different targets)
      const double cand = dp[j] + tab[(size_t)(j & (TAB - 1)) * ROW + ((j + len) & 15)];
      if (len == 16 || cand > dp[j + len]) { dp[j + len] = cand; st[j + len] = j; }  // (len 16: the target's first)
    }
  }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 262144;
  std::vector<double> tab((size_t)TAB * ROW);
  unsigned long long s = 88172645463325252ull;
  for (auto& v : tab) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const unsigned r = (unsigned)(s >> 40);
    v = (r % 3 == 0) ? -1.0 / 0.0 : -(double)(1 + r % 7);  // a third of the cells empty; small integers = exact ties
  }
  for (int j = 0; j < TAB; j++) tab[(size_t)j * ROW + ((j + 1) & 15)] = -(double)(3 + j % 5);  // every byte is a token
  double *d_tab, *d_best;
  int* d_start;
  long long* d_cyc;
  cudaMalloc(&d_tab, tab.size() * 8);
  cudaMalloc(&d_best, (size_t)n * 8);
  cudaMalloc(&d_start, (size_t)n * 4);
  cudaMalloc(&d_cyc, 8);
  cudaMemcpy(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)TAB * ROW * 8;
  cudaFuncSetAttribute(chain_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(chain_b<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(chain_b<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(chain_b<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(chain_c<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(chain_c<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(chain_d<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<double> want(n), got(n);
  std::vector<int> wants(n), gots(n);
  reference(tab, n, want, wants);
  for (int v = 0; v < 7; v++) {
    long long cyc = 0;
    for (int rep = 0; rep < 2; rep++) {
      cudaMemset(d_best, 0, (size_t)n * 8);
      if (v == 0) chain_a<<<1, 32, smem>>>(d_tab, n, d_best, d_start, d_cyc);
      if (v == 1) chain_b<2><<<1, 32, smem>>>(d_tab, n, d_best, d_start, d_cyc);
      if (v == 2) chain_b<3><<<1, 32, smem>>>(d_tab, n, d_best, d_start, d_cyc);
      if (v == 3) chain_b<4><<<1, 32, smem>>>(d_tab, n, d_best, d_start, d_cyc);
      if (v == 4) chain_c<4><<<1, 32, smem>>>(d_tab, n, d_best, d_start, d_cyc);
      if (v == 5) chain_c<5><<<1, 32, smem>>>(d_tab, n, d_best, d_start, d_cyc);
      if (v == 6) chain_d<4><<<1, 32, smem>>>(d_tab, n, d_best, d_start, d_cyc);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    }
    cudaMemcpy(got.data(), d_best, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(gots.data(), d_start, (size_t)n * 4, cudaMemcpyDeviceToHost);
    long long bad = 0, first = -1;
    for (int j = 0; j < n; j++)
      if (memcmp(&got[j], &want[j], 8) != 0 || (want[j] != -1.0 / 0.0 && j > 0 && gots[j] != wants[j])) { if (first < 0) first = j; bad++; }
    printf("%s: %.1f cycles per position (%d positions), %lld mismatches vs the host chain (first at %lld)\n",
           v == 0 ? "A shuffle on the chain (today)" : v == 1 ? "B K=2" : v == 2 ? "B K=3" : v == 3 ? "B K=4" : v == 4 ? "C K=4 (pipelined)" : v == 5 ? "C K=5 (pipelined)" : "D K=4 (pipelined, hand-overs through shared memory)", (double)cyc / n, n, bad,
           first);
  }
  return 0;
}
