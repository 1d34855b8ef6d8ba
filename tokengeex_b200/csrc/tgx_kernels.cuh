// sm_100a kernels of the TokenGeeX hot path (see DESIGN.md for the data layout).
//
// Work unit = one sample (encode / frequency pass) or one <= 81920-byte snippet (E-step).
// Units are sorted by length (descending) and handed to groups of G lanes (G = 1..32):
// a warp runs 32/G units side by side.  Each group alternates two phases over tiles of
// G positions:
//   phase A (parallel)  lane i walks the double-array trie from position p0+i and parks
//                       every match (score, len, id) in the warp's shared-memory buffer;
//   phase B (ordered)   positions p0 .. p0+G-1 are finalised one after the other; the
//                       matches of the current position are relaxed by the lanes of the
//                       group in parallel (distinct lengths -> distinct targets) against a
//                       rolling window of the next max_token_len positions in shared memory.
// Phase B replays the reference's evaluation order exactly (ascending start position,
// strict '>' on f64 sums built left to right), which is what makes ids bit-exact.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "tgx_libm.h"
#include "tgx_libm_tables.h"

namespace tgxk {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t F_OCC = 1u << 24, F_TERM = 2u << 24, F_HASCH = 4u << 24, ID_MASK = 0x00FFFFFFu;
constexpr int ROW_STRIDE = 33;  // padded row of the match buffer: conflict-free column reads
constexpr int WPB = 4;          // warps per block
constexpr uint32_t BT_STAGE = 1024;  // tokens parked per backtrack flush (CTA kernel)

struct UnitParams {
  const uint8_t* text;         // blob
  const uint64_t* unit_start;  // [U] absolute byte offset of the unit in text
  const uint32_t* unit_len;    // [U]
  const uint32_t* order;       // sorted unit indices; this launch handles order[first .. first+count)
  uint32_t first, count;
  const uint4* trie;
  uint32_t root_base;
  uint32_t rows;  // match-buffer rows = max token length
  uint32_t W;     // window slots = rows + 1
};

struct ViterbiParams {
  UnitParams u;
  uint32_t* bp;                  // [N] back-pointers (len << 24 | id) per end position; later ids, right-aligned
  unsigned long long* n_tokens;  // [U]
  int32_t* status;               // [U] 0 ok / 6 NoPath
  unsigned long long* freq;      // optional [V]: frequency pass
  int emit;                      // write ids in place (encode) or not (frequency pass only)
};

struct FbParams {
  UnitParams u;
  double* A;          // [N + U] forward log-probabilities: unit k owns A[start_k + k .. + n_k]
  int32_t* status;    // [U] 0 ok / 7 bad z
  double* expected;   // [V]
  // Hot tokens (small ids: vocabularies are score-sorted) would serialise every SM's atomics
  // on a handful of L2 addresses, so ids < hot_k accumulate into one of hot_r replicas
  // (picked per block) that fold_hot_kernel sums into expected[] afterwards.
  double* hot;        // [hot_r][hot_k]
  uint32_t hot_k, hot_r;
};

__host__ __device__ inline size_t warp_smem_bytes(uint32_t rows, uint32_t W, int G) {
  size_t ng = 32 / G;
  size_t b = (size_t)rows * ROW_STRIDE * 8;  // mscore
  b += (size_t)ng * W * 8;                   // window f64
  b += (size_t)rows * ROW_STRIDE * 4;        // mpack
  b += (size_t)ng * W * 4;                   // window u32
  b += 32 * 4;                               // mcnt
  return (b + 15) & ~(size_t)15;
}

struct WarpSmem {
  double* mscore;
  double* wf;       // window f64 [ng][W]
  uint32_t* mpack;
  uint32_t* wu;     // window u32 [ng][W]
  uint32_t* mcnt;
};

__device__ inline WarpSmem carve(unsigned char* base, uint32_t rows, uint32_t W, int G) {
  WarpSmem s;
  size_t ng = 32 / G;
  s.mscore = reinterpret_cast<double*>(base);
  s.wf = s.mscore + (size_t)rows * ROW_STRIDE;
  s.mpack = reinterpret_cast<uint32_t*>(s.wf + ng * W);
  s.wu = s.mpack + (size_t)rows * ROW_STRIDE;
  s.mcnt = s.wu + ng * W;
  return s;
}

// Phase A: common_prefix_search from `pos` (src/trie.rs:51-63 restated on the double-array):
// one 16-byte load per byte walked; stops at the first missing edge.
__device__ __forceinline__ uint32_t walk_matches(const UnitParams& u, const uint8_t* text, uint32_t pos,
                                                 uint32_t n, const WarpSmem& s, int lane) {
  uint32_t cnt = 0;
  if (pos < n) {
    uint32_t base = u.root_base;
    uint32_t d = 0;
    uint32_t maxd = min(n - pos, u.rows);
    while (d < maxd) {
      uint32_t c = __ldg(text + pos + d);
      uint4 e = __ldg(u.trie + (base ^ c));
      if ((e.x & 0xFFu) != c || !(e.y & F_OCC)) break;
      d++;
      if (e.y & F_TERM) {
        s.mscore[cnt * ROW_STRIDE + lane] = __hiloint2double((int)e.w, (int)e.z);
        s.mpack[cnt * ROW_STRIDE + lane] = (d << 24) | (e.y & ID_MASK);
        cnt++;
      }
      if (!(e.y & F_HASCH)) break;
      base = e.x >> 8;
    }
  }
  return cnt;
}

// -----------------------------------------------------------------------------------------
// K2/K3  Viterbi forward + in-kernel backtrack.  Model::encode, src/model.rs:59-129.
// -----------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(WPB * 32) viterbi_kernel(ViterbiParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NG = 32 / G;
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lig = lane & (G - 1), gid = lane / G;
  const uint32_t W = u.W;
  WarpSmem s = carve(smem + (size_t)warp * warp_smem_bytes(u.rows, W, G), u.rows, W, G);
  double* wbest = s.wf + (size_t)gid * W;
  uint32_t* wbp = s.wu + (size_t)gid * W;

  const uint64_t gidx = ((uint64_t)blockIdx.x * WPB + warp) * NG + gid;
  const bool has = gidx < u.count;
  const uint32_t unit = has ? u.order[u.first + gidx] : 0;
  const uint32_t n = has ? u.unit_len[unit] : 0;
  const uint64_t start = has ? u.unit_start[unit] : 0;
  const uint8_t* text = u.text + start;

  for (uint32_t i = lig; i < W; i += G) wbp[i] = NONE;
  uint32_t nmax = n;
#pragma unroll
  for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
  __syncwarp();

  // positions 0..n are visited (position n only emits its back-pointer)
  const uint32_t tiles = nmax / G + 1;
  uint32_t slot0 = 0;  // (tile * G) % W
  for (uint32_t tile = 0; tile < tiles; tile++) {
    const uint32_t p0 = tile * G;
    // ---- phase A
    s.mcnt[lane] = walk_matches(u, text, p0 + lig, n, s, lane);
    __syncwarp();
    // ---- phase B
    uint32_t my_bp = NONE;  // back-pointer of end position p0 + lig
    uint32_t sl = slot0;
#pragma unroll 1
    for (int j = 0; j < G; j++) {
      const uint32_t pp = p0 + j;
      const int src = gid * G + j;
      double best;
      uint32_t pk;
      bool reached;
      if (pp == 0) {  // dp[0].start = Some(0), score 0.0   (src/model.rs:72-81)
        best = 0.0; pk = 0; reached = true;
      } else {
        best = wbest[sl]; pk = wbp[sl]; reached = pk != NONE;
      }
      if (lig == j) my_bp = pk;
      __syncwarp();
      if (lig == 0) wbp[sl] = NONE;  // the slot now stands for position pp + W
      const uint32_t c = (pp < n) ? s.mcnt[src] : 0;
      if (reached) {  // unreachable positions are skipped (src/model.rs:85-87)
        for (uint32_t k = lig; k < c; k += G) {
          const double sc = s.mscore[k * ROW_STRIDE + src];
          const uint32_t mp = s.mpack[k * ROW_STRIDE + src];
          const double cand = __dadd_rn(best, sc);  // dp[pos].score + vocab[id].score  (:98)
          uint32_t ts = sl + (mp >> 24);
          if (ts >= W) ts -= W;
          const uint32_t opk = wbp[ts];
          if (opk == NONE || cand > wbest[ts]) {  // node.start.is_none() || score > node.score  (:100-101)
            wbest[ts] = cand;
            wbp[ts] = mp;
          }
        }
      }
      __syncwarp();
      if (++sl == W) sl = 0;
    }
    const uint32_t e = p0 + lig;
    if (has && e >= 1 && e <= n) p.bp[start + e - 1] = my_bp;
    slot0 += G;
    while (slot0 >= W) slot0 -= W;
  }
  __syncwarp();

  // ---- backtrack (src/model.rs:113-126): ids are written right-aligned into the unit's own
  // back-pointer region: token k from the end lands at index n-1-k >= the index just read.
  if (has && lig == 0) {
    unsigned long long k = 0;
    int st = 0;
    if (n > 0) {
      if (p.bp[start + n - 1] == NONE) {
        st = 6;  // Error::NoPath(n, n)
      } else {
        uint32_t pos = n;
        while (pos > 0) {
          const uint32_t v = p.bp[start + pos - 1];
          const uint32_t id = v & ID_MASK;
          if ((v >> 24) == 0 || (v >> 24) > pos) { st = 99; break; }  // corrupt chain: never loop forever
          if (p.freq) atomicAdd(p.freq + id, 1ull);  // src/prune.rs:223-225
          if (p.emit) p.bp[start + n - 1 - k] = id;
          pos -= v >> 24;
          k++;
        }
      }
    }
    p.n_tokens[unit] = st ? 0 : k;
    p.status[unit] = st;
  }
}

// -----------------------------------------------------------------------------------------
// K2c  CTA-cooperative Viterbi: one sample per CTA, producer/consumer.
//
// The per-sample critical path is the ordered relax chain (position p must be final before
// its matches are pushed), so the kernel is built around its latency, not its width:
//   * P producer warps run phase A (trie walks) one 32-position tile each, writing a DENSE
//     table  mpack[depth][column] / mscore[depth][column]  (0 = no token of that length);
//   * ONE consumer warp runs phase B entirely in registers: lane l owns the dp cell of
//     every position q with q % 32 == l (max_token_len <= 31), so finalising position p is
//     a shuffle-broadcast of lane p%32's (score, back-pointer) and the relax of the match
//     of length len is a predicated compare in lane (p+len)%32 — no shared-memory
//     round trip and no barrier inside the chain.
// Rounds of P tiles are double-buffered and separated by one __syncthreads().
// CTAs fetch samples from a global counter in length-descending order (LPT).
// -----------------------------------------------------------------------------------------
struct CtaStage {
  uint32_t* mpack;   // [rows][ROW_STRIDE]
  double* mscore;    // [rows][ROW_STRIDE]
};

__host__ __device__ inline size_t cta_stage_bytes(uint32_t rows) {
  return ((size_t)rows * ROW_STRIDE * 12 + 15) & ~(size_t)15;
}

// 24-byte text window starting at the 8-byte aligned address at or below `ptr`
// (w[0] bits 8*sh.. hold *ptr).  Never reads at or beyond blob_end.
__device__ __forceinline__ void load_window(const uint8_t* ptr, const uint8_t* blob_end, unsigned long long (&w)[3],
                                            uint32_t& sh) {
  const unsigned long long a = reinterpret_cast<unsigned long long>(ptr) & ~7ull;
  sh = (uint32_t)(reinterpret_cast<unsigned long long>(ptr) & 7ull);
  const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a);
  if (reinterpret_cast<const uint8_t*>(q + 3) <= blob_end) {
    w[0] = __ldg(q); w[1] = __ldg(q + 1); w[2] = __ldg(q + 2);
  } else {  // last bytes of the blob
    const uint8_t* b = reinterpret_cast<const uint8_t*>(q);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      unsigned long long v = 0;
      for (int k = 0; k < 8; k++) {
        const uint8_t* bp_ = b + i * 8 + k;
        const unsigned long long byte = (bp_ >= ptr && bp_ < blob_end) ? (unsigned long long)__ldg(bp_) : 0ull;
        v |= byte << (8 * k);
      }
      w[i] = v;
    }
  }
}

// Eight text bytes starting at ptr + 8*g, assembled from the 24-byte window.
__device__ __forceinline__ unsigned long long window_bytes(const unsigned long long (&w)[3], uint32_t sh, int g) {
  // g in {0,1}: bytes [8g, 8g+8) relative to ptr  =  (w[g] >> 8sh) | (w[g+1] << (64 - 8sh))
  const unsigned long long lo = w[g], hi = w[g + 1];
  return sh ? ((lo >> (8 * sh)) | (hi << (64 - 8 * sh))) : lo;
}

// Phase A for one 32-position tile.  Dense table per stage: row d (token length d+1), column =
// start position within the tile; empty entries hold score -inf (their mpack is never read).
__device__ __forceinline__ void produce_tile(const UnitParams& u, const uint8_t* text, const uint8_t* blob_end,
                                             uint32_t pos, uint32_t n, uint32_t* mpack, double* mscore, int lane) {
  const uint32_t rows = u.rows;
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
  {
    double* r = mscore + lane;
    for (uint32_t d = 0; d < rows; d++, r += ROW_STRIDE) *r = ninf;
  }
  if (pos >= n) return;
  const uint32_t maxd = min(n - pos, rows);
  uint32_t base = u.root_base;
  double* ms = mscore + lane;
  uint32_t* mp = mpack + lane;
  uint32_t dpk = 1u << 24;  // (depth + 1) << 24
  for (uint32_t d0 = 0; d0 < maxd; d0 += 16) {
    unsigned long long w[3];
    uint32_t sh;
    load_window(text + pos + d0, blob_end, w, sh);
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const unsigned long long a = window_bytes(w, sh, g);
      const uint32_t alo = (uint32_t)a, ahi = (uint32_t)(a >> 32);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (d0 + g * 8 + k >= maxd) return;
        const uint32_t c = __byte_perm(k < 4 ? alo : ahi, 0, 0x4440 + (k & 3));
        const uint4 e = __ldg(u.trie + (base ^ c));
        if ((e.x & 0xFFu) != c || !(e.y & F_OCC)) return;
        if (e.y & F_TERM) {
          *ms = __hiloint2double((int)e.w, (int)e.z);
          *mp = dpk | (e.y & ID_MASK);
        }
        if (!(e.y & F_HASCH)) return;
        base = e.x >> 8;
        ms += ROW_STRIDE;
        mp += ROW_STRIDE;
        dpk += 1u << 24;
      }
    }
  }
}

// Phase B over one 32-position tile, in registers.  Lane l owns the dp cell (best, pk) of
// every position q with q % 32 == l; "unreached" is best == -inf (scores are finite, so a
// candidate built on an unreached position is -inf and can never win, and the first finite
// candidate always replaces -inf — the reference's `start.is_none() ||` test, src/model.rs:100).
// Operands (score, packed len|id) are fetched three positions ahead of the ordered chain,
// which is then only  shuffle -> DADD -> compare -> select.
__device__ __forceinline__ void load_operand(const uint32_t* __restrict__ mpack, const double* __restrict__ mscore,
                                             uint32_t rows, int lane, int j, uint32_t& mp, double& sc) {
  const uint32_t len = (uint32_t)(lane - j) & 31u;
  const bool valid = (len - 1u) < rows;
  const uint32_t row = valid ? len - 1u : 0u;
  mp = mpack[row * ROW_STRIDE + j];
  const double v = mscore[row * ROW_STRIDE + j];
  sc = valid ? v : __longlong_as_double(0xFFF0000000000000ll);
}

__device__ __forceinline__ uint32_t consume_tile(const uint32_t* __restrict__ mpack, const double* __restrict__ mscore,
                                                 uint32_t rows, int lane, double& best, uint32_t& pk) {
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
  uint32_t my_bp = NONE;
  uint32_t mp[4];
  double sc[4];
#pragma unroll
  for (int j = 0; j < 3; j++) load_operand(mpack, mscore, rows, lane, j, mp[j], sc[j]);
#pragma unroll
  for (int j = 0; j < 32; j++) {
    if (j + 3 < 32) load_operand(mpack, mscore, rows, lane, j + 3, mp[(j + 3) & 3], sc[(j + 3) & 3]);
    const double bsrc = __shfl_sync(0xFFFFFFFFu, best, j);
    if (lane == j) { my_bp = pk; pk = NONE; best = ninf; }  // the cell now stands for position p + 32
    const double cand = __dadd_rn(bsrc, sc[j & 3]);           // dp[pos].score + vocab[id].score  (src/model.rs:98)
    if (cand > best) {                                        // (:100-101)
      best = cand;
      pk = mp[j & 3];
    }
  }
  return my_bp;
}

template <int P>
__global__ void __launch_bounds__(32 * (P + 1), (P <= 2 ? 8 : 4)) viterbi_cta_kernel(ViterbiParams p, unsigned int* work_counter,
                                                                   const uint8_t* blob_end, uint32_t chunk_cap) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t s_unit, s_pos, s_endpk;
  __shared__ unsigned long long s_k;
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rows = u.rows;
  const size_t stage_b = cta_stage_bytes(rows);
  uint32_t* chunk = reinterpret_cast<uint32_t*>(smem);

  for (;;) {
    if (threadIdx.x == 0) s_unit = atomicAdd(work_counter, 1u);
    __syncthreads();
    const uint32_t r = s_unit;
    if (r >= u.count) break;
    const uint32_t unit = u.order[u.first + r];
    const uint32_t n = u.unit_len[unit];
    const uint64_t start = u.unit_start[unit];
    const uint8_t* text = u.text + start;
    const uint32_t ntiles = n / 32 + 1;  // positions 0..n
    const uint32_t rounds = (ntiles + P - 1) / P;
    // dp[0] = { score 0.0, start Some(0) }  (src/model.rs:72-81); every other cell unreached
    double best = (lane == 0) ? 0.0 : __longlong_as_double(0xFFF0000000000000ll);
    uint32_t pk = (lane == 0) ? 0u : NONE;
    uint32_t last_bp = NONE;
    for (uint32_t round = 0; round <= rounds; round++) {
      if (warp > 0) {
        const uint32_t t = round * P + (warp - 1);
        if (round < rounds && t < ntiles) {
          unsigned char* st = smem + ((size_t)(round & 1) * P + (warp - 1)) * stage_b;
          double* ms = reinterpret_cast<double*>(st);
          uint32_t* mp = reinterpret_cast<uint32_t*>(st + (size_t)rows * ROW_STRIDE * 8);
          produce_tile(u, text, blob_end, t * 32 + lane, n, mp, ms, lane);
        }
      } else if (round > 0) {
        for (int k = 0; k < P; k++) {
          const uint32_t t = (round - 1) * P + k;
          if (t >= ntiles) break;
          const unsigned char* st = smem + ((size_t)((round - 1) & 1) * P + k) * stage_b;
          const double* ms = reinterpret_cast<const double*>(st);
          const uint32_t* mp = reinterpret_cast<const uint32_t*>(st + (size_t)rows * ROW_STRIDE * 8);
          const uint32_t my_bp = consume_tile(mp, ms, rows, lane, best, pk);
          const uint32_t e = t * 32 + lane;
          if (e >= 1 && e <= n) p.bp[start + e - 1] = my_bp;
          if (e == n) last_bp = my_bp;
        }
      }
      __syncthreads();
    }
    if (warp == 0 && lane == (int)(n & 31u)) s_endpk = (n == 0) ? 0u : last_bp;
    if (threadIdx.x == 0) { s_pos = n; s_k = 0; }
    __syncthreads();
    // ---- backtrack (src/model.rs:113-126).  Back-pointers are staged in shared memory a chunk
    // at a time; one thread follows the chain (a pure LDS -> subtract dependency) and parks the
    // visited entries in a second buffer, which all threads then flush: ids go right-aligned
    // into the sample's own back-pointer region (token k from the end at index n-1-k, always
    // >= any index still to be read), frequencies through atomics.
    int st_code = 0;
    if (n > 0 && s_endpk == NONE) {
      st_code = 6;  // Error::NoPath(n, n)
    } else {
      uint32_t* stage = chunk + chunk_cap;  // [BT_STAGE]
      uint32_t pos = n;
      unsigned long long kbase = 0;
      while (pos > 0) {
        const uint32_t lo = pos > chunk_cap ? pos - chunk_cap : 0;
        for (uint32_t i = threadIdx.x; i < pos - lo; i += blockDim.x) chunk[i] = p.bp[start + lo + i];
        __syncthreads();
        uint32_t q = pos;
        while (q > lo) {  // uniform: q is re-read from shared memory after every stage
          if (threadIdx.x == 0) {
            uint32_t qq = q, cnt = 0;
            while (qq > lo && cnt < BT_STAGE) {
              const uint32_t v = chunk[qq - 1 - lo];
              const uint32_t len = v >> 24;
              if (len == 0 || len > qq) { qq = 0xFFFFFFFFu; break; }  // corrupt chain: abort, never spin
              stage[cnt++] = v;
              qq -= len;
            }
            s_pos = qq;
            s_k = cnt;
          }
          __syncthreads();
          const uint32_t cnt = (uint32_t)s_k;
          for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
            const uint32_t id = stage[i] & ID_MASK;
            if (p.freq) atomicAdd(p.freq + id, 1ull);  // src/prune.rs:223-225
            if (p.emit) p.bp[start + n - 1 - (kbase + i)] = id;
          }
          kbase += cnt;
          q = s_pos;
          __syncthreads();
          if (q == 0xFFFFFFFFu) break;
        }
        if (q == 0xFFFFFFFFu) { st_code = 99; break; }
        pos = q;
      }
      if (threadIdx.x == 0) s_k = kbase;
    }
    if (threadIdx.x == 0) {
      p.n_tokens[unit] = st_code ? 0ull : s_k;
      p.status[unit] = st_code;
    }
    __syncthreads();
  }
}

// K3b: compact the right-aligned ids into the caller's id array (input order).
__global__ void gather_ids_kernel(const uint32_t* __restrict__ bp, const uint64_t* __restrict__ unit_start,
                                  const uint32_t* __restrict__ unit_len,
                                  const unsigned long long* __restrict__ id_off, uint32_t U,
                                  uint32_t* __restrict__ out, unsigned long long cap) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= U) return;
  const unsigned long long o = id_off[warp], T = id_off[warp + 1] - o;
  if (o + T > cap) return;
  const uint32_t* src = bp + unit_start[warp] + unit_len[warp] - T;
  for (unsigned long long i = lane; i < T; i += 32) out[o + i] = src[i];
}

// -----------------------------------------------------------------------------------------
// log_sum_exp, src/lattice.rs:321-333 (init_mode handled by the callers)
// -----------------------------------------------------------------------------------------
// exp / ln: glibc's algorithm restated bit for bit (tgx_libm.h), coefficients broadcast from
// constant memory, the two 2 KB lookup tables staged in shared memory by every block.
__constant__ unsigned long long c_exp_hdr[8];
__constant__ unsigned long long c_log_hdr[18];
__constant__ unsigned long long c_exp_tab[256];
__constant__ unsigned long long c_log_tab[256];

struct LibmTabs {
  const uint64_t* et;  // shared
  const double* lt;    // shared
};

__device__ __forceinline__ LibmTabs stage_libm_tables(unsigned long long* s_et, double* s_lt) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_et[i] = c_exp_tab[i];
    s_lt[i] = __longlong_as_double((long long)c_log_tab[i]);
  }
  __syncthreads();
  LibmTabs t;
  t.et = reinterpret_cast<const uint64_t*>(s_et);
  t.lt = s_lt;
  return t;
}

__device__ __forceinline__ double tgx_exp(double x, const LibmTabs& t) {
  return tgx_exp_impl(x, reinterpret_cast<const double*>(c_exp_hdr), t.et);
}
__device__ __forceinline__ double tgx_log(double x, const LibmTabs& t) {
  return tgx_log_impl(x, reinterpret_cast<const double*>(c_log_hdr), t.lt);
}

__device__ __forceinline__ double log_sum_exp(double x, double y, const LibmTabs& t) {
  double vmin, vmax;
  if (x > y) { vmin = y; vmax = x; } else { vmin = x; vmax = y; }
  if (vmax > __dadd_rn(vmin, 50.0)) return vmax;
  return __dadd_rn(vmax, tgx_log(__dadd_rn(tgx_exp(__dadd_rn(vmin, -vmax), t), 1.0), t));
}

// -----------------------------------------------------------------------------------------
// K4  forward pass: A[e] = fold over tokens ending at e, ascending start, of
//     log_sum_exp(., score + A[start]); first term assigns; nothing ends at e -> 0.0.
//     Lattice::populate_marginal alpha loop, src/lattice.rs:259-272 (per-position form).
// -----------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(WPB * 32) fb_forward_kernel(FbParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NG = 32 / G;
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lig = lane & (G - 1), gid = lane / G;
  const uint32_t W = u.W;
  WarpSmem s = carve(smem + (size_t)warp * warp_smem_bytes(u.rows, W, G), u.rows, W, G);
  double* wacc = s.wf + (size_t)gid * W;
  uint32_t* wseen = s.wu + (size_t)gid * W;
  __shared__ unsigned long long s_et[256];
  __shared__ double s_lt[256];
  const LibmTabs lt = stage_libm_tables(s_et, s_lt);

  const uint64_t gidx = ((uint64_t)blockIdx.x * WPB + warp) * NG + gid;
  const bool has = gidx < u.count;
  const uint32_t unit = has ? u.order[u.first + gidx] : 0;
  const uint32_t n = has ? u.unit_len[unit] : 0;
  const uint64_t start = has ? u.unit_start[unit] : 0;
  const uint8_t* text = u.text + start;
  double* A = p.A + start + unit;

  for (uint32_t i = lig; i < W; i += G) wseen[i] = 0;
  uint32_t nmax = n;
#pragma unroll
  for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
  __syncwarp();

  const uint32_t tiles = nmax / G + 1;
  uint32_t slot0 = 0;
  for (uint32_t tile = 0; tile < tiles; tile++) {
    const uint32_t p0 = tile * G;
    s.mcnt[lane] = walk_matches(u, text, p0 + lig, n, s, lane);
    __syncwarp();
    double my_a = 0.0;
    uint32_t sl = slot0;
#pragma unroll 1
    for (int j = 0; j < G; j++) {
      const uint32_t pp = p0 + j;
      const int src = gid * G + j;
      // alpha of every right node at pp; 0.0 when nothing ends here (src/lattice.rs:255, Q7)
      double a = 0.0;
      if (pp != 0 && wseen[sl]) a = wacc[sl];
      if (lig == j) my_a = a;
      __syncwarp();
      if (lig == 0) wseen[sl] = 0;
      const uint32_t c = (pp < n) ? s.mcnt[src] : 0;
      for (uint32_t k = lig; k < c; k += G) {
        const double y = __dadd_rn(s.mscore[k * ROW_STRIDE + src], a);  // nodes[lid].score + alpha[lid]
        uint32_t ts = sl + (s.mpack[k * ROW_STRIDE + src] >> 24);
        if (ts >= W) ts -= W;
        if (!wseen[ts]) {  // lid == end_nodes[pos][0]  -> init_mode
          wacc[ts] = y;
          wseen[ts] = 1;
        } else {
          wacc[ts] = log_sum_exp(wacc[ts], y, lt);
        }
      }
      __syncwarp();
      if (++sl == W) sl = 0;
    }
    const uint32_t e = p0 + lig;
    if (has && e <= n) A[e] = my_a;
    slot0 += G;
    while (slot0 >= W) slot0 -= W;
  }
  __syncwarp();
  if (has && lig == 0) {
    const double z = A[n];  // alpha[eos]  (src/lattice.rs:290-291)
    const double az = fabs(z);
    const bool normal = (az >= 2.2250738585072014e-308) && (az <= 1.7976931348623157e308);  // f64::is_normal
    p.status[unit] = normal ? 0 : 7;
  }
}

// -----------------------------------------------------------------------------------------
// K5  backward pass + expected counts.  B[p] = fold over tokens starting at p, ascending
//     length, of log_sum_exp(., score + B[p+len]) (src/lattice.rs:275-287); contribution
//     exp(A[p] + score + B[p+len] - z) added to expected[id] (:295-309).
// -----------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(WPB * 32) fb_backward_kernel(FbParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NG = 32 / G;
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lig = lane & (G - 1), gid = lane / G;
  const uint32_t W = u.W;
  WarpSmem s = carve(smem + (size_t)warp * warp_smem_bytes(u.rows, W, G), u.rows, W, G);
  double* wB = s.wf + (size_t)gid * W;
  __shared__ unsigned long long s_et[256];
  __shared__ double s_lt[256];
  const LibmTabs lt = stage_libm_tables(s_et, s_lt);

  const uint64_t gidx = ((uint64_t)blockIdx.x * WPB + warp) * NG + gid;
  bool has = gidx < u.count;
  const uint32_t unit = has ? u.order[u.first + gidx] : 0;
  if (has && p.status[unit] != 0) has = false;  // bad z: the reference panics; nothing is added
  const uint32_t n = has ? u.unit_len[unit] : 0;
  const uint64_t start = has ? u.unit_start[unit] : 0;
  const uint8_t* text = u.text + start;
  const double* A = p.A + start + unit;
  const double z = has ? A[n] : 0.0;

  uint32_t nmax = n;
#pragma unroll
  for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
  if (nmax == 0) return;
  // window slot of position q is q % W; beta at the end of the sentence is 0.0 (EOS)
  if (lig == 0) wB[n % W] = 0.0;
  __syncwarp();

  const uint32_t tiles = (nmax + G - 1) / G;
  for (uint32_t tile = tiles; tile-- > 0;) {
    const uint32_t p0 = tile * G;
    s.mcnt[lane] = walk_matches(u, text, p0 + lig, n, s, lane);
    const double a_mine = (has && p0 + lig < n) ? A[p0 + lig] : 0.0;
    __syncwarp();
    uint32_t sl = (p0 + G - 1) % W;  // slot of the tile's last position
#pragma unroll 1
    for (int j = G - 1; j >= 0; j--) {
      const uint32_t pp = p0 + j;
      const int src = gid * G + j;
      const double a = __shfl_sync(0xFFFFFFFFu, a_mine, src);
      const uint32_t c = (pp < n) ? s.mcnt[src] : 0;
      double b = 0.0;  // stays 0.0 when nothing begins at pp (Q7)
      for (uint32_t k = 0; k < c; k++) {  // ascending length = begin_nodes[pos] order
        const double sc = s.mscore[k * ROW_STRIDE + src];
        uint32_t ts = sl + (s.mpack[k * ROW_STRIDE + src] >> 24);
        if (ts >= W) ts -= W;
        const double y = __dadd_rn(sc, wB[ts]);  // nodes[rid].score + beta[rid]
        b = (k == 0) ? y : log_sum_exp(b, y, lt);
      }
      for (uint32_t k = lig; k < c; k += G) {
        const double sc = s.mscore[k * ROW_STRIDE + src];
        const uint32_t mp = s.mpack[k * ROW_STRIDE + src];
        uint32_t ts = sl + (mp >> 24);
        if (ts >= W) ts -= W;
        // total = a + score + b - z ; update = total.exp()   (src/lattice.rs:305-307)
        const double total = __dadd_rn(__dadd_rn(__dadd_rn(a, sc), wB[ts]), -z);
        const uint32_t id = mp & ID_MASK;
        double* dst = id < p.hot_k ? p.hot + (size_t)(blockIdx.x % p.hot_r) * p.hot_k + id : p.expected + id;
        atomicAdd(dst, tgx_exp(total, lt));
      }
      __syncwarp();
      if (pp < n && lig == 0) wB[sl] = b;
      __syncwarp();
      sl = sl == 0 ? W - 1 : sl - 1;
    }
  }
}

__global__ void fold_hot_kernel(const double* __restrict__ hot, uint32_t hot_k, uint32_t hot_r,
                                double* __restrict__ expected) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hot_k) return;
  double s = 0.0;
  for (uint32_t r = 0; r < hot_r; r++) s += hot[(size_t)r * hot_k + i];
  expected[i] += s;
}

}  // namespace tgxk
