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
#include "trie_build.h"

namespace tgxk {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t F_TERM = 2u << 24, F_HASCH = 4u << 24, ID_MASK = 0x00FFFFFFu;  // slot.y; see trie_build.h
constexpr int ROW_STRIDE = 33;  // padded row of the match buffer: conflict-free column reads
constexpr int WPB = 4;          // warps per block

struct UnitParams {
  const uint8_t* text;         // blob
  const uint64_t* unit_start;  // [U] absolute byte offset of the unit in text
  const uint32_t* unit_len;    // [U]
  const uint32_t* order;       // sorted unit indices (length-descending)
  // This launch handles order[first .. first+count).  When `counts` is set the range is read on the
  // device (counts[0] = #units at least as long as the "long" threshold, counts[1] = #non-empty
  // units), so that the host never has to wait for the sort in the middle of a call:
  // part 0 = every non-empty unit, 1 = the long ones, 2 = the short ones; 3 / 4 = the same split at
  // the second threshold (counts[2]: pair-CTA kernel vs lane kernel).  `count` is then only the
  // upper bound the grid was sized for.
  uint32_t first, count;
  const uint32_t* counts;
  int part;
  const uint4* trie;
  uint32_t root_base;  // xbase of the root
  uint32_t rows;  // match-buffer rows = max token length
  uint32_t W;     // window slots = rows + 1
};

struct ViterbiParams {
  UnitParams u;
  uint8_t* bp;  // [N] byte length of the best last token per end position (0 = unreachable)
  // dropout in (0, 1) (viterbi_kernel<G, true> only): see drop_draw
  double dropout;
  unsigned long long drop_seed, unit_base;  // unit_base = index of this launch's unit 0 within the caller's batch
};

// dropout (src/model.rs:100): the reference draws rand::random::<f64>() — an unseeded thread_rng — once per
// multi-byte candidate of a reachable position, so only the DISTRIBUTION of its output is defined.  Here the draw
// of candidate (sample, start position, length) is a pure function of a caller-provided seed (two rounds of the
// splitmix64 finaliser), which gives independent uniform draws with the same keep rule (`dropout < u`), is
// reproducible, and is restated in the oracle (orc_encode_keyed) for bit-exact parity tests.
__host__ __device__ __forceinline__ unsigned long long drop_mix(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ unsigned long long drop_unit_key(unsigned long long seed, unsigned long long sample) {
  return drop_mix(seed + 0x9E3779B97F4A7C15ULL * (sample + 1));
}
__host__ __device__ __forceinline__ double drop_draw(unsigned long long unit_key, unsigned long long pos, uint32_t len) {
  const unsigned long long z = drop_mix(unit_key + 0x9E3779B97F4A7C15ULL * ((pos << 8) | len));
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);  // uniform in [0, 1), 53 bits
}

struct FbParams {
  UnitParams u;
  double* A;          // [N + U] forward log-probabilities: unit k owns A[start_k + k .. + n_k]
  int32_t* status;    // [U] 0 ok / 7 bad z
  // Expected counts are accumulated in 192-bit fixed point with integer atomics (acc_add below): acc[3 * id + 0 / 1]
  // = fraction bits 65..128 / 1..64, acc[3 * id + 2] = integer part.  The sum of the contributions (each truncated to
  // 2^-128) is exact, so it does not depend on the order of the additions: counts are bit-identical from run to run,
  // for every chunking and every number of GPUs.
  unsigned long long* acc;  // [V][3]
  // Hot tokens (small ids: vocabularies are score-sorted) would serialise every SM's atomics
  // on a handful of L2 addresses, so ids < hot_k accumulate into one of hot_r replicas
  // (picked per block) that fold_hot_acc_kernel sums into acc[] afterwards.
  unsigned long long* hot_acc;  // [hot_r][hot_k][3]
  uint32_t hot_k, hot_r;
  // populate_nodes' dropout (src/model.rs:48-50; fb_*_kernel<G, .., true> only): the draw of the multi-byte match
  // (start byte, length) is keyed by the byte's offset in the call's text (+ drop_base, the shard's offset in a
  // sharded corpus), so the forward and the backward kernel — which both re-derive the matches — see one lattice.
  double dropout;
  unsigned long long drop_key, drop_base;
};

constexpr int ACC_LIMBS = 3;  // u64 words per accumulator
// c -> integer part + 128 fraction bits in two 64-bit words, added with integer atomics; a carry out of a word goes
// into the next one as a separate atomic (the additions commute, so the sum stays exact whatever the order).
// (Measured alternative: four 32-bit chunks in counters of their own, no carries and so no atomic that has to return
//  a value, but two reductions for a typical contribution instead of one: 87 ms per GB in the counts kernel against
//  69 for this form and 35 for f64 atomics, which are not reproducible.)
__device__ __forceinline__ void acc_add(unsigned long long* slot, double c) {
  unsigned long long ip = 0;
  double fr = c;
  if (c >= 1.0) {  // a probability: at most 1 up to rounding
    ip = __double2ull_rz(c);
    fr = c - (double)ip;
  }
  const double t = fr * 18446744073709551616.0;          // exact (a power of two)
  unsigned long long v1 = __double2ull_rz(t);            // fraction bits 1..64
  // t >= 2^53 is an integer, so the remainder is zero; below that the truncation and the difference are exact
  const unsigned long long v0 = __double2ull_rz((t - (double)v1) * 18446744073709551616.0);  // fraction bits 65..128
  if (v0) {
    const unsigned long long old = atomicAdd(slot, v0);
    if (old + v0 < old) v1++;  // (v1 <= 2^64 - 2^11: no overflow)
  }
  if (v1) {
    const unsigned long long old = atomicAdd(slot + 1, v1);
    if (old + v1 < old) ip++;
  }
  if (ip) atomicAdd(slot + 2, ip);
}
__device__ __forceinline__ unsigned long long* acc_slot(const FbParams& p, uint32_t bid, uint32_t id) {
  return id < p.hot_k ? p.hot_acc + ((size_t)(bid % p.hot_r) * p.hot_k + id) * ACC_LIMBS : p.acc + (size_t)id * ACC_LIMBS;
}

__device__ __forceinline__ void unit_range(const uint32_t* counts, int part, uint32_t& first, uint32_t& count) {
  if (!counts) return;
  const uint32_t nl = counts[0], nn = counts[1];
  if (part == 1) { first = 0; count = nl; }
  else if (part == 2) { first = nl; count = nn - nl; }
  else if (part == 3) { first = 0; count = min(counts[2], nn); }
  else if (part == 4) { const uint32_t n2 = min(counts[2], nn); first = n2; count = nn - n2; }
  else { first = 0; count = nn; }
}

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
template <bool DROP = false>
__device__ __forceinline__ uint32_t walk_matches(const UnitParams& u, const uint8_t* text, uint32_t pos,
                                                 uint32_t n, const WarpSmem& s, int lane, double dropout = 0.0,
                                                 unsigned long long unit_key = 0, unsigned long long pos_base = 0) {
  uint32_t cnt = 0;
  if (pos < n) {
    uint32_t xb = u.root_base;
    uint32_t d = 0;
    uint32_t maxd = min(n - pos, u.rows);
    while (d < maxd) {
      const uint32_t cw = 0x100u | __ldg(text + pos + d);
      uint4 e = __ldg(u.trie + (xb ^ cw));
      if ((e.x ^ cw) & 0x1FFu) break;
      d++;
      // (the draw of a dropped candidate does not depend on the position being reachable, so it can be taken here:
      //  an unreachable start never relaxes anything either way, src/model.rs:85-87)
      if ((e.y & F_TERM) && (!DROP || d <= 1 || dropout < drop_draw(unit_key, pos_base + pos, d))) {
        s.mscore[cnt * ROW_STRIDE + lane] = __hiloint2double((int)e.w, (int)e.z);
        s.mpack[cnt * ROW_STRIDE + lane] = (d << 24) | (e.y & ID_MASK);
        cnt++;
      }
      if (!(e.y & F_HASCH)) break;
      xb = e.x >> 9;
    }
  }
  return cnt;
}

// -----------------------------------------------------------------------------------------
// K2g  Viterbi forward, lane-group form (any max_token_len <= 64; the fallback when the
//      vocabulary has tokens longer than 16 bytes).  Model::encode forward loop,
//      src/model.rs:83-110.  Output: bp[start + e - 1] = byte length of the best last token
//      ending at position e (0 = position unreachable).
// -----------------------------------------------------------------------------------------
template <int G, bool DROP = false>
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
  uint32_t ufirst = u.first, ucount = u.count;
  unit_range(u.counts, u.part, ufirst, ucount);
  const bool has = gidx < ucount;
  const uint32_t unit = has ? u.order[ufirst + gidx] : 0;
  const uint32_t n = has ? u.unit_len[unit] : 0;
  const uint64_t start = has ? u.unit_start[unit] : 0;
  const uint8_t* text = u.text + start;
  const unsigned long long unit_key = DROP ? drop_unit_key(p.drop_seed, p.unit_base + unit) : 0ull;

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
    s.mcnt[lane] = walk_matches<DROP>(u, text, p0 + lig, n, s, lane, p.dropout, unit_key);
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
    if (has && e >= 1 && e <= n) p.bp[start + e - 1] = (my_bp == NONE) ? (uint8_t)0 : (uint8_t)(my_bp >> 24);
    slot0 += G;
    while (slot0 >= W) slot0 -= W;
  }
}

// -----------------------------------------------------------------------------------------
// K2  Viterbi forward, pair-CTA form (max_token_len <= 16): the encode hot loop.
//
// The ordered relax chain of a sample (position p must be final before its matches are
// pushed; src/model.rs:83-110) is one f64 add + compare per position and cannot be split
// without changing roundings, so the kernel is organised around it:
//   * a CTA works on TWO samples at a time.  ONE consumer warp owns both chains: lanes
//     0-15 the first sample, lanes 16-31 the second.  Lane g of a half owns the dp cell of
//     every position q with q % 16 == g (a token is at most 16 bytes, so the 16 cells
//     ahead of the current position are exactly the cells in flight).  Finalising position
//     p is a 16-wide shuffle broadcast of its owner's score; the relax of "token of length
//     len starting at p" is an add + compare in lane (p + len) % 16.  The cell of p is
//     re-used for p + 16 by letting its owner take the length-16 candidate unconditionally.
//   * 2R producer warps walk the double-array trie, one start position per lane, and park
//     the scores in a dense shared-memory table indexed [start position][target cell]
//     (row stride 17 doubles: conflict-free for both the row-wise producer and the
//     column-wise consumer); -inf = no such token.  The dp keeps only the START index of the
//     best candidate; the back length (1 byte per position, the only HBM output of this
//     kernel) is derived when the cell is final.  Token ids are recovered later, in
//     parallel, by emit_kernel.
//   * rounds of R tiles (32 positions) per sample are double-buffered and separated by one
//     __syncthreads(); the consumer's half-warp leaders schedule the next round and fetch
//     new samples from a global counter in length-descending order (LPT).
// -----------------------------------------------------------------------------------------
constexpr int PT_ROW = 17;
constexpr int PT_TILE = 32 * PT_ROW;  // doubles per (sample slot, tile)

struct __align__(16) PairInfo {
  unsigned long long start;
  uint32_t n, tile0, ntiles;
  int32_t unit;  // < 0: nothing to do
  uint32_t pad[2];
};
static_assert(sizeof(PairInfo) == 32, "PairInfo is 32 bytes");

struct PairParams {
  UnitParams u;
  const uint8_t* blob_end;
  uint8_t* bp;  // [N] back length per end position
  unsigned int* counter;
  uint32_t hot_slots;  // leading trie slots staged in shared memory (covers HOT levels)
  uint32_t groups;     // consumer/producer groups per CTA
  uint32_t dbg;        // developer timing experiments (tools/probe.py): 1 = skip walks, 2 = skip the dp
};
// dropout in (0, 1): a second kernel parameter of viterbi_pair_drop_kernel only (see drop_draw) — the default
// kernel's parameter block and code stay exactly what they were (its 64-register shape is sensitive to both)
struct DropInfo {
  double dropout;
  unsigned long long seed, unit_base;
};

// shared memory of one CTA: [hot trie prefix][per group: tables 2 x 2 x R tiles | 4 x 2 PairInfo]
__host__ __device__ inline size_t pair_group_bytes(int R) { return (size_t)2 * 2 * R * PT_TILE * 8 + 8 * 32; }
__host__ __device__ inline size_t pair_smem_bytes(int R, uint32_t groups, uint32_t hot_slots) {
  return (size_t)hot_slots * 16 + (size_t)groups * pair_group_bytes(R);
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
  const unsigned long long lo = w[g], hi = w[g + 1];
  return sh ? ((lo >> (8 * sh)) | (hi << (64 - 8 * sh))) : lo;
}

// Phase A for one start position: TrieIterator::next (src/trie.rs:51-63) unrolled over the 16
// possible depths.  The walk may run past the end of the sample (into the next sample's bytes):
// such a token lands on a dp cell beyond position n, which is never read.
template <int HOT, bool DROP = false>
__device__ __forceinline__ void pair_produce(const uint4* __restrict__ trie, const uint4* __restrict__ hot,
                                             uint32_t root, const unsigned long long (&w)[3], uint32_t sh, bool active,
                                             double* row, int lane, double dropout = 0.0,
                                             unsigned long long unit_key = 0, uint32_t pos = 0) {
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
#pragma unroll
  for (int c = 0; c < 16; c++) row[c] = ninf;
  if (!active) return;
  unsigned char* rb = reinterpret_cast<unsigned char*>(row);
  const uint32_t l18 = (uint32_t)(lane + 1) * 8u;
  uint32_t xb = root;
#pragma unroll
  for (int g = 0; g < 2; g++) {
    const unsigned long long a = window_bytes(w, sh, g);
    const uint32_t alo = (uint32_t)a, ahi = (uint32_t)(a >> 32);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int d = g * 8 + k;
      const uint32_t cw = __byte_perm(k < 4 ? alo : ahi, 1u, 0x5540 + (k & 3));  // 0x100 | byte
      // the first HOT levels of the trie live in shared memory (every probe out of a node of
      // depth < HOT lands in the staged prefix, hit or miss; see trie_build.h)
      const uint4 e = (d < HOT) ? hot[xb ^ cw] : __ldg(trie + (xb ^ cw));
      if ((e.x ^ cw) & 0x1FFu) return;
      // target cell (start + len) % 16 = (lane + d + 1) % 16; a dropped multi-byte candidate (src/model.rs:100) is
      // simply never parked
      if ((e.y & F_TERM) && (!DROP || d == 0 || dropout < drop_draw(unit_key, pos, (uint32_t)d + 1u)))
        *reinterpret_cast<double*>(rb + ((l18 + 8u * d) & 120u)) = __hiloint2double((int)e.w, (int)e.z);
      if (!(e.y & F_HASCH)) return;
      xb = e.x >> 9;
    }
  }
}

// Phase B over one 32-position tile for both halves of the consumer warp.  tb = this half's
// table + g, so the operand of step j is tb[j * PT_ROW].  "Unreached" is best == -inf: scores
// are finite, so a candidate built on an unreached position is -inf and can never win, and the
// first finite candidate always replaces -inf (the reference's `start.is_none() ||`, src/model.rs:100).
__device__ __forceinline__ void pair_consume(const double* __restrict__ tb, int g, double& best, uint32_t& ps,
                                             uint32_t& len0, uint32_t& len1) {
  double sc[4];
  sc[0] = tb[0];
  sc[1] = tb[PT_ROW];
  uint32_t sv_hi = 0, sv_ps = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    if (j + 2 < 32) sc[(j + 2) & 3] = tb[(j + 2) * PT_ROW];
    const double bs = __shfl_sync(0xFFFFFFFFu, best, j & 15, 16);  // dp[pos].score, final
    const bool own = g == (j & 15);
    if (own) {  // this lane's cell is position j of the tile: keep its result, the cell moves on to j + 16
      sv_hi = (uint32_t)__double2hiint(best);
      sv_ps = ps;
    }
    const double cand = __dadd_rn(bs, sc[j & 3]);  // dp[pos].score + vocab[id].score  (src/model.rs:98)
    if (cand > best || own) {                      // (:100-101); a fresh cell takes its first candidate
      best = cand;
      ps = j;
    }
    if ((j & 15) == 15) {
      const uint32_t l = (sv_hi == 0xFFF00000u) ? 0u : (((uint32_t)(j - 15 + g) - sv_ps) & 31u);
      if (j == 15) len0 = l; else len1 = l;
    }
  }
}

// Body shared by viterbi_pair_kernel and the hybrid kernel; called by every thread of the CTA
// (warps beyond p.groups * WG only help staging the hot trie prefix).
template <int R, int HOT, bool DROP = false>
__device__ __forceinline__ void pair_body(const PairParams& p, unsigned char* smem, const DropInfo* di = nullptr) {
  constexpr int WG = 2 * R + 1;  // warps per group: consumer + 2R producers
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / WG, wg = warp % WG;
  const bool spare = (uint32_t)grp >= p.groups;
  const int h = lane >> 4, g = lane & 15;
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);

  const uint4* hot = reinterpret_cast<const uint4*>(smem);
  unsigned char* gbase = smem + (size_t)p.hot_slots * 16 + (size_t)grp * pair_group_bytes(R);
  double* tab = reinterpret_cast<double*>(gbase);  // [2 stages][2 halves][R][PT_TILE]
  PairInfo* s_info = reinterpret_cast<PairInfo*>(gbase + (size_t)2 * 2 * R * PT_TILE * 8);  // [4][2]
  if (HOT > 0) {
    uint4* hw = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < p.hot_slots; i += blockDim.x) hw[i] = __ldg(u.trie + i);
  }

  // consumer state: one dp cell per lane
  double best = ninf;
  uint32_t ps = 0;
  // producer state: text window prefetched for the next round
  unsigned long long pw[3] = {0, 0, 0};
  uint32_t psh = 0, pf_tile = 0;
  int32_t pf_unit = -1;
  // scheduler state (half-warp leaders of the consumer warp)
  PairInfo cur;
  cur.unit = -1; cur.start = 0; cur.n = 0; cur.tile0 = 0; cur.ntiles = 0; cur.pad[0] = cur.pad[1] = 0;
  uint32_t ufirst = u.first, ucount = u.count;
  unit_range(u.counts, u.part, ufirst, ucount);
  auto fetch = [&]() {
    const uint32_t idx = atomicAdd(p.counter, 1u);
    if (idx < ucount) {
      cur.unit = (int32_t)u.order[ufirst + idx];
      cur.n = u.unit_len[cur.unit];
      cur.start = u.unit_start[cur.unit];
      cur.tile0 = 0;
      cur.ntiles = cur.n / 32 + 1;  // positions 0..n
    } else {
      cur.unit = -1;
    }
  };
  if (!spare && wg == 0 && g == 0) {
    fetch();
    s_info[0 * 2 + h] = cur;
    PairInfo none = cur;
    none.unit = -1;
    s_info[3 * 2 + h] = none;
  }
  __syncthreads();
  if (spare) return;

  for (uint32_t r = 0;; r++) {
    if (wg == 0) {
      if (g == 0) {  // what the producers do in round r + 1
        if (cur.unit >= 0) {
          cur.tile0 += R;
          if (cur.tile0 >= cur.ntiles) fetch();
        }
        s_info[((r + 1) & 3) * 2 + h] = cur;
      }
      const PairInfo ci = s_info[((r + 3) & 3) * 2 + h];  // what they did in round r - 1
#pragma unroll
      for (int k = 0; k < R; k++) {
        const uint32_t t = ci.tile0 + k;
        const bool act = ci.unit >= 0 && t < ci.ntiles;
        if (act && t == 0) {  // dp[0] = { score 0.0, start Some(0) }  (src/model.rs:72-81); the rest unreached
          best = (g == 0) ? 0.0 : ninf;
          ps = 0;
        }
        const double* tb = tab + ((size_t)(((r + 1) & 1) * 2 + h) * R + k) * PT_TILE + g;
        uint32_t len0, len1;
        if (p.dbg & 2u) { len0 = len1 = 1; } else
        pair_consume(tb, g, best, ps, len0, len1);  // both halves always run it (full-warp shuffles)
        if (act) {
          const uint32_t e0 = t * 32 + g, e1 = e0 + 16;
          if (e0 >= 1 && e0 <= ci.n) p.bp[ci.start + e0 - 1] = (uint8_t)len0;
          if (e1 <= ci.n) p.bp[ci.start + e1 - 1] = (uint8_t)len1;
        }
      }
    } else {
      const int w = wg - 1;
      const int ph = w & 1, k = w >> 1;
      const PairInfo pi = s_info[(r & 3) * 2 + ph];
      const uint32_t t = pi.tile0 + k;
      if (pi.unit >= 0 && t < pi.ntiles) {
        const uint32_t pos = t * 32 + lane;
        double* row = tab + ((size_t)((r & 1) * 2 + ph) * R + k) * PT_TILE + lane * PT_ROW;
        const uint8_t* ptr = u.text + pi.start + pos;
        // the text of this tile was requested a round ago (HBM latency off the round's critical path)
        if (!(pf_unit == pi.unit && pf_tile == t)) load_window(ptr, p.blob_end, pw, psh);
        unsigned long long w3[3] = {pw[0], pw[1], pw[2]};
        const uint32_t sh = psh;
        if (t + R < pi.ntiles) {
          load_window(ptr + 32 * R, p.blob_end, pw, psh);
          pf_unit = pi.unit;
          pf_tile = t + R;
        } else {
          pf_unit = -1;
        }
        if constexpr (DROP)
          pair_produce<HOT, true>(u.trie, hot, u.root_base, w3, sh, pos < pi.n, row, lane, di->dropout,
                                  drop_unit_key(di->seed, di->unit_base + (uint32_t)pi.unit), pos);
        else
          pair_produce<HOT>(u.trie, hot, u.root_base, w3, sh, pos < pi.n && !(p.dbg & 1u), row, lane);
      }
    }
    // group barrier: the groups of a CTA only share the read-only hot trie
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(WG * 32) : "memory");
    const bool more = s_info[(r & 3) * 2 + 0].unit >= 0 || s_info[(r & 3) * 2 + 1].unit >= 0 ||
                      s_info[((r + 1) & 3) * 2 + 0].unit >= 0 || s_info[((r + 1) & 3) * 2 + 1].unit >= 0;
    if (!more) break;
  }
}

// MAXT = threads the kernel is compiled for: 800 (5 groups of 5 warps, 72 registers, no spills: the lowest latency per
// chain) or 960 (6 groups when R = 2, 10 when R = 1; 64 registers).
template <int R, int HOT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) viterbi_pair_kernel(PairParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  pair_body<R, HOT>(p, smem);
}

// The same kernel with the keyed dropout draw in its producers (src/model.rs:100).
template <int R, int HOT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) viterbi_pair_drop_kernel(PairParams p, DropInfo di) {
  extern __shared__ __align__(16) unsigned char smem[];
  pair_body<R, HOT, true>(p, smem, &di);
}

// -----------------------------------------------------------------------------------------
// K3a  backtrack (src/model.rs:113-123): one thread per sample follows the back lengths from
//      position n and marks every token END with the token's length (mark[] is zeroed by the
//      caller and separate from bp[], so the chain's loads stay L1 hits).
// -----------------------------------------------------------------------------------------
struct BacktrackParams {
  const uint64_t* unit_start;
  const uint32_t* unit_len;
  const uint32_t* order;
  uint32_t first, count;   // see UnitParams
  const uint32_t* counts;
  int part;
  const uint8_t* bp;
  uint8_t* mark;                 // [N]
  unsigned long long* n_tokens;  // [U]
  int32_t* status;               // [U] 0 ok / 6 NoPath
};

// Short samples: one thread per sample follows the chain with plain dependent loads.
constexpr int BT_THREADS = 128;

__global__ void __launch_bounds__(BT_THREADS) backtrack_thread_kernel(BacktrackParams p) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t pfirst = p.first, pcount = p.count;
  unit_range(p.counts, p.part, pfirst, pcount);
  if (i >= pcount) return;
  const uint32_t unit = p.order[pfirst + i];
  const uint32_t n = p.unit_len[unit];
  const uint64_t start = p.unit_start[unit];
  const uint8_t* b = p.bp + start;
  uint8_t* mk = p.mark + start;
  unsigned long long k = 0;
  int st = 0;
  if (n > 0) {
    if (__ldg(b + n - 1) == 0) {
      st = 6;  // Error::NoPath(n, n)  (src/model.rs:119)
    } else {
      uint32_t pos = n;
      while (pos > 0) {
        const uint32_t l = __ldg(b + pos - 1);
        if (l == 0 || l > pos) { st = 99; break; }  // corrupt chain: never loop forever
        mk[pos - 1] = (uint8_t)l;
        pos -= l;
        k++;
      }
    }
  }
  p.n_tokens[unit] = st ? 0ull : k;
  p.status[unit] = st;
}

// Long samples: one WARP per sample, 1024 positions at a time, so that the serial part of the
// chain pos -> pos - len[pos] shrinks from one dependent step per token to one per 32 positions:
//   phase 1  lane s owns the 32-position segment s of the chunk and computes, for EVERY position
//            x of its segment, where the chain starting at x leaves the segment (exit[x] =
//            exit[x - len[x]] unless x - len[x] is already below the segment): 32 short,
//            lane-private steps, all lanes in parallel;
//   phase 2  the true chain enters the top segment at the chunk's top position; hopping
//            segment to segment through exit[] (one shared-memory load per segment) yields the
//            entry position of every segment;
//   phase 3  lane s re-walks its own segment from its entry and marks the token ends.
// Tokens are at most 64 bytes (tgx::MAX_TOKEN_LEN), so a hop lands at most two segments down
// and the next chunk's top at most 63 bytes below this chunk.  The back lengths of the next
// chunk are prefetched into registers by the whole warp (warp-uniform, so the per-warp
// scoreboard costs nothing).
constexpr int BW_WARPS = 8;
constexpr int BW_CHUNK = 1024;
constexpr int BW_VECS = 72;                 // staged 16-byte vectors: window [wbase, wbase + 1152)
constexpr int BW_SEG_STRIDE = 34;           // int16 per segment row (bank skew)

__global__ void __launch_bounds__(BW_WARPS * 32) backtrack_warp_kernel(BacktrackParams p) {
  __shared__ __align__(16) uint8_t s_len[BW_WARPS][BW_VECS * 16];
  __shared__ int16_t s_exit[BW_WARPS][32 * BW_SEG_STRIDE];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * BW_WARPS + w;
  uint32_t pfirst = p.first, pcount = p.count;
  unit_range(p.counts, p.part, pfirst, pcount);
  if (i >= pcount) return;
  const uint32_t unit = p.order[pfirst + i];
  const uint32_t n = p.unit_len[unit];
  const uint64_t start = p.unit_start[unit];
  const uint4* b16 = reinterpret_cast<const uint4*>(p.bp);
  uint8_t* sl = s_len[w];
  int16_t* ex = s_exit[w] + lane * BW_SEG_STRIDE;  // this lane's segment row, indices 1..32
  unsigned long long k = 0;
  int st = 0;
  if (n == 0) {  // (not scheduled: empty samples are filtered by the caller)
    if (lane == 0) { p.n_tokens[unit] = 0; p.status[unit] = 0; }
    return;
  }
  uint64_t gtop = start + n - 1;                                           // last byte of the current token
  uint64_t gbase = (gtop - start >= BW_CHUNK) ? gtop - (BW_CHUNK - 1) : start;  // first byte of the chunk
  uint64_t wbase = gbase & ~15ull;                                         // staged window
  uint4 v[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const uint32_t vi = lane + 32 * j;
    v[j] = (vi < BW_VECS && wbase + 16ull * vi <= gtop) ? __ldg(b16 + (wbase >> 4) + vi) : make_uint4(0, 0, 0, 0);
  }
  for (;;) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const uint32_t vi = lane + 32 * j;
      if (vi < BW_VECS) reinterpret_cast<uint4*>(sl)[vi] = v[j];
    }
    __syncwarp();
    const uint32_t C = (uint32_t)(gtop - gbase) + 1;   // positions x = 1..C <-> byte gbase + x - 1
    const uint32_t off = (uint32_t)(gbase - wbase);    // sl[off + x - 1] = back length of position x
    if (gtop == start + n - 1 && sl[off + C - 1] == 0) { st = 6; break; }  // Error::NoPath(n, n)  (src/model.rs:119)
    // prefetch the next chunk's window: its top is at most 63 bytes below gbase
    const bool last = gbase == start;
    const uint64_t nwbase = (gbase >= 1087 ? gbase - 1087 : 0) & ~15ull;
    if (!last) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const uint32_t vi = lane + 32 * j;
        v[j] = (vi < BW_VECS && nwbase + 16ull * vi < gbase) ? __ldg(b16 + (nwbase >> 4) + vi) : make_uint4(0, 0, 0, 0);
      }
    }
    // ---- phase 1
    const int lo = 32 * lane;
#pragma unroll 4
    for (int q = 1; q <= 32; q++) {
      const int x = lo + q;
      if (x > (int)C) break;
      const int l = sl[off + x - 1];
      const int t = x - l;
      int e;
      if (l == 0) e = x;            // unreachable position: never on the chain
      else if (t <= lo) e = t;      // leaves the segment (may be <= 0: leaves the chunk)
      else e = ex[t - lo];
      ex[q] = (int16_t)e;
    }
    __syncwarp();
    // ---- phase 2 (every lane follows the same hops; broadcast loads)
    int cur = (int)C, myent = 0;
    while (cur >= 1) {
      const int s = (cur - 1) >> 5;
      if (s == lane) myent = cur;
      const int nxt = (int)s_exit[w][s * BW_SEG_STRIDE + (cur - 32 * s)];
      if (nxt >= cur) { cur = -32768; break; }  // corrupt chain: never loop forever
      cur = nxt;
    }
    // ---- phase 3
    if (cur != -32768) {
      int x = myent;
      while (x > lo) {
        const int l = sl[off + x - 1];
        if (l == 0) break;
        p.mark[gbase + x - 1] = (uint8_t)l;
        k++;
        x -= l;
      }
    }
    __syncwarp();
    // cur <= 0 is where the chain left the chunk; as a sample position: (gbase - start) + cur
    const long long pe = (long long)(gbase - start) + cur;
    if (cur == -32768 || cur > 0 || pe < 0) { st = 99; break; }  // corrupt chain
    if (pe == 0) break;                                           // reached position 0: done
    gtop = gbase + cur - 1;  // the end of the first token that ends below this chunk
    gbase = (gtop - start >= BW_CHUNK) ? gtop - (BW_CHUNK - 1) : start;
    wbase = nwbase;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) k += __shfl_xor_sync(0xFFFFFFFFu, k, o);
  if (lane == 0) {
    p.n_tokens[unit] = st ? 0ull : k;
    p.status[unit] = st;
  }
}

// -----------------------------------------------------------------------------------------
// K3b/K3c  emit: stream-compact the marked token ends (in blob order = output order, since
//      samples are contiguous and in input order) and recover each token's id by walking the
//      trie over its bytes — thousands of independent walks instead of one id per dp relax.
//      With freq != nullptr it is the frequency pass of prune_vocab (src/prune.rs:223-225).
// -----------------------------------------------------------------------------------------
constexpr int EM_BLOCK = 256;
constexpr int EM_PER_THREAD = 16;
constexpr int EM_TILE = EM_BLOCK * EM_PER_THREAD;
constexpr uint32_t EM_HOT = 4096;  // ids below this are counted in shared memory first (frequency pass)

__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t w) {  // 0x80 in every non-zero byte
  return (((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u;
}

__device__ __forceinline__ uint32_t em_block_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  uint32_t before = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < EM_BLOCK / 32; i++) {
    const uint32_t ws = warp_sums[i];
    if (i < warp) before += ws;
    tot += ws;
  }
  total = tot;
  return before + x - v;
}

// mark is padded with zeros to a multiple of EM_TILE bytes.
__global__ void __launch_bounds__(EM_BLOCK) mark_count_kernel(const uint4* __restrict__ mark, uint64_t n_tiles,
                                                              unsigned long long* __restrict__ tile_cnt) {
  __shared__ uint32_t ws[EM_BLOCK / 32];
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint4 v = mark[tile * EM_BLOCK + threadIdx.x];
    uint32_t c = __popc(nonzero_bytes(v.x)) + __popc(nonzero_bytes(v.y)) + __popc(nonzero_bytes(v.z)) +
                 __popc(nonzero_bytes(v.w));
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
#pragma unroll
      for (int i = 0; i < EM_BLOCK / 32; i++) t += ws[i];
      tile_cnt[tile] = t;
    }
    __syncthreads();
  }
}

struct EmitParams {
  const uint4* mark;
  const uint8_t* text;
  uint64_t text_bytes;  // readable bytes at text (rounded up to 16)
  uint64_t n_tiles;
  const unsigned long long* tile_prefix;  // exclusive scan of tile_cnt
  const uint4* trie;
  uint32_t root_base;  // xbase of the root
  uint32_t* ids;             // may be null
  unsigned long long cap;
  unsigned long long* freq;  // may be null
  uint32_t V;
  // token hash (trie_build.h; built when max_token_len <= 16): ONE probe per token instead of one dependent trie
  // probe per byte.  hash_mask == 0: not available, walk the trie.
  const uint4* hash;
  uint32_t hash_mask;
  unsigned long long hash_seed;
};

__global__ void __launch_bounds__(EM_BLOCK) emit_kernel(EmitParams p) {
  __shared__ uint32_t ws[EM_BLOCK / 32];
  __shared__ uint32_t s_tok[EM_TILE];  // (len << 16) | offset of the token's last byte in the tile
  __shared__ __align__(16) uint8_t s_text[EM_TILE + 48];  // bytes tile * EM_TILE - 16 .. (hash path; +32: word reads past the tile)
  extern __shared__ uint32_t s_hot[];  // [EM_HOT] when freq
  const bool staged = p.hash_mask && (reinterpret_cast<unsigned long long>(p.text) & 15ull) == 0;
  if (p.freq) {
    for (uint32_t i = threadIdx.x; i < EM_HOT; i += EM_BLOCK) s_hot[i] = 0;
    __syncthreads();
  }
  for (uint64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const uint4 v = p.mark[tile * EM_BLOCK + threadIdx.x];
    if (staged) {  // (text is padded like mark: reading the whole last tile stays inside the allocation)
      const uint4* t16 = reinterpret_cast<const uint4*>(p.text) + tile * EM_BLOCK;
      // (the last vector of a caller-owned blob is read byte by byte: nothing beyond text_bytes is touched)
      const unsigned long long vb = tile * EM_TILE + 16ull * threadIdx.x;
      uint4 tv = make_uint4(0, 0, 0, 0);
      if (vb + 16 <= p.text_bytes) {
        tv = __ldg(t16 + threadIdx.x);
      } else if (vb < p.text_bytes) {
        uint32_t w[4] = {0, 0, 0, 0};
        for (uint32_t k = 0; vb + k < p.text_bytes; k++) w[k >> 2] |= (uint32_t)__ldg(p.text + vb + k) << (8 * (k & 3));
        tv = make_uint4(w[0], w[1], w[2], w[3]);
      }
      reinterpret_cast<uint4*>(s_text)[threadIdx.x + 1] = tv;
      if (threadIdx.x == 0) reinterpret_cast<uint4*>(s_text)[0] = tile ? __ldg(t16 - 1) : make_uint4(0, 0, 0, 0);
    }
    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) c += __popc(nonzero_bytes(wv[i]));
    uint32_t total;
    uint32_t o = em_block_scan(c, ws, total);
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const uint32_t l = (wv[i] >> (8 * b)) & 0xFFu;
        if (l) s_tok[o++] = (l << 16) | (uint32_t)(threadIdx.x * EM_PER_THREAD + i * 4 + b);
      }
    }
    __syncthreads();
    const unsigned long long base = p.tile_prefix[tile];
    const uint8_t* tt = p.text + tile * EM_TILE;
    for (uint32_t k = threadIdx.x; k < total; k += EM_BLOCK) {
      const uint32_t tk = s_tok[k];
      const uint32_t len = tk >> 16;
      uint32_t id;
      if (p.hash_mask) {
        unsigned long long lo = 0, hi = 0;
        if (staged) {  // token bytes from the staged tile (16-byte halo in front: a token may start in the previous tile)
          // 16 bytes from the token's first byte: five aligned words + funnel shifts, then cut to `len`
          const uint32_t o = 16u + (tk & 0xFFFFu) + 1u - len;
          const uint32_t* w = reinterpret_cast<const uint32_t*>(s_text) + (o >> 2);
          const uint32_t sh = 8u * (o & 3u);
          const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
          lo = (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
          hi = (unsigned long long)__funnelshift_r(w2, w3, sh) | ((unsigned long long)__funnelshift_r(w3, w4, sh) << 32);
          if (len < 8) {
            lo &= (1ull << (8 * len)) - 1ull;
            hi = 0;
          } else if (len < 16) {
            hi &= (1ull << (8 * (len - 8))) - 1ull;
          }
        } else {
          const uint8_t* q = tt + (tk & 0xFFFFu) + 1 - len;
          for (uint32_t d = 0; d < len && d < 8; d++) lo |= (unsigned long long)__ldg(q + d) << (8 * d);
          for (uint32_t d = 8; d < len; d++) hi |= (unsigned long long)__ldg(q + d) << (8 * (d - 8));
        }
        const unsigned long long key = tgx::token_key(lo, hi, len, p.hash_seed);
        uint32_t s = tgx::token_key_slot(key, p.hash_mask);
        id = NONE;
        for (int it = 0; it < 256; it++) {
          const uint4 e = __ldg(p.hash + s);
          if ((((unsigned long long)e.y << 32) | e.x) == key) { id = e.z; break; }
          if ((e.x | e.y) == 0) break;
          s = (s + 1) & p.hash_mask;
        }
      } else {
        const uint8_t* q = tt + (tk & 0xFFFFu) + 1 - len;  // first byte of the token (may lie in an earlier tile)
        uint32_t xb = p.root_base;
        uint4 e = make_uint4(0, 0, 0, 0);
        for (uint32_t d = 0; d < len; d++) {  // the token is in the vocabulary: no checks needed
          e = __ldg(p.trie + (xb ^ (0x100u | __ldg(q + d))));
          xb = e.x >> 9;
        }
        id = e.y & ID_MASK;
      }
      if (p.ids && base + k < p.cap) p.ids[base + k] = id;
      if (p.freq) {
        if (id < EM_HOT) atomicAdd(s_hot + id, 1u);
        else atomicAdd(p.freq + id, 1ull);
      }
    }
    __syncthreads();
  }
  if (p.freq) {
    for (uint32_t i = threadIdx.x; i < EM_HOT && i < p.V; i += EM_BLOCK) {
      const uint32_t c = s_hot[i];
      if (c) atomicAdd(p.freq + i, (unsigned long long)c);
    }
  }
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
  // ln(exp(vmin - vmax) + 1.0): exp and log fused over the only domain this call produces (tgx_libm.h)
  return __dadd_rn(vmax, tgx_softplus_impl(__dadd_rn(vmin, -vmax), reinterpret_cast<const double*>(c_exp_hdr), t.et,
                                           reinterpret_cast<const double*>(c_log_hdr), t.lt));
}

// -----------------------------------------------------------------------------------------
// K4  forward pass: A[e] = fold over tokens ending at e, ascending start, of
//     log_sum_exp(., score + A[start]); first term assigns; nothing ends at e -> 0.0.
//     Lattice::populate_marginal alpha loop, src/lattice.rs:259-272 (per-position form).
// -----------------------------------------------------------------------------------------
template <int G, bool DROP = false>
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
  uint32_t ufirst = u.first, ucount = u.count;
  unit_range(u.counts, u.part, ufirst, ucount);
  const bool has = gidx < ucount;
  const uint32_t unit = has ? u.order[ufirst + gidx] : 0;
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
    s.mcnt[lane] = walk_matches<DROP>(u, text, p0 + lig, n, s, lane, p.dropout, p.drop_key, p.drop_base + start);
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
// STORE_B: only the beta chain, written to Bout (layout of A); the counts are added by fb_contrib_kernel.
template <int G, bool STORE_B = false, bool DROP = false>
__global__ void __launch_bounds__(WPB * 32) fb_backward_kernel(FbParams p, double* Bout = nullptr) {
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
  uint32_t ufirst = u.first, ucount = u.count;
  unit_range(u.counts, u.part, ufirst, ucount);
  bool has = gidx < ucount;
  const uint32_t unit = has ? u.order[ufirst + gidx] : 0;
  if (!STORE_B && has && p.status[unit] != 0) has = false;  // bad z: the reference panics; nothing is added
  const uint32_t n = has ? u.unit_len[unit] : 0;
  const uint64_t start = has ? u.unit_start[unit] : 0;
  const uint8_t* text = u.text + start;
  const double* A = p.A + start + unit;
  double* Bo = Bout + start + unit;
  const double z = (has && !STORE_B) ? A[n] : 0.0;
  if (STORE_B && has && lig == 0) Bo[n] = 0.0;

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
    s.mcnt[lane] = walk_matches<DROP>(u, text, p0 + lig, n, s, lane, p.dropout, p.drop_key, p.drop_base + start);
    const double a_mine = (!STORE_B && has && p0 + lig < n) ? A[p0 + lig] : 0.0;
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
      for (uint32_t k = lig; !STORE_B && k < c; k += G) {
        const double sc = s.mscore[k * ROW_STRIDE + src];
        const uint32_t mp = s.mpack[k * ROW_STRIDE + src];
        uint32_t ts = sl + (mp >> 24);
        if (ts >= W) ts -= W;
        // total = a + score + b - z ; update = total.exp()   (src/lattice.rs:305-307)
        const double total = __dadd_rn(__dadd_rn(__dadd_rn(a, sc), wB[ts]), -z);
        const uint32_t id = mp & ID_MASK;
        acc_add(acc_slot(p, blockIdx.x, id), tgx_exp(total, lt));
      }
      __syncwarp();
      if (pp < n && lig == 0) {
        wB[sl] = b;
        if (STORE_B) Bo[pp] = b;
      }
      __syncwarp();
      sl = sl == 0 ? W - 1 : sl - 1;
    }
  }
}

// -----------------------------------------------------------------------------------------
// K4L / K5L  forward-backward, one LANE per snippet (snippets below the lane threshold, max_token_len <= 16).
//
// The lane-group kernels above advance 32 / G chains per warp instruction (ncu, G = 4: 11 of 32 threads active,
// 194 warp instructions per position); here every lane of a warp is its own snippet, so one warp instruction
// advances up to 32 chains.  The folds are the same operations in the same order (forward: ascending start per end
// position; backward: ascending length per start position), so A, B and every contribution are bit-identical to
// the lane-group kernels'.  The trie walk is fused with the fold: a terminal met at depth d is folded at once.
// Per lane: a 16-slot window of accumulators in shared memory ([slot][lane]: conflict-free), the "seen" flags in
// a register, the next 16 text bytes in two registers refilled one aligned 8-byte word per 8 positions.
// -----------------------------------------------------------------------------------------
constexpr int FL_WARPS = 4;

struct FbLaneParams {
  FbParams f;
  const uint8_t* blob_end;
  double* B;  // [N + U] backward log-probabilities (split form only)
};

__device__ __forceinline__ unsigned long long fl_word(const uint8_t* a, const uint8_t* blob_end) {
  return a < blob_end ? __ldg(reinterpret_cast<const unsigned long long*>(a)) : 0ull;
}

// Both kernels run a FLATTENED walk: one loop iteration = one trie probe of the lane's own (position, depth) state,
// whatever position the other lanes are at, so a lane never waits for the deepest walk of its warp and the fold
// below runs with most lanes active (ncu on the position-synchronous version: 8 of 32 threads per instruction).
// The next probe is issued before the fold of the current terminal, which hides its latency.
__device__ __forceinline__ uint32_t fl_byte(unsigned long long lo, unsigned long long hi, uint32_t d) {
  return (uint32_t)((d < 8 ? lo : hi) >> (8 * (d & 7u))) & 0xFFu;
}

// populate_nodes' dropout (src/model.rs:48-50): the keyed draw of the multi-byte match (start byte, length)
template <bool DROP>
__device__ __forceinline__ bool fb_dropped(const FbParams& p, unsigned long long byte_off, uint32_t l) {
  if (!DROP) return false;
  return l > 1u && !(p.dropout < drop_draw(p.drop_key, p.drop_base + byte_off, l));
}

template <bool DROP>
__device__ __forceinline__ void fb_forward_lane_body(const FbLaneParams& q, uint32_t bid, double* s_win,
                                                     const LibmTabs& lt) {
  const FbParams& p = q.f;
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* acc = s_win + warp * (16 * 32) + lane;  // slot s at acc[s * 32]
  const uint64_t gidx = (uint64_t)bid * (FL_WARPS * 32) + threadIdx.x;
  const bool has = gidx < u.count;
  const uint32_t unit = has ? u.order[u.first + gidx] : 0;
  const uint32_t n = has ? u.unit_len[unit] : 0;
  const uint64_t start = has ? u.unit_start[unit] : 0;
  double* A = p.A + start + unit;
  if (has) A[0] = 0.0;
  if (has && n == 0) p.status[unit] = 7;  // z = 0.0 is not normal (Q11)
  bool active = has && n != 0;
  const uint8_t* tp = u.text + start;
  const uint8_t* base = reinterpret_cast<const uint8_t*>(reinterpret_cast<unsigned long long>(tp) & ~7ull);
  uint32_t sh = (uint32_t)(tp - base);
  unsigned long long w0 = 0, w1 = 0, w2 = 0;
  if (active) {
    w0 = fl_word(base, q.blob_end);
    w1 = fl_word(base + 8, q.blob_end);
    w2 = fl_word(base + 16, q.blob_end);
  }
  unsigned long long lo = sh ? ((w0 >> (8 * sh)) | (w1 << (64 - 8 * sh))) : w0;
  unsigned long long hi = sh ? ((w1 >> (8 * sh)) | (w2 << (64 - 8 * sh))) : w1;
  uint32_t seen = 0, pos = 0, d = 0, limit = min(16u, n), xb = u.root_base;
  double a = 0.0;  // alpha of the nodes that start at pos; 0.0 when nothing ends there (src/lattice.rs:255, Q7)
  uint32_t cw = 0x100u | fl_byte(lo, hi, 0);
  uint4 e = __ldg(u.trie + (xb ^ cw));
  // The vote makes every iteration a convergence point: without it the lanes of a warp drift into separate
  // instruction streams (ncu: 7 of 32 threads active per instruction).
  while (__any_sync(0xFFFFFFFFu, active)) {
    if (active) {
      const bool hit = ((e.x ^ cw) & 0x1FFu) == 0;
      const bool term = hit && (e.y & F_TERM) && !fb_dropped<DROP>(p, start + pos, d + 1u);
      const double y = __dadd_rn(__hiloint2double((int)e.w, (int)e.z), a);  // nodes[lid].score + alpha[lid]
      const uint32_t ts = (pos + d + 1u) & 15u;
      const bool cont = hit && (e.y & F_HASCH) && (d + 1u < limit);
      if (cont) {
        d++;
        xb = e.x >> 9;
      } else {
        pos++;
        d = 0;
        xb = u.root_base;
        if (pos < n) {
          if (++sh == 8) {
            sh = 0;
            base += 8;
            w0 = w1;
            w1 = w2;
            w2 = fl_word(base + 16, q.blob_end);
          }
          lo = sh ? ((w0 >> (8 * sh)) | (w1 << (64 - 8 * sh))) : w0;
          hi = sh ? ((w1 >> (8 * sh)) | (w2 << (64 - 8 * sh))) : w1;
          limit = min(16u, n - pos);
        }
      }
      cw = 0x100u | fl_byte(lo, hi, d);
      e = __ldg(u.trie + (xb ^ cw));  // the next probe flies while the terminal below is folded
      if (term) {
        if ((seen >> ts) & 1u) {
          acc[ts * 32] = log_sum_exp(acc[ts * 32], y, lt);
        } else {  // lid == end_nodes[pos][0] -> init_mode
          acc[ts * 32] = y;
          seen |= 1u << ts;
        }
      }
      if (!cont) {  // pos is the NEXT position now: everything that ends there has been folded
        const uint32_t sl = pos & 15u;
        a = ((seen >> sl) & 1u) ? acc[sl * 32] : 0.0;
        seen &= ~(1u << sl);
        A[pos] = a;
        if (pos == n) {  // a = alpha[eos]
          const double az = fabs(a);
          const bool normal = (az >= 2.2250738585072014e-308) && (az <= 1.7976931348623157e308);  // f64::is_normal
          p.status[unit] = normal ? 0 : 7;
          active = false;
        }
      }
    }
  }
}

// STORE_B: only the beta chain, written to q.B (same layout as A) — it needs neither A nor z, so it runs BESIDE the
// forward kernel and fb_contrib_kernel adds the expected counts afterwards; the longest snippet then costs
// max(forward, backward) instead of their sum.  !STORE_B: the fused form (after the forward kernel).
template <bool STORE_B, bool DROP>
__device__ __forceinline__ void fb_backward_lane_body(const FbLaneParams& q, uint32_t bid, double* s_win,
                                                      const LibmTabs& lt) {
  const FbParams& p = q.f;
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* wB = s_win + warp * (16 * 32) + lane;
  const uint64_t gidx = (uint64_t)bid * (FL_WARPS * 32) + threadIdx.x;
  const bool has = gidx < u.count;
  const uint32_t unit = has ? u.order[u.first + gidx] : 0;
  const uint32_t n = has ? u.unit_len[unit] : 0;
  // bad z: the reference panics; nothing is added
  bool active = has && n != 0 && (STORE_B || p.status[unit] == 0);
  const uint64_t start = has ? u.unit_start[unit] : 0;
  const double* A = p.A + start + unit;
  double* Bout = q.B + start + unit;
  const double z = (active && !STORE_B) ? A[n] : 0.0;
  wB[(n & 15u) * 32] = 0.0;  // beta at the end of the sentence (EOS)
  if (STORE_B && has) Bout[n] = 0.0;
  uint32_t pos = active ? n - 1 : 0;
  const uint8_t* tp = u.text + start + pos;
  const uint8_t* base = reinterpret_cast<const uint8_t*>(reinterpret_cast<unsigned long long>(tp) & ~7ull);
  uint32_t sh = (uint32_t)(tp - base);
  unsigned long long w0 = 0, w1 = 0, w2 = 0;
  if (active) {
    w0 = fl_word(base, q.blob_end);
    w1 = fl_word(base + 8, q.blob_end);
    w2 = fl_word(base + 16, q.blob_end);
  }
  unsigned long long lo = sh ? ((w0 >> (8 * sh)) | (w1 << (64 - 8 * sh))) : w0;
  unsigned long long hi = sh ? ((w1 >> (8 * sh)) | (w2 << (64 - 8 * sh))) : w1;
  double a = (active && !STORE_B) ? A[pos] : 0.0, a_next = (active && !STORE_B && pos) ? A[pos - 1] : 0.0;
  double b = 0.0;  // stays 0.0 when nothing begins at pos (Q7)
  bool first = true;
  uint32_t d = 0, limit = 1, xb = u.root_base;
  uint32_t cw = 0x100u | fl_byte(lo, hi, 0);
  uint4 e = __ldg(u.trie + (xb ^ cw));
  while (__any_sync(0xFFFFFFFFu, active)) {
    if (active) {
      const bool hit = ((e.x ^ cw) & 0x1FFu) == 0;
      const bool term = hit && (e.y & F_TERM) && !fb_dropped<DROP>(p, start + pos, d + 1u);
      const double sc = __hiloint2double((int)e.w, (int)e.z);
      const double bt = wB[((pos + d + 1u) & 15u) * 32];
      const uint32_t id = e.y & ID_MASK;
      const bool cont = hit && (e.y & F_HASCH) && (d + 1u < limit);
      uint32_t nxb = u.root_base, nd = 0;
      if (cont) {
        nd = d + 1;
        nxb = e.x >> 9;
      } else if (pos != 0) {  // first probe of position pos - 1
        if (sh != 0) {
          sh--;
        } else {
          sh = 7;
          base -= 8;
          w2 = w1;
          w1 = w0;
          w0 = fl_word(base, q.blob_end);
        }
        lo = sh ? ((w0 >> (8 * sh)) | (w1 << (64 - 8 * sh))) : w0;
        hi = sh ? ((w1 >> (8 * sh)) | (w2 << (64 - 8 * sh))) : w1;
      }
      cw = 0x100u | fl_byte(lo, hi, nd);
      e = __ldg(u.trie + (nxb ^ cw));  // the next probe flies while the terminal below is folded
      if (term) {  // ascending length = begin_nodes[pos] order
        const double y = __dadd_rn(sc, bt);  // nodes[rid].score + beta[rid]
        b = first ? y : log_sum_exp(b, y, lt);
        first = false;
        if (!STORE_B) {
          // total = a + score + b - z ; update = total.exp()   (src/lattice.rs:305-307)
          const double total = __dadd_rn(__dadd_rn(__dadd_rn(a, sc), bt), -z);
          acc_add(acc_slot(p, bid, id), tgx_exp(total, lt));
        }
      }
      d = nd;
      xb = nxb;
      if (!cont) {
        wB[(pos & 15u) * 32] = b;
        if (STORE_B) Bout[pos] = b;
        if (pos == 0) {
          active = false;
        } else {
          pos--;
          if (!STORE_B) {
            a = a_next;
            a_next = pos ? A[pos - 1] : 0.0;
          }
          b = 0.0;
          first = true;
          limit = min(16u, n - pos);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(FL_WARPS * 32) fb_forward_lane_kernel(FbLaneParams q) {
  __shared__ double s_win[FL_WARPS * 16 * 32];
  __shared__ unsigned long long s_et[256];
  __shared__ double s_lt[256];
  const LibmTabs lt = stage_libm_tables(s_et, s_lt);
  fb_forward_lane_body<false>(q, blockIdx.x, s_win, lt);
}

__global__ void __launch_bounds__(FL_WARPS * 32) fb_backward_lane_kernel(FbLaneParams q) {
  __shared__ double s_win[FL_WARPS * 16 * 32];
  __shared__ unsigned long long s_et[256];
  __shared__ double s_lt[256];
  const LibmTabs lt = stage_libm_tables(s_et, s_lt);
  fb_backward_lane_body<false, false>(q, blockIdx.x, s_win, lt);
}

// Split form: even blocks run the forward chains of 128 snippets, odd blocks the beta chains of the same snippets,
// so the block scheduler starts the longest snippets of BOTH directions first (two kernels on two streams do not
// interleave: the second kernel's blocks wait for the first kernel's to be dispatched).
// DROP: with the keyed dropout draw (the reference's default `--dropout 0.01`, src/prune.rs:87): a terminal that is
// dropped is simply not a terminal for the folds; the walk goes on through it.
template <bool DROP>
__global__ void __launch_bounds__(FL_WARPS * 32) fb_split_lane_kernel(FbLaneParams q) {
  __shared__ double s_win[FL_WARPS * 16 * 32];
  __shared__ unsigned long long s_et[256];
  __shared__ double s_lt[256];
  const LibmTabs lt = stage_libm_tables(s_et, s_lt);
  if (blockIdx.x & 1u) fb_backward_lane_body<true, DROP>(q, blockIdx.x >> 1, s_win, lt);
  else fb_forward_lane_body<DROP>(q, blockIdx.x >> 1, s_win, lt);
}

// Expected counts from stored alpha and beta (split form): one warp per snippet, a lane per start position.
// exp(alpha[pos] + score + beta[pos + len] - z) per matched token, in the reference's operation order
// (src/lattice.rs:295-309), so every contribution equals the fused kernels' bit for bit.
constexpr int FC_WARPS = 8;

// (Measured alternative: the counts of the ~1800 smallest ids in shared memory, flushed once per block of a persistent
//  grid — 133 ms per GB against 69: every thread of a block then hammers the same few shared-memory words of the
//  hottest tokens, where the global accumulators have 64 replicas.)
template <bool DROP>
__global__ void __launch_bounds__(FC_WARPS * 32) fb_contrib_kernel(FbLaneParams q) {
  __shared__ unsigned long long s_et[256];
  __shared__ double s_lt[256];
  const FbParams& p = q.f;
  const UnitParams& u = p.u;
  const LibmTabs lt = stage_libm_tables(s_et, s_lt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t widx = (uint64_t)blockIdx.x * FC_WARPS + warp;
  if (widx >= u.count) return;
  const uint32_t unit = u.order[u.first + widx];
  if (p.status[unit] != 0) return;  // bad z: the reference panics; nothing is added
  const uint32_t n = u.unit_len[unit];
  const uint64_t start = u.unit_start[unit];
  const double* A = p.A + start + unit;
  const double* B = q.B + start + unit;
  const uint8_t* text = u.text + start;
  const double z = A[n];
  for (uint32_t pos = lane; pos < n; pos += 32) {
    const double a = A[pos];
    const uint32_t limit = min(16u, n - pos);
    uint32_t xb = u.root_base;
    for (uint32_t d = 0; d < limit; d++) {
      const uint32_t cw = 0x100u | __ldg(text + pos + d);
      const uint4 e = __ldg(u.trie + (xb ^ cw));
      if ((e.x ^ cw) & 0x1FFu) break;
      if ((e.y & F_TERM) && !fb_dropped<DROP>(p, start + pos, d + 1u)) {
        const double sc = __hiloint2double((int)e.w, (int)e.z);
        const double total = __dadd_rn(__dadd_rn(__dadd_rn(a, sc), B[pos + d + 1]), -z);
        const uint32_t id = e.y & ID_MASK;
        acc_add(acc_slot(p, blockIdx.x, id), tgx_exp(total, lt));
      }
      if (!(e.y & F_HASCH)) break;
      xb = e.x >> 9;
    }
  }
}

}  // namespace tgxk
