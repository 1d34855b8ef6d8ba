// Match stream + row consumer: the default Viterbi forward pass (max_token_len <= 16).
//
// Model::encode's inner loop (src/model.rs:83-110) interleaves two things: common_prefix_search from every
// position (src/trie.rs:51-63) — independent of the dp, embarrassingly parallel — and the ordered relax chain.
// Here they are two kernels coupled through a 4-byte record per position in HBM:
//
//   match_kernel        one thread per start position walks the 8-byte double-array (trie_build.h: slots8; the leading
//                       slots staged in shared memory, the rest from L2), no barriers, no dp, and writes
//                       rec[p] = (L - 1) << 28 | row offset   of the DEEPEST token that starts at p.
//   viterbi_rows_kernel the relax chain.  The row of that token (trie_build.h: rows) lists, dense by length, the score of
//                       every token on the trie path down to it, i.e. everything the reference's iterator yields at p;
//                       a half-warp owns one sample, lane g the dp cell of the positions = g (mod 16), as in the
//                       pair-CTA kernel — but the candidate scores come from ONE coalesced read of the row (hot rows
//                       from shared memory, the rest from L1/L2) instead of a dense f64 table that producer warps park
//                       in shared memory (136 B per buffered position: what capped that kernel at 10-12 chains per SM).
//
// The evaluation order is the reference's (ascending start, strict '>', f64 sums built left to right), so the back
// lengths — the only output — are bit-identical to the pair kernel's.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "tgx_kernels.cuh"

namespace tgxk {

constexpr uint32_t REC_NOMATCH = 15u << 28;  // row 0: empty header, 16 x -inf
constexpr uint32_t REC_OFF = 0x0FFFFFFFu;

// -----------------------------------------------------------------------------------------
// K2a  match_kernel: TrieIterator::next (src/trie.rs:51-63) from every byte of the blob.
// -----------------------------------------------------------------------------------------
struct MatchParams {
  const uint8_t* text;      // blob (after the processors)
  const uint8_t* blob_end;  // text + N
  unsigned long long N;
  const uint2* trie8;       // tgx::DoubleArray::slots8
  uint32_t root_base;
  uint32_t staged;          // leading slots staged in shared memory
  uint32_t* rec;            // [N + 64]
  unsigned long long slice; // positions per CTA (a multiple of blockDim.x * ILP)
  const uint2* pair2;       // match2_kernel: tgx::DoubleArray::pair2 (the state of a walk after its two first bytes)
  const uint8_t* skip;      // match2_kernel: [ceil(N / 128)] != 0 = nobody reads the records of these 128 positions
                            // (they lie inside a sample that the pair-CTA kernel takes): write row 0, do not walk
};

// skip[t] = 1 for every 128-position block that lies wholly inside one of the first n_long units of the
// length-descending order (the samples the pair-CTA kernel takes in forward pass 3): one warp per unit
__global__ void mark_skip_kernel(const uint64_t* __restrict__ unit_start, const uint32_t* __restrict__ unit_len,
                                 const uint32_t* __restrict__ order, uint32_t n_long, uint8_t* __restrict__ skip) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_long) return;
  const uint32_t unit = order[w];
  const unsigned long long a = unit_start[unit], b = a + unit_len[unit];
  const unsigned long long t0 = (a + 127) >> 7, t1 = b >> 7;  // blocks [t0, t1) are inside [a, b)
  for (unsigned long long t = t0 + lane; t < t1; t += 32) skip[t] = 1;
}

// A thread owns ILP consecutive start positions and walks them side by side, one trie level per step: ILP independent
// probes in flight per thread.  Every probe is two predicated loads — shared memory for the staged prefix of the
// array (the hottest nodes: trie_build.cpp lays the slots out by descending walk frequency, 87 % of the probes of the
// bench corpus land in the first 64 KB), L1/L2 for the rest — and a finished walk keeps probing slot 0, so the step
// has no branch.  The window slides one byte per level, so walk i always reads byte i of it (a rolled loop: the
// 16-fold unrolled form spilled ~100 bytes per walk).
// A walk may run past the end of a sample (into the next sample's bytes): such a match lands on a dp cell beyond
// the sample, which the consumer never reads; the E-step consumers cut it at the snippet's end.
__device__ __forceinline__ uint2 mk_probe(uint32_t t, uint32_t staged, uint32_t s_base, const uint2* trie8) {
  uint2 e;
  asm volatile("{\n\t"
      ".reg .pred ps;\n\t"
      ".reg .u32 sa;\n\t"
      ".reg .u64 ga;\n\t"
      "setp.lt.u32 ps, %2, %3;\n\t"
      "mad.lo.u32 sa, %2, 8, %4;\n\t"
      "mad.wide.u32 ga, %2, 8, %5;\n\t"
      "@ps ld.shared.v2.u32 {%0, %1}, [sa];\n\t"
      "@!ps ld.global.nc.v2.u32 {%0, %1}, [ga];\n\t"
      "}"
      : "=r"(e.x), "=r"(e.y)
      : "r"(t), "r"(staged), "r"(s_base), "l"(trie8));
  return e;
}

constexpr int mk_max_threads(int ilp) { return ilp <= 4 ? 1024 : 768; }  // registers: 64 / 85

template <int ILP>
__global__ void __launch_bounds__(mk_max_threads(ILP), 1) match_kernel(MatchParams p) {
  static_assert(ILP == 4, "positions per thread (1, 2 and 8 were measured: the time does not move)");
  extern __shared__ __align__(16) unsigned char smem[];
  {
    uint2* s_trie = reinterpret_cast<uint2*>(smem);
    for (uint32_t i = threadIdx.x; i < p.staged; i += blockDim.x) s_trie[i] = __ldg(p.trie8 + i);
  }
  __syncthreads();
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem);
  // CTA b owns the positions [b * slice, (b + 1) * slice): more CTAs than SMs, handed out as SMs come free (the
  // forward pass of the longest samples may hold some of them for the whole of this kernel)
  const unsigned long long stride = (unsigned long long)blockDim.x * ILP;
  unsigned long long pos = (unsigned long long)blockIdx.x * p.slice + (unsigned long long)threadIdx.x * ILP;
  const unsigned long long lim = min(p.N, ((unsigned long long)blockIdx.x + 1) * p.slice);
  if (blockIdx.x == 0 && threadIdx.x < 64) p.rec[p.N + threadIdx.x] = 0u;  // padding the consumers may read: row 0
  unsigned long long w[3] = {0, 0, 0};
  uint32_t sh = 0;
  if (pos < lim) load_window(p.text + pos, p.blob_end, w, sh);
  while (pos < lim) {
    // 24 bytes from `pos` on (the last sh of them read as zero: the walks need 16 + ILP - 1 <= 23)
    const unsigned long long a0 = sh ? ((w[0] >> (8 * sh)) | (w[1] << (64 - 8 * sh))) : w[0];
    const unsigned long long a1 = sh ? ((w[1] >> (8 * sh)) | (w[2] << (64 - 8 * sh))) : w[1];
    const unsigned long long a2 = sh ? (w[2] >> (8 * sh)) : w[2];
    const unsigned long long npos = pos + stride;
    if (npos < lim) load_window(p.text + npos, p.blob_end, w, sh);  // the next window flies during these walks
    uint32_t xb[ILP], best[ILP];
    bool go[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      xb[i] = p.root_base;
      best[i] = REC_NOMATCH;
      go[i] = true;
    }
    uint32_t b0 = (uint32_t)a0, b1 = (uint32_t)(a0 >> 32), b2 = (uint32_t)a1, b3 = (uint32_t)(a1 >> 32),
             b4 = (uint32_t)a2, b5 = (uint32_t)(a2 >> 32);
#pragma unroll 1
    for (uint32_t d = 0; d < 16u; d++) {
      uint2 e[ILP];
      uint32_t cw[ILP];
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        cw[i] = __byte_perm(i < 4 ? b0 : b1, 1u, 0x5540 + (i & 3));  // 0x100 | byte i
        e[i] = mk_probe(go[i] ? (xb[i] ^ cw[i]) : 0u, p.staged, s_base, p.trie8);
      }
      bool any = false;
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        const bool hit = go[i] && ((e[i].x ^ cw[i]) & 0x1FFu) == 0u;
        if (hit && (e[i].y & tgx::SLOT8_TERM)) best[i] = (d << 28) | (e[i].y & tgx::SLOT8_OFF_MASK);
        go[i] = hit && (e[i].y & tgx::SLOT8_HASCH);
        xb[i] = e[i].x >> 9;
        any |= go[i];
      }
      if (!any) break;
      b0 = __funnelshift_r(b0, b1, 8);
      b1 = __funnelshift_r(b1, b2, 8);
      b2 = __funnelshift_r(b2, b3, 8);
      b3 = __funnelshift_r(b3, b4, 8);
      b4 = __funnelshift_r(b4, b5, 8);
      b5 >>= 8;
    }
    if (ILP == 4 && pos + 4 <= p.N && (reinterpret_cast<unsigned long long>(p.rec) & 15ull) == 0) {
      *reinterpret_cast<uint4*>(p.rec + pos) = make_uint4(best[0], best[1 % ILP], best[2 % ILP], best[3 % ILP]);
    } else if (ILP == 8 && pos + 8 <= p.N && (reinterpret_cast<unsigned long long>(p.rec) & 15ull) == 0) {
      *reinterpret_cast<uint4*>(p.rec + pos) = make_uint4(best[0], best[1 % ILP], best[2 % ILP], best[3 % ILP]);
      *reinterpret_cast<uint4*>(p.rec + pos + 4) = make_uint4(best[4 % ILP], best[5 % ILP], best[6 % ILP], best[7 % ILP]);
    } else {
#pragma unroll
      for (int i = 0; i < ILP; i++)
        if (pos + i < p.N) p.rec[pos + i] = best[i];
    }
    pos = npos;
  }
}

// match2_kernel: the same records, with the walks COMPACTED twice inside their warp.
//
// match_kernel keeps a warp in its walk loop until the deepest of its 128 walks is done: 13 trips on average (of 16)
// for walks that need 5 probes on average (tools/analyse_walk_depth.py: 43 % of the starts are alive after four probes,
// 15 % after eight) — ncu: 16 of 32 threads per instruction.  Here a warp walks its 128 starts four levels deep (four
// per lane, side by side, no exit test), writes the survivors' states (start, base, deepest token so far) to a queue in
// shared memory, and walks those four levels further with ONE walk per lane (two trips of the queue on average), queues
// the survivors again and finishes them (a single trip, exit when the last one is done).  No barrier: the queues are
// the warp's own.
constexpr int MK2_THREADS = 768;
constexpr int MK2_Q = 128;  // entries per queue: the warp's starts of one trip

__device__ __forceinline__ uint32_t mk_byte_cw(unsigned long long a, int j) {  // 0x100 | byte j of a (0 <= j < 8)
  return __byte_perm(j < 4 ? (uint32_t)a : (uint32_t)(a >> 32), 1u, 0x5540 + (j & 3));
}

__global__ void __launch_bounds__(MK2_THREADS, 1) match2_kernel(MatchParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  {
    uint2* s_trie = reinterpret_cast<uint2*>(smem);
    for (uint32_t i = threadIdx.x; i < p.staged; i += blockDim.x) s_trie[i] = __ldg(p.trie8 + i);
  }
  __syncthreads();
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2* q1 = reinterpret_cast<uint2*>(smem + (size_t)p.staged * 8) + (size_t)warp * 2 * MK2_Q;
  uint2* q2 = q1 + MK2_Q;
  const uint32_t lt = (1u << lane) - 1u;
  const unsigned long long stride = (unsigned long long)blockDim.x * 4;
  unsigned long long pos = (unsigned long long)blockIdx.x * p.slice + (unsigned long long)threadIdx.x * 4;
  const unsigned long long lim = min(p.N, ((unsigned long long)blockIdx.x + 1) * p.slice);
  if (blockIdx.x == 0 && threadIdx.x < 64) p.rec[p.N + threadIdx.x] = 0u;  // padding the consumers may read: row 0
  const bool aligned = (reinterpret_cast<unsigned long long>(p.rec) & 15ull) == 0;
  unsigned long long w[3] = {0, 0, 0};
  uint32_t sh = 0;
  if (pos < lim) load_window(p.text + pos, p.blob_end, w, sh);
  // (a warp's 128 starts lie in one slice: the slice is a multiple of blockDim.x * 4)
  while (pos - (unsigned long long)lane * 4 < lim) {
    const unsigned long long wpos = pos - (unsigned long long)lane * 4;  // the warp's first start (a multiple of 128)
    const unsigned long long a0 = window_bytes(w, sh, 0);               // 8 bytes from `pos` on (zero beyond the blob)
    const unsigned long long npos = pos + stride;
    if (npos < lim) load_window(p.text + npos, p.blob_end, w, sh);  // the next window flies during these walks
    if (p.skip && __ldg(p.skip + (wpos >> 7))) {  // (warp-uniform)
      if (pos + 4 <= p.N && aligned) {
        *reinterpret_cast<uint4*>(p.rec + pos) = make_uint4(0u, 0u, 0u, 0u);
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (pos + i < p.N) p.rec[pos + i] = 0u;
      }
      pos = npos;
      continue;
    }
    // ---- levels 1..4: four walks per lane, side by side; the two first levels come from the pair table
    uint32_t xb[4], best[4];
    bool go[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint2 t2 = __ldg(p.pair2 + (uint32_t)((a0 >> (8 * i)) & 0xFFFFull));
      xb[i] = t2.x & 0x7FFFFFFFu;
      best[i] = t2.y;
      go[i] = (t2.x >> 31) != 0u && pos + i < lim;
    }
#pragma unroll
    for (int d = 2; d < 4; d++) {
      uint2 e[4];
      uint32_t cw[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        cw[i] = mk_byte_cw(a0, i + d);
        e[i] = mk_probe(go[i] ? (xb[i] ^ cw[i]) : 0u, p.staged, s_base, p.trie8);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const bool hit = go[i] && ((e[i].x ^ cw[i]) & 0x1FFu) == 0u;
        if (hit && (e[i].y & tgx::SLOT8_TERM)) best[i] = ((uint32_t)d << 28) | (e[i].y & tgx::SLOT8_OFF_MASK);
        go[i] = hit && (e[i].y & tgx::SLOT8_HASCH);
        xb[i] = e[i].x >> 9;
      }
    }
    if (pos + 4 <= p.N && aligned) {
      *reinterpret_cast<uint4*>(p.rec + pos) = make_uint4(best[0], best[1], best[2], best[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (pos + i < p.N) p.rec[pos + i] = best[i];
    }
    // ---- the walks that go on: into the warp's queue (their records are written again when they end)
    uint32_t c1 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t b = __ballot_sync(0xFFFFFFFFu, go[i]);
      if (go[i]) q1[c1 + __popc(b & lt)] = make_uint2(xb[i] | (uint32_t)(lane * 4 + i) << 24, best[i]);
      c1 += __popc(b);
    }
    __syncwarp();
    // ---- levels 5..8: one walk per lane
    uint32_t c2 = 0;
    for (uint32_t e0 = 0; e0 < c1; e0 += 32) {
      const bool has = e0 + lane < c1;
      const uint2 ent = has ? q1[e0 + lane] : make_uint2(0u, 0u);
      const uint32_t rel = ent.x >> 24;
      uint32_t x = ent.x & 0xFFFFFFu, bst = ent.y;
      unsigned long long v[3] = {0, 0, 0};
      uint32_t vs = 0;
      if (has) load_window(p.text + wpos + rel + 4, p.blob_end, v, vs);
      const unsigned long long a = window_bytes(v, vs, 0);
      bool g = has;
#pragma unroll
      for (int d = 0; d < 4; d++) {
        const uint32_t cw = mk_byte_cw(a, d);
        const uint2 e = mk_probe(g ? (x ^ cw) : 0u, p.staged, s_base, p.trie8);
        const bool hit = g && ((e.x ^ cw) & 0x1FFu) == 0u;
        if (hit && (e.y & tgx::SLOT8_TERM)) bst = ((uint32_t)(d + 4) << 28) | (e.y & tgx::SLOT8_OFF_MASK);
        g = hit && (e.y & tgx::SLOT8_HASCH);
        x = e.x >> 9;
      }
      const uint32_t b = __ballot_sync(0xFFFFFFFFu, g);
      if (g) q2[c2 + __popc(b & lt)] = make_uint2(x | rel << 24, bst);
      else if (has) p.rec[wpos + rel] = bst;
      c2 += __popc(b);
    }
    __syncwarp();
    // ---- levels 9..16
    for (uint32_t e0 = 0; e0 < c2; e0 += 32) {
      const bool has = e0 + lane < c2;
      const uint2 ent = has ? q2[e0 + lane] : make_uint2(0u, 0u);
      const uint32_t rel = ent.x >> 24;
      uint32_t x = ent.x & 0xFFFFFFu, bst = ent.y;
      unsigned long long v[3] = {0, 0, 0};
      uint32_t vs = 0;
      if (has) load_window(p.text + wpos + rel + 8, p.blob_end, v, vs);
      const unsigned long long a = window_bytes(v, vs, 0);
      bool g = has;
#pragma unroll 1
      for (int d = 0; d < 8; d++) {
        const uint32_t cw = 0x100u | (uint32_t)((a >> (8 * d)) & 0xFFull);
        const uint2 e = mk_probe(g ? (x ^ cw) : 0u, p.staged, s_base, p.trie8);
        const bool hit = g && ((e.x ^ cw) & 0x1FFu) == 0u;
        if (hit && (e.y & tgx::SLOT8_TERM)) bst = ((uint32_t)(d + 8) << 28) | (e.y & tgx::SLOT8_OFF_MASK);
        g = hit && (e.y & tgx::SLOT8_HASCH);
        x = e.x >> 9;
        if (!__any_sync(0xFFFFFFFFu, g)) break;
      }
      if (has) p.rec[wpos + rel] = bst;
    }
    __syncwarp();  // (the queues are written again in the next trip)
    pos = npos;
  }
}

// -----------------------------------------------------------------------------------------
// K2b  viterbi_rows_kernel: the relax chain of Model::encode (src/model.rs:83-110) over the match stream.
// -----------------------------------------------------------------------------------------
struct RowsParams {
  UnitParams u;         // unit_start / unit_len / order / counts+part (text and trie unused)
  const uint32_t* rec;  // [N] match stream
  const double* rows;   // row table
  uint32_t hot16;       // leading 16-byte units of the row table staged in shared memory
  uint8_t* bp;          // [N] back length per end position (0 = unreachable)
  unsigned int* counter;
};

struct RowsUnit {
  unsigned long long start;
  uint32_t n, ntiles, tile;
  int32_t unit;  // < 0: nothing to do
};

// rows[idx] from the staged prefix of the row table (shared memory) or from L1/L2: two predicated loads, no branch
// (as C++ the select compiled to divergent branches in front of every step of the chain).
__device__ __forceinline__ double ld_row(uint32_t idx, uint32_t hot_dbl, uint32_t s_base, const double* rows) {
  double v;
  asm("{\n\t"
      ".reg .pred ph;\n\t"
      ".reg .u32 sa;\n\t"
      ".reg .u64 ga;\n\t"
      "setp.lt.u32 ph, %1, %2;\n\t"
      "mad.lo.u32 sa, %1, 8, %3;\n\t"
      "mad.wide.u32 ga, %1, 8, %4;\n\t"
      "@ph ld.shared.f64 %0, [sa];\n\t"
      "@!ph ld.global.nc.f64 %0, [ga];\n\t"
      "}"
      : "=d"(v)
      : "r"(idx), "r"(hot_dbl), "r"(s_base), "l"(rows));
  return v;
}

// One 32-position tile of both halves of the warp.  r0 / r1: this lane's records of positions g and 16 + g of the
// tile.  Lane g's candidate at step j is the token of length ((g - j - 1) & 15) + 1 starting at position j — the one
// that lands on its cell (a length beyond the row's reads row 0: -inf); "unreached" is best == -inf exactly as in
// pair_consume (tgx_kernels.cuh).
__device__ __forceinline__ void rows_consume(const double* __restrict__ rows, uint32_t s_base,
                                             uint32_t hot_dbl, uint32_t r0, uint32_t r1, int g, double& best,
                                             uint32_t& ps, uint32_t& len0, uint32_t& len1) {
  auto fetch = [&](int j) -> double {
    const uint32_t rs = __shfl_sync(0xFFFFFFFFu, j < 16 ? r0 : r1, j & 15, 16);  // record of position j
    const uint32_t li = (uint32_t)(g - j - 1) & 15u;                            // candidate length - 1
    const uint32_t idx = 1u + ((li <= (rs >> 28)) ? (rs & REC_OFF) * 2u + li : li);  // (a row starts with its header)
    return ld_row(idx, hot_dbl, s_base, rows);  // (element-wise: a row may straddle the staged prefix)
  };
  double sc[4];
  sc[0] = fetch(0);
  sc[1] = fetch(1);
  sc[2] = fetch(2);
  uint32_t sv_hi = 0, sv_ps = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    if (j + 3 < 32) sc[(j + 3) & 3] = fetch(j + 3);
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

__global__ void __launch_bounds__(512, 1) viterbi_rows_kernel(RowsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.rows);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < p.hot16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31;
  const int g = lane & 15;
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
  uint32_t ufirst = u.first, ucount = u.count;
  unit_range(u.counts, u.part, ufirst, ucount);

  RowsUnit cur;
  cur.unit = -1; cur.start = 0; cur.n = 0; cur.ntiles = 0; cur.tile = 0;
  // the half's leader takes the next sample of the length-descending order (LPT); everybody gets a copy
  // (called by all 32 lanes: the shuffle is a full-warp one)
  auto fetch_unit = [&](bool need) {
    uint32_t idx = 0;
    if (need && g == 0) idx = atomicAdd(p.counter, 1u);
    idx = __shfl_sync(0xFFFFFFFFu, idx, 0, 16);
    if (!need) return;
    if (idx < ucount) {
      cur.unit = (int32_t)u.order[ufirst + idx];
      cur.n = u.unit_len[cur.unit];
      cur.start = u.unit_start[cur.unit];
      cur.tile = 0;
      cur.ntiles = cur.n / 32 + 1;  // positions 0..n
    } else {
      cur.unit = -1;
    }
  };
  auto load_rec = [&](uint32_t tile, uint32_t& a, uint32_t& b) {
    const uint32_t q0 = tile * 32 + g, q1 = q0 + 16;
    a = (cur.unit >= 0 && q0 < cur.n) ? __ldg(p.rec + cur.start + q0) : REC_NOMATCH;
    b = (cur.unit >= 0 && q1 < cur.n) ? __ldg(p.rec + cur.start + q1) : REC_NOMATCH;
  };
  auto pull_rows = [&](uint32_t a, uint32_t b) {  // cold rows of a tile that is about to be consumed: into L1
    if ((a & REC_OFF) >= p.hot16) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.rows + (size_t)(a & REC_OFF) * 2));
    if ((b & REC_OFF) >= p.hot16) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.rows + (size_t)(b & REC_OFF) * 2));
  };
  fetch_unit(true);
  double best = ninf;
  uint32_t ps = 0;
  // records of the current tile, the next one (its cold rows are pulled into L1 while this tile is consumed) and the
  // one after (in flight)
  uint32_t r0, r1, n0, n1, m0 = REC_NOMATCH, m1 = REC_NOMATCH;
  load_rec(0, r0, r1);
  load_rec(1, n0, n1);
  while (__any_sync(0xFFFFFFFFu, cur.unit >= 0)) {
    const bool act = cur.unit >= 0;
    if (act && cur.tile == 0) {  // dp[0] = { score 0.0, start Some(0) }  (src/model.rs:72-81); the rest unreached
      best = (g == 0) ? 0.0 : ninf;
      ps = 0;
    }
    const bool more = act && cur.tile + 1 < cur.ntiles;
    if (more) {
      load_rec(cur.tile + 2, m0, m1);
      pull_rows(n0, n1);
    }
    uint32_t len0, len1;
    rows_consume(p.rows, s_base, p.hot16 * 2u, r0, r1, g, best, ps, len0, len1);  // both halves always run it (full-warp shuffles)
    if (act) {
      const uint32_t e0 = cur.tile * 32 + g, e1 = e0 + 16;
      if (e0 >= 1 && e0 <= cur.n) p.bp[cur.start + e0 - 1] = (uint8_t)len0;
      if (e1 <= cur.n) p.bp[cur.start + e1 - 1] = (uint8_t)len1;
    }
    if (more) {
      cur.tile++;
      r0 = n0; r1 = n1;
      n0 = m0; n1 = m1;
    }
    fetch_unit(act && !more);
    if (!more) {
      load_rec(0, r0, r1);
      load_rec(1, n0, n1);
    }
  }
}

}  // namespace tgxk
