// Segment-parallel exact Viterbi (forward algo 4): Model::encode (src/model.rs:59-129) without the per-sample chain.
//
// A boundary c of a sample is a CUT when no vocabulary token matched anywhere starts before c and ends after c.
// Every segmentation passes through every cut, so a sample is a chain of small independent lattices ("segments",
// 3.7 bytes on average on the bench corpus) coupled only through the exact f64 value dp[cut].  Inside a segment the
// reference's decisions (src/model.rs:98-101: cand = dp[s] + score, strict '>', ascending start) compare sums that
// differ from the segment-local sums  L[s] + score  (local dp started at 0.0) by at most (depth + 1) roundings of
// half an ulp of the largest |dp| of the sample.  Hence:
//   P1  seg_solve_kernel   position-parallel over the whole blob: walk the trie from every byte, find the cuts
//       (prefix max of the reach), solve every segment locally, one lane per segment, and accept the local decisions
//       when EVERY comparison made in the segment has a margin above T = 256 ulp(n_max * max|score|) — then they are
//       the reference's decisions whatever dp[cut] is.  Accepted segments get their token-end marks and ids at once
//       (ids through the token hash, trie_build.h).  Anything else — an exact or near tie, an unreachable end, a
//       segment longer than SG_MAXSEG — is left as a "hard" marker 0xFE on the segment's first byte.
//   P2  seg_chain_kernel   one warp per sample, longest first: the exact chain dp[cut] is ONE f64 add per best-path
//       token (fl(dp[s] + score), left to right, exactly the value the reference holds at that position), and each
//       hard segment is solved exactly in place, in the reference's order, from the exact dp at its cut.
// Output: mark[] (token length at every token END of the best path), ids_at[] (token id at the same byte), token
// counts and NoPath status per sample — what backtrack_*_kernel + the re-walk of emit_kernel produce for algo 0.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "tgx_kernels.cuh"
#include "trie_build.h"

namespace tgxk {

constexpr int SG_TP = 512;                     // segment starts (boundaries) a tile owns
constexpr int SG_MAXSEG = 48;                  // longest segment solved by P1
constexpr int SG_LH = 16;                      // left halo: starts whose reach decides the first cuts of the tile
constexpr int SG_ROWS = SG_TP + SG_MAXSEG;     // score table rows: starts t0 .. t0 + ROWS - 1
constexpr int SG_W = SG_LH + SG_ROWS;          // walked starts t0 - 16 .. t0 + ROWS - 1  (= threads per CTA)
constexpr int SG_THREADS = SG_W;               // 576 = 18 warps
constexpr int SG_WARPS = SG_THREADS / 32;
constexpr int SG_TW = SG_W + 32;               // text window: bytes t0 - 32 .. t0 + ROWS + 16
constexpr uint8_t SG_HARD = 0xFE;
static_assert(SG_THREADS % 32 == 0, "whole warps");

struct SegParams {
  const uint8_t* text;
  unsigned long long N;
  const uint32_t* bitmap;  // bit b set: a sample starts at byte b (b > 0)
  unsigned long long bitmap_words;
  const uint4* trie;
  uint32_t root_base, hot_slots, max_len;
  const uint4* hash;  // token hash (trie_build.h): x,y = key, z = id
  uint32_t hash_mask;
  unsigned long long hash_seed;
  const uint32_t* sorted_len;  // [U] sample lengths, descending
  double wmax;                 // max |score| of the vocabulary
  uint8_t* mark;
  uint32_t* ids_at;
  unsigned long long n_tiles;
  unsigned long long* dbg;  // developer counters (tools/probe.py); null in production
};

__host__ __device__ inline size_t seg_smem_bytes(uint32_t hot_slots) {
  size_t b = (size_t)hot_slots * 16;
  b += (size_t)16 * SG_ROWS * 8;   // score table [len - 1][row]
  b += (size_t)(SG_W + 8) * 8;     // local dp
  b += (size_t)SG_W * 2;           // length masks
  b += (size_t)SG_TP * 2;          // segment starts
  b += SG_TW;                      // text
  b += SG_W + 16;                  // back lengths
  b += SG_W + 16;                  // marks
  b += 24 * 4 * 2 + 40 * 4;        // sample-start bits, cut bits, scan scratch
  return (b + 15) & ~(size_t)15;
}

// margin below which a local comparison is not trusted: 256 ulp of the largest |dp| any sample of the batch can
// reach (|dp| <= n * max|score|).  A candidate is off its local value by <= (depth + 1) roundings of ulp/2 in the
// reference's chain and as many in the local one, two candidates per comparison: <= 2 (SG_MAXSEG + 1) ulp < 256 ulp.
__device__ __forceinline__ double seg_margin(uint32_t longest, double wmax) {
  const double B = (double)max(longest, 64u) * wmax;
  const long long bits = __double_as_longlong(B);
  long long e = ((bits >> 52) & 0x7FF) - 44;  // biased exponent of 2^(ilogb(B) - 52 + 8)
  if (e < 1) return 0.0;                      // (all scores ~0: only exact ties are hard)
  if (e > 2046) e = 2046;
  return __longlong_as_double(e << 52);
}

__device__ __forceinline__ uint32_t hash_lookup(const uint4* __restrict__ hash, uint32_t mask, unsigned long long seed,
                                                unsigned long long lo, unsigned long long hi, uint32_t len) {
  const unsigned long long key = tgx::token_key(lo, hi, len, seed);
  uint32_t s = tgx::token_key_slot(key, mask);
  for (int it = 0; it < 256; it++) {
    const uint4 e = __ldg(hash + s);
    if ((((unsigned long long)e.y << 32) | e.x) == key) return e.z;
    if ((e.x | e.y) == 0) break;
    s = (s + 1) & mask;
  }
  return NONE;  // not a vocabulary token (cannot happen for tokens the trie matched)
}

template <int HOT>
__global__ void __launch_bounds__(SG_THREADS, 2) seg_solve_kernel(SegParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint4* s_hot = reinterpret_cast<uint4*>(smem);
  double* s_tab = reinterpret_cast<double*>(smem + (size_t)p.hot_slots * 16);
  double* s_d = s_tab + 16 * SG_ROWS;
  uint16_t* s_mask = reinterpret_cast<uint16_t*>(s_d + SG_W + 8);
  uint16_t* s_seg = s_mask + SG_W;
  uint8_t* s_text = reinterpret_cast<uint8_t*>(s_seg + SG_TP);
  uint8_t* s_bp = s_text + SG_TW;
  uint8_t* s_mark = s_bp + SG_W + 16;
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_mark + SG_W + 16);
  uint32_t* s_cut = s_bits + 24;
  uint32_t* s_ws = s_cut + 24;  // [0,18) warp maxima, [18,36) warp segment counts, [37] end of the tile's output

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
  if (HOT > 0)
    for (uint32_t i = tid; i < p.hot_slots; i += SG_THREADS) s_hot[i] = __ldg(p.trie + i);
  const double T = seg_margin(__ldg(p.sorted_len), p.wmax);
  const bool text_aligned = (reinterpret_cast<unsigned long long>(p.text) & 3ull) == 0;
  const long long N = (long long)p.N;

  for (unsigned long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const long long t0 = (long long)tile * SG_TP;
    const long long tb = t0 - 32;  // blob byte of s_text[0]
    if (tid < SG_TW / 4) {
      const long long g = tb + 4 * tid;
      uint32_t w = 0;
      if (g >= 0 && g + 4 <= N && text_aligned) {
        w = __ldg(reinterpret_cast<const uint32_t*>(p.text + g));
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (g + k >= 0 && g + k < N) w |= (uint32_t)__ldg(p.text + g + k) << (8 * k);
      }
      reinterpret_cast<uint32_t*>(s_text)[tid] = w;
    }
    if (tid < 24) {  // sample-start bits of boundaries tb + 32 k ..
      const long long wi = (long long)tile * (SG_TP / 32) - 1 + tid;
      s_bits[tid] = (wi >= 0 && (unsigned long long)wi < p.bitmap_words) ? __ldg(p.bitmap + wi) : 0u;
      s_cut[tid] = 0;
    }
    s_mark[tid] = 0;
    if (tid < 16) s_mark[SG_W + tid] = 0;
    __syncthreads();

    // ---- walk: thread j owns start byte i = t0 - 16 + j  (TrieIterator::next, src/trie.rs:51-63)
    const int j = tid;
    const long long i = t0 - SG_LH + j;
    uint32_t mask = 0;
    if (i >= 0 && i < N) {
      const uint32_t r = (uint32_t)j + 16u;  // boundary i relative to tb
      const unsigned long long wb =
          ((unsigned long long)s_bits[r >> 5] | ((unsigned long long)s_bits[(r >> 5) + 1] << 32)) >> (r & 31);
      const uint32_t m16 = (uint32_t)(wb >> 1) & 0xFFFFu;  // sample starts at boundaries i + 1 .. i + 16
      uint32_t limit = m16 ? (uint32_t)__ffs(m16) : 16u;
      limit = (uint32_t)min((long long)min(limit, p.max_len), N - i);
      const uint8_t* tx = s_text + 16 + j;
      const int row = j - SG_LH;
      uint32_t xb = p.root_base;
#pragma unroll
      for (int d = 0; d < 16; d++) {
        if ((uint32_t)d >= limit) break;
        const uint32_t cw = 0x100u | tx[d];
        const uint4 e = (d < HOT) ? s_hot[xb ^ cw] : __ldg(p.trie + (xb ^ cw));
        if ((e.x ^ cw) & 0x1FFu) break;
        if (e.y & F_TERM) {
          mask |= 1u << d;
          if (row >= 0) s_tab[d * SG_ROWS + row] = __hiloint2double((int)e.w, (int)e.z);
        }
        if (!(e.y & F_HASCH)) break;
        xb = e.x >> 9;
      }
    }
    s_mask[j] = (uint16_t)mask;

    // ---- cuts: boundary j is a cut iff no earlier start reaches beyond it (exclusive prefix max of the reach)
    uint32_t incl = (uint32_t)j + (32u - (uint32_t)__clz(mask));
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl = max(incl, y);
    }
    if (lane == 31) s_ws[warp] = incl;
    __syncthreads();
    uint32_t prev = 0;
    for (int w = 0; w < warp; w++) prev = max(prev, s_ws[w]);
    incl = max(incl, prev);
    uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) excl = prev;
    const bool cut = excl <= (uint32_t)j;
    const uint32_t cb = __ballot_sync(0xFFFFFFFFu, cut);
    if (lane == 0) s_cut[warp] = cb;
    if (tid == SG_THREADS - 1) s_cut[SG_WARPS] = (incl <= (uint32_t)SG_W) ? 1u : 0u;  // boundary j = SG_W
    const bool is_start = cut && j >= SG_LH && j < SG_LH + SG_TP && i < N;
    const uint32_t sb = __ballot_sync(0xFFFFFFFFu, is_start);
    if (lane == 0) s_ws[18 + warp] = __popc(sb);
    __syncthreads();
    uint32_t before = 0, nseg = 0;
    for (int w = 0; w < SG_WARPS; w++) {
      const uint32_t c = s_ws[18 + w];
      if (w < warp) before += c;
      nseg += c;
    }
    if (is_start) s_seg[before + __popc(sb & ((1u << lane) - 1u))] = (uint16_t)j;
    __syncthreads();

    // ---- solve: one lane per segment
    if ((uint32_t)tid < nseg) {
      const int c0 = s_seg[tid];
      const uint32_t r = (uint32_t)c0 + 1u, k = r >> 5, sh = r & 31u;
      unsigned long long w = ((unsigned long long)s_cut[k] | ((unsigned long long)s_cut[k + 1] << 32)) >> sh;
      if (sh) w |= (unsigned long long)s_cut[k + 2] << (64 - sh);
      const bool known = w != 0;
      const int len = known ? __ffsll((long long)w) : 65;  // next cut = c0 + len
      bool hard = !known || len > SG_MAXSEG;
      int why = hard ? 1 : 0;
      if (!hard) {
        for (int q = c0 + 1; q <= c0 + len; q++) s_d[q] = ninf;
        for (int q = c0; q < c0 + len; q++) {
          const double dq = (q == c0) ? 0.0 : s_d[q];
          if (dq == ninf) continue;  // unreachable start (src/model.rs:85-87)
          uint32_t m = s_mask[q];
          const int row = q - SG_LH;
          while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int t = q + l + 1;
            if (t > c0 + len) { hard = true; continue; }  // (a token across a cut: cannot happen)
            const double cand = __dadd_rn(dq, s_tab[l * SG_ROWS + row]);
            const double diff = __dadd_rn(cand, -s_d[t]);
            if (fabs(diff) <= T) { hard = true; why = 2; }  // tie or near tie: the reference's rounding decides (P2)
            if (diff > 0.0) {
              s_d[t] = cand;
              s_bp[t] = (uint8_t)(l + 1);
            }
          }
        }
        if (s_d[c0 + len] == ninf) { hard = true; why = 3; }  // end of the segment unreachable: NoPath, reported by P2
      }
      if (!hard) {
        int q = c0 + len;
        while (q > c0) {
          const int l = s_bp[q];
          s_mark[q - 1] = (uint8_t)l;
          const int s = q - l;
          unsigned long long lo = 0, hi = 0;
          const uint8_t* tx = s_text + 16 + s;
          for (int d = 0; d < l; d++) {
            const unsigned long long b = tx[d];
            if (d < 8) lo |= b << (8 * d); else hi |= b << (8 * (d - 8));
          }
          p.ids_at[t0 - SG_LH + q - 1] = hash_lookup(p.hash, p.hash_mask, p.hash_seed, lo, hi, (uint32_t)l);
          q = s;
        }
      } else {
        s_mark[c0] = SG_HARD;
      }
      if (p.dbg) {
        atomicAdd(p.dbg + 0, 1ull);
        if (hard) {
          atomicAdd(p.dbg + why, 1ull);
          atomicAdd(p.dbg + 4, (unsigned long long)(known ? len : 64));
        }
      }
      if ((uint32_t)tid == nseg - 1) s_ws[37] = (uint32_t)(c0 + (known ? len : 1));
    }
    __syncthreads();

    // ---- the tile's bytes [first owned cut, end of the last owned segment) are contiguous
    if (nseg) {
      const int jfirst = s_seg[0], jend = (int)s_ws[37];
      for (int x = jfirst + tid; x < jend; x += SG_THREADS) p.mark[t0 - SG_LH + x] = s_mark[x];
    }
    __syncthreads();
  }
}

// -----------------------------------------------------------------------------------------
// P2: exact chain + hard segments, one warp per sample.
// -----------------------------------------------------------------------------------------
constexpr int CH_WARPS = 16;
constexpr int CH_CHUNK = 256;  // mark bytes per step (8 per lane)
constexpr size_t CH_WARP_BYTES = 256 * 8 + 16 * 16 * 8 + 16 * 16 * 4 + 64 * 8 + 256 + 256;  // 6144

struct ChainParams {
  const uint8_t* text;
  const uint64_t* unit_start;
  const uint32_t* unit_len;
  const uint32_t* order;   // length-descending
  const uint32_t* counts;  // counts[1] = #non-empty units
  const uint4* trie;
  uint32_t root_base, max_len;
  uint8_t* mark;
  uint32_t* ids_at;
  const double* scores;  // [V]
  uint32_t V;
  unsigned long long* n_tokens;
  int32_t* status;
  unsigned int* counter;
  unsigned long long* dbg;
};

__global__ void __launch_bounds__(CH_WARPS * 32, 2) seg_chain_kernel(ChainParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wb = smem + (size_t)warp * CH_WARP_BYTES;
  double* s_w = reinterpret_cast<double*>(wb);            // [256] score of entry k
  double* s_tab = s_w + 256;                              // [16 starts][16 lens]
  double* s_dp = s_tab + 256;                             // [64] ring
  uint32_t* s_tid = reinterpret_cast<uint32_t*>(s_dp + 64);  // [16][16]
  uint8_t* s_pos = reinterpret_cast<uint8_t*>(s_tid + 256);  // [256] byte offset of entry k in the chunk
  uint8_t* s_val = s_pos + 256;                           // [256] mark value
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
  const uint32_t nn = p.counts[1];

  for (;;) {
    uint32_t idx = 0;
    if (lane == 0) idx = atomicAdd(p.counter, 1u);
    idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
    if (idx >= nn) break;
    const uint32_t unit = p.order[idx];
    const unsigned long long start = p.unit_start[unit];
    const unsigned long long E = start + p.unit_len[unit];
    double X = 0.0;  // dp[pos].score, exact (src/model.rs:72-81: dp[0] = 0.0)
    unsigned long long cnt = 0;
    int st = 0;
    unsigned long long pos = start;
    while (pos < E && st == 0) {
      const unsigned long long cb = pos & ~7ull;
      const unsigned long long a = cb + 8ull * lane;
      if (p.dbg && lane == 0) atomicAdd(p.dbg + 7, 1ull);
      unsigned long long v = 0;
      if (a < E) {
        v = *reinterpret_cast<const volatile unsigned long long*>(p.mark + a);
        if (a < pos) v &= ~0ull << (8 * (pos - a));
        if (a + 8 > E) v &= ~0ull >> (8 * (a + 8 - E));
      }
      const uint32_t vlo = (uint32_t)v, vhi = (uint32_t)(v >> 32);
      const uint32_t c = __popc(nonzero_bytes(vlo)) + __popc(nonzero_bytes(vhi));
      uint32_t o = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, o, d);
        if (lane >= d) o += y;
      }
      const uint32_t total = __shfl_sync(0xFFFFFFFFu, o, 31);
      o -= c;
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
        const uint32_t b = (uint32_t)(v >> (8 * kk)) & 0xFFu;
        if (b) {
          s_pos[o] = (uint8_t)(8 * lane + kk);
          s_val[o] = (uint8_t)b;
          o++;
        }
      }
      __syncwarp();
      uint32_t hm[8];
#pragma unroll
      for (int bi = 0; bi < 8; bi++) {
        const uint32_t k = bi * 32 + lane;
        bool h = false;
        if (k < total) {
          if (s_val[k] == SG_HARD) {
            h = true;
            s_w[k] = 0.0;
          } else {
            const uint32_t id = p.ids_at[cb + s_pos[k]];
            s_w[k] = id < p.V ? __ldg(p.scores + id) : 0.0;
          }
        }
        hm[bi] = __ballot_sync(0xFFFFFFFFu, h);
      }
      __syncwarp();
      uint32_t k = 0;
      unsigned long long next_pos = cb + CH_CHUNK;
      while (k < total) {
        uint32_t h = total;
#pragma unroll
        for (int bi = 7; bi >= 0; bi--) {
          uint32_t mm = hm[bi];
          const int lo_k = (int)k - bi * 32;
          if (lo_k > 0) mm = (lo_k >= 32) ? 0u : (mm & (0xFFFFFFFFu << lo_k));
          if (mm) h = bi * 32 + __ffs(mm) - 1;
        }
        cnt += h - k;
        for (; k + 4 <= h; k += 4) {  // dp[e].score = dp[s].score + vocab[id].score  (src/model.rs:98)
          const double w0 = s_w[k], w1 = s_w[k + 1], w2 = s_w[k + 2], w3 = s_w[k + 3];
          X = __dadd_rn(X, w0);
          X = __dadd_rn(X, w1);
          X = __dadd_rn(X, w2);
          X = __dadd_rn(X, w3);
        }
        for (; k < h; k++) X = __dadd_rn(X, s_w[k]);
        if (h >= total) break;

        // ---- hard segment starting at boundary q: the reference's forward loop (src/model.rs:83-110) from the
        //      exact dp[q] = X until the next cut, 16 starts walked in parallel, relaxed in order.
        const unsigned long long q = cb + s_pos[h];
        s_dp[lane] = ninf;
        s_dp[lane + 32] = ninf;
        if (lane == 0) p.mark[q] = 0;  // the marker itself
        __syncwarp();
        if (lane == 0) s_dp[q & 63] = X;
        unsigned long long maxreach = q, e = 0;
        if (p.dbg && lane == 0) atomicAdd(p.dbg + 5, 1ull);
        for (unsigned long long p0 = q; e == 0; p0 += 16) {
          uint32_t mask = 0;
          if (p.dbg && lane == 0) atomicAdd(p.dbg + 6, 1ull);
          __syncwarp();
          if (lane < 16) {
            const unsigned long long sp = p0 + lane;
            if (sp < E) {
              const uint32_t limit = (uint32_t)min((unsigned long long)min(16u, p.max_len), E - sp);
              uint32_t xb = p.root_base;
              for (uint32_t d = 0; d < limit; d++) {
                const uint32_t cw = 0x100u | __ldg(p.text + sp + d);
                const uint4 en = __ldg(p.trie + (xb ^ cw));
                if ((en.x ^ cw) & 0x1FFu) break;
                if (en.y & F_TERM) {
                  mask |= 1u << d;
                  s_tab[lane * 16 + d] = __hiloint2double((int)en.w, (int)en.z);
                  s_tid[lane * 16 + d] = en.y & ID_MASK;
                }
                if (!(en.y & F_HASCH)) break;
                xb = en.x >> 9;
              }
            }
          }
          __syncwarp();
          for (int s = 0; s < 16; s++) {
            const unsigned long long sp = p0 + s;
            if (sp > q && maxreach <= sp) { e = sp; break; }
            const uint32_t m = __shfl_sync(0xFFFFFFFFu, mask, s);
            maxreach = max(maxreach, sp + (32u - (uint32_t)__clz(m)));
            const double dq = s_dp[sp & 63];
            __syncwarp();
            if (dq != ninf && lane < 16 && ((m >> lane) & 1u)) {
              const double cand = __dadd_rn(dq, s_tab[s * 16 + lane]);
              const unsigned long long t = sp + lane + 1;
              if (cand > s_dp[t & 63]) {  // unset (-inf) or strictly better  (src/model.rs:100)
                s_dp[t & 63] = cand;
                p.mark[t - 1] = (uint8_t)(lane + 1);
                p.ids_at[t - 1] = s_tid[s * 16 + lane];
              }
            }
            if (lane == 0) s_dp[sp & 63] = ninf;  // the slot serves position sp + 64 next
            __syncwarp();
          }
        }
        const double xe = s_dp[e & 63];
        if (xe == ninf) { st = 6; break; }  // a cut no path reaches: Error::NoPath(n, n)  (src/model.rs:119)
        X = xe;
        __syncwarp();
        uint32_t ctok = 0;
        if (lane == 0) {  // backtrack inside the segment: flag the token ends of the best path
          unsigned long long x = e;
          while (x > q) {
            const uint32_t l = *reinterpret_cast<volatile uint8_t*>(p.mark + x - 1);
            if (l == 0 || l > 64 || l > x - q) { ctok = 0xFFFFFFFFu; break; }
            *reinterpret_cast<volatile uint8_t*>(p.mark + x - 1) = (uint8_t)(l | 0x80u);
            x -= l;
            ctok++;
          }
        }
        ctok = __shfl_sync(0xFFFFFFFFu, ctok, 0);
        if (ctok == 0xFFFFFFFFu) { st = 99; break; }
        __syncwarp();
        for (unsigned long long b = q + lane; b < e; b += 32) {
          const uint32_t l = *reinterpret_cast<volatile uint8_t*>(p.mark + b);
          *reinterpret_cast<volatile uint8_t*>(p.mark + b) = (l & 0x80u) ? (uint8_t)(l & 0x7Fu) : (uint8_t)0;
        }
        __syncwarp();
        cnt += ctok;
        k = h + 1;
        while (k < total && cb + s_pos[k] < e) k++;
        if (e > next_pos) next_pos = e;
      }
      __syncwarp();
      pos = next_pos;
    }
    if (st) {  // failed sample: no marks, no tokens (like backtrack_*_kernel)
      for (unsigned long long b = start + lane; b < E; b += 32) p.mark[b] = 0;
      cnt = 0;
    }
    if (lane == 0) {
      p.n_tokens[unit] = cnt;
      p.status[unit] = st;
    }
    __syncwarp();
  }
}

}  // namespace tgxk
