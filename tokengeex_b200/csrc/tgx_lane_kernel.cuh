// K2L  Viterbi forward, thread-per-sample ("lane") form: the throughput path of encode.
//
// The ordered relax chain of Model::encode (src/model.rs:83-110) is serial per sample, so the
// only bulk parallelism is ACROSS samples.  Here every lane of every warp owns one sample at a
// time and runs two decoupled state machines over a lane-private shared-memory ring:
//
//   walker    TrieIterator::next (src/trie.rs:51-63) as a uniform loop body: one double-array
//             probe per step, whatever (position, depth) the lane is at.  A terminal hit appends
//             {score, len, FIRST?} to the ring; a miss / leaf / length limit ends the walk and
//             the lane moves to the next start position.  Lanes never wait for each other's
//             walk depth (the pair kernel's producers do).
//   consumer  pops one ring entry per step.  FIRST marks the first entry of a start position:
//             that position is now final, its back length is emitted (packed, 8 bytes per
//             store) and its score read into a register.  Every entry is one relax:
//             cand = dp[pos] + score, strict '>' against the cell of pos + len, in the
//             reference's order (ascending start, ascending length).  The dp window is 16 (32,
//             64) cells per lane in shared memory, columns indexed by lane: conflict-free.
//
// There is no inter-lane communication at all: no shuffles, no barriers.  The text of a sample
// streams through a 32-byte register window (aligned 8-byte loads, prefetched one word ahead).
// Samples are handed out in length-descending order through a global counter (LPT).
#pragma once
#include "tgx_kernels.cuh"

namespace tgxk {

constexpr int LN_E = 16;  // ring entries per lane (power of two)
constexpr uint32_t LN_FIRST = 0x100u, LN_LAST = 0x200u, LN_HDR = 0x400u;
constexpr int LN_MAX_WARPS = 16;

struct LaneParams {
  UnitParams u;
  const uint8_t* blob_end;
  uint8_t* bp;  // [N] back length per end position
  unsigned int* counter;
  uint32_t hot_slots;  // leading trie slots staged in shared memory
};

template <int CELLS>
__host__ __device__ constexpr size_t lane_warp_bytes() {
  return (size_t)32 * (LN_E * 12 + CELLS * 12);
}

// Eight text bytes at the 8-byte aligned address q; bytes at or beyond blob_end read as zero.
__device__ __forceinline__ uint2 lane_load8(const uint8_t* q, const uint8_t* blob_end) {
  if (q + 8 <= blob_end) return __ldg(reinterpret_cast<const uint2*>(q));
  uint32_t lo = 0, hi = 0;
  for (int k = 0; k < 8; k++) {
    if (q + k < blob_end) {
      const uint32_t b = __ldg(q + k);
      if (k < 4) lo |= b << (8 * k); else hi |= b << (8 * (k - 4));
    }
  }
  return make_uint2(lo, hi);
}

enum : int { LW_IDLE = 0, LW_FETCH = 1, LW_WALK = 2, LW_TERM = 3 };

// Body shared by viterbi_lane_kernel and the hybrid kernel.  Called by every thread of the CTA
// (the hot trie prefix is staged cooperatively); warps >= nwarps return after the staging.
template <int CELLS, int KW, int KC>
__device__ __forceinline__ void lane_body(const LaneParams& p, unsigned char* smem, uint32_t nwarps, bool stage_hot) {
  const UnitParams& u = p.u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);

  if (stage_hot) {
    uint4* hw = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < p.hot_slots; i += blockDim.x) hw[i] = __ldg(u.trie + i);
  }
  __syncthreads();
  if ((uint32_t)warp >= nwarps) return;
  const uint4* hot = reinterpret_cast<const uint4*>(smem);
  unsigned char* wb = smem + (size_t)p.hot_slots * 16 + (size_t)warp * lane_warp_bytes<CELLS>();
  double* ring_sc = reinterpret_cast<double*>(wb) + lane;                  // [e * 32]
  double* dp_sc = reinterpret_cast<double*>(wb + LN_E * 32 * 8) + lane;    // [c * 32]
  uint32_t* ring_mt = reinterpret_cast<uint32_t*>(wb + (LN_E + CELLS) * 32 * 8) + lane;
  uint32_t* dp_len = ring_mt + LN_E * 32;

  uint32_t ufirst = u.first, ucount = u.count;
  unit_range(u.counts, u.part, ufirst, ucount);
  const uint32_t maxlen = u.rows, root = u.root_base, hot_slots = p.hot_slots;
  const uint4* __restrict__ trie = u.trie;

  // ---- walker state
  int wstate = LW_FETCH;
  uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;  // the 16 bytes at the current start position
  uint32_t n0 = 0, n1 = 0, p0 = 0, p1 = 0;  // the bytes after them / one more word, prefetched
  uint32_t ncnt = 8;
  const uint8_t* nptr = u.text;
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;  // remaining bytes of the current walk
  uint32_t d = 0, xb = root, wrem = 0, lim = 0;
  bool first = true;
  uint32_t wr = 0, rd = 0;
  // ---- consumer state
  uint8_t* cbp = p.bp;
  uint32_t cur = 0;
  double best = ninf;
  unsigned long long acc = 0;

  auto push = [&](double sc, uint32_t mt) {
    const uint32_t wi = wr & (LN_E - 1);
    ring_sc[wi * 32] = sc;
    ring_mt[wi * 32] = mt;
    wr++;
  };
  auto advance_text = [&]() {
    b0 = __funnelshift_r(b0, b1, 8);
    b1 = __funnelshift_r(b1, b2, 8);
    b2 = __funnelshift_r(b2, b3, 8);
    b3 = __funnelshift_r(b3, n0, 8);
    n0 = __funnelshift_r(n0, n1, 8);
    n1 >>= 8;
    if (--ncnt == 0) {
      n0 = p0; n1 = p1;
      const uint2 v = lane_load8(nptr, p.blob_end);
      p0 = v.x; p1 = v.y;
      nptr += 8;
      ncnt = 8;
    }
  };

  auto walker_step = [&]() {
    const bool space = (wr - rd) < (uint32_t)LN_E;
    if (wstate == LW_WALK) {
      if (!space) return;
      const uint32_t cw = __byte_perm(a0, 1u, 0x5540);  // 0x100 | next byte
      const uint32_t slot = xb ^ cw;
      // the first levels of the trie live in shared memory (BFS slot order, trie_build.h)
      const uint4* q = slot < hot_slots ? hot + slot : trie + slot;
      const uint4 e = *q;
      const bool hit = ((e.x ^ cw) & 0x1FFu) == 0;
      if (hit && (e.y & F_TERM)) {
        push(__hiloint2double((int)e.w, (int)e.z), (d + 1) | (first ? LN_FIRST : 0u));
        first = false;
      }
      if (hit && (e.y & F_HASCH) && d + 1 < lim) {
        xb = e.x >> 9;
        d++;
        a0 = __funnelshift_r(a0, a1, 8);
        a1 = __funnelshift_r(a1, a2, 8);
        a2 = __funnelshift_r(a2, a3, 8);
        a3 >>= 8;
      } else {
        // end of the walk from this start position; a position without any match still has to
        // tell the consumer that it exists (the candidate -inf never wins)
        if (first) push(ninf, 1u | LN_FIRST);
        if (--wrem == 0) {
          wstate = LW_TERM;
        } else {
          advance_text();
          a0 = b0; a1 = b1; a2 = b2; a3 = b3;
          d = 0; xb = root; first = true;
          lim = min(maxlen, wrem);
        }
      }
    } else if (wstate != LW_IDLE) {
      if (!space) return;
      if (wstate == LW_TERM) {  // position n: emits the last back length and closes the sample
        push(ninf, 1u | LN_FIRST | LN_LAST);
        wstate = LW_FETCH;
      } else {  // LW_FETCH
        const uint32_t idx = atomicAdd(p.counter, 1u);
        if (idx >= ucount) {
          wstate = LW_IDLE;
        } else {
          const uint32_t unit = u.order[ufirst + idx];
          const uint32_t n = u.unit_len[unit];
          const unsigned long long start = u.unit_start[unit];
          push(__longlong_as_double((long long)start), LN_HDR);
          const uint8_t* s = u.text + start;
          const uint8_t* base = reinterpret_cast<const uint8_t*>(reinterpret_cast<unsigned long long>(s) & ~7ull);
          uint2 v = lane_load8(base, p.blob_end);
          b0 = v.x; b1 = v.y;
          v = lane_load8(base + 8, p.blob_end);
          b2 = v.x; b3 = v.y;
          v = lane_load8(base + 16, p.blob_end);
          n0 = v.x; n1 = v.y;
          v = lane_load8(base + 24, p.blob_end);
          p0 = v.x; p1 = v.y;
          nptr = base + 32;
          ncnt = 8;
          for (uint32_t k = (uint32_t)(reinterpret_cast<unsigned long long>(s) & 7ull); k; k--) advance_text();
          a0 = b0; a1 = b1; a2 = b2; a3 = b3;
          d = 0; xb = root; first = true;
          wrem = n;
          lim = min(maxlen, n);
          wstate = n ? LW_WALK : LW_TERM;
        }
      }
    }
  };

  auto consumer_step = [&]() {
    if (rd == wr) return;
    const uint32_t ri = rd & (LN_E - 1);
    const double sc = ring_sc[ri * 32];
    const uint32_t mt = ring_mt[ri * 32];
    rd++;
    if (mt & LN_HDR) {
      cbp = p.bp + (unsigned long long)__double_as_longlong(sc) - 1;  // back length of position e goes to cbp[e]
      cur = 0xFFFFFFFFu;
      acc = 0;
      // dp[0] = { score 0.0, start Some(0) }, everything else unreached  (src/model.rs:72-81)
#pragma unroll
      for (int c = 0; c < CELLS; c++) dp_sc[c * 32] = c ? ninf : 0.0;
      return;
    }
    if (mt & LN_FIRST) {  // position cur + 1 is final
      cur++;
      const uint32_t c = cur & (CELLS - 1);
      best = dp_sc[c * 32];
      const uint32_t bl = ((uint32_t)__double2hiint(best) == 0xFFF00000u) ? 0u : dp_len[c * 32];
      dp_sc[c * 32] = ninf;  // the cell now stands for position cur + CELLS
      if (cur) {
        acc = (acc >> 8) | ((unsigned long long)bl << 56);
        uint8_t* A = cbp + cur;
        if ((reinterpret_cast<unsigned long long>(A) & 7ull) == 7ull) {
          if (cur >= 8) {
            *reinterpret_cast<unsigned long long*>(A - 7) = acc;
          } else {  // the word started before this sample: only this sample's bytes
            for (uint32_t i = 0; i < cur; i++) A[-(int)i] = (uint8_t)(acc >> (56 - 8 * i));
          }
        }
      }
    }
    if (mt & LN_LAST) {
      uint8_t* A = cbp + cur;
      const uint32_t r = (uint32_t)((reinterpret_cast<unsigned long long>(A) + 1ull) & 7ull);
      const uint32_t cnt = min(r, cur);
      for (uint32_t i = 0; i < cnt; i++) A[-(int)i] = (uint8_t)(acc >> (56 - 8 * i));
      return;
    }
    const uint32_t len = mt & 0xFFu;
    const uint32_t t = (cur + len) & (CELLS - 1);
    const double cand = __dadd_rn(best, sc);  // dp[pos].score + vocab[id].score  (src/model.rs:98)
    const double old = dp_sc[t * 32];
    if (cand > old) {  // start.is_none() || score > node.score  (:100-101); unreached = -inf
      dp_sc[t * 32] = cand;
      dp_len[t * 32] = len;
    }
  };

  for (;;) {
#pragma unroll
    for (int i = 0; i < (KW > KC ? KW : KC); i++) {
      if (i < KW) walker_step();
      if (i < KC) consumer_step();
    }
    const bool done = wstate == LW_IDLE && rd == wr;
    if (__all_sync(0xFFFFFFFFu, done)) break;
  }
}

template <int CELLS, int KW, int KC>
__global__ void __launch_bounds__(LN_MAX_WARPS * 32, 1) viterbi_lane_kernel(LaneParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  lane_body<CELLS, KW, KC>(p, smem, blockDim.x >> 5, true);
}

// K2H  hybrid forward kernel (max_token_len <= 16): ONE persistent launch, one CTA per SM.  The first
// pair_ctas CTAs start on the long samples with the latency-oriented pair-CTA body (a long sample
// is one long dependent chain: it has to start first and run on the fastest path), then join
// the others, which run the throughput-oriented lane body over the remaining samples.
struct HybridParams {
  PairParams pair;
  LaneParams lane;
  uint32_t pair_ctas, lane_warps;
};

template <int R, int HOT, int KW, int KC>
__global__ void __launch_bounds__(800, 1) viterbi_hybrid_kernel(HybridParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  bool staged = false;
  if (blockIdx.x < p.pair_ctas) {
    pair_body<R, HOT>(p.pair, smem);
    staged = HOT > 0;  // same prefix, same place
    __syncthreads();
  }
  lane_body<16, KW, KC>(p.lane, smem, p.lane_warps, !staged);
}

}  // namespace tgxk
