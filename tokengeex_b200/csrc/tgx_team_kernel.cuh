// viterbi_team_kernel: the relax chain of Model::encode (src/model.rs:83-110) over the match stream, FOUR LANES PER
// SAMPLE, eight samples per warp.
//
// Between the two extremes measured in round 2 — 16 lanes per sample (pair_consume / rows_consume: the shortest chain,
// ~95-120 cycles per position, but 15 warp instructions per position) and one lane per sample (the 16 open cells in a
// register ring: 4.6 warp instructions per position, but ~150 instructions in a lane's stream per position: a warp that
// is left alone advances a position per ~800 cycles) — a team of four: lane t of a team owns the candidate lengths 4t .. 4t + 3
// (lane 0: 1, 2, 3 and 16), i.e. two 16-byte pairs of the start's row, and keeps the dp cells those candidates land on
// in a four-slot register ring.  A cell is born in lane 0 (its first candidate is the token of length 16 from the
// earliest start), travels to lane 3, 2, 1 and back to lane 0 — one shuffle per step, always towards the lanes that own
// the LATER starts, so every lane's relax is the reference's `score > node.score` in ascending order of the start —
// and leaves lane 0 as dp[s + 1], which a second shuffle hands to the team as the next step's dp[pos].score.
//
// Same bits as the reference: a candidate is dp[s] + score, one rounding; the cell keeps the strictly greater one, the
// earlier start on ties (src/model.rs:98-101).  -inf = unreached (scores are finite, trie_build.cpp:56).
//
// Records come in aligned groups of four (the team walks the virtual steps sv = 0, 1, .. of the blob positions
// (start & ~3) + sv; the steps in front of the sample see dp = -inf and change nothing), rows from the staged prefix of
// the row table in shared memory or from L1 / L2 through one generic load.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "tgx_kernels.cuh"
#include "tgx_match_kernels.cuh"

namespace tgxk {

struct TeamParams {
  UnitParams u;         // unit_start / unit_len / order / counts+part
  const uint32_t* rec;  // [N + 64] match stream (the 64 entries past N hold 0 = row 0)
  const double* rows;   // row table (padded by 144 bytes: a lane may read up to nine pairs from the start of a row)
  uint32_t hot16;       // leading 16-byte units of the row table staged in shared memory (>= 9: row 0)
  uint8_t* bp;          // [N] back length per end position (0 = unreachable)
  unsigned int* counter;
};

constexpr int TM_PF = 0;  // steps between the L1 prefetch of a cold row and its use (<= 8: the record ring holds 12 ahead)

// 16 warps per SM: measured against 12 / 14 / 20 / 24 (fewer lose throughput, more lengthen every chain and spill below
// 96 registers).  Measured and removed: two lanes per sample (twice the chains per warp, but 400 cycles
// per step for a warp alone: slower), one lane per sample (viterbi_thread_kernel: 146 instructions per position in a
// lane's stream, only samples below 16 KB fit).
constexpr int TM_WARPS = 16;
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) viterbi_team_kernel(TeamParams p) {
  constexpr int G = 4;         // lanes per sample
  constexpr int LPL = 16 / G;  // entries (candidate lengths) per lane = slots of its ring
  constexpr int NP = LPL / 2;  // 16-byte pairs per lane
  extern __shared__ __align__(16) unsigned char smem[];
  const double2* s_rows = reinterpret_cast<const double2*>(smem);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.rows);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < p.hot16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const double2* g_rows = reinterpret_cast<const double2*>(p.rows);
  const int lane = threadIdx.x & 31;
  const uint32_t t = (uint32_t)lane & (uint32_t)(G - 1);  // place in the team
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
  // a row (generic pointer: shared or global); this lane reads its pairs 2t, 2t + 1 and entry 16
  auto row_of = [&](uint32_t r) -> const double2* {  // (a row is at most 9 pairs: staged ones lie wholly inside)
    const uint32_t o = r & REC_OFF;
    return o + 9u <= p.hot16 ? s_rows + o : g_rows + o;
  };
  // candidate lengths of this lane: slot i holds entry 4t + i of the row; lane 0's slot 0 is the length 16 (entry 16)
  uint32_t len[LPL];
#pragma unroll
  for (int i = 0; i < LPL; i++) len[i] = (uint32_t)LPL * t + (uint32_t)i;
  if (t == 0u) len[0] = 16u;
  uint32_t ufirst = p.u.first, ucount = p.u.count;
  unit_range(p.u.counts, p.u.part, ufirst, ucount);
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(p.counter, (unsigned)(32 / G));  // 32 / G samples of the length-descending order per warp (LPT)
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= ucount) break;
    const uint32_t idx = base + (uint32_t)lane / (uint32_t)G;
    const bool have = idx < ucount;
    const uint32_t unit = have ? p.u.order[ufirst + idx] : 0u;
    const uint32_t n = have ? p.u.unit_len[unit] : 0u;
    const unsigned long long start = have ? p.u.unit_start[unit] : 0ull;
    const uint32_t K = (uint32_t)start & 3u;
    const uint32_t total = have ? K + n : 0u;  // starts are the steps K .. K + n - 1; step K + n only emits
    // (+ 4: the word that holds the sample's last byte is stored at the first multiple of four at or after step K + n)
    const uint32_t nsteps = __reduce_max_sync(0xFFFFFFFFu, have ? total + 4u : 0u);
    const uint4* rp = reinterpret_cast<const uint4*>(p.rec + (start - K));
    uint8_t* outp = p.bp + (start - K);  // the back length emitted at step sv belongs to byte sv - 1
    const bool writer = have && t == 0u;

    double c[LPL];
    uint32_t bl[LPL];
#pragma unroll
    for (int i = 0; i < LPL; i++) {
      c[i] = (t == 0u && (uint32_t)i == K) ? 0.0 : ninf;  // (K < 4 <= LPL)  // dp[0] = { score 0.0, start Some(0) }  (src/model.rs:72-81)
      bl[i] = 0u;
    }
    const uint4 none4 = make_uint4(0u, 0u, 0u, 0u);  // row 0: -inf candidates
    uint4 gq[4];  // records: a ring of four groups of four steps; group G of the sample = rp[G]
#pragma unroll
    for (int q = 0; q < 3; q++) gq[q] = (4u * q < total) ? __ldg(rp + q) : none4;
    gq[3] = none4;
    uint32_t pack = 0;
    // scores of the step to come, loaded one step ahead
    // Only the lanes whose pair lies inside the row load it (a team's row is short most of the time: the lanes of the
    // long lengths stay out of the load — the shared-memory / L1 data pipe is what bounds this kernel, ncu: 70 % busy);
    // a register that is not loaded keeps an older score, and the relax of an entry beyond the row is predicated off.
    // Two sets, loaded TWO steps ahead (set k & 1 serves step k): a step of a warp sees 8 rows, one of them cold (L2)
    // most of the time, and one step of lead (~70 instructions) does not cover an L2 access.
    double sc[2][LPL];
    double s16[2] = {ninf, ninf};
#pragma unroll
    for (int i = 0; i < LPL; i++) sc[0][i] = sc[1][i] = ninf;
    auto load_scores = [&](uint32_t r, int w) {
      const double2* b = row_of(r) + (uint32_t)NP * t;
      const uint32_t L1n = (r >> 28) + 1u;
#pragma unroll
      for (int j = 0; j < NP; j++) {
        const uint32_t need = (uint32_t)LPL * t + 2u * (uint32_t)j <= L1n ? 1u : 0u;
        asm("{\n\t"
            ".reg .pred q;\n\t"
            "setp.ne.u32 q, %3, 0;\n\t"
            "@q ld.v2.f64 {%0, %1}, [%2];\n\t"
            "}"
            : "+d"(sc[w][2 * j]), "+d"(sc[w][2 * j + 1])
            : "l"(b + j), "r"(need));
      }
      {  // entry 16: lane 0's candidate of length 16
        const uint32_t need = (t == 0u && L1n >= 16u) ? 1u : 0u;
        asm("{\n\t"
            ".reg .pred q;\n\t"
            "setp.ne.u32 q, %2, 0;\n\t"
            "@q ld.f64 %0, [%1];\n\t"
            "}"
            : "+d"(s16[w])
            : "l"(reinterpret_cast<const double*>(b) + 16), "r"(need));
      }
    };
    load_scores(gq[0].x, 0);
    load_scores(gq[0].y, 1);
    for (uint32_t sv0 = 0; sv0 < nsteps; sv0 += 16) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        if ((k & 7) == 0) {  // the record stream, a line and a half ahead: into L1 (a line = 32 steps of this team)
          const uint32_t G = (sv0 >> 2) + (uint32_t)(k >> 2) + 12u;
          asm volatile("{\n\t"
                       ".reg .pred q;\n\t"
                       "setp.lt.u32 q, %1, %2;\n\t"
                       "@q prefetch.global.L1 [%0];\n\t"
                       "}" ::"l"(rp + G), "r"(4u * G), "r"(total));
        }
        if ((k & 3) == 0) {  // the group three ahead replaces the one that was just finished
          const uint32_t G = (sv0 >> 2) + (uint32_t)(k >> 2) + 3u;
          gq[((k >> 2) + 3) & 3] = (4u * G < total) ? __ldg(rp + G) : none4;
        }
        const uint4 r4 = gq[(k >> 2) & 3];
        const uint32_t rs = (k & 3) == 0 ? r4.x : (k & 3) == 1 ? r4.y : (k & 3) == 2 ? r4.z : r4.w;
        const uint4 q4 = gq[((k + 2) >> 2) & 3];  // the record of the step after next
        const uint32_t rn = ((k + 2) & 3) == 0 ? q4.x : ((k + 2) & 3) == 1 ? q4.y : ((k + 2) & 3) == 2 ? q4.z : q4.w;
        if (TM_PF > 0) {  // a cold row of the step TM_PF ahead: into L1
          const int kp = k + TM_PF;
          const uint4 p4 = gq[(kp >> 2) & 3];
          const uint32_t rq = (kp & 3) == 0 ? p4.x : (kp & 3) == 1 ? p4.y : (kp & 3) == 2 ? p4.z : p4.w;
          asm volatile("{\n\t"
                       ".reg .pred q;\n\t"
                       "setp.gt.u32 q, %1, %2;\n\t"
                       "@q prefetch.global.L1 [%0];\n\t"
                       "}" ::"l"(g_rows + (rq & REC_OFF) + (uint32_t)NP * t), "r"((rq & REC_OFF) + 9u), "r"(p.hot16));
        }
        const int k0 = k % LPL;
        // dp[sv].score: lane 0's slot k0 — the cell that left its window at the last step, final
        const double cur = __shfl_sync(0xFFFFFFFFu, c[k0], 0, G);
        if ((k & 3) == 0) {
          pack |= bl[k0] << 24;  // (written by lane 0 only: its bl[k0] is the back length of position sv)
          const uint32_t sv = sv0 + k;
          if (sv >= 4u && writer) {  // bytes w0 .. w0 + 3 are complete (the steps w0 + 1 .. w0 + 4 = sv emitted them)
            const uint32_t w0 = sv - 4u;
            if (w0 >= K && sv <= total) {
              *reinterpret_cast<uint32_t*>(outp + w0) = pack;
            } else if (w0 + 3u >= K && w0 < total) {
#pragma unroll
              for (int b = 0; b < 4; b++)
                if (w0 + b >= K && w0 + b < total) outp[w0 + b] = (uint8_t)(pack >> (8 * b));
            }
          }
          pack = 0;
        } else {
          pack |= bl[k0] << (8 * ((k + 3) & 3));
        }
        // slot 0: the oldest cell of this lane's window (lane 0: a new cell, born at sv + 16)
        double x = t == 0u ? ninf : c[k0];
        uint32_t xb = t == 0u ? 0u : bl[k0];
        const uint32_t L1 = (rs >> 28) + 1u;  // entries 1 .. L1 of the row are real (scores or -inf), the rest is not this row
        // candidate: dp[pos].score + vocab[id].score (src/model.rs:98), kept if strictly greater (:100-101)
        {
          const double cand = __dadd_rn(cur, t == 0u ? s16[k & 1] : sc[k & 1][0]);
          if (len[0] <= L1 && cand > x) {
            x = cand;
            xb = len[0];
          }
        }
#pragma unroll
        for (int i = 1; i < LPL; i++) {
          const double cand = __dadd_rn(cur, sc[k & 1][i]);
          if (len[i] <= L1 && cand > c[(k + i) % LPL]) {
            c[(k + i) % LPL] = cand;
            bl[(k + i) % LPL] = len[i];
          }
        }
        load_scores(rn, k & 1);  // the scores of the step after next, into the set that was just used
        // the cell moves on to the lane that owns the later starts (lane 0's newborn to the last lane): it is the
        // newest cell of that lane's window from the next step on, in the slot this step's oldest one just left
        c[k0] = __shfl_sync(0xFFFFFFFFu, x, (int)((t + 1u) & (uint32_t)(G - 1)), G);
        bl[k0] = __shfl_sync(0xFFFFFFFFu, xb, (int)((t + 1u) & (uint32_t)(G - 1)), G);
      }
    }
  }
}

}  // namespace tgxk
