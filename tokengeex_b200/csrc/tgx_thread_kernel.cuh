// viterbi_thread_kernel: the relax chain of Model::encode (src/model.rs:83-110) over the match stream, ONE LANE PER
// SAMPLE.
//
// The cooperative consumers (pair_consume, rows_consume) give a sample 16 lanes and put a shuffle on the chain:
// 2 chains per warp, ~15-31 warp instructions per position.  Here a lane owns a whole sample and keeps the 16 open dp
// cells (the positions a token that starts at `s` can end at: s + 1 .. s + 16) in REGISTERS — a ring indexed by the
// unrolled step — so the chain of a sample is DADD -> DSETP -> SEL with no shuffle, no shared memory and no barrier,
// and a warp carries 32 chains: ~3-4 warp instructions per position.
//
// Same bits as the reference: the candidate of (start s, length l) is dp[s] + score, one rounding, and the cell keeps
// the strictly greater one in ascending order of s (first wins on ties) — src/model.rs:98-101.  A cell that no token
// reaches stays -inf with back length 0 ("unreachable": scores are finite, trie_build.cpp:56).
//
// Alignment trick: lane's sample starts at byte `start`; the lane walks the virtual steps sv = 0, 1, ... of the blob
// positions (start & ~3) + sv, so that its records come in aligned 16-byte groups of four; the K = start & 3 steps in
// front of the sample see dp = -inf and change nothing, dp[0] = 0.0 sits in ring cell K.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "tgx_kernels.cuh"
#include "tgx_match_kernels.cuh"

namespace tgxk {

struct ThreadParams {
  UnitParams u;         // unit_start / unit_len / order / counts+part
  const uint32_t* rec;  // [N + 64] match stream (the 64 entries past N hold 0 = row 0)
  const double* rows;   // row table
  uint32_t hot16;       // leading 16-byte units of the row table staged in shared memory (>= 9: row 0)
  uint8_t* bp;          // [N] back length per end position (0 = unreachable)
  unsigned int* counter;
};

constexpr int TK_PF = 8;  // steps between the L1 prefetch of a row and its use (<= 16)

// Entries 2j, 2j + 1 of the row at `base` with one 16-byte load where the row reaches that far (`in`), else the same
// entries of row 0 (-inf; `zero` = its copy in shared memory: one address for all such lanes, a broadcast).  `base` is a
// GENERIC pointer: into the staged prefix of the row table (shared memory) for the hot rows, into the table itself
// (L1 / L2) for the rest — one LD per pair whichever it is, nothing predicated, ptxas is free to hoist the loads.
// A pair that starts inside a row lies inside its padding: its second entry is a score or -inf.
// (Why shared memory: 32 lanes x 32 different rows is a gather; L1 serves ~1.5 sixteen-byte gathers per clock and SM
// (profiles/r01_ubench_gather.txt), shared memory ~3-4, and a step needs ~3.)
__device__ __forceinline__ void tk_ld2(const double2* base, const double2* zero, bool in, int j, double& a, double& b) {
  const double2 v = (in ? base : zero)[j];
  a = v.x;
  b = v.y;
}

__device__ __forceinline__ void thread_body(const ThreadParams& p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const double2* s_rows = reinterpret_cast<const double2*>(smem);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.rows);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < p.hot16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const double2* g_rows = reinterpret_cast<const double2*>(p.rows);
  auto row_of = [&](uint32_t r) -> const double2* {  // (generic: shared or global)
    const uint32_t o = r & REC_OFF;
    return o + 9u <= p.hot16 ? s_rows + o : g_rows + o;  // (a row is at most 9 pairs: staged ones lie wholly inside)
  };
  const int lane = threadIdx.x & 31;
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
  uint32_t ufirst = p.u.first, ucount = p.u.count;
  unit_range(p.u.counts, p.u.part, ufirst, ucount);
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(p.counter, 32u);  // 32 samples of the length-descending order per warp (LPT)
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= ucount) break;
    const uint32_t idx = base + lane;
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

    double c[16];
    uint32_t bl[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      c[i] = (i < 4 && (uint32_t)i == K) ? 0.0 : ninf;  // dp[0] = { score 0.0, start Some(0) }  (src/model.rs:72-81)
      bl[i] = 0u;
    }
    const uint4 none4 = make_uint4(0u, 0u, 0u, 0u);  // row 0 with L = 1: one -inf candidate (REC_NOMATCH would hold the warp for 16)
    // records: a ring of four groups of four steps; group G of the sample = rp[G]
    uint4 gq[4];
#pragma unroll
    for (int q = 0; q < 3; q++) gq[q] = (4u * q < total) ? __ldg(rp + q) : none4;
    gq[3] = none4;
    uint32_t pack = 0;
    // The scores of a step are loaded ONE STEP AHEAD (a warp waits for the slowest of its 32 rows, and the relaxes of a
    // step are all that can hide a load): f = pairs 0..2 of the row (lengths 1..5), always; gB / gC / gD = pairs 3..4,
    // 5..6, 7..8 (lengths 6..9, 10..13, 14..16) when somebody in the warp has a token that long at that step.
    double f[6], gB[4], gC[4], gD[4];
    uint32_t Lw;
    {
      const uint32_t r0 = gq[0].x;
      const uint32_t L0 = r0 >> 28;
      const double2* b0 = row_of(r0);
      Lw = __reduce_max_sync(0xFFFFFFFFu, L0);
      tk_ld2(b0, s_rows, true, 0, f[0], f[1]);
      tk_ld2(b0, s_rows, L0 >= 1u, 1, f[2], f[3]);
      tk_ld2(b0, s_rows, L0 >= 3u, 2, f[4], f[5]);
      tk_ld2(b0, s_rows, L0 >= 5u, 3, gB[0], gB[1]);
      tk_ld2(b0, s_rows, L0 >= 7u, 4, gB[2], gB[3]);
      tk_ld2(b0, s_rows, L0 >= 9u, 5, gC[0], gC[1]);
      tk_ld2(b0, s_rows, L0 >= 11u, 6, gC[2], gC[3]);
      tk_ld2(b0, s_rows, L0 >= 13u, 7, gD[0], gD[1]);
      tk_ld2(b0, s_rows, L0 >= 15u, 8, gD[2], gD[3]);
    }
    for (uint32_t sv0 = 0; sv0 < nsteps; sv0 += 16) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        if ((k & 7) == 0) {  // the record stream, a line and a half ahead: into L1 (a line = 32 steps of this lane)
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
        const uint4 q4 = gq[((k + 1) >> 2) & 3];
        const uint32_t rn = ((k + 1) & 3) == 0 ? q4.x : ((k + 1) & 3) == 1 ? q4.y : ((k + 1) & 3) == 2 ? q4.z : q4.w;
        if (TK_PF > 0) {  // a cold row of the step TK_PF ahead: into L1
          const int kp = k + TK_PF;
          const uint4 p4 = gq[(kp >> 2) & 3];
          const uint32_t rq = (kp & 3) == 0 ? p4.x : (kp & 3) == 1 ? p4.y : (kp & 3) == 2 ? p4.z : p4.w;
          const double* rowp = p.rows + (size_t)(rq & REC_OFF) * 2u;
          asm volatile("{\n\t"
                       ".reg .pred q;\n\t"
                       "setp.gt.u32 q, %1, %2;\n\t"
                       "@q prefetch.global.L1 [%0];\n\t"
                       "}" ::"l"(rowp), "r"((rq & REC_OFF) + 9u), "r"(p.hot16));
        }
        const double cur = c[k];  // dp[sv].score, final: every start < sv has been relaxed into it
        pack |= bl[k] << (8 * ((k + 3) & 3));
        c[k] = ninf;  // the cell moves on to position sv + 16
        bl[k] = 0u;
        if ((k & 3) == 0) {  // bytes w0 .. w0 + 3 are complete (the steps w0 + 1 .. w0 + 4 = sv emitted them)
          const uint32_t sv = sv0 + k;
          if (sv >= 4u && have) {
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
        }
        // candidate of length l: dp[pos].score + vocab[id].score (src/model.rs:98), kept if strictly greater (:100-101)
        auto relax = [&](double sc, int l) {
          const double cand = __dadd_rn(cur, sc);
          const int t = (k + l) & 15;
          if (cand > c[t]) {
            c[t] = cand;
            bl[t] = (uint32_t)l;
          }
        };
        const uint32_t Ln = rn >> 28;  // the step to come
        const double2* bn = row_of(rn);
        const uint32_t Lwn = __reduce_max_sync(0xFFFFFFFFu, Ln);
        relax(f[1], 1);
        relax(f[2], 2);
        relax(f[3], 3);
        relax(f[4], 4);
        relax(f[5], 5);
        tk_ld2(bn, s_rows, true, 0, f[0], f[1]);
        tk_ld2(bn, s_rows, Ln >= 1u, 1, f[2], f[3]);
        tk_ld2(bn, s_rows, Ln >= 3u, 2, f[4], f[5]);
        if (Lw >= 5u) {  // somebody in the warp has a token of 6 bytes or more here
          relax(gB[0], 6);
          relax(gB[1], 7);
          relax(gB[2], 8);
          relax(gB[3], 9);
          if (Lw >= 9u) {
            relax(gC[0], 10);
            relax(gC[1], 11);
            relax(gC[2], 12);
            relax(gC[3], 13);
            if (Lw >= 13u) {
              relax(gD[0], 14);
              relax(gD[1], 15);
              relax(gD[2], 16);
            }
          }
        }
        if (Lwn >= 5u) {
          tk_ld2(bn, s_rows, Ln >= 5u, 3, gB[0], gB[1]);
          tk_ld2(bn, s_rows, Ln >= 7u, 4, gB[2], gB[3]);
          if (Lwn >= 9u) {
            tk_ld2(bn, s_rows, Ln >= 9u, 5, gC[0], gC[1]);
            tk_ld2(bn, s_rows, Ln >= 11u, 6, gC[2], gC[3]);
            if (Lwn >= 13u) {
              tk_ld2(bn, s_rows, Ln >= 13u, 7, gD[0], gD[1]);
              tk_ld2(bn, s_rows, Ln >= 15u, 8, gD[2], gD[3]);
            }
          }
        }
        Lw = Lwn;
      }
    }
  }
}

// One CTA per SM (it holds the staged rows): <16, 1> = 16 warps in 128 registers, <12, 1> = 12 warps in 168,
// <8, 1> = 8 warps, no register limit to speak of.
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) viterbi_thread_kernel(ThreadParams p) {
  thread_body(p);
}

}  // namespace tgxk
