// Host half of the EM pruning loop (product code; the GPU does the E-step and the
// frequency pass).  Follows /root/reference/src/prune.rs:
//   tgx_m_step            run_m_step              :124-170  (+ digamma :322-335)
//   tgx_prune_select      prune_vocab             :173-319  minus the frequency pass (:205-246),
//                                                  which arrives as an input computed on the GPU
// The per-token "alternatives" (:179-203) need Lattice::nbest(2) (src/lattice.rs:152-238) over
// the token's own bytes.  Those lattices are tiny (<= 64 positions), independent per token and
// — unlike the reference, which walks all V tokens serially — are spread over host threads.
// The A* search keeps the reference's exact pop order, including Rust's BinaryHeap sift rules
// for hypotheses with equal fx, because ties change which path is
// "second best" and with it the pruning loss.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tokengeex_b200.h"
#include "trie_build.h"

namespace {

// src/prune.rs:322-335
double digamma(double x) {
  double result = 0.0;
  while (x < 7.0) {
    result -= 1.0 / x;
    x += 1.0;
  }
  x -= 1.0 / 2.0;
  const double xx = 1.0 / x;
  const double xx2 = xx * xx;
  const double xx4 = xx2 * xx2;
  result += std::log(x) + (1.0 / 24.0) * xx2 - 7.0 / 960.0 * xx4 + (31.0 / 8064.0) * xx4 * xx2 -
            (127.0 / 30720.0) * xx4 * xx4;
  return result;
}

// One token's segmentation lattice: arcs sorted by (start asc, len asc) = insertion order of
// Model::populate_nodes (src/model.rs:34-55); node 0 = BOS, node 1 = EOS as in Lattice::from.
struct Arc {
  uint32_t pos, len, id;
  double score;
  double backtrack;  // Lattice::viterbi's backtrack_score
};

struct TokenLattice {
  std::vector<Arc> nodes;                     // [0]=BOS, [1]=EOS, then arcs
  std::vector<std::vector<uint32_t>> ending;  // end_nodes[pos]
  std::vector<std::vector<uint32_t>> beginning;
  size_t n = 0;

  void build(const tgx::DoubleArray& da, const double* scores, const uint8_t* s, size_t len) {
    n = len;
    nodes.clear();
    ending.assign(len + 1, {});
    beginning.assign(len + 1, {});
    nodes.push_back(Arc{0, 0, 0xFFFFFFFEu, 0.0, 0.0});
    nodes.push_back(Arc{(uint32_t)len, 0, 0xFFFFFFFFu, 0.0, 0.0});
    ending[0].push_back(0);
    beginning[len].push_back(1);
    for (size_t pos = 0; pos < len; pos++) {
      tgx::da_common_prefix_search(da, s + pos, len - pos, [&](uint32_t id, uint32_t l) {
        uint32_t idx = (uint32_t)nodes.size();
        beginning[pos].push_back(idx);
        ending[pos + l].push_back(idx);
        nodes.push_back(Arc{(uint32_t)pos, l, id, scores[id], 0.0});
      });
    }
  }

  // Lattice::viterbi (src/lattice.rs:112-138): fills backtrack scores; false if disconnected
  bool viterbi_scores() {
    for (size_t pos = 0; pos <= n; pos++) {
      for (uint32_t r : beginning[pos]) {
        double best = 0.0;
        bool have = false;
        for (uint32_t l : ending[pos]) {
          const double sc = nodes[l].backtrack + nodes[r].score;
          if (!have || sc > best) {
            have = true;
            best = sc;
          }
        }
        if (!have) return false;
        nodes[r].backtrack = best;
      }
    }
    return true;
  }
};

struct Hyp {
  uint32_t node;
  int32_t next;
  double fx, gx;
};

// Rust std::collections::BinaryHeap<Hypothesis>; `a <= b` is `a.fx < b.fx` (src/lattice.rs:370-378)
struct Heap {
  std::vector<int32_t> d;
  const std::vector<Hyp>* h;
  bool le(int32_t a, int32_t b) const { return (*h)[a].fx < (*h)[b].fx; }
  void sift_up(size_t start, size_t pos) {
    const int32_t e = d[pos];
    while (pos > start) {
      const size_t parent = (pos - 1) / 2;
      if (le(e, d[parent])) break;
      d[pos] = d[parent];
      pos = parent;
    }
    d[pos] = e;
  }
  void push(int32_t x) {
    d.push_back(x);
    sift_up(0, d.size() - 1);
  }
  int32_t pop() {
    int32_t item = d.back();
    d.pop_back();
    if (!d.empty()) {
      std::swap(item, d[0]);
      const size_t end = d.size();
      size_t pos = 0;
      const int32_t e = d[0];
      size_t child = 1;
      while (child + 1 < end) {
        if (le(d[child], d[child + 1])) child++;
        d[pos] = d[child];
        pos = child;
        child = 2 * pos + 1;
      }
      if (child == end - 1) {
        d[pos] = d[child];
        pos = child;
      }
      d[pos] = e;
      sift_up(0, pos);
    }
    return item;
  }
};

// Lattice::nbest(2) (src/lattice.rs:152-238).  Returns the number of paths found (<= 2);
// path k's token ids in paths[k] (BOS/EOS excluded, as nbest(n >= 2) does).
int nbest2(TokenLattice& lat, std::vector<Hyp>& arena, Heap& agenda, std::vector<uint32_t> (&paths)[2]) {
  arena.clear();
  agenda.d.clear();
  agenda.h = &arena;
  arena.push_back(Hyp{1, -1, 0.0, 0.0});
  agenda.push(0);
  lat.viterbi_scores();  // a disconnected lattice leaves partial scores, exactly as the reference does
  int found = 0;
  while (!agenda.d.empty()) {
    const int32_t top = agenda.pop();
    const uint32_t node = arena[top].node;
    if (lat.nodes[node].id == 0xFFFFFFFEu) {  // BOS reached: one complete hypothesis
      paths[found].clear();
      int32_t nx = arena[top].next;
      while (arena[nx].next >= 0) {
        paths[found].push_back(lat.nodes[arena[nx].node].id);
        nx = arena[nx].next;
      }
      if (++found == 2) return found;
    } else {
      const double gx0 = arena[top].gx;
      for (uint32_t l : lat.ending[lat.nodes[node].pos]) {
        arena.push_back(Hyp{l, top, lat.nodes[l].backtrack + gx0, lat.nodes[l].score + gx0});
        agenda.push((int32_t)arena.size() - 1);
      }
      if (agenda.d.size() > 100000) {  // k_max_agenda_size; unreachable for <= 64-byte tokens, kept for fidelity
        Heap na;
        na.h = &arena;
        for (int i = 0; i < 20; i++) na.push(agenda.pop());  // min(512, n * 10), n = 2
        agenda.d.swap(na.d);
      }
    }
  }
  return found;
}

// std::stable_sort over `threads` contiguous chunks, then rounds of pairwise std::inplace_merge (stable: the left
// chunk's elements stay in front of equal ones from the right) — the same order as one stable sort.
template <class T, class Cmp>
void stable_sort_threads(std::vector<T>& v, Cmp cmp, int threads) {
  const size_t n = v.size();
  size_t chunks = std::min<size_t>((size_t)std::max(1, threads), n / 64);
  if (chunks < 2) {
    std::stable_sort(v.begin(), v.end(), cmp);
    return;
  }
  std::vector<size_t> cut(chunks + 1);
  for (size_t c = 0; c <= chunks; c++) cut[c] = n * c / chunks;
  {
    std::vector<std::thread> ts;
    for (size_t c = 1; c < chunks; c++)
      ts.emplace_back([&, c] { std::stable_sort(v.begin() + cut[c], v.begin() + cut[c + 1], cmp); });
    std::stable_sort(v.begin(), v.begin() + cut[1], cmp);
    for (auto& t : ts) t.join();
  }
  while (cut.size() > 2) {  // merge neighbours: (0,1) (2,3) ..; an odd chunk at the end waits for the next round
    std::vector<size_t> nc;
    std::vector<std::thread> ts;
    size_t c = 0;
    for (; c + 2 < cut.size(); c += 2) {
      const size_t a = cut[c], m = cut[c + 1], b = cut[c + 2];
      ts.emplace_back([&v, cmp, a, m, b] { std::inplace_merge(v.begin() + a, v.begin() + m, v.begin() + b, cmp); });
      nc.push_back(a);
    }
    for (; c + 1 < cut.size(); c++) nc.push_back(cut[c]);
    nc.push_back(n);
    for (auto& t : ts) t.join();
    cut.swap(nc);
  }
}

}  // namespace

extern "C" {

// run_m_step.  kept[i] = 1 iff token i survives (freq >= 0.5 || keep); new_scores[i] is its new
// score (digamma(max(freq, 0.5)) - digamma(sum)), undefined for dropped tokens.  The sum runs
// over survivors in vocabulary order, left to right, as `iter().map().sum::<f64>()` does.
// Returns TGX_ERR_INVALID if a score is NaN/inf (the reference panics, :154-163).
int tgx_m_step(const double* expected, const uint8_t* keep, uint64_t V, uint8_t* kept, double* new_scores,
               uint64_t* n_kept) {
  if (!expected || !kept || !new_scores) return TGX_ERR_INVALID;
  double sum = 0.0;
  uint64_t k = 0;
  for (uint64_t i = 0; i < V; i++) {
    const double f = expected[i];
    const bool kp = keep && keep[i];
    if (f < 0.5 && !kp) {
      kept[i] = 0;
      continue;
    }
    kept[i] = 1;
    new_scores[i] = std::fmax(f, 0.5);
    sum += new_scores[i];
    k++;
  }
  const double logsum = digamma(sum);
  // (the sum above is the reference's, left to right; the scores below are independent of each other: host threads)
  std::atomic<int> bad{0};
  auto scores_of = [&](uint64_t lo, uint64_t hi) {
    int b = 0;
    for (uint64_t i = lo; i < hi; i++) {
      if (!kept[i]) continue;
      const double s = digamma(new_scores[i]) - logsum;
      if (std::isnan(s) || std::isinf(s)) b = 1;
      new_scores[i] = s;
    }
    if (b) bad.store(1);
  };
  const uint64_t T = V < 256 ? 1 : std::min<uint64_t>(8, std::max(1u, std::thread::hardware_concurrency()));
  std::vector<std::thread> ts;
  for (uint64_t t = 1; t < T; t++) ts.emplace_back(scores_of, V * t / T, V * (t + 1) / T);
  scores_of(0, V / T);
  for (auto& t : ts) t.join();
  if (n_kept) *n_kept = k;
  return bad.load() ? TGX_ERR_INVALID : TGX_OK;
}

// prune_vocab without its frequency pass.  Inputs: the vocabulary (blob/offsets/scores/keep),
// token frequencies from tgx_token_frequencies, the number of samples.  Output: out_ids =
// indices (into the input vocabulary) of the pruned vocabulary IN ITS FINAL ORDER (score
// descending; exact ties keep input order), *out_n its size.  audit[8] (optional):
// always_keep=false count, tokens with alternatives, silent drops, zero-frequency drops,
// candidates, loss tie at the cut (0/1), loss gap at the cut.
// Returns TGX_ERR_INVALID if a loss is not normal (the reference panics, :291-296).
int tgx_prune_select(const uint8_t* token_bytes, const uint64_t* token_offsets, const double* scores,
                     const uint8_t* keep, uint64_t V, const uint64_t* freq, uint64_t n_samples,
                     uint64_t target_vocab_size, double shrink_factor, int threads, uint32_t* out_ids,
                     uint64_t* out_n, double* audit) {
  if (!token_offsets || !scores || !freq || !out_ids || !out_n) return TGX_ERR_INVALID;
  tgx::DoubleArray da;
  std::string err = tgx::build_double_array(token_bytes, token_offsets, scores, V, &da);
  if (!err.empty()) return TGX_ERR_UNSUPPORTED;
  return tgx::prune_select_with(da, token_bytes, token_offsets, scores, keep, V, freq, n_samples, target_vocab_size,
                                shrink_factor, threads, out_ids, out_n, audit);
}

}  // extern "C"

// prune_vocab over a double-array that already holds this vocabulary (tgx_model_prune_select: the model's own trie,
// so the EM loop does not build it a second time per iteration).
int tgx::prune_select_with(const tgx::DoubleArray& da, const uint8_t* token_bytes, const uint64_t* token_offsets,
                           const double* scores, const uint8_t* keep, uint64_t V, const uint64_t* freq,
                           uint64_t n_samples, uint64_t target_vocab_size, double shrink_factor, int threads,
                           uint32_t* out_ids, uint64_t* out_n, double* audit) {
  if (!token_offsets || !scores || !freq || !out_ids || !out_n) return TGX_ERR_INVALID;

  size_t pruned_size = (size_t)((double)V * shrink_factor);         // :174  (truncation)
  pruned_size = std::max<size_t>(pruned_size, target_vocab_size);   // :175

  // ---- alternatives (:179-203) and losses (:247-300), parallel over tokens: a token's loss reads only the input
  // frequencies of its alternative's ids, so it is computed by the thread that found the alternative; the lists below
  // are then filled serially in id order, as the reference's loop does.
  uint64_t sum_u = 0;
  for (uint64_t i = 0; i < V; i++) sum_u += freq[i];
  const double sum_f = (double)sum_u;
  const double logsum = std::log(sum_f);
  enum : uint8_t { KEPT = 0, ZERO_DROP = 1, CANDIDATE = 2, SILENT = 3, BAD_LOSS = 4 };
  std::vector<uint8_t> state(V, KEPT), flags(V, 0);  // flags: 1 = always_keep is false, 2 = has alternatives
  std::vector<double> loss_of(V, 0.0);
  const int T = std::max(1, threads);
  {
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
      TokenLattice lat;
      std::vector<Hyp> arena;
      Heap agenda;
      std::vector<uint32_t> paths[2];
      for (;;) {
        const uint64_t lo = next.fetch_add(256);
        if (lo >= V) break;
        const uint64_t hi = std::min<uint64_t>(V, lo + 256);
        for (uint64_t id = lo; id < hi; id++) {
          lat.build(da, scores, token_bytes + token_offsets[id], (size_t)(token_offsets[id + 1] - token_offsets[id]));
          const int nb = nbest2(lat, arena, agenda, paths);
          const bool always_keep = !(nb > 1 && paths[0].size() > 1);  // :191-195
          const bool has_alt = nb > 1 && paths[0].size() == 1 && !paths[1].empty();  // :197-202
          flags[id] = (uint8_t)((always_keep ? 0 : 1) | (has_alt ? 2 : 0));
          if (keep && keep[id]) continue;  // KEPT
          if (freq[id] == 0 && !always_keep) {
            state[id] = ZERO_DROP;
          } else if (!has_alt) {
            // KEPT
          } else if (freq[id] != 0) {
            const double f = (double)freq[id];
            const double logprob = std::log(f) - logsum;
            // `alternatives.len()` is the OUTER Vec's length, i.e. V (:279)
            const double alt_logsum = std::log(sum_f + f * (double)(V - 1));
            double alt_logprob = 0.0;
            for (uint32_t a : paths[1]) alt_logprob += std::log((double)freq[a] + f) - alt_logsum;
            const double loss = (f / (double)n_samples) * (logprob - alt_logprob);
            loss_of[id] = loss;
            state[id] = std::isnormal(loss) ? CANDIDATE : BAD_LOSS;
          } else {
            state[id] = SILENT;  // freq == 0 && always_keep && has alternatives: falls through every branch
          }
        }
      }
    };
    std::vector<std::thread> ts;
    for (int t = 1; t < T; t++) ts.emplace_back(work);
    work();
    for (auto& t : ts) t.join();
  }

  std::vector<std::pair<uint32_t, double>> candidates;
  std::vector<std::pair<uint32_t, double>> pruned;  // (id, score)
  uint64_t n_akf = 0, n_alt = 0, n_silent = 0, n_zero = 0;
  for (uint64_t id = 0; id < V; id++) {
    if (flags[id] & 1) n_akf++;
    if (flags[id] & 2) n_alt++;
    switch (state[id]) {
      case KEPT: pruned.emplace_back((uint32_t)id, scores[id]); break;
      case ZERO_DROP: n_zero++; break;
      case CANDIDATE: candidates.emplace_back((uint32_t)id, loss_of[id]); break;
      case SILENT: n_silent++; break;
      default: return TGX_ERR_INVALID;  // a loss that is not normal: the reference panics (:291-296)
    }
  }
  // :308 loss descending (exact ties: id ascending)
  const auto by_second_desc = [](const std::pair<uint32_t, double>& a, const std::pair<uint32_t, double>& b) {
    return a.second > b.second;
  };
  stable_sort_threads(candidates, by_second_desc, T);
  size_t taken = 0;
  for (auto& c : candidates) {  // :309-314  (`==`, so an already over-full list takes every candidate)
    if (pruned.size() == pruned_size) break;
    pruned.emplace_back(c.first, scores[c.first]);
    taken++;
  }
  double gap = 0.0, tie = 0.0;
  if (taken > 0 && taken < candidates.size()) {
    gap = candidates[taken - 1].second - candidates[taken].second;
    tie = gap == 0.0 ? 1.0 : 0.0;
  }
  // :316 score descending (exact ties keep the order above)
  stable_sort_threads(pruned, by_second_desc, T);
  for (size_t i = 0; i < pruned.size(); i++) out_ids[i] = pruned[i].first;
  *out_n = pruned.size();
  if (audit) {
    audit[0] = (double)n_akf;
    audit[1] = (double)n_alt;
    audit[2] = (double)n_silent;
    audit[3] = (double)n_zero;
    audit[4] = (double)candidates.size();
    audit[5] = tie;
    audit[6] = gap;
    audit[7] = (double)pruned_size;
  }
  return TGX_OK;
}
