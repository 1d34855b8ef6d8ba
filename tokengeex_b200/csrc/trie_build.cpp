// XOR double-array construction — see trie_build.h.
#include "trie_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <functional>
#include <numeric>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace tgx {
namespace {

struct Node {
  uint32_t first_child = 0;   // 0 = none (node 0 is the root and never a child)
  uint32_t last_child = 0;
  uint32_t next_sibling = 0;
  int32_t term_id = -1;
  uint32_t slot = 0;
  uint8_t label = 0;
};

// Free slots and unused base values of one 256-entry block, as bitmaps (bit set = free / unused).
struct Block {
  uint64_t free_slots[4];
  uint64_t free_bases[4];
};

// out bit f = in bit (f ^ L): the XOR with a constant permutes the 256 positions — whole words by L's two high
// bits, and inside a word by one butterfly stage per set low bit.
inline void xor_permute(const uint64_t (&in)[4], uint32_t L, uint64_t (&out)[4]) {
  static const uint64_t M[6] = {0x5555555555555555ull, 0x3333333333333333ull, 0x0F0F0F0F0F0F0F0Full,
                                0x00FF00FF00FF00FFull, 0x0000FFFF0000FFFFull, 0x00000000FFFFFFFFull};
  for (int w = 0; w < 4; w++) {
    uint64_t x = in[w ^ (L >> 6)];
    for (int j = 0; j < 6; j++)
      if ((L >> j) & 1u) {
        const int sft = 1 << j;
        x = ((x & M[j]) << sft) | ((x >> sft) & M[j]);
      }
    out[w] = x;
  }
}

}  // namespace

std::string build_double_array(const uint8_t* bytes, const uint64_t* off, const double* scores,
                               uint64_t V, DoubleArray* out, bool hot_order) {
  if (V >= MAX_VOCAB) return "vocabulary too large (ids must fit 24 bits)";
  uint32_t max_len = 0;
  for (uint64_t i = 0; i < V; i++) {
    uint64_t l = off[i + 1] - off[i];
    if (l > MAX_TOKEN_LEN) return "token longer than 64 bytes is not supported by the device kernels";
    if (!std::isfinite(scores[i])) return "non-finite token score";
    max_len = std::max<uint32_t>(max_len, (uint32_t)l);
  }

  auto T0 = std::chrono::steady_clock::now();
  const bool timing = getenv("TGX_BUILD_TIMING") != nullptr;  // developer aid: phase times on stderr
  auto lap = [&](const char* w) {
    if (!timing) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "  build_double_array %s %.1f ms\n", w, std::chrono::duration<double, std::milli>(t - T0).count());
    T0 = t;
  };
  // 1. sort token indices by bytes (ties: id ascending, so the last duplicate is seen last)
  //    (the first 8 bytes, big endian, decide most comparisons without touching the blob)
  struct Key {
    uint64_t k;
    uint32_t id;
  };
  std::vector<Key> keyed;
  keyed.reserve(V);
  for (uint64_t i = 0; i < V; i++) {
    const uint64_t l = off[i + 1] - off[i];
    if (!l) continue;  // the empty token never matches
    uint64_t k = 0;
    for (uint64_t d = 0; d < 8; d++) k = (k << 8) | (d < l ? bytes[off[i] + d] : 0u);
    keyed.push_back(Key{k, (uint32_t)i});
  }
  std::sort(keyed.begin(), keyed.end(), [&](const Key& x, const Key& y) {
    if (x.k != y.k) return x.k < y.k;
    const uint32_t a = x.id, b = y.id;
    size_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b];
    int c = std::memcmp(bytes + off[a], bytes + off[b], std::min(la, lb));
    if (c) return c < 0;
    if (la != lb) return la < lb;
    return a < b;
  });
  std::vector<uint32_t> order;
  order.reserve(keyed.size());
  for (const Key& x : keyed) order.push_back(x.id);

  lap("sort");
  // 2. pointer trie by sorted insertion (a matching child is always the last child)
  std::vector<Node> nodes(1);
  nodes.reserve(V * 4 + 16);
  uint32_t n_term = 0;
  for (uint32_t id : order) {
    const uint8_t* p = bytes + off[id];
    size_t l = off[id + 1] - off[id];
    uint32_t cur = 0;
    for (size_t d = 0; d < l; d++) {
      uint32_t lc = nodes[cur].last_child;
      if (lc && nodes[lc].label == p[d]) {
        cur = lc;
      } else {
        uint32_t nn = (uint32_t)nodes.size();
        nodes.emplace_back();
        nodes[nn].label = p[d];
        if (lc) nodes[lc].next_sibling = nn; else nodes[cur].first_child = nn;
        nodes[cur].last_child = nn;
        cur = nn;
      }
    }
    if (nodes[cur].term_id < 0) n_term++;
    nodes[cur].term_id = std::max(nodes[cur].term_id, (int32_t)id);  // last id wins
  }

  lap("trie");
  // 3. BFS order: shallow nodes get the lowest slots (the part staged in shared memory)
  std::vector<uint32_t> bfs;
  bfs.reserve(nodes.size());
  bfs.push_back(0);
  std::vector<uint8_t> depth(nodes.size(), 0);
  std::vector<uint32_t> parent(nodes.size(), 0);
  for (size_t i = 0; i < bfs.size(); i++)
    for (uint32_t c = nodes[bfs[i]].first_child; c; c = nodes[c].next_sibling) {
      depth[c] = (uint8_t)std::min<uint32_t>(depth[bfs[i]] + 1u, 255u);
      parent[c] = bfs[i];
      bfs.push_back(c);
    }

  lap("bfs");
  // 4. slot allocation.  A node needs a base b that no other node uses (trie_build.h) with every child slot
  //    b ^ label free; both stay inside b's 256-entry block.  Per block the admissible bases are one bitmap
  //    expression: unused_bases & AND_i permute(free_slots, label_i), so a block is tested in O(k) word operations
  //    (a walk over the free-slot list spent 85 % of its tries on bases that were taken: 150 ms for 230k nodes).
  std::vector<Block> blocks;
  std::vector<uint32_t> n_free;  // free slots per block
  auto add_block = [&]() {
    Block nb;
    for (int w = 0; w < 4; w++) nb.free_slots[w] = nb.free_bases[w] = ~0ull;
    blocks.push_back(nb);
    n_free.push_back(256);
  };
  add_block();
  blocks[0].free_slots[0] &= ~1ull;  // slot 0 = root
  n_free[0]--;
  nodes[0].slot = 0;
  std::vector<uint32_t> base_of(nodes.size(), 0);
  uint8_t labels[256];
  uint32_t first_open = 0;  // blocks below it have (next to) no free slot left
  // Allocation order: the root and its children first (BFS, so that hot[1] / hot[2] below hold), then every deeper
  // node by descending weight = sum of exp(score) over the tokens below it, i.e. roughly how often a walk reaches it:
  // a parent's children land in the lowest block that has room when the parent is processed, so the nodes walks
  // visit most sit at the front of the array — the part match_kernel stages in shared memory.
  std::vector<uint32_t> alloc_order;
  if (!hot_order) {
    alloc_order = bfs;
  } else {
    double smax = -INFINITY;
    for (uint64_t i = 0; i < V; i++) smax = std::max(smax, scores[i]);
    std::vector<float> weight(nodes.size(), 0.f);
    for (size_t i = bfs.size(); i-- > 1;) {  // children before parents
      const uint32_t nd = bfs[i];
      if (nodes[nd].term_id >= 0) weight[nd] += (float)std::exp(scores[nodes[nd].term_id] - smax);
      weight[parent[nd]] += weight[nd];
    }
    alloc_order.reserve(bfs.size());
    size_t shallow = 0;
    while (shallow < bfs.size() && depth[bfs[shallow]] < 2) alloc_order.push_back(bfs[shallow++]);
    // (weights are non-negative floats: their bit patterns order like the values; ties keep BFS order)
    std::vector<uint64_t> deep;
    deep.reserve(bfs.size() - shallow);
    for (size_t i = shallow; i < bfs.size(); i++)
      if (nodes[bfs[i]].first_child) {
        uint32_t wb;
        std::memcpy(&wb, &weight[bfs[i]], 4);
        deep.push_back(((uint64_t)(~wb) << 32) | (uint32_t)i);
      }
    std::sort(deep.begin(), deep.end());
    for (uint64_t k : deep) alloc_order.push_back(bfs[(uint32_t)k]);
  }
  lap("order");
  for (uint32_t nidx : alloc_order) {
    uint32_t k = 0;
    for (uint32_t c = nodes[nidx].first_child; c; c = nodes[c].next_sibling) labels[k++] = nodes[c].label;
    if (!k) continue;
    auto try_block = [&](uint32_t bi) -> uint32_t {  // an admissible base inside block bi, or END
      if (n_free[bi] < k) return 0xFFFFFFFFu;
      const Block& B = blocks[bi];
      uint64_t cand[4] = {B.free_bases[0], B.free_bases[1], B.free_bases[2], B.free_bases[3]};
      for (uint32_t i = 0; i < k; i++) {
        uint64_t pm[4];
        xor_permute(B.free_slots, labels[i], pm);
        uint64_t any = 0;
        for (int w = 0; w < 4; w++) any |= (cand[w] &= pm[w]);
        if (!any) return 0xFFFFFFFFu;
      }
      for (int w = 0; w < 4; w++)
        if (cand[w]) return (bi << 8) | (uint32_t)(w * 64 + __builtin_ctzll(cand[w]));
      return 0xFFFFFFFFu;
    };
    while (first_open < blocks.size() && n_free[first_open] < 4) first_open++;  // (gives up on <= 3 slots of 256)
    uint32_t base = 0xFFFFFFFFu;
    const uint32_t nb = (uint32_t)blocks.size();
    // Narrow nodes fill the holes wide nodes left behind: they look at the lowest open blocks first; every node
    // then looks at the most recent blocks only, so the search stays short.
    const uint32_t back = k >= 64 ? 1 : k >= 8 ? 2 : 4;
    const uint32_t recent = nb > back ? nb - back : 0;
    if (k <= 2)
      for (uint32_t bi = first_open; bi < std::min(nb, first_open + 2) && base == 0xFFFFFFFFu; bi++) base = try_block(bi);
    for (uint32_t bi = std::max(recent, first_open); bi < nb && base == 0xFFFFFFFFu; bi++) base = try_block(bi);
    if (base == 0xFFFFFFFFu) {
      // open a fresh block: any base inside it works (all slots free, no base used)
      if ((blocks.size() + 1) * 256 > MAX_SLOTS) return "trie too large (slot index must fit 23 bits)";
      base = (uint32_t)blocks.size() << 8;
      add_block();
    }
    Block& B = blocks[base >> 8];
    B.free_bases[(base & 255u) >> 6] &= ~(1ull << (base & 63u));
    base_of[nidx] = base;
    if (depth[nidx] < 2) {  // every probe out of this node stays inside base's 256-slot block
      uint32_t end = (base | 255u) + 1u;
      for (int d = depth[nidx] + 1; d <= 2; d++) out->hot[d] = std::max(out->hot[d], end);
    }
    for (uint32_t c = nodes[nidx].first_child; c; c = nodes[c].next_sibling) {
      const uint32_t sl = base ^ nodes[c].label;
      B.free_slots[(sl & 255u) >> 6] &= ~(1ull << (sl & 63u));
      n_free[base >> 8]--;
      nodes[c].slot = sl;
    }
  }
  const uint32_t n_slots = (uint32_t)blocks.size() * 256u;

  lap("alloc");
  // 5. emit slots
  out->slots.assign(n_slots, Slot{0, 0, 0, 0});
  for (size_t i = 0; i < nodes.size(); i++) {
    const Node& nd = nodes[i];
    Slot s{0, 0, 0, 0};
    // the root is never a transition target: it stays "empty" (bit 8 clear)
    s.x = ((base_of[i] ^ 0x100u) << 9) | (i == 0 ? 0u : SLOT_OCC) | nd.label;
    uint32_t flags = 0;
    if (nd.first_child) flags |= SLOT_HASCH;
    if (nd.term_id >= 0 && i != 0) {
      flags |= SLOT_TERM;
      s.y = (uint32_t)nd.term_id & SLOT_ID_MASK;
      uint64_t bits;
      std::memcpy(&bits, &scores[nd.term_id], 8);
      s.z = (uint32_t)bits;
      s.w = (uint32_t)(bits >> 32);
    }
    s.y |= flags;
    out->slots[nd.slot] = s;
  }
  lap("emit");
  // 6. what build_match_tables needs later (the match tables are built on first use: trie_build.h)
  out->slots8.clear();
  out->pair2.clear();
  out->rows.clear();
  out->row_ids.clear();
  out->node_parent.clear();
  if (max_len >= 1 && max_len <= 16) {
    out->node_parent = std::move(parent);
    out->node_depth = std::move(depth);
    out->node_slot.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); i++) out->node_slot[i] = nodes[i].slot;
  }
  lap("rows");
  out->root_base = base_of[0] ^ 0x100u;
  out->max_token_len = max_len;
  out->n_nodes = (uint32_t)nodes.size();
  out->n_terminals = n_term;
  for (int d = 1; d <= 2; d++) out->hot[d] = std::min<uint32_t>(std::max(out->hot[d], 256u), n_slots);
  return "";
}

std::string retarget_double_array(DoubleArray* da, const uint8_t* bytes, const uint64_t* off, const double* scores,
                                  uint64_t V) {
  if (da->slots.empty() || V > MAX_VOCAB) return "miss";
  const size_t n_slots = da->slots.size();
  const uint32_t depth_max = da->max_token_len;
  const size_t T = V < 2048 ? 1 : std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency()));
  // f(lo, hi) over T equal ranges of [0, n), one per host thread (n = T: T workers that share work themselves)
  auto over_threads = [T](size_t n, const std::function<void(size_t, size_t)>& f) {
    std::vector<std::thread> ts;
    for (size_t t = 1; t < T; t++) ts.emplace_back(f, n * t / T, n * (t + 1) / T);
    f(0, n / T);
    for (auto& x : ts) x.join();
  };
  // 1. every token walked to its slot (read-only: a failure leaves the array as it was).  Blocks of ids handed out to
  //    host threads.  "miss" sends the caller to build_double_array, which reports a non-finite score itself, so which
  //    of several failures is seen first does not matter.
  std::vector<uint32_t> slot_of(V, 0xFFFFFFFFu);
  std::atomic<uint64_t> next{0};
  std::atomic<int> failed{0};  // 1 = miss, 2 = non-finite score
  over_threads(T, [&](size_t, size_t) {  // T workers; the blocks of ids come from `next`
    const Slot* slots = da->slots.data();
    while (!failed.load(std::memory_order_relaxed)) {
      const uint64_t lo = next.fetch_add(4096);
      if (lo >= V) break;
      const uint64_t hi = std::min<uint64_t>(V, lo + 4096);
      for (uint64_t i = lo; i < hi; i++) {
        const uint64_t len = off[i + 1] - off[i];
        int kind = 0;
        if (!std::isfinite(scores[i])) {
          kind = 2;
        } else if (len == 0) {
          continue;  // the empty token never matches
        } else if (len > depth_max) {
          kind = 1;
        } else {
          const uint8_t* tk = bytes + off[i];
          uint32_t xbase = da->root_base, at = 0;
          for (uint64_t d = 0; d < len && !kind; d++) {
            const uint32_t cw = 0x100u | tk[d];
            at = xbase ^ cw;
            if (at >= n_slots) {
              kind = 1;
              break;
            }
            const Slot& e = slots[at];
            if (((e.x ^ cw) & 0x1FFu) || (d + 1 < len && !(e.y & SLOT_HASCH))) kind = 1;
            xbase = e.x >> 9;
          }
          if (!kind) slot_of[i] = at;
        }
        if (kind) {
          int none = 0;
          if (kind == 2) failed.store(2);
          else failed.compare_exchange_strong(none, 1);
          return;
        }
      }
    }
  });
  if (failed.load()) return failed.load() == 2 ? "non-finite token score" : "miss";
  // 2. nothing can fail from here on: terminals / ids / scores rewritten in place
  da->slots8.clear();
  da->pair2.clear();
  da->rows.clear();
  da->row_ids.clear();
  over_threads(n_slots, [&](size_t lo, size_t hi) {
    Slot* slots = da->slots.data();
    for (size_t k = lo; k < hi; k++) {
      slots[k].y &= ~(SLOT_TERM | SLOT_ID_MASK);
      slots[k].z = slots[k].w = 0;
    }
  });
  uint32_t n_term = 0;
  for (uint64_t i = 0; i < V; i++) {  // in id order: the last of equal byte strings wins (src/trie.rs:19)
    if (slot_of[i] == 0xFFFFFFFFu) continue;
    Slot& s = da->slots[slot_of[i]];
    if (!(s.y & SLOT_TERM)) n_term++;
    s.y = (s.y & ~SLOT_ID_MASK) | SLOT_TERM | ((uint32_t)i & SLOT_ID_MASK);
    uint64_t bits;
    std::memcpy(&bits, &scores[i], 8);
    s.z = (uint32_t)bits;
    s.w = (uint32_t)(bits >> 32);
  }
  // (max_token_len stays the depth of the ARRAY: the kernels size their windows by it)
  da->n_terminals = n_term;
  return "";
}

std::string build_match_tables(DoubleArray* out) {
  if (!out->slots8.empty()) return "";
  if (out->node_parent.empty()) return "no match tables for this vocabulary (tokens longer than 16 bytes)";
  const std::vector<uint32_t>& parent = out->node_parent;
  const std::vector<uint8_t>& depth = out->node_depth;
  const size_t n_nodes = parent.size();
  const size_t n_slots = out->slots.size();
  // the token that ends at a node and its score are what the node's slot says (node 0, the root, ends none)
  auto term_of = [&](uint32_t nd) -> int32_t {
    const Slot& sl = out->slots[out->node_slot[nd]];
    return nd != 0 && (sl.y & SLOT_TERM) ? (int32_t)(sl.y & SLOT_ID_MASK) : -1;
  };
  auto score_of = [&](uint32_t nd) -> double {
    const Slot& sl = out->slots[out->node_slot[nd]];
    const uint64_t bits = (uint64_t)sl.z | ((uint64_t)sl.w << 32);
    double v;
    std::memcpy(&v, &bits, 8);
    return v;
  };
  // terminal nodes in id order (ids are unique per terminal: the last duplicate owns the node) ...
  std::vector<std::pair<int32_t, uint32_t>> byid;
  byid.reserve(out->n_terminals);
  for (uint32_t i = 1; i < n_nodes; i++)
    if (term_of(i) >= 0) byid.emplace_back(term_of(i), i);
  std::sort(byid.begin(), byid.end());
  std::vector<uint32_t> terms;
  terms.reserve(byid.size());
  for (auto& pr : byid) terms.push_back(pr.second);
  // ... which is already hottest-first for a score-sorted vocabulary (the usual case); otherwise sort
  auto hotter = [&](uint32_t a, uint32_t b) { return score_of(a) > score_of(b); };
  if (!std::is_sorted(terms.begin(), terms.end(), hotter)) std::stable_sort(terms.begin(), terms.end(), hotter);
  const double ninf = -INFINITY;
  std::vector<uint32_t> row_of(n_nodes, 0);
  uint64_t units16 = 9;  // row 0: header + 16 x -inf (+ pad)
  for (uint32_t nd : terms) {
    row_of[nd] = (uint32_t)units16;
    units16 += (depth[nd] + 2u) / 2u;  // header + depth scores, padded to 16 bytes
  }
  if (units16 > SLOT8_OFF_MASK) return "row table too large";
  out->rows.assign(units16 * 2, ninf);
  out->row_ids.assign(units16 * 2, 0xFFFFFFFFu);
  const uint64_t zero_bits = 0;
  std::memcpy(&out->rows[0], &zero_bits, 8);
  auto fill = [&](size_t lo, size_t hi) {
    for (size_t t = lo; t < hi; t++) {
      const uint32_t nd = terms[t];
      const size_t base = (size_t)row_of[nd] * 2;
      uint64_t mask = 0;
      for (uint32_t a = nd; a; a = parent[a])
        if (term_of(a) >= 0) {
          out->rows[base + depth[a]] = score_of(a);
          out->row_ids[base + depth[a]] = (uint32_t)term_of(a);
          mask |= 1ull << (depth[a] - 1);
        }
      std::memcpy(&out->rows[base], &mask, 8);
      out->row_ids[base] = (uint32_t)mask;
    }
  };
  {
    const size_t T = terms.size() >= 65536 ? 4 : 1;  // rows are disjoint: a few host threads
    std::vector<std::thread> th;
    for (size_t k = 1; k < T; k++) th.emplace_back(fill, terms.size() * k / T, terms.size() * (k + 1) / T);
    fill(0, terms.size() / T);
    for (auto& x : th) x.join();
  }
  out->slots8.assign(n_slots, 0);
  for (size_t i = 0; i < n_nodes; i++) {
    const Slot& sl = out->slots[out->node_slot[i]];
    uint32_t y = 0;
    if (sl.y & SLOT_HASCH) y |= SLOT8_HASCH;
    if (term_of((uint32_t)i) >= 0) y |= SLOT8_TERM | row_of[i];
    out->slots8[out->node_slot[i]] = (uint64_t)sl.x | ((uint64_t)y << 32);
  }
  // the two first levels of every walk, tabulated (the probes of match_kernel, restated)
  out->pair2.assign(65536, 0);
  for (uint32_t b0 = 0; b0 < 256; b0++) {
    uint32_t xb1 = 0, best1 = 15u << 28;
    bool go1 = false;
    {
      const uint32_t cw = 0x100u | b0, t = out->root_base ^ cw;
      if (t < n_slots) {
        const uint32_t ex = (uint32_t)out->slots8[t], ey = (uint32_t)(out->slots8[t] >> 32);
        const bool hit = ((ex ^ cw) & 0x1FFu) == 0;
        if (hit && (ey & SLOT8_TERM)) best1 = (0u << 28) | (ey & SLOT8_OFF_MASK);
        go1 = hit && (ey & SLOT8_HASCH);
        xb1 = ex >> 9;
      }
    }
    for (uint32_t b1 = 0; b1 < 256; b1++) {
      uint32_t xb = xb1, best = best1;
      bool go = go1;
      if (go) {
        const uint32_t cw = 0x100u | b1, t = xb ^ cw;
        go = false;
        if (t < n_slots) {
          const uint32_t ex = (uint32_t)out->slots8[t], ey = (uint32_t)(out->slots8[t] >> 32);
          const bool hit = ((ex ^ cw) & 0x1FFu) == 0;
          if (hit && (ey & SLOT8_TERM)) best = (1u << 28) | (ey & SLOT8_OFF_MASK);
          go = hit && (ey & SLOT8_HASCH);
          xb = ex >> 9;
        }
      }
      out->pair2[b0 | (b1 << 8)] = (uint64_t)(go ? (xb | (1u << 31)) : 0u) | ((uint64_t)best << 32);
    }
  }
  return "";
}

std::string build_token_hash(const uint8_t* bytes, const uint64_t* off, uint64_t V, TokenHash* out) {
  out->slots.clear();
  out->mask = 0;
  uint64_t n = 0;
  for (uint64_t i = 0; i < V; i++) {
    uint64_t len = off[i + 1] - off[i];
    if (len >= 1 && len <= 16) n++;
  }
  uint64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  if (cap > (1ull << 30)) return "vocabulary too large for the token hash";
  std::vector<std::pair<uint64_t, uint64_t>> lohi(V);
  for (uint64_t i = 0; i < V; i++) {
    uint64_t len = off[i + 1] - off[i], lo = 0, hi = 0;
    if (len >= 1 && len <= 16) {
      for (uint64_t k = 0; k < len; k++) {
        uint64_t b = bytes[off[i] + k];
        if (k < 8) lo |= b << (8 * k); else hi |= b << (8 * (k - 8));
      }
    }
    lohi[i] = {lo, hi};
  }
  for (uint64_t attempt = 0; attempt < 64; attempt++) {
    const uint64_t seed = 0x243F6A8885A308D3ull + attempt * 0x13198A2E03707344ull;
    std::vector<Slot> slots(cap, Slot{0, 0, 0, 0});
    const uint32_t mask = (uint32_t)(cap - 1);
    // ids of a key -> the bytes that own it (to tell a duplicate token from a key collision)
    bool collision = false;
    for (uint64_t i = 0; i < V && !collision; i++) {
      uint32_t len = (uint32_t)(off[i + 1] - off[i]);
      if (len < 1 || len > 16) continue;
      const uint64_t key = token_key(lohi[i].first, lohi[i].second, len, seed);
      uint32_t s = token_key_slot(key, mask);
      for (;;) {
        Slot& e = slots[s];
        const uint64_t k = ((uint64_t)e.y << 32) | e.x;
        if (k == 0) {
          e.x = (uint32_t)key; e.y = (uint32_t)(key >> 32); e.z = (uint32_t)i;
          break;
        }
        if (k == key) {
          const uint64_t j = e.z;  // same key: must be the same bytes (a duplicate token: last id wins)
          if (off[j + 1] - off[j] == len && lohi[j] == lohi[i]) { e.z = (uint32_t)i; break; }
          collision = true;
          break;
        }
        s = (s + 1) & mask;
      }
    }
    if (!collision) {
      out->slots.swap(slots);
      out->mask = mask;
      out->seed = seed;
      return "";
    }
  }
  return "token hash: could not find a collision-free seed";
}

}  // namespace tgx
