// XOR double-array construction — see trie_build.h.
#include "trie_build.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

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

// Doubly linked list of free slots, kept in increasing order.
struct FreeList {
  std::vector<uint32_t> prev, next;  // valid for free slots; index n_slots is the sentinel
  std::vector<uint8_t> used;
  std::vector<uint8_t> base_used;
  uint32_t head = 0;  // first free slot or END
  static constexpr uint32_t END = 0xFFFFFFFFu;
  uint32_t tail = END;

  uint32_t size() const { return (uint32_t)used.size(); }

  void add_block() {
    uint32_t lo = size();
    used.resize(lo + 256, 0);
    base_used.resize(lo + 256, 0);
    prev.resize(lo + 256);
    next.resize(lo + 256);
    for (uint32_t s = lo; s < lo + 256; s++) {
      prev[s] = (s == lo) ? tail : s - 1;
      next[s] = (s == lo + 255) ? END : s + 1;
    }
    if (tail != END) next[tail] = lo; else head = lo;
    tail = lo + 255;
  }

  void take(uint32_t s) {
    uint32_t p = prev[s], n = next[s];
    if (p != END) next[p] = n; else head = n;
    if (n != END) prev[n] = p; else tail = p;
    used[s] = 1;
  }
};

}  // namespace

std::string build_double_array(const uint8_t* bytes, const uint64_t* off, const double* scores,
                               uint64_t V, DoubleArray* out) {
  if (V >= MAX_VOCAB) return "vocabulary too large (ids must fit 24 bits)";
  uint32_t max_len = 0;
  for (uint64_t i = 0; i < V; i++) {
    uint64_t l = off[i + 1] - off[i];
    if (l > MAX_TOKEN_LEN) return "token longer than 64 bytes is not supported by the device kernels";
    if (!std::isfinite(scores[i])) return "non-finite token score";
    max_len = std::max<uint32_t>(max_len, (uint32_t)l);
  }

  // 1. sort token indices by bytes (ties: id ascending, so the last duplicate is seen last)
  std::vector<uint32_t> order;
  order.reserve(V);
  for (uint64_t i = 0; i < V; i++)
    if (off[i + 1] > off[i]) order.push_back((uint32_t)i);  // the empty token never matches
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    size_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b];
    int c = std::memcmp(bytes + off[a], bytes + off[b], std::min(la, lb));
    if (c) return c < 0;
    if (la != lb) return la < lb;
    return a < b;
  });

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

  // 3. BFS order: shallow nodes get the lowest slots (the part staged in shared memory)
  std::vector<uint32_t> bfs;
  bfs.reserve(nodes.size());
  bfs.push_back(0);
  std::vector<uint8_t> depth(nodes.size(), 0);
  for (size_t i = 0; i < bfs.size(); i++)
    for (uint32_t c = nodes[bfs[i]].first_child; c; c = nodes[c].next_sibling) {
      depth[c] = (uint8_t)std::min<uint32_t>(depth[bfs[i]] + 1u, 255u);
      bfs.push_back(c);
    }

  // 4. slot allocation
  FreeList fl;
  fl.add_block();
  fl.take(0);  // root
  nodes[0].slot = 0;
  std::vector<uint32_t> base_of(nodes.size(), 0);
  uint8_t labels[256];
  uint32_t hint[3] = {0, 0, 0};
  for (uint32_t nidx : bfs) {
    uint32_t k = 0;
    for (uint32_t c = nodes[nidx].first_child; c; c = nodes[c].next_sibling) labels[k++] = nodes[c].label;
    if (!k) continue;
    uint32_t base = FreeList::END;
    // Candidates: free slots f taken as the home of the first label.  Narrow nodes first
    // try the global head of the free list (they fill the holes wide nodes left behind);
    // then every node scans the recent blocks only, so the search stays short.
    auto scan = [&](uint32_t f, uint32_t max_tries) {
      uint32_t tries = 0;
      for (; f != FreeList::END && tries < max_tries; f = fl.next[f], tries++) {
        uint32_t b = f ^ labels[0];
        if (fl.base_used[b]) continue;
        bool ok = true;
        for (uint32_t i = 1; i < k; i++)
          if (fl.used[b ^ labels[i]]) { ok = false; break; }
        if (ok) return b;
      }
      return FreeList::END;
    };
    if (k <= 2) base = scan(fl.head, 64);
    if (base == FreeList::END) {
      int cls = k >= 64 ? 0 : k >= 8 ? 1 : 2;
      static const uint32_t BACK[3] = {2, 16, 64};
      uint32_t nblocks = fl.size() >> 8;
      uint32_t start = nblocks > BACK[cls] ? (nblocks - BACK[cls]) << 8 : 0;
      uint32_t& h = hint[cls];
      if (h < start) h = start;
      while (h < fl.size() && fl.used[h]) h++;  // used slots never become free: monotone
      if (h < fl.size()) base = scan(h, 2048);
    }
    if (base == FreeList::END) {
      // open a fresh block: any base inside it works (all slots free, no base used)
      uint32_t lo = fl.size();
      if (lo + 256 > MAX_SLOTS) return "trie too large (slot index must fit 23 bits)";
      fl.add_block();
      base = lo;
    }
    fl.base_used[base] = 1;
    base_of[nidx] = base;
    if (depth[nidx] < 2) {  // every probe out of this node stays inside base's 256-slot block
      uint32_t end = (base | 255u) + 1u;
      for (int d = depth[nidx] + 1; d <= 2; d++) out->hot[d] = std::max(out->hot[d], end);
    }
    for (uint32_t c = nodes[nidx].first_child; c; c = nodes[c].next_sibling) {
      uint32_t s = base ^ nodes[c].label;
      fl.take(s);
      nodes[c].slot = s;
    }
  }

  // 5. emit slots
  out->slots.assign(fl.size(), Slot{0, 0, 0, 0});
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
  out->root_base = base_of[0] ^ 0x100u;
  out->max_token_len = max_len;
  out->n_nodes = (uint32_t)nodes.size();
  out->n_terminals = n_term;
  for (int d = 1; d <= 2; d++) out->hot[d] = std::min<uint32_t>(std::max(out->hot[d], 256u), (uint32_t)fl.size());
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
