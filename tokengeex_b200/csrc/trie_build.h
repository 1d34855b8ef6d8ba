// Host-side construction of the device vocabulary structure.
//
// Replaces /root/reference/src/trie.rs (pointer trie, one FNV HashMap per node) and
// the trie half of Model::from (/root/reference/src/model.rs:16-30) with a single
// XOR double-array laid out for the GPU: one 16-byte slot per trie node, so that one
// 128-bit load performs the transition, the terminal test and the score fetch.
//
//   slot.x = (xbase << 9) | 0x100 | label   occupied slots carry bit 8; an empty slot is all zero.
//                                      xbase = base ^ 0x100: with cw = 0x100 | byte the children of
//                                      this node live at xbase ^ cw (= base ^ byte)
//   slot.y = id | flags << 24         id: token id (24 bit), flags: TERM | HASCH
//   slot.zw = f64 score bits          score of vocab[id] (valid when TERM)
//
// A transition s --c--> t is valid iff t = xbase(s) ^ cw satisfies (slot[t].x ^ cw) & 0x1FF == 0,
// i.e. t is occupied and label(t) == c: ONE masked compare on the word that also carries the
// next base.  Bases are unique per parent, which makes the one-byte label a sufficient check
// (two parents reaching the same slot with the same label would share a base).
// Semantics kept from the reference: duplicate byte strings keep the LAST id
// (src/trie.rs:19); the empty token is never matched (src/trie.rs:51-63).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tgx {

constexpr uint32_t SLOT_OCC = 0x100u;         // in slot.x
constexpr uint32_t SLOT_TERM = 2u << 24;      // in slot.y
constexpr uint32_t SLOT_HASCH = 4u << 24;
constexpr uint32_t SLOT_ID_MASK = 0x00FFFFFFu;
constexpr uint32_t MAX_TOKEN_LEN = 64;        // rows of the per-lane match buffer
constexpr uint32_t MAX_VOCAB = 1u << 24;      // 24-bit ids inside slots / back-pointers
constexpr uint32_t MAX_SLOTS = 1u << 23;      // 23-bit bases

struct Slot {
  uint32_t x, y, z, w;
};

struct DoubleArray {
  std::vector<Slot> slots;   // slot 0 is the root (never a transition target)
  uint32_t root_base = 0;    // xbase of the root (see above)
  uint32_t max_token_len = 0;
  uint32_t n_nodes = 0;      // trie nodes incl. root
  uint32_t n_terminals = 0;  // distinct token byte strings
  // Slots are handed out in BFS order, so shallow nodes sit at the front: every transition out of
  // a node of depth < d (hit or miss: base ^ byte stays inside base's 256-slot block) lands in
  // [0, hot[d]).  hot[1] covers the root's children, hot[2] also their children.  The kernels
  // stage that prefix in shared memory.
  uint32_t hot[3] = {0, 0, 0};

  // ---- match tables (max_token_len <= 16 only; empty otherwise): what match_kernel / viterbi_rows_kernel read.
  // slots8[t] = slots[t].x | y8 << 32 with y8 = TERM8 | HASCH8 | row offset: the 8-byte form of the same double-array
  // (transition check + flags + the ROW of the token that ends at this node), half the footprint of `slots`.
  // A row lists, dense by length, the score of every vocabulary token that is a prefix of the row's token
  // (= everything common_prefix_search yields on the way down to it, src/trie.rs:51-63).  With off = 2 * (row offset in
  // 16-byte units): rows[off] = header (as integer bits: bit l - 1 set iff the prefix of length l is a token),
  // rows[off + l] = score of the prefix of length l, or -inf when that prefix is not a token; padded with -inf to a
  // multiple of 16 bytes.  row_ids has the same layout with the token ids (0xFFFFFFFF = no token; [off] = the mask).
  // Rows are laid out by descending score of their own token (frequent tokens first), so the first bytes of the table
  // are the hot ones.  Row 0 = empty mask + 16 x -inf = "no token starts here".
  // Built on first use (build_match_tables) from the node arrays below, which build_double_array keeps: the EM loop
  // rebuilds the model ~30 times and only some of its passes read the tables.
  std::vector<uint64_t> slots8;
  // pair2[b0 | b1 << 8]: the walk's state after the two bytes b0 b1 — low word = base of the depth-2 node | 1 << 31 when
  // the walk goes on, 0 when it has ended; high word = the record (match_kernel's format) of the deepest token so far.
  // match2_kernel starts its walks here: one 8-byte load instead of the two first probes.
  std::vector<uint64_t> pair2;
  std::vector<double> rows;
  std::vector<uint32_t> row_ids;
  // per trie node (node 0 = root); empty when max_token_len > 16.  (The token that ends at a node, and its score, are in
  // the node's slot: a retargeted array needs no second copy of them.)
  std::vector<uint32_t> node_parent, node_slot;
  std::vector<uint8_t> node_depth;
};
// Fills slots8 / rows / row_ids (no-op when they are there).  Returns "" on success.
std::string build_match_tables(DoubleArray* da);
constexpr uint32_t SLOT8_TERM = 1u << 31, SLOT8_HASCH = 1u << 30, SLOT8_OFF_MASK = 0x0FFFFFFFu;

// Returns "" on success, else an error message.
// hot_order: slots beyond the first two levels are handed out by descending walk frequency instead of BFS order (a
// few per cent fewer L1 misses in the kernels, ~30 ms more per 250k tokens here: on for models that are built once,
// off for the rebuilds of the EM loop).
std::string build_double_array(const uint8_t* token_bytes, const uint64_t* token_offsets,
                               const double* scores, uint64_t vocab_size, DoubleArray* out, bool hot_order = true);

// The EM loop's rebuilds (src/prune.rs:48,53: `*model = Model::from(vocab)` after every M-step and every prune step)
// hand over a SUBSET of the vocabulary the array was built for.  This keeps the array's layout — every token of the
// new vocabulary is walked to its slot (host threads), terminal flags / ids / scores are rewritten IN PLACE, the nodes
// of the tokens that went stay as non-terminal nodes — instead of sorting, building and packing again.
// Same matches, ids and scores as a fresh build: the kernels only ever yield terminals.  Returns "" on success, "miss"
// when some token has no node in the array (not a subset: build afresh), else an error message; the array is
// untouched unless "" is returned.
std::string retarget_double_array(DoubleArray* da, const uint8_t* token_bytes, const uint64_t* token_offsets,
                                  const double* scores, uint64_t vocab_size);

// ---- token hash: bytes of a vocabulary token (1..16 bytes) -> id, ONE probe instead of one trie probe per byte.
// Used by the emit kernel, which only ever looks up strings that ARE vocabulary tokens (the marked tokens of the
// best path), so a 64-bit key stands for the bytes: the builder re-seeds until no two distinct tokens share a key.
// Same duplicate rule as the trie: the LAST id of equal byte strings wins (src/trie.rs:19).
#if defined(__CUDACC__)
#define TGX_HD __host__ __device__
#else
#define TGX_HD
#endif
// lo = bytes 0..7, hi = bytes 8..15 (little endian, zero beyond len); never returns 0 (0 = empty slot)
TGX_HD inline uint64_t token_key(uint64_t lo, uint64_t hi, uint32_t len, uint64_t seed) {
  uint64_t h = (lo ^ seed) * 0x9E3779B97F4A7C15ull;
  h ^= h >> 31;
  h += (hi ^ (seed >> 7)) * 0xC2B2AE3D27D4EB4Full + (uint64_t)len * 0xD6E8FEB86659FD93ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h | 1ull;
}
TGX_HD inline uint32_t token_key_slot(uint64_t key, uint32_t mask) { return (uint32_t)(key >> 24) & mask; }

struct TokenHash {
  std::vector<Slot> slots;  // x,y = key (0 = empty), z = id; open addressing, linear probing
  uint32_t mask = 0;        // slots.size() - 1 (power of two); 0 = not built
  uint64_t seed = 0;
};
// Builds the table over every token of 1..16 bytes.  Returns "" on success.
std::string build_token_hash(const uint8_t* token_bytes, const uint64_t* token_offsets, uint64_t vocab_size,
                             TokenHash* out);

// prune_vocab (src/prune.rs:173-319) minus its frequency pass, over a double-array that already holds the vocabulary
// (prune_host.cpp; see tgx_prune_select in include/tokengeex_b200.h for the arguments).
int prune_select_with(const DoubleArray& da, const uint8_t* token_bytes, const uint64_t* token_offsets,
                      const double* scores, const uint8_t* keep, uint64_t V, const uint64_t* freq, uint64_t n_samples,
                      uint64_t target_vocab_size, double shrink_factor, int threads, uint32_t* out_ids, uint64_t* out_n,
                      double* audit);

// Host walk (used by Tokenizer::common_prefix_search, src/model.rs:132-138).
template <class F>
inline void da_common_prefix_search(const DoubleArray& da, const uint8_t* s, size_t n, F&& f) {
  uint32_t xbase = da.root_base;
  for (size_t d = 0; d < n; d++) {
    uint32_t cw = 0x100u | s[d];
    uint32_t t = xbase ^ cw;
    if (t >= da.slots.size()) return;
    const Slot& e = da.slots[t];
    if ((e.x ^ cw) & 0x1FFu) return;
    if (e.y & SLOT_TERM) f(e.y & SLOT_ID_MASK, (uint32_t)d + 1);
    if (!(e.y & SLOT_HASCH)) return;
    xbase = e.x >> 9;
  }
}

}  // namespace tgx
