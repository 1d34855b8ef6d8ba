// Device self-check of tgx_model_rebuild in C++ alone (no Python start-up: runs in a second on a GPU box):
// a rebuilt model — layout kept for a subset (src/prune.rs:48,53 hands over subsets), built afresh below option 45's
// share — must encode exactly like a model created from the same vocabulary.
//   g++ -O2 -std=c++17 -I include -o tools/rebuild_check tools/rebuild_check.cpp -ldl && tools/rebuild_check
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "tokengeex_b200.h"

#define SYM(name) auto p_##name = reinterpret_cast<decltype(&name)>(dlsym(lib, #name)); if (!p_##name) { printf("missing %s\n", #name); return 2; }

struct Vocab {
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> off{0};
  std::vector<double> scores;
  void add(const std::string& t, double s) {
    bytes.insert(bytes.end(), t.begin(), t.end());
    off.push_back(bytes.size());
    scores.push_back(s);
  }
};

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "tokengeex_b200/csrc/libtokengeex_b200.so";
  void* lib = dlopen(path, RTLD_NOW);
  if (!lib) { printf("dlopen: %s\n", dlerror()); return 2; }
  SYM(tgx_model_create) SYM(tgx_model_rebuild) SYM(tgx_model_destroy) SYM(tgx_encode_batch) SYM(tgx_last_error)
  std::mt19937_64 rng(7);
  const std::string alpha = "abcdefgh";
  std::set<std::string> uniq;
  for (char c : alpha) uniq.insert(std::string(1, c));
  while (uniq.size() < 6000) {
    std::string t;
    const int n = 1 + (int)(rng() % 12);
    for (int i = 0; i < n; i++) t += alpha[rng() % alpha.size()];
    uniq.insert(t);
  }
  std::vector<std::string> toks(uniq.begin(), uniq.end());
  std::vector<double> sc(toks.size());
  for (auto& s : sc) s = -(0.1 + (double)(rng() % 80000) / 10000.0);
  std::vector<uint8_t> text;
  std::vector<uint64_t> off{0};
  for (int s = 0; s < 300; s++) {
    const int n = (int)(rng() % 4000);
    for (int i = 0; i < n; i++) text.push_back((uint8_t)alpha[rng() % alpha.size()]);
    off.push_back(text.size());
  }
  const uint64_t S = off.size() - 1;
  auto encode = [&](tgx_model* m, std::vector<uint32_t>& ids, std::vector<uint64_t>& id_off) {
    ids.assign(text.size() + 4, 0);
    id_off.assign(S + 1, 0);
    int64_t bad = -1;
    const int rc = p_tgx_encode_batch(m, text.data(), off.data(), S, 0, ids.data(), ids.size(), id_off.data(), nullptr, nullptr, &bad);
    ids.resize(id_off[S]);
    return rc;
  };
  auto make = [&](const std::vector<size_t>& keep, double shift) {
    Vocab v;
    for (size_t i : keep) v.add(toks[i], sc[i] + shift);
    return v;
  };
  std::vector<size_t> all(toks.size());
  for (size_t i = 0; i < all.size(); i++) all[i] = i;
  Vocab v0 = make(all, 0.0);
  tgx_model* m = nullptr;
  if (p_tgx_model_create(v0.bytes.data(), v0.off.data(), v0.scores.data(), v0.scores.size(), 0, &m)) { printf("create: %s\n", p_tgx_last_error()); return 1; }
  int fails = 0;
  std::vector<size_t> cur = all;
  for (double frac : {1.0, 0.8, 0.7, 0.3}) {  // 1.0 / 0.8 / 0.56 of the built size keep the layout, 0.17 builds afresh
    std::vector<size_t> nx;
    for (size_t i : cur)
      if (toks[i].size() == 1 || (double)(rng() % 1000) < frac * 1000.0) nx.push_back(i);
    Vocab v = make(nx, -0.5 * frac);
    if (p_tgx_model_rebuild(m, v.bytes.data(), v.off.data(), v.scores.data(), v.scores.size())) { printf("rebuild: %s\n", p_tgx_last_error()); return 1; }
    tgx_model* f = nullptr;
    if (p_tgx_model_create(v.bytes.data(), v.off.data(), v.scores.data(), v.scores.size(), 0, &f)) { printf("create: %s\n", p_tgx_last_error()); return 1; }
    std::vector<uint32_t> a, b;
    std::vector<uint64_t> ao, bo;
    const int ra = encode(m, a, ao), rb = encode(f, b, bo);
    const bool same = ra == rb && a == b && ao == bo;
    printf("rebuild to %zu of %zu tokens: rc %d/%d, %zu ids, %s\n", nx.size(), toks.size(), ra, rb, a.size(), same ? "same as a fresh model" : "DIFFERENT");
    fails += same ? 0 : 1;
    p_tgx_model_destroy(f);
    cur = nx;
  }
  p_tgx_model_destroy(m);
  printf(fails ? "FAILED\n" : "rebuild_check ok\n");
  return fails ? 1 : 0;
}
