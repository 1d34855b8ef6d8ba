// Synthetic corpora and initial vocabularies for tests and benchmarks (host only).
//
// There is no network for datasets, so the workloads BASELINE.json names are
// synthesised deterministically: NUL-free valid-UTF-8 samples whose
// lengths follow lognormal(median 3000 B, sigma 1.2) clipped to [16, 262144] (bounds
// from /root/reference/scripts/datagen.py:100), drawn from Zipf-distributed identifier
// / keyword / operator inventories, with 5 % CRLF samples and optional CJK prose.
// The vocabulary builder follows /root/reference/src/generate.rs:54-243 (random
// substrings admitted by the "exact" allow-rule of data/exact.regex, per-sample
// dedup, frequency*len scores, bytes 0..254 kept, ln-probabilities) with a seeded
// PRNG instead of thread_rng so that every box builds the same vocabulary.
//
// PRNG: splitmix64-seeded xoshiro256**; every sample owns a stream keyed by
// (seed, sample index), so the output does not depend on the thread count.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed, uint64_t stream = 0) {
    uint64_t x = seed * 0xD1342543DE82EF95ULL + stream * 0x2545F4914F6CDD1DULL + 0x1234567ULL;
    for (auto& v : s) v = splitmix(x);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
  double normal() {
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

// Walker alias table for O(1) Zipf draws.
struct Alias {
  std::vector<double> prob;
  std::vector<uint32_t> alias;
  void build(const std::vector<double>& w) {
    size_t n = w.size();
    prob.assign(n, 0.0);
    alias.assign(n, 0);
    double sum = 0;
    for (double x : w) sum += x;
    std::vector<double> p(n);
    std::vector<uint32_t> small, large;
    for (size_t i = 0; i < n; i++) {
      p[i] = w[i] * n / sum;
      (p[i] < 1.0 ? small : large).push_back((uint32_t)i);
    }
    while (!small.empty() && !large.empty()) {
      uint32_t s = small.back(), l = large.back();
      small.pop_back();
      prob[s] = p[s];
      alias[s] = l;
      p[l] = p[l] + p[s] - 1.0;
      if (p[l] < 1.0) { large.pop_back(); small.push_back(l); }
    }
    for (uint32_t i : large) prob[i] = 1.0;
    for (uint32_t i : small) prob[i] = 1.0;
  }
  uint32_t draw(Rng& r) const {
    uint32_t i = r.below((uint32_t)prob.size());
    return r.uniform() < prob[i] ? i : alias[i];
  }
};

// Keyword / type / operator inventories (the kinds of lexemes the reference's
// allow-regexes are written for: /root/reference/src/regex.rs:82-172).
const char* const KW[8][24] = {
    {"def", "return", "if", "else", "elif", "for", "in", "while", "import", "from", "class", "self", "None", "True", "False", "with", "as", "try", "except", "lambda", "yield", "pass", "raise", "not"},
    {"function", "return", "if", "else", "for", "const", "let", "var", "new", "this", "class", "import", "export", "from", "async", "await", "null", "undefined", "true", "false", "typeof", "switch", "case", "break"},
    {"public", "private", "static", "void", "int", "String", "return", "if", "else", "for", "new", "class", "final", "import", "package", "this", "null", "boolean", "throws", "try", "catch", "extends", "implements", "long"},
    {"fn", "let", "mut", "pub", "impl", "struct", "enum", "match", "use", "mod", "self", "Self", "return", "if", "else", "for", "in", "while", "loop", "Some", "None", "Ok", "Err", "usize"},
    {"int", "char", "void", "return", "if", "else", "for", "while", "struct", "static", "const", "unsigned", "sizeof", "typedef", "include", "define", "NULL", "double", "float", "long", "switch", "case", "break", "continue"},
    {"func", "return", "if", "else", "for", "range", "var", "type", "struct", "interface", "package", "import", "nil", "err", "string", "int", "map", "chan", "go", "defer", "switch", "case", "default", "bool"},
    {"template", "typename", "class", "public", "private", "virtual", "const", "auto", "return", "if", "else", "for", "while", "namespace", "using", "std", "nullptr", "void", "int", "bool", "static", "inline", "override", "size_t"},
    {"SELECT", "FROM", "WHERE", "AND", "OR", "INSERT", "INTO", "VALUES", "UPDATE", "SET", "DELETE", "JOIN", "ON", "GROUP", "BY", "ORDER", "LIMIT", "NULL", "NOT", "AS", "CREATE", "TABLE", "INDEX", "IN"}};
const char* const OPS[] = {" = ", " == ", " != ", " + ", " - ", " * ", " / ", " < ", " > ", " <= ", " >= ", " && ", " || ", " += ", " -= ", " => ", " -> ", "::", ".", ", ", ": ", "++", "--", " % ", " & ", " | ", " << ", " >> ", "!", " := ", "...", "?"};
const char LETTER_FREQ[] = "eeeeeeeeeeeetttttttttaaaaaaaaoooooooiiiiiiinnnnnnnsssssshhhhhhrrrrrrddddllllcccuuummmwwffggyyppbbvkjxqz";

struct Lexicon {
  std::vector<std::string> idents;
  Alias ident_alias, cjk_alias;
  Lexicon(uint64_t seed) {
    Rng r(seed, 0xABCDEF);
    const int N = 50000;
    idents.reserve(N);
    std::vector<double> w(N);
    size_t nl = sizeof(LETTER_FREQ) - 1;
    for (int i = 0; i < N; i++) {
      int parts = 1 + (r.below(100) < 45) + (r.below(100) < 15);
      int style = r.below(10);  // 0-5 lower/snake, 6-8 camel, 9 UPPER
      std::string s;
      for (int p = 0; p < parts; p++) {
        int len = 2 + r.below(7);
        if (parts == 1) len = 2 + r.below(11);
        std::string part;
        for (int k = 0; k < len; k++) part.push_back(LETTER_FREQ[r.below((uint32_t)nl)]);
        if (style >= 6 && style <= 8) {
          if (p > 0 || r.below(2)) part[0] = (char)(part[0] - 32);
        } else if (style == 9) {
          for (auto& c : part) c = (char)(c - 32);
          if (p > 0) s.push_back('_');
        } else if (p > 0) {
          s.push_back('_');
        }
        s += part;
      }
      idents.push_back(s);
      w[i] = 1.0 / std::pow((double)(i + 1), 1.1);
    }
    ident_alias.build(w);
    std::vector<double> cw(6000);
    for (int i = 0; i < 6000; i++) cw[i] = 1.0 / std::pow((double)(i + 1), 1.0);
    cjk_alias.build(cw);
  }
};

void put_cjk(std::string& o, uint32_t rank) {
  // a fixed pseudo-random permutation of ranks onto U+4E00.. keeps frequent
  // ideographs scattered over the block
  uint32_t cp = 0x4E00 + (uint32_t)((rank * 2654435761u) % 20000u);
  o.push_back((char)(0xE0 | (cp >> 12)));
  o.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
  o.push_back((char)(0x80 | (cp & 0x3F)));
}

void gen_number(Rng& r, std::string& o) {
  int k = r.below(10);
  if (k < 5) o += std::to_string(r.below(10));
  else if (k < 8) o += std::to_string(r.below(1000));
  else { o += "0x"; const char* h = "0123456789abcdef"; for (int i = 0; i < 4; i++) o.push_back(h[r.below(16)]); }
}

void gen_expr(Rng& r, const Lexicon& lx, int lang, std::string& o, int depth) {
  int terms = 1 + r.below(3);
  for (int t = 0; t < terms; t++) {
    if (t) o += OPS[r.below(16)];
    int k = r.below(100);
    if (k < 55) {
      o += lx.idents[lx.ident_alias.draw(r)];
      if (r.below(100) < 25) { o.push_back('.'); o += lx.idents[lx.ident_alias.draw(r)]; }
      if (r.below(100) < 30 && depth < 2) {
        o.push_back('(');
        int a = r.below(3);
        for (int i = 0; i < a; i++) { if (i) o += ", "; gen_expr(r, lx, lang, o, depth + 1); }
        o.push_back(')');
      } else if (r.below(100) < 8) {
        o.push_back('['); gen_number(r, o); o.push_back(']');
      }
    } else if (k < 70) {
      gen_number(r, o);
    } else if (k < 82) {
      char q = r.below(2) ? '"' : '\'';
      o.push_back(q);
      int w = 1 + r.below(4);
      for (int i = 0; i < w; i++) { if (i) o.push_back(' '); o += lx.idents[lx.ident_alias.draw(r)]; }
      o.push_back(q);
    } else {
      o += KW[lang][r.below(24)];
    }
  }
}

void gen_code_line(Rng& r, const Lexicon& lx, int lang, int indent, bool tabs, std::string& o) {
  if (tabs) o.append((size_t)indent, '\t'); else o.append((size_t)indent * 4, ' ');
  int k = r.below(100);
  if (k < 8) {  // comment
    o += (lang == 0 ? "# " : "// ");
    int w = 2 + r.below(8);
    for (int i = 0; i < w; i++) { if (i) o.push_back(' '); o += lx.idents[lx.ident_alias.draw(r)]; }
  } else if (k < 30) {
    o += KW[lang][r.below(24)]; o.push_back(' ');
    gen_expr(r, lx, lang, o, 0);
    o += (lang == 0 ? ":" : " {");
  } else if (k < 80) {
    o += lx.idents[lx.ident_alias.draw(r)];
    o += OPS[r.below(2) ? 0 : 13 + r.below(2)];
    gen_expr(r, lx, lang, o, 0);
    if (lang != 0 && lang != 5) o.push_back(';');
  } else if (k < 90) {
    o += KW[lang][r.below(24)]; o.push_back(' ');
    o += lx.idents[lx.ident_alias.draw(r)]; o.push_back('(');
    int a = r.below(4);
    for (int i = 0; i < a; i++) { if (i) o += ", "; o += lx.idents[lx.ident_alias.draw(r)]; }
    o += (lang == 0 ? "):" : ") {");
  } else if (k < 95 && lang != 0) {
    o.push_back('}');
  }
}

void gen_prose_line(Rng& r, const Lexicon& lx, std::string& o) {
  int k = r.below(100);
  if (k < 10) o += "# "; else if (k < 20) o += "- ";
  int segs = 1 + r.below(6);
  for (int s = 0; s < segs; s++) {
    if (r.below(100) < 75) {
      int run = 1 + r.below(16);
      for (int i = 0; i < run; i++) put_cjk(o, lx.cjk_alias.draw(r));
      int p = r.below(100);
      if (p < 30) o += "\xEF\xBC\x8C";       // ，
      else if (p < 45) o += "\xE3\x80\x82";  // 。
    } else {
      o.push_back(' '); o += lx.idents[lx.ident_alias.draw(r)]; o.push_back(' ');
      if (r.below(100) < 30) { o.push_back('`'); o += lx.idents[lx.ident_alias.draw(r)]; o += "()` "; }
    }
  }
}

// kind: 0 code (one language), 1 multi-language code, 2 code + Chinese prose mix
void gen_sample(int kind, uint64_t seed, uint64_t index, const Lexicon& lx, uint8_t* out, size_t n) {
  Rng r(seed, index + 1);
  int lang = kind == 0 ? 0 : (int)r.below(8);
  bool prose = kind == 2 && r.below(100) < 30;
  bool crlf = r.below(100) < 5;
  bool tabs = r.below(100) < 20;
  std::string buf;
  buf.reserve(n + 256);
  int indent = 0;
  while (buf.size() < n) {
    if (prose) {
      gen_prose_line(r, lx, buf);
    } else {
      int d = (int)r.below(5);
      if (d == 0 && indent < 6) indent++; else if (d == 1 && indent > 0) indent--;
      gen_code_line(r, lx, lang, indent, tabs, buf);
    }
    if (crlf) buf.push_back('\r');
    buf.push_back('\n');
    if (r.below(100) < 8) { if (crlf) buf.push_back('\r'); buf.push_back('\n'); }
  }
  // cut at a UTF-8 scalar boundary, pad the tail so that the sample has exactly n bytes
  size_t cut = n;
  while (cut > 0 && ((uint8_t)buf[cut] & 0xC0) == 0x80) cut--;
  std::memcpy(out, buf.data(), cut);
  for (size_t i = cut; i < n; i++) out[i] = ' ';
}

template <class F>
void parallel_for(size_t n, int threads, F&& f) {
  std::atomic<size_t> next{0};
  const size_t grain = 64;
  auto w = [&]() {
    for (;;) {
      size_t lo = next.fetch_add(grain);
      if (lo >= n) break;
      size_t hi = std::min(n, lo + grain);
      for (size_t i = lo; i < hi; i++) f(i);
    }
  };
  if (threads <= 1) { w(); return; }
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; t++) ts.emplace_back(w);
  for (auto& t : ts) t.join();
}

// ---- the "exact" allow-rule (/root/reference/data/exact.regex) -----------------
inline bool is_cjk(uint32_t cp) { return (cp >= 0x3400 && cp <= 0x4DBF) || (cp >= 0x4E00 && cp <= 0x9FFF); }
inline bool is_punct(uint8_t c) { return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126); }
const char* const ALLOW_OPS[] = {"+", "-", "*", "/", "%", "&", "|", "^", "!", "~", "&&", "||", "==", "!=", "!==", "<", ">", "<=", ">=", "<<", ">>", ">>>", "++", "--", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "=>", "->", ".", "...", "?", "=", ":=", "[]", "()"};

bool allow_exact(const uint8_t* s, size_t n) {
  if (n == 0) return false;
  // ^(?:.)$ : one scalar that is not '\n'
  {
    size_t l = s[0] < 0x80 ? 1 : (s[0] >> 5) == 6 ? 2 : (s[0] >> 4) == 14 ? 3 : 4;
    if (l == n && s[0] != '\n') return true;
  }
  bool lower = true, upper = true, spaces = true, tabsr = true;
  for (size_t i = 0; i < n; i++) {
    lower &= (s[i] >= 'a' && s[i] <= 'z');
    upper &= (s[i] >= 'A' && s[i] <= 'Z');
    spaces &= s[i] == ' ';
    tabsr &= s[i] == '\t';
  }
  if (lower || upper || spaces || tabsr) return true;
  if (n >= 2 && s[0] >= 'A' && s[0] <= 'Z') {  // [A-Z][a-z]+
    bool ok = true;
    for (size_t i = 1; i < n; i++) ok &= (s[i] >= 'a' && s[i] <= 'z');
    if (ok) return true;
  }
  if (s[0] == '\'') {
    static const char* const C[] = {"re", "ve", "s", "d", "ll", "t", "m"};
    for (auto c : C) if (n - 1 == std::strlen(c) && !std::memcmp(s + 1, c, n - 1)) return true;
  }
  if (n % 3 == 0) {  // CJK run
    bool ok = true;
    for (size_t i = 0; i < n && ok; i += 3) {
      if ((s[i] >> 4) != 14) { ok = false; break; }
      uint32_t cp = ((s[i] & 0x0F) << 12) | ((s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F);
      ok = is_cjk(cp);
    }
    if (ok) return true;
  }
  // " ?(op) ?"  and  " ?[[:punct:]] ?"
  size_t a = 0, b = n;
  if (s[a] == ' ') a++;
  if (b > a && s[b - 1] == ' ') b--;
  for (int pass = 0; pass < 4; pass++) {  // optional spaces are optional: try all 4 trims
    size_t aa = (pass & 1) ? a : 0, bb = (pass & 2) ? b : n;
    if (bb <= aa) continue;
    size_t m = bb - aa;
    if (m == 1 && is_punct(s[aa])) return true;
    for (auto op : ALLOW_OPS) if (m == std::strlen(op) && !std::memcmp(s + aa, op, m)) return true;
  }
  return false;
}

}  // namespace

extern "C" {

// Sample lengths: lognormal(ln 3000, 1.2) clipped to [16, 262144], until the sum
// reaches total_bytes (last one trimmed, min 16).  Returns S (<= cap).
uint64_t tgx_synth_sample_lengths(uint64_t seed, uint64_t total_bytes, uint64_t* lens, uint64_t cap) {
  Rng r(seed, 0x5EED);
  uint64_t sum = 0, S = 0;
  while (sum < total_bytes && S < cap) {
    double l = std::exp(std::log(3000.0) + 1.2 * r.normal());
    uint64_t n = (uint64_t)std::llround(std::min(262144.0, std::max(16.0, l)));
    if (sum + n > total_bytes) n = std::max<uint64_t>(16, total_bytes - sum);
    lens[S++] = n;
    sum += n;
  }
  return S;
}

void tgx_synth_corpus(int kind, uint64_t seed, const uint64_t* offsets, uint64_t S, uint8_t* blob, int threads) {
  Lexicon lx(seed);
  parallel_for(S, threads, [&](size_t i) {
    gen_sample(kind, seed, i, lx, blob + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
  });
}

// Samples [first, first + count) of the same corpus, written from blob[0] on (offsets = the offsets of the WHOLE
// corpus): every rank of a sharded run generates exactly its own shard of ONE corpus.
void tgx_synth_corpus_range(int kind, uint64_t seed, const uint64_t* offsets, uint64_t first, uint64_t count,
                            uint8_t* blob, int threads) {
  Lexicon lx(seed);
  const uint64_t b0 = offsets[first];
  parallel_for(count, threads, [&](size_t i) {
    gen_sample(kind, seed, first + i, lx, blob + (offsets[first + i] - b0),
               (size_t)(offsets[first + i + 1] - offsets[first + i]));
  });
}

int tgx_synth_allow_exact(const uint8_t* s, uint64_t n) { return allow_exact(s, n) ? 1 : 0; }

// VocabularyGenerator::feed + generate (/root/reference/src/generate.rs:54-243).
// Returns the vocabulary size actually produced (<= vocab_size), or -1 if tok_cap is short.
int64_t tgx_synth_vocab(const uint8_t* blob, const uint64_t* offsets, uint64_t S, uint64_t seed,
                        uint64_t vocab_size, uint32_t max_token_length, double insert_probability,
                        int threads, uint8_t* tok_blob, uint64_t tok_cap, uint64_t* tok_off,
                        double* scores, uint8_t* keep) {
  int T = std::max(1, threads);
  std::vector<std::unordered_map<std::string, uint64_t>> maps((size_t)T);
  std::atomic<size_t> next{0};
  auto worker = [&](int tid) {
    auto& fm = maps[(size_t)tid];
    std::vector<std::string> sample_tokens;
    for (;;) {
      size_t si = next.fetch_add(1);
      if (si >= S) break;
      Rng r(seed ^ 0xF00DULL, si + 1);
      const uint8_t* p = blob + offsets[si];
      size_t n = (size_t)(offsets[si + 1] - offsets[si]);
      sample_tokens.clear();
      for (size_t i = 0; i < n; i++) {  // generate.rs:91-112
        if ((p[i] & 0xC0) == 0x80) continue;  // char_indices
        size_t len = 0, j = i;
        while (j < n) {
          size_t cl = p[j] < 0x80 ? 1 : (p[j] >> 5) == 6 ? 2 : (p[j] >> 4) == 14 ? 3 : 4;
          len += cl;
          if (len > max_token_length || j + cl > n) break;
          j += cl;
          if (r.uniform() < insert_probability && allow_exact(p + i, len))
            sample_tokens.emplace_back((const char*)p + i, len);
        }
      }
      std::sort(sample_tokens.begin(), sample_tokens.end());  // per-sample dedup (HashSet) :118-133
      sample_tokens.erase(std::unique(sample_tokens.begin(), sample_tokens.end()), sample_tokens.end());
      for (auto& t : sample_tokens) fm[t] += 1;
    }
  };
  {
    std::vector<std::thread> ts;
    for (int t = 0; t < T; t++) ts.emplace_back(worker, t);
    for (auto& t : ts) t.join();
  }
  auto& all = maps[0];
  for (int t = 1; t < T; t++) {
    for (auto& kv : maps[(size_t)t]) all[kv.first] += kv.second;
    maps[(size_t)t].clear();
  }
  // generate.rs:150-151  sort by frequency desc (ties: bytes asc, for determinism)
  std::vector<std::pair<const std::string*, uint64_t>> freq;
  freq.reserve(all.size());
  for (auto& kv : all) freq.emplace_back(&kv.first, kv.second);
  std::sort(freq.begin(), freq.end(), [](const auto& a, const auto& b) {
    return a.second != b.second ? a.second > b.second : *a.first < *b.first;
  });
  uint64_t highest = freq.empty() ? 1 : freq[0].second;  // :159-163
  struct Tok { std::string v; double score; bool keep; };
  std::vector<Tok> vocab;
  for (int b = 0; b < 255; b++) vocab.push_back(Tok{std::string(1, (char)b), (double)highest, true});  // :164-169
  for (auto& f : freq) {  // :194-208
    if (vocab.size() >= vocab_size) break;
    if (f.first->size() > 1) vocab.push_back(Tok{*f.first, (double)(f.second * f.first->size()), false});
  }
  std::stable_sort(vocab.begin(), vocab.end(), [](const Tok& a, const Tok& b) { return a.score > b.score; });  // :211-216
  double sum = 0.0;  // logprobs :237-243
  for (auto& t : vocab) sum += t.score;
  double logsum = std::log(sum);
  uint64_t o = 0;
  for (size_t i = 0; i < vocab.size(); i++) {
    if (o + vocab[i].v.size() > tok_cap) return -1;
    tok_off[i] = o;
    std::memcpy(tok_blob + o, vocab[i].v.data(), vocab[i].v.size());
    o += vocab[i].v.size();
    scores[i] = std::log(vocab[i].score) - logsum;
    keep[i] = vocab[i].keep;
  }
  tok_off[vocab.size()] = o;
  return (int64_t)vocab.size();
}

}  // extern "C"
