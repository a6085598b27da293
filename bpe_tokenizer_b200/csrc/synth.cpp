// synth.cpp -- compiled twin of bpe_tokenizer_b200/synth.py (the specification): seeded Zipf-word
// documents wrapped '\r' ... '\n' (reference core.ts:61-64 linesToCorpus convention).  Host only.
#include "../../include/bpe_b200.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr uint64_t GAMMA = 0x9E3779B97F4A7C15ull;
inline uint64_t mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t draw(uint64_t seed, uint64_t i) { return mix(seed + (i + 1) * GAMMA); }

struct Words {
  std::vector<std::string> w;
  std::vector<uint64_t> cum;
  uint64_t total = 0;
};

Words make_words(int32_t vocab, uint64_t seed) {
  Words W;
  W.w.reserve(vocab);
  W.cum.reserve(vocab);
  uint64_t p = 0, acc = 0;
  for (int32_t k = 0; k < vocab; k++) {
    int len = 2 + (int)(draw(seed, p++) % 9);
    std::string s(len, 'a');
    for (int i = 0; i < len; i++) s[i] = (char)('a' + draw(seed, p++) % 26);
    W.w.push_back(std::move(s));
    acc += (1ull << 40) / (uint64_t)(k + 1);
    W.cum.push_back(acc);
  }
  W.total = acc;
  return W;
}

inline size_t pick(const Words& W, uint64_t r) {
  return (size_t)(std::upper_bound(W.cum.begin(), W.cum.end(), r % W.total) - W.cum.begin());
}

int64_t doc_len(const Words& W, uint64_t seed, uint64_t d) {
  uint64_t ds = draw(seed, d);
  int nw = 8 + (int)(draw(ds, 0) % 57);
  int64_t len = nw + 1;  // '\r' + (nw-1) spaces + '\n'
  for (int i = 0; i < nw; i++) len += (int64_t)W.w[pick(W, draw(ds, (uint64_t)i + 1))].size();
  return len;
}

void doc_fill(const Words& W, uint64_t seed, uint64_t d, uint8_t* out) {
  uint64_t ds = draw(seed, d);
  int nw = 8 + (int)(draw(ds, 0) % 57);
  *out++ = '\r';
  for (int i = 0; i < nw; i++) {
    const std::string& s = W.w[pick(W, draw(ds, (uint64_t)i + 1))];
    if (i) *out++ = ' ';
    std::memcpy(out, s.data(), s.size());
    out += s.size();
  }
  *out = '\n';
}

template <typename F>
void parallel_for(int64_t n, F f) {
  unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (n < 4096) nt = 1;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=] {
      int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
      for (int64_t i = lo; i < hi; i++) f(i);
    });
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" int bpe_synth_corpus(int64_t target_bytes, uint64_t seed, int32_t vocab, uint64_t word_seed, uint8_t* text,
                                int64_t text_cap, int64_t* doc_offsets, int64_t offsets_cap, int64_t* n_bytes,
                                int64_t* n_docs) {
  if (target_bytes <= 0 || vocab <= 0 || !n_bytes || !n_docs) return BPE_E_INVALID;
  static thread_local Words tl_words;  // per calling thread; workers below get it through a plain reference
  static thread_local int32_t w_vocab = 0;
  static thread_local uint64_t w_seed = 0;
  if (w_vocab != vocab || w_seed != word_seed || tl_words.w.empty()) {
    tl_words = make_words(vocab, word_seed);
    w_vocab = vocab;
    w_seed = word_seed;
  }
  const Words& W = tl_words;
  std::vector<int64_t> lens;
  int64_t total = 0;
  uint64_t d0 = 0;
  const int64_t batch = 1 << 16;
  while (total < target_bytes) {
    size_t base = lens.size();
    lens.resize(base + batch);
    parallel_for(batch, [&W, &lens, base, seed, d0](int64_t i) { lens[base + i] = doc_len(W, seed, d0 + (uint64_t)i); });
    int64_t k = 0;
    for (; k < batch && total < target_bytes; k++) total += lens[base + k];
    lens.resize(base + k);
    d0 += (uint64_t)k;
  }
  int64_t nd = (int64_t)lens.size();
  *n_bytes = total;
  *n_docs = nd;
  if (!text && !doc_offsets) return BPE_OK;
  if (!text || !doc_offsets || text_cap < total || offsets_cap < nd + 1) return BPE_E_CAPACITY;
  doc_offsets[0] = 0;
  for (int64_t d = 0; d < nd; d++) doc_offsets[d + 1] = doc_offsets[d] + lens[d];
  parallel_for(nd, [&W, seed, text, doc_offsets](int64_t d) { doc_fill(W, seed, (uint64_t)d, text + doc_offsets[d]); });
  return BPE_OK;
}
