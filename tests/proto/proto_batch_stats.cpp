// PROTOTYPE / MEASUREMENT ONLY (test infrastructure, never linked into the product).
//
// How many consecutive merges of the exact mergeUntil sequence (core.ts:365-383) could one barrier round of the device
// loop commit?  The device takes the top-K pairs of ONE decision and runs their site passes side by side; merge i+1 of
// the batch is exact iff it is the true next winner after merges <= i (no born pair out-ranks it, its count did not move)
// -- this program replays the true sequence with the incremental CPU oracle and, for several batching rules, measures the
// batch lengths a greedy batcher would reach and why its batches end.
//
//   g++ -O2 -std=c++17 -pthread -o /tmp/proto_batch_stats tests/proto/proto_batch_stats.cpp bpe_tokenizer_b200/csrc/synth.cpp
//   /tmp/proto_batch_stats <bytes> <merges> [max_length]
#include "../../oracle/fast_oracle.cpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>

extern "C" int bpe_synth_corpus(int64_t target_bytes, uint64_t seed, int32_t vocab, uint64_t word_seed, uint8_t* text, int64_t text_cap,
                                int64_t* doc_offsets, int64_t offsets_cap, int64_t* n_bytes, int64_t* n_docs);

namespace {

struct Cand {
  uint64_t key;
  int64_t count, negsum;
};

// the K best valid entries of the heap, best first (ties on (count, sum) in heap order)
std::vector<Cand> topk(Fast& F, size_t K, int32_t max_length) {
  std::vector<Fast::Entry> taken;
  std::vector<Cand> out;
  std::set<uint64_t> seen;
  while (!F.heap.empty() && out.size() < K) {
    Fast::Entry e = F.heap.top();
    F.heap.pop();
    uint64_t key = std::get<2>(e);
    int32_t a = (int32_t)(key >> 32), b = (int32_t)(key & 0xFFFFFFFFu);
    auto it = F.pairs.find(key);
    bool stale = it == F.pairs.end() || it->second.count != std::get<0>(e);
    bool too_long = max_length > 0 && F.len16[a] + F.len16[b] > max_length;
    if (stale || too_long || seen.count(key)) continue;  // (dropped for good: a stale entry is never a winner)
    seen.insert(key);
    taken.push_back(e);
    out.push_back(Cand{key, std::get<0>(e), std::get<1>(e)});
  }
  for (auto& e : taken) F.heap.push(e);
  return out;
}

// true winner (the oracle's findNextMerge part); pops it.  tie_out: more than one pair shared (count, sum)
bool find_next(Fast& F, int64_t min_weight, int32_t max_length, uint64_t* key_out, int64_t* cnt_out, bool* tie_out) {
  std::vector<Fast::Entry> tied;
  while (!F.heap.empty()) {
    Fast::Entry e = F.heap.top();
    uint64_t key = std::get<2>(e);
    int32_t a = (int32_t)(key >> 32), b = (int32_t)(key & 0xFFFFFFFFu);
    auto it = F.pairs.find(key);
    bool stale = it == F.pairs.end() || it->second.count != std::get<0>(e);
    bool too_long = max_length > 0 && F.len16[a] + F.len16[b] > max_length;
    if (!tied.empty() && !(std::get<0>(e) == std::get<0>(tied[0]) && std::get<1>(e) == std::get<1>(tied[0]))) break;
    F.heap.pop();
    if (stale || too_long) continue;
    bool dup = false;
    for (const Fast::Entry& t : tied) dup = dup || std::get<2>(t) == key;
    if (!dup) tied.push_back(e);
  }
  if (tied.empty() || std::get<0>(tied[0]) < min_weight) {
    for (const Fast::Entry& t : tied) F.heap.push(t);
    return false;
  }
  size_t win = 0;
  if (tied.size() > 1) {
    uint32_t best = NONE;
    for (size_t i = 0; i < tied.size(); i++) {
      uint64_t key = std::get<2>(tied[i]);
      uint32_t p = last_counted(F, F.pairs[key], (int32_t)(key >> 32), (int32_t)(key & 0xFFFFFFFFu));
      if (p < best) {
        best = p;
        win = i;
      }
    }
  }
  for (size_t i = 0; i < tied.size(); i++)
    if (i != win) F.heap.push(tied[i]);
  *key_out = std::get<2>(tied[win]);
  *cnt_out = std::get<0>(tied[win]);
  *tie_out = tied.size() > 1;
  return true;
}

// largest token index among the neighbours of the sites of (a, b) -- a neighbour >= c_first was created inside the batch
int32_t max_neighbour(const Fast& F, int32_t a, int32_t b) {
  auto it = F.pairs.find(key_of(a, b));
  if (it == F.pairs.end()) return -1;
  int32_t m = -1;
  for (uint32_t p : it->second.occ) {
    if (!occurs_at(F, p, a, b)) continue;
    uint32_t q = F.nxt[p];
    uint32_t x = left_of(F, p), y = right_of(F, q);
    if (x != NONE) m = std::max(m, F.tok[x]);
    if (y != NONE) m = std::max(m, F.tok[y]);
  }
  return m;
}

enum Reason { R_ORDER = 0, R_TIE, R_TOKENS, R_ADJ, R_BIG, R_CAP, R_N };
const char* reason_name[R_N] = {"not-the-next-candidate", "tie-on-primary", "shares-a-token", "adjacent-to-batch-site", "big-merge", "cap-K"};

struct Batcher {
  std::string name;
  size_t K;
  bool need_disjoint, forbid_adjacent;
  int64_t small_limit;
  // state
  std::vector<Cand> cands;
  size_t i = 0;
  int32_t c_first = 0;
  std::set<int32_t> toks;
  // stats
  int64_t merges_small = 0, rounds_small = 0, merges_big = 0;
  int64_t ends[R_N] = {0};
  std::map<size_t, int64_t> len_hist;
  void close() {
    if (i) len_hist[i]++;
  }
};

}  // namespace

int main(int argc, char** argv) {
  int64_t bytes = argc > 1 ? atoll(argv[1]) : 16000000;
  int64_t M = argc > 2 ? atoll(argv[2]) : 4000;
  int32_t max_length = argc > 3 ? atoi(argv[3]) : 0;
  int64_t nb = 0, nd = 0;
  bpe_synth_corpus(bytes, 43, 50000, 42, nullptr, 0, nullptr, 0, &nb, &nd);
  std::vector<uint8_t> text(nb);
  std::vector<int64_t> off(nd + 1);
  bpe_synth_corpus(bytes, 43, 50000, 42, text.data(), nb, off.data(), nd + 1, &nb, &nd);
  std::vector<int32_t> ids(nb);
  int32_t map[256];
  for (int i = 0; i < 256; i++) map[i] = -1;
  int32_t T = 0;
  for (int64_t i = 0; i < nb; i++) {
    if (map[text[i]] < 0) map[text[i]] = T++;
    ids[i] = map[text[i]];
  }
  Fast F;
  fast_add_documents(&F, ids.data(), off.data(), nd);
  std::vector<int32_t> l16(T, 1);
  F.len16 = l16;
  build(F);
  fprintf(stderr, "corpus %lld bytes, %lld docs, %d chars\n", (long long)nb, (long long)nd, T);

  std::vector<Batcher> B;
  for (size_t K : {2, 4, 8, 16}) {
    B.push_back(Batcher{"K" + std::to_string(K) + " identity-only", K, false, false, 16384});
    B.push_back(Batcher{"K" + std::to_string(K) + " disjoint-tokens (adjacency handled)", K, true, false, 16384});
    B.push_back(Batcher{"K" + std::to_string(K) + " disjoint-tokens + no-adjacency", K, true, true, 16384});
  }
  int64_t done = 0;
  const int32_t first_new = T;
  std::vector<int64_t> phase_marks;
  while (done < M) {
    // candidates of a decision taken NOW (before the winner is popped), for the batchers that may start a batch here
    std::vector<Cand> snap = topk(F, 18, max_length);
    uint64_t key;
    int64_t cnt;
    bool tie;
    if (!find_next(F, 2, max_length, &key, &cnt, &tie)) break;
    const int32_t a = (int32_t)(key >> 32), b = (int32_t)(key & 0xFFFFFFFFu), c = first_new + (int32_t)done;
    const int32_t maxnb = max_neighbour(F, a, b);
    for (Batcher& x : B) {
      const bool big = cnt > x.small_limit;
      int reason = -1;
      if (x.i == 0) reason = -2;  // nothing open
      else if (big) reason = R_BIG;
      else if (x.i >= x.K) reason = R_CAP;
      else if (x.i >= x.cands.size() || x.cands[x.i].key != key || x.cands[x.i].count != cnt) reason = R_ORDER;
      else if (tie || (x.i + 1 < x.cands.size() && x.cands[x.i + 1].count == cnt && x.cands[x.i + 1].negsum == x.cands[x.i].negsum)) reason = R_TIE;
      else if (x.need_disjoint && (a >= x.c_first || b >= x.c_first || x.toks.count(a) || x.toks.count(b))) reason = R_TOKENS;
      else if (x.forbid_adjacent && maxnb >= x.c_first) reason = R_ADJ;
      if (reason != -1) {  // this merge opens a new round
        if (reason >= 0) x.ends[reason]++;
        x.close();
        x.cands = snap;
        x.i = 0;
        x.c_first = c;
        x.toks.clear();
        if (big) {
          x.merges_big++;
          x.i = 0;  // a big merge runs alone and is not counted as a round of the latency-bound regime
          continue;
        }
        x.rounds_small++;
      }
      x.merges_small++;
      x.toks.insert(a);
      x.toks.insert(b);
      x.i++;
    }
    if ((int32_t)F.len16.size() <= c) F.len16.resize((size_t)c + 1, 0);
    F.len16[c] = F.len16[a] + F.len16[b];
    F.changed.clear();
    apply_merge(F, a, b, c);
    for (uint64_t k : F.changed) {
      const PairInfo& pi = F.pairs[k];
      if (pi.count > 0) F.heap.emplace(pi.count, -(int64_t)((k >> 32) + (k & 0xFFFFFFFFu)), k);
    }
    F.merge_no++;
    done++;
    if ((done & (done - 1)) == 0 || done == M || done % 4000 == 0) {
      printf("---- after %lld merges (last weight %lld) ----\n", (long long)done, (long long)cnt);
      for (Batcher& x : B) {
        printf("%-48s small merges %lld in %lld rounds = %.2f per round; big %lld; ends:", x.name.c_str(), (long long)x.merges_small,
               (long long)x.rounds_small, x.rounds_small ? (double)x.merges_small / x.rounds_small : 0.0, (long long)x.merges_big);
        for (int r = 0; r < R_N; r++) printf(" %s=%lld", reason_name[r], (long long)x.ends[r]);
        printf("\n");
      }
      fflush(stdout);
    }
  }
  return 0;
}
