// TEST INFRASTRUCTURE ONLY -- compiled CPU oracle for the BPE hot path.
//
// Int-level literal restatement of /root/reference/core.ts: tokens are their
// `index` (the reference stores `code = fromCodePoint(index+1)`, core.ts:149,189),
// a document is an int32 array (the reference: one code string, core.ts:106,206).
// The ALGORITHM and its data-structure shapes are kept on purpose, because this
// file is also the timed CPU baseline (bench.py cpu_baseline / --impl reference):
//   * find_next_merge: fresh two-level hash map per call + the RUNNING arg-max
//     inside the counting loop (core.ts:259-310) -- not a post-hoc max;
//   * apply_merge: full left-to-right non-overlapping rewrite of every document
//     (String.replaceAll semantics, core.ts:356-359);
//   * encode: one full replace pass per merge, in merge order (core.ts:404-406).
// Nothing here is linked into, or called by, the product library.
//
// Build: see oracle/Makefile  (g++ -O2 -shared -fPIC).
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

struct Oracle {
  std::vector<std::vector<int32_t>> corpus;  // core.ts:106 corpus_in_code
  std::vector<int32_t> len16;                // UTF-16 length of token.chars, per index
  std::vector<int32_t> merges;               // a,b,c triples in training order (core.ts:88-91)
};

// String.prototype.replaceAll(a.code+b.code, c.code) on one document:
// scan left to right, on a match emit c and resume AFTER the match.
static void replace_all(std::vector<int32_t>& doc, int32_t a, int32_t b, int32_t c) {
  size_t n = doc.size(), w = 0, i = 0;
  while (i < n) {
    if (i + 1 < n && doc[i] == a && doc[i + 1] == b) {
      doc[w++] = c;
      i += 2;
    } else {
      doc[w++] = doc[i++];
    }
  }
  doc.resize(w);
}

}  // namespace

extern "C" {

void* orc_create() { return new Oracle(); }
void orc_destroy(void* h) { delete static_cast<Oracle*>(h); }

// core.ts:206 / :215 -- push one sample (ids are token indices).
void orc_add_document(void* h, const int32_t* ids, int64_t n) {
  static_cast<Oracle*>(h)->corpus.emplace_back(ids, ids + n);
}

void orc_clear_corpus(void* h) { static_cast<Oracle*>(h)->corpus.clear(); }

void orc_set_len16(void* h, const int32_t* len16, int32_t n) {
  static_cast<Oracle*>(h)->len16.assign(len16, len16 + n);
}

int64_t orc_num_documents(void* h) { return (int64_t)static_cast<Oracle*>(h)->corpus.size(); }

int64_t orc_document_length(void* h, int64_t d) { return (int64_t)static_cast<Oracle*>(h)->corpus[d].size(); }

void orc_get_document(void* h, int64_t d, int32_t* out) {
  auto& doc = static_cast<Oracle*>(h)->corpus[d];
  if (!doc.empty()) std::memcpy(out, doc.data(), doc.size() * sizeof(int32_t));
}

int64_t orc_total_tokens(void* h) {
  int64_t s = 0;
  for (auto& d : static_cast<Oracle*>(h)->corpus) s += (int64_t)d.size();
  return s;
}

// core.ts:247-313.  Returns 1 and fills (a,b,weight) or returns 0 for `null`.
// min_weight: caller passes the already-defaulted value (options?.min_weight || 2).
// max_length: 0 = falsy = unlimited (core.ts:272).
int orc_find_next_merge(void* h, int64_t min_weight, int32_t max_length, int32_t* out_a, int32_t* out_b,
                        int64_t* out_weight) {
  Oracle* o = static_cast<Oracle*>(h);
  std::unordered_map<int32_t, std::unordered_map<int32_t, int64_t>> a_b_c_weights;  // core.ts:259
  int32_t max_a = -1, max_b = -1;
  int64_t max_c_index = 0, max_c_weight = 0;  // 0 == "null" (core.ts:297 `!max_c_weight`)
  const int32_t* len16 = o->len16.data();
  for (auto& sample : o->corpus) {  // core.ts:265
    int32_t last_a = -1, a = -1;    // -1 == null
    for (int32_t b : sample) {      // core.ts:268
      if (a >= 0 && (!max_length || len16[a] + len16[b] <= max_length)) {  // core.ts:270-273
        auto& b_c_weights = a_b_c_weights[a];                              // core.ts:274-278
        int64_t& slot = b_c_weights[b];
        int64_t c_weight = slot;
        if (!c_weight) {  // core.ts:281-283
          slot = 1;
          c_weight = 1;
        } else {
          if (a == b && last_a == a) {  // core.ts:285-290
            last_a = -1;
            a = b;
            continue;
          }
          c_weight++;
          slot = c_weight;
        }
        int64_t c_index = (int64_t)a + (int64_t)b;  // core.ts:294
        if (!max_c_weight || c_weight > max_c_weight ||
            (c_weight == max_c_weight && c_index < max_c_index)) {  // core.ts:296-305
          max_a = a;
          max_b = b;
          max_c_weight = c_weight;
          max_c_index = c_index;
        }
      }
      last_a = a;  // core.ts:307-308
      a = b;
    }
  }
  if (!max_c_weight) return 0;                             // core.ts:312
  if (min_weight && max_c_weight < min_weight) return 0;   // core.ts:313
  *out_a = max_a;
  *out_b = max_b;
  *out_weight = max_c_weight;
  return 1;
}

// core.ts:350-359 (corpus + merge list part; weight bookkeeping stays with the caller).
void orc_apply_merge(void* h, int32_t a, int32_t b, int32_t c) {
  Oracle* o = static_cast<Oracle*>(h);
  o->merges.push_back(a);
  o->merges.push_back(b);
  o->merges.push_back(c);
  if ((int32_t)o->len16.size() <= c) o->len16.resize(c + 1, 0);
  o->len16[c] = o->len16[a] + o->len16[b];  // chars = a.chars + b.chars (core.ts:318)
  for (auto& doc : o->corpus) replace_all(doc, a, b, c);
}

// core.ts:365-383.  `first_new_index` = token_table.length before the loop.
// log receives (a, b, weight) per merge.  Returns merges done.
int64_t orc_merge_until(void* h, int64_t min_weight, int32_t max_length, int64_t max_iterations,
                        int32_t first_new_index, int32_t* log_a, int32_t* log_b, int64_t* log_w, int64_t cap) {
  int64_t done = 0;
  for (int64_t iteration = 1; !max_iterations || iteration <= max_iterations; iteration++) {
    if (done >= cap) break;
    int32_t a, b;
    int64_t w;
    if (!orc_find_next_merge(h, min_weight, max_length, &a, &b, &w)) break;
    log_a[done] = a;
    log_b[done] = b;
    log_w[done] = w;
    orc_apply_merge(h, a, b, first_new_index + (int32_t)done);
    done++;
  }
  return done;
}

void orc_load_merges(void* h, const int32_t* abc, int64_t n) {
  static_cast<Oracle*>(h)->merges.assign(abc, abc + 3 * n);
}

// core.ts:404-406: one replaceAll pass per merge, in order.  `ids` already hold
// single-character token indices (core.ts:396-402 is a host dictionary lookup).
int64_t orc_encode(void* h, const int32_t* ids, int64_t n, int32_t* out) {
  Oracle* o = static_cast<Oracle*>(h);
  std::vector<int32_t> doc(ids, ids + n);
  for (size_t m = 0; m + 2 < o->merges.size(); m += 3)
    replace_all(doc, o->merges[m], o->merges[m + 1], o->merges[m + 2]);
  if (!doc.empty()) std::memcpy(out, doc.data(), doc.size() * sizeof(int32_t));
  return (int64_t)doc.size();
}

// A faster *equivalent* encoder used only to produce expected outputs for big
// inputs in tests ("repeatedly merge every occurrence of the lowest-rank pair
// present"); checked against orc_encode in the CPU tests.
int64_t orc_encode_fast(void* h, const int32_t* ids, int64_t n, int32_t* out) {
  Oracle* o = static_cast<Oracle*>(h);
  std::unordered_map<uint64_t, std::pair<int32_t, int32_t>> rank;  // (a,b) -> (rank, c), first rule wins
  for (size_t m = 0; m < o->merges.size(); m += 3) {
    uint64_t k = ((uint64_t)(uint32_t)o->merges[m] << 32) | (uint32_t)o->merges[m + 1];
    rank.emplace(k, std::make_pair((int32_t)(m / 3), o->merges[m + 2]));
  }
  std::vector<int32_t> doc(ids, ids + n);
  for (;;) {
    int32_t best = INT32_MAX, ba = 0, bb = 0, bc = 0;
    for (size_t i = 0; i + 1 < doc.size(); i++) {
      auto it = rank.find(((uint64_t)(uint32_t)doc[i] << 32) | (uint32_t)doc[i + 1]);
      if (it != rank.end() && it->second.first < best) {
        best = it->second.first;
        ba = doc[i];
        bb = doc[i + 1];
        bc = it->second.second;
      }
    }
    if (best == INT32_MAX) break;
    replace_all(doc, ba, bb, bc);
  }
  if (!doc.empty()) std::memcpy(out, doc.data(), doc.size() * sizeof(int32_t));
  return (int64_t)doc.size();
}

}  // extern "C"
