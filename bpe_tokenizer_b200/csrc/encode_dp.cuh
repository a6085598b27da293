// encode_dp.cuh -- K4, forward path: exact encodeToCode (core.ts:392-409) in ONE left-to-right pass per document.
//
// The reference applies every merge in training order with replaceAll over the whole string (core.ts:404-406).  For a merge
// table built by that very process (token indices grow with the rank, characters first) the result can be computed without
// replaying the merges: with last[i] = the token the encoding of the prefix text[0..i) ends with,
//
//     last[i] = the LONGEST token t ending at i such that (last[i - len(t)], t) is a pair the sequential process leaves
//               standing ("compatible"), or len(t) == i,
//
// and the encoding of the document is the chain last[n], last[n - len(last[n])], ... read backwards.  The tokens ending at
// i, longest first, come out of an Aho-Corasick automaton over the tokens' characters (one transition per character, then
// the chain of dictionary suffixes).  Compatibility of two tokens is decided on their merge trees alone: walk down the right
// spine of t1 and the left spine of t2, at each level the rule across the boundary (if there is one) must be YOUNGER than the
// rule that built the part just opened -- otherwise the sequential process would have fired it first.
// Prototype, fuzzed against the literal restatement (runs, chains, 1-letter alphabets, unseen text):
// tests/proto/proto_encode_dp.py.
//
// Mapping: one THREAD per document, a warp owns 32 consecutive documents and steps through them in lock step (a lane whose
// document is over idles).  last[] goes to a scratch area of the warp, transposed ([position][lane]: one coalesced 64-byte
// store per step); the last 64 values also sit in a thread-local ring, which is where the compatibility test reads them.
// Two walks back along the chain give the token count and then the tokens in forward order.
// Documents longer than DP_MAX_LEN, and warps that find no scratch space, leave their documents to the per-document kernel
// (out_len = EL_LONG), like the lane path does.
#pragma once
#include "encode_lanes.cuh"

namespace bpe {

constexpr uint32_t DP_MAX_LEN = 8192;   // characters of a document on this path
constexpr uint32_t DP_NONE = 0xFFFFu;
constexpr int DP_RING = 64;

constexpr uint32_t DP_CACHE = 8192;   // entries of a block's cache of compatibility verdicts (shared memory, 8 bytes each)

struct DpTables {
  const unsigned long long* dfa;  // [states][n_alpha]: next state | longest token that is a suffix of its string << 32 | that
                                  // token's length << 48 -- one load per character gives all three
  const uint16_t* tok_len;    // [tokens] characters of the token
  const uint16_t* shorter;    // [tokens] longest token that is a proper suffix of the token's string (DP_NONE: a character)
  const uint32_t* split;      // [tokens] left part | right part << 16 (a character: itself twice)
  uint32_t n_alpha;           // single-character tokens = indices 0 .. n_alpha-1 (the automaton's alphabet)
};

// Would the sequential process, run over chars(t1) + chars(t2), end with exactly [t1, t2]?
__device__ __forceinline__ bool dp_compatible(const LaneTables& T, const uint2* s_dense, const DpTables& D, uint32_t t1, uint32_t t2) {
  uint32_t limit = 0x7FFFFFFFu;
  for (;;) {
    const uint32_t x = el_lookup(T, s_dense, t1, t2).x;  // rk | c << 16
    if ((x & 0xFFFFu) != EL_NONE && (x >> 16) < limit) return false;
    if (t1 > t2) {
      limit = t1;
      if (t1 < D.n_alpha) return true;
      t1 = __ldg(D.split + t1) >> 16;
    } else {
      limit = t2 + 1u;
      if (t2 < D.n_alpha) return true;
      t2 = __ldg(D.split + t2) & 0xFFFFu;
    }
  }
}

// ... with the verdict of a pair of tokens remembered per block: adjacent token pairs of natural text repeat (Zipf), and a
// verdict read from shared memory replaces a walk of ~10 dependent table loads.  One 8-byte entry per slot (pair key << 1 |
// verdict, written whole), direct mapped: a lost race or a collision only costs the walk.
__device__ __forceinline__ bool dp_compatible_cached(const LaneTables& T, const uint2* s_dense, const DpTables& D, unsigned long long* s_cache, uint32_t t1,
                                                     uint32_t t2) {
  const uint32_t key = (t1 << 16) | t2;
  unsigned long long* slot = s_cache + ((key * 0x9E3779B1u) >> 19);  // 13 bits
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(slot);
  if ((uint32_t)(e >> 1) == key && e != ~0ull) return (e & 1ull) != 0;
  const bool ok = dp_compatible(T, s_dense, D, t1, t2);
  *reinterpret_cast<volatile unsigned long long*>(slot) = ((unsigned long long)key << 1) | (ok ? 1ull : 0ull);
  return ok;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_encode_dp(const int32_t* __restrict__ ids, const int64_t* __restrict__ doc_off, int64_t n_docs,
                                                           LaneTables T, DpTables D, uint16_t* __restrict__ scratch, unsigned long long scratch_cap,
                                                           unsigned long long* __restrict__ scratch_cursor, int32_t* __restrict__ out_tmp,
                                                           uint32_t* __restrict__ out_len, uint32_t* __restrict__ n_long, uint32_t* __restrict__ err) {
  extern __shared__ unsigned long long dp_smem[];
  unsigned long long* s_cache = dp_smem;                                   // [DP_CACHE]
  uint2* s_dense = reinterpret_cast<uint2*>(dp_smem + DP_CACHE);           // [EL_DENSE * EL_DENSE]
  for (int i = threadIdx.x; i < EL_DENSE * EL_DENSE; i += WARPS * 32) s_dense[i] = T.dense[i];
  for (int i = threadIdx.x; i < (int)DP_CACHE; i += WARPS * 32) s_cache[i] = ~0ull;
  __syncthreads();
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t gwarp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * WARPS;
  uint16_t ring[DP_RING];
  for (int64_t d0 = gwarp * 32; d0 < n_docs; d0 += nwarps * 32) {
    const int64_t d = d0 + lane;
    int64_t base = 0;
    uint32_t len = 0;
    bool is_long = false;
    if (d < n_docs) {
      base = doc_off[d];
      const int64_t l64 = doc_off[d + 1] - base;
      is_long = l64 > (int64_t)DP_MAX_LEN;
      len = is_long ? 0u : (uint32_t)l64;
    }
    uint32_t maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(FULL, maxlen, o));
    // scratch rows 1 .. maxlen of this warp
    unsigned long long sb = 0;
    if (lane == 0 && maxlen) sb = atomicAdd(scratch_cursor, (unsigned long long)(maxlen + 1u) * 32ull);
    sb = __shfl_sync(FULL, sb, 0);
    const bool no_space = maxlen && sb + (unsigned long long)(maxlen + 1u) * 32ull > scratch_cap;
    if (d < n_docs) {
      if (is_long || (no_space && len)) {
        out_len[d] = EL_LONG;
        atomicAdd(n_long, 1u);
      } else if (len == 0) {
        out_len[d] = 0;
      }
    }
    if (!maxlen || no_space) continue;
    uint16_t* const col = scratch + sb + lane;  // last[i] of this lane's document at col[i * 32]
    uint32_t state = 0;
    for (uint32_t i = 1; i <= maxlen; i++) {
      if (i <= len) {
        const uint32_t ch = (uint32_t)__ldg(ids + base + (i - 1));
        const unsigned long long tr = __ldg(D.dfa + (size_t)state * D.n_alpha + ch);
        state = (uint32_t)tr;
        uint32_t t = (uint32_t)(tr >> 32) & 0xFFFFu, L = (uint32_t)(tr >> 48);
        for (;;) {
          if (t == DP_NONE) {  // (cannot happen: a character is a token and is compatible with whatever precedes it)
            atomicOr(err, 1u);
            t = ch;
            break;
          }
          if (L >= i) break;  // the token starts the document (L == i)
          const uint32_t prev = (L < (uint32_t)DP_RING) ? (uint32_t)ring[(i - L) & (DP_RING - 1)] : (uint32_t)col[(size_t)(i - L) * 32];
          if (dp_compatible_cached(T, s_dense, D, s_cache, prev, t)) break;
          t = __ldg(D.shorter + t);
          L = (t == DP_NONE) ? 0u : (uint32_t)__ldg(D.tok_len + t);
        }
        ring[i & (DP_RING - 1)] = (uint16_t)t;
        col[(size_t)i * 32] = (uint16_t)t;
      }
    }
    if (len) {
      // the chain from the end: count, then the tokens in forward order
      uint32_t K = 0;
      for (uint32_t i = len; i > 0;) {
        const uint32_t t = col[(size_t)i * 32];
        K++;
        i -= __ldg(D.tok_len + t);
      }
      out_len[d] = K;
      uint32_t k = K;
      for (uint32_t i = len; i > 0;) {
        const uint32_t t = col[(size_t)i * 32];
        out_tmp[base + --k] = (int32_t)t;
        i -= __ldg(D.tok_len + t);
      }
    }
  }
}

}  // namespace bpe
