// encode_kernels.cuh -- K4 (apply the learned merge list to new text) and K5 (vector-index map).
//
// Reference: encodeToCode applies every merge in training order with replaceAll over the whole
// string (core.ts:404-406).  Because a merge of rank r can only create pairs of rank > r, that is
// equivalent to "repeat: take the lowest-rank pair present, replace all its occurrences left to right
// without overlap" (cross-checked against the literal form in tests).  One warp owns
// one document; tokens and cached pair ranks live in shared memory (global scratch for documents
// longer than ENC_WARP_MAX); the merge list is an open-addressing hash that stays L1/L2 resident.
#pragma once
#include "common.cuh"

namespace bpe {

constexpr uint32_t RK_NONE = 0xFFFFFFFFu;
constexpr uint32_t RK_DIRTY = 0xFFFFFFFEu;
constexpr unsigned long long MT_EMPTY = ~0ull;
constexpr int ENC_WARPS = 4;
constexpr int ENC_THREADS = ENC_WARPS * 32;
constexpr int ENC_WARP_MAX = 1024;  // tokens per document handled in shared memory

struct MergeTable {
  const unsigned long long* ent;  // (pair_key << 32) | (rank << 16) | c
  uint32_t mask;
  uint32_t shift;
};

__device__ __forceinline__ uint32_t mt_lookup(const MergeTable& mt, uint32_t a, uint32_t b) {
  uint32_t key = pair_key(a, b);
  uint32_t h = (key * 0x9E3779B1u) >> mt.shift;
  for (;;) {
    unsigned long long e = __ldg(mt.ent + h);
    if ((uint32_t)(e >> 32) == key) return (uint32_t)e;
    if (e == MT_EMPTY) return RK_NONE;
    h = (h + 1) & mt.mask;
  }
}

// Encode one document held in tok[0..n) (token indices) with rank cache rk[0..n).  Returns new length.
template <typename TokT>
__device__ __forceinline__ uint32_t encode_doc(TokT* tok, uint32_t* rk, uint32_t n, const MergeTable& mt,
                                               uint32_t lane) {
  for (uint32_t i = lane; i < n; i += 32) rk[i] = (i + 1 < n) ? mt_lookup(mt, tok[i], tok[i + 1]) : RK_NONE;
  __syncwarp();
  for (;;) {
    uint32_t m = RK_NONE;
    for (uint32_t i = lane; i + 1 < n; i += 32) m = min(m, rk[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if (m == RK_NONE) break;
    uint32_t c = m & 0xFFFFu;
    // pass 1: select non-overlapping occurrences left to right, compact in place
    uint32_t wr = 0, carry_run = 0, carry_sel = 0;
    for (uint32_t base = 0; base < n; base += 32) {
      uint32_t i = base + lane;
      uint32_t t = (i < n) ? (uint32_t)tok[i] : 0u;
      uint32_t r = (i < n) ? rk[i] : RK_NONE;
      bool match = (r == m);
      uint32_t M = __ballot_sync(0xFFFFFFFFu, match);
      uint32_t zeros_below = ~M & ((1u << lane) - 1u);
      uint32_t off = zeros_below ? (lane - (32u - __clz(zeros_below))) : (lane + carry_run);
      bool sel = match && ((off & 1u) == 0);
      uint32_t S = __ballot_sync(0xFFFFFFFFu, sel);
      uint32_t removed_mask = (S << 1) | carry_sel;
      bool keep = (i < n) && !((removed_mask >> lane) & 1u);
      uint32_t K = __ballot_sync(0xFFFFFFFFu, keep);
      if (keep) {
        uint32_t j = wr + __popc(K & ((1u << lane) - 1u));
        tok[j] = (TokT)(sel ? c : t);
        rk[j] = sel ? RK_DIRTY : r;
      }
      wr += __popc(K);
      carry_sel = S >> 31;
      carry_run = (M == 0xFFFFFFFFu) ? carry_run + 32u : (uint32_t)__clz(~M);
      __syncwarp();
    }
    n = wr;
    // pass 2: refresh the ranks of pairs that touch a new token
    for (uint32_t base = 0; base < n; base += 32) {
      uint32_t i = base + lane;
      uint32_t r = (i < n) ? rk[i] : RK_NONE;
      uint32_t rn = (i + 1 < n) ? rk[i + 1] : RK_NONE;
      __syncwarp();
      if (i < n && (r == RK_DIRTY || rn == RK_DIRTY)) rk[i] = (i + 1 < n) ? mt_lookup(mt, tok[i], tok[i + 1]) : RK_NONE;
      __syncwarp();
    }
  }
  return n;
}

// One warp per document.  out_tmp holds each document's tokens at the document's INPUT offset.
__global__ void __launch_bounds__(ENC_THREADS) k_encode(const int32_t* __restrict__ ids,
                                                         const int64_t* __restrict__ doc_off, int64_t n_docs,
                                                         MergeTable mt, int32_t* __restrict__ out_tmp,
                                                         uint32_t* __restrict__ out_len, uint32_t* __restrict__ g_tok,
                                                         uint32_t* __restrict__ g_rk) {
  __shared__ uint16_t s_tok[ENC_WARPS][ENC_WARP_MAX];
  __shared__ uint32_t s_rk[ENC_WARPS][ENC_WARP_MAX];
  uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t wid = (int64_t)blockIdx.x * ENC_WARPS + warp;
  int64_t nw = (int64_t)gridDim.x * ENC_WARPS;
  int64_t base0 = doc_off[0];
  for (int64_t d = wid; d < n_docs; d += nw) {
    int64_t s = doc_off[d], e = doc_off[d + 1];
    uint32_t n = (uint32_t)(e - s);
    const int32_t* src = ids + s;
    int32_t* dst = out_tmp + s;
    if (n <= ENC_WARP_MAX) {
      for (uint32_t i = lane; i < n; i += 32) s_tok[warp][i] = (uint16_t)__ldg(src + i);
      __syncwarp();
      n = encode_doc<uint16_t>(s_tok[warp], s_rk[warp], n, mt, lane);
      for (uint32_t i = lane; i < n; i += 32) dst[i] = (int32_t)s_tok[warp][i];
    } else {
      uint32_t* tok = g_tok + (s - base0);
      uint32_t* rk = g_rk + (s - base0);
      for (uint32_t i = lane; i < n; i += 32) tok[i] = (uint32_t)__ldg(src + i);
      __syncwarp();
      n = encode_doc<uint32_t>(tok, rk, n, mt, lane);
      for (uint32_t i = lane; i < n; i += 32) dst[i] = (int32_t)tok[i];
    }
    if (lane == 0) out_len[d] = n;
    __syncwarp();
  }
}

// K5: gather each document's tokens to its final offset, mapping through to_vector_index
// (core.ts:434-442); a hole (-1) marks the document's first_bad and is emitted as -(index+1).
__global__ void __launch_bounds__(ENC_THREADS) k_gather_map(const int32_t* __restrict__ out_tmp,
                                                             const int64_t* __restrict__ doc_off,
                                                             const uint64_t* __restrict__ out_off, int64_t n_docs,
                                                             const int32_t* __restrict__ tvi, int32_t n_tvi,
                                                             int32_t* __restrict__ out, int64_t* __restrict__ out_offsets,
                                                             int64_t* __restrict__ first_bad) {
  uint32_t lane = threadIdx.x & 31;
  int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t d = wid; d <= n_docs; d += nw) {
    if (d == n_docs) {
      if (lane == 0) out_offsets[d] = (int64_t)out_off[d];
      continue;
    }
    const int32_t* src = out_tmp + doc_off[d];
    uint64_t o = out_off[d];
    uint32_t n = (uint32_t)(out_off[d + 1] - o);
    uint32_t bad = 0xFFFFFFFFu;
    for (uint32_t i = lane; i < n; i += 32) {
      int32_t t = __ldg(src + i);
      int32_t v = t;
      if (tvi) {
        v = (t < n_tvi) ? __ldg(tvi + t) : -1;
        if (v < 0) {
          bad = min(bad, i);
          v = -(t + 1);
        }
      }
      out[o + i] = v;
    }
    if (first_bad) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) bad = min(bad, __shfl_xor_sync(0xFFFFFFFFu, bad, s));
      if (lane == 0) first_bad[d] = (bad == 0xFFFFFFFFu) ? -1 : (int64_t)bad;
    }
    if (lane == 0) out_offsets[d] = (int64_t)o;
  }
}

}  // namespace bpe
