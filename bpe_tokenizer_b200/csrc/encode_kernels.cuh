// encode_kernels.cuh -- K4 (apply the learned merge list to new text) and K5 (vector-index map).
//
// Reference: encodeToCode applies every merge in training order with replaceAll over the whole
// string (core.ts:404-406).  Because a merge of rank r can only create pairs of rank > r, that is
// equivalent to "repeat: take the lowest-rank pair present, replace all its occurrences left to right
// without overlap" (cross-checked against the literal form in tests).
//
// One warp owns one document.  Tokens never move: the document is a doubly linked list over its
// original positions (tok/next/prev in shared memory) with the rank of every adjacent pair cached
// (rk).  A round = warp-min over the cached ranks, then only the lanes that own an occurrence of that
// pair do work: relink, write the new token, refresh the two ranks next to it (two L2-resident hash
// probes).  Pairs with a == b (runs, core.ts:285-290 semantics of replaceAll) take a short serial path
// on lane 0.  Documents longer than ENC_WARP_MAX use the same code over global scratch.
#pragma once
#include "common.cuh"

namespace bpe {

constexpr uint32_t RK_NONE = 0xFFFFFFFFu;
constexpr unsigned long long MT_EMPTY = ~0ull;
constexpr int ENC_WARPS = 8;
constexpr int ENC_THREADS = ENC_WARPS * 32;
constexpr int ENC_WARP_MAX = 512;  // tokens per document handled in shared memory (10 B per token)

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

// IdxT: position type (uint16_t in shared memory, uint32_t in global scratch).  END = "no neighbour".
template <typename TokT, typename IdxT>
__device__ __forceinline__ void encode_doc(TokT* tok, IdxT* nxt, IdxT* prv, uint32_t* rk, uint32_t n,
                                           const MergeTable& mt, uint32_t lane) {
  const IdxT END = (IdxT)~(IdxT)0;
  for (uint32_t i = lane; i < n; i += 32) {
    nxt[i] = (i + 1 < n) ? (IdxT)(i + 1) : END;
    prv[i] = (i > 0) ? (IdxT)(i - 1) : END;
    rk[i] = (i + 1 < n) ? mt_lookup(mt, tok[i], tok[i + 1]) : RK_NONE;
  }
  __syncwarp();
  for (;;) {
    uint32_t m = RK_NONE;
    for (uint32_t i = lane; i < n; i += 32) m = min(m, rk[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if (m == RK_NONE) break;
    const uint32_t c = m & 0xFFFFu;
    // is this a run pair (a == b)?  every lane learns it from the first occurrence it can see
    uint32_t first = 0xFFFFFFFFu;
    for (uint32_t i = lane; i < n; i += 32)
      if (rk[i] == m) {
        first = i;
        break;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, o));
    const bool run_pair = tok[first] == tok[nxt[first]];
    constexpr bool kSmem = sizeof(IdxT) == 2;  // shared-memory path: n <= ENC_WARP_MAX, <= 16 positions per lane
    if (!run_pair || kSmem) {
      // Select the occurrences to replace (replaceAll: left to right, non-overlapping, core.ts:405).
      // a != b: occurrences cannot overlap, all are taken.  a == b: inside a run of matches only the even ones.
      uint32_t sel = 0;  // bit k: position lane + 32*k (shared-memory path only)
      if (run_pair) {
        for (uint32_t i = lane, k = 0; i < n; i += 32, k++) {
          if (rk[i] == m) {
            uint32_t preds = 0;
            IdxT p = prv[i];
            while (p != END && rk[p] == m) {
              preds++;
              p = prv[p];
            }
            if ((preds & 1u) == 0) sel |= 1u << k;
          }
        }
        __syncwarp();
      }
      for (uint32_t i = lane, k = 0; i < n; i += 32, k++) {
        if (run_pair ? ((sel >> k) & 1u) != 0 : rk[i] == m) {
          uint32_t j = nxt[i];
          IdxT jn = nxt[j];
          tok[i] = (TokT)c;
          nxt[i] = jn;
          if (jn != END) prv[jn] = (IdxT)i;
          rk[j] = RK_NONE;
          rk[i] = RK_NONE - 1;  // marks "new token here" until the refresh below
        }
      }
      __syncwarp();
      for (uint32_t i = lane; i < n; i += 32) {
        if (rk[i] == RK_NONE - 1) {
          IdxT jn = nxt[i], ip = prv[i];
          // a neighbour that is a new token too already holds its final value (after the __syncwarp above)
          if (ip != END && rk[ip] != RK_NONE - 1) rk[ip] = mt_lookup(mt, tok[ip], c);
          rk[i] = (jn != END) ? mt_lookup(mt, c, tok[jn]) : RK_NONE;
        }
      }
      __syncwarp();
    } else {
      // a == b in a document too long for shared memory: walk the list once on lane 0
      if (lane == 0) {
        uint32_t i = first;
        while (i != (uint32_t)END) {
          IdxT jn = nxt[i];
          if (rk[i] == m) {
            uint32_t j = nxt[i];
            IdxT ip = prv[i];
            jn = nxt[j];
            tok[i] = (TokT)c;
            nxt[i] = jn;
            if (jn != END) prv[jn] = (IdxT)i;
            rk[j] = RK_NONE;
            if (ip != END) rk[ip] = mt_lookup(mt, tok[ip], c);
            rk[i] = (jn != END) ? mt_lookup(mt, c, tok[jn]) : RK_NONE;  // refreshed again if jn merges next
          }
          i = (jn != END) ? (uint32_t)jn : (uint32_t)END;
        }
      }
      __syncwarp();
    }
  }
}

// One warp per document.  out_tmp holds each document's tokens at the document's INPUT offset.
__global__ void __launch_bounds__(ENC_THREADS) k_encode(const int32_t* __restrict__ ids,
                                                         const int64_t* __restrict__ doc_off, int64_t n_docs,
                                                         MergeTable mt, int32_t* __restrict__ out_tmp,
                                                         uint32_t* __restrict__ out_len, uint32_t* __restrict__ g_tok,
                                                         uint32_t* __restrict__ g_nxt, uint32_t* __restrict__ g_prv,
                                                         uint32_t* __restrict__ g_rk) {
  __shared__ uint16_t s_tok[ENC_WARPS][ENC_WARP_MAX];
  __shared__ uint16_t s_nxt[ENC_WARPS][ENC_WARP_MAX];
  __shared__ uint16_t s_prv[ENC_WARPS][ENC_WARP_MAX];
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
    uint32_t wr = 0;
    if (n <= ENC_WARP_MAX) {
      for (uint32_t i = lane; i < n; i += 32) s_tok[warp][i] = (uint16_t)__ldg(src + i);
      __syncwarp();
      encode_doc<uint16_t, uint16_t>(s_tok[warp], s_nxt[warp], s_prv[warp], s_rk[warp], n, mt, lane);
      // survivors in position order: position 0 always survives; i survives iff it is still linked
      for (uint32_t base = 0; base < n; base += 32) {
        uint32_t i = base + lane;
        bool keep = false;
        if (i < n) keep = (i == 0) || (s_prv[warp][i] != 0xFFFFu && s_nxt[warp][s_prv[warp][i]] == i);
        uint32_t K = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) dst[wr + __popc(K & ((1u << lane) - 1u))] = (int32_t)s_tok[warp][i];
        wr += __popc(K);
      }
    } else {
      uint32_t* tok = g_tok + (s - base0);
      uint32_t* nx = g_nxt + (s - base0);
      uint32_t* pv = g_prv + (s - base0);
      uint32_t* rk = g_rk + (s - base0);
      for (uint32_t i = lane; i < n; i += 32) tok[i] = (uint32_t)__ldg(src + i);
      __syncwarp();
      encode_doc<uint32_t, uint32_t>(tok, nx, pv, rk, n, mt, lane);
      for (uint32_t base = 0; base < n; base += 32) {
        uint32_t i = base + lane;
        bool keep = false;
        if (i < n) keep = (i == 0) || (pv[i] != 0xFFFFFFFFu && nx[pv[i]] == i);
        uint32_t K = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) dst[wr + __popc(K & ((1u << lane) - 1u))] = (int32_t)tok[i];
        wr += __popc(K);
      }
    }
    if (lane == 0) out_len[d] = wr;
    __syncwarp();
  }
}

// K5: gather each document's tokens to its final offset, mapping through to_vector_index
// (core.ts:434-442); a hole (-1) marks the document's first_bad and is emitted as -(index+1).
__global__ void __launch_bounds__(128) k_gather_map(const int32_t* __restrict__ out_tmp,
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
