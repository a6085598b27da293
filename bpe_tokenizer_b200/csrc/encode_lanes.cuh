// encode_lanes.cuh -- K4, fast path: lane-sequential rounds with exact stability scans.
//
// Reference: encodeToCode applies every merge in training order with replaceAll over the whole string
// (core.ts:404-406).  Here one WARP owns a batch of whole documents (<= 32*L tokens) laid out over its 32 lanes:
// lane l holds the contiguous segment [l*L, (l+1)*L) in shared memory and walks it sequentially; lanes exchange only
// clamp composites (two warp-shuffle scans) and a few boundary flags per round.  No block-level synchronisation.
//
// Every pair record j = (tok_j, rk_j, rs_j, ls_j) comes from ONE probe of the pair table with (tok_j, tok_j+1):
//   rk  rank of the rule (tok_j, tok_j+1), NONE when there is none
//   rs  lowest rank of any rule (y, tok_j+1) where tok_j lies on the RIGHT spine of y's merge tree (y = tok_j
//       included): whatever is ever built ending in tok_j cannot take tok_j+1 from its left before time rs
//   ls  lowest rank of any rule (tok_j, z) where tok_j+1 lies on the LEFT spine of z
// Stability values, all on the state at the start of the round (sequential process = core.ts:404-406):
//   T[j]    = min(rk_j, SL[j])            SL[j+1] = max(rs_j, T[j])          SL = INF at a document start
//   SR[j]   = max(ls_j, min(rk_j, SR[j+1]))                                   SR = INF at a document end
//   SL[j] >= r: token j is certainly not consumed from its left strictly before time r (SR: from its right)
// A pair j of rank r is merged by the sequential process exactly as it stands iff r <= SL[j] and r <= SR[j+1]; a
// pair (x,x) follows the replaceAll parity of its run (core.ts:285-290) and needs r <= T[s-1] at the run start s (the
// run must not grow at its left end before time r).  The lowest rank present always qualifies, so every round makes
// progress; with the spine bounds the number of rounds is close to the depth of the merge trees (~8 for 32k merges).
// SL and SR are scans of clamp functions x -> max(lo, min(hi, x)), which compose into clamps.
// Prototype, fuzzed against the literal oracle: tests/proto/proto_encode_lanes.py.
#pragma once
#include "common.cuh"

namespace bpe {

constexpr uint32_t EL_INF = 0xFFFFu, EL_NONE = 0xFFFFu, EL_BOUNDARY = 0xFFFEu, EL_DIRTY = 0xFFFDu;
constexpr uint32_t EL_MAX_RANK = 0xFFFCu;  // ranks 0..EL_MAX_RANK
constexpr int EL_DENSE = 32;               // pairs of ids < EL_DENSE are looked up in a dense shared-memory table
constexpr uint32_t EL_LONG = 0xFFFFFFFFu;  // out_len marker: document too long for the lane path

struct LaneTables {
  const uint4* ent;        // x = pair_key, y = rk | c << 16, z = rs | ls << 16
  uint32_t mask, shift;
  const uint2* dense;      // [EL_DENSE * EL_DENSE]: x = rk | c << 16, y = rs | ls << 16
  const uint16_t* rule_c;  // [n_merges] token produced by the rule of each rank
  int32_t c_affine;        // >= 0: rule_c[r] == c_affine + r for every rank (the usual case: no load needed)
};

// (rk | c << 16, rs | ls << 16) of the pair (a, b)
__device__ __forceinline__ uint2 el_lookup(const LaneTables& T, const uint2* s_dense, uint32_t a, uint32_t b) {
  if (a < EL_DENSE && b < EL_DENSE) return s_dense[a * EL_DENSE + b];
  uint32_t key = pair_key(a, b);
  uint32_t h = (key * 0x9E3779B1u) >> T.shift;
  for (;;) {
    uint4 e = __ldg(T.ent + h);
    if (e.x == key) return make_uint2(e.y, e.z);
    if (e.x == EMPTY_KEY) return make_uint2(EL_NONE, 0xFFFFFFFFu);
    h = (h + 1) & T.mask;
  }
}

// clamp functions packed as lo | hi << 16, evaluated on both halves at once with the packed u16x2 min/max of sm_100
// (VIMNMX.U16x2): outer o inner = (outer(inner.lo), outer(inner.hi))
__device__ __forceinline__ uint32_t cl_apply(uint32_t f, uint32_t x) { return max(f & 0xFFFFu, min(f >> 16, x)); }
__device__ __forceinline__ uint32_t cl_compose(uint32_t outer, uint32_t inner) {
  return __vmaxu2(__byte_perm(outer, 0, 0x1010), __vminu2(__byte_perm(outer, 0, 0x3232), inner));
}
constexpr uint32_t CL_IDENT = 0xFFFF0000u;
// step of one record: lo | max(rk, lo) << 16, with a = tok | rk << 16
__device__ __forceinline__ uint32_t cl_make2(uint32_t lo2 /* lo | lo << 16 */, uint32_t a) { return __vmaxu2(lo2, a & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t cl_make(uint32_t lo, uint32_t rk) { return lo | (max(rk, lo) << 16); }

// index of the range's first document: lower bound of base0 + k*stride in doc_off[0..n_docs]
__global__ void k_range_starts(const int64_t* __restrict__ doc_off, int64_t n_docs, uint32_t stride, uint32_t n_ranges,
                               uint32_t* __restrict__ range_first) {
  int64_t base0 = doc_off[0];
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k <= n_ranges; k += gridDim.x * blockDim.x) {
    if (k == n_ranges) {
      range_first[k] = (uint32_t)n_docs;
      continue;
    }
    int64_t target = base0 + (int64_t)k * stride;
    int64_t lo = 0, hi = n_docs;  // first d with doc_off[d] >= target
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (doc_off[mid] < target) lo = mid + 1;
      else hi = mid;
    }
    range_first[k] = (uint32_t)lo;
  }
}

template <int LMAX>
struct DirtyMask {
  typedef unsigned long long type;
  static __device__ __forceinline__ int popc(type m) { return __popcll(m); }
  static __device__ __forceinline__ int ffs(type m) { return __ffsll((long long)m); }
};
template <>
struct DirtyMask<32> {
  typedef uint32_t type;
  static __device__ __forceinline__ int popc(type m) { return __popc(m); }
  static __device__ __forceinline__ int ffs(type m) { return __ffs((int)m); }
};

template <int LMAX, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_encode_lanes(const int32_t* __restrict__ ids,
                                                              const int64_t* __restrict__ doc_off, int64_t n_docs,
                                                              const uint32_t* __restrict__ range_first, uint32_t n_ranges,
                                                              LaneTables T, int32_t* __restrict__ out_tmp,
                                                              uint32_t* __restrict__ out_len, uint32_t* __restrict__ n_long,
                                                              uint32_t* __restrict__ err, uint32_t* __restrict__ next_range) {
  constexpr uint32_t CAP = 32u * LMAX;
  extern __shared__ uint32_t el_smem[];
  uint2* s_dense = reinterpret_cast<uint2*>(el_smem);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t* sA = el_smem + 2 * EL_DENSE * EL_DENSE + warp * (LMAX * 32 * 2 + LMAX * 16);  // tok | rk << 16
  uint32_t* sB = sA + LMAX * 32;                                                        // rs | ls << 16
  uint16_t* sS = reinterpret_cast<uint16_t*>(sB + LMAX * 32);                           // SR (and u16 staging of the ids)
  for (int i = threadIdx.x; i < EL_DENSE * EL_DENSE; i += WARPS * 32) s_dense[i] = T.dense[i];
  __syncthreads();
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t gwarp = blockIdx.x * WARPS + warp, nwarps = gridDim.x * WARPS;

  // ranges (spans of consecutive documents, several batches each) are claimed dynamically: the first nwarps statically
  for (uint32_t k = gwarp;;) {
    if (k >= n_ranges) break;
    int64_t d = range_first[k];
    const int64_t d1 = range_first[k + 1];
    while (d < d1) {
      // ---- form a batch [d, de): consecutive non-empty documents, total <= CAP tokens ----
      const int64_t base = doc_off[d];
      int64_t de = d;
      uint32_t n = 0;
      for (;;) {
        int64_t dd = de + lane;
        int64_t e_i = (dd < d1) ? doc_off[dd + 1] : (int64_t)-1;
        int64_t s_i = __shfl_up_sync(FULL, e_i, 1);
        if (lane == 0) s_i = base + n;
        bool ok = (dd < d1) && e_i > s_i && (uint64_t)(e_i - base) <= CAP;
        uint32_t okm = __ballot_sync(FULL, ok);
        uint32_t nb = (okm == FULL) ? 32u : (uint32_t)(__ffs(~okm) - 1);
        if (nb) n = (uint32_t)(__shfl_sync(FULL, e_i, nb - 1) - base);
        de += nb;
        if (nb < 32u) break;
      }
      if (de == d) {  // the first document is empty or too long for the lane path
        if (lane == 0) {
          int64_t len = doc_off[d + 1] - base;
          if (len == 0) out_len[d] = 0;
          else {
            out_len[d] = EL_LONG;
            atomicAdd(n_long, 1u);
          }
        }
        d++;
        continue;
      }
      // ---- load: coalesced ids -> u16 staging -> lane segments ----
      const uint32_t L = (n + 31u) >> 5;
      __syncwarp();
      for (uint32_t p = lane; p < n; p += 32) sS[p] = (uint16_t)__ldg(ids + base + p);
      __syncwarp();
      uint32_t head = 0;
      uint32_t cnt = (lane * L < n) ? min(L, n - lane * L) : 0u;
      for (uint32_t j = 0; j < cnt; j++) sA[j * 32 + lane] = (uint32_t)sS[lane * L + j] | (EL_DIRTY << 16);
      __syncwarp();
      for (int64_t dd = d + lane; dd < de; dd += 32) {  // last token of every document: no pair to its right
        uint32_t p = (uint32_t)(doc_off[dd + 1] - base) - 1u;
        uint32_t idx = (p % L) * 32 + p / L;
        sA[idx] = (sA[idx] & 0xFFFFu) | (EL_BOUNDARY << 16);
        sB[idx] = 0xFFFFFFFFu;
      }
      __syncwarp();

      // ---- initial ranks: every record but the document ends is DIRTY; each lane resolves its own ----
      uint32_t nonempty = __ballot_sync(FULL, cnt > 0);
      int nlane = (nonempty >> lane) >> 1 ? (int)lane + __ffs((nonempty >> lane) >> 1) : -1;
      uint32_t firstA = cnt ? sA[lane] : 0u;
      uint32_t nf_tok = __shfl_sync(FULL, firstA, nlane < 0 ? (int)lane : nlane) & 0xFFFFu;
      for (uint32_t j = 0; j < cnt; j++) {
        uint32_t idx = j * 32 + lane;
        uint32_t a = sA[idx];
        if ((a >> 16) == EL_DIRTY) {
          uint32_t nt = (j + 1 < cnt) ? (sA[idx + 32] & 0xFFFFu) : nf_tok;
          uint2 r2 = el_lookup(T, s_dense, a & 0xFFFFu, nt);
          sA[idx] = (a & 0xFFFFu) | ((r2.x & 0xFFFFu) << 16);
          sB[idx] = r2.y;
        }
      }
      typedef typename DirtyMask<(LMAX <= 32 ? 32 : 64)>::type mask_t;
      typedef DirtyMask<(LMAX <= 32 ? 32 : 64)> DM;
      mask_t dm = 0;  // rows of this lane whose record is DIRTY (set by pass 3, cleared by the probe phase)

      // ---- rounds ----
      for (uint32_t round = 0;; round++) {
        // composites F (all records but the last) and G (all records) of this lane's segment
        uint32_t F = CL_IDENT, G = CL_IDENT;
        bool anyvalid = false;
        for (uint32_t j = 0; j < cnt; j++) {
          uint32_t idx = (head + j) * 32 + lane;
          uint32_t a = sA[idx], b = sB[idx];
          uint32_t rk = a >> 16;
          anyvalid |= rk < EL_DIRTY;
          G = cl_compose(G, cl_make2(__byte_perm(b, 0, 0x3232), a));
          if (j + 1 < cnt) F = cl_compose(cl_make2(__byte_perm(b, 0, 0x1010), a), F);
        }
        if (!__any_sync(FULL, anyvalid)) break;
        if (round > n + 8u) {  // every round merges at least one pair
          if (lane == 0) atomicOr(err, 1u);
          break;
        }
        // neighbour records (pre-round state)
        int plane = (nonempty & lt_mask) ? 31 - __clz(nonempty & lt_mask) : -1;
        uint32_t lastA = cnt ? sA[(head + cnt - 1) * 32 + lane] : 0u;
        uint32_t lastB = cnt ? sB[(head + cnt - 1) * 32 + lane] : 0u;
        firstA = cnt ? sA[head * 32 + lane] : 0u;
        uint32_t pA = __shfl_sync(FULL, lastA, plane < 0 ? (int)lane : plane);
        uint32_t pB = __shfl_sync(FULL, lastB, plane < 0 ? (int)lane : plane);
        uint32_t nA = __shfl_sync(FULL, firstA, nlane < 0 ? (int)lane : nlane);
        if (nlane < 0) nA = EL_BOUNDARY << 16;
        // x_in = SL of the previous non-empty lane's last token: exclusive prefix scan of the shifted composites
        uint32_t Fs = CL_IDENT;
        if (cnt) Fs = (plane >= 0) ? cl_compose(F, cl_make(pB & 0xFFFFu, pA >> 16)) : F;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t other = __shfl_up_sync(FULL, Fs, o);
          if ((int)lane >= o) Fs = cl_compose(Fs, other);
        }
        uint32_t x_in = __shfl_up_sync(FULL, Fs, 1) >> 16;  // composite applied to INF
        if (lane == 0) x_in = EL_INF;
        // y_in = SR of the next non-empty lane's first token: exclusive suffix scan
        uint32_t Gs = G;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t other = __shfl_down_sync(FULL, Gs, o);
          if ((int)lane + o < 32) Gs = cl_compose(Gs, other);
        }
        uint32_t y_in = __shfl_down_sync(FULL, Gs, 1) >> 16;
        if (lane == 31) y_in = EL_INF;
        // pass 2: SR values, right to left
        {
          uint32_t y = y_in;
          for (uint32_t j = cnt; j-- > 0;) {
            uint32_t idx = (head + j) * 32 + lane;
            uint32_t a = sA[idx], b = sB[idx];
            y = max(b >> 16, min(a >> 16, y));
            sS[idx] = (uint16_t)y;
          }
        }
        // pass 3: decide + compact, left to right (straight-line, predicated)
        uint32_t sl, g, prk;
        if (plane < 0) {
          sl = EL_INF;
          g = EL_INF;
          prk = EL_BOUNDARY;
        } else {
          prk = pA >> 16;
          g = (prk == EL_BOUNDARY) ? EL_INF : min(prk, x_in);
          sl = (prk == EL_BOUNDARY) ? EL_INF : max(pB & 0xFFFFu, g);
        }
        uint32_t par = 0, w = 0;
        bool run_ok = false, consumed = false, took_straddle = false, first_is_new = false, have_pend = false;
        uint32_t pendA = 0, pendB = 0;
        uint32_t a = cnt ? sA[head * 32 + lane] : 0u;
        for (uint32_t j = 0; j < cnt; j++) {
          uint32_t idx = (head + j) * 32 + lane;
          uint32_t b = sB[idx];
          bool last = (j + 1 == cnt);
          uint32_t an = last ? nA : sA[idx + 32];
          uint32_t srn = last ? y_in : (uint32_t)sS[idx + 32];
          uint32_t r = a >> 16, rs = b & 0xFFFFu;
          bool valid = r < EL_DIRTY;
          bool cont = valid && r == prk;          // same rule as the pair to the left: inside a run x x x
          bool isxx = ((a ^ an) & 0xFFFFu) == 0;
          bool ok_new = r <= (isxx ? g : sl);     // a run start also needs T[j-1] >= r (the run must not grow leftwards)
          par = cont ? (par ^ 1u) : 0u;
          run_ok = cont ? (run_ok && j != 0) : ok_new;  // j == 0: run carried in from the previous lane, parity unknown
          bool take = valid && !consumed && r <= srn && (cont ? (par == 0 && run_ok) : ok_new);
          bool emit = !consumed;
          if (take && have_pend && (pendA >> 16) != EL_BOUNDARY) pendA = (pendA & 0xFFFFu) | (EL_DIRTY << 16);
          first_is_new |= take && !have_pend;
          if (emit && have_pend) {
            sA[w * 32 + lane] = pendA;
            sB[w * 32 + lane] = pendB;
            if ((pendA >> 16) == EL_DIRTY) dm |= (mask_t)1 << w;
            w++;
          }
          if (emit) {
            have_pend = true;
            uint32_t c = (T.c_affine >= 0) ? (uint32_t)T.c_affine + r : (take ? (uint32_t)__ldg(T.rule_c + r) : 0u);
            uint32_t nrk = ((an >> 16) == EL_BOUNDARY) ? EL_BOUNDARY : EL_DIRTY;
            pendA = take ? (c | (nrk << 16)) : a;
            pendB = take ? 0xFFFFFFFFu : b;
            took_straddle = take && last;
          }
          consumed = take;
          g = (r == EL_BOUNDARY) ? EL_INF : min(r, sl);
          sl = max(rs, g);
          prk = r;
          a = an;
        }
        if (have_pend) {
          sA[w * 32 + lane] = pendA;
          sB[w * 32 + lane] = pendB;
          if ((pendA >> 16) == EL_DIRTY) dm |= (mask_t)1 << w;
          w++;
        }
        // post: first tokens consumed by the lane to the left; ranks next to new tokens across lanes
        bool eaten = __shfl_sync(FULL, took_straddle, plane < 0 ? (int)lane : plane) && plane >= 0 && cnt > 0;
        head = eaten ? 1u : 0u;
        cnt = eaten ? w - 1u : w;
        if (eaten) dm &= ~(mask_t)1;
        nonempty = __ballot_sync(FULL, cnt > 0);
        nlane = (nonempty >> lane) >> 1 ? (int)lane + __ffs((nonempty >> lane) >> 1) : -1;
        bool fin = __shfl_sync(FULL, first_is_new, nlane < 0 ? (int)lane : nlane) && nlane >= 0;
        uint32_t lastrow = head + cnt - 1u;
        if (fin && cnt) {
          uint32_t idx = lastrow * 32 + lane;
          uint32_t v = sA[idx];
          if ((v >> 16) != EL_BOUNDARY) {
            sA[idx] = (v & 0xFFFFu) | (EL_DIRTY << 16);
            dm |= (mask_t)1 << lastrow;
          }
        }
        // ---- probe phase: the warp resolves all DIRTY records together (32 independent table loads in flight) ----
        firstA = cnt ? sA[head * 32 + lane] : 0u;
        nf_tok = __shfl_sync(FULL, firstA, nlane < 0 ? (int)lane : nlane) & 0xFFFFu;
        uint32_t nd = (uint32_t)DM::popc(dm), qpos = nd;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t other = __shfl_up_sync(FULL, qpos, o);
          if ((int)lane >= o) qpos += other;
        }
        uint32_t Q = __shfl_sync(FULL, qpos, 31);
        qpos -= nd;
        __syncwarp();
        while (dm) {
          uint32_t row = (uint32_t)DM::ffs(dm) - 1u;
          dm &= dm - (mask_t)1;
          sS[qpos++] = (uint16_t)(row * 32 + lane);
        }
        __syncwarp();
        for (uint32_t q0 = 0; q0 < Q; q0 += 32) {
          uint32_t q = q0 + lane;
          bool act = q < Q;
          uint32_t cell = act ? (uint32_t)sS[q] : lane;
          uint32_t owner = cell & 31u;
          uint32_t o_last = __shfl_sync(FULL, lastrow, owner);
          uint32_t o_nf = __shfl_sync(FULL, nf_tok, owner);
          if (act) {
            uint32_t av = sA[cell];
            uint32_t nt = ((cell >> 5) == o_last) ? o_nf : (sA[cell + 32] & 0xFFFFu);
            uint2 r2 = el_lookup(T, s_dense, av & 0xFFFFu, nt);
            sA[cell] = (av & 0xFFFFu) | ((r2.x & 0xFFFFu) << 16);
            sB[cell] = r2.y;
          }
        }
        __syncwarp();
      }

      // ---- output: token k of document dd -> out_tmp[doc_off[dd] + k], out_len[dd] ----
      {
        uint32_t nb = 0, tail = 0;  // documents ending in this lane; tokens after the last such end
        for (uint32_t j = 0; j < cnt; j++) {
          uint32_t v = sA[(head + j) * 32 + lane];
          tail++;
          if ((v >> 16) == EL_BOUNDARY) {
            nb++;
            tail = 0;
          }
        }
        // exclusive prefix: ordinal of the document open at my first token, and its tokens so far
        uint32_t ord = nb, v = tail, f = nb ? 1u : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t ord_o = __shfl_up_sync(FULL, ord, o);
          uint32_t v_o = __shfl_up_sync(FULL, v, o);
          uint32_t f_o = __shfl_up_sync(FULL, f, o);
          if ((int)lane >= o) {
            ord += ord_o;
            if (!f) v += v_o;
            f |= f_o;
          }
        }
        uint32_t ord0 = __shfl_up_sync(FULL, ord, 1);
        uint32_t pos = __shfl_up_sync(FULL, v, 1);
        if (lane == 0) {
          ord0 = 0;
          pos = 0;
        }
        int64_t dd = d + ord0;
        int64_t obase = (cnt && dd < de) ? doc_off[dd] : 0;
        for (uint32_t j = 0; j < cnt; j++) {
          uint32_t v2 = sA[(head + j) * 32 + lane];
          out_tmp[obase + pos] = (int32_t)(v2 & 0xFFFFu);
          pos++;
          if ((v2 >> 16) == EL_BOUNDARY) {
            out_len[dd] = pos;
            dd++;
            pos = 0;
            if (j + 1 < cnt) obase = doc_off[dd];
          }
        }
      }
      __syncwarp();
      d = de;
    }
    if (lane == 0) k = nwarps + atomicAdd(next_range, 1u);
    k = __shfl_sync(FULL, k, 0);
  }
}

}  // namespace bpe
