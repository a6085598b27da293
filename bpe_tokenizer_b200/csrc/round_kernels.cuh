// round_kernels.cuh -- mergeUntil (core.ts:365-383) as ROUNDS of several exact merges per pair of grid barriers.
//
// k_merge_loop (train_kernels.cuh) pays ~26 us of dependent L2/DRAM round trips and two grid barriers for EVERY merge,
// however small.  93 % of the merges of a 1 GB run are that small, and the next winner is almost always the runner-up of
// the current decision (tests/proto/proto_batch_stats.cpp: with the rules below a decision of the 1 GB-style Zipf corpus
// yields 6.3 exact merges on average when up to 8 are tried, 8.8 when up to 16).  This kernel therefore takes the best K
// pairs of ONE decision and runs their site passes side by side:
//
//   decide      every block folds the per-block top-2 partials into the exact, strictly ordered list of the best pairs
//               (exact down to the largest "second best" any block published) and builds the same batch m_0 .. m_{k-1}:
//               strictly decreasing (count, -(a+b)) keys, no tie inside the batch, pairwise disjoint tokens, no token born in
//               the previous round (its occurrence lists are still being written), capacities.
//   P1          site passes of all k merges over the corpus as it was at the decision.  A merge changes counts only of
//               pairs that contain a, b or c, so ALL its count deltas live in dense per-token rows (no hashing in P1):
//               DEC_L[x] = decrements of (x,a), DEC_R[y] of (b,y), NEW_L[x] / NEW_R[y] = occurrences and counted
//               occurrences of the born pairs (x,c) / (c,y).  Token sets are disjoint, so the sites of m_j are the same
//               before and after m_0..m_{j-1}; only a NEIGHBOUR of a site can have been rewritten by an earlier merge of the
//               batch, and the site pass of m_j looks for exactly that: a left neighbour b_i preceded by a_i, or a right
//               neighbour a_i followed by b_i (i < j), is the token c_i there (virtual neighbour), its pair with a_j / b_j
//               is the born pair (c_i, a_j) / (b_j, c_i) of merge i, and the new adjacency is (c_i, c_j) / (c_j, c_i).
//               Per merge and side the pass also sums, over its warps, the largest number of lanes that share one
//               neighbour: an upper bound U of the count of ANY pair born by that merge (fire-and-forget adds).
//   -- barrier --
//   P2          every block computes the same valid prefix v: m_j is the exact next winner iff its count did not move
//               (true by token disjointness), every older pair outside the batch ranks below it (the list was exact) and no
//               pair born by m_0..m_{j-1} reaches its count (U_i < W_j for all i < j -- conservative).  Merges >= v are
//               dropped: they only ever wrote to their own rows.  For the v valid merges, side by side: the born pairs
//               enter the table (count minus what later valid merges took from them), the decrements are applied -- one
//               thread per (pair), which hands the pair's new key to the arg-max --, the corpus is rewritten, and the
//               arg-max runs over the old hot pairs that no valid merge touches.  -> per-block top-2 partials.
//   -- barrier --
//   The occurrence lists of the born pairs are filled next to P1 of the following round (as in k_merge_loop), and the rows
//   of a round are zeroed by the scan of the following one (rows are double-buffered by round parity).
//
// Exactness argument for the order inside a batch (core.ts:294-305): counts of existing pairs never grow
// under merging, so an old pair that ranked below m_j at the decision still does; born pairs are bounded by U; m_j's own
// count is unchanged; equal keys never enter a batch (position tie-breaks run alone, through the path k_merge_loop uses).
#pragma once
#include "train_kernels.cuh"

namespace bpe {

constexpr int RB = 16;                   // merges per round at most
constexpr uint32_t R_SMALL = 16384;      // a merge with more counted occurrences than this runs alone
enum { RW_DEC_L = 0, RW_DEC_R, RW_NL_LEN, RW_NL_CNT, RW_NR_LEN, RW_NR_CNT, RW_NL_SLOT, RW_NR_SLOT, RW_ROWS };
constexpr int RW_CLEAR_ROWS = 6;         // rows that accumulate (the SLOT rows are written before they are read)
constexpr uint32_t R_QCAP = 640;         // per-block top-2 entries a decision can fold (2 x blocks)
constexpr uint32_t ERR_ROUND_MISMATCH = 2048u;  // a merge of a round found a different number of sites than its count

struct RoundState {
  uint32_t n_sites[2][RB];
  uint32_t ub[2][RB][2];  // U bounds (left / right born pairs) per merge
  unsigned long long rounds, round_merges, rounds_cut_born, rounds_single, tried;
  unsigned long long stop_reason[8];  // why a batch was not extended: 0 cap, 1 no exact candidate, 2 tie, 3 big, 4 token, 5 fresh token, 6 limits
};

struct RoundArgs {
  LoopArgs L;
  uint32_t* rows;    // [2][RB][RW_ROWS][ND_STRIDE]
  SiteRec* bsites;   // [2][RB][R_SMALL]: merges 1.. of a round (merge 0 uses L.A.sites / L.sites2, which the host sizes)
  uint4* gp;         // [2 * blocks] per-block top-2: (primary lo, primary hi, slot, mult)
  uint32_t* gk;      // [2 * blocks] ... and the pair key of that slot
  RoundState* rs;
  uint32_t kmax;     // merges per round (1 .. RB)
};

__device__ __forceinline__ uint32_t* round_row(const RoundArgs& R, uint32_t par, uint32_t j, int row) {
  return R.rows + (((size_t)par * RB + j) * RW_ROWS + (size_t)row) * ND_STRIDE;
}

// ---- top-2 groups: the two largest DISTINCT primaries, each with the number of pairs that share it and the smallest slot ----
struct Top2 {
  unsigned long long p0, p1;
  uint32_t s0, s1, m0, m1, k0, k1;
};
__device__ __forceinline__ Top2 top2_empty() { return Top2{0ull, 0ull, NOSLOT, NOSLOT, 0u, 0u, 0u, 0u}; }
__device__ __forceinline__ void top2_add(Top2& t, unsigned long long p, uint32_t slot, uint32_t mult, uint32_t key) {
  if (!p) return;
  if (p > t.p0) {
    t.p1 = t.p0; t.s1 = t.s0; t.m1 = t.m0; t.k1 = t.k0;
    t.p0 = p; t.s0 = slot; t.m0 = mult; t.k0 = key;
  } else if (p == t.p0) {
    t.m0 += mult;
    if (slot < t.s0) { t.s0 = slot; t.k0 = key; }
  } else if (p > t.p1) {
    t.p1 = p; t.s1 = slot; t.m1 = mult; t.k1 = key;
  } else if (p == t.p1) {
    t.m1 += mult;
    if (slot < t.s1) { t.s1 = slot; t.k1 = key; }
  }
}
__device__ __forceinline__ Top2 top2_warp_reduce(Top2 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long q0 = __shfl_xor_sync(0xFFFFFFFFu, v.p0, o), q1 = __shfl_xor_sync(0xFFFFFFFFu, v.p1, o);
    uint32_t s0 = __shfl_xor_sync(0xFFFFFFFFu, v.s0, o), s1 = __shfl_xor_sync(0xFFFFFFFFu, v.s1, o);
    uint32_t m0 = __shfl_xor_sync(0xFFFFFFFFu, v.m0, o), m1 = __shfl_xor_sync(0xFFFFFFFFu, v.m1, o);
    uint32_t k0 = __shfl_xor_sync(0xFFFFFFFFu, v.k0, o), k1 = __shfl_xor_sync(0xFFFFFFFFu, v.k1, o);
    top2_add(v, q0, s0, m0, k0);
    top2_add(v, q1, s1, m1, k1);
  }
  return v;
}

// shared memory of one block of k_merge_rounds
struct RoundSm {
  // the batch of the current round (same in every block)
  uint32_t k, status, mult0, pad0;
  uint32_t a[RB], b[RB], w[RB], slot[RB], lstart[RB], llen[RB], lenc[RB];
  uint32_t iter0[RB + 1];  // warp-iterations of P1: merge j owns [iter0[j], iter0[j+1])
  uint32_t fill_n[RB];     // sites of the previous round's merges whose born adjacencies are not in their lists yet
  // decision scratch
  uint32_t qn;
  unsigned long long qp[R_QCAP];
  uint32_t qs[R_QCAP], qm[R_QCAP], qk[R_QCAP];
  unsigned long long cp[RB];
  uint32_t cs[RB], cm[RB], ck[RB], cls[RB], cll[RB];
  Top2 t2[32];
  unsigned long long red[32];
  uint32_t s_max[32];
  uint32_t rslot;
};

__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long u = __shfl_xor_sync(0xFFFFFFFFu, v, o);
    v = u > v ? u : v;
  }
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const uint32_t nw = blockDim.x >> 5;
  unsigned long long u = lane_id() < nw ? red[lane_id()] : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long x = __shfl_xor_sync(0xFFFFFFFFu, u, o);
    u = x > u ? x : u;
  }
  return u;
}

__device__ __forceinline__ Top2 top2_block_reduce(Top2 v, Top2* s_t2) {
  v = top2_warp_reduce(v);
  __syncthreads();
  if (lane_id() == 0) s_t2[threadIdx.x >> 5] = v;
  __syncthreads();
  const uint32_t nw = blockDim.x >> 5;
  Top2 u = lane_id() < nw ? s_t2[lane_id()] : top2_empty();
  return top2_warp_reduce(u);
}

// all 32 lanes call; `has` lanes add one to row[tok]; lanes that agree on tok elect a leader which issues ONE atomic.
// Returns, in the leader lane, the number of lanes it stands for (0 elsewhere).
__device__ __forceinline__ uint32_t row_add_warp(uint32_t* row, uint32_t tok, bool has) {
  const uint32_t lane = lane_id();
  const uint32_t peers = __match_any_sync(0xFFFFFFFFu, has ? tok : (0xFFFFFF00u + lane));
  const bool leader = has && lane == (uint32_t)(__ffs(peers) - 1);
  if (!leader) return 0;
  const uint32_t n = (uint32_t)__popc(peers);
  atomicAdd(row + tok, n);
  return n;
}

// the born pair (tok, c) / (c, tok): occurrences and counted occurrences; returns the leader's counted lanes (0 elsewhere)
__device__ __forceinline__ uint32_t row_new_warp(uint32_t* len_row, uint32_t* cnt_row, uint32_t tok, bool has, bool counted) {
  const uint32_t lane = lane_id();
  const uint32_t peers = __match_any_sync(0xFFFFFFFFu, has ? tok : (0xFFFFFF00u + lane));
  const uint32_t cmask = __ballot_sync(0xFFFFFFFFu, has && counted);
  const bool leader = has && lane == (uint32_t)(__ffs(peers) - 1);
  if (!leader) return 0;
  atomicAdd(len_row + tok, (uint32_t)__popc(peers));
  const uint32_t nc = (uint32_t)__popc(peers & cmask);
  if (nc) atomicAdd(cnt_row + tok, nc);
  return nc;
}

// One warp-iteration of the site pass of merge j of the round: 32 entries of its occurrence list.
// The logic of the neighbourhoods is phase_sites' (train_kernels.cuh); what differs is where the deltas go (the merge's
// dense rows) and the virtual neighbours c_i of the earlier merges of the batch.
__device__ __forceinline__ void round_sites_iter(const RoundArgs& R, const RoundSm& S, uint32_t par, uint32_t j, uint32_t c_first,
                                                 uint32_t i, SiteRec* site_out, uint32_t sites_cap) {
  const ApplyArgs& A = R.L.A;
  const uint32_t* slots = A.slots;
  const uint32_t n = A.n;
  DevState* st = A.st;
  const uint32_t a = S.a[j], b = S.b[j], c = c_first + j;
  const uint32_t total = S.llen[j];
  const uint32_t lane = lane_id();
  uint32_t* const dec_l = round_row(R, par, j, RW_DEC_L);
  uint32_t* const dec_r = round_row(R, par, j, RW_DEC_R);
  bool site = false;
  uint32_t p = 0, w = 0, q = 0, koff = 0;
  if (i < total) {
    p = A.pool[S.lstart[j] + i];
    w = ld_slot(slots + p);
    if (slot_is_id(w) && slot_val(w) == a && right_token(slots, n, p, &q) == (int)b) {
      site = true;
      if (a == b) {
        koff = run_left(slots, p, w, (int)a);
        if (koff & 1u) site = false;  // overlaps the occurrence to its left (replaceAll is non-overlapping)
      }
    }
  }
  const uint32_t smask = __ballot_sync(0xFFFFFFFFu, site);
  if (!smask) return;
  uint32_t site_base = 0;
  if (lane == (uint32_t)(__ffs(smask) - 1)) site_base = atomicAdd(&R.rs->n_sites[par][j], (uint32_t)__popc(smask));
  SiteRec rec{p, NOPOS, NOTOKV, NOTOKV};
  // ---- adjacency on the left of the new token ----
  uint32_t dec1_tok = 0, new1_tok = 0;
  bool dec1 = false, new1 = false, new1_counted = true;
  if (site) {
    uint32_t lpos;
    int x = left_token(slots, p, w, &lpos);
    if (x != NOTOK) {
      bool chained = false;
      uint32_t chain_j = 0, llpos = NOPOS;
      if (a == b) {
        chained = koff >= 2;
        chain_j = koff >> 1;
        if (chained) left_token(slots, lpos, ld_slot(slots + lpos), &llpos);
      } else if ((uint32_t)x == b) {
        int xx = left_token(slots, lpos, ld_slot(slots + lpos), &llpos);
        if (xx == (int)a) {  // the pair to the left is itself a site: ... a b a b
          chained = true;
          chain_j = 1;
          uint32_t cur = llpos;
          for (;;) {  // index of this site within its chain of back-to-back sites
            uint32_t l1, l2;
            if (left_token(slots, cur, ld_slot(slots + cur), &l1) != (int)b) break;
            if (left_token(slots, l1, ld_slot(slots + l1), &l2) != (int)a) break;
            chain_j++;
            cur = l2;
          }
        }
      }
      new1 = true;
      if (chained) {
        if (a != b) {
          dec1 = true;
          dec1_tok = b;  // the pair (b,a)
        }
        new1_tok = c;  // the pair (c,c): runs of c count every other pair (:285-290)
        new1_counted = (chain_j & 1u) != 0;
        rec.lpos = llpos;
      } else {
        // is the left neighbour the right half of a site of an EARLIER merge of this batch?  Then it is c_i by now.
        int vi = -1;
        uint32_t vpos = NOPOS;
        for (uint32_t i2 = 0; i2 < j; i2++) {
          if ((uint32_t)x != S.b[i2]) continue;
          const uint32_t wl = ld_slot(slots + lpos);
          if (S.a[i2] != S.b[i2]) {
            uint32_t l2;
            if (left_token(slots, lpos, wl, &l2) == (int)S.a[i2]) {
              vi = (int)i2;
              vpos = l2;
            }
          } else if (run_left(slots, lpos, wl, x) & 1u) {  // odd offset inside its run: the second half of a replaced pair
            vi = (int)i2;
            left_token(slots, lpos, wl, &vpos);
          }
          break;  // token sets are disjoint: at most one merge owns x as its b
        }
        if (vi >= 0) {
          dec1 = true;
          dec1_tok = c_first + (uint32_t)vi;  // the born pair (c_i, a) of merge i loses this adjacency
          new1_tok = c_first + (uint32_t)vi;  // ... and (c_i, c) is born
          rec.lpos = vpos;
        } else if ((uint32_t)x == a) {  // (a != b here) the run of a's ending at p loses its last element
          uint32_t Lr = 1 + run_left(slots, p, w, (int)a);
          dec1 = (Lr & 1u) == 0;
          dec1_tok = a;  // the pair (a,a)
          new1_tok = (uint32_t)x;
          rec.lpos = lpos;
        } else {
          dec1 = true;
          dec1_tok = (uint32_t)x;  // the pair (x,a)
          new1_tok = (uint32_t)x;  // the pair (x,c)
          rec.lpos = lpos;
        }
      }
    }
  }
  const uint32_t nc1 = row_new_warp(round_row(R, par, j, RW_NL_LEN), round_row(R, par, j, RW_NL_CNT), new1_tok, new1, new1_counted);
  rec.lslot = new1 ? new1_tok : NOTOKV;

  // ---- adjacency on the right of the new token (left to the next site when that one is chained) ----
  uint32_t dec2_tok = 0, new2_tok = 0;
  bool dec2 = false, new2 = false;
  if (site) {
    uint32_t r;
    int y = right_token(slots, n, q, &r);
    if (y != NOTOK) {
      bool chained_right = false;
      if ((uint32_t)y == a) {
        uint32_t r2;
        chained_right = right_token(slots, n, r, &r2) == (int)b;
      }
      if (!chained_right) {
        int vi = -1;
        for (uint32_t i2 = 0; i2 < j; i2++) {
          if ((uint32_t)y != S.a[i2]) continue;
          uint32_t r2;
          if (right_token(slots, n, r, &r2) == (int)S.b[i2]) vi = (int)i2;  // y starts a site of merge i (for a_i == b_i: offset 0 of its run)
          break;
        }
        if (vi >= 0) {
          dec2 = true;
          dec2_tok = c_first + (uint32_t)vi;  // the born pair (b, c_i) of merge i
          new2_tok = c_first + (uint32_t)vi;  // (c, c_i)
        } else if ((uint32_t)y == b && a != b) {  // the run of b's starting at q loses its first element
          uint32_t Lr = 1 + run_right(slots, n, q, (int)b);
          dec2 = (Lr & 1u) == 0;
          dec2_tok = b;  // the pair (b,b)
          new2_tok = (uint32_t)y;
        } else {
          if (!(a == b && (uint32_t)y == a)) {  // (a,a) itself is zeroed in P2
            dec2 = true;
            dec2_tok = (uint32_t)y;  // the pair (b,y)
          }
          new2_tok = (uint32_t)y;  // the pair (c,y)
        }
        new2 = true;
      }
    }
  }
  row_add_warp(dec_l, dec1_tok, dec1);
  row_add_warp(dec_r, dec2_tok, dec2);
  const uint32_t nc2 = row_new_warp(round_row(R, par, j, RW_NR_LEN), round_row(R, par, j, RW_NR_CNT), new2_tok, new2, true);
  rec.rslot = new2 ? new2_tok : NOTOKV;
  // upper bounds of the born pairs' counts: the largest group of lanes that share a neighbour, summed over the warps
  const uint32_t u1 = __reduce_max_sync(0xFFFFFFFFu, nc1), u2 = __reduce_max_sync(0xFFFFFFFFu, nc2);
  if (lane == 0) {
    if (u1) atomicAdd(&R.rs->ub[par][j][0], u1);
    if (u2) atomicAdd(&R.rs->ub[par][j][1], u2);
  }
  // ---- record the site ----
  const uint32_t base = __shfl_sync(0xFFFFFFFFu, site_base, __ffs(smask) - 1);
  if (site) {
    const uint32_t k = base + __popc(smask & ((1u << lane) - 1u));
    if (k < sites_cap) reinterpret_cast<uint4*>(site_out)[k] = make_uint4(rec.p, rec.lpos, rec.lslot, rec.rslot);
    else atomicOr(&st->err, ERR_SITE_OVERFLOW);
  }
}

// phase_fill with the slot rows of a round's merge
__device__ __forceinline__ void round_fill(const ApplyArgs& A, const uint32_t* lrow, const uint32_t* rrow, const SiteRec* sites, uint32_t n_sites,
                                           uint32_t vt, uint32_t nvt) {
  const PairTable& t = A.t;
  const uint32_t round = (n_sites + 31u) & ~31u;
  for (uint32_t i = vt; i < round; i += nvt) {
    const bool has = i < n_sites;
    const uint4 rv = has ? ld_cg4(reinterpret_cast<const uint4*>(sites) + i) : make_uint4(0, NOPOS, NOTOKV, NOTOKV);
    const uint32_t lslot = (has && rv.z != NOTOKV) ? ld_cg(lrow + rv.z) : NOSLOT;
    const uint32_t rslot = (has && rv.w != NOTOKV) ? ld_cg(rrow + rv.w) : NOSLOT;
    const bool hl = has && lslot != NOSLOT && t.occ_len[lslot];
    const bool hr = has && rslot != NOSLOT && t.occ_len[rslot];
    const uint32_t il = agg_cursor(t, lslot, hl);
    const uint32_t ir = agg_cursor(t, rslot, hr);
    if (hl) A.pool[il] = rv.y;
    if (hr) A.pool[ir] = rv.x;
  }
}

__device__ __forceinline__ SiteRec* round_sites_buf(const RoundArgs& R, uint32_t par, uint32_t j) {
  if (j == 0) return par ? R.L.sites2 : R.L.A.sites;
  return R.bsites + ((size_t)par * RB + j) * R_SMALL;
}

struct RoundFill {  // what the previous round left to do
  uint32_t v, par, k, c_first;  // merges committed / row parity / merges tried (rows to clear) / first token it created
};

// the lists of the pairs born by the previous round's merges (runs on the helper warps next to P1, or on everybody)
__device__ __forceinline__ void round_fill_all(const RoundArgs& R, const RoundSm& S, const RoundFill& F, uint32_t vt, uint32_t nvt) {
  for (uint32_t j = 0; j < F.v; j++)
    if (S.fill_n[j])
      round_fill(R.L.A, round_row(R, F.par, j, RW_NL_SLOT), round_row(R, F.par, j, RW_NR_SLOT), round_sites_buf(R, F.par, j), S.fill_n[j], vt, nvt);
}

// zero the accumulating rows of a finished round (only needed when no later round's scan does it: at kernel exit)
__device__ __forceinline__ void round_clear_rows(const RoundArgs& R, uint32_t par, uint32_t k, uint32_t c_hi, uint32_t vt, uint32_t nvt) {
  const uint32_t T = (c_hi + 31u) & ~31u;
  for (uint32_t j = 0; j < k; j++)
    for (int row = 0; row < RW_CLEAR_ROWS; row++) {
      uint32_t* r = round_row(R, par, j, row);
      for (uint32_t i = vt; i < T; i += nvt) r[i] = 0;
    }
}

__global__ void k_rounds_prepare(RoundState* rs) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int p = 0; p < 2; p++)
      for (int j = 0; j < RB; j++) {
        rs->n_sites[p][j] = 0;
        rs->ub[p][j][0] = rs->ub[p][j][1] = 0;
      }
  }
}

__global__ void __launch_bounds__(ML_THREADS, 1) k_merge_rounds(RoundArgs R) {
  __shared__ RoundSm S;
  const LoopArgs& L = R.L;
  const ApplyArgs& A = L.A;
  DevState* st = A.st;
  RoundState* rs = R.rs;
  const PairTable& t = A.t;
  const uint32_t bid = blockIdx.x, nblk = gridDim.x;
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5, lane = tid & 31u;
  const uint32_t gt = bid * blockDim.x + tid, gn = nblk * blockDim.x;
  unsigned long long epoch = 0;
  const uint32_t n_tokens0 = ld_cg(&st->n_tokens);
  const uint32_t thresh = ld_cg(&st->hot_thresh);
  const bool lead = (bid == 0 && tid == 0);

  // ---- first partials: per-block top-2 over the hot list ----
  if (lead) st->snap_err = st->err;
  {
    Top2 mine = top2_empty();
    const uint32_t hn = ld_cg(&st->hot_n);
    for (uint32_t i = gt; i < hn; i += gn) {
      const uint32_t hs = L.hot[i];
      uint32_t hk;
      const unsigned long long pr = slot_primary_k(t, A.len16, hs, L.max_length, &hk);
      top2_add(mine, pr, hs, 1, hk);
    }
    const Top2 v = top2_block_reduce(mine, S.t2);
    if (tid == 0) {
      R.gp[2 * bid] = make_uint4((uint32_t)v.p0, (uint32_t)(v.p0 >> 32), v.s0, v.m0);
      R.gp[2 * bid + 1] = make_uint4((uint32_t)v.p1, (uint32_t)(v.p1 >> 32), v.s1, v.m1);
      R.gk[2 * bid] = v.k0;
      R.gk[2 * bid + 1] = v.k1;
    }
    if (tid < RB) S.fill_n[tid] = 0;
  }
  grid_barrier(L.barrier, ++epoch * nblk, L.bar_ns);

  const bool prof = lead;
  unsigned long long tp0 = prof ? now_ns() : 0, tp1;
  // block 0's view, per phase (decide, P1, wait, P2, wait, -, -, tie path); fine_ns splits it into rounds of small merges
  // ([0..4], rounds in [5]) and rounds of one big merge ([6..10], rounds in [11])
  uint32_t prof_big = 0;
#define RPROF(i)                                              \
  if (prof) {                                                 \
    tp1 = now_ns();                                           \
    st->prof_ns[i] += tp1 - tp0;                              \
    if ((i) < 5) st->fine_ns[(prof_big ? 6 : 0) + (i)] += tp1 - tp0; \
    tp0 = tp1;                                                \
  }
  RoundFill F{0, 0, 0, 0};
  uint32_t it = 0;  // merges committed by this launch
  for (uint32_t round = 0;; round++) {
    const uint32_t par = round & 1u;
    const uint32_t c_first = n_tokens0 + it;
    const uint32_t hot_pre = ld_cg(&st->hot_n);
    // ================= decide =================
    {
      unsigned long long myp = 0;
      uint32_t mys = NOSLOT, mym = 0, myk = 0;
      if (tid < 2 * nblk) {
        const uint4 v = ld_cg4(R.gp + tid);
        myp = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
        mys = v.z;
        mym = v.w;
        myk = ld_cg(R.gk + tid);
      }
      if (tid == 0) S.qn = 0;
      if (tid < RB) {
        S.cp[tid] = 0;
        S.cm[tid] = 0;
      }
      // the list of candidates is exact down to the largest second-best primary any block published
      const unsigned long long Lcut = block_max_u64((tid & 1u) ? myp : 0ull, S.red);  // (syncs inside: qn / cp are visible)
      if (myp && myp >= Lcut) {
        const uint32_t k = atomicAdd(&S.qn, 1u);
        S.qp[k] = myp;
        S.qs[k] = mys;
        S.qm[k] = mym;
        S.qk[k] = myk;
      }
      __syncthreads();
      const uint32_t qn = S.qn;
      if (tid < qn) {  // rank by counting; entries that share a primary form one group (sum of mults, smallest slot)
        const unsigned long long p = S.qp[tid];
        uint32_t rank = 0, mult = 0, slot = NOSLOT, key = 0;
        bool leader = true;
        for (uint32_t u = 0; u < qn; u++) {
          const unsigned long long pu = S.qp[u];
          rank += pu > p;
          if (pu == p) {
            mult += S.qm[u];
            if (S.qs[u] < slot) {
              slot = S.qs[u];
              key = S.qk[u];
            }
            if (u < tid) leader = false;
          }
        }
        if (leader && rank < RB) {
          S.cp[rank] = p;
          S.cs[rank] = slot;
          S.cm[rank] = mult;
          S.ck[rank] = key;
        }
      }
      __syncthreads();
      if (tid < RB && S.cp[tid]) {  // list of every candidate, loaded side by side
        S.cls[tid] = t.occ_start[S.cs[tid]];
        S.cll[tid] = t.occ_len[S.cs[tid]];
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t status = LOOP_RUNNING, k = 0;
        const unsigned long long p0 = S.cp[0];
        const uint32_t w0 = (uint32_t)(p0 >> 20);
        const uint32_t n_keys = ld_cg(&st->n_keys), pool_cursor = ld_cg(&st->pool_cursor);
        if (ld_cg(&st->snap_err)) status = LOOP_ERROR;
        else if (!p0) status = (thresh <= 1) ? LOOP_EMPTY : LOOP_NEED_REBUILD;
        else if (w0 < thresh) status = LOOP_NEED_REBUILD;
        else if (w0 < L.min_weight) status = LOOP_DONE;  // core.ts:313
        else if (it >= L.log_cap) status = LOOP_LIMIT;
        else if (c_first >= L.max_tokens) status = LOOP_NEED_HOST;
        unsigned long long keys_sum = 0, w_sum = 0;
        int stop = 0;
        if (status == LOOP_RUNNING) {
          for (uint32_t j = 0; j < R.kmax; j++) {
            const unsigned long long pj = S.cp[j];
            if (!pj) { stop = 1; break; }
            const uint32_t wj = (uint32_t)(pj >> 20), key = S.ck[j];
            const uint32_t aj = key >> 16, bj = key & 0xFFFFu, cj = c_first + j;
            const unsigned long long new_keys = min(2ull * wj + 2ull, 2ull * (cj + 1ull) + 2ull);
            bool ok = true;
            uint32_t why = LOOP_NEED_HOST;
            // capacity the host guarantees (k_merge_loop's checks, cumulative over the batch)
            if ((unsigned long long)n_keys + keys_sum + new_keys > (unsigned long long)(L.tbl_cap >> 1)) ok = false;
            else if ((unsigned long long)pool_cursor + 2ull * (w_sum + wj) > L.pool_cap) ok = false;
            else if (j == 0 && wj > A.sites_cap) ok = false;
            else if ((unsigned long long)hot_pre + keys_sum + new_keys > min(L.hot_cap, L.hot_limit)) { ok = false; why = LOOP_NEED_REBUILD; }
            else if (cj + 1 > L.len16_cap) ok = false;
            if (j == 0) {
              if (!ok) { status = why; break; }
              if (S.cm[0] > 1 && S.cm[0] > L.cand_cap) { status = LOOP_NEED_HOST; break; }
            } else {
              if (!ok) { stop = 6; break; }
              if (S.cm[0] > 1 || S.cm[j] > 1) { stop = 2; break; }       // position tie-breaks run alone
              if (S.w[0] > R_SMALL) { stop = 3; break; }                  // a big merge keeps the whole grid
              if (wj < thresh || wj < L.min_weight) { stop = 6; break; }  // the next decision handles the status
              if (it + j >= L.log_cap || cj >= L.max_tokens) { stop = 6; break; }
              bool shares = false;
              for (uint32_t i2 = 0; i2 < j; i2++) shares = shares || aj == S.a[i2] || aj == S.b[i2] || bj == S.a[i2] || bj == S.b[i2];
              if (shares) { stop = 4; break; }
              if (F.v && (aj >= F.c_first || bj >= F.c_first)) { stop = 5; break; }  // its list is still being written
            }
            S.a[j] = aj;
            S.b[j] = bj;
            S.w[j] = wj;
            S.slot[j] = S.cs[j];
            S.lstart[j] = S.cls[j];
            S.llen[j] = S.cll[j];
            S.lenc[j] = A.len16[aj] + A.len16[bj];
            keys_sum += new_keys;
            w_sum += wj;
            k = j + 1;
          }
          if (status == LOOP_RUNNING && k == R.kmax) stop = 0;
        }
        S.iter0[0] = 0;
        for (uint32_t j = 0; j < k; j++) S.iter0[j + 1] = S.iter0[j] + ((S.llen[j] + 31u) >> 5);
        S.k = k;
        S.status = status;
        S.mult0 = S.cm[0];
        if (lead && status == LOOP_RUNNING) {
          rs->rounds++;
          rs->tried += k;
          rs->stop_reason[stop]++;
          if (k == 1) rs->rounds_single++;
        }
      }
      __syncthreads();
    }
    uint32_t status = S.status;
    uint32_t k = S.k;
    prof_big = (status == LOOP_RUNNING && S.w[0] > R_SMALL) ? 1u : 0u;
    if (prof && status == LOOP_RUNNING) st->fine_ns[prof_big ? 11 : 5] += 1;
    RPROF(0)
    // ---- tie on (weight, a.index+b.index): the pair whose last counted occurrence comes first wins (core.ts:294-305) ----
    if (status == LOOP_RUNNING && S.mult0 > 1) {
      if (F.v) {  // the tie-break reads occurrence lists: those of the last round must be complete
        round_fill_all(R, S, F, gt, gn);
        __syncthreads();
        if (tid < RB) S.fill_n[tid] = 0;
        F.v = 0;  // (the rows of that round still wait for their clearing: F.k stays)
        grid_barrier(L.barrier, ++epoch * nblk, L.bar_ns);
      }
      phase_collect(t, A.len16, L.max_length, 1, L.hot, hot_pre, S.cp[0], L.cands, L.cand_cap, st, bid, nblk);
      grid_barrier(L.barrier, ++epoch * nblk, L.bar_ns);
      phase_tie(A.slots, A.n, t, A.pool, L.cands, ld_cg(&st->n_cand), st, S.s_max, bid, nblk);
      grid_barrier(L.barrier, ++epoch * nblk, L.bar_ns);
      const unsigned long long tp = ld_cg(&st->tie_pos);
      __syncthreads();
      if (tp == ~0ull) {
        status = LOOP_ERROR;
      } else if (tid == 0) {
        const uint32_t s = (uint32_t)(tp & 0xFFFFFFFFu);
        const uint32_t key = t.keys[s];
        S.slot[0] = s;
        S.a[0] = key >> 16;
        S.b[0] = key & 0xFFFFu;
        S.lstart[0] = t.occ_start[s];
        S.llen[0] = t.occ_len[s];
        S.lenc[0] = A.len16[key >> 16] + A.len16[key & 0xFFFFu];
        S.iter0[1] = (S.llen[0] + 31u) >> 5;
      }
      __syncthreads();
    }
    RPROF(7)
    if (status != LOOP_RUNNING) {
      // the kernel leaves every list complete and every row zero
      round_fill_all(R, S, F, gt, gn);
      if (F.k) round_clear_rows(R, F.par, F.k, F.c_first + F.k, gt, gn);
      if (lead) {
        st->status = status;
        st->iters_done = it;
        st->n_tokens = n_tokens0 + it;
        Best wb{S.cp[0], S.cp[0] ? S.cs[0] : NOSLOT, S.cm[0]};
        publish_best(t, st, wb);
        st->n_cand = 0;
        st->tie_pos = ~0ull;
      }
      return;
    }
    if (F.v && (S.a[0] >= F.c_first || S.b[0] >= F.c_first)) {
      // the winner contains a token the previous round created: its list is what fill still has to write
      round_fill_all(R, S, F, gt, gn);
      __syncthreads();
      if (tid < RB) S.fill_n[tid] = 0;
      F.v = 0;
      grid_barrier(L.barrier, ++epoch * nblk, L.bar_ns);
      if (tid == 0) {  // (the list length was read before the fill; it does not change, the start neither)
        S.lstart[0] = t.occ_start[S.slot[0]];
        S.llen[0] = t.occ_len[S.slot[0]];
      }
      __syncthreads();
    }
    // ================= P1: the site passes of the batch, next to the list filling of the previous round =================
    if (lead) {
      for (uint32_t j = 0; j < k; j++) A.len16[c_first + j] = S.lenc[j];  // chars = a.chars + b.chars (:318)
      for (uint32_t j = 0; j < RB; j++) {
        rs->n_sites[par ^ 1u][j] = 0;
        rs->ub[par ^ 1u][j][0] = rs->ub[par ^ 1u][j][1] = 0;
      }
      if (S.mult0 > 1) st->tie_breaks++;
    }
    {
      const uint32_t iters = S.iter0[k];
      uint32_t fill_total = 0;
      for (uint32_t j = 0; j < F.v; j++) fill_total += S.fill_n[j];
      const bool split = fill_total <= 8192u * 4u && S.w[0] <= R_SMALL && blockDim.x == 512u;
      const uint32_t ws = split ? L.p1_sites : 16u, wh = 16u - ws;
      if (warp < ws) {
        for (uint32_t wi = bid * ws + warp; wi < iters; wi += nblk * ws) {
          uint32_t j = 0;
          while (j + 1 < k && wi >= S.iter0[j + 1]) j++;
          round_sites_iter(R, S, par, j, c_first, (wi - S.iter0[j]) * 32u + lane, round_sites_buf(R, par, j), j == 0 ? A.sites_cap : R_SMALL);
        }
        if (!split && F.v) round_fill_all(R, S, F, gt, gn);
      } else if (F.v) {
        round_fill_all(R, S, F, (bid * wh + warp - ws) * 32u + lane, nblk * wh * 32u);
      }
    }
    RPROF(1)
    grid_barrier(L.barrier, ++epoch * nblk, L.bar_ns);
    RPROF(2)
    // ================= P2 =================
    // valid prefix: no pair born by an earlier merge of the batch may reach the count of a later one
    uint32_t v = k;
    {
      const uint32_t ubv = (lane < 2 * k) ? ld_cg(&rs->ub[par][lane >> 1][lane & 1u]) : 0u;
      uint32_t run = 0;
      for (uint32_t j = 0; j + 1 < k; j++) {
        run = max(run, max(__shfl_sync(0xFFFFFFFFu, ubv, 2 * j), __shfl_sync(0xFFFFFFFFu, ubv, 2 * j + 1)));
        if (run >= S.w[j + 1]) {
          v = j + 1;
          break;
        }
      }
    }
    uint32_t ns[RB];
#pragma unroll
    for (int j = 0; j < RB; j++) ns[j] = ((uint32_t)j < v) ? ld_cg(&rs->n_sites[par][j]) : 0u;
    if (lead) {
      unsigned long long live = 0;
      uint32_t bad = 0;
#pragma unroll
      for (int j = 0; j < RB; j++) {
        if ((uint32_t)j >= v) break;
        MergeRec r;
        r.a = (int32_t)S.a[j];
        r.b = (int32_t)S.b[j];
        r.c = (int32_t)(c_first + j);
        r.reserved = 0;
        r.weight = (long long)S.w[j];
        L.log[it + j] = r;
        t.cnt[S.slot[j]] = 0;  // every counted occurrence of the winner is being replaced
        live += ns[j];
        bad |= ns[j] != S.w[j];
      }
      st->live_tokens -= live;
      st->sites_total += live;
      st->n_cand = 0;
      st->tie_pos = ~0ull;
      if (bad) atomicOr(&st->err, ERR_ROUND_MISMATCH);
      st->snap_err = st->err;
      rs->round_merges += v;
      if (v < k) rs->rounds_cut_born++;
    }
    Top2 mine = top2_empty();
    {
      const uint32_t sites_all = [&] { uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < RB; j++) s += ns[j];
        return s; }();
      const bool split = sites_all <= 4u * 16384u && S.w[0] <= R_SMALL && blockDim.x == 512u;
      const uint32_t wn = L.p2_new, wr = L.p2_rw, wm = 16u - wn - wr;
      // ---- job 1: the dense rows -> born pairs and decrements of the valid merges; the previous round's rows are zeroed ----
      if (!split || warp < wn) {
        const uint32_t vt = split ? (bid * wn + warp) * 32u + lane : gt, nvt = split ? nblk * wn * 32u : gn;
        const uint32_t c_hi = c_first + k;
        const uint32_t T = (c_hi + 31u) & ~31u;
        const uint32_t kk = max(k, F.k);
        const uint32_t total = kk * 2u * T;
        const uint32_t opar = par ^ 1u;
        for (uint32_t i0 = vt; i0 < total; i0 += 4u * nvt) {
          uint32_t dec[4], len[4], cnt[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const uint32_t i = i0 + (uint32_t)u * nvt;
            dec[u] = len[u] = cnt[u] = 0;
            if (i < total) {
              const uint32_t qd = i / T, tok = i - qd * T, jj = qd >> 1, side = qd & 1u;
              if (jj < v) {
                dec[u] = ld_cg(round_row(R, par, jj, side ? RW_DEC_R : RW_DEC_L) + tok);
                len[u] = ld_cg(round_row(R, par, jj, side ? RW_NR_LEN : RW_NL_LEN) + tok);
                cnt[u] = ld_cg(round_row(R, par, jj, side ? RW_NR_CNT : RW_NL_CNT) + tok);
              }
              if (jj < F.k) {  // rows of the round before this one (the other parity): nobody reads them any more
                round_row(R, opar, jj, side ? RW_DEC_R : RW_DEC_L)[tok] = 0;
                round_row(R, opar, jj, side ? RW_NR_LEN : RW_NL_LEN)[tok] = 0;
                round_row(R, opar, jj, side ? RW_NR_CNT : RW_NL_CNT)[tok] = 0;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const uint32_t i = i0 + (uint32_t)u * nvt;
            if (!__any_sync(0xFFFFFFFFu, (dec[u] | len[u]) != 0)) continue;
            const uint32_t qd = i / T, tok = i - qd * T, jj = min(qd >> 1, (uint32_t)RB - 1u), side = qd & 1u;
            const uint32_t cj = c_first + jj;
            // ---- a pair born by merge jj: (tok, c) or (c, tok) ----
            {
              const bool act = len[u] != 0;
              uint32_t s = NOSLOT, cntv = cnt[u];
              bool ins = false;
              if (act) {
                // adjacencies a LATER valid merge of the batch took away again (its virtual neighbour c_jj)
                for (uint32_t j2 = jj + 1; j2 < v; j2++) {
                  if (side) {
                    if (tok == S.a[j2]) cntv -= ld_cg(round_row(R, par, j2, RW_DEC_L) + cj);
                  } else {
                    if (tok == S.b[j2]) cntv -= ld_cg(round_row(R, par, j2, RW_DEC_R) + cj);
                  }
                }
                s = tbl_find_or_insert_ex(t, side ? pair_key(cj, tok) : pair_key(tok, cj), &ins);
                if (s == NOSLOT) atomicOr(&st->err, ERR_TABLE_FULL);
              }
              const uint32_t im = __ballot_sync(0xFFFFFFFFu, ins);
              if (im && lane == (uint32_t)(__ffs(im) - 1)) atomicAdd(&st->n_keys, (uint32_t)__popc(im));
              uint32_t mylen = (act && s != NOSLOT) ? len[u] : 0u;
              uint32_t inc = mylen;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if ((int)lane >= o) inc += x;
              }
              const uint32_t wtotal = __shfl_sync(0xFFFFFFFFu, inc, 31);
              uint32_t wbase = 0;
              if (lane == 31 && wtotal) wbase = atomicAdd(&st->pool_cursor, wtotal);
              wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
              if (act && s != NOSLOT) {
                uint32_t start = wbase + (inc - mylen);
                if (start > L.pool_cap || mylen > L.pool_cap - start) {
                  atomicOr(&st->err, ERR_POOL_FULL);
                  start = 0;
                  mylen = 0;
                }
                t.occ_start[s] = start;
                t.occ_len[s] = mylen;
                t.occ_fill[s] = 0;
                round_row(R, par, jj, side ? RW_NR_SLOT : RW_NL_SLOT)[tok] = s;
                t.cnt[s] = cntv;  // the key is new: nobody else touches its count in this phase
                const uint32_t pa = side ? cj : tok, pb = side ? tok : cj;
                const uint32_t la = (pa >= c_first) ? S.lenc[pa - c_first] : A.len16[pa], lb = (pb >= c_first) ? S.lenc[pb - c_first] : A.len16[pb];
                const unsigned long long pr = (cntv && !(L.max_length && la + lb > L.max_length)) ? make_primary(cntv, pa, pb) : 0ull;
                if (pr && (uint32_t)(pr >> 20) >= thresh) {
                  const uint32_t hk = atomicAdd(&st->hot_n, 1u);
                  if (hk < L.hot_cap) L.hot[hk] = s;
                  else atomicOr(&st->err, ERR_HOT_OVERFLOW);
                  top2_add(mine, pr, s, 1, pair_key(pa, pb));
                }
              }
            }
            // ---- decrements of an OLD pair: (tok, a_jj) on the left side, (b_jj, tok) on the right side ----
            if (dec[u] != 0 && tok < c_first) {
              uint32_t totald = dec[u];
              bool handle = true;
              const uint32_t ajj = S.a[jj], bjj = S.b[jj];
              if (side) {  // (b_jj, tok): also decremented from the left side of the valid merge whose a is tok
                for (uint32_t j2 = 0; j2 < v; j2++)
                  if (tok == S.a[j2]) totald += ld_cg(round_row(R, par, j2, RW_DEC_L) + bjj);
              } else {     // (tok, a_jj): when tok is the b of a valid merge and that merge's right side holds the pair too, it handles both
                for (uint32_t j2 = 0; j2 < v; j2++)
                  if (tok == S.b[j2] && ld_cg(round_row(R, par, j2, RW_DEC_R) + ajj) != 0) handle = false;
              }
              if (handle) {
                const uint32_t pa = side ? bjj : tok, pb = side ? tok : ajj;
                const uint32_t s = tbl_find(t, pair_key(pa, pb));
                if (s == NOSLOT) {
                  atomicOr(&st->err, ERR_MISSING_KEY);
                } else {
                  const uint32_t old = t.cnt[s];
                  if (old < totald) atomicOr(&st->err, ERR_ROUND_MISMATCH);
                  const uint32_t nv = old - totald;
                  t.cnt[s] = nv;  // this thread is the only one that touches the pair in this phase
                  if (nv && !(L.max_length && A.len16[pa] + A.len16[pb] > L.max_length)) top2_add(mine, make_primary(nv, pa, pb), s, 1, pair_key(pa, pb));
                }
              }
            }
          }
        }
      }
      // ---- job 2: rewrite the corpus at the sites of the valid merges ----
      if (!split || (warp >= wn && warp < wn + wr)) {
        const uint32_t vt = split ? (bid * wr + warp - wn) * 32u + lane : gt, nvt = split ? nblk * wr * 32u : gn;
#pragma unroll
        for (int j = 0; j < RB; j++) {
          if ((uint32_t)j >= v) break;
          const SiteRec* sites = round_sites_buf(R, par, j);
          uint32_t* slots = A.slots;
          const uint32_t cj = c_first + j;
          for (uint32_t i = vt; i < ns[j]; i += nvt) {
            const uint32_t p = ld_cg(&sites[i].p);
            const uint32_t q = next_pos(slots, A.n, p);
            const uint32_t e = next_pos(slots, A.n, q) - 1;
            const uint32_t span = e - p + 1;
            const uint32_t w = ld_slot(slots + p);
            if (span > VAL_MASK) atomicOr(&st->err, ERR_SPAN_OVERFLOW);
            slots[p] = (w & DOCSTART) | cj;
            if (span == 2) {
              slots[p + 1] = mk_back(1);
            } else {
              if (q != p + 1 && q != e) slots[q] = mk_hole();
              slots[p + 1] = mk_span(span);
              slots[e] = mk_back(span - 1);
            }
          }
        }
      }
      // ---- job 3: arg-max over the pairs that were already hot and that no valid merge touches ----
      if (!split || warp >= wn + wr) {
        const uint32_t vt = split ? (bid * wm + warp - wn - wr) * 32u + lane : gt, nvt = split ? nblk * wm * 32u : gn;
        for (uint32_t i = vt; i < hot_pre; i += nvt) {
          const uint32_t hs = L.hot[i];
          const uint32_t key = t.keys[hs];
          if (key == EMPTY_KEY) continue;
          const uint32_t x = key >> 16, y = key & 0xFFFFu;
          bool skip = false;
          for (uint32_t j2 = 0; j2 < v; j2++) {
            if (hs == S.slot[j2]) skip = true;  // a winner: its count is being zeroed
            if (y == S.a[j2] && ld_cg(round_row(R, par, j2, RW_DEC_L) + x) != 0) skip = true;  // job 1 hands in its new count
            if (x == S.b[j2] && ld_cg(round_row(R, par, j2, RW_DEC_R) + y) != 0) skip = true;
          }
          if (skip) continue;
          const uint32_t cv = t.cnt[hs];
          if (!cv) continue;
          if (L.max_length && A.len16[x] + A.len16[y] > L.max_length) continue;
          top2_add(mine, make_primary(cv, x, y), hs, 1, key);
        }
      }
    }
    {
      const Top2 tv = top2_block_reduce(mine, S.t2);
      if (tid == 0) {
        R.gp[2 * bid] = make_uint4((uint32_t)tv.p0, (uint32_t)(tv.p0 >> 32), tv.s0, tv.m0);
        R.gp[2 * bid + 1] = make_uint4((uint32_t)tv.p1, (uint32_t)(tv.p1 >> 32), tv.s1, tv.m1);
        R.gk[2 * bid] = tv.k0;
        R.gk[2 * bid + 1] = tv.k1;
      }
      if (tid < RB) S.fill_n[tid] = (tid < v) ? ld_cg(&rs->n_sites[par][tid]) : 0u;
    }
    F.v = v;
    F.par = par;
    F.k = k;
    F.c_first = c_first;
    it += v;
    RPROF(3)
    grid_barrier(L.barrier, ++epoch * nblk, L.bar_ns);
    RPROF(4)
  }
#undef RPROF
}

}  // namespace bpe
