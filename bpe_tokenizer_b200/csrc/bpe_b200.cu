// bpe_b200.cu -- engine state, host orchestration and the C ABI declared in include/bpe_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (build.py).
#include "../../include/bpe_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "encode_kernels.cuh"
#include "encode_lanes.cuh"
#include "encode_dp.cuh"
#include "train_kernels.cuh"
#include "mg_kernels.cuh"
#include "round_kernels.cuh"
#include "text_kernels.cuh"

using namespace bpe;

namespace {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) {
    o.p = nullptr;
    o.cap = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      cap = o.cap;
      o.p = nullptr;
      o.cap = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // grow to at least n elements; keep_n elements are preserved (stream-ordered copy)
  cudaError_t reserve(size_t n, size_t keep_n = 0, cudaStream_t s = 0, double growth = 1.0) {
    if (n <= cap) return cudaSuccess;
    size_t want = std::max(n, (size_t)((double)cap * growth));
    T* q = nullptr;
    cudaError_t err = cudaMalloc(&q, want * sizeof(T));
    if (err != cudaSuccess) {
      (void)cudaGetLastError();  // the runtime keeps the failure as its "last error": the next launch check must not see it
      want = n;
      err = cudaMalloc(&q, want * sizeof(T));
      if (err != cudaSuccess) {
        (void)cudaGetLastError();
        return err;
      }
    }
    if (p && keep_n) {
      err = cudaMemcpyAsync(q, p, keep_n * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (err != cudaSuccess) {
        cudaFree(q);
        return err;
      }
      cudaStreamSynchronize(s);
    }
    if (p) cudaFree(p);
    p = q;
    cap = want;
    return cudaSuccess;
  }
};

// grow-only pinned host buffer (staging of the small per-document arrays of the encode pipeline)
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t err = cudaHostAlloc((void**)&p, n * sizeof(T), cudaHostAllocDefault);
    if (err == cudaSuccess) cap = n;
    else (void)cudaGetLastError();
    return err;
  }
};

uint32_t pow2_at_least(uint64_t x) {
  uint32_t p = 1;
  while (p < x && p < 0x80000000u) p <<= 1;
  return p;
}
int ilog2(uint32_t p) {
  int l = 0;
  while ((1u << l) < p) l++;
  return l;
}

}  // namespace

struct EncodeScratch {
  DevBuf<int32_t> out_tmp;
  DevBuf<uint32_t> out_len;
  DevBuf<uint64_t> out_off;
  DevBuf<uint32_t> g_tok, g_nxt, g_prv, g_rk, g_sel;
  DevBuf<uint32_t> range_first, tile_sums;
  DevBuf<uint64_t> tile_off;
  DevBuf<uint32_t> flags;  // [0] long documents left for the per-document kernel, [1] error, [2] next range to claim
  DevBuf<uint16_t> dp_last;              // forward path (encode_dp.cuh): last[] of every document, per warp [position][lane]
  DevBuf<unsigned long long> dp_cursor;  // ... and where the next warp's rows start
};

struct bpe_engine {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  std::string err;
  bool profiling = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  // corpus_in_code (core.ts:106): slots + document boundaries
  DevBuf<uint32_t> slots;
  uint64_t n_slots = 0;
  std::vector<int64_t> doc_off{0};
  uint64_t live_tokens = 0;

  // vocabulary
  std::vector<int32_t> h_len16;
  DevBuf<uint32_t> d_len16;
  int32_t n_tokens = 0;

  // merge list (core.ts:88-91) + device lookup table for encode
  std::vector<int32_t> h_merges;
  bool mt_dirty = true;
  DevBuf<unsigned long long> d_mt;
  DevBuf<uint32_t> d_minlr;  // per token: lowest rank as left operand << 16 | lowest rank as right operand
  uint32_t mt_cap = 0;
  // lane path (encode_lanes.cuh): pair table with spine bounds, dense table of the low ids, token of each rank
  bool lt_dirty = true;
  DevBuf<uint4> d_lt;
  DevBuf<uint2> d_lt_dense;
  DevBuf<uint16_t> d_rule_c;
  uint32_t lt_cap = 0;
  int32_t lt_c_affine = -1;
  // forward path (encode_dp.cuh): Aho-Corasick automaton over the tokens' characters + merge trees
  bool dp_dirty = true, dp_ok = false;
  DevBuf<unsigned long long> d_dfa;
  DevBuf<uint32_t> d_split;
  DevBuf<uint16_t> d_tok_len, d_shorter;
  bool dp_attr_set = false;
  uint32_t dp_alpha = 0;
  int enc_dp = 0;        // BPE_ENC_DP=1: forward path (exact, measured slower than the lane path: 8.5 vs 26.9 GB/s -- thread-per-document
                         // keeps 5 of 32 lanes busy; kept as an independent second implementation)
  int enc_lmax = 32;     // rows per lane of the lane path (32 or 48); BPE_ENC_LMAX overrides
  int enc_lmax_forced = 0;
  int enc_force_old = 0; // debug: BPE_ENC_OLD=1 routes every document through the per-document kernel

  // pair index
  bool index_valid = false;
  uint32_t tbl_cap = 0;
  DevBuf<uint32_t> t_ent;  // 5 fields x capacity words (field-major when TBL_STRIDE == 1, common.cuh)
  DevBuf<uint32_t> u_ent;  // second buffer, target of the next rehash (kept at high water)
  DevBuf<uint32_t> pool;
  DevBuf<DevState> d_st;
  DevState* h_st = nullptr;  // pinned
  DevBuf<Best> partials;
  DevBuf<uint32_t> partial_keys;
  DevBuf<SiteRec> sites, sites2;  // k_merge_loop alternates between the two (deferred list filling)
  DevBuf<uint32_t> newslots;
  DevBuf<uint32_t> nd;  // dense accumulator of the pairs born by the current merge (train_kernels.cuh, ND_*)
  DevBuf<uint32_t> hot;
  bool hot_valid = false;
  uint32_t hot_max_length = 0;
  uint32_t hot_thresh = 0;
  uint32_t hot_limit = 0;
  DevBuf<uint32_t> cands;
  DevBuf<MergeRec> dev_log;
  DevBuf<unsigned long long> barrier;  // own 128-byte line
  int loop_blocks = 0;   // co-resident grid of k_merge_loop
  // several exact merges per barrier round (round_kernels.cuh): dense delta rows, site buffers of merges 1.., per-block top-2 partials
  DevBuf<unsigned long long> r_cells;
  DevBuf<uint32_t> r_slotrows, r_lists;
  DevBuf<SiteRec> r_bsites;
  DevBuf<uint4> r_gp;
  DevBuf<uint32_t> r_gk;
  DevBuf<RoundState> r_state;
  DevBuf<unsigned long long> r_gcells;  // sharded rounds: the deltas summed over the ranks
  DevBuf<uint32_t> r_glists;
  int round_blocks = 0;  // co-resident grid of k_merge_rounds
  int host_loop = 0;     // debug: drive mergeUntil from the host, one launch per phase
  int scan_mode = 0;     // debug: walk all slots instead of occurrence lists

  // sharded training (mg_kernels.cuh): this rank's mailbox + the peers' mailboxes mapped through cudaIpc
  int mg_rank = 0, mg_world = 1;
  void* mg_mailbox = nullptr;            // cudaMalloc'ed, exported with cudaIpcGetMemHandle
  size_t mg_mailbox_bytes = 0;
  void* mg_peer[MG_MAX_WORLD] = {};      // [rank] == mg_mailbox
  bool mg_connected = false;
  bool mg_counts_global = false;         // the table's counts are the sum over all ranks (import done for this index)
  uint32_t mg_inbox_stride = 0, mg_tie_cap = 0;
  DevBuf<int32_t> mg_dlt;
  DevBuf<uint32_t> mg_mark, mg_touched, mg_tie_sorted, mg_export_n, mg_newpair;
  int mg_loop_blocks = 0;
  unsigned long long mg_epoch = 0, mg_tie_epoch = 0;

  // text front end (text_kernels.cuh): code point -> single-character token index, -1 = unknown
  DevBuf<int32_t> d_cpmap;
  DevBuf<uint32_t> d_firstpos;
  DevBuf<uint8_t> x_text;
  DevBuf<uint32_t> x_tilecnt;
  DevBuf<uint64_t> x_tileoff;
  DevBuf<unsigned long long> x_counts;
  DevBuf<uint32_t> x_doccnt;
  DevBuf<uint64_t> x_docoff;

  // staging of the host-buffer encode / restore calls (grow-only, reused across calls)
  DevBuf<int32_t> x_ids, x_out, x_tvi;
  DevBuf<int64_t> x_off, x_ooff, x_ooff2, x_bad;
  DevBuf<unsigned long long> x_flag;
  std::vector<int64_t> h_rel;
  EncodeScratch x_scratch;
  EncodeScratch dev_scratch;             // scratch of bpe_encode_batch_dev (device buffers in, device buffers out)
  uint32_t lanes_attr_set = 0;           // bit per k_encode_lanes instantiation whose dynamic shared memory limit was raised on
                                         // THIS engine's device (the attribute is per device, not per process)
  // the host-buffer encode calls run as a three-stream pipeline over chunks of whole documents: copy in | encode | copy out
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_done = nullptr, ev_res[2] = {nullptr, nullptr};
  unsigned long long* h_pipe = nullptr;  // pinned: 3 slots x 4 words of per-chunk flags, [12..13] flags of the lane kernel
  PinBuf<int64_t> h_in_off, h_out_off, h_bad;  // pinned staging of document offsets in, output offsets and first offenders out
  int64_t enc_chunk = 128ll << 20;       // input units (ids or bytes) per full-size chunk; BPE_ENC_CHUNK overrides

  // scratch
  DevBuf<int32_t> stage_ids;
  DevBuf<int64_t> stage_off;

  bpe_stats stats{};

  PairTable table() const {
    PairTable t;
    const size_t fs = (TBL_STRIDE == 1) ? (size_t)tbl_cap : 1;  // distance between fields
    t.keys.p = t_ent.p;
    t.cnt.p = t_ent.p + fs;
    t.occ_start.p = t_ent.p + 2 * fs;
    t.occ_len.p = t_ent.p + 3 * fs;
    t.occ_fill.p = t_ent.p + 4 * fs;
    t.mask = tbl_cap - 1;
    t.shift = 32 - ilog2(tbl_cap);
    return t;
  }
  int grid(int per_sm = 4) const { return sm_count * per_sm; }
};

namespace {

int fail(bpe_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  return code;
}

#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t _err = (call);                                                                                \
    if (_err != cudaSuccess)                                                                                  \
      return fail(e, _err == cudaErrorMemoryAllocation ? BPE_E_NOMEM : BPE_E_CUDA, "%s:%d %s: %s", __FILE__, \
                  __LINE__, #call, cudaGetErrorString(_err));                                                 \
  } while (0)

#define CKL()                                \
  do {                                       \
    e->stats.kernel_launches++;              \
    CK(cudaGetLastError());                  \
  } while (0)

#define TRY(call)             \
  do {                        \
    int _rc = (call);         \
    if (_rc != BPE_OK) return _rc; \
  } while (0)

int fetch_state(bpe_engine* e) {
  CK(cudaMemcpyAsync(e->h_st, e->d_st.p, sizeof(DevState), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return BPE_OK;
}

int check_dev_err(bpe_engine* e) {
  uint32_t f = e->h_st->err & ~ERR_HOT_OVERFLOW;  // hot-list overflow is recoverable (rebuild), handled by the caller
  if (!f) return BPE_OK;
  if (f & ERR_POOL_FULL) return fail(e, BPE_E_INTERNAL, "occurrence pool exhausted (flags 0x%x)", f);
  if (f & ERR_MISSING_KEY) return fail(e, BPE_E_INTERNAL, "pair missing from the table (flags 0x%x)", f);
  if (f & ERR_SPAN_OVERFLOW) return fail(e, BPE_E_DOMAIN, "token span exceeds 2^29 positions");
  return fail(e, BPE_E_INTERNAL, "device error flags 0x%x", f);
}

int sync_len16(bpe_engine* e) {
  size_t n = e->h_len16.size();
  CK(e->d_len16.reserve(std::max<size_t>(n + 1024, 4096), 0, e->stream, 2.0));
  if (n) {
    std::vector<uint32_t> tmp(e->h_len16.begin(), e->h_len16.end());
    CK(cudaMemcpyAsync(e->d_len16.p, tmp.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return BPE_OK;
}

// ---- pair table allocation / growth -------------------------------------------------------------
__global__ void k_init_table(uint4* __restrict__ ent, uint64_t cap) {  // key = EMPTY_KEY, every other field 0
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
    ent[2 * i] = make_uint4(EMPTY_KEY, 0u, 0u, 0u);
    ent[2 * i + 1] = make_uint4(0u, 0u, 0u, 0u);
  }
}

int alloc_table(bpe_engine* e, uint32_t cap) {
  // buffers only ever grow (cudaFree / cudaMalloc of GBs costs up to seconds); the logical capacity is `cap`
  CK(e->t_ent.reserve((size_t)cap * (TBL_STRIDE == 1 ? 5 : TBL_STRIDE)));
  e->tbl_cap = cap;
  if (TBL_STRIDE == 1) {
    CK(cudaMemsetAsync(e->t_ent.p, 0xFF, (size_t)cap * 4, e->stream));
    CK(cudaMemsetAsync(e->t_ent.p + cap, 0, (size_t)cap * 16, e->stream));
  } else {
    k_init_table<<<e->grid(8), 256, 0, e->stream>>>(reinterpret_cast<uint4*>(e->t_ent.p), (uint64_t)cap);
    CKL();
  }
  if (e->mg_world > 1) {  // per-slot side arrays of the sharded loop follow the table (all zero between merges)
    CK(e->mg_dlt.reserve(cap));
    CK(cudaMemsetAsync(e->mg_dlt.p, 0, (size_t)cap * 4, e->stream));
  }
  return BPE_OK;
}

__global__ void k_rehash(PairTable src, PairTable dst, DevState* st) {
  uint32_t cap = src.mask + 1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
    uint32_t key = src.keys[i];
    if (key == EMPTY_KEY) continue;
    uint32_t h = (key * 0x9E3779B1u) >> dst.shift;
    for (;;) {
      uint32_t old = atomicCAS(dst.keys + h, EMPTY_KEY, key);
      if (old == EMPTY_KEY) break;
      h = (h + 1) & dst.mask;
    }
    dst.cnt[h] = src.cnt[i];
    dst.occ_start[h] = src.occ_start[i];
    dst.occ_len[h] = src.occ_len[i];
    dst.occ_fill[h] = src.occ_fill[i];
  }
}

int grow_table(bpe_engine* e, uint32_t new_cap) {
  PairTable old = e->table();
  std::swap(e->t_ent, e->u_ent);
  TRY(alloc_table(e, new_cap));  // the (former) alternate set becomes the live table
  k_rehash<<<e->grid(), 256, 0, e->stream>>>(old, e->table(), e->d_st.p);
  CKL();
  CK(cudaStreamSynchronize(e->stream));
  e->hot_valid = false;  // slot numbers changed
  return BPE_OK;
}

// ---- K1: build histogram + occurrence lists from the corpus -------------------------------------
int build_index(bpe_engine* e) {
  e->index_valid = false;
  e->hot_valid = false;
  if (!e->h_st) CK(cudaHostAlloc((void**)&e->h_st, sizeof(DevState), cudaHostAllocDefault));
  CK(e->d_st.reserve(1));
  CK(e->partials.reserve((size_t)e->grid(8)));
  if (!e->nd.p) CK(e->nd.reserve((size_t)ND_ROWS * ND_STRIDE));
  CK(cudaMemsetAsync(e->nd.p, 0, (size_t)ND_ROWS * ND_STRIDE * 4, e->stream));  // all rows zero between merges
  TRY(sync_len16(e));
  uint64_t n = e->n_slots;
  if (n >= 0xFFFFFFF0ull) return fail(e, BPE_E_DOMAIN, "corpus of %llu positions exceeds the 2^32 engine limit", (unsigned long long)n);
  uint64_t pool_need = 3 * n + 1024;
  if (pool_need > 0xFFFFFFF0ull) pool_need = 0xFFFFFFF0ull;
  CK(e->pool.reserve((size_t)pool_need));
  uint64_t t2 = (uint64_t)e->n_tokens * (uint64_t)e->n_tokens;
  uint32_t cap = pow2_at_least(std::min<uint64_t>(std::max<uint64_t>(2 * std::min(n, t2), 1u << 16), 1u << 24));
  {
    if (!e->ev0) {
      CK(cudaEventCreate(&e->ev0));
      CK(cudaEventCreate(&e->ev1));
    }
    CK(cudaEventRecord(e->ev0, e->stream));
  }
  for (;;) {
    TRY(alloc_table(e, cap));
    DevState init{};
    init.live_tokens = e->live_tokens;
    init.tie_pos = ~0ull;
    init.mg_epoch = e->mg_epoch;  // the peers' flags keep counting across index rebuilds
    init.mg_tie_epoch = e->mg_tie_epoch;
    CK(cudaMemcpyAsync(e->d_st.p, &init, sizeof init, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (n) {
      k_hist<<<e->grid(4), K1_THREADS, 0, e->stream>>>(e->slots.p, (uint32_t)n, e->table(), e->d_st.p);
      CKL();
    }
    TRY(fetch_state(e));
    if (e->h_st->err & ERR_TABLE_FULL) {
      if (cap >= 0x80000000u) return fail(e, BPE_E_NOMEM, "pair table cannot grow further");
      cap <<= 2;
      continue;
    }
    TRY(check_dev_err(e));
    if ((uint64_t)e->h_st->n_keys * 2 > cap && cap < 0x80000000u) {  // keep the load factor <= 0.5
      cap <<= 1;
      continue;
    }
    break;
  }
  if (n) {
    k_alloc_lists<<<e->grid(4), 256, 0, e->stream>>>(e->table(), e->d_st.p, (uint32_t)e->pool.cap);
    CKL();
    int k1b_per_sm = 8;
    if (const char* v = getenv("BPE_K1B_PER_SM")) k1b_per_sm = std::max(1, std::min(atoi(v), 8));  // tuning knob
    k_scatter<<<e->grid(k1b_per_sm), K1_THREADS, 0, e->stream>>>(e->slots.p, (uint32_t)n, e->table(), e->pool.p, e->d_st.p);
    CKL();
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  TRY(fetch_state(e));
  TRY(check_dev_err(e));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->stats.ms_index_build += ms;
  e->stats.index_builds++;
  e->index_valid = true;
  e->mg_counts_global = false;
  return BPE_OK;
}

int ensure_index(bpe_engine* e) {
  if (e->index_valid) return BPE_OK;
  return build_index(e);
}

// ---- hot list -----------------------------------------------------------------------------------
constexpr uint32_t HOT_TARGET = 1u << 15;

// returns found=0 when no pair with a count exists at all
int rebuild_hot(bpe_engine* e, uint32_t max_length, bool* any) {
  PairTable t = e->table();
  CK(cudaMemsetAsync(&e->d_st.p->bins[0], 0, sizeof(uint32_t) * 36, e->stream));
  k_count_bins<<<e->grid(4), 256, 0, e->stream>>>(t, e->d_len16.p, max_length, e->d_st.p);
  CKL();
  TRY(fetch_state(e));
  uint64_t acc = 0;
  int pick = 0;
  for (int k = 32; k >= 1; k--) {
    uint64_t nk = e->h_st->bins[k];
    if (acc + nk > HOT_TARGET && acc > 0) break;
    acc += nk;
    if (nk) pick = k;
    if (acc > HOT_TARGET) break;
  }
  *any = pick != 0;
  e->hot_valid = false;
  if (!pick) return BPE_OK;
  // everything from bin `pick` up: counts >= 2^(pick-1); extend down to the lowest bin that adds nothing
  int lo = pick;
  while (lo > 1 && e->h_st->bins[lo - 1] == 0) lo--;
  uint32_t thresh = 1u << (lo - 1);
  CK(e->hot.reserve(std::max<size_t>((size_t)acc * 2 + 4096 + 2 * (size_t)BPE_MAX_TOKENS, 1u << 18)));
  e->h_st->hot_n = 0;
  e->h_st->hot_thresh = thresh;
  uint32_t two[2] = {0, thresh};
  CK(cudaMemcpyAsync(&e->d_st.p->hot_n, two, sizeof two, cudaMemcpyHostToDevice, e->stream));
  k_build_hot<<<e->grid(4), 256, 0, e->stream>>>(t, e->d_len16.p, max_length, e->hot.p, (uint32_t)e->hot.cap, e->d_st.p);
  CKL();
  e->hot_thresh = thresh;
  e->hot_limit = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(4ull * HOT_TARGET, 2 * acc + 4096), 0xFFFFFFF0u);
  e->hot_max_length = max_length;
  e->hot_valid = true;
  e->stats.hot_rebuilds++;
  return BPE_OK;
}

// ---- K2 driver: leaves the winner in h_st (best_primary == 0: none) -------------------------------
int run_argmax(bpe_engine* e, uint32_t max_length, int use_hot) {
  PairTable t = e->table();
  int blocks = use_hot ? e->sm_count : e->grid(4);
  k_argmax<<<blocks, AM_THREADS, 0, e->stream>>>(t, e->d_len16.p, max_length, use_hot, e->hot.p, e->partials.p, e->d_st.p);
  CKL();
  TRY(fetch_state(e));
  TRY(check_dev_err(e));
  if (e->h_st->best_primary && e->h_st->best_mult > 1 && !(use_hot && e->h_st->best_cnt < e->hot_thresh)) {
    // tie on (weight, a.index + b.index): the pair whose last counted occurrence comes first wins (core.ts:294-305)
    uint32_t mult = e->h_st->best_mult;
    CK(e->cands.reserve(std::max<size_t>(mult, 1024), 0, e->stream, 2.0));
    k_collect_cands<<<blocks, 256, 0, e->stream>>>(t, e->d_len16.p, max_length, use_hot, e->hot.p, e->cands.p, (uint32_t)e->cands.cap, e->d_st.p);
    CKL();
    k_tie_break<<<std::min<uint32_t>(mult, (uint32_t)e->grid(4)), 256, 0, e->stream>>>(e->slots.p, (uint32_t)e->n_slots, t, e->pool.p, e->cands.p, e->d_st.p);
    CKL();
    k_tie_finish<<<1, 32, 0, e->stream>>>(t, e->d_st.p);
    CKL();
    TRY(fetch_state(e));
    TRY(check_dev_err(e));
    e->stats.tie_breaks++;
  }
  return BPE_OK;
}

ApplyArgs apply_args(bpe_engine* e) {
  ApplyArgs A;
  A.slots = e->slots.p;
  A.n = (uint32_t)e->n_slots;
  A.t = e->table();
  A.pool = e->pool.p;
  A.st = e->d_st.p;
  A.sites = e->sites.p;
  A.sites_cap = (uint32_t)std::min<size_t>(e->sites.cap, 0xFFFFFFFFu);
  A.newslots = e->newslots.p;
  A.nd = e->nd.p;
  A.newpair = nullptr;
  A.newpair_cap = 0;
  A.new_cap = (uint32_t)std::min<size_t>(e->newslots.cap, 0xFFFFFFFFu);
  A.len16 = e->d_len16.p;
  A.scan_mode = e->scan_mode;
  A.dlt = nullptr;
  A.touched = nullptr;
  A.touched_cap = 0;
  for (int q = 0; q < 8; q++) A.push[q] = nullptr;
  A.push_world = 0;
  A.push_cap = 0;
  return A;
}

// ---- K3 driver ------------------------------------------------------------------------------------
// bound = upper bound on the number of sites (count of the pair, or its list length)
int run_apply(bpe_engine* e, uint32_t a, uint32_t b, uint32_t c, uint32_t bound) {
  // capacity: new keys <= 2*bound, new list cells <= 2*bound
  uint64_t keys_after = (uint64_t)e->h_st->n_keys + std::min<uint64_t>(2ull * bound + 2, 2ull * ((uint64_t)e->n_tokens + 1) + 2);
  if (keys_after * 2 > e->tbl_cap) {
    uint64_t want = keys_after * 5 / 2;
    if (want > 0x80000000ull) want = 0x80000000ull;
    if (keys_after * 2 > pow2_at_least(want)) return fail(e, BPE_E_NOMEM, "pair table cannot grow further");
    TRY(grow_table(e, pow2_at_least(want)));
  }
  uint64_t pool_after = (uint64_t)e->h_st->pool_cursor + 2ull * bound;
  if (pool_after > e->pool.cap) {
    if (pool_after > 0xFFFFFFF0ull) return fail(e, BPE_E_DOMAIN, "occurrence pool exceeds 2^32 cells");
    CK(e->pool.reserve((size_t)pool_after, e->h_st->pool_cursor, e->stream, 1.5));
  }
  CK(e->sites.reserve(std::max<size_t>(bound, 4096), 0, e->stream, 2.0));
  CK(e->newslots.reserve(std::max<size_t>(2ull * bound + 2, 8192), 0, e->stream, 2.0));
  if ((size_t)c + 1 > e->d_len16.cap) CK(e->d_len16.reserve((size_t)c + 1, (size_t)c, e->stream, 2.0));
  PairTable t = e->table();
  uint32_t work = e->scan_mode ? (uint32_t)e->n_slots : bound;
  int blocks = (int)std::min<uint64_t>((uint64_t)e->grid(8), std::max<uint64_t>(1, (work + 255) / 256));
  ApplyArgs A = apply_args(e);
  k_sites<<<blocks, 256, 0, e->stream>>>(A, a, b, c);
  CKL();
  int blocks2 = (int)std::min<uint64_t>((uint64_t)e->grid(8), std::max<uint64_t>(1, (2ull * bound + 255) / 256));
  k_alloc_new<<<blocks2, 256, 0, e->stream>>>(A, c, e->hot_max_length, e->hot_valid ? 1 : 0, e->hot.p, (uint32_t)e->hot.cap, (uint32_t)e->pool.cap);
  CKL();
  int blocks3 = (int)std::min<uint64_t>((uint64_t)e->grid(8), std::max<uint64_t>(1, ((uint64_t)bound + 255) / 256));
  k_apply<<<blocks3, 256, 0, e->stream>>>(A, a, b, c);
  CKL();
  e->h_len16.push_back(e->h_len16[a] + e->h_len16[b]);
  e->h_merges.push_back((int32_t)a);
  e->h_merges.push_back((int32_t)b);
  e->h_merges.push_back((int32_t)c);
  e->mt_dirty = e->lt_dirty = e->dp_dirty = true;
  e->n_tokens++;
  e->stats.merges_applied++;
  return BPE_OK;
}

// ---- merge lookup table for encode ------------------------------------------------------------------
int ensure_merge_table(bpe_engine* e) {
  if (!e->mt_dirty && e->d_mt.p) return BPE_OK;
  size_t m = e->h_merges.size() / 3;
  uint32_t cap = pow2_at_least(std::max<size_t>(4 * m, 1024));
  std::vector<unsigned long long> h(cap, MT_EMPTY);
  int shift = 32 - ilog2(cap);
  for (size_t r = 0; r < m; r++) {
    uint32_t a = (uint32_t)e->h_merges[3 * r], b = (uint32_t)e->h_merges[3 * r + 1], c = (uint32_t)e->h_merges[3 * r + 2];
    uint32_t key = pair_key(a, b);
    uint32_t i = (key * 0x9E3779B1u) >> shift;
    bool dup = false;
    while (h[i] != MT_EMPTY) {
      if ((uint32_t)(h[i] >> 32) == key) {
        dup = true;  // a second rule for the same pair can never fire (the first one removed every occurrence)
        break;
      }
      i = (i + 1) & (cap - 1);
    }
    if (!dup) h[i] = ((unsigned long long)key << 32) | ((unsigned long long)(uint32_t)r << 16) | c;
  }
  std::vector<uint32_t> minlr((size_t)std::max(e->n_tokens, 1), 0xFFFFFFFFu);
  for (size_t r = 0; r < m; r++) {
    uint32_t a = (uint32_t)e->h_merges[3 * r], b = (uint32_t)e->h_merges[3 * r + 1];
    if (a >= minlr.size() || b >= minlr.size()) return fail(e, BPE_E_INVALID, "merge %zu refers to a token outside the table", r);
    uint32_t rr = (uint32_t)std::min<size_t>(r, 0xFFFE);
    if ((minlr[a] >> 16) > rr) minlr[a] = (minlr[a] & 0xFFFFu) | (rr << 16);
    if ((minlr[b] & 0xFFFFu) > rr) minlr[b] = (minlr[b] & 0xFFFF0000u) | rr;
  }
  CK(e->d_minlr.reserve(minlr.size()));
  CK(cudaMemcpyAsync(e->d_minlr.p, minlr.data(), minlr.size() * 4, cudaMemcpyHostToDevice, e->stream));
  CK(e->d_mt.reserve(cap));
  CK(cudaMemcpyAsync(e->d_mt.p, h.data(), (size_t)cap * 8, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->mt_cap = cap;
  e->mt_dirty = false;
  return BPE_OK;
}

// ---- pair table of the lane path: rank + spine bounds per pair (see encode_lanes.cuh) ---------------------------
// rank, token and spine bounds of every pair that is a rule or lies on a spine of one (encode_lanes.cuh) -- pure host work,
// also reachable without a device through bpe_debug_lane_table (CPU parity test against the prototype of the algorithm)
struct LaneEnt {
  uint16_t rk = 0xFFFF, c = 0, rs = 0xFFFF, ls = 0xFFFF;
};

int build_lane_entries(bpe_engine* e, const std::vector<int32_t>& merges, int32_t n_tokens, std::unordered_map<uint32_t, LaneEnt>& M,
                       std::vector<uint16_t>& rule_c) {
  size_t m = merges.size() / 3;
  if (m > EL_MAX_RANK + 1) return fail(e, BPE_E_DOMAIN, "merge list too long");
  const int32_t nt = std::max(n_tokens, 1);
  M.clear();
  M.reserve(m * 3 + 16);
  // definition of every merged token (core.ts:315-325: c is a fresh index, so each c has one definition)
  std::vector<int32_t> def_a((size_t)nt, -1), def_b((size_t)nt, -1);
  std::vector<uint8_t> first(m, 0);
  rule_c.assign(std::max<size_t>(m, 1), 0);
  for (size_t r = 0; r < m; r++) {
    int32_t a = merges[3 * r], b = merges[3 * r + 1], c = merges[3 * r + 2];
    if (a < 0 || b < 0 || c < 0 || a >= nt || b >= nt || c >= nt)
      return fail(e, BPE_E_INVALID, "merge %zu refers to a token outside the table", r);
    rule_c[r] = (uint16_t)c;
    if (def_a[c] < 0 && a < c && b < c) {
      def_a[c] = a;
      def_b[c] = b;
    } else if (def_a[c] >= 0 && (def_a[c] != a || def_b[c] != b)) {
      return fail(e, BPE_E_INVALID, "token %d is produced by two different merges", c);
    }
    LaneEnt& x = M[pair_key((uint32_t)a, (uint32_t)b)];
    if (x.rk == 0xFFFF) {  // a second rule for the same pair can never fire (the first one removed every occurrence)
      x.rk = (uint16_t)r;
      x.c = (uint16_t)c;
      first[r] = 1;
    }
  }
  for (size_t r = 0; r < m; r++) {
    if (!first[r]) continue;
    int32_t a = merges[3 * r], b = merges[3 * r + 1];
    for (int32_t u = a;;) {  // every token on the right spine of a: anything ending in u may grow into a
      LaneEnt& x = M[pair_key((uint32_t)u, (uint32_t)b)];
      if (x.rs > r) x.rs = (uint16_t)r;
      if (def_a[u] < 0) break;
      u = def_b[u];
    }
    for (int32_t w = b;;) {  // every token on the left spine of b
      LaneEnt& x = M[pair_key((uint32_t)a, (uint32_t)w)];
      if (x.ls > r) x.ls = (uint16_t)r;
      if (def_a[w] < 0) break;
      w = def_a[w];
    }
  }
  return BPE_OK;
}

int ensure_lane_tables(bpe_engine* e) {
  if (!e->lt_dirty && e->d_lt.p) return BPE_OK;
  size_t m = e->h_merges.size() / 3;
  std::unordered_map<uint32_t, LaneEnt> M;
  std::vector<uint16_t> rule_c;
  TRY(build_lane_entries(e, e->h_merges, e->n_tokens, M, rule_c));
  uint32_t cap = pow2_at_least(std::max<size_t>(2 * M.size(), 1024));
  std::vector<uint4> h(cap, make_uint4(EMPTY_KEY, EL_NONE, 0xFFFFFFFFu, 0));
  std::vector<uint2> dense((size_t)EL_DENSE * EL_DENSE, make_uint2(EL_NONE, 0xFFFFFFFFu));
  int shift = 32 - ilog2(cap);
  for (auto& kv : M) {
    uint32_t key = kv.first;
    const LaneEnt& x = kv.second;
    uint32_t y = (uint32_t)x.rk | ((uint32_t)x.c << 16), z = (uint32_t)x.rs | ((uint32_t)x.ls << 16);
    uint32_t a = key >> 16, b = key & 0xFFFFu;
    if (a < (uint32_t)EL_DENSE && b < (uint32_t)EL_DENSE) dense[a * EL_DENSE + b] = make_uint2(y, z);
    uint32_t i = (key * 0x9E3779B1u) >> shift;
    while (h[i].x != EMPTY_KEY) i = (i + 1) & (cap - 1);
    h[i] = make_uint4(key, y, z, 0);
  }
  CK(e->d_lt.reserve(cap));
  CK(cudaMemcpyAsync(e->d_lt.p, h.data(), (size_t)cap * sizeof(uint4), cudaMemcpyHostToDevice, e->stream));
  CK(e->d_lt_dense.reserve(dense.size()));
  CK(cudaMemcpyAsync(e->d_lt_dense.p, dense.data(), dense.size() * sizeof(uint2), cudaMemcpyHostToDevice, e->stream));
  CK(e->d_rule_c.reserve(rule_c.size()));
  CK(cudaMemcpyAsync(e->d_rule_c.p, rule_c.data(), rule_c.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->lt_cap = cap;
  e->lt_c_affine = m ? (int32_t)rule_c[0] : 0;
  for (size_t r = 0; r < m; r++)
    if ((int64_t)rule_c[r] != (int64_t)rule_c[0] + (int64_t)r) e->lt_c_affine = -1;
  e->lt_dirty = false;
  return BPE_OK;
}

// ---- tables of the forward encode path (encode_dp.cuh): tokens ending at a position, longest first, and merge trees ----
// Supported when the table is what mergeUntil / fromJSON produce: the single characters are the tokens 0 .. n_alpha-1, merge r
// creates token n_alpha + r out of older tokens, no two tokens spell the same string, and the dense automaton fits.
int ensure_dp_tables(bpe_engine* e) {
  if (!e->dp_dirty) return BPE_OK;
  e->dp_dirty = false;
  e->dp_ok = false;
  const size_t m = e->h_merges.size() / 3;
  const int64_t n_tok = e->n_tokens;
  const int64_t A = n_tok - (int64_t)m;
  if (A <= 0 || A > 4096 || n_tok > 0xFFF0) return BPE_OK;
  std::vector<uint32_t> split((size_t)n_tok), off((size_t)n_tok + 1, 0);
  std::vector<uint16_t> tok_len((size_t)n_tok, 1);
  for (int64_t t = 0; t < A; t++) split[(size_t)t] = (uint32_t)t | ((uint32_t)t << 16);
  for (size_t r = 0; r < m; r++) {
    const int64_t a = e->h_merges[3 * r], b = e->h_merges[3 * r + 1], c = e->h_merges[3 * r + 2];
    if (c != A + (int64_t)r || a < 0 || b < 0 || a >= c || b >= c) return BPE_OK;  // (characters added after merges: lane path)
    const uint32_t L = (uint32_t)tok_len[(size_t)a] + tok_len[(size_t)b];
    if (L > 0xFFF0u) return BPE_OK;
    tok_len[(size_t)c] = (uint16_t)L;
    split[(size_t)c] = (uint32_t)a | ((uint32_t)b << 16);
  }
  // the tokens' characters, then the trie
  for (int64_t t = 0; t < n_tok; t++) off[(size_t)t + 1] = off[(size_t)t] + tok_len[(size_t)t];
  if (off[(size_t)n_tok] > (64u << 20)) return BPE_OK;
  std::vector<uint16_t> text(off[(size_t)n_tok]);
  for (int64_t t = 0; t < A; t++) text[off[(size_t)t]] = (uint16_t)t;
  for (size_t r = 0; r < m; r++) {
    const size_t a = (size_t)e->h_merges[3 * r], b = (size_t)e->h_merges[3 * r + 1], c = (size_t)e->h_merges[3 * r + 2];
    std::copy(text.begin() + off[a], text.begin() + off[a + 1], text.begin() + off[c]);
    std::copy(text.begin() + off[b], text.begin() + off[b + 1], text.begin() + off[c] + tok_len[a]);
  }
  std::unordered_map<uint64_t, uint32_t> child;  // (node << 16 | char) -> node
  child.reserve(text.size() * 2);
  std::vector<uint16_t> node_tok(1, (uint16_t)DP_NONE);
  std::vector<uint32_t> node_of((size_t)n_tok), parent(1, 0);
  std::vector<uint16_t> pch(1, 0);
  for (int64_t t = 0; t < n_tok; t++) {
    uint32_t v = 0;
    for (uint32_t i = off[(size_t)t]; i < off[(size_t)t + 1]; i++) {
      const uint64_t key = ((uint64_t)v << 16) | text[i];
      auto it = child.find(key);
      if (it == child.end()) {
        const uint32_t nv = (uint32_t)node_tok.size();
        child.emplace(key, nv);
        node_tok.push_back((uint16_t)DP_NONE);
        parent.push_back(v);
        pch.push_back(text[i]);
        v = nv;
      } else {
        v = it->second;
      }
    }
    if (node_tok[v] != DP_NONE) return BPE_OK;  // two tokens spell the same string: the string does not identify the token
    node_tok[v] = (uint16_t)t;
    node_of[(size_t)t] = v;
  }
  const size_t S = node_tok.size();
  if (S * (size_t)A > (64u << 20)) return BPE_OK;  // dense automaton too large (huge alphabets): lane path
  // Aho-Corasick as a dense automaton, nodes in order of depth: the row of a node is the row of its failure node (shallower,
  // hence complete) with its own children written over it; the failure node of a child u = (v, ch) is row[fail(v)][ch]
  std::vector<uint32_t> depth(S, 0), order(S);
  for (size_t v = 1; v < S; v++) depth[v] = depth[parent[v]] + 1;  // (a parent's index is smaller than its children's)
  for (size_t v = 0; v < S; v++) order[v] = (uint32_t)v;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return depth[x] < depth[y]; });
  std::vector<uint32_t> kid_begin(S + 1, 0), kids(S > 0 ? S - 1 : 0);
  for (size_t v = 1; v < S; v++) kid_begin[parent[v] + 1]++;
  for (size_t v = 0; v < S; v++) kid_begin[v + 1] += kid_begin[v];
  {
    std::vector<uint32_t> fill(kid_begin.begin(), kid_begin.end() - 1);
    for (size_t v = 1; v < S; v++) kids[fill[parent[v]]++] = (uint32_t)v;
  }
  std::vector<uint32_t> dfa(S * (size_t)A, 0), fail_(S, 0);
  std::vector<uint16_t> out_tok(S, (uint16_t)DP_NONE);
  for (uint32_t v : order) {
    if (v != 0) {
      out_tok[v] = node_tok[v] != DP_NONE ? node_tok[v] : out_tok[fail_[v]];
      std::copy(dfa.begin() + (size_t)fail_[v] * A, dfa.begin() + (size_t)(fail_[v] + 1) * A, dfa.begin() + (size_t)v * A);
    }
    for (uint32_t i = kid_begin[v]; i < kid_begin[v + 1]; i++) {
      const uint32_t u = kids[i];
      fail_[u] = v == 0 ? 0u : dfa[(size_t)fail_[v] * A + pch[u]];
      dfa[(size_t)v * A + pch[u]] = u;
    }
  }
  std::vector<uint16_t> shorter((size_t)n_tok, (uint16_t)DP_NONE);
  for (int64_t t = 0; t < n_tok; t++) shorter[(size_t)t] = out_tok[fail_[node_of[(size_t)t]]];
  // one 8-byte entry per transition: next state | longest token ending there << 32 | its length << 48
  std::vector<unsigned long long> dfa64(dfa.size());
  for (size_t i = 0; i < dfa.size(); i++) {
    const uint32_t nx = dfa[i];
    const uint16_t ot = out_tok[nx];
    dfa64[i] = (unsigned long long)nx | ((unsigned long long)ot << 32) | ((unsigned long long)(ot == DP_NONE ? 0 : tok_len[ot]) << 48);
  }
  CK(e->d_dfa.reserve(dfa64.size()));
  CK(cudaMemcpyAsync(e->d_dfa.p, dfa64.data(), dfa64.size() * 8, cudaMemcpyHostToDevice, e->stream));
  CK(e->d_tok_len.reserve((size_t)n_tok));
  CK(cudaMemcpyAsync(e->d_tok_len.p, tok_len.data(), (size_t)n_tok * 2, cudaMemcpyHostToDevice, e->stream));
  CK(e->d_shorter.reserve((size_t)n_tok));
  CK(cudaMemcpyAsync(e->d_shorter.p, shorter.data(), (size_t)n_tok * 2, cudaMemcpyHostToDevice, e->stream));
  CK(e->d_split.reserve((size_t)n_tok));
  CK(cudaMemcpyAsync(e->d_split.p, split.data(), (size_t)n_tok * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->dp_alpha = (uint32_t)A;
  e->dp_ok = true;
  return BPE_OK;
}

template <int LMAX, int WARPS>
int launch_encode_lanes(bpe_engine* e, const int32_t* dev_ids, const int64_t* dev_doc_off, int64_t n_docs, const uint32_t* range_first,
                        uint32_t n_ranges, const LaneTables& lt, int32_t* out_tmp, uint32_t* out_len, uint32_t* n_long, uint32_t* err,
                        uint32_t* next_range) {
  size_t smem = (size_t)2 * EL_DENSE * EL_DENSE * 4 + (size_t)WARPS * (LMAX * 32 * 2 + LMAX * 16) * 4;
  auto kern = k_encode_lanes<LMAX, WARPS>;
  const uint32_t inst_bit = 1u << (LMAX / 4 - 4);  // LMAX = 16, 20, 24, 32, 48 (one WARPS each): bits 0, 1, 2, 4, 8
  if (!(e->lanes_attr_set & inst_bit)) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    e->lanes_attr_set |= inst_bit;
  }
  int per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
  per_sm = std::max(per_sm, 1);
  int64_t blocks = std::min<int64_t>(((int64_t)n_ranges + WARPS - 1) / WARPS, (int64_t)e->sm_count * per_sm);
  kern<<<(int)std::max<int64_t>(blocks, 1), WARPS * 32, smem, e->stream>>>(dev_ids, dev_doc_off, n_docs, range_first, n_ranges, lt, out_tmp,
                                                                       out_len, n_long, err, next_range);
  CKL();
  return BPE_OK;
}


// streams, events and pinned flag words of the host-buffer encode pipeline (encode_pipeline below)
int ensure_pipe(bpe_engine* e) {
  if (e->h_pipe) return BPE_OK;
  CK(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
  for (cudaEvent_t* ev : {&e->ev_in[0], &e->ev_in[1], &e->ev_in[2], &e->ev_done, &e->ev_res[0], &e->ev_res[1]})
    CK(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  if (const char* v = getenv("BPE_ENC_CHUNK"))
    if (atoll(v) > 0) e->enc_chunk = atoll(v);
  CK(cudaHostAlloc((void**)&e->h_pipe, 16 * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
  return BPE_OK;
}

// device-resident encode; leaves compacted output in dev_out / dev_out_offsets
int encode_dev(bpe_engine* e, EncodeScratch& sc, const int32_t* dev_ids, const int64_t* dev_doc_off, int64_t n_docs,
               int64_t n_ids, int64_t max_doc_len, const int32_t* dev_tvi, int32_t n_tvi, int32_t* dev_out,
               int64_t* dev_out_offsets, int64_t* dev_first_bad, int64_t* n_out, const std::function<int()>* overlap = nullptr) {
  // `overlap`: host work of the caller that runs while the lane kernel does (the encode pipeline prepares its next chunk there)
  int overlap_rc = BPE_OK;
  if (e->h_merges.size() / 3 > EL_MAX_RANK + 1) return fail(e, BPE_E_DOMAIN, "merge list too long");
  TRY(ensure_lane_tables(e));
  TRY(ensure_dp_tables(e));
  TRY(ensure_pipe(e));
  CK(sc.out_tmp.reserve((size_t)std::max<int64_t>(n_ids, 1)));
  CK(sc.out_len.reserve((size_t)n_docs + 1));
  CK(sc.out_off.reserve((size_t)n_docs + 2));
  CK(sc.flags.reserve(4));
  if (!e->ev0) {
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
  }
  CK(cudaEventRecord(e->ev0, e->stream));
  bool run_old = e->enc_force_old != 0;
  if (n_docs > 0 && !run_old) {
    // lane path: one warp per batch of whole documents; position ranges of `stride` ids name the batches
    // rows per lane: the smallest batch that still holds the longest document keeps the most warps resident (the kernel is
    // latency bound: 26.4 GB/s with 20 rows / 32 warps per SM vs 22.5 GB/s with 32 rows / 20 warps on the 1 GB text)
    int lmax = e->enc_lmax;
    if (!e->enc_lmax_forced) lmax = max_doc_len <= 512 ? 16 : max_doc_len <= 640 ? 20 : max_doc_len <= 768 ? 24 : max_doc_len <= 1024 ? 32 : 48;
    const uint32_t cap = 32u * (uint32_t)lmax;
    uint32_t stride = 4u * cap;  // a range = the documents starting in `stride` consecutive positions, packed greedily into batches
    uint64_t nr64 = ((uint64_t)n_ids + stride - 1) / stride;
    if (nr64 == 0) nr64 = 1;
    if (nr64 > 0x7FFFFFF0ull || n_docs > 0x7FFFFFF0ll) return fail(e, BPE_E_DOMAIN, "batch too large for one encode call");
    uint32_t n_ranges = (uint32_t)nr64;
    CK(sc.range_first.reserve((size_t)n_ranges + 1));
    CK(cudaMemsetAsync(sc.flags.p, 0, 4 * sizeof(uint32_t), e->stream));
    k_range_starts<<<(int)std::min<uint32_t>((n_ranges + 256) / 256, (uint32_t)e->sm_count * 8), 256, 0, e->stream>>>(dev_doc_off, n_docs, stride, n_ranges,
                                                                                                             sc.range_first.p);
    CKL();
    LaneTables lt{e->d_lt.p, e->lt_cap - 1, (uint32_t)(32 - ilog2(e->lt_cap)), e->d_lt_dense.p, e->d_rule_c.p, e->lt_c_affine};
    e->stats.encode_path = (e->enc_dp && e->dp_ok) ? 1 : 2;
    if (e->enc_dp && e->dp_ok) {
      // forward path: one left-to-right pass per document (encode_dp.cuh), one thread per document
      constexpr int DPW = 16;
      const size_t dp_smem_bytes = (size_t)DP_CACHE * 8 + (size_t)EL_DENSE * EL_DENSE * sizeof(uint2);
      if (!e->dp_attr_set) {
        CK(cudaFuncSetAttribute(k_encode_dp<DPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp_smem_bytes));
        e->dp_attr_set = true;
      }
      const unsigned long long cap_entries = 3ull * (unsigned long long)n_ids + (1ull << 20);  // (a warp of 32 documents needs 32 x its longest)
      CK(sc.dp_last.reserve((size_t)cap_entries));
      CK(sc.dp_cursor.reserve(1));
      CK(cudaMemsetAsync(sc.dp_cursor.p, 0, 8, e->stream));
      DpTables dt{e->d_dfa.p, e->d_tok_len.p, e->d_shorter.p, e->d_split.p, e->dp_alpha};
      int per_sm = 1;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_dp<DPW>, DPW * 32, dp_smem_bytes));
      const int64_t want = (n_docs + DPW * 32 - 1) / (DPW * 32);
      const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)e->sm_count * std::max(per_sm, 1)));
      k_encode_dp<DPW><<<blocks, DPW * 32, dp_smem_bytes, e->stream>>>(dev_ids, dev_doc_off, n_docs, lt, dt, sc.dp_last.p, cap_entries, sc.dp_cursor.p, sc.out_tmp.p,
                                                          sc.out_len.p, sc.flags.p, sc.flags.p + 1);
      CKL();
    } else if (lmax == 16)
      TRY((launch_encode_lanes<16, 20>(e, dev_ids, dev_doc_off, n_docs, sc.range_first.p, n_ranges, lt, sc.out_tmp.p, sc.out_len.p, sc.flags.p, sc.flags.p + 1, sc.flags.p + 2)));
    else if (lmax == 20)
      TRY((launch_encode_lanes<20, 16>(e, dev_ids, dev_doc_off, n_docs, sc.range_first.p, n_ranges, lt, sc.out_tmp.p, sc.out_len.p, sc.flags.p, sc.flags.p + 1, sc.flags.p + 2)));
    else if (lmax == 24)
      TRY((launch_encode_lanes<24, 13>(e, dev_ids, dev_doc_off, n_docs, sc.range_first.p, n_ranges, lt, sc.out_tmp.p, sc.out_len.p, sc.flags.p, sc.flags.p + 1, sc.flags.p + 2)));
    else if (lmax == 32)
      TRY((launch_encode_lanes<32, 10>(e, dev_ids, dev_doc_off, n_docs, sc.range_first.p, n_ranges, lt, sc.out_tmp.p, sc.out_len.p, sc.flags.p, sc.flags.p + 1, sc.flags.p + 2)));
    else
      TRY((launch_encode_lanes<48, 6>(e, dev_ids, dev_doc_off, n_docs, sc.range_first.p, n_ranges, lt, sc.out_tmp.p, sc.out_len.p, sc.flags.p, sc.flags.p + 1, sc.flags.p + 2)));
    volatile uint32_t* hf = reinterpret_cast<volatile uint32_t*>(e->h_pipe + 12);  // flags[0..1], written by the device (k_copy_u64)
    k_copy_u64<<<1, 32, 0, e->stream>>>(e->h_pipe + 12, reinterpret_cast<const unsigned long long*>(sc.flags.p), 1);
    CKL();
    if (overlap) {
      overlap_rc = (*overlap)();
      overlap = nullptr;
    }
    CK(cudaStreamSynchronize(e->stream));
    if (overlap_rc != BPE_OK) return overlap_rc;
    if (hf[1]) return fail(e, BPE_E_INTERNAL, "encode rounds did not converge");
    run_old = hf[0] != 0;
  }
  if (e->enc_force_old) e->stats.encode_path = 3;
  if (n_docs > 0 && run_old) {
    // per-document kernel: documents longer than a lane batch (marked EL_LONG), or everything when forced
    TRY(ensure_merge_table(e));
    if (max_doc_len > ENC_WARP_MAX) {
      CK(sc.g_tok.reserve((size_t)n_ids));
      CK(sc.g_nxt.reserve((size_t)n_ids));
      CK(sc.g_prv.reserve((size_t)n_ids));
      CK(sc.g_rk.reserve((size_t)n_ids));
      CK(sc.g_sel.reserve((size_t)n_ids));
    }
    EncTables mt{MergeTable{e->d_mt.p, e->mt_cap - 1, (uint32_t)(32 - ilog2(e->mt_cap))}, e->d_minlr.p};
    int64_t blocks = std::min<int64_t>((n_docs + ENC_WARPS - 1) / ENC_WARPS, (int64_t)e->sm_count * 16);
    k_encode<<<(int)blocks, ENC_THREADS, 0, e->stream>>>(dev_ids, dev_doc_off, n_docs, mt, sc.out_tmp.p, sc.out_len.p, sc.g_tok.p, sc.g_nxt.p, sc.g_prv.p,
                                                         sc.g_rk.p, sc.g_sel.p, e->enc_force_old ? 0 : 1);
    CKL();
  }
  if (n_docs <= 65536) {
    k_scan_counts<<<1, 1024, 0, e->stream>>>(sc.out_len.p, sc.out_off.p, (uint32_t)n_docs);
    CKL();
  } else {
    uint32_t n_tiles = (uint32_t)((n_docs + SC_TILE - 1) / SC_TILE);
    CK(sc.tile_sums.reserve(n_tiles));
    CK(sc.tile_off.reserve((size_t)n_tiles + 1));
    k_sum_tiles<<<n_tiles, SC_THREADS, 0, e->stream>>>(sc.out_len.p, (uint32_t)n_docs, sc.tile_sums.p);
    CKL();
    k_scan_counts<<<1, 1024, 0, e->stream>>>(sc.tile_sums.p, sc.tile_off.p, n_tiles);
    CKL();
    k_scan_tiles<<<n_tiles, SC_THREADS, 0, e->stream>>>(sc.out_len.p, (uint32_t)n_docs, sc.tile_off.p, n_tiles, sc.out_off.p);
    CKL();
  }
  {
    int64_t warps = n_docs + 1;
    int64_t blocks = std::min<int64_t>((warps + 3) / 4, (int64_t)e->sm_count * 16);
    k_gather_map<<<(int)blocks, 128, 0, e->stream>>>(sc.out_tmp.p, dev_doc_off, sc.out_off.p, n_docs, dev_tvi, n_tvi, dev_out, dev_out_offsets, dev_first_bad);
    CKL();
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  if (overlap) overlap_rc = (*overlap)();  // the lane kernel did not run
  k_copy_u64<<<1, 32, 0, e->stream>>>(e->h_pipe + 13, reinterpret_cast<const unsigned long long*>(sc.out_off.p + n_docs), 1);
  CKL();
  CK(cudaStreamSynchronize(e->stream));
  if (overlap_rc != BPE_OK) return overlap_rc;
  const uint64_t total = *reinterpret_cast<volatile unsigned long long*>(e->h_pipe + 13);
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->stats.ms_encode = ms;
  *n_out = (int64_t)total;
  return BPE_OK;
}

int check_offsets(bpe_engine* e, const int64_t* off, int64_t n_docs) {
  if (n_docs < 0 || (!off && n_docs > 0)) return fail(e, BPE_E_INVALID, "bad document offsets");
  for (int64_t d = 0; d < n_docs; d++)
    if (off[d + 1] < off[d]) return fail(e, BPE_E_INVALID, "document offsets must be non-decreasing (doc %lld)", (long long)d);
  return BPE_OK;
}

// append ids (device pointer, already merged or raw) as documents
int append_docs_dev(bpe_engine* e, const int32_t* dev_ids, const int64_t* host_off, int64_t n_docs) {
  int64_t base_in = host_off[0];
  int64_t total = host_off[n_docs] - base_in;
  uint64_t new_n = e->n_slots + (uint64_t)total;
  if (new_n >= 0xFFFFFFF0ull) return fail(e, BPE_E_DOMAIN, "corpus exceeds 2^32 positions per engine");
  CK(e->slots.reserve((size_t)new_n + 4, (size_t)e->n_slots, e->stream, 1.5));
  CK(e->d_st.reserve(1));
  if (total > 0) {
    CK(cudaMemsetAsync(&e->d_st.p->err, 0, sizeof(uint32_t), e->stream));
    int blocks = (int)std::min<int64_t>((total / 4 + 255) / 256 + 1, (int64_t)e->grid(8));
    if (reinterpret_cast<const uint32_t*>(dev_ids + base_in) == e->slots.p + e->n_slots)
      k_validate_ids<<<blocks, 256, 0, e->stream>>>(e->slots.p + e->n_slots, (uint64_t)total, (uint32_t)e->n_tokens, &e->d_st.p->err);
    else
      k_ingest_ids<<<blocks, 256, 0, e->stream>>>(dev_ids + base_in, e->slots.p + e->n_slots, (uint64_t)total, (uint32_t)e->n_tokens, &e->d_st.p->err);
    CKL();
    CK(e->stage_off.reserve((size_t)n_docs + 1, 0, e->stream, 1.5));
    CK(cudaMemcpyAsync(e->stage_off.p, host_off, (size_t)(n_docs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, e->stream));
    int blocks2 = (int)std::min<int64_t>((n_docs + 255) / 256, (int64_t)e->grid(8));
    k_mark_docstarts<<<blocks2, 256, 0, e->stream>>>(e->slots.p, e->stage_off.p, n_docs, (int64_t)e->n_slots - base_in);
    CKL();
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, &e->d_st.p->err, sizeof bad, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (bad) {
      CK(cudaMemsetAsync(&e->d_st.p->err, 0, sizeof(uint32_t), e->stream));
      return fail(e, BPE_E_INVALID, "document holds a token index >= n_tokens (%d); call bpe_set_tokens first", e->n_tokens);
    }
  }
  for (int64_t d = 0; d < n_docs; d++) e->doc_off.push_back((int64_t)e->n_slots + (host_off[d + 1] - base_in));
  e->n_slots = new_n;
  e->live_tokens += (uint64_t)total;
  e->index_valid = false;
  e->hot_valid = false;
  return BPE_OK;
}


// buffers of k_merge_rounds (round_kernels.cuh): delta cells, slot rows, per-block cell lists, site buffers, partials
int ensure_round_buffers(bpe_engine* e) {
  if (!e->round_blocks) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_merge_rounds, RD_THREADS, 0));
    if (per_sm < 1) return fail(e, BPE_E_CUDA, "k_merge_rounds does not fit on an SM");
    // (a decision folds RT partial entries per block with one thread each)
    const int most = std::min({e->sm_count * per_sm, (int)(R_QCAP / RT), RD_THREADS / RT});
    e->round_blocks = std::min(e->sm_count, most);
    if (const char* v = getenv("BPE_LOOP_BLOCKS")) e->round_blocks = std::max(1, std::min(atoi(v), most));
  }
  const size_t cells = (size_t)2 * RB * 2 * ND_STRIDE;
  if (!e->r_cells.p) {
    CK(e->r_cells.reserve(cells));
    CK(cudaMemsetAsync(e->r_cells.p, 0, cells * 8, e->stream));  // the kernel keeps the cells zero between launches
    CK(e->r_slotrows.reserve(cells));
    CK(e->r_lists.reserve((size_t)2 * e->round_blocks * R_LISTCAP));
    CK(e->r_bsites.reserve((size_t)2 * RB * R_SMALL));
    CK(e->r_state.reserve(1));
    CK(cudaMemsetAsync(e->r_state.p, 0, sizeof(RoundState), e->stream));
  }
  if (e->mg_world > 1 && !e->r_gcells.p) {
    CK(e->r_gcells.reserve(cells));
    CK(cudaMemsetAsync(e->r_gcells.p, 0, cells * 8, e->stream));
    CK(e->r_glists.reserve((size_t)2 * e->round_blocks * R_LISTCAP));
  }
  CK(e->r_gp.reserve((size_t)RT * e->round_blocks));
  CK(e->r_gk.reserve((size_t)RT * e->round_blocks));
  return BPE_OK;
}

RoundArgs round_args(bpe_engine* e, const LoopArgs& L, int round_k) {
  RoundArgs RA{};
  RA.L = L;
  RA.cells = e->r_cells.p;
  RA.slotrows = e->r_slotrows.p;
  RA.lists = e->r_lists.p;
  RA.bsites = e->r_bsites.p;
  RA.gp = e->r_gp.p;
  RA.gk = e->r_gk.p;
  RA.rs = e->r_state.p;
  RA.kmax = (uint32_t)round_k;
  RA.bar_mode = 1;
  if (const char* v = getenv("BPE_LOOP_BAR")) RA.bar_mode = atoi(v) ? 1 : 0;
  RA.mg_on = 0;
  RA.gcells = e->r_gcells.p;
  RA.glists = e->r_glists.p;
  return RA;
}

void round_stats(bpe_engine* e, const RoundState& hrs) {
  e->stats.loop_rounds = (int64_t)hrs.rounds;
  e->stats.loop_round_merges = (int64_t)hrs.round_merges;
  e->stats.loop_round_tried = (int64_t)hrs.tried;
  e->stats.loop_rounds_cut = (int64_t)hrs.rounds_cut_born;
}

// ---- mergeUntil: persistent cooperative kernel, the host only grows buffers / rebuilds the hot list -----------
int merge_until_device(bpe_engine* e, int64_t min_weight, int32_t max_length, int64_t max_iterations, bpe_merge* log,
                       int64_t log_cap, int64_t* n_done, const int32_t* dev_replay = nullptr) {
  *n_done = 0;
  CK(cudaSetDevice(e->device));
  if (e->n_slots == 0) return BPE_OK;
  TRY(ensure_index(e));
  uint32_t ml = max_length > 0 ? (uint32_t)max_length : 0;
  int64_t mw = min_weight > 0 ? min_weight : 2;
  if (!e->loop_blocks) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_merge_loop, ML_THREADS, 0));
    if (per_sm < 1) return fail(e, BPE_E_CUDA, "k_merge_loop does not fit on an SM");
    // one 512-thread block per SM: measured best on cfg3 (1.28 s vs 1.41 s with two per SM, 1.33 s with half an SM count,
    // 1.55 s with 1024-thread blocks) -- a merge is latency bound, fewer arrivals per grid barrier and fewer partials win
    int want = 1;
    if (const char* v = getenv("BPE_LOOP_PER_SM")) want = std::max(1, atoi(v));  // tuning knob: co-resident blocks per SM
    e->loop_blocks = e->sm_count * std::min(per_sm, want);
    if (const char* v = getenv("BPE_LOOP_BLOCKS")) e->loop_blocks = std::max(1, std::min(atoi(v), e->sm_count * per_sm));
  }
  // several exact merges per barrier round (round_kernels.cuh) unless a merge list is replayed; BPE_LOOP_ROUNDS=0 keeps k_merge_loop
  bool use_rounds = !dev_replay;
  int round_k = RB;
  if (const char* v = getenv("BPE_LOOP_ROUNDS")) use_rounds = use_rounds && atoi(v) != 0;
  if (const char* v = getenv("BPE_LOOP_K")) round_k = std::max(1, std::min(atoi(v), (int)RB));
  if (use_rounds) TRY(ensure_round_buffers(e));
  CK(e->partials.reserve((size_t)std::max(e->loop_blocks, e->grid(8))));
  CK(e->partial_keys.reserve((size_t)std::max(e->loop_blocks, e->grid(8))));
  CK(e->sites.reserve(1u << 16, 0, e->stream, 1.0));
  CK(e->sites2.reserve(std::max<size_t>(e->sites.cap, 1u << 16), 0, e->stream, 1.0));
  CK(e->newslots.reserve(1u << 16, 0, e->stream, 1.0));
  CK(e->cands.reserve(4096, 0, e->stream, 1.0));
  CK(e->barrier.reserve(64));
  cudaEvent_t t0, t1;
  CK(cudaEventCreate(&t0));
  CK(cudaEventCreate(&t1));
  CK(cudaEventRecord(t0, e->stream));
  int rc = BPE_OK;
  int64_t done = 0;
  bool legacy_above = false;
  std::vector<MergeRec> tmp;
  for (;;) {
    int64_t remaining = log_cap - done;
    if (max_iterations > 0) remaining = std::min(remaining, max_iterations - done);  // core.ts:374-377
    if (remaining <= 0) break;
    if (!dev_replay && (!e->hot_valid || e->hot_max_length != ml)) {
      bool any = false;
      if ((rc = rebuild_hot(e, ml, &any)) != BPE_OK) break;
      if (!any) break;  // nothing countable left (core.ts:312)
    }
    uint32_t chunk = (uint32_t)std::min<int64_t>(remaining, 1 << 16);
    if (e->n_tokens + (int64_t)chunk > BPE_MAX_TOKENS) chunk = (uint32_t)std::max<int64_t>(0, BPE_MAX_TOKENS - e->n_tokens);
    if (chunk == 0) {
      rc = fail(e, BPE_E_DOMAIN, "token table would exceed %d", BPE_MAX_TOKENS);
      break;
    }
    cudaError_t ce;
    if ((ce = e->dev_log.reserve(chunk)) != cudaSuccess ||
        (ce = e->d_len16.reserve((size_t)e->n_tokens + chunk + 1, (size_t)e->n_tokens, e->stream, 1.5)) != cudaSuccess) {
      rc = fail(e, BPE_E_NOMEM, "%s", cudaGetErrorString(ce));
      break;
    }
    LoopArgs L;
    L.A = apply_args(e);
    L.pool_cap = (uint32_t)std::min<size_t>(e->pool.cap, 0xFFFFFFF0u);
    L.len16_cap = (uint32_t)std::min<size_t>(e->d_len16.cap, 0xFFFFFFF0u);
    L.hot = e->hot.p;
    L.hot_cap = (uint32_t)std::min<size_t>(e->hot.cap, 0xFFFFFFF0u);
    L.hot_limit = e->hot_limit;
    L.cands = e->cands.p;
    L.cand_cap = (uint32_t)std::min<size_t>(e->cands.cap, 0xFFFFFFF0u);
    L.partials = e->partials.p;
    L.partial_keys = e->partial_keys.p;
    L.barrier = e->barrier.p;
    L.log = e->dev_log.p;
    L.log_cap = chunk;
    L.max_length = ml;
    L.min_weight = (uint32_t)std::min<int64_t>(mw, 0xFFFFFFFFll);
    L.max_tokens = BPE_MAX_TOKENS;
    L.tbl_cap = e->tbl_cap;
    L.sites2 = e->sites2.p;
    L.A.sites_cap = (uint32_t)std::min<size_t>(L.A.sites_cap, e->sites2.cap);  // (only this kernel alternates between the two)
    {  // warp split of the latency-bound phases and barrier back-off (tuning knobs, read per launch)
      // measured on cfg3 (profiles/r01i): born pairs / rewrite / old-hot arg-max = 10/2/4 warps beats 8/4/4 by 1 %, 6/x loses 1-2 %;
      // 10..14 site warps and 64..512 ns of back-off are all within noise
      int a = 12, b = 10, c = 2, ns = 256;
      if (const char* v = getenv("BPE_LOOP_P1_SITES")) a = atoi(v);
      if (const char* v = getenv("BPE_LOOP_P2_NEW")) b = atoi(v);
      if (const char* v = getenv("BPE_LOOP_P2_RW")) c = atoi(v);
      if (const char* v = getenv("BPE_LOOP_BAR_NS")) ns = atoi(v);
      if (a < 1 || a > 15) a = 12;
      if (b < 1 || c < 1 || b + c > 15) b = 10, c = 2;
      L.p1_sites = (uint32_t)a;
      L.p2_new = (uint32_t)b;
      L.p2_rw = (uint32_t)c;
      L.bar_ns = (uint32_t)std::max(32, std::min(ns, 4096));
    }
    const char* pf = getenv("BPE_LOOP_PREFETCH");  // read per launch: a tuning knob
    L.prefetch = pf ? std::max(0, std::min(atoi(pf), 2)) : (use_rounds ? 0 : 1);  // (k_merge_rounds: measured, the helper warps become the tail)
    L.replay = dev_replay ? dev_replay + 2 * done : nullptr;
    if (dev_replay) e->hot_valid = false;  // replayed merges do not feed the hot list
    static const bool trace = getenv("BPE_TRACE") != nullptr;
    auto tw0 = std::chrono::steady_clock::now();
    k_loop_prepare<<<1, 32, 0, e->stream>>>(e->d_st.p, (uint32_t)e->n_tokens, e->barrier.p);
    e->stats.kernel_launches++;
    if (use_rounds && legacy_above) {
      // the winner has more sites than the packed delta cells of a round can count: k_merge_loop takes the merges above
      // R_HUGE (it stops, "done", at the first winner at or below it) and the rounds continue from there
      L.min_weight = (uint32_t)std::max<int64_t>(mw, (int64_t)R_HUGE + 1);
      void* args[] = {&L};
      ce = cudaLaunchCooperativeKernel((void*)k_merge_loop, dim3(e->loop_blocks), dim3(ML_THREADS), args, 0, e->stream);
    } else if (use_rounds) {
      RoundArgs RA = round_args(e, L, round_k);
      k_rounds_prepare<<<1, 32, 0, e->stream>>>(e->r_state.p);
      e->stats.kernel_launches++;
      void* rargs[] = {&RA};
      ce = cudaLaunchCooperativeKernel((void*)k_merge_rounds, dim3(e->round_blocks), dim3(RD_THREADS), rargs, 0, e->stream);
    } else {
      void* args[] = {&L};
      ce = cudaLaunchCooperativeKernel((void*)k_merge_loop, dim3(e->loop_blocks), dim3(ML_THREADS), args, 0, e->stream);
    }
    if (ce != cudaSuccess) {
      rc = fail(e, BPE_E_CUDA, "cooperative launch of the mergeUntil kernel: %s", cudaGetErrorString(ce));
      break;
    }
    e->stats.kernel_launches++;
    if ((rc = fetch_state(e)) != BPE_OK) break;
    uint32_t iters = e->h_st->iters_done;
    if (trace) {
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
      fprintf(stderr, "[bpe] k_merge_loop: %u merges in %.2f ms, status %u, best_cnt %u, hot_n %u thresh %u, keys %u/%u\n", iters, ms,
              e->h_st->status, e->h_st->best_cnt, e->h_st->hot_n, e->h_st->hot_thresh, e->h_st->n_keys, e->tbl_cap);
    }
    if (iters) {
      tmp.resize(iters);
      ce = cudaMemcpyAsync(tmp.data(), e->dev_log.p, (size_t)iters * sizeof(MergeRec), cudaMemcpyDeviceToHost, e->stream);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
      if (ce != cudaSuccess) {
        rc = fail(e, BPE_E_CUDA, "merge log read-back: %s", cudaGetErrorString(ce));
        break;
      }
      for (uint32_t i = 0; i < iters; i++) {
        const MergeRec& r = tmp[i];
        log[done].a = r.a;
        log[done].b = r.b;
        log[done].c = r.c;
        log[done].reserved = 0;
        log[done].weight = r.weight;
        done++;
        e->h_len16.push_back(e->h_len16[r.a] + e->h_len16[r.b]);
        e->h_merges.push_back(r.a);
        e->h_merges.push_back(r.b);
        e->h_merges.push_back(r.c);
      }
      e->n_tokens += (int32_t)iters;
      e->mt_dirty = e->lt_dirty = e->dp_dirty = true;
      e->stats.merges_applied += iters;
    }
    uint32_t status = e->h_st->status;
    if (use_rounds && legacy_above && status == LOOP_DONE && (int64_t)e->h_st->best_cnt >= mw) {
      legacy_above = false;  // below R_HUGE now: back to the rounds
      continue;
    }
    if (status == LOOP_NEED_LEGACY) {
      legacy_above = true;
      continue;
    }
    if (status == LOOP_DONE || status == LOOP_EMPTY) break;
    if (status == LOOP_LIMIT) continue;
    if (status == LOOP_NEED_REBUILD) {
      e->hot_valid = false;
      continue;
    }
    if (status == LOOP_ERROR) {
      rc = check_dev_err(e);
      if (rc == BPE_OK) rc = fail(e, BPE_E_INTERNAL, "merge loop stopped with an unexplained error (flags 0x%x)", e->h_st->err);
      break;
    }
    if (status == LOOP_NEED_HOST) {
      if (e->n_tokens >= BPE_MAX_TOKENS) {
        rc = fail(e, BPE_E_DOMAIN, "token table would exceed %d", BPE_MAX_TOKENS);
        break;
      }
      uint64_t w = e->h_st->best_cnt;
      uint64_t new_keys = std::min<uint64_t>(2 * w + 2, 2 * ((uint64_t)e->n_tokens + 1) + 2);
      uint64_t keys_after = (uint64_t)e->h_st->n_keys + new_keys;
      if (keys_after * 2 > e->tbl_cap) {
        uint64_t want = std::min<uint64_t>(keys_after * 5 / 2, 0x80000000ull);  // next power of two: load 0.2 .. 0.4
        // a run over a big corpus ends with distinct pairs ~ 5 % of its positions (48 M for the 1 GB Zipf corpus): the first
        // growth goes most of the way instead of doubling eleven times (a relaunch + rehash each)
        want = std::max<uint64_t>(want, std::min<uint64_t>(e->n_slots / 16, 1ull << 27));
        if (keys_after * 2 > pow2_at_least(want)) {
          rc = fail(e, BPE_E_NOMEM, "pair table cannot grow further");
          break;
        }
        if ((rc = grow_table(e, pow2_at_least(want))) != BPE_OK) break;
      }
      uint64_t pool_after = (uint64_t)e->h_st->pool_cursor + 2 * w;
      if (use_rounds) pool_after += (uint64_t)e->round_blocks * R_POOL_CHUNK;  // every block of k_merge_rounds may open a private chunk
      if (pool_after > 0xFFFFFFF0ull) {
        rc = fail(e, BPE_E_DOMAIN, "occurrence pool exceeds 2^32 cells");
        break;
      }
      if ((ce = e->pool.reserve((size_t)pool_after, e->h_st->pool_cursor, e->stream, 1.5)) != cudaSuccess ||
          (ce = e->sites.reserve((size_t)w, 0, e->stream, 1.25)) != cudaSuccess ||
          (ce = e->sites2.reserve(e->sites.cap, 0, e->stream, 1.0)) != cudaSuccess ||
          (ce = e->newslots.reserve((size_t)new_keys, 0, e->stream, 1.25)) != cudaSuccess ||
          (ce = e->hot.reserve((size_t)e->h_st->hot_n + new_keys, e->h_st->hot_n, e->stream, 1.5)) != cudaSuccess ||
          (ce = e->cands.reserve((size_t)e->h_st->best_mult, 0, e->stream, 1.5)) != cudaSuccess) {
        rc = fail(e, BPE_E_NOMEM, "growing merge buffers: %s", cudaGetErrorString(ce));
        break;
      }
      continue;
    }
    rc = fail(e, BPE_E_INTERNAL, "merge loop returned status %u", status);
    break;
  }
  *n_done = done;
  if (rc == BPE_OK) {
    CK(cudaEventRecord(t1, e->stream));
    TRY(fetch_state(e));
    TRY(check_dev_err(e));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    e->stats.ms_last_merge_until = ms;
    e->live_tokens = e->h_st->live_tokens;
    e->stats.sites_merged = (int64_t)e->h_st->sites_total;
    e->stats.tie_breaks = e->h_st->tie_breaks;
    if (use_rounds) {
      RoundState hrs;
      CK(cudaMemcpyAsync(&hrs, e->r_state.p, sizeof(RoundState), cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
      round_stats(e, hrs);
      static const bool trace_r = getenv("BPE_TRACE") != nullptr;
      if (trace_r)
        fprintf(stderr, "[bpe] rounds %llu, merges %llu (tried %llu), cut by the born-pair bound %llu, single %llu; batch ends: cap %llu, no-candidate %llu, tie %llu, big %llu, "
                        "token %llu, fresh-token %llu, limits %llu\n", hrs.rounds, hrs.round_merges, hrs.tried, hrs.rounds_cut_born, hrs.rounds_single, hrs.stop_reason[0],
                hrs.stop_reason[1], hrs.stop_reason[2], hrs.stop_reason[3], hrs.stop_reason[4], hrs.stop_reason[5], hrs.stop_reason[6]);
      if (trace_r)
        fprintf(stderr, "[bpe] P1 warp-iterations small %llu big %llu; touched cells small %llu big %llu; sites small %llu big %llu\n", hrs.iters_small, hrs.iters_big,
                hrs.cells_small, hrs.cells_big, hrs.sites_small, hrs.sites_big);
      if (trace_r) {
        const unsigned long long* m = e->h_st->mg_prof_ns;
        const double n = std::max(1.0, (double)m[4]);
        fprintf(stderr, "[bpe] P2 of rounds of small merges, block 0, us after the barrier: cells done %.1f, rewrite done %.1f, arg-max done %.1f, partials out %.1f\n",
                m[0] / n * 1e-3, m[1] / n * 1e-3, m[2] / n * 1e-3, m[3] / n * 1e-3);
      }
      if (trace_r) {
        const unsigned long long* f = e->h_st->fine_ns;
        const double ns = std::max(1.0, (double)f[5]), nb = std::max(1.0, (double)f[11]);
        fprintf(stderr, "[bpe] block 0, us per round whose first merge has <= 16384 sites (%.0f rounds): decide %.1f, P1 %.1f, wait %.1f, P2 %.1f, wait %.1f; per round whose first merge has more (%.0f): "
                        "decide %.1f, P1 %.1f, wait %.1f, P2 %.1f, wait %.1f\n", (double)f[5], f[0] / ns * 1e-3, f[1] / ns * 1e-3, f[2] / ns * 1e-3, f[3] / ns * 1e-3, f[4] / ns * 1e-3,
                (double)f[11], f[6] / nb * 1e-3, f[7] / nb * 1e-3, f[8] / nb * 1e-3, f[9] / nb * 1e-3, f[10] / nb * 1e-3);
      }
    }
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return rc;
}


// ---- mergeUntil on a sharded corpus: k_merge_loop_mg, same host protocol as merge_until_device ---------------------
MgArgs mg_args(bpe_engine* e) {
  MgArgs M{};
  M.rank = e->mg_rank;
  M.world = e->mg_world;
  M.inbox_stride = e->mg_inbox_stride;
  M.tie_cap = e->mg_tie_cap;
  const size_t flags_bytes = (size_t)MG_MAX_WORLD * 128;
  const size_t inbox_bytes = (size_t)2 * e->mg_world * e->mg_inbox_stride * 8;
  for (int q = 0; q < e->mg_world; q++) {
    char* base = static_cast<char*>(e->mg_peer[q]);
    M.flag_data[q] = reinterpret_cast<unsigned long long*>(base);
    M.flag_tie[q] = reinterpret_cast<unsigned long long*>(base + flags_bytes);
    M.inbox[q] = reinterpret_cast<unsigned long long*>(base + 2 * flags_bytes);
    M.tiebox[q] = reinterpret_cast<uint32_t*>(base + 2 * flags_bytes + inbox_bytes);
  }
  M.tie_sorted = e->mg_tie_sorted.p;
  M.newpair = e->mg_newpair.p;
  M.newpair_cap = (uint32_t)e->mg_newpair.cap;
  return M;
}

int merge_until_mg(bpe_engine* e, int64_t min_weight, int32_t max_length, int64_t max_iterations, bpe_merge* log, int64_t log_cap,
                   int64_t* n_done) {
  *n_done = 0;
  CK(cudaSetDevice(e->device));
  if (!e->mg_connected) return fail(e, BPE_E_INVALID, "sharded engine: call bpe_mg_connect first");
  TRY(ensure_index(e));
  if (!e->mg_counts_global) return fail(e, BPE_E_INVALID, "sharded engine: exchange the pair counts first (bpe_mg_export_counts / bpe_mg_import_counts)");
  uint32_t ml = max_length > 0 ? (uint32_t)max_length : 0;
  int64_t mw = min_weight > 0 ? min_weight : 2;
  if (!e->mg_loop_blocks) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_merge_loop_mg, ML_THREADS, 0));
    if (per_sm < 1) return fail(e, BPE_E_CUDA, "k_merge_loop_mg does not fit on an SM");
    int want = 1;
    if (const char* v = getenv("BPE_LOOP_PER_SM")) want = std::max(1, atoi(v));
    e->mg_loop_blocks = e->sm_count * std::min(per_sm, want);
  }
  CK(e->partials.reserve((size_t)std::max(e->mg_loop_blocks, e->grid(8))));
  CK(e->sites.reserve(1u << 16, 0, e->stream, 1.0));
  CK(e->newslots.reserve(1u << 17, 0, e->stream, 1.0));
  CK(e->cands.reserve(4096, 0, e->stream, 1.0));
  CK(e->barrier.reserve(64));
  CK(e->mg_touched.reserve((size_t)4 * BPE_MAX_TOKENS + 64));
  CK(e->mg_tie_sorted.reserve(e->mg_tie_cap));
  CK(e->mg_newpair.reserve((size_t)2 * BPE_MAX_TOKENS + 64));
  // several exact merges per barrier round and ONE exchange per round (round_kernels.cuh); BPE_LOOP_ROUNDS=0 keeps k_merge_loop_mg
  bool use_rounds = true;
  int round_k = RB;
  if (const char* v = getenv("BPE_LOOP_ROUNDS")) use_rounds = atoi(v) != 0;
  if (const char* v = getenv("BPE_LOOP_K")) round_k = std::max(1, std::min(atoi(v), (int)RB));
  if (use_rounds) {
    TRY(ensure_round_buffers(e));
    CK(e->sites2.reserve(std::max<size_t>(e->sites.cap, 1u << 16), 0, e->stream, 1.0));
  }
  bool legacy_above = false;
  cudaEvent_t t0, t1;
  CK(cudaEventCreate(&t0));
  CK(cudaEventCreate(&t1));
  CK(cudaEventRecord(t0, e->stream));
  int rc = BPE_OK;
  int64_t done = 0;
  std::vector<MergeRec> tmp;
  double host_ms[4] = {0, 0, 0, 0};  // hot rebuilds, growth (NEED_HOST), launch+state fetch, log read-back
  auto clk = [] { return std::chrono::steady_clock::now(); };
  auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  for (;;) {
    int64_t remaining = log_cap - done;
    if (max_iterations > 0) remaining = std::min(remaining, max_iterations - done);  // core.ts:374-377
    if (remaining <= 0) break;
    if (!e->hot_valid || e->hot_max_length != ml) {
      bool any = false;
      auto th = clk();
      rc = rebuild_hot(e, ml, &any);
      host_ms[0] += since(th);
      if (rc != BPE_OK) break;
      if (!any) break;  // nothing countable left on ANY rank: the counts are global (core.ts:312)
    }
    uint32_t chunk = (uint32_t)std::min<int64_t>(remaining, 1 << 16);
    if (e->n_tokens + (int64_t)chunk > BPE_MAX_TOKENS) chunk = (uint32_t)std::max<int64_t>(0, BPE_MAX_TOKENS - e->n_tokens);
    if (chunk == 0) {
      rc = fail(e, BPE_E_DOMAIN, "token table would exceed %d", BPE_MAX_TOKENS);
      break;
    }
    cudaError_t ce;
    if ((ce = e->dev_log.reserve(chunk)) != cudaSuccess ||
        (ce = e->d_len16.reserve((size_t)e->n_tokens + chunk + 1, (size_t)e->n_tokens, e->stream, 1.5)) != cudaSuccess) {
      rc = fail(e, BPE_E_NOMEM, "%s", cudaGetErrorString(ce));
      break;
    }
    LoopArgsMg P;
    LoopArgs& L = P.L;
    L.A = apply_args(e);
    L.A.newpair = e->mg_newpair.p;
    L.A.newpair_cap = (uint32_t)e->mg_newpair.cap;
    L.A.dlt = e->mg_dlt.p;
    L.A.touched = e->mg_touched.p;
    L.A.touched_cap = (uint32_t)e->mg_touched.cap;
    L.pool_cap = (uint32_t)std::min<size_t>(e->pool.cap, 0xFFFFFFF0u);
    L.len16_cap = (uint32_t)std::min<size_t>(e->d_len16.cap, 0xFFFFFFF0u);
    L.hot = e->hot.p;
    L.hot_cap = (uint32_t)std::min<size_t>(e->hot.cap, 0xFFFFFFF0u);
    L.hot_limit = e->hot_limit;
    L.cands = e->cands.p;
    L.cand_cap = (uint32_t)std::min<size_t>(e->cands.cap, 0xFFFFFFF0u);
    L.partials = e->partials.p;
    L.barrier = e->barrier.p;
    L.log = e->dev_log.p;
    L.log_cap = chunk;
    L.max_length = ml;
    L.min_weight = (uint32_t)std::min<int64_t>(mw, 0xFFFFFFFFll);
    L.max_tokens = BPE_MAX_TOKENS;
    L.tbl_cap = e->tbl_cap;
    L.replay = nullptr;
    P.M = mg_args(e);
    static const bool trace = getenv("BPE_TRACE") != nullptr;
    auto tw0 = std::chrono::steady_clock::now();
    k_loop_prepare<<<1, 32, 0, e->stream>>>(e->d_st.p, (uint32_t)e->n_tokens, e->barrier.p);
    e->stats.kernel_launches++;
    if (use_rounds && !legacy_above) {
      L.sites2 = e->sites2.p;
      L.A.sites_cap = (uint32_t)std::min<size_t>(L.A.sites_cap, e->sites2.cap);
      L.p1_sites = 12;
      L.p2_new = 10;
      L.p2_rw = 2;
      L.bar_ns = 256;
      L.prefetch = 0;
      RoundArgs RA = round_args(e, L, round_k);
      RA.mg_on = 1;
      RA.mg = P.M;
      k_rounds_prepare<<<1, 32, 0, e->stream>>>(e->r_state.p);
      e->stats.kernel_launches++;
      void* rargs[] = {&RA};
      ce = cudaLaunchCooperativeKernel((void*)k_merge_rounds, dim3(e->round_blocks), dim3(RD_THREADS), rargs, 0, e->stream);
    } else {
      // (rounds: the winner has more sites than the packed delta cells can count -- k_merge_loop_mg takes the merges above
      // R_HUGE, it stops, "done", at the first winner at or below it)
      if (use_rounds) L.min_weight = (uint32_t)std::max<int64_t>(mw, (int64_t)R_HUGE + 1);
      void* args[] = {&P};
      ce = cudaLaunchCooperativeKernel((void*)k_merge_loop_mg, dim3(e->mg_loop_blocks), dim3(ML_THREADS), args, 0, e->stream);
    }
    if (ce != cudaSuccess) {
      rc = fail(e, BPE_E_CUDA, "cooperative launch of the sharded mergeUntil kernel: %s", cudaGetErrorString(ce));
      break;
    }
    e->stats.kernel_launches++;
    if ((rc = fetch_state(e)) != BPE_OK) break;
    host_ms[2] += since(tw0);
    e->mg_epoch = e->h_st->mg_epoch;
    e->mg_tie_epoch = e->h_st->mg_tie_epoch;
    uint32_t iters = e->h_st->iters_done;
    if (trace) {
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
      fprintf(stderr, "[bpe r%d] k_merge_loop_mg: %u merges in %.2f ms, status %u, best_cnt %u, hot_n %u thresh %u, keys %u/%u err 0x%x gerr 0x%x\n",
              e->mg_rank, iters, ms, e->h_st->status, e->h_st->best_cnt, e->h_st->hot_n, e->h_st->hot_thresh, e->h_st->n_keys, e->tbl_cap,
              e->h_st->err, e->h_st->g_vals[0]);
    }
    auto tl = clk();
    if (iters) {
      tmp.resize(iters);
      ce = cudaMemcpyAsync(tmp.data(), e->dev_log.p, (size_t)iters * sizeof(MergeRec), cudaMemcpyDeviceToHost, e->stream);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
      if (ce != cudaSuccess) {
        rc = fail(e, BPE_E_CUDA, "merge log read-back: %s", cudaGetErrorString(ce));
        break;
      }
      for (uint32_t i = 0; i < iters; i++) {
        const MergeRec& r = tmp[i];
        log[done].a = r.a;
        log[done].b = r.b;
        log[done].c = r.c;
        log[done].reserved = 0;
        log[done].weight = r.weight;
        done++;
        e->h_len16.push_back(e->h_len16[r.a] + e->h_len16[r.b]);
        e->h_merges.push_back(r.a);
        e->h_merges.push_back(r.b);
        e->h_merges.push_back(r.c);
      }
      e->n_tokens += (int32_t)iters;
      e->mt_dirty = e->lt_dirty = e->dp_dirty = true;
      e->stats.merges_applied += iters;
    }
    host_ms[3] += since(tl);
    uint32_t status = e->h_st->status;
    if (use_rounds && legacy_above && status == LOOP_DONE && (int64_t)e->h_st->best_cnt >= mw) {
      legacy_above = false;  // below R_HUGE now: back to the rounds
      continue;
    }
    if (status == LOOP_NEED_LEGACY) {
      legacy_above = true;
      continue;
    }
    if (status == LOOP_DONE || status == LOOP_EMPTY) break;
    if (status == LOOP_LIMIT) continue;
    if (status == LOOP_NEED_REBUILD) {
      e->hot_valid = false;
      continue;
    }
    if (status == LOOP_ERROR) {
      uint32_t f = e->h_st->err | e->h_st->g_vals[0];
      if (f & ERR_PEER_TIMEOUT) rc = fail(e, BPE_E_INTERNAL, "a peer GPU did not answer within %.0f s (flags 0x%x)", MG_TIMEOUT_NS / 1e9, f);
      else {
        rc = check_dev_err(e);
        if (rc == BPE_OK) rc = fail(e, BPE_E_INTERNAL, "sharded merge loop stopped: local flags 0x%x, all ranks 0x%x", e->h_st->err, e->h_st->g_vals[0]);
      }
      break;
    }
    if (status == LOOP_NEED_HOST) {  // every rank is here with the same winner: each one grows what IT lacks
      auto tg = clk();
      if (e->n_tokens >= BPE_MAX_TOKENS) {
        rc = fail(e, BPE_E_DOMAIN, "token table would exceed %d", BPE_MAX_TOKENS);
        break;
      }
      uint64_t w = e->h_st->best_cnt;
      uint64_t new_keys = std::min<uint64_t>(2 * w + 2, 2 * ((uint64_t)e->n_tokens + 1) + 2);
      uint64_t keys_after = (uint64_t)e->h_st->n_keys + new_keys;
      if (keys_after * 2 > e->tbl_cap) {
        uint64_t want = std::min<uint64_t>(keys_after * 5 / 2, 0x80000000ull);
        want = std::max<uint64_t>(want, std::min<uint64_t>((uint64_t)e->n_slots * e->mg_world / 16, 1ull << 27));  // (the keys are global)
        if (keys_after * 2 > pow2_at_least(want)) {
          rc = fail(e, BPE_E_NOMEM, "pair table cannot grow further");
          break;
        }
        auto tt = clk();
        if ((rc = grow_table(e, pow2_at_least(want))) != BPE_OK) break;
        if (trace) fprintf(stderr, "[bpe r%d] grow_table -> %u slots: %.1f ms\n", e->mg_rank, e->tbl_cap, since(tt));
      }
      uint64_t pool_after = (uint64_t)e->h_st->pool_cursor + 4 * w;  // the in-kernel test is one merge conservative
      // ... one ROUND conservative there: the first merge of a round has at most R_HUGE sites (larger ones take the other kernel),
      // the previous round's batch -- whose allocation the figure in the header does not show yet -- at most as many
      if (use_rounds) pool_after += 2ull * R_HUGE + 2ull * std::max(R_HUGE, R_BATCH_SITES) + 2ull * e->round_blocks * R_POOL_CHUNK;
      auto tp = clk();
      if (pool_after > 0xFFFFFFF0ull) {
        rc = fail(e, BPE_E_DOMAIN, "occurrence pool exceeds 2^32 cells");
        break;
      }
      if ((ce = e->pool.reserve((size_t)pool_after, e->h_st->pool_cursor, e->stream, 1.5)) != cudaSuccess ||
          (ce = e->sites.reserve((size_t)w, 0, e->stream, 1.25)) != cudaSuccess ||
          (use_rounds && (ce = e->sites2.reserve(e->sites.cap, 0, e->stream, 1.0)) != cudaSuccess) ||
          (ce = e->newslots.reserve((size_t)new_keys, 0, e->stream, 1.25)) != cudaSuccess ||
          (ce = e->hot.reserve((size_t)e->h_st->hot_n + 2 * new_keys, e->h_st->hot_n, e->stream, 1.5)) != cudaSuccess ||
          (ce = e->cands.reserve((size_t)e->h_st->best_mult, 0, e->stream, 1.5)) != cudaSuccess) {
        rc = fail(e, BPE_E_NOMEM, "growing merge buffers: %s", cudaGetErrorString(ce));
        break;
      }
      if (trace && since(tp) > 1.0) fprintf(stderr, "[bpe r%d] buffer growth (pool %zu cells): %.1f ms\n", e->mg_rank, e->pool.cap, since(tp));
      if (e->h_st->best_mult > e->mg_tie_cap) {
        rc = fail(e, BPE_E_DOMAIN, "%u pairs tie on (weight, index sum): more than the tie mailbox holds (%u)", e->h_st->best_mult, e->mg_tie_cap);
        break;
      }
      host_ms[1] += since(tg);
      continue;
    }
    rc = fail(e, BPE_E_INTERNAL, "merge loop returned status %u", status);
    break;
  }
  *n_done = done;
  if (rc == BPE_OK) {
    CK(cudaEventRecord(t1, e->stream));
    TRY(fetch_state(e));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    e->stats.ms_last_merge_until = ms;
    e->live_tokens = e->h_st->live_tokens;
    e->stats.sites_merged = (int64_t)e->h_st->sites_total;
    e->stats.tie_breaks = e->h_st->tie_breaks;
    if (use_rounds) {
      RoundState hrs;
      CK(cudaMemcpyAsync(&hrs, e->r_state.p, sizeof(RoundState), cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
      round_stats(e, hrs);
      if (getenv("BPE_TRACE") && e->mg_rank == 0) {
        const unsigned long long* f = e->h_st->fine_ns;
        const double ns = std::max(1.0, (double)f[5]), nb = std::max(1.0, (double)f[11]);
        fprintf(stderr, "[bpe r0] rounds %llu, merges %llu (tried %llu); block 0, us per round whose first merge has <= 16384 sites (%.0f): decide %.1f, P1 %.1f, wait %.1f, P2 %.1f, wait %.1f; "
                        "whose first merge has more (%.0f): decide %.1f, P1 %.1f, wait %.1f, P2 %.1f, wait %.1f; exchange (emit..summed) ms %.1f\n", hrs.rounds, hrs.round_merges, hrs.tried,
                (double)f[5], f[0] / ns * 1e-3, f[1] / ns * 1e-3, f[2] / ns * 1e-3, f[3] / ns * 1e-3, f[4] / ns * 1e-3, (double)f[11], f[6] / nb * 1e-3, f[7] / nb * 1e-3,
                f[8] / nb * 1e-3, f[9] / nb * 1e-3, f[10] / nb * 1e-3, (double)e->h_st->prof_ns[5] * 1e-6);
        const unsigned long long* m = e->h_st->mg_prof_ns;
        fprintf(stderr, "[bpe r0] exchange, block 0, ms: header + records out + report %.1f, wait + sum as they arrive + fold %.1f, barrier %.1f\n", m[5] * 1e-6, m[6] * 1e-6, m[7] * 1e-6);
      }
    }
    if (getenv("BPE_TRACE") && e->mg_rank == 0) {
      fprintf(stderr, "[bpe r0] mg phases ms (decide, P1, wait, M1, wait, send+local P2, barrier+peer wait, apply, wait, P3, wait, tie):");
      for (int i = 0; i < 12; i++) fprintf(stderr, " %.1f", (double)e->h_st->mg_prof_ns[i] * 1e-6);
      fprintf(stderr, "  total %.1f ms, %lld merges; host ms: hot rebuild %.1f, growth %.1f, launch..fetch %.1f, log %.1f\n", ms, (long long)done,
              host_ms[0], host_ms[1], host_ms[2], host_ms[3]);
    }
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return rc;
}

int merge_until_host(bpe_engine* e, int64_t min_weight, int32_t max_length, int64_t max_iterations, bpe_merge* log,
                            int64_t log_cap, int64_t* n_done) {
  if (!e || !n_done || (log_cap > 0 && !log) || log_cap < 0) return BPE_E_INVALID;
  *n_done = 0;
  CK(cudaSetDevice(e->device));
  if (e->n_slots == 0) return BPE_OK;
  TRY(ensure_index(e));
  uint32_t ml = max_length > 0 ? (uint32_t)max_length : 0;
  int64_t mw = min_weight > 0 ? min_weight : 2;
  if (!e->ev0) {
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
  }
  cudaEvent_t t0, t1;
  CK(cudaEventCreate(&t0));
  CK(cudaEventCreate(&t1));
  CK(cudaEventRecord(t0, e->stream));
  int rc = BPE_OK;
  int64_t done = 0;
  // core.ts:374-382: for (iteration = 1; !max_iterations || iteration <= max_iterations; iteration++)
  while ((max_iterations <= 0 || done < max_iterations) && done < log_cap) {
    if (!e->hot_valid || e->hot_max_length != ml) {
      bool any = false;
      if ((rc = rebuild_hot(e, ml, &any)) != BPE_OK) break;
      if (!any) break;  // nothing countable left (core.ts:312)
    }
    if ((rc = run_argmax(e, ml, 1)) != BPE_OK) break;
    if ((e->h_st->err & ERR_HOT_OVERFLOW) || e->h_st->hot_n > e->hot_limit) {
      e->hot_valid = false;
      CK(cudaMemsetAsync(&e->d_st.p->err, 0, sizeof(uint32_t), e->stream));
      continue;
    }
    if (!e->h_st->best_primary || e->h_st->best_cnt < e->hot_thresh) {
      if (e->hot_thresh <= 1 && !e->h_st->best_primary) break;
      e->hot_valid = false;  // the maximum fell below the list's threshold: rebuild lower
      continue;
    }
    uint32_t a = e->h_st->best_a, b = e->h_st->best_b, w = e->h_st->best_cnt;
    if ((int64_t)w < mw) break;  // core.ts:313
    int32_t c = e->n_tokens;
    if (c >= BPE_MAX_TOKENS) {
      rc = fail(e, BPE_E_DOMAIN, "token table would exceed %d", BPE_MAX_TOKENS);
      break;
    }
    if ((rc = run_apply(e, a, b, (uint32_t)c, e->scan_mode ? w : e->h_st->list_len)) != BPE_OK) break;
    log[done].a = (int32_t)a;
    log[done].b = (int32_t)b;
    log[done].c = c;
    log[done].reserved = 0;
    log[done].weight = (int64_t)w;
    e->stats.sites_merged += w;
    done++;
  }
  *n_done = done;
  if (rc != BPE_OK) return rc;
  CK(cudaEventRecord(t1, e->stream));
  TRY(fetch_state(e));
  TRY(check_dev_err(e));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, t0, t1));
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  e->stats.ms_last_merge_until = ms;
  e->live_tokens = e->h_st->live_tokens;
  return BPE_OK;
}


}  // namespace

// =====================================================================================================
extern "C" {

int bpe_abi_version(void) { return BPE_ABI_VERSION; }

int bpe_create(int device, bpe_engine** out) {
  if (!out) return BPE_E_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return BPE_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return BPE_E_CUDA;
  bpe_engine* e = new bpe_engine();
  e->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) e->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete e;
    return BPE_E_CUDA;
  }
  e->stream = e->own_stream;
  if (getenv("BPE_HOST_LOOP")) e->host_loop = 1;
  if (getenv("BPE_ENC_OLD")) e->enc_force_old = 1;
  if (const char* v = getenv("BPE_ENC_DP")) e->enc_dp = atoi(v) != 0;
  if (getenv("BPE_ENC_LMAX")) e->enc_lmax_forced = 1;
  if (const char* v = getenv("BPE_ENC_LMAX")) e->enc_lmax = (atoi(v) == 48 || atoi(v) == 24 || atoi(v) == 20 || atoi(v) == 16) ? atoi(v) : 32;
  *out = e;
  return BPE_OK;
}

void bpe_destroy(bpe_engine* e) {
  if (e && e->mg_mailbox) {
    cudaSetDevice(e->device);
    for (int q = 0; q < e->mg_world; q++)
      if (q != e->mg_rank && e->mg_peer[q]) cudaIpcCloseMemHandle(e->mg_peer[q]);
    cudaFree(e->mg_mailbox);
    e->mg_mailbox = nullptr;
  }
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (e->h_st) cudaFreeHost(e->h_st);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  for (cudaEvent_t ev : {e->ev_in[0], e->ev_in[1], e->ev_in[2], e->ev_done, e->ev_res[0], e->ev_res[1]})
    if (ev) cudaEventDestroy(ev);
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  if (e->h_pipe) cudaFreeHost(e->h_pipe);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
}

const char* bpe_last_error(bpe_engine* e) { return e ? e->err.c_str() : "null engine"; }

int bpe_set_stream(bpe_engine* e, void* cuda_stream) {
  if (!e) return BPE_E_INVALID;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
  return BPE_OK;
}

int bpe_synchronize(bpe_engine* e) {
  if (!e) return BPE_E_INVALID;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  return BPE_OK;
}

int bpe_set_profiling(bpe_engine* e, int enabled) {
  if (!e) return BPE_E_INVALID;
  e->profiling = enabled != 0;
  e->scan_mode = (enabled & 2) ? 1 : 0;  // bit 1: debug full-scan site discovery
  e->host_loop = (enabled & 4) ? 1 : 0;  // bit 2: debug host-driven mergeUntil (one launch per phase)
  return BPE_OK;
}

int bpe_get_stats(bpe_engine* e, bpe_stats* out) {
  if (!e || !out) return BPE_E_INVALID;
  CK(cudaSetDevice(e->device));
  if (e->index_valid && e->d_st.p) {
    TRY(fetch_state(e));
    e->live_tokens = e->h_st->live_tokens;
    e->stats.distinct_pairs = e->h_st->n_keys;
    for (int i = 0; i < 8; i++) e->stats.ms_loop_phase[i] = (double)e->h_st->prof_ns[i] * 1e-6;
#ifdef BPE_FINE_PROF
    fprintf(stderr, "[bpe] phase_sites sub-steps ms (pool, slot+right, left, dec1, new1, right, dec2, new2, record):");
    for (int i = 0; i < 9; i++) fprintf(stderr, " %.1f", (double)e->h_st->fine_ns[i] * 1e-6);
    fprintf(stderr, "\n[bpe] per log2(weight) bucket: merges, P1 us/merge, P2 us/merge, P3 us/merge, P1 ns/site\n");
    for (int b = 0; b < 32; b++) {
      double m = (double)e->h_st->bucket_ns[b][3];
      if (m > 0) fprintf(stderr, "  2^%-2d %8.0f  %8.1f %8.1f %8.1f  %8.2f\n", b, m, e->h_st->bucket_ns[b][0] / m * 1e-3, e->h_st->bucket_ns[b][1] / m * 1e-3,
                         e->h_st->bucket_ns[b][2] / m * 1e-3, e->h_st->bucket_ns[b][0] / m / (1.5 * (double)(1u << b)));
    }
#endif
    e->stats.pool_used = e->h_st->pool_cursor;
  }
  e->stats.corpus_positions = (int64_t)e->n_slots;
  e->stats.corpus_tokens = (int64_t)e->live_tokens;
  *out = e->stats;
  return BPE_OK;
}

int bpe_set_tokens(bpe_engine* e, const int32_t* utf16_len, int32_t n_tokens) {
  if (!e || n_tokens < 0 || (!utf16_len && n_tokens > 0)) return fail(e, BPE_E_INVALID, "bad token table");
  if (n_tokens > BPE_MAX_TOKENS) return fail(e, BPE_E_DOMAIN, "token table of %d exceeds %d", n_tokens, BPE_MAX_TOKENS);
  for (int32_t i = 0; i < n_tokens; i++)
    if (utf16_len[i] < 0) return fail(e, BPE_E_INVALID, "negative token length");
  CK(cudaSetDevice(e->device));
  e->h_len16.assign(utf16_len, utf16_len + n_tokens);
  e->n_tokens = n_tokens;
  e->mt_dirty = e->lt_dirty = e->dp_dirty = true;
  TRY(sync_len16(e));
  e->hot_valid = false;
  return BPE_OK;
}

int bpe_num_tokens(bpe_engine* e, int32_t* n_tokens) {
  if (!e || !n_tokens) return BPE_E_INVALID;
  *n_tokens = e->n_tokens;
  return BPE_OK;
}

int bpe_load_merges(bpe_engine* e, const int32_t* abc, int64_t n_merges) {
  if (!e || n_merges < 0 || (!abc && n_merges > 0)) return fail(e, BPE_E_INVALID, "bad merge list");
  for (int64_t i = 0; i < 3 * n_merges; i++)
    if (abc[i] < 0 || abc[i] >= BPE_MAX_TOKENS) return fail(e, BPE_E_INVALID, "merge %lld holds index %d", (long long)(i / 3), abc[i]);
  e->h_merges.assign(abc, abc + 3 * n_merges);
  e->mt_dirty = e->lt_dirty = e->dp_dirty = true;
  return BPE_OK;
}

int bpe_add_documents_dev(bpe_engine* e, const int32_t* dev_ids, const int64_t* host_doc_offsets, int64_t n_docs) {
  if (!e) return BPE_E_INVALID;
  TRY(check_offsets(e, host_doc_offsets, n_docs));
  if (n_docs == 0) return BPE_OK;
  if (!dev_ids && host_doc_offsets[n_docs] > host_doc_offsets[0]) return fail(e, BPE_E_INVALID, "null ids");
  CK(cudaSetDevice(e->device));
  return append_docs_dev(e, dev_ids, host_doc_offsets, n_docs);
}

int bpe_add_documents(bpe_engine* e, const int32_t* ids, const int64_t* doc_offsets, int64_t n_docs) {
  if (!e) return BPE_E_INVALID;
  TRY(check_offsets(e, doc_offsets, n_docs));
  if (n_docs == 0) return BPE_OK;
  int64_t base = doc_offsets[0], total = doc_offsets[n_docs] - base;
  if (!ids && total > 0) return fail(e, BPE_E_INVALID, "null ids");
  CK(cudaSetDevice(e->device));
  uint64_t new_n = e->n_slots + (uint64_t)total;
  if (new_n >= 0xFFFFFFF0ull) return fail(e, BPE_E_DOMAIN, "corpus exceeds 2^32 positions per engine");
  CK(e->slots.reserve((size_t)new_n + 4, (size_t)e->n_slots, e->stream, 1.5));
  // int32 token ids ARE the slot encoding of single-slot tokens: copy them into place, then validate in place
  if (total > 0)
    CK(cudaMemcpyAsync(e->slots.p + e->n_slots, ids + base, (size_t)total * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  std::vector<int64_t> rel((size_t)n_docs + 1);
  for (int64_t d = 0; d <= n_docs; d++) rel[d] = doc_offsets[d] - base;
  return append_docs_dev(e, reinterpret_cast<const int32_t*>(e->slots.p + e->n_slots), rel.data(), n_docs);
}

int bpe_clear_corpus(bpe_engine* e) {
  if (!e) return BPE_E_INVALID;
  e->n_slots = 0;
  e->doc_off.assign(1, 0);
  e->live_tokens = 0;
  e->index_valid = false;
  e->hot_valid = false;
  return BPE_OK;
}

int bpe_corpus_size(bpe_engine* e, int64_t* n_docs, int64_t* n_tokens) {
  if (!e) return BPE_E_INVALID;
  CK(cudaSetDevice(e->device));
  if (e->index_valid && e->d_st.p) {
    TRY(fetch_state(e));
    e->live_tokens = e->h_st->live_tokens;
  }
  if (n_docs) *n_docs = (int64_t)e->doc_off.size() - 1;
  if (n_tokens) *n_tokens = (int64_t)e->live_tokens;
  return BPE_OK;
}

int bpe_get_corpus(bpe_engine* e, int64_t doc_begin, int64_t doc_end, int32_t* out, int64_t out_cap, int64_t* out_offsets,
                   int64_t* n_out) {
  if (!e || !n_out) return BPE_E_INVALID;
  int64_t nd = (int64_t)e->doc_off.size() - 1;
  if (doc_begin < 0 || doc_end < doc_begin || doc_end > nd) return fail(e, BPE_E_INVALID, "document range [%lld,%lld) outside [0,%lld)", (long long)doc_begin, (long long)doc_end, (long long)nd);
  CK(cudaSetDevice(e->device));
  uint64_t begin = (uint64_t)e->doc_off[doc_begin], end = (uint64_t)e->doc_off[doc_end];
  int64_t n_bounds = doc_end - doc_begin + 1;
  uint64_t span = end - begin;
  uint32_t nblk = (uint32_t)((span + CP_TILE - 1) / CP_TILE);
  DevBuf<uint32_t> bc;
  DevBuf<uint64_t> bo;
  DevBuf<int64_t> dpos, doffs;
  DevBuf<int32_t> dout;
  CK(bc.reserve(nblk + 1));
  CK(bo.reserve(nblk + 2));
  CK(dpos.reserve((size_t)n_bounds));
  CK(doffs.reserve((size_t)n_bounds));
  if (nblk) {
    k_count_ids<<<nblk, CP_THREADS, 0, e->stream>>>(e->slots.p, begin, end, bc.p);
    CKL();
  }
  k_scan_counts<<<1, 1024, 0, e->stream>>>(bc.p, bo.p, nblk);
  CKL();
  uint64_t total = 0;
  CK(cudaMemcpyAsync(&total, bo.p + nblk, sizeof total, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *n_out = (int64_t)total;
  if ((int64_t)total > out_cap || (!out && total) || !out_offsets) {
    if (!out_offsets || (!out && total)) return fail(e, BPE_E_CAPACITY, "output buffers missing; need %llu values", (unsigned long long)total);
    return fail(e, BPE_E_CAPACITY, "output buffer holds %lld values, need %llu", (long long)out_cap, (unsigned long long)total);
  }
  CK(cudaMemcpyAsync(dpos.p, e->doc_off.data() + doc_begin, (size_t)n_bounds * sizeof(int64_t), cudaMemcpyHostToDevice, e->stream));
  k_doc_ranks<<<(int)std::min<int64_t>((n_bounds * 32 + 255) / 256, (int64_t)e->grid(8)), 256, 0, e->stream>>>(e->slots.p, begin, bo.p, dpos.p, n_bounds, doffs.p);
  CKL();
  CK(cudaMemcpyAsync(out_offsets, doffs.p, (size_t)n_bounds * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
  if (total) {
    CK(dout.reserve((size_t)total));
    k_compact_ids<<<nblk, CP_THREADS, 0, e->stream>>>(e->slots.p, begin, end, bo.p, dout.p);
    CKL();
    CK(cudaMemcpyAsync(out, dout.p, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  }
  CK(cudaStreamSynchronize(e->stream));
  return BPE_OK;
}

int bpe_find_next_merge(bpe_engine* e, int64_t min_weight, int32_t max_length, bpe_merge* out, int* found) {
  if (!e || !out || !found) return BPE_E_INVALID;
  *found = 0;
  // a shard alone would answer with ITS best pair, not the corpus's (core.ts:265-310 counts over all documents)
  if (e->mg_world > 1) return fail(e, BPE_E_INVALID, "bpe_find_next_merge is not available on a sharded engine; use bpe_merge_until(max_iterations = 1)");
  CK(cudaSetDevice(e->device));
  if (e->n_slots == 0) return BPE_OK;
  TRY(ensure_index(e));
  uint32_t ml = max_length > 0 ? (uint32_t)max_length : 0;
  int64_t mw = min_weight > 0 ? min_weight : 2;  // core.ts:256
  TRY(run_argmax(e, ml, 0));
  if (!e->h_st->best_primary) return BPE_OK;            // core.ts:312
  if ((int64_t)e->h_st->best_cnt < mw) return BPE_OK;   // core.ts:313
  out->a = (int32_t)e->h_st->best_a;
  out->b = (int32_t)e->h_st->best_b;
  out->c = e->n_tokens;
  out->reserved = 0;
  out->weight = (int64_t)e->h_st->best_cnt;
  *found = 1;
  return BPE_OK;
}

int bpe_apply_merge(bpe_engine* e, int32_t a, int32_t b, int32_t c, int64_t* n_replaced) {
  if (!e) return BPE_E_INVALID;
  if (a < 0 || b < 0 || a >= e->n_tokens || b >= e->n_tokens) return fail(e, BPE_E_INVALID, "merge operands (%d,%d) outside the token table (%d)", a, b, e->n_tokens);
  if (c != e->n_tokens) return fail(e, BPE_E_INVALID, "new token index must be %d, got %d", e->n_tokens, c);
  if (c >= BPE_MAX_TOKENS) return fail(e, BPE_E_DOMAIN, "token table would exceed %d", BPE_MAX_TOKENS);
  CK(cudaSetDevice(e->device));
  if (n_replaced) *n_replaced = 0;
  if (e->n_slots == 0) {  // no corpus: only the vocabulary grows (example/import-merge-log-to-ram.ts:22-31)
    e->h_len16.push_back(e->h_len16[a] + e->h_len16[b]);
    e->h_merges.push_back(a);
    e->h_merges.push_back(b);
    e->h_merges.push_back(c);
    e->mt_dirty = e->lt_dirty = e->dp_dirty = true;
    e->n_tokens++;
    e->index_valid = false;
    return BPE_OK;
  }
  // a single step would update this shard's counts with its local deltas only (the sharded loop exchanges them)
  if (e->mg_world > 1) return fail(e, BPE_E_INVALID, "bpe_apply_merge on a corpus is not available on a sharded engine; use bpe_merge_until");
  TRY(ensure_index(e));
  k_lookup_pair<<<1, 32, 0, e->stream>>>(e->table(), (uint32_t)a, (uint32_t)b, e->d_st.p);
  CKL();
  TRY(fetch_state(e));
  TRY(run_apply(e, (uint32_t)a, (uint32_t)b, (uint32_t)c, e->h_st->list_len));
  TRY(fetch_state(e));
  TRY(check_dev_err(e));
  e->stats.sites_merged += e->h_st->n_sites[0];
  if (n_replaced) *n_replaced = e->h_st->n_sites[0];
  return BPE_OK;
}


// Batched restoreMerge (core.ts:477-494; example/import-merge-log-to-ram.ts:24-31 replays one line per call): applies
// merges (a_i, b_i) -> n_tokens + i in order inside the persistent loop kernel, no arg-max, no host round trip per merge.
int bpe_apply_merges(bpe_engine* e, const int32_t* ab, int64_t n, int64_t* n_replaced) {
  if (!e || n < 0 || (n > 0 && !ab)) return fail(e, BPE_E_INVALID, "bad merge list");
  if (n == 0) return BPE_OK;
  CK(cudaSetDevice(e->device));
  if (e->mg_world > 1) return fail(e, BPE_E_INVALID, "bpe_apply_merges is not available on a sharded engine");
  if ((int64_t)e->n_tokens + n > BPE_MAX_TOKENS) return fail(e, BPE_E_DOMAIN, "token table would exceed %d", BPE_MAX_TOKENS);
  for (int64_t i = 0; i < n; i++)
    if (ab[2 * i] < 0 || ab[2 * i + 1] < 0 || ab[2 * i] >= e->n_tokens + i || ab[2 * i + 1] >= e->n_tokens + i)
      return fail(e, BPE_E_INVALID, "merge %lld refers to a token that does not exist yet", (long long)i);
  if (e->n_slots == 0) {  // no corpus: only the vocabulary / merge list grow (import-merge-log-to-ram.ts:22 empties the corpus first)
    for (int64_t i = 0; i < n; i++) {
      e->h_len16.push_back(e->h_len16[ab[2 * i]] + e->h_len16[ab[2 * i + 1]]);
      e->h_merges.push_back(ab[2 * i]);
      e->h_merges.push_back(ab[2 * i + 1]);
      e->h_merges.push_back(e->n_tokens);
      e->n_tokens++;
      if (n_replaced) n_replaced[i] = 0;
    }
    e->mt_dirty = e->lt_dirty = e->dp_dirty = true;
    TRY(sync_len16(e));
    return BPE_OK;
  }
  DevBuf<int32_t> d_ab;
  CK(d_ab.reserve((size_t)2 * n));
  CK(cudaMemcpyAsync(d_ab.p, ab, (size_t)2 * n * 4, cudaMemcpyHostToDevice, e->stream));
  std::vector<bpe_merge> log((size_t)n);
  int64_t done = 0;
  TRY(merge_until_device(e, 1, 0, n, log.data(), n, &done, d_ab.p));
  if (done != n) return fail(e, BPE_E_INTERNAL, "replayed %lld of %lld merges", (long long)done, (long long)n);
  if (n_replaced)
    for (int64_t i = 0; i < n; i++) n_replaced[i] = log[(size_t)i].weight;
  return BPE_OK;
}

int bpe_merge_until(bpe_engine* e, int64_t min_weight, int32_t max_length, int64_t max_iterations, bpe_merge* log,
                    int64_t log_cap, int64_t* n_done) {
  if (!e || !n_done || (log_cap > 0 && !log) || log_cap < 0) return BPE_E_INVALID;
  if (e->mg_world > 1) return merge_until_mg(e, min_weight, max_length, max_iterations, log, log_cap, n_done);
  if (e->host_loop) return merge_until_host(e, min_weight, max_length, max_iterations, log, log_cap, n_done);
  return merge_until_device(e, min_weight, max_length, max_iterations, log, log_cap, n_done);
}


// ---- sharded training: mailbox set-up and the initial pair-count exchange ------------------------------------------
int bpe_mg_init(bpe_engine* e, int rank, int world, void* handle_out64) {
  if (!e || !handle_out64 || world < 1 || world > MG_MAX_WORLD || rank < 0 || rank >= world) return fail(e, BPE_E_INVALID, "bad rank/world");
  CK(cudaSetDevice(e->device));
  if (e->mg_mailbox) return fail(e, BPE_E_INVALID, "bpe_mg_init called twice");
  e->mg_rank = rank;
  e->mg_world = world;
  // records one rank can send per exchange (k_merge_rounds: the cells of up to 16 merges a round; k_merge_loop_mg: of one merge)
  e->mg_inbox_stride = 1u << 20;
  e->mg_tie_cap = 4096;
  size_t bytes = (size_t)2 * MG_MAX_WORLD * 128 + (size_t)2 * world * e->mg_inbox_stride * 8 + (size_t)2 * world * e->mg_tie_cap * 4;
  CK(cudaMalloc(&e->mg_mailbox, bytes));
  CK(cudaMemset(e->mg_mailbox, 0, bytes));
  e->mg_mailbox_bytes = bytes;
  e->mg_peer[rank] = e->mg_mailbox;
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaIpcGetMemHandle(&h, e->mg_mailbox));
  memcpy(handle_out64, &h, sizeof h);
  e->index_valid = false;  // per-slot side arrays are allocated with the table
  e->mg_connected = (world == 1);
  return BPE_OK;
}

int bpe_mg_connect(bpe_engine* e, const char* handles /* world x 64 bytes, by rank */) {
  if (!e || !handles || !e->mg_mailbox) return fail(e, BPE_E_INVALID, "call bpe_mg_init first");
  CK(cudaSetDevice(e->device));
  for (int q = 0; q < e->mg_world; q++) {
    if (q == e->mg_rank || e->mg_peer[q]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)q * 64, sizeof h);
    CK(cudaIpcOpenMemHandle(&e->mg_peer[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  e->mg_connected = true;
  return BPE_OK;
}

int bpe_mg_state(bpe_engine* e, int* counts_global) {
  if (!e || !counts_global) return BPE_E_INVALID;
  *counts_global = (e->index_valid && e->mg_counts_global) ? 1 : 0;
  return BPE_OK;
}

int bpe_mg_export_counts(bpe_engine* e, uint32_t* dev_keys, uint32_t* dev_counts, int64_t cap, int64_t* n) {
  if (!e || !n || cap < 0 || (cap > 0 && (!dev_keys || !dev_counts))) return fail(e, BPE_E_INVALID, "bad export arguments");
  CK(cudaSetDevice(e->device));
  TRY(ensure_index(e));
  if (e->mg_counts_global) return fail(e, BPE_E_INVALID, "pair counts of this index were already exchanged");
  TRY(fetch_state(e));
  *n = e->h_st->n_keys;
  if (cap == 0) return BPE_OK;  // size query
  CK(e->mg_export_n.reserve(1));
  CK(cudaMemsetAsync(e->mg_export_n.p, 0, 4, e->stream));
  k_mg_export_counts<<<e->grid(4), 256, 0, e->stream>>>(e->table(), dev_keys, dev_counts, (uint32_t)std::min<int64_t>(cap, 0xFFFFFFFFll), e->mg_export_n.p);
  CKL();
  uint32_t got = 0;
  CK(cudaMemcpyAsync(&got, e->mg_export_n.p, 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *n = got;
  if ((int64_t)got > cap) return fail(e, BPE_E_CAPACITY, "export buffers hold %lld pairs, need %u", (long long)cap, got);
  return BPE_OK;
}

int bpe_mg_import_counts(bpe_engine* e, const uint32_t* dev_keys, const uint32_t* dev_counts, int64_t n, int last) {
  if (!e || n < 0 || (n > 0 && (!dev_keys || !dev_counts))) return fail(e, BPE_E_INVALID, "bad import arguments");
  CK(cudaSetDevice(e->device));
  if (!e->index_valid) return fail(e, BPE_E_INVALID, "no pair index to import into");
  TRY(fetch_state(e));
  uint64_t keys_after = (uint64_t)e->h_st->n_keys + (uint64_t)n;
  if (keys_after * 2 > e->tbl_cap) {
    uint64_t want = std::min<uint64_t>(keys_after * 5 / 2, 0x80000000ull);
    if (keys_after * 2 > pow2_at_least(want)) return fail(e, BPE_E_NOMEM, "pair table cannot grow further");
    TRY(grow_table(e, pow2_at_least(want)));
  }
  if (n) {
    k_mg_import_counts<<<e->grid(4), 256, 0, e->stream>>>(e->table(), dev_keys, dev_counts, (uint32_t)n, e->d_st.p);
    CKL();
  }
  TRY(fetch_state(e));
  TRY(check_dev_err(e));
  e->hot_valid = false;
  if (last) e->mg_counts_global = true;
  return BPE_OK;
}

int bpe_pair_counts(bpe_engine* e, int32_t* a, int32_t* b, int64_t* count, int64_t cap, int64_t* n) {
  if (!e || !n || cap < 0) return BPE_E_INVALID;
  *n = 0;
  CK(cudaSetDevice(e->device));
  if (e->n_slots == 0) return BPE_OK;
  TRY(ensure_index(e));
  DevBuf<int32_t> da, db;
  DevBuf<int64_t> dc;
  DevBuf<uint32_t> dn;
  size_t c = (size_t)std::max<int64_t>(cap, 1);
  CK(da.reserve(c));
  CK(db.reserve(c));
  CK(dc.reserve(c));
  CK(dn.reserve(1));
  CK(cudaMemsetAsync(dn.p, 0, 4, e->stream));
  k_dump_pairs<<<e->grid(4), 256, 0, e->stream>>>(e->table(), da.p, db.p, dc.p, (uint32_t)std::min<int64_t>(cap, 0x7FFFFFFF), dn.p);
  CKL();
  uint32_t cnt = 0;
  CK(cudaMemcpyAsync(&cnt, dn.p, 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *n = cnt;
  if ((int64_t)cnt > cap) return fail(e, BPE_E_CAPACITY, "need room for %u pairs", cnt);
  if (cnt) {
    CK(cudaMemcpy(a, da.p, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b, db.p, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(count, dc.p, cnt * sizeof(int64_t), cudaMemcpyDeviceToHost));
  }
  return BPE_OK;
}

int bpe_encode_batch_dev(bpe_engine* e, const int32_t* dev_ids, const int64_t* dev_doc_offsets, int64_t n_docs, int64_t n_ids,
                         int64_t max_doc_len, const int32_t* dev_to_vector_index, int32_t n_tvi, int32_t* dev_out,
                         int64_t* dev_out_offsets, int64_t* dev_first_bad, int64_t* n_out) {
  if (!e || !n_out || n_docs < 0 || n_ids < 0 || !dev_doc_offsets || !dev_out_offsets || (n_ids > 0 && (!dev_ids || !dev_out)))
    return fail(e, BPE_E_INVALID, "bad encode arguments");
  CK(cudaSetDevice(e->device));
  // scratch owned by the engine (freed by bpe_destroy, on the engine's device): reused across calls so that steady-state
  // encode does not allocate
  return encode_dev(e, e->dev_scratch, dev_ids, dev_doc_offsets, n_docs, n_ids, max_doc_len, dev_to_vector_index, n_tvi, dev_out,
                    dev_out_offsets, dev_first_bad, n_out);
}

// host buffers in, host buffers out: stage through grow-only device buffers owned by the engine
int stage_encode_inputs(bpe_engine* e, const int32_t* ids, const int64_t* doc_offsets, int64_t n_docs, int64_t* total_out, int64_t* max_len_out) {
  int64_t base = n_docs ? doc_offsets[0] : 0, total = n_docs ? doc_offsets[n_docs] - base : 0;
  if (total > 0 && !ids) return fail(e, BPE_E_INVALID, "null ids");
  int64_t max_len = 0;
  e->h_rel.resize((size_t)n_docs + 1);
  e->h_rel[0] = 0;
  for (int64_t d = 0; d < n_docs; d++) {
    e->h_rel[d + 1] = doc_offsets[d + 1] - base;
    max_len = std::max(max_len, doc_offsets[d + 1] - doc_offsets[d]);
  }
  CK(e->x_ids.reserve((size_t)std::max<int64_t>(total, 1)));
  CK(e->x_out.reserve((size_t)std::max<int64_t>(total, 1)));
  CK(e->x_off.reserve((size_t)n_docs + 1));
  CK(e->x_ooff.reserve((size_t)n_docs + 1));
  CK(e->x_flag.reserve(1));
  if (total) CK(cudaMemcpyAsync(e->x_ids.p, ids + base, (size_t)total * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->x_off.p, e->h_rel.data(), (size_t)(n_docs + 1) * 8, cudaMemcpyHostToDevice, e->stream));
  if (total) {  // ids must name existing tokens; report the first offender like the reference's left-to-right scan would
    CK(cudaMemsetAsync(e->x_flag.p, 0xFF, 8, e->stream));
    k_first_bad_id<<<e->grid(8), 256, 0, e->stream>>>(e->x_ids.p, (uint64_t)total, (uint32_t)e->n_tokens, e->x_flag.p);
    CKL();
    unsigned long long fb = ~0ull;
    CK(cudaMemcpyAsync(&fb, e->x_flag.p, 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (fb != ~0ull) return fail(e, BPE_E_INVALID, "id %d at %lld outside the token table", ids[base + (int64_t)fb], (long long)fb);
  }
  *total_out = total;
  *max_len_out = max_len;
  return BPE_OK;
}

int encode_host(bpe_engine* e, bool is_text, const int32_t* ids, const uint8_t* utf8, const int64_t* off, int64_t n_docs, const int32_t* to_vector_index,
                int32_t n_tvi, int32_t* out, int64_t out_cap, int64_t* out_offsets, int64_t* first_bad, int64_t* n_out, int64_t* unknown_pos,
                int32_t* unknown_code_point);

int bpe_encode_batch(bpe_engine* e, const int32_t* ids, const int64_t* doc_offsets, int64_t n_docs, const int32_t* to_vector_index,
                     int32_t n_tvi, int32_t* out, int64_t out_cap, int64_t* out_offsets, int64_t* first_bad, int64_t* n_out) {
  if (!e || !n_out || !out_offsets) return fail(e, BPE_E_INVALID, "bad encode arguments");
  if (n_docs < 0 || (!doc_offsets && n_docs > 0)) return fail(e, BPE_E_INVALID, "bad document offsets");
  if (n_docs > 0 && doc_offsets[n_docs] > doc_offsets[0] && !ids) return fail(e, BPE_E_INVALID, "null ids");
  CK(cudaSetDevice(e->device));
  return encode_host(e, false, ids, nullptr, doc_offsets, n_docs, to_vector_index, n_tvi, out, out_cap, out_offsets, first_bad, n_out, nullptr, nullptr);
}


// exclusive scan u32 -> u64 (n + 1 outputs) with the tiled kernels of encode_kernels.cuh
int scan_u32(bpe_engine* e, EncodeScratch& sc, const uint32_t* in, uint64_t* out, uint32_t n) {
  if (n <= 65536) {
    k_scan_counts<<<1, 1024, 0, e->stream>>>(in, out, n);
    CKL();
    return BPE_OK;
  }
  uint32_t n_tiles = (n + SC_TILE - 1) / SC_TILE;
  CK(sc.tile_sums.reserve(n_tiles));
  CK(sc.tile_off.reserve((size_t)n_tiles + 1));
  k_sum_tiles<<<n_tiles, SC_THREADS, 0, e->stream>>>(in, n, sc.tile_sums.p);
  CKL();
  k_scan_counts<<<1, 1024, 0, e->stream>>>(sc.tile_sums.p, sc.tile_off.p, n_tiles);
  CKL();
  k_scan_tiles<<<n_tiles, SC_THREADS, 0, e->stream>>>(in, n, sc.tile_off.p, n_tiles, out);
  CKL();
  return BPE_OK;
}


// ---- text front end: UTF-8 documents -> token indices on the device ----------------------------------------------
int ensure_cpmap(bpe_engine* e) {
  if (e->d_cpmap.p) return BPE_OK;
  CK(e->d_cpmap.reserve(CP_LIMIT));
  CK(e->d_firstpos.reserve(CP_LIMIT));
  CK(cudaMemsetAsync(e->d_cpmap.p, 0xFF, (size_t)CP_LIMIT * 4, e->stream));
  CK(cudaMemsetAsync(e->d_firstpos.p, 0xFF, (size_t)CP_LIMIT * 4, e->stream));
  return BPE_OK;
}

// launches only, no synchronisation: the UTF-8 bytes of whole documents (d_text, boundaries d_byte_off relative to it) become
// ids (placeholders -(cp+1) for unknown code points) and document boundaries in code points (d_char_off, n_docs + 1).
// track_new: keep the first positions of unknown code points (addToCorpus); otherwise d_flags[0] receives
// (pos << 32 | cp) of the first unknown code point or ~0.  d_flags[1] = code points of the longest document,
// d_flags[2] = code points in total.
int text_decode_dev(bpe_engine* e, const uint8_t* d_text, int64_t nbytes, const int64_t* d_byte_off, int64_t n_docs, int32_t* d_ids,
                    int64_t* d_char_off, bool track_new, unsigned long long* d_flags) {
  if ((uint64_t)nbytes >= 0xFFFFFFF0ull) return fail(e, BPE_E_DOMAIN, "text batch too large for one call");
  uint32_t n_tiles = (uint32_t)((nbytes + TX_TILE - 1) / TX_TILE);
  CK(e->x_tilecnt.reserve((size_t)n_tiles + 1));
  CK(e->x_tileoff.reserve((size_t)n_tiles + 2));
  CK(e->x_doccnt.reserve((size_t)n_docs + 1));
  CK(e->x_docoff.reserve((size_t)n_docs + 2));
  CK(cudaMemsetAsync(d_flags, 0xFF, 8, e->stream));
  CK(cudaMemsetAsync(d_flags + 1, 0, 8, e->stream));
  if (n_tiles) {
    k_utf8_count<<<n_tiles, TX_THREADS, 0, e->stream>>>(d_text, (uint64_t)nbytes, e->x_tilecnt.p);
    CKL();
  }
  TRY(scan_u32(e, e->x_scratch, e->x_tilecnt.p, e->x_tileoff.p, n_tiles));
  if (n_tiles) {
    k_utf8_decode<<<n_tiles, TX_THREADS, 0, e->stream>>>(d_text, (uint64_t)nbytes, e->x_tileoff.p, e->d_cpmap.p, d_ids,
                                                         track_new ? e->d_firstpos.p : nullptr, track_new ? nullptr : d_flags);
    CKL();
  }
  // document boundaries in code points: per-document counts, then a scan
  if (n_docs) {
    k_utf8_doc_counts<<<(int)std::min<int64_t>((n_docs + 7) / 8, (int64_t)e->grid(16)), 256, 0, e->stream>>>(d_text, d_byte_off, n_docs, e->x_doccnt.p);
    CKL();
    k_max_u32<<<(int)std::min<int64_t>((n_docs + 255) / 256, (int64_t)e->grid(4)), 256, 0, e->stream>>>(e->x_doccnt.p, n_docs, d_flags + 1);
    CKL();
  }
  TRY(scan_u32(e, e->x_scratch, e->x_doccnt.p, e->x_docoff.p, (uint32_t)n_docs));
  k_u64_to_i64<<<(int)std::min<int64_t>((n_docs + 256) / 256, (int64_t)e->grid(8)), 256, 0, e->stream>>>(e->x_docoff.p, d_char_off, n_docs + 1);
  CKL();
  CK(cudaMemcpyAsync(d_flags + 2, e->x_docoff.p + n_docs, 8, cudaMemcpyDeviceToDevice, e->stream));
  return BPE_OK;
}

// one un-pipelined batch (bpe_add_text): decodes into e->x_ids, document boundaries in code points into e->x_off (device)
// and e->h_rel (host)
int text_to_ids(bpe_engine* e, const uint8_t* utf8, const int64_t* doc_byte_offsets, int64_t n_docs, bool track_new, int64_t* n_chars,
                unsigned long long* unknown) {
  TRY(check_offsets(e, doc_byte_offsets, n_docs));
  TRY(ensure_cpmap(e));
  int64_t base = n_docs ? doc_byte_offsets[0] : 0, nbytes = n_docs ? doc_byte_offsets[n_docs] - base : 0;
  if (nbytes > 0 && !utf8) return fail(e, BPE_E_INVALID, "null text");
  if ((uint64_t)nbytes >= 0xFFFFFFF0ull) return fail(e, BPE_E_DOMAIN, "text batch too large for one call");
  std::vector<int64_t> rel((size_t)n_docs + 1, 0);
  for (int64_t d = 0; d < n_docs; d++) rel[d + 1] = doc_byte_offsets[d + 1] - base;
  CK(e->x_text.reserve((size_t)std::max<int64_t>(nbytes, 1)));
  CK(e->x_ids.reserve((size_t)std::max<int64_t>(nbytes, 1)));  // at most one code point per byte
  CK(e->x_off.reserve((size_t)n_docs + 1));
  CK(e->x_ooff.reserve((size_t)n_docs + 1));
  CK(e->x_flag.reserve(8));
  if (nbytes) CK(cudaMemcpyAsync(e->x_text.p, utf8 + base, (size_t)nbytes, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->x_ooff.p, rel.data(), (size_t)(n_docs + 1) * 8, cudaMemcpyHostToDevice, e->stream));  // byte offsets
  TRY(text_decode_dev(e, e->x_text.p, nbytes, e->x_ooff.p, n_docs, e->x_ids.p, e->x_off.p, track_new, e->x_flag.p));
  e->h_rel.resize((size_t)n_docs + 1);
  CK(cudaMemcpyAsync(e->h_rel.data(), e->x_off.p, (size_t)(n_docs + 1) * 8, cudaMemcpyDeviceToHost, e->stream));
  unsigned long long unk = ~0ull;
  CK(cudaMemcpyAsync(&unk, e->x_flag.p, 8, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *n_chars = n_docs ? e->h_rel[(size_t)n_docs] : 0;
  if (unknown) *unknown = unk;
  return BPE_OK;
}

// ---- host buffers in, host buffers out: chunks of whole documents flow through three streams ----------------------------
// copy in (s_in) | front end + encode (e->stream) | copy out (s_out).  Chunk c+1 is uploaded while chunk c is encoded and the
// vectors of chunk c-1 travel back; every chunk owns its own region of the staging buffers, so no stage waits for a buffer.
// With pinned host buffers the call costs max(PCIe in, encode, PCIe out) instead of their sum; with pageable buffers the
// copies are synchronous and the call degrades to the serial order.
struct PipeChunk {
  int64_t d0 = 0, d1 = 0;    // documents [d0, d1)
  int64_t in0 = 0, in1 = 0;  // input units (ids or bytes) relative to the first document of the batch
  int64_t max_len = 0;       // longest document in input units
  int64_t reg = 0, oreg = 0; // where the chunk lives in the unit staging buffers / the offset staging buffers
};

// the next chunk starting at document d0: about `target` input units, at least one document.  The offsets are checked on the
// way (nothing is read through an offset that was not) and written, relative to the chunk, to `rel` (d1 - d0 + 1 values).
int next_chunk(bpe_engine* e, const int64_t* off, int64_t n_docs, int64_t d0, int64_t target, int64_t* rel, PipeChunk* c) {
  const int64_t first = off[d0], limit = first + target;
  int64_t d = d0, max_len = 0;
  rel[0] = 0;
  while (d < n_docs && (d == d0 || off[d + 1] <= limit)) {
    int64_t len = off[d + 1] - off[d];
    if (len < 0) return fail(e, BPE_E_INVALID, "document offsets must be non-decreasing (doc %lld)", (long long)d);
    max_len = std::max(max_len, len);
    d++;
    rel[d - d0] = off[d] - first;
  }
  c->d0 = d0;
  c->d1 = d;
  c->in0 = first - off[0];
  c->in1 = off[d] - off[0];
  c->max_len = max_len;
  return BPE_OK;
}

// Sizes of the pipeline for a batch of `total` input units in n_docs documents.  The first chunks are small and double up to
// `target` (the first copy in is exposed), the last ones halve again (so are the last encode and the last copy out).  A chunk
// closes when the next document would pass its target, so two consecutive chunks always hold more than the smallest target
// -- which bounds their number and with it the staging buffers (tests/test_abi_symbols.py checks the bound through
// bpe_debug_plan_chunks).
struct PipePlan {
  int64_t target, min_target, max_chunks, pad;
  size_t unit_cap, off_cap;
};

PipePlan pipe_plan(int64_t total, int64_t n_docs, int64_t chunk_units) {
  PipePlan P;
  P.target = std::min<int64_t>(std::max<int64_t>(chunk_units, 1), 1ll << 40);
  P.min_target = std::max<int64_t>(P.target / 8, 1);
  P.max_chunks = std::min<int64_t>(2 * (total / P.min_target) + 8, n_docs + 1);
  P.pad = 256;  // every chunk's region of the unit buffers starts on a multiple of this
  P.unit_cap = (size_t)(total + P.pad * P.max_chunks);
  P.off_cap = (size_t)(n_docs + P.max_chunks + 1);
  return P;
}

// target of chunk number `index` when `remaining` units are left
int64_t pipe_want(const PipePlan& P, int64_t index, int64_t remaining) {
  int64_t want = P.target >> std::max<int64_t>(0, 3 - index);          // ramp up: 1/8, 1/4, 1/2, 1
  want = std::min(want, std::max(remaining / 2, P.min_target));        // ramp down
  return std::max(want, P.min_target);
}

int encode_pipeline(bpe_engine* e, const bool is_text, const int32_t* ids, const uint8_t* utf8, const int64_t* off, int64_t n_docs,
                    const int32_t* to_vector_index, int32_t n_tvi, int32_t* out, int64_t out_cap, int64_t* out_offsets, int64_t* first_bad,
                    int64_t* n_out, int64_t* unknown_pos, int32_t* unknown_code_point) {
  const int64_t base = off[0];
  if (off[n_docs] < base) return fail(e, BPE_E_INVALID, "document offsets must be non-decreasing");
  const int64_t total = off[n_docs] - base;
  const PipePlan P = pipe_plan(total, n_docs, e->enc_chunk);
  const int64_t pad = P.pad;
  const size_t unit_cap = P.unit_cap, off_cap = P.off_cap;
  // everything a chunk in flight may touch is allocated before the first copy starts
  if (is_text) {
    TRY(ensure_cpmap(e));
    CK(e->x_text.reserve(unit_cap));
    CK(e->x_ooff2.reserve(off_cap));
  }
  CK(e->x_ids.reserve(unit_cap));  // text: at most one code point per byte
  CK(e->x_out.reserve(unit_cap));
  CK(e->x_off.reserve(off_cap));
  CK(e->x_ooff.reserve(off_cap));
  CK(e->x_flag.reserve(12));
  CK(e->h_in_off.reserve(off_cap));
  CK(e->h_out_off.reserve((size_t)n_docs + 1));
  if (first_bad) {
    CK(e->x_bad.reserve((size_t)n_docs));
    CK(e->h_bad.reserve((size_t)n_docs));
  }
  const int32_t* d_tvi = nullptr;
  if (to_vector_index && n_tvi > 0) {
    CK(e->x_tvi.reserve((size_t)n_tvi));
    CK(cudaMemcpyAsync(e->x_tvi.p, to_vector_index, (size_t)n_tvi * 4, cudaMemcpyHostToDevice, e->stream));
  }
  if (to_vector_index) d_tvi = e->x_tvi.p;
  int64_t* d_in_off = is_text ? e->x_ooff.p : e->x_off.p;  // the caller's offsets relative to the chunk (text: in bytes)
  int64_t* d_out_off = is_text ? e->x_ooff2.p : e->x_ooff.p;

  int64_t n_up = 0;  // chunks sent so far
  // s_in: the chunk after `prev` goes to the device (units, offsets; ids are checked against the token table there)
  auto upload_next = [&](const PipeChunk& prev, int slot, PipeChunk* c) -> int {
    const int64_t reg = (prev.reg + (prev.in1 - prev.in0) + pad - 1) / pad * pad, oreg = prev.oreg + (prev.d1 - prev.d0) + (prev.d1 ? 1 : 0);
    TRY(next_chunk(e, off, n_docs, prev.d1, pipe_want(P, n_up, total - prev.in1), e->h_in_off.p + oreg, c));
    c->reg = reg;
    c->oreg = oreg;
    const int64_t len = c->in1 - c->in0, nd = c->d1 - c->d0;
    CK(cudaMemcpyAsync(d_in_off + oreg, e->h_in_off.p + oreg, (size_t)(nd + 1) * 8, cudaMemcpyHostToDevice, e->s_in));
    if (len && is_text) CK(cudaMemcpyAsync(e->x_text.p + reg, utf8 + base + c->in0, (size_t)len, cudaMemcpyHostToDevice, e->s_in));
    if (len && !is_text) CK(cudaMemcpyAsync(e->x_ids.p + reg, ids + base + c->in0, (size_t)len * 4, cudaMemcpyHostToDevice, e->s_in));
    if (!is_text) {  // ids must name existing tokens; the first offender is reported like the reference's left-to-right scan would
      CK(cudaMemsetAsync(e->x_flag.p + slot * 4, 0xFF, 8, e->s_in));
      if (len) {
        k_first_bad_id<<<e->grid(8), 256, 0, e->s_in>>>(e->x_ids.p + reg, (uint64_t)len, (uint32_t)e->n_tokens, e->x_flag.p + slot * 4);
        CKL();
      }
      k_copy_u64<<<1, 32, 0, e->s_in>>>(e->h_pipe + slot * 4, e->x_flag.p + slot * 4, 1);
      CKL();
    }
    CK(cudaEventRecord(e->ev_in[slot], e->s_in));
    return BPE_OK;
  };
  // the per-document results of a finished chunk leave the pinned staging for the caller's arrays
  auto deliver = [&](const PipeChunk& c, int res_slot, bool last) -> int {
    CK(cudaEventSynchronize(e->ev_res[res_slot]));
    const int64_t nd = c.d1 - c.d0;
    memcpy(out_offsets + c.d0, e->h_out_off.p + c.d0, (size_t)(nd + (last ? 1 : 0)) * 8);
    if (first_bad) memcpy(first_bad + c.d0, e->h_bad.p + c.d0, (size_t)nd * 8);
    return BPE_OK;
  };

  PipeChunk chunks[3], done;  // chunks[c % 3]: current, next, the one after
  int64_t cum = 0, char_cum = 0;
  bool overflow = false, have_done = false;
  float ms_sum = 0;
  TRY(upload_next(PipeChunk(), 0, &chunks[0]));
  n_up = 1;
  if (chunks[0].d1 < n_docs) {
    TRY(upload_next(chunks[0], 1, &chunks[1]));
    n_up = 2;
  }
  for (int64_t c = 0;; c++) {
    const PipeChunk cur = chunks[c % 3];
    const int slot = (int)(c % 3);
    const int64_t nd = cur.d1 - cur.d0, len = cur.in1 - cur.in0;
    const bool last = cur.d1 >= n_docs;
    int64_t n_units = len, max_len = cur.max_len;
    if (!is_text) {
      CK(cudaEventSynchronize(e->ev_in[slot]));
      unsigned long long fb = *static_cast<volatile unsigned long long*>(e->h_pipe + slot * 4);
      if (fb != ~0ull)
        return fail(e, BPE_E_INVALID, "id %d at %lld outside the token table", ids[base + cur.in0 + (int64_t)fb], (long long)(cur.in0 + (int64_t)fb));
    } else {
      CK(cudaStreamWaitEvent(e->stream, e->ev_in[slot], 0));
      TRY(text_decode_dev(e, e->x_text.p + cur.reg, len, d_in_off + cur.oreg, nd, e->x_ids.p + cur.reg, e->x_off.p + cur.oreg, false,
                          e->x_flag.p + slot * 4));
      k_copy_u64<<<1, 32, 0, e->stream>>>(e->h_pipe + slot * 4, e->x_flag.p + slot * 4, 3);
      CKL();
      CK(cudaStreamSynchronize(e->stream));
      const volatile unsigned long long* hp = e->h_pipe;
      unsigned long long unk = hp[slot * 4];
      if (unk != ~0ull) {  // encodeToCode throws at the first unknown character (core.ts:398-400)
        int64_t pos = char_cum + (int64_t)(unk >> 32);
        if (unknown_pos) *unknown_pos = pos;
        if (unknown_code_point) *unknown_code_point = (int32_t)(unk & 0xFFFFFFFFu);
        return fail(e, BPE_E_INVALID, "unknown token, char: U+%04X at code point %lld", (unsigned)(unk & 0xFFFFFFFFu), (long long)pos);
      }
      max_len = (int64_t)hp[slot * 4 + 1];
      n_units = (int64_t)hp[slot * 4 + 2];
    }
    // host work hidden under the encode kernel: the chunk after the next one starts its way in, the previous one is handed over
    const std::function<int()> overlap = [&]() -> int {
      if (n_up == c + 2 && chunks[(c + 1) % 3].d1 < n_docs) {
        TRY(upload_next(chunks[(c + 1) % 3], (int)((c + 2) % 3), &chunks[(c + 2) % 3]));
        n_up++;
      }
      if (have_done) TRY(deliver(done, (int)((c - 1) & 1), false));
      have_done = false;
      return BPE_OK;
    };
    int64_t k = 0;
    TRY(encode_dev(e, e->x_scratch, e->x_ids.p + cur.reg, e->x_off.p + cur.oreg, nd, n_units, max_len, d_tvi, n_tvi, e->x_out.p + cur.reg,
                   d_out_off + cur.oreg, first_bad ? e->x_bad.p + cur.d0 : nullptr, &k, &overlap));
    ms_sum += e->stats.ms_encode;
    if (cum) {  // the chunk's output offsets take their place in the batch
      k_shift_i64<<<(int)std::min<int64_t>((nd + 256) / 256, (int64_t)e->grid(4)), 256, 0, e->stream>>>(d_out_off + cur.oreg, nd + 1, cum);
      CKL();
    }
    CK(cudaEventRecord(e->ev_done, e->stream));
    CK(cudaStreamWaitEvent(e->s_out, e->ev_done, 0));
    CK(cudaMemcpyAsync(e->h_out_off.p + cur.d0, d_out_off + cur.oreg, (size_t)(nd + (last ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, e->s_out));
    if (first_bad && nd) CK(cudaMemcpyAsync(e->h_bad.p + cur.d0, e->x_bad.p + cur.d0, (size_t)nd * 8, cudaMemcpyDeviceToHost, e->s_out));
    CK(cudaEventRecord(e->ev_res[c & 1], e->s_out));
    if (cum + k > out_cap || (k && !out)) overflow = true;  // keep going: the caller learns the size it needs
    if (!overflow && k) CK(cudaMemcpyAsync(out + cum, e->x_out.p + cur.reg, (size_t)k * 4, cudaMemcpyDeviceToHost, e->s_out));
    cum += k;
    char_cum += n_units;
    done = cur;
    have_done = true;
    if (last) {
      TRY(deliver(done, (int)(c & 1), true));
      break;
    }
  }
  *n_out = cum;
  e->stats.ms_encode = ms_sum;
  if (overflow) return fail(e, BPE_E_CAPACITY, "output buffer holds %lld values, need %lld", (long long)out_cap, (long long)cum);
  return BPE_OK;
}

int encode_host(bpe_engine* e, bool is_text, const int32_t* ids, const uint8_t* utf8, const int64_t* off, int64_t n_docs, const int32_t* to_vector_index,
                int32_t n_tvi, int32_t* out, int64_t out_cap, int64_t* out_offsets, int64_t* first_bad, int64_t* n_out, int64_t* unknown_pos,
                int32_t* unknown_code_point) {
  *n_out = 0;
  if (n_docs == 0) {
    out_offsets[0] = 0;
    return BPE_OK;
  }
  TRY(ensure_pipe(e));
  int rc = encode_pipeline(e, is_text, ids, utf8, off, n_docs, to_vector_index, n_tvi, out, out_cap, out_offsets, first_bad, n_out, unknown_pos, unknown_code_point);
  // whatever happened, nothing may still read or write the caller's buffers when the call returns
  cudaError_t c0 = cudaStreamSynchronize(e->s_in), c1 = cudaStreamSynchronize(e->stream), c2 = cudaStreamSynchronize(e->s_out);
  if (rc == BPE_OK && (c0 != cudaSuccess || c1 != cudaSuccess || c2 != cudaSuccess))
    rc = fail(e, BPE_E_CUDA, "encode pipeline: %s", cudaGetErrorString(c0 != cudaSuccess ? c0 : c1 != cudaSuccess ? c1 : c2));
  return rc;
}

int bpe_decode_batch(bpe_engine* e, const int32_t* values, const int64_t* doc_offsets, int64_t n_docs, const int32_t* from_vector_index,
                     int32_t n_fvi, const uint8_t* token_bytes, const int64_t* token_byte_offsets, int32_t n_tokens, uint8_t* out,
                     int64_t out_cap, int64_t* out_offsets, int64_t* first_bad, int64_t* n_out) {
  if (!e || !n_out || !out_offsets || n_tokens < 0 || (n_tokens > 0 && (!token_byte_offsets))) return fail(e, BPE_E_INVALID, "bad decode arguments");
  TRY(check_offsets(e, doc_offsets, n_docs));
  CK(cudaSetDevice(e->device));
  int64_t base = n_docs ? doc_offsets[0] : 0, total = n_docs ? doc_offsets[n_docs] - base : 0;
  if (total > 0 && !values) return fail(e, BPE_E_INVALID, "null values");
  if ((uint64_t)total >= 0xFFFFFFF0ull) return fail(e, BPE_E_DOMAIN, "batch too large for one decode call");
  const int64_t arena = n_tokens ? token_byte_offsets[n_tokens] : 0;
  if (arena > 0 && !token_bytes) return fail(e, BPE_E_INVALID, "null token bytes");
  e->h_rel.resize((size_t)n_docs + 1);
  e->h_rel[0] = 0;
  for (int64_t d = 0; d < n_docs; d++) e->h_rel[d + 1] = doc_offsets[d + 1] - base;
  DevBuf<uint8_t> d_arena, d_out;
  DevBuf<int64_t> d_tokoff;
  DevBuf<uint32_t> d_lens;
  DevBuf<uint64_t> d_boff;
  DevBuf<unsigned long long> d_bad;
  CK(e->x_ids.reserve((size_t)std::max<int64_t>(total, 1)));
  CK(e->x_off.reserve((size_t)n_docs + 1));
  CK(e->x_ooff.reserve((size_t)n_docs + 1));
  CK(d_arena.reserve((size_t)std::max<int64_t>(arena, 1)));
  CK(d_tokoff.reserve((size_t)n_tokens + 1));
  CK(d_lens.reserve((size_t)std::max<int64_t>(total, 1)));
  CK(d_boff.reserve((size_t)total + 1));
  if (total) CK(cudaMemcpyAsync(e->x_ids.p, values + base, (size_t)total * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->x_off.p, e->h_rel.data(), (size_t)(n_docs + 1) * 8, cudaMemcpyHostToDevice, e->stream));
  if (arena) CK(cudaMemcpyAsync(d_arena.p, token_bytes, (size_t)arena, cudaMemcpyHostToDevice, e->stream));
  if (n_tokens) CK(cudaMemcpyAsync(d_tokoff.p, token_byte_offsets, (size_t)(n_tokens + 1) * 8, cudaMemcpyHostToDevice, e->stream));
  else CK(cudaMemsetAsync(d_tokoff.p, 0, 8, e->stream));
  if (from_vector_index && n_fvi > 0) {
    CK(e->x_tvi.reserve((size_t)n_fvi));
    CK(cudaMemcpyAsync(e->x_tvi.p, from_vector_index, (size_t)n_fvi * 4, cudaMemcpyHostToDevice, e->stream));
  }
  if (first_bad) {
    CK(d_bad.reserve((size_t)std::max<int64_t>(n_docs, 1)));
    CK(cudaMemsetAsync(d_bad.p, 0xFF, (size_t)std::max<int64_t>(n_docs, 1) * 8, e->stream));
  }
  const int32_t* fvi = from_vector_index ? e->x_tvi.p : nullptr;
  int blocks = (int)std::min<int64_t>((total + 255) / 256 + 1, (int64_t)e->grid(8));
  k_decode_lens<<<blocks, 256, 0, e->stream>>>(e->x_ids.p, (uint64_t)total, fvi, from_vector_index ? n_fvi : 0, d_tokoff.p, n_tokens, e->x_off.p, n_docs, d_lens.p,
                                            first_bad ? d_bad.p : nullptr);
  CKL();
  TRY(scan_u32(e, e->x_scratch, d_lens.p, d_boff.p, (uint32_t)total));
  uint64_t nbytes = 0;
  CK(cudaMemcpyAsync(&nbytes, d_boff.p + total, 8, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *n_out = (int64_t)nbytes;
  k_decode_doc_offsets<<<(int)std::min<int64_t>((n_docs + 256) / 256, (int64_t)e->grid(8)), 256, 0, e->stream>>>(e->x_off.p, n_docs, d_boff.p, e->x_ooff.p);
  CKL();
  CK(cudaMemcpyAsync(out_offsets, e->x_ooff.p, (size_t)(n_docs + 1) * 8, cudaMemcpyDeviceToHost, e->stream));
  if (first_bad && n_docs) {
    CK(cudaMemcpyAsync(first_bad, d_bad.p, (size_t)n_docs * 8, cudaMemcpyDeviceToHost, e->stream));  // ~0 (== -1) = no offender
  }
  if ((int64_t)nbytes > out_cap || (nbytes && !out)) {
    CK(cudaStreamSynchronize(e->stream));
    return fail(e, BPE_E_CAPACITY, "output buffer holds %lld bytes, need %llu", (long long)out_cap, (unsigned long long)nbytes);
  }
  if (nbytes) {
    CK(d_out.reserve((size_t)nbytes));
    k_decode_gather<<<blocks, 256, 0, e->stream>>>(e->x_ids.p, (uint64_t)total, fvi, from_vector_index ? n_fvi : 0, d_arena.p, d_tokoff.p, n_tokens, d_boff.p, d_out.p);
    CKL();
    CK(cudaMemcpyAsync(out, d_out.p, (size_t)nbytes, cudaMemcpyDeviceToHost, e->stream));
  }
  CK(cudaStreamSynchronize(e->stream));
  return BPE_OK;
}


// ---- text in, tokens out ---------------------------------------------------------------------------------------------
int bpe_set_chars(bpe_engine* e, const int32_t* code_points, const int32_t* indices, int32_t n) {
  if (!e || n < 0 || (n > 0 && (!code_points || !indices))) return fail(e, BPE_E_INVALID, "bad character table");
  CK(cudaSetDevice(e->device));
  TRY(ensure_cpmap(e));
  CK(cudaMemsetAsync(e->d_cpmap.p, 0xFF, (size_t)CP_LIMIT * 4, e->stream));
  if (n) {
    DevBuf<int32_t> a, b;
    CK(a.reserve((size_t)n));
    CK(b.reserve((size_t)n));
    CK(cudaMemcpyAsync(a.p, code_points, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(b.p, indices, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    k_set_cpmap<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_cpmap.p, a.p, b.p, (uint32_t)n);
    CKL();
    CK(cudaStreamSynchronize(e->stream));
  }
  return BPE_OK;
}

int bpe_add_text(bpe_engine* e, const uint8_t* utf8, const int64_t* doc_byte_offsets, int64_t n_docs, int32_t* new_code_points, int32_t new_cap,
                 int32_t* n_new, int64_t* counts, int64_t counts_cap) {
  if (!e || !n_new) return fail(e, BPE_E_INVALID, "bad arguments");
  *n_new = 0;
  CK(cudaSetDevice(e->device));
  int64_t n_chars = 0;
  const bool trace_t = getenv("BPE_TRACE") != nullptr;
  auto clk = [] { return std::chrono::steady_clock::now(); };
  auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(clk() - t).count(); };
  auto tt0 = clk();
  TRY(text_to_ids(e, utf8, doc_byte_offsets, n_docs, true, &n_chars, nullptr));
  const double ms_ids = since(tt0);
  // unknown code points become tokens in first-appearance order (core.ts:186-199)
  DevBuf<uint32_t> d_cp, d_pos, d_n;
  const uint32_t cap = 1u << 16;
  CK(d_cp.reserve(cap));
  CK(d_pos.reserve(cap));
  CK(d_n.reserve(1));
  CK(cudaMemsetAsync(d_n.p, 0, 4, e->stream));
  k_collect_new_cps<<<e->grid(4), 256, 0, e->stream>>>(e->d_firstpos.p, d_cp.p, d_pos.p, cap, d_n.p);
  CKL();
  uint32_t nn = 0;
  CK(cudaMemcpyAsync(&nn, d_n.p, 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (nn > cap || (int64_t)e->n_tokens + nn > BPE_MAX_TOKENS) return fail(e, BPE_E_DOMAIN, "token table would exceed %d", BPE_MAX_TOKENS);
  if ((int32_t)nn > new_cap) {
    *n_new = (int32_t)nn;
    return fail(e, BPE_E_CAPACITY, "%u new characters, room for %d", nn, new_cap);
  }
  std::vector<uint32_t> cps(nn), pos(nn);
  if (nn) {
    CK(cudaMemcpy(cps.data(), d_cp.p, (size_t)nn * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pos.data(), d_pos.p, (size_t)nn * 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> order(nn);
    for (uint32_t i = 0; i < nn; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return pos[a] < pos[b]; });
    std::vector<int32_t> h_cp(nn), h_idx(nn);
    for (uint32_t k = 0; k < nn; k++) {
      h_cp[k] = (int32_t)cps[order[k]];
      h_idx[k] = e->n_tokens + (int32_t)k;
      new_code_points[k] = h_cp[k];
      e->h_len16.push_back(h_cp[k] >= 0x10000 ? 2 : 1);  // `chars.length` in UTF-16 units (core.ts:272)
    }
    e->n_tokens += (int32_t)nn;
    e->mt_dirty = e->lt_dirty = e->dp_dirty = true;
    TRY(sync_len16(e));
    DevBuf<int32_t> a, b;
    CK(a.reserve(nn));
    CK(b.reserve(nn));
    CK(cudaMemcpyAsync(a.p, h_cp.data(), (size_t)nn * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(b.p, h_idx.data(), (size_t)nn * 4, cudaMemcpyHostToDevice, e->stream));
    k_set_cpmap<<<(nn + 255) / 256, 256, 0, e->stream>>>(e->d_cpmap.p, a.p, b.p, nn);
    CKL();
    CK(cudaStreamSynchronize(e->stream));
  }
  *n_new = (int32_t)nn;
  if (counts_cap < e->n_tokens && counts) return fail(e, BPE_E_CAPACITY, "counts holds %lld entries, need %d", (long long)counts_cap, e->n_tokens);
  CK(e->x_counts.reserve((size_t)std::max(e->n_tokens, 1)));
  CK(cudaMemsetAsync(e->x_counts.p, 0, (size_t)std::max(e->n_tokens, 1) * 8, e->stream));
  if (n_chars) {
    k_fix_and_count<<<e->grid(8), 256, 0, e->stream>>>(e->x_ids.p, (uint64_t)n_chars, e->d_cpmap.p, e->x_counts.p, (uint32_t)e->n_tokens);
    CKL();
  }
  if (counts) CK(cudaMemcpyAsync(counts, e->x_counts.p, (size_t)e->n_tokens * 8, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const double ms_fix = since(tt0) - ms_ids;
  int rc_app = append_docs_dev(e, e->x_ids.p, e->h_rel.data(), n_docs);
  if (trace_t) {
    cudaStreamSynchronize(e->stream);
    fprintf(stderr, "[bpe] add_text: %lld chars: copy in + decode %.1f ms, new characters + counts %.1f ms, append %.1f ms\n", (long long)n_chars, ms_ids, ms_fix,
            since(tt0) - ms_ids - ms_fix);
  }
  return rc_app;
}

int bpe_encode_text_batch(bpe_engine* e, const uint8_t* utf8, const int64_t* doc_byte_offsets, int64_t n_docs, const int32_t* to_vector_index,
                          int32_t n_tvi, int32_t* out, int64_t out_cap, int64_t* out_offsets, int64_t* first_bad, int64_t* n_out,
                          int64_t* unknown_pos, int32_t* unknown_code_point) {
  if (!e || !n_out || !out_offsets) return fail(e, BPE_E_INVALID, "bad encode arguments");
  if (n_docs < 0 || (!doc_byte_offsets && n_docs > 0)) return fail(e, BPE_E_INVALID, "bad document offsets");
  if (unknown_pos) *unknown_pos = -1;
  if (!utf8 && n_docs > 0 && doc_byte_offsets[n_docs] > doc_byte_offsets[0]) return fail(e, BPE_E_INVALID, "null text");
  CK(cudaSetDevice(e->device));
  return encode_host(e, true, nullptr, utf8, doc_byte_offsets, n_docs, to_vector_index, n_tvi, out, out_cap, out_offsets, first_bad, n_out, unknown_pos,
                     unknown_code_point);
}

int bpe_debug_lane_table(const int32_t* abc, int64_t n_merges, int32_t n_tokens, int32_t* a, int32_t* b, int32_t* rank, int32_t* c,
                         int32_t* right_spine_bound, int32_t* left_spine_bound, int64_t cap, int64_t* n) {
  if (n_merges < 0 || (n_merges > 0 && !abc) || !n) return BPE_E_INVALID;
  std::vector<int32_t> merges(abc, abc + 3 * n_merges);
  std::unordered_map<uint32_t, LaneEnt> M;
  std::vector<uint16_t> rule_c;
  TRY(build_lane_entries(nullptr, merges, n_tokens, M, rule_c));
  *n = (int64_t)M.size();
  if ((int64_t)M.size() > cap) return BPE_E_CAPACITY;
  int64_t i = 0;
  for (const auto& kv : M) {
    a[i] = (int32_t)(kv.first >> 16);
    b[i] = (int32_t)(kv.first & 0xFFFFu);
    rank[i] = kv.second.rk == 0xFFFF ? -1 : (int32_t)kv.second.rk;
    c[i] = kv.second.rk == 0xFFFF ? -1 : (int32_t)kv.second.c;
    right_spine_bound[i] = kv.second.rs == 0xFFFF ? -1 : (int32_t)kv.second.rs;
    left_spine_bound[i] = kv.second.ls == 0xFFFF ? -1 : (int32_t)kv.second.ls;
    i++;
  }
  return BPE_OK;
}

int bpe_debug_plan_chunks(const int64_t* doc_offsets, int64_t n_docs, int64_t chunk_units, int64_t* first_doc, int64_t cap, int64_t* n_chunks,
                          int64_t* bounds) {
  if (!doc_offsets || n_docs <= 0 || !n_chunks || !bounds) return BPE_E_INVALID;
  if (doc_offsets[n_docs] < doc_offsets[0]) return BPE_E_INVALID;
  const int64_t total = doc_offsets[n_docs] - doc_offsets[0];
  const PipePlan P = pipe_plan(total, n_docs, chunk_units);
  std::vector<int64_t> rel((size_t)n_docs + 2);
  PipeChunk prev, c;
  int64_t n = 0, reg = 0, oreg = 0;
  while (prev.d1 < n_docs) {
    reg = (prev.reg + (prev.in1 - prev.in0) + P.pad - 1) / P.pad * P.pad;
    oreg = prev.oreg + (prev.d1 - prev.d0) + (prev.d1 ? 1 : 0);
    TRY(next_chunk(nullptr, doc_offsets, n_docs, prev.d1, pipe_want(P, n, total - prev.in1), rel.data(), &c));
    c.reg = reg;
    c.oreg = oreg;
    if (first_doc && n < cap) first_doc[n] = c.d0;
    n++;
    prev = c;
  }
  *n_chunks = n;
  bounds[0] = P.max_chunks;
  bounds[1] = (int64_t)P.unit_cap;
  bounds[2] = prev.reg + (prev.in1 - prev.in0);       // units of the staging buffers the last chunk reaches
  bounds[3] = (int64_t)P.off_cap;
  bounds[4] = prev.oreg + (prev.d1 - prev.d0) + 1;    // offset entries the last chunk reaches
  return BPE_OK;
}

int bpe_restore_documents(bpe_engine* e, const int32_t* ids, const int64_t* doc_offsets, int64_t n_docs) {
  if (!e) return BPE_E_INVALID;
  TRY(check_offsets(e, doc_offsets, n_docs));
  if (n_docs == 0) return BPE_OK;
  CK(cudaSetDevice(e->device));
  int64_t total = 0, max_len = 0;
  TRY(stage_encode_inputs(e, ids, doc_offsets, n_docs, &total, &max_len));
  int64_t n_out = 0;
  TRY(encode_dev(e, e->x_scratch, e->x_ids.p, e->x_off.p, n_docs, total, max_len, nullptr, 0, e->x_out.p, e->x_ooff.p, nullptr, &n_out));
  std::vector<int64_t> ooff((size_t)n_docs + 1);
  CK(cudaMemcpyAsync(ooff.data(), e->x_ooff.p, (size_t)(n_docs + 1) * 8, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return append_docs_dev(e, e->x_out.p, ooff.data(), n_docs);
}

}  // extern "C"
