// Internal declarations shared by the translation units of libdewi_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/dewi_b200.h"

namespace dewi {

// ---- error plumbing (thread-local message behind dewi_last_error) -----------------------------
void set_error(const std::string& msg);
int fail(const std::string& msg);
#define DEWI_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::dewi::fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                          std::to_string(__LINE__) + ")");                                           \
  } while (0)
#define DEWI_TRY(expr)     \
  do {                     \
    int _r = (expr);       \
    if (_r != 0) return _r; \
  } while (0)

constexpr int kQueryBlock = 128;  // queries per MMA M-tile == TMEM lanes
constexpr int kKBlock = 64;       // bf16 elements per 128-byte swizzle row
constexpr int kMaxStages = 8;

// Query j of a 128-query block sits on TMEM lane / partial-list column query_lane(j): consecutive
// queries go to different epilogue warps, so a small batch keeps all four of them busy instead of one.
__host__ __device__ inline int query_lane(int j) { return (j & 3) * 32 + (j >> 2); }

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// SM count of the CURRENT device, looked up once per device (launch geometry: grids are sized in
// multiples of it instead of a hard-coded 148).  Falls back to 148 only if the query itself fails.
int current_sm_count();
// Experiment switches (DEWI_* environment variables): read once per name, then served from a cache.
// Returns `dflt` when the variable is unset.
int env_int(const char* name, int dflt);
bool env_set(const char* name);

// Partial candidate lists: [n_items][kc][128] (score) / (row index); item = chunk * n_qb + qb.
struct Partials {
  float* s = nullptr;
  int* i = nullptr;
  int n_chunks = 0;
  int n_qb = 0;
  int kc = 0;
};

// Per-query admission thresholds for a sweep: values[b * stride + off] (device), or none.  A sweep given
// `max_out` instead runs as the pre-pass: it only records each work item's maximum score per query.
struct SweepSeed {
  const float* values = nullptr;
  int stride = 0, off = 0, n_queries = 0;
  float* max_out = nullptr;
  // pre-pass granularity: 0 = one maximum per work item, [item][128]; 1 = one per 32-row group of every tile,
  // [tile * (N_TILE / 32) + group][n_qb][128] -- many more (still distinct-row) maxima from a small sample, which is
  // what a long candidate list (the certified sweep) needs for a useful seed
  int max_groups = 0;
};

// rerank_scope = "full" (opt-in, NOT the reference's semantics -- SURVEY.md section 0.2): the sweep selects by the
// BLENDED key  w_sim * sim + w_dewi * dewi[row] (+ pref * ent[row])  over the whole corpus instead of by similarity.
// The key is evaluated in the sweep's epilogue from the two payload columns (8 more bytes per row streamed).
struct SweepBlend {
  int enabled = 0;
  float w_sim = 1.f, w_dewi = 0.f, pref = 0.f;
  int use_pref = 0;
  const float* dewi = nullptr;
  const float* ent = nullptr;
};

// ---- tcgen05 sweep (search_tc.cu) ------------------------------------------------------------
// mode 0: S = Q0.E0 ; mode 1: S = (Q0+Q1).E0 ; mode 2: S = Q0.E0 + Q1.E0 + Q0.E1
struct TcPlan {
  int mode;
  int q_rows;    // query rows per MMA (M): 128, or 64 for batches of at most 64 queries
  int n_tile;    // corpus rows per MMA (N): 128 or 256
  int n_stages;  // corpus ring depth
  int q_stages;  // query ring depth (or the number of resident query k-blocks)
  int q_resident;  // the single <= 64-query block stays in shared memory for the whole launch
  int rows_on_m;   // search_tcr.cu: corpus rows on the MMA M dimension, q_rows = queries on N (16 / 32 / 64), natural plane order
  int n_chunks;  // corpus chunks (work items per query block)
  int grid;
  size_t smem_bytes;
};
int tc_supported(int dim, int64_t n_rows);
// swap_queries > 0 (the batch size, at most 64): prefer the rows-on-M sweep (search_tcr.cu) when it applies
int tc_make_plan(int mode, int dim, int64_t n_rows, int n_qb, int kc, int sm_count, TcPlan* plan, int force_chunks = 0,
                 int q_rows = 128, int swap_queries = 0);
int tc_encode_rows_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int box_rows);
int tc_launch(const TcPlan& plan, const CUtensorMap& e0, const CUtensorMap& e1, const CUtensorMap& q0,
              const CUtensorMap& q1, int64_t n_rows, int dim, int n_qb, int kc, float* part_s, int* part_i,
              const SweepSeed& seed, cudaStream_t stream, int fp16_planes = 0, int n_queries = 0,
              const SweepBlend* blend = nullptr);

// ---- tcgen05 sweep with the corpus rows on M, for at most 64 queries (search_tcr.cu) ------------
int tcr_make_plan(int dim, int64_t n_rows, int B, int kc, int sm_count, TcPlan* plan, int force_chunks);
int tcr_launch(const TcPlan& plan, const CUtensorMap& e0, const CUtensorMap& q0, int64_t n_rows, int dim, int B, int kc,
               float* part_s, int* part_i, const SweepSeed& seed, cudaStream_t stream, int fp16_planes,
               const SweepBlend* blend = nullptr);

// ---- tcgen05 sweep on CTA pairs, for more than one query block (search_tc2.cu) -----------------
struct Tc2Plan {
  int mode;
  int n_stages;
  int q_stages;
  int n_chunks;
  int grid;
  int first_items;   // staged sweep: items [0, first_items) run in a first launch whose lists seed the second (0: one launch)
  size_t smem_bytes;
};
int tc2_make_plan(int mode, int dim, int64_t n_rows, int n_qb, int kc, int sm_count, Tc2Plan* plan, int force_chunks = 0,
                  int64_t tiles_per_item = 0, int a_resident = 0, int stage_first = 0);
int tc2_box_rows();
int64_t tc2_sync_words(const Tc2Plan& plan, int64_t n_rows, int n_qb);
int tc2_launch(const Tc2Plan& plan, const CUtensorMap& e0, const CUtensorMap& e1, const CUtensorMap& q0,
               const CUtensorMap& q1, int64_t n_rows, int dim, int n_qb, int kc, float* part_s, int* part_i,
               const SweepSeed& seed, cudaStream_t stream, unsigned int* sync_cnt = nullptr, int fp16_planes = 0, int item0 = 0,
               int item1 = -1);

int tc2_join_launch(int mode, const CUtensorMap& b0, const CUtensorMap& b1, const CUtensorMap& a0, const CUtensorMap& a1,
                    int64_t m_rows, int64_t m_pad, int64_t n_rows, int dim, int sm_count, float tau, int self_join, int64_t a_offset,
                    int sym, unsigned long long* row_best, int* row_count, int64_t* pair_i, int64_t* pair_j, float* pair_sim,
                    int64_t pair_cap, unsigned long long* pair_count, cudaStream_t stream);

// ---- CUDA-core exact sweep (search_simt.cu) --------------------------------------------------
int simt_plan(int64_t n_rows, int B, int sm_count, int* n_chunks);
int simt_launch(const void* rows, int rows_are_bf16, int64_t n_rows, int dim, int space, const float* qn, int B, int kc,
                int n_chunks, float* part_s, int* part_i, cudaStream_t stream, const SweepBlend* blend = nullptr);

// ---- selection / re-rank (select.cu) ---------------------------------------------------------
// seed[b] = the kc-th largest of query b's per-item maxima [n_chunks][n_qb][128] (-inf when n_chunks < kc)
int launch_seed_from_maxima(const float* maxima, int n_chunks, int n_qb, int B, int kc, float* seed, cudaStream_t stream);
// Staged sweep: seed_out[b] = max(seed_in[b] (optional), the kc-th largest entry of query b's partial lists of chunks
// [0, n_chunks_done)) -- kc distinct rows score at least that much, so it bounds the query's kc-th best from below.
constexpr int kSeedWindow = 2048;   // list entries per query the selection ranks in shared memory (select.cu: kWindow)
int launch_seed_from_partials(const float* part_s, int n_chunks_done, int n_qb, int B, int kc, const float* seed_in, float* seed_out,
                              cudaStream_t stream);
int launch_merge_select(const Partials& p, int B, int kc_out, int* cand_idx, float* cand_sim, cudaStream_t stream);
int launch_rescore(const void* rows, int rows_are_bf16, int dim, const float* qn, int* cand_idx, int B, int kc,
                   float* cand_sim, cudaStream_t stream, const float* bar = nullptr);
// any dim / cosine or l2: exact scores of the candidates a blended-key sweep selected (rerank_scope = "full")
int launch_rescore_any(const void* rows, int rows_are_bf16, int dim, int is_l2, const float* qn, const int* cand_idx, int B, int kc,
                       float* cand_sim, cudaStream_t stream);
// Certified single-plane sweep (fp32 corpus swept through its bf16 hi plane): counts in *fails the queries whose list
// of the kc best swept scores cannot be PROVEN to contain the exact top-`need` (see select.cu).
int launch_certificate(const float* cand_sim, const int* cand_idx, int B, int kc, int need, const float* q_stats, int q_planes,
                       const unsigned int* plane_max, int dim, int* fails, float* bar, cudaStream_t stream);
// Fused exchange of the multi-GPU search (one process per GPU, peers' buffers mapped through symmetric memory):
// finalize_local writes this rank's candidate block `[id i64 | sim | dewi | ent]` straight into EVERY rank's gather
// buffer with peer stores over NVLink (slot `my_rank`), and the last block releases `flags[r][my_rank] = seq` on
// every rank; the re-rank kernel of rank r acquires all `world` flags of its own buffer before it reads.
constexpr int kMaxPeers = 16;
struct PeerPush {
  int world = 0;                       // 0: plain local output
  int my_rank = 0;
  unsigned int seq = 0;
  long long block_stride = 0;          // bytes between two ranks' blocks in a gather buffer
  unsigned long long base[kMaxPeers];  // rank r's gather buffer (as mapped in THIS process)
  unsigned long long flags[kMaxPeers]; // rank r's ready flags, uint32[world]
  unsigned int* ticket = nullptr;      // local block counter (self-resetting)
};
int launch_finalize_local(const int* cand_idx, const float* cand_sim, int B, int kc_in, int kcand, int64_t id_base,
                          const float* dewi, const float* ent, float* out_sim, int64_t* out_id, float* out_dewi,
                          float* out_ent, cudaStream_t stream, const PeerPush* push = nullptr, const SweepBlend* blend = nullptr);
// Fused tail (select.cu: tail_kernel): merge -> [certificate] -> exact re-score -> order + payload (+ peer push) -> [blend +
// top-k on a single shard] in ONE launch, one block per query.
struct TailCert {          // certificate of the single-plane sweep
  const float* q_stats;
  int q_planes;
  const unsigned int* plane_max;
  int need;
  int* fails;
};
struct TailRerank {        // single shard: DEWI blend and final top-k in the same launch
  int k;
  float w_sim, w_dewi, pref;
  int use_pref;
  int64_t* out_id;
  float* out_score;
};
int tail_supported(int kc, int kcand, int k);
int launch_tail(const Partials& p, int B, const TailCert* cert, const void* rows, int rows_are_bf16, int dim, const float* qn, int kcand,
                int64_t id_base, const float* dewi, const float* ent, float* out_sim, int64_t* out_id, float* out_dewi, float* out_ent,
                const PeerPush* push, const TailRerank* rr, cudaStream_t stream);
int launch_rerank(const float* sim, const int64_t* id, const float* dewi, const float* ent, int B, int n_shards, int kcand,
                  int64_t shard_stride_bytes, int cand_count, int k, float w_sim, float w_dewi, float pref, int use_pref,
                  int64_t* out_id, float* out_score, cudaStream_t stream, const unsigned int* ready_flags = nullptr,
                  unsigned int seq = 0, unsigned int* status = nullptr, double timeout_s = 120.0);

// ---- operand preparation (prep.cu) -----------------------------------------------------------
// rows fp32 [n, dim] -> optional fp32 copy (normalised), 16-bit hi plane, optional 16-bit lo plane.  The planes are
// bf16 (hi = bf16(x), lo = bf16(x - hi)) or, with fp16_planes, fp16 (the same split in fp16; typed __nv_bfloat16*
// here only as "16-bit storage").  fp16 is used for fp32 corpora: unit-norm rows stay inside its range and its
// 11-bit significand rounds 8x finer than bf16.
// plane_max (optional, device, 2 words): running maxima over all rows ever prepared of ||x - hi|| and ||hi|| as
// float bit patterns -- the corpus side of the certified single-plane sweep's error bound.
int launch_prep_corpus(const float* src, int64_t n, int dim, int normalize, float* dst_f32, __nv_bfloat16* hi,
                       __nv_bfloat16* lo, int* bad_flag, cudaStream_t stream, unsigned int* plane_max = nullptr, int fp16_planes = 0);
// queries fp32 [B, dim] -> qn fp32 [b_pad, dim] (natural order), q hi/lo bf16 [b_pad, dim] (rows >= B
// zeroed).  lane_order 1: plane row of query b is (b / 128) * 128 + query_lane(b % 128) (M = 128 sweeps);
// lane_order 2 (b < 64 only): row (b & 3) * 16 + (b >> 2), the A-row whose M = 64 accumulator lane is query_lane(b).
// q_stats (optional, device, [b_pad][4]): per query { ||q||, ||q - hi||, ||q - hi - lo||, ||hi|| }, rounded upwards.
int launch_prep_queries(const float* q, int B, int b_pad, int dim, int normalize, float* qn, __nv_bfloat16* hi,
                        __nv_bfloat16* lo, cudaStream_t stream, int lane_order = 0, float* q_stats = nullptr, int fp16_planes = 0);
// bf16 -> fp32 (exact widening) of `count` contiguous elements: bulk export of a bf16-storage corpus.
int launch_widen_bf16(const __nv_bfloat16* src, int64_t count, float* dst, cudaStream_t stream);

}  // namespace dewi
