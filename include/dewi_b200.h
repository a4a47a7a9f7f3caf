/*
 * dewi_b200.h -- C ABI of the B200-native DEWI retrieval hot path (libdewi_b200.so).
 *
 * The reference (lexsightllc/DEWI) is pure Python: it has no FFI boundary of its own.  Its plugin
 * boundary is the Python class `BaseIndex` (src/dewi/backends.py:54-163) selected by `DewiIndex`
 * (src/dewi/index.py:44-60).  This header is the native layer a `BaseIndex` subclass binds through
 * ctypes; every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; `dewi_last_error()` then returns a
 *     thread-local message.  No C++ exception crosses this boundary.
 *   - plain pointers and sizes only.  Unless a parameter says "host", pointers are DEVICE pointers
 *     on the handle's device; the caller owns them.  The library owns only the handle, its corpus
 *     planes and its workspace.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls on one
 *     handle must be issued from one thread at a time.
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute entry point
 *     fails with an error.
 */
#ifndef DEWI_B200_H
#define DEWI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEWI_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define DEWI_API __attribute__((visibility("default")))
#else
#define DEWI_API
#endif

typedef struct dewi_index dewi_index_t;

/* space: how ExactIndex scores (backends.py:392,431-436) */
enum { DEWI_SPACE_COSINE = 0, DEWI_SPACE_L2 = 1 };
/* corpus storage */
enum {
  DEWI_DTYPE_FP32 = 0, /* rows kept in fp32 + a bf16 hi/lo plane pair for the tensor-core sweep  */
  DEWI_DTYPE_BF16 = 1  /* rows rounded to bf16 after normalisation (2 B / element)               */
};
/* search flags */
enum {
  DEWI_FLAG_QUERY_NORMALIZED = 1 << 0, /* queries were already normalised as backends.py:420-424  */
  DEWI_FLAG_FORCE_SIMT = 1 << 1,       /* use the CUDA-core exact sweep even when tcgen05 applies  */
  DEWI_FLAG_FORCE_TC = 1 << 2,         /* fail instead of falling back to the CUDA-core sweep      */
  DEWI_FLAG_HOST_IO = 1 << 3,          /* queries / outputs are HOST pointers (copies inside call) */
  DEWI_FLAG_PRECISE_QUERY = 1 << 4,    /* bf16 corpus: hi+lo query planes (2 MMAs) not 1 + rescore */
  DEWI_FLAG_SCOPE_FULL = 1 << 5,       /* (non-reference, opt-in) blend over the WHOLE corpus       */
  DEWI_FLAG_NO_PAIR = 1 << 6,          /* B > 128: keep the 1-CTA sweep instead of the CTA-pair one  */
  DEWI_FLAG_NO_SEED = 1 << 7,          /* skip the sample pre-pass that seeds admission thresholds    */
  DEWI_FLAG_NO_M64 = 1 << 8,           /* B <= 64: keep M = 128 MMAs instead of M = 64                */
  DEWI_FLAG_NO_CERT = 1 << 9,          /* fp32 corpus: always sweep the full hi/lo product (3 MMAs, both planes) instead
                                          of the certified single-plane sweep (hi plane + proof + exact re-score)       */
  DEWI_FLAG_FORCE_CERT = 1 << 10,      /* fp32 corpus: use the certified sweep even on a corpus too small to repay its
                                          host synchronisation (tests)                                                  */
  DEWI_FLAG_NO_ROWS_ON_M = 1 << 11     /* B <= 64: keep the queries on the MMA M dimension instead of the corpus rows    */
};

/* ---- library ------------------------------------------------------------------------------ */
DEWI_API int dewi_abi_version(void);
DEWI_API const char* dewi_last_error(void);
/* Fails unless `device` is compute capability 10.x (B200); reports SM count and memory. */
DEWI_API int dewi_device_check(int device, int* sm_count, size_t* free_bytes, size_t* total_bytes);

/* ---- index lifetime: replaces ExactIndex.__init__ / add / build (backends.py:389-412) ------ */
DEWI_API int dewi_index_create(int dim, int space, int dtype, int device, dewi_index_t** out);
DEWI_API int dewi_index_destroy(dewi_index_t* h);
/* Pre-size the corpus planes for `rows` rows (avoids regrowth while appending). */
DEWI_API int dewi_index_reserve(dewi_index_t* h, int64_t rows);
/* Global id of this shard's row 0 (row-sharded corpus). */
DEWI_API int dewi_index_set_id_base(dewi_index_t* h, int64_t id_base);
/* Append `n` fp32 rows [n, dim] (device pointer, or host pointer when src_is_host).  When
 * `normalized` is 0 and space is cosine each row is divided by its L2 norm in fp32 as
 * backends.py:403-405 does; a zero-norm row is an error (the reference stores a NaN row).  */
DEWI_API int dewi_index_append(dewi_index_t* h, const float* rows, int64_t n, int normalized, int src_is_host, void* stream);
/* Per-row payload columns read by the re-rank: dewi[i] = payload.dewi and ent[i] =
 * float32((ht_mean + hi_mean) * 0.5) (backends.py:457-458).  Writes rows [offset, offset+n). */
DEWI_API int dewi_index_set_payload(dewi_index_t* h, const float* dewi, const float* ent, int64_t offset, int64_t n,
                           int src_is_host, void* stream);
DEWI_API int dewi_index_size(const dewi_index_t* h, int64_t* rows);
/* Copy the stored (normalised) row back as fp32 -- DewiIndex.get_embedding (index.py:101-116). */
DEWI_API int dewi_index_get_row(dewi_index_t* h, int64_t row, float* out_host);
/* Bulk form: rows [row0, row0 + n) as fp32 [n, dim] into `out` (host memory when dst_is_host, else device) --
 * what ExactIndex.save writes as embeddings.npy (backends.py:483-515) and `_embeddings` exposes.  bf16-storage
 * rows are widened on the device (exactly) and copied in >= 64 MB chunks; never one transfer per row.        */
DEWI_API int dewi_index_export_rows(dewi_index_t* h, int64_t row0, int64_t n, float* out, int dst_is_host, void* stream);
/* bf16-storage index only: the stored rows as raw bf16 bit patterns [n, dim] (2 B / element) -- the sharded
 * sidecar CudaIndex.save writes next to the ExactIndex directory -- and the matching ingest, which appends rows
 * that are ALREADY normalised and bf16-rounded without touching them.                                        */
DEWI_API int dewi_index_export_bf16(dewi_index_t* h, int64_t row0, int64_t n, uint16_t* out, int dst_is_host, void* stream);
DEWI_API int dewi_index_append_bf16(dewi_index_t* h, const uint16_t* rows, int64_t n, int src_is_host, void* stream);
/* Read back the two payload columns of rows [offset, offset + n) (see dewi_index_set_payload). */
DEWI_API int dewi_index_get_payload(dewi_index_t* h, int64_t offset, int64_t n, float* dewi_out, float* ent_out,
                           int dst_is_host, void* stream);

/* ---- search: replaces ExactIndex.search (backends.py:414-481) ------------------------------ */
/* Stage 1+2 on one shard: similarity sweep (backends.py:431-436) and candidate selection
 * (backends.py:439-447).  For each of the B queries emits the shard's `kcand` best rows by exact
 * similarity, sorted descending: similarity, GLOBAL id (id_base + row, -1 = empty slot) and the
 * two payload columns of that row.  All outputs are [B, kcand] device arrays.               */
DEWI_API int dewi_index_search_local(dewi_index_t* h, const float* queries, int B, int kcand, int flags, float* out_sim,
                            int64_t* out_id, float* out_dewi, float* out_ent, void* stream);
/* Stage 3: DEWI blend and final select (backends.py:461-481) over the candidates gathered from
 * `n_shards` shards.  Each shard contributed [B, kcand] arrays laid out as stage 1 writes them;
 * shard g's arrays start g * shard_stride_bytes after shard 0's (`sim`, `id`, `dewi`, `ent` point
 * at shard 0's) -- the layout an all-gather of per-rank blocks produces; n_shards = 1 reads
 * stage-1 output in place.  Keeps the `cand_count` (= min(2k, N_total), backends.py:440) best by
 * similarity, computes in fp32
 *     adj = float32(1 - eta) * sim + float32(eta) * dewi  [+ float32(pref) * ent  when pref != 0]
 * with the weak-scalar rounding numpy applies, and writes the k best by adj, sorted descending,
 * into out_id / out_score [B, k].  Does not need an index handle.                              */
DEWI_API int dewi_rerank(const float* sim, const int64_t* id, const float* dewi, const float* ent, int B, int n_shards,
                int kcand, int64_t shard_stride_bytes, int cand_count, int k, double eta, double entropy_pref,
                int64_t* out_id, float* out_score, int device, void* stream);
/* Fused exchange for the row-sharded index (one process per GPU; every rank's gather buffer is mapped into
 * every process, e.g. through torch symmetric memory).  dewi_index_search_local_push is stage 1 + 2 with the
 * all-gather folded in: the finalize kernel writes this rank's block
 *     [ id int64 [B, kcand] | sim f32 [B, kcand] | dewi f32 [B, kcand] | ent f32 [B, kcand] ]
 * into slot `my_rank` (at my_rank * block_stride_bytes) of EVERY rank's gather buffer with peer stores over
 * NVLink, and its last block releases `flags_r[my_rank] = seq` in every rank's flag array (uint32[world]).
 * `peer_bases` / `peer_flags` are HOST arrays of `world` device addresses as mapped in this process.
 * dewi_rerank_gathered is dewi_rerank preceded by an acquire on `ready_flags[0..n_shards) >= seq` (this rank's
 * own flag array): no collective call and no extra launch sits between the sweep and the re-rank.  Callers
 * alternate between two buffers (seq parity) so that a rank one search ahead never overwrites a block a peer
 * is still reading; `seq` must increase by one per search on all ranks.
 * The acquire is bounded in wall time (`timeout_s` seconds, <= 0: DEWI_PUSH_TIMEOUT_S or 120): a block that gives up
 * writes ids of -1 / scores of -inf for its query and stores `seq` into `*status_word` (host-mapped or device
 * memory, may be NULL) instead of trapping, so a slow or dead peer costs one failed search, not the CUDA context. */
DEWI_API int dewi_index_search_local_push(dewi_index_t* h, const float* queries, int B, int kcand, int flags, int world,
                            int my_rank, const uint64_t* peer_bases, const uint64_t* peer_flags, int64_t block_stride_bytes,
                            uint32_t seq, void* stream);
DEWI_API int dewi_rerank_gathered(const float* sim, const int64_t* id, const float* dewi, const float* ent, int B,
                int n_shards, int kcand, int64_t shard_stride_bytes, int cand_count, int k, double eta, double entropy_pref,
                int64_t* out_id, float* out_score, const uint32_t* ready_flags, uint32_t seq, uint32_t* status_word,
                double timeout_s, int device, void* stream);
/* rerank_scope = "full" (DEWI_FLAG_SCOPE_FULL; opt-in, NOT the reference's two-stage semantics of backends.py:439-481,
 * see SURVEY.md section 0.2): the sweep selects by the blended score  (1 - eta) * sim + eta * dewi (+ entropy_pref * ent)
 * over the WHOLE corpus, search_local returns the kcand best rows by that score (still as sim / id / dewi / ent, so the
 * same dewi_rerank finishes the job -- also across shards: the global top-k by the blend is in the union of the local
 * ones).  The weights are per handle: set them before dewi_index_search_local; dewi_index_search sets them itself. */
DEWI_API int dewi_index_set_blend(dewi_index_t* h, double eta, double entropy_pref);
/* Limit: min(kcand, N) <= 400 candidates per query (k <= 200 for dewi_index_search on more than 400 rows); beyond it the
 * calls fail with a message instead of truncating (the reference accepts any k <= N).
 * Whole single-shard search = search_local + rerank.  With DEWI_FLAG_HOST_IO `queries`,
 * `out_id`, `out_score` are host pointers and the call returns after the results have landed. */
DEWI_API int dewi_index_search(dewi_index_t* h, const float* queries, int B, int k, double eta, double entropy_pref,
                      int flags, int64_t* out_id, float* out_score, void* stream);
/* Test aid (no GPU needed): the launch plan the host-side planners choose for a sweep shape, so that CPU tests can
 * walk the shape space and check every plan against the kernels the library instantiates (tests/test_planner_cpu.py).
 * which = 0: single-CTA sweeps (mode 0 = one plane, 1 = two query planes, 2 = hi/lo; q_rows = 64 | 128 query rows on
 *   MMA M; opt = batch size <= 64 to prefer the rows-on-M sweep, else 0) ->
 *   out = { rows_on_m, q_rows, n_tile, n_stages, q_stages, q_resident, n_chunks, grid, smem_bytes, mode }.
 * which = 1: CTA-pair sweep (n_qb even; opt = 1 asks for the staged form) ->
 *   out = { mode, n_stages, q_stages, n_chunks, grid, first_items, smem_bytes, rows_per_tile, 0, 0 }.
 * Non-zero when the planner has no plan for the shape (the search then falls back, see dewi_index_search_local). */
DEWI_API int dewi_plan_probe(int which, int mode, int dim, int64_t n_rows, int n_qb, int kc, int sm_count, int q_rows,
                    int opt, int* out);
/* Kernel launches issued by the last search on this handle (bench.py's `gpu_launches`). */
DEWI_API int dewi_index_last_launches(const dewi_index_t* h, int* launches);
/* fp32 corpus: searches answered by the certified single-plane sweep so far, and how many of those had to be re-run
 * with the full hi/lo product because the certificate could not be given.                                   */
DEWI_API int dewi_index_cert_stats(const dewi_index_t* h, int64_t* used, int64_t* failed);

/* Measurement aid: while enabled, every search brackets its sweep kernel (stage 1's dominant launch)
 * with a CUDA event pair on the caller's stream, kept in a ring of 64 (no synchronisation is added to
 * the search).  sweep_ms waits for the `back`-th most recent bracket (0 = last) and returns its device
 * time and which sweep ran (1 = tcgen05 sweep, 2 = CUDA-core sweep, 3 = tcgen05 CTA-pair sweep).                    */
DEWI_API int dewi_index_set_profiling(dewi_index_t* h, int enable);
DEWI_API int dewi_index_sweep_ms(dewi_index_t* h, int back, float* ms, int* kind);

/* ---- scorer: replaces RobustStats.fit and DewiScorer.score (scorer.py:18-31,49-89) ---------- */
/* Median and MAD of `f` fp32 columns of `n` values each (column c starts at cols + c*ld).
 * Exact order statistics: even n averages the two middle values in fp32 like np.median; MAD is
 * the fp32 median of |v - med|; a zero MAD is returned as 1e-8 (scorer.py:22-25).
 * med_host / mad_host: host arrays of `f` doubles; the call synchronises the stream.           */
DEWI_API int dewi_fit_stats(const float* cols, int64_t n, int f, int64_t ld, double* med_host, double* mad_host, int device,
                   void* stream);
/* Seven signal columns in the order ht_mean, ht_q90, hi_mean, hi_q90, I_hat, redundancy, noise
 * (column c at cols + c*ld elements; floats, or doubles when in_f64 -- the reference scores
 * un-rounded Python floats, which matters when a MAD is tiny).  med7/mad7 (host) as fitted; w6
 * (host) = alpha_t, alpha_i, alpha_m, alpha_r, alpha_n, delta.  float64 arithmetic of
 * scorer.py:28-31,49-89; `out` is n floats, or n doubles when out_f64.                          */
DEWI_API int dewi_score(const void* cols, int in_f64, int64_t n, int64_t ld, const double* med7, const double* mad7,
               const double* w6, int conditional, void* out, int out_f64, int device, void* stream);

/* ---- redundancy: replaces RedundancyEstimator's normalise + matmul (redundancy.py:36-38) ---- */
/* out[m, n] = normalize(a)[m, d] @ normalize(b)[n, d]^T, fp32, eps 1e-12 as torch F.normalize. */
DEWI_API int dewi_similarity_dense(const float* a, int64_t m, const float* b, int64_t n, int d, float* out, int device,
                          void* stream);
/* Thresholded join (this repository's definition): per row of `a` the
 * best similarity against rows of `b`, its index, and the number of rows with sim >= tau; plus up
 * to `pair_cap` (i, j, sim) pairs with sim >= tau appended to pair_i/pair_j/pair_sim, the total
 * number found in *pair_count_host.  self_join: b == a, diagonal excluded, pairs only j > i.
 * a_offset >= 0 (with self_join == 0): `a` is the slice of `b` starting at row a_offset -- one rank's
 * shard of a row-sharded self-join: column a_offset + i is excluded for row i, pairs are emitted only
 * for j > a_offset + i and carry the global row a_offset + i.  a_offset < 0: plain cross join.
 * Large inputs with d % 64 == 0 run on the tensor cores (CTA-pair sweep with a threshold epilogue):
 * rows as bf16 hi+lo planes, three MMAs, similarities good to ~1e-5 (the lo.lo term is dropped); DEWI_JOIN_BF16 keeps one bf16
 * plane (one MMA, similarities carry bf16 rounding ~1e-3).  Otherwise an fp32 CUDA-core kernel.
 * A self-join on the tensor cores multiplies every unordered pair of 256-row blocks ONCE (circulant
 * half of the block grid; an off-diagonal tile updates the statistics of its rows and of its columns);
 * DEWI_JOIN_NO_SYMMETRY evaluates the full M x N product instead (tests / comparison).             */
enum { DEWI_JOIN_BF16 = 1 << 0, DEWI_JOIN_FORCE_SIMT = 1 << 1, DEWI_JOIN_FORCE_TC = 1 << 2, DEWI_JOIN_NO_SYMMETRY = 1 << 3 };
/* Scratch memory (normalised operand planes, packed row statistics): `workspace` is a device buffer of at
 * least dewi_join_workspace_bytes(...) bytes owned by the caller -- the Python wrapper passes a torch tensor,
 * so repeated joins cost no cudaMalloc / cudaFree; NULL makes the call allocate and free its own.
 * For dewi_self_join_range use dewi_join_workspace_bytes(n, n, d, 1, flags).                          */
DEWI_API int64_t dewi_join_workspace_bytes(int64_t m, int64_t n, int d, int self_join, int flags);
DEWI_API int dewi_join(const float* a, int64_t m, const float* b, int64_t n, int d, float tau, int self_join,
              int64_t a_offset, int flags, float* row_max, int64_t* row_argmax, int32_t* row_count, int64_t* pair_i, int64_t* pair_j, float* pair_sim,
              int64_t pair_cap, int64_t* pair_count_host, void* workspace, int64_t workspace_bytes, int device, void* stream);

/* One rank's share of a row-sharded SYMMETRIC self-join of x[n, d] (every rank holds all rows): the row
 * blocks of [row_lo, row_hi) (multiples of 256, or ending at n) against their half of the block grid --
 * block I meets blocks I .. I + (T-1)/2 (mod T), T = ceil(n / 256), plus I + T/2 for I < T/2 when T is
 * even -- so contiguous ranges of equal length carry equal work and the ranges of all ranks cover every
 * unordered row pair exactly once.  row_max / row_argmax / row_count have n entries and hold the
 * contribution of this range only (-inf / -1 / 0 where it has none): combine ranks with max (ties: any
 * argmax) and sum.  Each pair with sim >= tau is emitted once, as (min, max).  Tensor cores only
 * (d % 64 == 0); flags: DEWI_JOIN_BF16.  [row_lo, row_hi) = [0, n) is the whole self-join.            */
DEWI_API int dewi_self_join_range(const float* x, int64_t n, int d, float tau, int64_t row_lo, int64_t row_hi, int flags,
              float* row_max, int64_t* row_argmax, int32_t* row_count, int64_t* pair_i, int64_t* pair_j, float* pair_sim,
              int64_t pair_cap, int64_t* pair_count_host, void* workspace, int64_t workspace_bytes, int device, void* stream);

/* ---- neighbours of the path that reuse its kernels -------------------------- */
/* local_weights_from_surprisal (src/dewi/local_weights.py:5-26): float32 median / MAD (+1e-8), z-score,
 * clip to +-5, softplus.  `s` and `out` are n floats on the device; the call synchronises the stream.  */
DEWI_API int dewi_local_weights(const float* s, int64_t n, float* out, int device, void* stream);
/* Connected components of a pair list (e.g. the join's near-duplicate pairs) over documents 0..n-1:
 * labels[i] = smallest document index of i's cluster.  The clusters are what metrics.duplicate_rate /
 * cluster_coverage consume (src/dewi/metrics.py:173-212).  Synchronises the stream.                    */
DEWI_API int dewi_cluster_pairs(const int64_t* pair_i, const int64_t* pair_j, int64_t n_pairs, int64_t n, int32_t* labels,
                       int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEWI_B200_H */
