// K2 -- candidate merge, exact re-score, DEWI blend and final selection.
//
// Everything after the sweep in ExactIndex.search (reference src/dewi/backends.py:439-481):
//   merge_select   : per-CTA partial lists -> the query's best `kc` rows by (approximate) similarity
//   rescore        : exact fp32 similarity of those rows against the fp32 query (tensor path only)
//   finalize_local : sort, keep the shard's top min(2k, n) (:439-447), attach global id + payload columns
//   rerank         : top-2k of the gathered shards' candidates, `adj = (1-eta)*sim + eta*dewi
//                    (+ entropy_pref*entropy)` in fp32 with numpy's rounding (:461-465), top-k, sort (:468-471)
// One thread block per query; candidate counts are tiny next to the sweep, so these are latency-, not
// bandwidth-bound kernels.
#include <algorithm>

#include "internal.h"

namespace dewi {
namespace {

constexpr int kSelThreads = 128;

__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// Descending sort key: score first, then lower id first.
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t id_lo) {
  return (static_cast<unsigned long long>(orderable(score)) << 32) | static_cast<unsigned long long>(~id_lo);
}

// In-place descending bitonic sort of p (power of two) (key, val) pairs in shared memory.
__device__ void bitonic_sort_desc(unsigned long long* key, int* val, int p) {
  for (int size = 2; size <= p; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (p >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = key[lo], b = key[hi];
        if ((a < b) == desc) {
          key[lo] = b; key[hi] = a;
          const int va = val[lo]; val[lo] = val[hi]; val[hi] = va;
        }
      }
    }
  }
  __syncthreads();
}

__host__ __device__ inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- merge_select -------------------------------------------------------------------------------
// Streams the query's n_chunks*kc partial candidates through a shared-memory window that holds the
// running best kc_out (sorted) plus newly admitted candidates.  Only candidates that beat the current
// kc_out-th best are admitted (block-wide ballot compaction), so after the first window fills and is
// sorted, later candidates trickle in and the number of bitonic sorts stays at one or two -- and a
// sweep that ran with seeded thresholds, whose partial lists are mostly empty, needs a single small one.
constexpr int kWindow = 2048;
static_assert(kWindow == kSeedWindow, "tc2_make_plan sizes the staged sweep's first round by this window");
constexpr int kMergeThreads = 1024;  // one compare-exchange per thread per bitonic stage
static_assert(kWindow <= 2 * kMergeThreads, "the seed selections hold the window two keys per thread");

__global__ void __launch_bounds__(kMergeThreads)
merge_select_kernel(const float* __restrict__ part_s, const int* __restrict__ part_i, int n_chunks, int n_qb, int kc,
                    int kc_out, int* __restrict__ cand_idx, float* __restrict__ cand_sim) {
  __shared__ unsigned long long key[kWindow];
  __shared__ int val[kWindow];
  __shared__ int warp_base[kMergeThreads / 32];
  __shared__ int fill_s;
  const int b = blockIdx.x;
  const int qb = b / kQueryBlock, ql = query_lane(b % kQueryBlock);
  const int total = n_chunks * kc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < kWindow; t += blockDim.x) { key[t] = 0ull; val[t] = -1; }
  if (threadIdx.x == 0) fill_s = kc_out;
  __syncthreads();
  unsigned long long floor_key = 0ull;  // admission bar: the kc_out-th best key so far (0 = any valid key)
  for (int base = 0; base < total; base += blockDim.x) {
    const int e = base + threadIdx.x;
    unsigned long long kk = 0ull;
    int vv = -1;
    if (e < total) {
      const int chunk = e / kc, k = e - chunk * kc;
      const size_t off = ((static_cast<size_t>(chunk) * n_qb + qb) * kc + k) * kQueryBlock + ql;
      const int idx = part_i[off];
      if (idx >= 0) { kk = make_key(part_s[off], static_cast<uint32_t>(idx)); vv = idx; }
    }
    const bool pass = kk > floor_key;
    const unsigned int m = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) warp_base[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 32 warp counts
      const int c = warp_base[lane];
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      warp_base[lane] = fill_s + incl - c;
      if (lane == 31) fill_s += incl;
    }
    __syncthreads();
    if (pass) {
      const int slot = warp_base[warp] + __popc(m & ((1u << lane) - 1u));
      key[slot] = kk;   // slot < kWindow: the window is sorted down whenever fewer than blockDim slots remain
      val[slot] = vv;
    }
    __syncthreads();
    const int fill = fill_s;
    if (fill > kWindow - static_cast<int>(blockDim.x) || base + static_cast<int>(blockDim.x) >= total) {
      const int p = next_pow2(fill);
      bitonic_sort_desc(key, val, p);  // slots [fill, p) hold zeros
      floor_key = key[kc_out - 1];
      __syncthreads();
      for (int t = kc_out + threadIdx.x; t < p; t += blockDim.x) { key[t] = 0ull; val[t] = -1; }
      if (threadIdx.x == 0) fill_s = kc_out;
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < kc_out; t += blockDim.x) {
    const int idx = val[t];
    cand_idx[static_cast<size_t>(b) * kc_out + t] = idx;
    float s = -INFINITY;
    if (idx >= 0) {
      const uint32_t o = static_cast<uint32_t>(key[t] >> 32);
      s = __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
    }
    cand_sim[static_cast<size_t>(b) * kc_out + t] = s;
  }
}

// ---- seed_from_maxima -----------------------------------------------------------------------------
// Pre-pass result -> admission thresholds.  Every work item of the pre-pass reported the best score it
// saw for the query; those maxima belong to distinct corpus rows, so the kc-th largest of them is a
// lower bound of the query's kc-th best score over the corpus.
//
// The kc-th largest of up to 2 * blockDim values (each thread holds two order-preserving 32-bit keys, 0 = absent) by
// bisection on the key: 32 rounds, each one or two barrier-popcounts (__syncthreads_count), no shared memory and no
// n^2 work.  (Rank counting -- every key against every other -- took 684 us for 4096 queries x 1024 maxima, 14 % of
// the C2 B = 4096 search: profiles/r2_zy_launches_c2_b4096_1Mrows.csv.)  Returns 0 when fewer than kc keys are present.
__device__ __forceinline__ uint32_t kth_largest_key(uint32_t k0, uint32_t k1, bool two, int kc) {
  uint32_t prefix = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = prefix | (1u << bit);
    int c = __syncthreads_count(k0 >= cand);
    if (two) c += __syncthreads_count(k1 >= cand);   // (`two` is uniform over the block)
    if (c >= kc) prefix = cand;                      // at least kc keys are >= cand: the kc-th largest is too
  }
  return prefix;
}
__device__ __forceinline__ float key_to_float(uint32_t o) {
  return o ? __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o) : -INFINITY;   // 0: fewer than kc values -> no seed
}

__global__ void __launch_bounds__(kMergeThreads)
seed_from_maxima_kernel(const float* __restrict__ maxima, int n_chunks, int n_qb, int kc, float* __restrict__ seed) {
  const int b = blockIdx.x;
  const int qb = b / kQueryBlock, ql = query_lane(b % kQueryBlock);
  uint32_t k[2] = {0u, 0u};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int t = static_cast<int>(threadIdx.x) + s * static_cast<int>(blockDim.x);
    if (t < n_chunks) {
      const float v = maxima[(static_cast<size_t>(t) * n_qb + qb) * kQueryBlock + ql];
      if (v > -INFINITY) k[s] = orderable(v);   // (every finite score maps to a non-zero key; NaN is dropped)
    }
  }
  const uint32_t kth = kth_largest_key(k[0], k[1], n_chunks > static_cast<int>(blockDim.x), kc);
  if (threadIdx.x == 0) seed[b] = key_to_float(kth);
}

// The same selection by rank counting (the key beaten by exactly kc - 1 others): kept for A/B runs (DEWI_SEED_RANKCOUNT=1).
__global__ void __launch_bounds__(kMergeThreads)
seed_from_maxima_rank_kernel(const float* __restrict__ maxima, int n_chunks, int n_qb, int kc, float* __restrict__ seed) {
  __shared__ unsigned long long key[kWindow];
  const int b = blockIdx.x;
  const int qb = b / kQueryBlock, ql = query_lane(b % kQueryBlock);
  for (int t = threadIdx.x; t < n_chunks; t += blockDim.x) {
    const float v = maxima[(static_cast<size_t>(t) * n_qb + qb) * kQueryBlock + ql];
    key[t] = (v > -INFINITY) ? make_key(v, static_cast<uint32_t>(t)) : 0ull;
  }
  if (threadIdx.x == 0) seed[b] = -INFINITY;   // fewer than kc finite maxima: no seed
  __syncthreads();
  for (int e = threadIdx.x; e < n_chunks; e += blockDim.x) {
    const unsigned long long x = key[e];
    if (x == 0ull) continue;
    int rank = 0;
    for (int f = 0; f < n_chunks; ++f) rank += (key[f] > x) ? 1 : 0;
    if (rank == kc - 1) {
      const uint32_t o = static_cast<uint32_t>(x >> 32);
      seed[b] = __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
    }
  }
}

// Staged sweep (api.cu): the same selection over the finished partial lists [chunk][n_qb][kc][128] of the first chunks.
__global__ void __launch_bounds__(kMergeThreads)
seed_from_partials_kernel(const float* __restrict__ part_s, int n_vals, int n_qb, int kc, const float* __restrict__ seed_in,
                          float* __restrict__ seed_out) {
  const int b = blockIdx.x;
  const int qb = b / kQueryBlock, ql = query_lane(b % kQueryBlock);
  uint32_t k[2] = {0u, 0u};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int t = static_cast<int>(threadIdx.x) + s * static_cast<int>(blockDim.x);
    if (t < n_vals) {
      const int chunk = t / kc, e = t - chunk * kc;
      const float v = part_s[((static_cast<size_t>(chunk) * n_qb + qb) * kc + e) * kQueryBlock + ql];
      if (v > -INFINITY) k[s] = orderable(v);
    }
  }
  const float found = key_to_float(kth_largest_key(k[0], k[1], n_vals > static_cast<int>(blockDim.x), kc));
  if (threadIdx.x == 0) {
    const float prev = seed_in ? seed_in[b] : -INFINITY;
    seed_out[b] = (prev > found) ? prev : found;   // (a NaN or missing earlier seed compares false: `found` stands)
  }
}

// (rank-counting form, DEWI_SEED_RANKCOUNT=1)
__global__ void __launch_bounds__(512)
seed_from_partials_rank_kernel(const float* __restrict__ part_s, int n_vals, int n_qb, int kc, const float* __restrict__ seed_in,
                               float* __restrict__ seed_out) {
  __shared__ unsigned long long key[kWindow];
  __shared__ float found;
  const int b = blockIdx.x;
  const int qb = b / kQueryBlock, ql = query_lane(b % kQueryBlock);
  for (int t = threadIdx.x; t < n_vals; t += blockDim.x) {
    const int chunk = t / kc, k = t - chunk * kc;
    const float v = part_s[((static_cast<size_t>(chunk) * n_qb + qb) * kc + k) * kQueryBlock + ql];
    key[t] = (v > -INFINITY) ? make_key(v, static_cast<uint32_t>(t)) : 0ull;
  }
  if (threadIdx.x == 0) found = -INFINITY;
  __syncthreads();
  for (int e = threadIdx.x; e < n_vals; e += blockDim.x) {
    const unsigned long long x = key[e];
    if (x == 0ull) continue;
    int rank = 0;
    for (int f = 0; f < n_vals; ++f) rank += (key[f] > x) ? 1 : 0;
    if (rank == kc - 1) {
      const uint32_t o = static_cast<uint32_t>(x >> 32);
      found = __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float prev = seed_in ? seed_in[b] : -INFINITY;
    seed_out[b] = (prev > found) ? prev : found;
  }
}

// ---- certificate of the single-plane sweep ----------------------------------------------------------------
// fp32 corpus, swept through its bf16 hi plane only (api.cu: certified mode).  With s = q . e the exact similarity
// and S the swept one (bf16 products are exact, fp32 accumulation),
//     |s - S| <= ||q|| * max_r ||e_r - hi_r|| + ||q - q_planes|| * max_r ||hi_r|| + gamma * ||q_planes|| * max_r ||hi_r|| =: eps
// (Cauchy-Schwarz on the two rounding residuals; gamma = dim * 2^-22 bounds the fp32 accumulation of `dim` exact
// products, with a factor two of slack for a truncating adder).  Let S_(j) be the j-th largest swept score.  At least
// `need` rows have s >= S_(need) - eps, so every row of the exact top-`need` has s >= S_(need) - eps and therefore
// S >= S_(need) - 2 eps: the list of the kc best swept scores contains the exact top-`need` whenever it is not full or
// S_(kc) < S_(need) - 2 eps.  Queries for which that cannot be shown are counted in *fails (the caller re-runs the
// batch with the full hi/lo product).  cand_sim is sorted descending per query, cand_idx < 0 marks empty slots.
__global__ void certificate_kernel(const float* __restrict__ cand_sim, const int* __restrict__ cand_idx, int B, int kc, int need,
                                   const float* __restrict__ q_stats, int q_planes, const unsigned int* __restrict__ plane_max,
                                   float gamma, int* __restrict__ fails, float* __restrict__ bar_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* cs = cand_sim + static_cast<size_t>(b) * kc;
  const int* ci = cand_idx + static_cast<size_t>(b) * kc;
  const float row_err = __uint_as_float(plane_max[0]), row_hi = __uint_as_float(plane_max[1]);
  const float qn = q_stats[b * 4 + 0], qe = q_stats[b * 4 + (q_planes == 1 ? 1 : 2)];
  const float qp = qn + qe;                         // >= ||q_planes||
  const float eps = __fmul_ru(qn, row_err) + __fmul_ru(qe, row_hi) + __fmul_ru(gamma, __fmul_ru(qp, row_hi));
  const float bar = __fsub_rd(cs[need - 1], __fmul_ru(2.f, eps));
  bar_out[b] = bar;                                 // list entries below the bar cannot be in the exact top-`need`
  if (ci[kc - 1] < 0) return;                       // fewer than kc rows in play: the list holds them all
  if (!(cs[kc - 1] < bar)) atomicAdd(fails, 1);     // (a NaN score fails the certificate too)
}

// ---- rescore ------------------------------------------------------------------------------------
// `bar` (optional, per query): candidates whose swept score lies below it are dropped instead of re-scored (certified
// sweep: they provably cannot belong to the exact top-2k).
template <typename RowT>
__global__ void rescore_kernel(const RowT* __restrict__ rows, int dim, const float* __restrict__ qn,
                               int* __restrict__ cand_idx, int total, int kc, float* __restrict__ cand_sim,
                               const float* __restrict__ bar) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= total) return;
  const int idx = cand_idx[w];
  const int b = w / kc;
  if (idx < 0 || (bar && cand_sim[w] < bar[b])) {
    __syncwarp();
    if (lane == 0) { cand_sim[w] = -INFINITY; cand_idx[w] = -1; }
    return;
  }
  const RowT* r = rows + static_cast<size_t>(idx) * dim;
  const float* q = qn + static_cast<size_t>(b) * dim;
  float acc = 0.f;
  for (int d = lane * 8; d < dim; d += 256) {
    float x[8];
    if (sizeof(RowT) == 2) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(r + d));
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        x[2 * i] = __uint_as_float(u[i] << 16);
        x[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
      }
    } else {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(r + d));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(r + d) + 1);
      x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
      x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
    }
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(q + d));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(q + d) + 1);
    acc = fmaf(x[0], q0.x, acc); acc = fmaf(x[1], q0.y, acc); acc = fmaf(x[2], q0.z, acc); acc = fmaf(x[3], q0.w, acc);
    acc = fmaf(x[4], q1.x, acc); acc = fmaf(x[5], q1.y, acc); acc = fmaf(x[6], q1.z, acc); acc = fmaf(x[7], q1.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) cand_sim[w] = acc;
}

// General form (any dim, cosine or l2 space; scalar loads): re-derives the exact scores of candidates that the
// CUDA-core sweep selected by a blended key (rerank_scope = "full"), in the sweep's own arithmetic (fmaf chains over
// d = lane, lane + 32, ... are NOT reproduced -- the re-rank only needs the exact similarity to within rounding).
template <typename RowT>
__global__ void rescore_any_kernel(const RowT* __restrict__ rows, int dim, const float* __restrict__ qn, const int* __restrict__ cand_idx,
                                   int total, int kc, float* __restrict__ cand_sim, int is_l2) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= total) return;
  const int idx = cand_idx[w];
  if (idx < 0) {
    if (lane == 0) cand_sim[w] = -INFINITY;
    return;
  }
  const RowT* r = rows + static_cast<size_t>(idx) * dim;
  const float* q = qn + static_cast<size_t>(w / kc) * dim;
  float acc = 0.f;
  for (int d = lane; d < dim; d += 32) {
    const float x = sizeof(RowT) == 2 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(r)[d]) : static_cast<float>(r[d]);
    if (is_l2) {
      const float t = x - q[d];
      acc = fmaf(t, t, acc);
    } else {
      acc = fmaf(x, q[d], acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) cand_sim[w] = is_l2 ? -acc : acc;
}

// ---- finalize_local -----------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// PUSH = 1: the block is not written to out_* but to slot `my_rank` of every rank's gather buffer (peer stores over
// NVLink, fire-and-forget), then the last block to finish publishes the ready flags -- finalize and all-gather in one kernel.
template <int PUSH>
__global__ void __launch_bounds__(kSelThreads)
finalize_local_kernel(const int* __restrict__ cand_idx, const float* __restrict__ cand_sim, int kc_in, int kcand,
                      long long id_base, const float* __restrict__ dewi, const float* __restrict__ ent,
                      float* __restrict__ out_sim, long long* __restrict__ out_id, float* __restrict__ out_dewi,
                      float* __restrict__ out_ent, const PeerPush push, const SweepBlend blend) {
  extern __shared__ unsigned long long sh[];
  __shared__ unsigned int ticket;
  const int p = next_pow2(kc_in);
  unsigned long long* key = sh;
  int* val = reinterpret_cast<int*>(sh + p);
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < p; t += blockDim.x) {
    unsigned long long kk = 0ull;
    int vv = -1;
    if (t < kc_in) {
      const int idx = cand_idx[static_cast<size_t>(b) * kc_in + t];
      if (idx >= 0) {
        float score = cand_sim[static_cast<size_t>(b) * kc_in + t];
        if (blend.enabled) {   // rerank_scope = "full": keep the kcand best by the blended score, evaluated as the re-rank does
          score = __fadd_rn(__fmul_rn(blend.w_sim, score), __fmul_rn(blend.w_dewi, dewi[idx]));
          if (blend.use_pref) score = __fadd_rn(score, __fmul_rn(blend.pref, ent[idx]));
        }
        kk = make_key(score, static_cast<uint32_t>(idx));
        vv = t;
      }
    }
    key[t] = kk;
    val[t] = vv;
  }
  bitonic_sort_desc(key, val, p);
  const size_t bk = static_cast<size_t>(gridDim.x) * kcand;   // elements per section of a block
  for (int t = threadIdx.x; t < kcand; t += blockDim.x) {
    const size_t o = static_cast<size_t>(b) * kcand + t;
    const int slot = (t < p) ? val[t] : -1;
    float v_sim = -INFINITY, v_dewi = 0.f, v_ent = 0.f;
    long long v_id = -1;
    if (slot >= 0) {
      const int idx = cand_idx[static_cast<size_t>(b) * kc_in + slot];
      v_sim = cand_sim[static_cast<size_t>(b) * kc_in + slot];
      v_id = id_base + idx;
      v_dewi = dewi[idx];
      v_ent = ent[idx];
    }
    if (!PUSH) {
      out_sim[o] = v_sim;
      out_id[o] = v_id;
      out_dewi[o] = v_dewi;
      out_ent[o] = v_ent;
    } else {
      for (int r = 0; r < push.world; ++r) {  // block layout: [id i64 bk | sim bk | dewi bk | ent bk]
        char* blk = reinterpret_cast<char*>(push.base[r]) + static_cast<long long>(push.my_rank) * push.block_stride;
        reinterpret_cast<long long*>(blk)[o] = v_id;
        reinterpret_cast<float*>(blk + 8 * bk)[o] = v_sim;
        reinterpret_cast<float*>(blk + 12 * bk)[o] = v_dewi;
        reinterpret_cast<float*>(blk + 16 * bk)[o] = v_ent;
      }
    }
  }
  if (PUSH) {
    // every block's stores are fenced at system scope before its ticket; the block that draws the last ticket
    // therefore publishes flags that cover all of them (threadfence reduction pattern, system-wide)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(push.ticket, 1u);
    __syncthreads();
    if (ticket == gridDim.x - 1) {
      __threadfence_system();
      if (threadIdx.x < push.world)
        st_release_sys(reinterpret_cast<unsigned int*>(push.flags[threadIdx.x]) + push.my_rank, push.seq);
      if (threadIdx.x == 0) *push.ticket = 0u;
    }
  }
}

// ---- rerank -------------------------------------------------------------------------------------
// Candidate c of query b lives in shard g = c / kcand, slot j = c % kcand, at element
// b * kcand + j of that shard's arrays; shard g's arrays start g * shard_stride BYTES after shard 0's
// (the layout an all-gather of per-rank [B, kcand] blocks produces).
template <typename T>
__device__ __forceinline__ T shard_at(const T* base, int c, int b, int kcand, long long shard_stride) {
  const int g = c / kcand, j = c - g * kcand;
  // (L2 loads: in the fused exchange these blocks are written by peer GPUs while the kernel waits)
  const char* p = reinterpret_cast<const char*>(base) + static_cast<long long>(g) * shard_stride;
  return __ldcg(reinterpret_cast<const T*>(p) + static_cast<size_t>(b) * kcand + j);
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(kSelThreads)
rerank_kernel(const float* __restrict__ sim, const long long* __restrict__ id, const float* __restrict__ dewi,
              const float* __restrict__ ent, int n_shards, int kcand, long long shard_stride, int cand_count, int k,
              float w_sim, float w_dewi, float pref, int use_pref, long long* __restrict__ out_id,
              float* __restrict__ out_score, const unsigned int* ready_flags, unsigned int seq,
              unsigned int* status, unsigned long long timeout_ns) {
  extern __shared__ unsigned long long sh[];
  if (ready_flags) {
    // fused exchange: every rank's block must have landed in this buffer (flags are released by the peers'
    // finalize kernels).  The wait is bounded in WALL time (%globaltimer, independent of the SM clock): host-side
    // skew between ranks (GC, logging, a debugger) is unbounded in principle, and a peer may be dead.  A block
    // that gives up does NOT trap -- that would poison the context and lose the resident corpus -- it reports
    // `*status = seq` (host-visible memory) and returns ids of -1 for its query; the host raises on its next call.
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if (threadIdx.x < n_shards) {
      const unsigned long long t0 = globaltimer_ns();
      while (static_cast<int>(ld_acquire_sys(ready_flags + threadIdx.x) - seq) < 0) {
        if (globaltimer_ns() - t0 > timeout_ns) { timed_out = 1; break; }
      }
    }
    __syncthreads();
    if (timed_out) {
      for (int t = threadIdx.x; t < k; t += blockDim.x) {
        out_id[static_cast<size_t>(blockIdx.x) * k + t] = -1;
        out_score[static_cast<size_t>(blockIdx.x) * k + t] = -INFINITY;
      }
      if (threadIdx.x == 0 && status) {
        *reinterpret_cast<volatile unsigned int*>(status) = seq;
        __threadfence_system();
      }
      return;
    }
  }
  // Selections by rank counting (keys are (score, ~id): distinct among valid candidates): a couple of hundred broadcast
  // shared-memory reads per thread instead of two bitonic sorts of ~40 barrier rounds each.
  const int ncand = n_shards * kcand;
  const int p = next_pow2(ncand);
  unsigned long long* key = sh;                                   // [ncand] similarity keys, then adjusted-score keys
  int* val = reinterpret_cast<int*>(sh + p);                      // [ncand] candidate slot of the entry ranked t-th by similarity
  float* adj = reinterpret_cast<float*>(val + p);
  const int b = blockIdx.x;
  // 1. candidate set: the cand_count best by similarity (backends.py:439-447)
  for (int t = threadIdx.x; t < ncand; t += blockDim.x) {
    const long long g = shard_at(id, t, b, kcand, shard_stride);
    key[t] = (g >= 0) ? make_key(shard_at(sim, t, b, kcand, shard_stride), static_cast<uint32_t>(g)) : 0ull;
    val[t] = -1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ncand; e += blockDim.x) {
    const unsigned long long x = key[e];
    if (x == 0ull) continue;
    int rank = 0;
    for (int f = 0; f < ncand; ++f) rank += (key[f] > x) ? 1 : 0;
    if (rank < cand_count) val[rank] = e;
  }
  __syncthreads();
  // 2. blend in fp32, one rounding per operation as numpy does (backends.py:461-465)
  for (int t = threadIdx.x; t < ncand; t += blockDim.x) {
    const int slot = (t < cand_count) ? val[t] : -1;
    unsigned long long kk = 0ull;
    if (slot >= 0) {
      float a = __fadd_rn(__fmul_rn(w_sim, shard_at(sim, slot, b, kcand, shard_stride)),
                          __fmul_rn(w_dewi, shard_at(dewi, slot, b, kcand, shard_stride)));
      if (use_pref) a = __fadd_rn(a, __fmul_rn(pref, shard_at(ent, slot, b, kcand, shard_stride)));
      kk = make_key(a, static_cast<uint32_t>(shard_at(id, slot, b, kcand, shard_stride)));
      adj[t] = a;
    }
    key[t] = kk;   // (the similarity keys are no longer needed: every rank has been taken)
  }
  for (int t = threadIdx.x; t < k; t += blockDim.x) {   // defaults: fewer than k candidates
    out_id[static_cast<size_t>(b) * k + t] = -1;
    out_score[static_cast<size_t>(b) * k + t] = -INFINITY;
  }
  __syncthreads();
  // 3. top-k by adjusted score, descending, ties: lower id first (backends.py:468-471)
  for (int e = threadIdx.x; e < ncand; e += blockDim.x) {
    const unsigned long long x = key[e];
    if (x == 0ull) continue;
    int rank = 0;
    for (int f = 0; f < ncand; ++f) rank += (key[f] > x) ? 1 : 0;
    if (rank < k) {
      out_id[static_cast<size_t>(b) * k + rank] = shard_at(id, val[e], b, kcand, shard_stride);
      out_score[static_cast<size_t>(b) * k + rank] = adj[e];
    }
  }
}


// ---- fused tail ---------------------------------------------------------------------------------------------
// Everything after the sweep for ONE query per block, in one launch instead of four or five: merge the per-CTA
// partial lists (merge_select), give the certificate of a single-plane sweep (certificate_kernel), re-score the
// candidates exactly (rescore), order them and attach the payload columns (finalize_local, incl. the peer push of
// the multi-GPU exchange) and -- on a single shard -- blend and pick the final top-k (rerank).  The stages are
// per-query independent, so block-level barriers replace the launches; selections are done by RANK COUNTING
// (every thread counts how many keys beat its own: a few hundred broadcast shared-memory reads) instead of bitonic
// sorts (dozens of block-wide barrier rounds each), which is what made the separate kernels latency-bound.
constexpr int kTailThreads = 512;      // per block when there are blocks to spare; small batches use kTailThreadsWide
constexpr int kTailThreadsWide = 1024; // B <= #SMs: one block per SM at most, so the block itself should be wide
constexpr int kTailWindow = 2048;     // candidate keys gathered between two selections
constexpr int kTailMaxKc = 512;       // list capacity the fused tail supports (longer lists use the separate kernels)

struct TailArgs {
  // merge: partial lists [item][kc][128] of (score, row); item = chunk * n_qb + qb
  const float* part_s;
  const int* part_i;
  int n_chunks, n_qb, kc;
  // certificate of the single-plane sweep (q_planes == 0: none)
  const float* q_stats;
  int q_planes;
  const unsigned int* plane_max;
  float gamma;
  int need;
  int* fails;
  // exact re-score (rows == nullptr: the swept scores are exact already)
  const void* rows;
  int rows_are_bf16, dim;
  const float* qn;
  // local result: the shard's best kcand rows by exact similarity, with payload columns
  int kcand;
  long long id_base;
  const float* dewi;
  const float* ent;
  float* out_sim;
  long long* out_id;
  float* out_dewi;
  float* out_ent;
  PeerPush push;          // world > 0: write the block into every rank's gather buffer instead
  // single shard: DEWI blend + final top-k in the same launch (backends.py:461-471)
  int fuse_rerank, k;
  float w_sim, w_dewi, pref;
  int use_pref;
  long long* fin_id;
  float* fin_score;
};

__device__ __forceinline__ float key_score(unsigned long long key) {
  const uint32_t o = static_cast<uint32_t>(key >> 32);
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ uint32_t key_id(unsigned long long key) { return ~static_cast<uint32_t>(key); }

// dst[rank] = key for the `keep` largest of the m distinct keys src[0 .. m) (block-wide; dst and src disjoint).
__device__ __forceinline__ void select_by_rank(const unsigned long long* src, int m, unsigned long long* dst, int keep) {
  for (int e = threadIdx.x; e < m; e += blockDim.x) {
    const unsigned long long x = src[e];
    int rank = 0;
    for (int f = 0; f < m; ++f) rank += (src[f] > x) ? 1 : 0;   // same address in every lane of a warp: broadcast
    if (rank < keep) dst[rank] = x;
  }
}

template <typename RowT>
__device__ __forceinline__ float exact_dot(const RowT* __restrict__ r, const float* __restrict__ q, int dim, int lane) {
  float acc = 0.f;
  for (int d = lane * 8; d < dim; d += 256) {
    float x[8];
    if (sizeof(RowT) == 2) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(r + d));
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        x[2 * i] = __uint_as_float(u[i] << 16);
        x[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
      }
    } else {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(r + d));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(r + d) + 1);
      x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
      x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
    }
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(q + d));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(q + d) + 1);
    acc = fmaf(x[0], q0.x, acc); acc = fmaf(x[1], q0.y, acc); acc = fmaf(x[2], q0.z, acc); acc = fmaf(x[3], q0.w, acc);
    acc = fmaf(x[4], q1.x, acc); acc = fmaf(x[5], q1.y, acc); acc = fmaf(x[6], q1.z, acc); acc = fmaf(x[7], q1.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

__global__ void __launch_bounds__(kTailThreadsWide)
tail_kernel(const TailArgs a) {
  __shared__ unsigned long long win[kTailWindow];      // keys gathered from the partial lists
  __shared__ unsigned long long best[2][kTailMaxKc];   // running best (sorted), double-buffered
  __shared__ float s_sim[kTailMaxKc], s_dewi[kTailMaxKc], s_ent[kTailMaxKc];
  __shared__ int fill_s, nb_s;
  __shared__ float bar_s;
  __shared__ unsigned int ticket;
  const int b = blockIdx.x;
  const int qb = b / kQueryBlock, ql = query_lane(b % kQueryBlock);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kc = a.kc;
  if (threadIdx.x == 0) { fill_s = 0; nb_s = 0; bar_s = -INFINITY; }
  __syncthreads();

  // 1. merge: stream the query's n_chunks * kc partial entries through the window; whenever it is (nearly) full, and at
  //    the end, keep the kc best of {running best, window} by rank counting.  Keys are (score, ~row): distinct.
  int cur = 0;
  unsigned long long floor_key = 0ull;   // admission bar: the kc-th best so far (0: any valid key)
  const int total = a.n_chunks * kc;
  for (int base = 0; base < total; base += blockDim.x) {
    const int e = base + threadIdx.x;
    unsigned long long kk = 0ull;
    if (e < total) {
      const int chunk = e / kc, k = e - chunk * kc;
      const size_t off = ((static_cast<size_t>(chunk) * a.n_qb + qb) * kc + k) * kQueryBlock + ql;
      const int idx = __ldg(a.part_i + off);
      const float sc = __ldg(a.part_s + off);   // (issued with the index, not after it: the round is two L2 latencies otherwise)
      if (idx >= 0) kk = make_key(sc, static_cast<uint32_t>(idx));
    }
    const bool pass = kk > floor_key;
    const unsigned int m = __ballot_sync(0xffffffffu, pass);
    int wbase = 0;
    if (lane == 0 && m) wbase = atomicAdd(&fill_s, __popc(m));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (pass) win[wbase + __popc(m & ((1u << lane) - 1u))] = kk;   // (fits: a selection runs whenever < blockDim slots remain)
    __syncthreads();
    const int fill = fill_s;
    __syncthreads();   // everybody has read the fill count before the next round's appends move it
    // A selection costs (fill + nb)^2 / blockDim comparisons per thread, so it runs EARLY: as soon as a few hundred keys
    // are waiting -- after the first one the admission bar is the kc-th best so far and later rounds add a handful
    // (measured at 12.5M rows, B = 64, ~4600 valid partial entries per query: 54 us with selections of ~1000 keys, see
    // profiles/).  It must run while the window still has room for the next round's keys and the running best.
    if (fill >= 256 || fill > kTailWindow - 2 * static_cast<int>(blockDim.x) - kTailMaxKc ||
        base + static_cast<int>(blockDim.x) >= total) {
      const int nb = nb_s;
      // append the running best to the window, then select into the other buffer
      for (int t = threadIdx.x; t < nb; t += blockDim.x) win[fill + t] = best[cur][t];
      __syncthreads();
      const int m_all = fill + nb;
      select_by_rank(win, m_all, best[cur ^ 1], kc);
      __syncthreads();
      cur ^= 1;
      if (threadIdx.x == 0) { nb_s = min(kc, m_all); fill_s = 0; }
      __syncthreads();
      if (nb_s == kc) floor_key = best[cur][kc - 1];
    }
  }
  const int nb = nb_s;   // candidates: best[cur][0 .. nb), sorted by swept score, descending

  // 2. certificate of the single-plane sweep (see certificate_kernel)
  if (a.q_planes && threadIdx.x == 0 && nb > 0) {
    const float row_err = __uint_as_float(a.plane_max[0]), row_hi = __uint_as_float(a.plane_max[1]);
    const float qn = a.q_stats[b * 4 + 0], qe = a.q_stats[b * 4 + (a.q_planes == 1 ? 1 : 2)];
    const float eps = __fmul_ru(qn, row_err) + __fmul_ru(qe, row_hi) + __fmul_ru(a.gamma, __fmul_ru(qn + qe, row_hi));
    const int need = min(a.need, nb);
    const float bar = __fsub_rd(key_score(best[cur][need - 1]), __fmul_ru(2.f, eps));
    bar_s = bar;
    if (nb == kc && !(key_score(best[cur][kc - 1]) < bar)) atomicAdd(a.fails, 1);
  }
  __syncthreads();

  // 3. exact re-score (one warp per candidate), dropping what the certificate proves irrelevant; keys by exact score
  unsigned long long* exact = win;   // [nb]
  {
    const float bar = bar_s;
    for (int t = warp; t < nb; t += blockDim.x / 32) {
      const unsigned long long kk = best[cur][t];
      const uint32_t idx = key_id(kk);
      float s = key_score(kk);
      unsigned long long out = 0ull;
      if (!(a.q_planes && s < bar)) {
        if (a.rows) {
          const float* q = a.qn + static_cast<size_t>(b) * a.dim;
          s = a.rows_are_bf16 ? exact_dot(static_cast<const __nv_bfloat16*>(a.rows) + static_cast<size_t>(idx) * a.dim, q, a.dim, lane)
                              : exact_dot(static_cast<const float*>(a.rows) + static_cast<size_t>(idx) * a.dim, q, a.dim, lane);
        }
        out = make_key(s, idx);
      }
      if (lane == 0) exact[t] = out;
    }
  }
  __syncthreads();

  // 4. order by exact similarity (ties: lower row first), keep the shard's top kcand, attach id + payload columns
  unsigned long long* sorted = best[cur ^ 1];
  for (int t = threadIdx.x; t < kc; t += blockDim.x) sorted[t] = 0ull;
  __syncthreads();
  for (int e = threadIdx.x; e < nb; e += blockDim.x) {
    const unsigned long long x = exact[e];
    if (x == 0ull) continue;
    int rank = 0;
    for (int f = 0; f < nb; ++f) rank += (exact[f] > x) ? 1 : 0;
    sorted[rank] = x;
  }
  __syncthreads();
  const size_t bk = static_cast<size_t>(gridDim.x) * a.kcand;   // elements per section of a pushed block
  for (int t = threadIdx.x; t < a.kcand; t += blockDim.x) {
    const unsigned long long kk = (t < kc) ? sorted[t] : 0ull;
    float v_sim = -INFINITY, v_dewi = 0.f, v_ent = 0.f;
    long long v_id = -1;
    if (kk != 0ull) {
      const uint32_t idx = key_id(kk);
      v_sim = key_score(kk);
      v_id = a.id_base + idx;
      v_dewi = a.dewi[idx];
      v_ent = a.ent[idx];
    }
    const size_t o = static_cast<size_t>(b) * a.kcand + t;
    if (a.push.world > 0) {
      for (int r = 0; r < a.push.world; ++r) {  // block layout: [id i64 bk | sim bk | dewi bk | ent bk]
        char* blk = reinterpret_cast<char*>(a.push.base[r]) + static_cast<long long>(a.push.my_rank) * a.push.block_stride;
        reinterpret_cast<long long*>(blk)[o] = v_id;
        reinterpret_cast<float*>(blk + 8 * bk)[o] = v_sim;
        reinterpret_cast<float*>(blk + 12 * bk)[o] = v_dewi;
        reinterpret_cast<float*>(blk + 16 * bk)[o] = v_ent;
      }
    } else if (a.out_sim) {
      a.out_sim[o] = v_sim;
      a.out_id[o] = v_id;
      a.out_dewi[o] = v_dewi;
      a.out_ent[o] = v_ent;
    }
    if (a.fuse_rerank && t < kTailMaxKc) { s_sim[t] = v_sim; s_dewi[t] = v_dewi; s_ent[t] = v_ent; }
  }
  if (a.push.world > 0) {
    // every block's stores are fenced at system scope before its ticket; the block that draws the last ticket publishes
    // the ready flags -- unless a certificate failed somewhere in the batch: the host re-runs it, and the re-run publishes
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(a.push.ticket, 1u);
    __syncthreads();
    if (ticket == gridDim.x - 1) {
      __threadfence_system();
      const bool ok = !a.q_planes || *reinterpret_cast<volatile int*>(a.fails) == 0;
      if (ok && threadIdx.x < a.push.world)
        st_release_sys(reinterpret_cast<unsigned int*>(a.push.flags[threadIdx.x]) + a.push.my_rank, a.push.seq);
      if (threadIdx.x == 0) *a.push.ticket = 0u;
    }
    return;
  }
  if (!a.fuse_rerank) return;
  __syncthreads();

  // 5. single shard: adj = w_sim * sim + w_dewi * dewi (+ pref * ent) in fp32, one rounding per operation as numpy
  //    (backends.py:461-465); top-k by adj, descending, ties: lower id first (backends.py:468-471)
  unsigned long long* adjk = win;   // [kcand]
  const int nc = min(a.kcand, kTailMaxKc);
  for (int t = threadIdx.x; t < nc; t += blockDim.x) {
    const unsigned long long kk = (t < kc) ? sorted[t] : 0ull;
    unsigned long long out = 0ull;
    if (kk != 0ull) {
      float adj = __fadd_rn(__fmul_rn(a.w_sim, s_sim[t]), __fmul_rn(a.w_dewi, s_dewi[t]));
      if (a.use_pref) adj = __fadd_rn(adj, __fmul_rn(a.pref, s_ent[t]));
      out = make_key(adj, static_cast<uint32_t>(a.id_base + key_id(kk)));
    }
    adjk[t] = out;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < a.k; t += blockDim.x) {   // defaults: fewer than k candidates
    a.fin_id[static_cast<size_t>(b) * a.k + t] = -1;
    a.fin_score[static_cast<size_t>(b) * a.k + t] = -INFINITY;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nc; e += blockDim.x) {
    const unsigned long long x = adjk[e];
    if (x == 0ull) continue;
    int rank = 0;
    for (int f = 0; f < nc; ++f) rank += (adjk[f] > x) ? 1 : 0;
    if (rank < a.k) {
      a.fin_id[static_cast<size_t>(b) * a.k + rank] = a.id_base + key_id(sorted[e]);
      a.fin_score[static_cast<size_t>(b) * a.k + rank] = key_score(x);
    }
  }
}

}  // namespace

int launch_seed_from_maxima(const float* maxima, int n_chunks, int n_qb, int B, int kc, float* seed, cudaStream_t stream) {
  if (n_chunks < kc || n_chunks > kWindow) return fail("seed_from_maxima: item count outside [kc, 2048]");
  // 1024 threads hold the (at most 2048) maxima two per thread
  if (env_int("DEWI_SEED_RANKCOUNT", 0) != 0) seed_from_maxima_rank_kernel<<<B, 512, 0, stream>>>(maxima, n_chunks, n_qb, kc, seed);  // experiments
  else seed_from_maxima_kernel<<<B, kMergeThreads, 0, stream>>>(maxima, n_chunks, n_qb, kc, seed);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int launch_seed_from_partials(const float* part_s, int n_chunks_done, int n_qb, int B, int kc, const float* seed_in, float* seed_out,
                              cudaStream_t stream) {
  const int n_vals = n_chunks_done * kc;
  if (n_vals < kc || n_vals > kWindow) return fail("seed_from_partials: list count out of range");
  if (env_int("DEWI_SEED_RANKCOUNT", 0) != 0)
    seed_from_partials_rank_kernel<<<B, 512, 0, stream>>>(part_s, n_vals, n_qb, kc, seed_in, seed_out);  // experiments
  else
    seed_from_partials_kernel<<<B, kMergeThreads, 0, stream>>>(part_s, n_vals, n_qb, kc, seed_in, seed_out);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int launch_merge_select(const Partials& p, int B, int kc_out, int* cand_idx, float* cand_sim, cudaStream_t stream) {
  if (kc_out > kWindow / 2) return fail("candidate count too large for merge window");
  merge_select_kernel<<<B, kMergeThreads, 0, stream>>>(p.s, p.i, p.n_chunks, p.n_qb, p.kc, kc_out, cand_idx, cand_sim);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int launch_certificate(const float* cand_sim, const int* cand_idx, int B, int kc, int need, const float* q_stats, int q_planes,
                       const unsigned int* plane_max, int dim, int* fails, float* bar, cudaStream_t stream) {
  if (need < 1 || need > kc) return fail("certificate: need must lie in [1, kc]");
  const float gamma = static_cast<float>(dim) * 2.384185791015625e-07f;   // dim * 2^-22
  certificate_kernel<<<static_cast<int>(ceil_div(B, 128)), 128, 0, stream>>>(cand_sim, cand_idx, B, kc, need, q_stats, q_planes,
                                                                              plane_max, gamma, fails, bar);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int launch_rescore(const void* rows, int rows_are_bf16, int dim, const float* qn, int* cand_idx, int B, int kc,
                   float* cand_sim, cudaStream_t stream, const float* bar) {
  if (dim % 8 != 0) return fail("rescore needs dim % 8 == 0");
  const int total = B * kc;
  const int threads = 256;
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(total) * 32, threads));
  if (rows_are_bf16)
    rescore_kernel<__nv_bfloat16><<<blocks, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(rows), dim, qn,
                                                                   cand_idx, total, kc, cand_sim, bar);
  else
    rescore_kernel<float><<<blocks, threads, 0, stream>>>(static_cast<const float*>(rows), dim, qn, cand_idx, total, kc,
                                                           cand_sim, bar);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int launch_rescore_any(const void* rows, int rows_are_bf16, int dim, int is_l2, const float* qn, const int* cand_idx, int B, int kc,
                       float* cand_sim, cudaStream_t stream) {
  const int total = B * kc;
  const int threads = 256;
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(total) * 32, threads));
  if (rows_are_bf16)
    rescore_any_kernel<__nv_bfloat16><<<blocks, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(rows), dim, qn, cand_idx, total,
                                                                       kc, cand_sim, is_l2);
  else
    rescore_any_kernel<float><<<blocks, threads, 0, stream>>>(static_cast<const float*>(rows), dim, qn, cand_idx, total, kc, cand_sim,
                                                               is_l2);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int launch_finalize_local(const int* cand_idx, const float* cand_sim, int B, int kc_in, int kcand, int64_t id_base,
                          const float* dewi, const float* ent, float* out_sim, int64_t* out_id, float* out_dewi,
                          float* out_ent, cudaStream_t stream, const PeerPush* push, const SweepBlend* blend_) {
  const SweepBlend blend = blend_ ? *blend_ : SweepBlend();
  const int p = next_pow2(kc_in);
  if (p > 4096) return fail("too many candidates per query");
  const size_t smem = static_cast<size_t>(p) * 12;
  if (push && push->world > 0) {
    if (push->world > kMaxPeers || !push->ticket) return fail("invalid peer-push descriptor");
    finalize_local_kernel<1><<<B, kSelThreads, smem, stream>>>(cand_idx, cand_sim, kc_in, kcand, id_base, dewi, ent, nullptr,
                                                               nullptr, nullptr, nullptr, *push, blend);
  } else {
    finalize_local_kernel<0><<<B, kSelThreads, smem, stream>>>(cand_idx, cand_sim, kc_in, kcand, id_base, dewi, ent, out_sim,
                                                               reinterpret_cast<long long*>(out_id), out_dewi, out_ent, PeerPush(), blend);
  }
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int tail_supported(int kc, int kcand, int k) { return kc >= 1 && kc <= kTailMaxKc && kcand >= 1 && kcand <= kTailMaxKc && k <= kcand; }

int launch_tail(const Partials& p, int B, const TailCert* cert, const void* rows, int rows_are_bf16, int dim, const float* qn, int kcand,
                int64_t id_base, const float* dewi, const float* ent, float* out_sim, int64_t* out_id, float* out_dewi, float* out_ent,
                const PeerPush* push, const TailRerank* rr, cudaStream_t stream) {
  if (!tail_supported(p.kc, kcand, rr ? rr->k : 1)) return fail("fused tail: unsupported list capacity");
  TailArgs a;
  a.part_s = p.s;
  a.part_i = p.i;
  a.n_chunks = p.n_chunks;
  a.n_qb = p.n_qb;
  a.kc = p.kc;
  a.q_stats = cert ? cert->q_stats : nullptr;
  a.q_planes = cert ? cert->q_planes : 0;
  a.plane_max = cert ? cert->plane_max : nullptr;
  a.gamma = static_cast<float>(dim) * 2.384185791015625e-07f;   // dim * 2^-22
  a.need = cert ? cert->need : 0;
  a.fails = cert ? cert->fails : nullptr;
  a.rows = rows;
  a.rows_are_bf16 = rows_are_bf16;
  a.dim = dim;
  a.qn = qn;
  a.kcand = kcand;
  a.id_base = id_base;
  a.dewi = dewi;
  a.ent = ent;
  a.out_sim = out_sim;
  a.out_id = reinterpret_cast<long long*>(out_id);
  a.out_dewi = out_dewi;
  a.out_ent = out_ent;
  a.push = (push && push->world > 0) ? *push : PeerPush();
  a.fuse_rerank = rr ? 1 : 0;
  a.k = rr ? rr->k : 0;
  a.w_sim = rr ? rr->w_sim : 0.f;
  a.w_dewi = rr ? rr->w_dewi : 0.f;
  a.pref = rr ? rr->pref : 0.f;
  a.use_pref = rr ? rr->use_pref : 0;
  a.fin_id = rr ? reinterpret_cast<long long*>(rr->out_id) : nullptr;
  a.fin_score = rr ? rr->out_score : nullptr;
  if (rows && dim % 8 != 0) return fail("fused tail: the exact re-score needs dim % 8 == 0");
  tail_kernel<<<B, B <= current_sm_count() ? kTailThreadsWide : kTailThreads, 0, stream>>>(a);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int launch_rerank(const float* sim, const int64_t* id, const float* dewi, const float* ent, int B, int n_shards, int kcand,
                  int64_t shard_stride_bytes, int cand_count, int k, float w_sim, float w_dewi, float pref, int use_pref,
                  int64_t* out_id, float* out_score, cudaStream_t stream, const unsigned int* ready_flags, unsigned int seq,
                  unsigned int* status, double timeout_s) {
  const int p = next_pow2(n_shards * kcand);
  if (ready_flags && n_shards > kSelThreads) return fail("too many shards for the fused exchange");
  if (p > 8192) return fail("too many gathered candidates per query");
  const size_t smem = static_cast<size_t>(p) * 16;
  auto kern = rerank_kernel;
  if (smem > 48 * 1024)
    DEWI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<B, kSelThreads, smem, stream>>>(sim, reinterpret_cast<const long long*>(id), dewi, ent, n_shards, kcand,
                                         shard_stride_bytes, cand_count, k, w_sim, w_dewi, pref, use_pref,
                                         reinterpret_cast<long long*>(out_id), out_score, ready_flags, seq, status,
                                         static_cast<unsigned long long>(std::max(timeout_s, 1e-3) * 1e9));
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dewi
