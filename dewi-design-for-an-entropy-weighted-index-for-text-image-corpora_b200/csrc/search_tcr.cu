// K1r -- the similarity sweep for batches of at most 64 queries with the operands SWAPPED: corpus rows sit on
// the MMA M dimension (128 per instruction), the queries on N (16 / 32 / 64).
//
// Same statement of the reference as search_tc.cu (`np.dot(self._embeddings, query.T)` followed by the first half
// of `np.argpartition`, src/dewi/backends.py:431-447); same partial-list output.  Why a second kernel: with queries on
// M a batch of 1..64 queries still executes M = 64 x N = 256 MMAs -- 128 tensor-pipe cycles per 256 rows per k-step
// whatever B is (ncu: tensor pipe 45 % active in an HBM-bound kernel), which on the power-capped single-GPU 100M-row
// run is paid for in SM clock.  Here one k-step of 128 rows costs N / 2 cycles: 8 at B <= 16, 16 at B <= 32, 32 at
// B <= 64 -- 8x / 4x / 2x less tensor work per corpus row -- and the epilogue reads N columns per row instead of 256
// per query.
//
//     D[128 corpus rows, QN queries] (fp32, TMEM) = sum_k  E[128, k] (16-bit plane, smem) . Q[QN, k]^T (smem)
//
// TMEM lane == corpus row, column == query: an epilogue thread owns one ROW per half-tile and tests its QN scores
// against the queries' admission thresholds (shared memory, broadcast reads).  The per-query candidate lists are shared
// by the four epilogue warps: a hit takes a slot with a shared-memory atomicAdd; when a list is full the four warps
// prune together (one named barrier per half-tile keeps them in step, see `scan_half`).  After threshold seeding
// (api.cu) a CTA sees ~kc hits per query over its whole share of the corpus, so the slow path is rare.
//
// Warp roles (192 threads, one CTA per SM, persistent over corpus chunks):
//   warp 0      TMA producer : the query block once (resident: dim / 64 k-blocks), then the corpus ring (32 KB stages)
//   warp 1      MMA issuer   : per k-block two M = 128 halves x four K = 16 steps into one of two accumulator pairs
//   warps 2..5  epilogue     : tcgen05.ld -> threshold test -> shared candidate lists / pre-pass maxima
#include <algorithm>

#include "internal.h"
#include "ptx.cuh"
#include "sweep_epilogue.cuh"

namespace dewi {

namespace {

using namespace sweep;

constexpr int kThreads = 192;
constexpr int kRowTile = 256;                 // corpus rows per ring stage: two M = 128 halves
constexpr uint32_t kEStageBytes = kRowTile * 128;
constexpr uint32_t kHalfBytes = 128 * 128;

struct TcrArgs {
  int n_rows;
  int n_tiles;
  int n_kb;
  int n_chunks;
  int kc;
  int e_stages;
  int n_queries;      // real queries (columns beyond them never admit anything)
  float* part_s;
  int* part_i;
  const float* seed;  // optional admission thresholds from the sample pre-pass
  int seed_stride, seed_off, seed_queries;
  float* max_out;     // pre-pass mode: maxima instead of candidate lists (internal.h: SweepSeed)
  int max_groups;
  int fp16;
  SweepBlend blend;   // rerank_scope = "full": select by the blended key instead of by similarity
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// order-preserving float <-> signed int (an involution), so maxima can use redux.sync / atomicMax
__device__ __forceinline__ int f2key(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

template <int CH>
__device__ __forceinline__ void load_cols(uint32_t taddr, float (&v)[CH]);
template <>
__device__ __forceinline__ void load_cols<16>(uint32_t taddr, float (&v)[16]) { ptx::tmem_ld_32x16(taddr, v); }
template <>
__device__ __forceinline__ void load_cols<32>(uint32_t taddr, float (&v)[32]) { ptx::tmem_ld_32x32(taddr, v); }

// The epilogue's shared state: QN candidate lists of `cap` slots laid out [slot][query], their fill counts and
// admission thresholds, and two overflow flags used alternately by consecutive half-tiles.
template <int QN>
struct RowLists {
  float* ls;
  int* li;
  float* thr;
  int* cnt;
  int* ovf;
  int kc, cap;
};

// One chunk of CH queries of this thread's row: every score above its query's threshold takes a slot of that query's
// list.  `only` restricts the attempt to a set of columns (retries).  Returns the columns whose list was full.
template <int QN, int CH>
__device__ __forceinline__ uint32_t append_chunk(const RowLists<QN>& L, const float (&v)[CH], const float (&t)[CH], uint32_t only,
                                                 int q0, int row) {
  uint32_t full = 0;
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    if (((only >> j) & 1u) && v[j] > t[j]) {
      const int slot = atomicAdd(&L.cnt[q0 + j], 1);
      if (slot < L.cap) {
        L.ls[slot * QN + q0 + j] = v[j];
        L.li[slot * QN + q0 + j] = row;
      } else {
        full |= 1u << j;
      }
    }
  }
  return full;
}

template <int CH>
__device__ __forceinline__ void load_thr(const float* thr, float (&t)[CH]) {
#pragma unroll
  for (int j = 0; j < CH; j += 4) {
    const float4 x = *reinterpret_cast<const float4*>(thr + j);
    t[j] = x.x; t[j + 1] = x.y; t[j + 2] = x.z; t[j + 3] = x.w;
  }
}

// One half-tile (128 rows x QN queries) of the main sweep.  `tc` addresses this warp's TMEM lanes at the half's first
// column; `row` is this thread's corpus row.  All four epilogue warps call this for the same half-tile.
//   1. fast path: QN compares per row against thresholds read as shared-memory broadcasts; a warp with no hit is done.
//   2. hits take list slots with atomicAdd; a full list leaves the hit pending and raises this half-tile's flag.
//   3. barrier; no flag -> done.  Otherwise the four warps prune the full lists (rank counting, coop_prune: the kc
//      best stay, the kc-th becomes the threshold), pending hits that still beat their threshold try again, repeat.
// Thresholds only rise and admission is strict, so a stale threshold merely admits a row that the next prune drops.
// The flag of half-tile p is zero again when its loop ends, and is next written after the first barrier of half-tile
// p + 1 -- which every warp reaches only after its last read of it: two flags used alternately suffice.
// Scores of one chunk of queries for this thread's row; with `key.on` they become the blended keys
// key.w * sim + key.bias (rerank_scope = "full": key.bias = w_dewi * dewi[row] + pref * ent[row]).
struct RowKey {
  bool on;
  float w, bias;
};
__device__ __forceinline__ RowKey row_key(const SweepBlend& b, int row, bool valid) {
  RowKey k{b.enabled != 0, b.w_sim, 0.f};
  if (k.on && valid) {
    k.bias = b.w_dewi * __ldg(b.dewi + row);
    if (b.use_pref) k.bias = fmaf(b.pref, __ldg(b.ent + row), k.bias);
  }
  return k;
}
template <int CH>
__device__ __forceinline__ void load_scores(uint32_t taddr, const RowKey& key, float (&v)[CH]) {
  load_cols<CH>(taddr, v);
  if (key.on) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = fmaf(key.w, v[j], key.bias);
  }
}

template <int QN, int CH>
__device__ __forceinline__ void scan_half(const RowLists<QN>& L, uint32_t tc, int row, bool valid, int e, int lane, uint32_t& par,
                                          const RowKey& key) {
  constexpr int NCH = QN / CH;
  uint32_t pend[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float v[CH], t[CH];
    load_scores<CH>(tc + c * CH, key, v);
    load_thr<CH>(L.thr + c * CH, t);
    bool hit = false;
#pragma unroll
    for (int j = 0; j < CH; ++j) hit |= v[j] > t[j];
    hit = hit && valid;   // rows past the end of the corpus were zero-filled by TMA
    pend[c] = 0;
    if (__any_sync(0xffffffffu, hit)) pend[c] = hit ? append_chunk<QN, CH>(L, v, t, 0xffffffffu, c * CH, row) : 0u;
  }
  bool any_pend = false;
#pragma unroll
  for (int c = 0; c < NCH; ++c) any_pend |= pend[c] != 0;
  volatile int* flag = L.ovf + par;
  if (any_pend) *flag = 1;
  for (;;) {
    epi_bar();
    if (*flag == 0) break;
    epi_bar();   // everyone has seen the flag
    if (e == 0 && lane == 0) *flag = 0;
    for (int q = e; q < QN; q += 4) {
      if (L.cnt[q] >= L.cap) {   // (same address in every lane: warp-uniform)
        const float t = coop_prune(L.ls + q, L.li + q, L.cap, L.kc, lane, QN);
        if (lane == 0) { L.cnt[q] = L.kc; L.thr[q] = t; }
      }
    }
    epi_bar();   // prunes and the flag reset are visible
    any_pend = false;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (__any_sync(0xffffffffu, pend[c] != 0)) {
        float v[CH], t[CH];
        load_scores<CH>(tc + c * CH, key, v);
        load_thr<CH>(L.thr + c * CH, t);
        pend[c] = pend[c] ? append_chunk<QN, CH>(L, v, t, pend[c], c * CH, row) : 0u;
        any_pend |= pend[c] != 0;
      }
    }
    if (any_pend) *flag = 1;
  }
  par ^= 1u;
}

// Pre-pass half-tile: this warp's 32 rows are one 32-row group; lane j ends up with query (q0 + j)'s maximum over them.
template <int QN, int CH>
__device__ __forceinline__ void max_half(const TcrArgs& a, int* imax, uint32_t tc, bool valid, int group, int lane, const RowKey& key) {
#pragma unroll
  for (int c = 0; c < QN / CH; ++c) {
    float v[CH];
    load_scores<CH>(tc + c * CH, key, v);
    int mine = f2key(-INFINITY);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int m = __reduce_max_sync(0xffffffffu, f2key(valid ? v[j] : -INFINITY));
      if (lane == j) mine = m;
    }
    if (lane < CH) {
      const int q = c * CH + lane;
      if (a.max_groups) a.max_out[static_cast<size_t>(group) * kQueryBlock + query_lane(q)] = key2f(mine);
      else atomicMax(&imax[q], mine);
    }
  }
}

template <int QN>
__global__ void __launch_bounds__(kThreads, 1)
search_tcr_kernel(const __grid_constant__ CUtensorMap map_e, const __grid_constant__ CUtensorMap map_q, const TcrArgs a) {
  constexpr int CH = QN < 32 ? QN : 32;
  constexpr uint32_t kQBlockBytes = QN * 128;   // one resident query k-block
  constexpr uint32_t kTmemCols = 4 * QN;        // two accumulator pairs (double buffer) x two halves x QN columns
  const uint32_t kIdesc = a.fp16 ? ptx::make_idesc_f16(128, QN) : ptx::make_idesc_bf16(128, QN);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring_e = smem;
  uint8_t* q_res = ring_e + static_cast<size_t>(a.e_stages) * kEStageBytes;
  RowLists<QN> L;
  L.kc = a.kc;
  L.cap = a.kc + kPending;
  L.ls = reinterpret_cast<float*>(q_res + static_cast<size_t>(a.n_kb) * kQBlockBytes);
  L.li = reinterpret_cast<int*>(L.ls + L.cap * QN);
  L.thr = reinterpret_cast<float*>(L.li + L.cap * QN);   // (16-byte aligned: every piece so far is a multiple of 64 B)
  L.cnt = reinterpret_cast<int*>(L.thr + QN);
  int* imax = L.cnt + QN;
  L.ovf = imax + QN;
  uint64_t* bar_e_full = reinterpret_cast<uint64_t*>(L.ovf + 4);
  uint64_t* bar_e_empty = bar_e_full + kMaxStages;
  uint64_t* bar_q_full = bar_e_empty + kMaxStages;
  uint64_t* bar_acc_full = bar_q_full + 1;
  uint64_t* bar_acc_empty = bar_acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_e);
    ptx::prefetch_tmap(&map_q);
    for (int s = 0; s < a.e_stages; ++s) {
      ptx::mbar_init(&bar_e_full[s], 1);
      ptx::mbar_init(&bar_e_empty[s], 1);
    }
    ptx::mbar_init(bar_q_full, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bar_acc_full[b], 1);
      ptx::mbar_init(&bar_acc_empty[b], 4);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: the query block once, then the corpus ring =====================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_q_full, static_cast<uint32_t>(a.n_kb) * kQBlockBytes);
      for (int kb = 0; kb < a.n_kb; ++kb)
        ptx::tma_load_2d(q_res + static_cast<size_t>(kb) * kQBlockBytes, &map_q, bar_q_full, kb * kKBlock, 0, ptx::kEvictLast);
    }
    __syncwarp();
    int se = 0;
    uint32_t pe = 0;
    for (int item = blockIdx.x; item < a.n_chunks; item += gridDim.x) {
      int t0, t1;
      tile_range(item, a.n_chunks, a.n_tiles, t0, t1);
      for (int t = t0; t < t1; ++t) {
        for (int kb = 0; kb < a.n_kb; ++kb) {
          ptx::mbar_wait(&bar_e_empty[se], pe ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&bar_e_full[se], kEStageBytes);
            ptx::tma_load_2d(ring_e + static_cast<size_t>(se) * kEStageBytes, &map_e, &bar_e_full[se], kb * kKBlock, t * kRowTile,
                             ptx::kEvictFirst);   // the corpus is read once
          }
          __syncwarp();
          if (++se == a.e_stages) { se = 0; pe ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
    int se = 0;
    uint32_t pe = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t ring_e_addr = ptx::smem_u32(ring_e), q_addr = ptx::smem_u32(q_res);
    ptx::mbar_wait(bar_q_full, 0);
    for (int item = blockIdx.x; item < a.n_chunks; item += gridDim.x) {
      int t0, t1;
      tile_range(item, a.n_chunks, a.n_tiles, t0, t1);
      for (int t = t0; t < t1; ++t) {
        ptx::mbar_wait(&bar_acc_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * 2 * QN);
        for (int kb = 0; kb < a.n_kb; ++kb) {
          ptx::mbar_wait(&bar_e_full[se], pe);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t ste = ring_e_addr + static_cast<uint32_t>(se) * kEStageBytes;
            const uint64_t dq = ptx::make_desc_sw128(q_addr + static_cast<uint32_t>(kb) * kQBlockBytes);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t de = ptx::make_desc_sw128(ste + h * kHalfBytes);
#pragma unroll
              for (int k = 0; k < kKBlock / 16; ++k) {
                const uint64_t adv = static_cast<uint64_t>(k * 2);  // 16 elements = 32 B = 2 x 16 B units
                ptx::mma_bf16_ss(tmem_d + h * QN, de + adv, dq + adv, kIdesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            ptx::mma_commit(&bar_e_empty[se]);
            if (kb == a.n_kb - 1) ptx::mma_commit(&bar_acc_full[acc]);
          }
          __syncwarp();
          if (++se == a.e_stages) { se = 0; pe ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: TMEM lane == corpus row =====================
    const int quarter = warp & 3;   // TMEM lane quarter this warp may read == 32-row group of the half-tile
    const int e = warp - 2;         // which lists this warp prunes and flushes
    const int et = e * 32 + lane;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    if (et < 4) L.ovf[et] = 0;
    int acc = 0;
    uint32_t acc_phase = 0, par = 0;
    for (int item = blockIdx.x; item < a.n_chunks; item += gridDim.x) {
      int t0, t1;
      tile_range(item, a.n_chunks, a.n_tiles, t0, t1);
      if (et < QN) {
        L.cnt[et] = 0;
        L.thr[et] = et < a.n_queries ? seed_threshold_query(a.seed, a.seed_stride, a.seed_off, a.seed_queries, et) : INFINITY;
        imax[et] = f2key(-INFINITY);
      }
      epi_bar();
      for (int t = t0; t < t1; ++t) {
        // (full scope: the two payload loads of both halves are issued BEFORE the wait for the accumulator, so their
        // DRAM latency overlaps the MMAs instead of sitting in front of every half-tile)
        RowKey keys[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = t * kRowTile + h * 128 + quarter * 32 + lane;
          keys[h] = row_key(a.blend, row, row < a.n_rows);
        }
        ptx::mbar_wait(&bar_acc_full[acc], acc_phase);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const uint32_t tc = tmem_lane + static_cast<uint32_t>(acc * 2 * QN + h * QN);
          const int row = t * kRowTile + h * 128 + quarter * 32 + lane;
          const bool valid = row < a.n_rows;
          const RowKey key = h ? keys[1] : keys[0];
          if (a.max_out) max_half<QN, CH>(a, imax, tc, valid, t * (kRowTile / 32) + h * 4 + quarter, lane, key);
          else scan_half<QN, CH>(L, tc, row, valid, e, lane, par, key);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar_acc_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (a.max_out) {
        if (!a.max_groups) {
          epi_bar();
          if (et < QN) a.max_out[static_cast<size_t>(item) * kQueryBlock + query_lane(et)] = key2f(imax[et]);
        }
      } else {
        // end of the item (every append is behind scan_half's last barrier): kc best of every list -> [k][query] partials
        for (int q = e; q < QN; q += 4) {
          int c = min(L.cnt[q], L.cap);
          if (c > L.kc) {
            coop_prune(L.ls + q, L.li + q, c, L.kc, lane, QN);
            c = L.kc;
          }
          __syncwarp();
          float* ps = a.part_s + static_cast<size_t>(item) * L.kc * kQueryBlock + query_lane(q);
          int* pi = a.part_i + static_cast<size_t>(item) * L.kc * kQueryBlock + query_lane(q);
          for (int k = lane; k < L.kc; k += 32) {
            const bool live = k < c;
            ps[k * kQueryBlock] = live ? L.ls[k * QN + q] : -INFINITY;
            pi[k * kQueryBlock] = live ? L.li[k * QN + q] : -1;
          }
        }
      }
      epi_bar();   // the lists are re-initialised by other threads than the ones that flushed them
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

size_t fixed_bytes(int kc, int qn) {
  return static_cast<size_t>(kc + kPending) * qn * 8 + static_cast<size_t>(qn) * 12 + 16 + (2 * kMaxStages + 5) * 8 + 16 +
         1024 /*alignment slack*/;
}

template <int QN>
int launch_one(const TcPlan& plan, const CUtensorMap& e0, const CUtensorMap& q0, const TcrArgs& args, cudaStream_t stream) {
  auto kern = search_tcr_kernel<QN>;
  DEWI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.smem_bytes)));
  kern<<<plan.grid, kThreads, plan.smem_bytes, stream>>>(e0, q0, args);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// Plan for the rows-on-M sweep, or non-zero (no error message) when it does not apply: more than 64 queries, a
// candidate list longer than the cooperative prune handles, or a query block too large to stay resident.
int tcr_make_plan(int dim, int64_t n_rows, int B, int kc, int sm_count, TcPlan* plan, int force_chunks) {
  if (!tc_supported(dim, n_rows) || B < 1 || B > 64 || kc + kPending > 64) return 1;
  if (env_int("DEWI_TCR", 1) == 0) return 1;   // experiments: keep the queries-on-M sweep
  const int qn = B <= 16 ? 16 : (B <= 32 ? 32 : 64);
  const size_t smem_max = 227 * 1024;
  const size_t used = fixed_bytes(kc, qn) + static_cast<size_t>(dim / kKBlock) * qn * 128;
  int stages = used < smem_max ? static_cast<int>((smem_max - used) / kEStageBytes) : 0;
  if (stages < 3) return 1;
  // Measured at 100M rows, B = 1: 3 stages 21.08 ms, 4 stages 20.94 ms, 5 and more 22.4 ms -- beyond ~128 KB in flight
  // per SM the 148 concurrent streams start to cost DRAM efficiency (profiles/r2_t_ab_stages_*.log).
  stages = std::min(stages, 4);
  if (env_set("DEWI_TC_STAGES")) stages = std::min(stages, std::max(2, env_int("DEWI_TC_STAGES", stages)));  // experiments
  const int64_t n_tiles = ceil_div(n_rows, kRowTile);
  int64_t chunks = std::min<int64_t>(sm_count, n_tiles);
  if (force_chunks > 0) chunks = force_chunks;
  chunks = std::max<int64_t>(1, std::min<int64_t>(chunks, n_tiles));
  plan->mode = 0;
  plan->q_rows = qn;
  plan->n_tile = kRowTile;
  plan->n_stages = stages;
  plan->q_stages = dim / kKBlock;
  plan->q_resident = 1;
  plan->rows_on_m = 1;
  plan->n_chunks = static_cast<int>(chunks);
  plan->grid = static_cast<int>(std::min<int64_t>(sm_count, chunks));
  plan->smem_bytes = used + static_cast<size_t>(stages) * kEStageBytes;
  return 0;
}

int tcr_launch(const TcPlan& plan, const CUtensorMap& e0, const CUtensorMap& q0, int64_t n_rows, int dim, int B, int kc,
               float* part_s, int* part_i, const SweepSeed& seed, cudaStream_t stream, int fp16_planes, const SweepBlend* blend) {
  TcrArgs a;
  if (blend) a.blend = *blend;
  a.n_rows = static_cast<int>(n_rows);
  a.n_tiles = static_cast<int>(ceil_div(n_rows, kRowTile));
  a.n_kb = dim / kKBlock;
  a.n_chunks = plan.n_chunks;
  a.kc = kc;
  a.e_stages = plan.n_stages;
  a.n_queries = B;
  a.part_s = part_s;
  a.part_i = part_i;
  a.seed = seed.values;
  a.seed_stride = seed.stride;
  a.seed_off = seed.off;
  a.seed_queries = seed.n_queries;
  a.max_out = seed.max_out;
  a.max_groups = seed.max_groups;
  a.fp16 = fp16_planes;
  if (plan.q_rows == 16) return launch_one<16>(plan, e0, q0, a, stream);
  if (plan.q_rows == 32) return launch_one<32>(plan, e0, q0, a, stream);
  if (plan.q_rows == 64) return launch_one<64>(plan, e0, q0, a, stream);
  return fail("unsupported rows-on-M sweep configuration");
}

}  // namespace dewi
