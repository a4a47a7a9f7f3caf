// K1 -- similarity sweep on the 5th-gen tensor cores with the candidate selection fused in.
//
// Replaces the hot statement of ExactIndex.search, `scores = np.dot(self._embeddings, query.T)`
// (reference src/dewi/backends.py:431-433) and the first half of `np.argpartition(scores,
// -candidate_count)` (:444): the [queries x corpus] score matrix is produced tile by tile in TMEM
// and reduced to per-query candidate lists on chip; it never reaches HBM.
//
// Shape of the contraction (per CTA, per corpus tile):
//     D[128 queries, N_TILE corpus rows] (fp32, TMEM) = sum_k  Q[128, k] (bf16, smem) . E[N_TILE, k]^T (bf16, smem)
// Queries sit on the MMA M dimension so that TMEM lane == query: an epilogue thread owns one query
// and scans its lane's columns against that query's running threshold -- no cross-thread traffic.
//
// Warp roles (224 threads, one CTA per SM, persistent over work items):
//   warp 0      TMA producer   : corpus (E) tile k-blocks into a deep smem ring (HBM stream)
//   warp 1      MMA issuer     : tcgen05.mma.kind::f16 into one of two TMEM accumulators; owns TMEM alloc
//   warps 2..5  epilogue       : tcgen05.ld -> threshold test -> append to smem candidate lists, lockstep prune
//   warp 6      TMA producer   : query (Q) k-blocks into a shallow ring (L2-resident, re-read per tile)
//
// Precision modes (operand planes are bf16, products are exact, accumulation fp32):
//   mode 0  Q0.E0                      bf16 corpus, single query plane (over-fetch + exact re-score follow)
//   mode 1  Q0.E0 + Q1.E0              bf16 corpus, query split hi+lo
//   mode 2  Q0.E0 + Q1.E0 + Q0.E1      fp32 corpus as hi+lo planes (drops only the lo.lo term, ~2^-18)
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "internal.h"
#include "ptx.cuh"
#include "sweep_epilogue.cuh"

namespace dewi {

namespace {

using namespace sweep;

constexpr int kThreads = 224;
constexpr int kQWarp = 6;  // query-ring producer (kept apart so the corpus ring can run ahead)
struct TcArgs {
  int n_rows;
  int n_tiles;
  int n_kb;
  int n_qb;
  int n_chunks;
  int n_items;
  int kc;
  int e_stages;
  int q_stages;
  float* part_s;
  int* part_i;
  const float* seed;  // optional admission thresholds from the sample pre-pass (see seed_threshold)
  int seed_stride, seed_off, n_queries;
  float* max_out;     // pre-pass mode: [item][128] maximum score per query, no candidate lists
  int max_groups;     // pre-pass mode: one maximum per 32-row group of every tile instead (SweepSeed::max_groups)
  int fp16;           // operand planes hold fp16 (fp32 corpus) instead of bf16
};

// QRES = 1 ("queries resident"; single query block, M = 64): all dim/64 query k-blocks are loaded ONCE and stay in
// shared memory (96 KB at dim 768) instead of being re-read from L2 with every corpus tile -- the corpus ring
// gives up one of its four stages for it (measured: three stages stream as fast as four, two do not).  L2 -> SM
// traffic per sweep drops from 1.5x to 1.0x the shard.
template <int MODE, int N_TILE, int Q_ROWS, int QRES>
__global__ void __launch_bounds__(kThreads, 1)
search_tc_kernel(const __grid_constant__ CUtensorMap map_e0, const __grid_constant__ CUtensorMap map_e1,
                 const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                 const TcArgs a) {
  using T = ModeTraits<MODE>;
  constexpr uint32_t kEPlaneBytes = N_TILE * 128;
  // Q_ROWS = 64: a batch of at most 64 queries runs M = 64 MMAs (half the tensor work and half the query
  // re-stream of M = 128).  The accumulator then lives in TMEM lanes 0-15 of each 32-lane quarter (rows
  // 16q .. 16q+15 -> lanes 32q .. 32q+15), which is exactly where query_lane() puts queries 0..63.
  constexpr uint32_t kQPlaneBytes = Q_ROWS * 128;
  constexpr uint32_t kEStageBytes = T::PE * kEPlaneBytes;
  constexpr uint32_t kQStageBytes = T::PQ * kQPlaneBytes;
  constexpr uint32_t kTmemCols = 2 * N_TILE;
  const uint32_t kIdesc = a.fp16 ? ptx::make_idesc_f16(Q_ROWS, N_TILE) : ptx::make_idesc_bf16(Q_ROWS, N_TILE);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B operand tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Two rings: corpus k-blocks come from HBM and need depth; query k-blocks come from L2 (the same
  // 12 blocks over and over) and need little.
  uint8_t* ring_e = smem;
  uint8_t* ring_q = ring_e + static_cast<size_t>(a.e_stages) * kEStageBytes;
  const int kl = a.kc + kPending;
  // candidate lists: [slot][column]; an M = 64 sweep has accumulator rows only in lanes 0-15 of each TMEM lane
  // quarter, so its lists keep 64 columns per slot (column = quarter * 16 + lane)
  constexpr int kListCols = (Q_ROWS == 64) ? 64 : kQueryBlock;
  float* list_s = reinterpret_cast<float*>(ring_q + static_cast<size_t>(a.q_stages) * kQStageBytes);
  int* list_i = reinterpret_cast<int*>(list_s + kl * kListCols);
  uint64_t* bar_e_full = reinterpret_cast<uint64_t*>(list_i + kl * kListCols);
  uint64_t* bar_e_empty = bar_e_full + kMaxStages;
  uint64_t* bar_q_full = bar_e_empty + kMaxStages;
  uint64_t* bar_q_empty = bar_q_full + kQStagesMax;
  uint64_t* bar_acc_full = bar_q_empty + kQStagesMax;
  uint64_t* bar_acc_empty = bar_acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

  // warp-uniform role index (the shuffle tells the compiler so); whole warps walk the role loops and
  // the TMA / MMA issue sites are predicated on elect.sync, which keeps their operands in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_e0);
    ptx::prefetch_tmap(&map_q0);
    if (T::PE > 1) ptx::prefetch_tmap(&map_e1);
    if (T::PQ > 1) ptx::prefetch_tmap(&map_q1);
    for (int s = 0; s < a.e_stages; ++s) {
      ptx::mbar_init(&bar_e_full[s], 1);
      ptx::mbar_init(&bar_e_empty[s], 1);
    }
    for (int s = 0; s < (QRES ? 1 : a.q_stages); ++s) {  // (resident queries: one barrier, q_stages = n_kb buffers)
      ptx::mbar_init(&bar_q_full[s], 1);
      ptx::mbar_init(&bar_q_empty[s], 1);
    }
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
    // ===================== TMA producer: corpus ring =====================
    {
      int se = 0;
      uint32_t pe = 0;
      // single query block: the corpus is streamed once, do not let it displace the queries in L2
      const uint64_t e_policy = a.n_qb > 1 ? ptx::kEvictNormal : ptx::kEvictFirst;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        int t0, t1;
        tile_range(item / a.n_qb, a.n_chunks, a.n_tiles, t0, t1);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < a.n_kb; ++kb) {
            ptx::mbar_wait(&bar_e_empty[se], pe ^ 1);
            if (ptx::elect_one()) {
              uint8_t* st = ring_e + static_cast<size_t>(se) * kEStageBytes;
              ptx::mbar_arrive_expect_tx(&bar_e_full[se], kEStageBytes);
              ptx::tma_load_2d(st, &map_e0, &bar_e_full[se], kb * kKBlock, t * N_TILE, e_policy);
              if (T::PE > 1)
                ptx::tma_load_2d(st + kEPlaneBytes, &map_e1, &bar_e_full[se], kb * kKBlock, t * N_TILE, e_policy);
            }
            __syncwarp();
            if (++se == a.e_stages) { se = 0; pe ^= 1; }
          }
        }
      }
    }
  } else if (warp == kQWarp) {
    // ===================== TMA producer: query ring =====================
    {
      int sq = 0;
      uint32_t pq = 0;
      if (QRES) {
        // the one query block of this launch: every k-block once, one barrier for all of them (q_stages == n_kb)
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&bar_q_full[0], static_cast<uint32_t>(a.n_kb) * kQStageBytes);
          for (int kb = 0; kb < a.n_kb; ++kb) {
            uint8_t* sqp = ring_q + static_cast<size_t>(kb) * kQStageBytes;
            ptx::tma_load_2d(sqp, &map_q0, &bar_q_full[0], kb * kKBlock, 0, ptx::kEvictFirst);
            if (T::PQ > 1) ptx::tma_load_2d(sqp + kQPlaneBytes, &map_q1, &bar_q_full[0], kb * kKBlock, 0, ptx::kEvictFirst);
          }
        }
        __syncwarp();
      }
      for (int item = blockIdx.x; !QRES && item < a.n_items; item += gridDim.x) {
        const int qb = item % a.n_qb;
        int t0, t1;
        tile_range(item / a.n_qb, a.n_chunks, a.n_tiles, t0, t1);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < a.n_kb; ++kb) {
            ptx::mbar_wait(&bar_q_empty[sq], pq ^ 1);
            if (ptx::elect_one()) {
              uint8_t* sqp = ring_q + static_cast<size_t>(sq) * kQStageBytes;
              ptx::mbar_arrive_expect_tx(&bar_q_full[sq], kQStageBytes);
              ptx::tma_load_2d(sqp, &map_q0, &bar_q_full[sq], kb * kKBlock, qb * kQueryBlock, ptx::kEvictLast);  // Q_ROWS-row box
              if (T::PQ > 1)
                ptx::tma_load_2d(sqp + kQPlaneBytes, &map_q1, &bar_q_full[sq], kb * kKBlock, qb * kQueryBlock,
                                 ptx::kEvictLast);
            }
            __syncwarp();
            if (++sq == a.q_stages) { sq = 0; pq ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
    {
      int se = 0, sq = 0;
      uint32_t pe = 0, pq = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t ring_e_addr = ptx::smem_u32(ring_e), ring_q_addr = ptx::smem_u32(ring_q);
      if (QRES) ptx::mbar_wait(&bar_q_full[0], 0);  // the resident query block has landed (once per launch)
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        int t0, t1;
        tile_range(item / a.n_qb, a.n_chunks, a.n_tiles, t0, t1);
        for (int t = t0; t < t1; ++t) {
          ptx::mbar_wait(&bar_acc_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * N_TILE);
          for (int kb = 0; kb < a.n_kb; ++kb) {
            if (QRES) sq = kb;
            else ptx::mbar_wait(&bar_q_full[sq], pq);
            ptx::mbar_wait(&bar_e_full[se], pe);  // TMA bytes have landed
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint32_t ste = ring_e_addr + static_cast<uint32_t>(se) * kEStageBytes;
              const uint32_t stq = ring_q_addr + static_cast<uint32_t>(sq) * kQStageBytes;
              const uint64_t de0 = ptx::make_desc_sw128(ste);
              const uint64_t de1 = ptx::make_desc_sw128(ste + kEPlaneBytes);
              const uint64_t dq0 = ptx::make_desc_sw128(stq);
              const uint64_t dq1 = ptx::make_desc_sw128(stq + kQPlaneBytes);
#pragma unroll
              for (int k = 0; k < kKBlock / 16; ++k) {
                const uint64_t adv = static_cast<uint64_t>(k * 2);  // 16 bf16 = 32 B = 2 x 16 B units
                ptx::mma_bf16_ss(tmem_d, dq0 + adv, de0 + adv, kIdesc, (kb | k) != 0 ? 1u : 0u);
                if (MODE >= 1) ptx::mma_bf16_ss(tmem_d, dq1 + adv, de0 + adv, kIdesc, 1u);
                if (MODE == 2) ptx::mma_bf16_ss(tmem_d, dq0 + adv, de1 + adv, kIdesc, 1u);
              }
              ptx::mma_commit(&bar_e_empty[se]);  // smem slots reusable once these MMAs retire
              if (!QRES) ptx::mma_commit(&bar_q_empty[sq]);
              if (kb == a.n_kb - 1) ptx::mma_commit(&bar_acc_full[acc]);
            }
            __syncwarp();
            if (++se == a.e_stages) { se = 0; pe ^= 1; }
            if (!QRES && ++sq == a.q_stages) { sq = 0; pq ^= 1; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> per-query candidate lists =====================
    const int quarter = warp & 3;             // TMEM lane quarter this warp may read
    const int qlane = quarter * 32 + lane;    // query within the 128-query block == TMEM lane
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int kc = a.kc;
    LaneList l;
    const int col = (Q_ROWS == 64) ? quarter * 16 + (lane & 15) : qlane;  // (lanes 16-31 of an M = 64 sweep never append)
    l.s = list_s + col;
    l.i = list_i + col;
    l.ws = list_s + ((Q_ROWS == 64) ? quarter * 16 : quarter * 32);
    l.wi = list_i + ((Q_ROWS == 64) ? quarter * 16 : quarter * 32);
    l.stride = kListCols;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      int t0, t1;
      tile_range(item / a.n_qb, a.n_chunks, a.n_tiles, t0, t1);
      l.cnt = 0;
      l.thr = seed_threshold(a.seed, a.seed_stride, a.seed_off, a.n_queries, item % a.n_qb, qlane);
      if (Q_ROWS == 64 && lane >= 16) l.thr = INFINITY;  // these TMEM lanes hold no accumulator row
      float best = -INFINITY;
      for (int t = t0; t < t1; ++t) {
        ptx::mbar_wait(&bar_acc_full[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t tcol = tmem_lane + static_cast<uint32_t>(acc * N_TILE);
        const int row_base = t * N_TILE;
        if (a.max_out && a.max_groups)
          max_groups_tile<N_TILE>(a.max_out + static_cast<size_t>(item % a.n_qb) * kQueryBlock + qlane,
                                  static_cast<size_t>(a.n_qb) * kQueryBlock, t, tcol, row_base, a.n_rows);
        else if (a.max_out) best = max_tile<N_TILE>(best, tcol, row_base, a.n_rows);
        else scan_tile<N_TILE>(l, kc, tcol, row_base, a.n_rows);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar_acc_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (a.max_out) { if (!a.max_groups) a.max_out[static_cast<size_t>(item) * kQueryBlock + qlane] = best; }
      else flush_item(l, kc, a.part_s + static_cast<size_t>(item) * kc * kQueryBlock + qlane,
                      a.part_i + static_cast<size_t>(item) * kc * kQueryBlock + qlane);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side ---------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

size_t e_stage_bytes(int mode, int n_tile) { return static_cast<size_t>((mode == 2) ? 2 : 1) * n_tile * 128; }
size_t q_stage_bytes(int mode, int q_rows) { return static_cast<size_t>((mode == 0) ? 1 : 2) * q_rows * 128; }
size_t fixed_bytes(int kc, int q_rows) {
  const int list_cols = (q_rows == 64) ? 64 : kQueryBlock;  // M = 64 sweeps keep 64 list columns
  return static_cast<size_t>(kc + kPending) * list_cols * 8 + (2 * kMaxStages + 2 * kQStagesMax + 4) * 8 + 16 +
         1024 /*alignment slack*/;
}

template <int MODE, int N_TILE, int Q_ROWS, int QRES = 0>
int launch_one(const TcPlan& plan, const CUtensorMap& e0, const CUtensorMap& e1, const CUtensorMap& q0,
               const CUtensorMap& q1, const TcArgs& args, cudaStream_t stream) {
  auto kern = search_tc_kernel<MODE, N_TILE, Q_ROWS, QRES>;
  DEWI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.smem_bytes)));
  kern<<<plan.grid, kThreads, plan.smem_bytes, stream>>>(e0, e1, q0, q1, args);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int tc_supported(int dim, int64_t n_rows) {
  return dim % kKBlock == 0 && dim >= kKBlock && dim <= 8192 && n_rows >= 1 && n_rows < (int64_t(1) << 31);
}

int tc_make_plan(int mode, int dim, int64_t n_rows, int n_qb, int kc, int sm_count, TcPlan* plan, int force_chunks,
                 int q_rows, int swap_queries) {
  if (!tc_supported(dim, n_rows)) return fail("tcgen05 sweep needs dim % 64 == 0 and rows < 2^31");
  if (swap_queries > 0 && mode == 0 && n_qb == 1 &&
      tcr_make_plan(dim, n_rows, swap_queries, kc, sm_count, plan, force_chunks) == 0)
    return 0;
  plan->rows_on_m = 0;
  const size_t smem_max = 227 * 1024;
  int n_tile = (mode == 2) ? 128 : 256;
  const size_t fixed = fixed_bytes(kc, q_rows);
  int q_stages = 2;
  int stages = 0;
  // Single query block of at most 64 queries: keep all its k-blocks resident if the corpus ring still gets 3 stages.
  bool q_res = n_qb == 1 && q_rows == 64 && mode != 2;
  q_res = q_res && env_int("DEWI_TC_QRES", 1) != 0;  // experiments
  if (q_res) {
    const size_t used = fixed + static_cast<size_t>(dim / kKBlock) * q_stage_bytes(mode, q_rows);
    const int st = used < smem_max ? static_cast<int>((smem_max - used) / e_stage_bytes(mode, n_tile)) : 0;
    if (st >= 3) {
      q_stages = dim / kKBlock;
      stages = st;
    } else {
      q_res = false;
    }
  }
  for (; !q_res;) {
    const size_t used = fixed + q_stages * q_stage_bytes(mode, q_rows);
    stages = used < smem_max ? static_cast<int>((smem_max - used) / e_stage_bytes(mode, n_tile)) : 0;
    if (stages >= 3 || n_tile == 128) break;
    n_tile = 128;
  }
  if (stages < 2) return fail("candidate list capacity too large for the tcgen05 sweep's shared memory");
  // M = 64 sweeps of one or two planes exist for 256-row tiles only (tc_launch): lists long enough to force 128-row
  // tiles there (k above ~105 with at most 64 queries) have no plan -- the caller takes the CUDA-core sweep
  if (q_rows == 64 && mode != 2 && n_tile != 256)
    return fail("candidate list capacity too large for the M = 64 tcgen05 sweep's shared memory");
  // four 32 KB stages in flight per SM stream fastest (12.5M rows, B = 128: 3.37 / 3.04 / 3.21 ms with 3 / 4 / 5 stages):
  // a deeper ring only adds concurrent DRAM streams
  stages = std::min(stages, 4);
  if (env_set("DEWI_TC_STAGES")) stages = std::min(stages, std::max(2, env_int("DEWI_TC_STAGES", stages)));  // experiments
  const int64_t n_tiles = ceil_div(n_rows, n_tile);
  // Work items = chunks x query blocks, dealt round-robin to `grid` persistent CTAs.  Choose the
  // number of chunks so the item count is a multiple of the grid (equal work per CTA).
  int grid = static_cast<int>(std::min<int64_t>(sm_count, n_tiles * n_qb));
  int64_t want = ceil_div(static_cast<int64_t>(grid), n_qb);  // >= one item per CTA
  // make chunks * n_qb a multiple of grid when the corpus is long enough
  int64_t chunks = want;
  for (int64_t c = want; c <= want + grid && c <= n_tiles; ++c) {
    if ((c * n_qb) % grid == 0) { chunks = c; break; }
  }
  if (force_chunks > 0) chunks = force_chunks;
  chunks = std::max<int64_t>(1, std::min<int64_t>(chunks, n_tiles));
  plan->mode = mode;
  plan->q_rows = q_rows;
  plan->n_tile = n_tile;
  plan->n_stages = stages;
  plan->q_stages = q_stages;
  plan->q_resident = q_res ? 1 : 0;
  plan->n_chunks = static_cast<int>(chunks);
  plan->grid = static_cast<int>(std::min<int64_t>(grid, chunks * n_qb));
  plan->smem_bytes = fixed + static_cast<size_t>(stages) * e_stage_bytes(mode, n_tile) + q_stages * q_stage_bytes(mode, q_rows);
  return 0;
}

int tc_encode_rows_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kKBlock), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
  return 0;
}

int tc_launch(const TcPlan& plan, const CUtensorMap& e0, const CUtensorMap& e1, const CUtensorMap& q0,
              const CUtensorMap& q1, int64_t n_rows, int dim, int n_qb, int kc, float* part_s, int* part_i,
              const SweepSeed& seed, cudaStream_t stream, int fp16_planes, int n_queries, const SweepBlend* blend) {
  if (plan.rows_on_m) return tcr_launch(plan, e0, q0, n_rows, dim, n_queries, kc, part_s, part_i, seed, stream, fp16_planes, blend);
  if (blend && blend->enabled) return fail("the full-corpus blend needs the rows-on-M or the CUDA-core sweep");
  TcArgs a;
  a.fp16 = fp16_planes;
  a.n_rows = static_cast<int>(n_rows);
  a.n_tiles = static_cast<int>(ceil_div(n_rows, plan.n_tile));
  a.n_kb = dim / kKBlock;
  a.n_qb = n_qb;
  a.n_chunks = plan.n_chunks;
  a.n_items = plan.n_chunks * n_qb;
  a.kc = kc;
  a.e_stages = plan.n_stages;
  a.q_stages = plan.q_stages;
  a.part_s = part_s;
  a.part_i = part_i;
  a.seed = seed.values;
  a.seed_stride = seed.stride;
  a.seed_off = seed.off;
  a.n_queries = seed.n_queries;
  a.max_out = seed.max_out;
  a.max_groups = seed.max_groups;
  if (plan.q_rows == 64) {
    if (plan.q_resident) {
      if (plan.mode == 0 && plan.n_tile == 256) return launch_one<0, 256, 64, 1>(plan, e0, e1, q0, q1, a, stream);
      if (plan.mode == 1 && plan.n_tile == 256) return launch_one<1, 256, 64, 1>(plan, e0, e1, q0, q1, a, stream);
      return fail("unsupported tcgen05 sweep configuration (resident queries)");
    }
    if (plan.mode == 0 && plan.n_tile == 256) return launch_one<0, 256, 64>(plan, e0, e1, q0, q1, a, stream);
    if (plan.mode == 1 && plan.n_tile == 256) return launch_one<1, 256, 64>(plan, e0, e1, q0, q1, a, stream);
    if (plan.mode == 2 && plan.n_tile == 128) return launch_one<2, 128, 64>(plan, e0, e1, q0, q1, a, stream);
    return fail("unsupported tcgen05 sweep configuration (M = 64)");
  }
  if (plan.mode == 0 && plan.n_tile == 256) return launch_one<0, 256, 128>(plan, e0, e1, q0, q1, a, stream);
  if (plan.mode == 0 && plan.n_tile == 128) return launch_one<0, 128, 128>(plan, e0, e1, q0, q1, a, stream);
  if (plan.mode == 1 && plan.n_tile == 256) return launch_one<1, 256, 128>(plan, e0, e1, q0, q1, a, stream);
  if (plan.mode == 1 && plan.n_tile == 128) return launch_one<1, 128, 128>(plan, e0, e1, q0, q1, a, stream);
  if (plan.mode == 2 && plan.n_tile == 128) return launch_one<2, 128, 128>(plan, e0, e1, q0, q1, a, stream);
  return fail("unsupported tcgen05 sweep configuration");
}

}  // namespace dewi
