// K1b -- the similarity sweep for large query batches on CTA PAIRS (tcgen05 cta_group::2).
//
// Same contract as search_tc.cu (reference src/dewi/backends.py:431-447 fused on chip), for B > 128
// where the contraction is tensor-bound and the ceiling of the 1-CTA kernel is L2 -> shared-memory
// operand traffic (48 KB per 128x256x64 MMA block).  Two CTAs of a cluster issue ONE
// tcgen05.mma.cta_group::2 with M = 256 (two 128-query blocks, one per CTA) x N = 256 corpus rows:
// each CTA stages its own query block and only HALF of the corpus tile (128 rows), the tensor core
// reads the other half from the peer's shared memory.  Per CTA that is 32 KB of operands per
// 128x256x64 block of work -- 1.5x less L2 traffic per flop -- and the corpus tile is fetched once
// per pair instead of once per CTA.
//
// Roles per CTA (192 threads): w0 TMA producer (own query k-block + own half of the corpus k-block into
// one ring stage), w1 MMA issuer (leader CTA only) + TMEM allocation, w2-5 epilogue over the CTA's own
// 128 TMEM lanes (identical to the 1-CTA kernel: sweep_epilogue.cuh).  Role loops are walked by whole
// warps with the issue sites predicated on elect.sync, so operands stay in uniform registers.
// Barriers: "full" barriers live in the leader (both CTAs' TMA transactions are accounted there);
// "empty" and "accumulator full" are signalled in both CTAs by a multicast tcgen05.commit;
// "accumulator empty" collects 8 epilogue-warp arrivals (4 local + 4 remote) in the leader.
#include <algorithm>
#include <cstdlib>

#include "internal.h"
#include "ptx.cuh"
#include "sweep_epilogue.cuh"
#include "join_epilogue.cuh"

namespace dewi {
namespace {

using namespace sweep;

constexpr int kThreads = 192;            // 2 role warps + 4 epilogue warps (top-k epilogue)
// The join epilogue is the longer one (row and column direction per tile) and sets the pace of the MMAs when a
// single warp per TMEM lane quarter has to walk all 256 columns: it runs TWO warps per quarter, 128 columns each.
constexpr int kJoinEpiWarps = 8;
constexpr int kJoinThreads = 64 + 32 * kJoinEpiWarps;
constexpr int kNTile = 256;           // corpus rows per MMA (N); each CTA stages kNTile / 2
constexpr int kHalfRows = kNTile / 2;

constexpr int kSyncEvery = 16;          // tiles between two rendezvous (2 .. 32 measured within 3 % of each other)
constexpr long long kSyncSpin = 60000;   // cycles (~30-45 us)

// The tiles of one work item: `len` offsets starting at o0 from tile `tbase` (wrapping at n_tiles), walked in
// the rotated order o0 + (s + rot) % len.  Symmetric join: the clusters running concurrently hold consecutive
// row blocks I0 + cluster_id whose tile ranges are shifted by one tile each; rotating cluster k's walk by
// (n_clusters - 1 - k) makes all of them read tile I0 + n_clusters - 1 + o0 + s at step s -- one DRAM read
// feeds every cluster through L2, exactly as in the plain sweep -- and leaves each cluster its private
// leftovers (n_clusters - 1 tiles, the diagonal among them) for the end of the item.
struct ItemSpan {
  int qpair, o0, len, rot, tbase;
};
__device__ __forceinline__ ItemSpan item_span(const Tc2Args& a, int item, int cluster_id, int n_clusters) {
  ItemSpan sp;
  sp.qpair = item % a.n_qpairs;
  const int chunk = item / a.n_qpairs;
  int o1;
  if (!a.sym) {
    tile_range(chunk, a.n_chunks, a.n_tiles, sp.o0, o1);
    sp.len = o1 - sp.o0;
    sp.rot = 0;
    sp.tbase = 0;
    return sp;
  }
  const int ig = a.blk0 + sp.qpair;
  const int T = a.n_tiles;
  // block pairs at distance d <= (T-1)/2 belong to the lower block; for even T the distance T/2 is met
  // from both sides and belongs to the block in the first half
  const int L = 1 + (T - 1) / 2 + ((((T & 1) == 0) && ig < T / 2) ? 1 : 0);
  tile_range(chunk, a.n_chunks, L, sp.o0, o1);
  sp.len = o1 - sp.o0;
  sp.rot = (a.rotate && sp.len > 0) ? (n_clusters - 1 - cluster_id) % sp.len : 0;
  sp.tbase = ig;
  return sp;
}
__device__ __forceinline__ int span_offset(const ItemSpan& sp, int s) {
  const int o = sp.o0 + s + sp.rot;
  return o >= sp.o0 + sp.len ? o - sp.len : o;
}
__device__ __forceinline__ int tile_at(const Tc2Args& a, int tbase, int o) {
  const int t = tbase + o;
  return t >= a.n_tiles ? t - a.n_tiles : t;
}


// ARES = 1 ("A resident", join only): the row block on the M side -- 128 rows x dim per CTA -- is loaded ONCE per
// work item into its own shared-memory buffer instead of being re-streamed from L2 with every corpus tile; the
// ring then carries only the corpus halves.  That halves the L2 -> SM operand traffic (at D = 512 the re-streamed
// row block is half of the ~10 TB/s the sweep pulls through the crossbar), which is what the power-capped chip
// converts into clock.  Needs 128 * dim * 2 * planes bytes next to a ring of >= 4 stages: bf16 planes up to dim 640.
template <int MODE, int EPI, int ARES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((EPI == 1) ? kJoinThreads : kThreads, 1)
search_tc2_kernel(const __grid_constant__ CUtensorMap map_e0, const __grid_constant__ CUtensorMap map_e1,
                  const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                  const Tc2Args a) {
  using T = ModeTraits<MODE>;
  constexpr uint32_t kEPlaneBytes = kHalfRows * 128;   // this CTA's half of a corpus k-block
  constexpr uint32_t kQPlaneBytes = kQueryBlock * 128;
  constexpr uint32_t kEBytes = T::PE * kEPlaneBytes;
  constexpr uint32_t kQBytes = T::PQ * kQPlaneBytes;
  constexpr uint32_t kStageBytes = kEBytes + (ARES ? 0u : kQBytes);   // [corpus half planes | query planes]
  constexpr uint32_t kTmemCols = 2 * kNTile;
  constexpr int kEpiWarps = (EPI == EPI_JOIN) ? kJoinEpiWarps : 4;
  constexpr int kColsPerWarp = kNTile * 4 / kEpiWarps;   // columns of a tile one epilogue warp walks
  const uint32_t kIdesc = a.fp16 ? ptx::make_idesc_f16(2 * kQueryBlock, kNTile) : ptx::make_idesc_bf16(2 * kQueryBlock, kNTile);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // ONE ring: both operands of a k-block arrive from L2 at the same rate here, and one full / one empty
  // barrier per k-block keeps the single-threaded MMA issue loop short.
  uint8_t* a_buf = smem;                                   // ARES: [k-block][plane] row-block tiles, n_kb * kQBytes
  uint8_t* ring = smem + (ARES ? static_cast<size_t>(a.n_kb) * kQBytes : 0);
  const int kl = a.kc + kPending;
  float* list_s = reinterpret_cast<float*>(ring + static_cast<size_t>(a.e_stages) * kStageBytes);
  int* list_i = reinterpret_cast<int*>(list_s + kl * kQueryBlock);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(list_i + kl * kQueryBlock);
  uint64_t* bar_empty = bar_full + kMaxStages;
  uint64_t* bar_acc_full = bar_empty + kMaxStages;
  uint64_t* bar_acc_empty = bar_acc_full + 2;
  uint64_t* bar_a_full = bar_acc_empty + 2;    // ARES: the item's row block has landed (leader; both CTAs' bytes)
  uint64_t* bar_a_empty = bar_a_full + 1;      // ARES: the item's last MMA has read it (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a_empty + 1);

  // warp-uniform role index (the shuffle tells the compiler so)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();   // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_e0);
    ptx::prefetch_tmap(&map_q0);
    if (T::PE > 1) ptx::prefetch_tmap(&map_e1);
    if (T::PQ > 1) ptx::prefetch_tmap(&map_q1);
    for (int s = 0; s < a.e_stages; ++s) {
      ptx::mbar_init(&bar_full[s], 2);   // one producer arrival per CTA of the pair
      ptx::mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bar_acc_full[b], 1);
      ptx::mbar_init(&bar_acc_empty[b], 2 * kEpiWarps);  // every epilogue warp of both CTAs
    }
    ptx::mbar_init(bar_a_full, 2);
    ptx::mbar_init(bar_a_empty, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, kTmemCols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();   // barrier inits and TMEM allocation visible to the peer before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: this CTA's query k-block + its half of the corpus k-block =====================
    int st = 0;
    uint32_t ph = 0;
    // a corpus tile is read again by the clusters working on the other query-block pairs of the same
    // chunk: leave it in L2 at normal priority unless this is the only pair
    const uint64_t e_policy = a.n_qpairs > 1 ? ptx::kEvictNormal : ptx::kEvictFirst;
    uint32_t a_ph = 0;
    for (int item = a.item0 + cluster_id; item < a.item1; item += n_clusters) {
      const ItemSpan sp = item_span(a, item, cluster_id, n_clusters);
      const int qrow = (sp.qpair * 2 + static_cast<int>(rank)) * kQueryBlock + (a.sym ? static_cast<int>(a.a_offset) : 0);
      if (ARES && sp.len > 0) {
        ptx::mbar_wait(bar_a_empty, a_ph ^ 1);   // the previous item's MMAs are done with the buffer
        if (ptx::elect_one()) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(bar_a_full, 2 * static_cast<uint32_t>(a.n_kb) * kQBytes);
          else ptx::mbar_arrive_leader(bar_a_full);
          for (int kb = 0; kb < a.n_kb; ++kb) {
            uint8_t* dst = a_buf + static_cast<size_t>(kb) * kQBytes;
            ptx::tma_load_2d_pair(dst, &map_q0, bar_a_full, kb * kKBlock, qrow, ptx::kEvictFirst);
            if (T::PQ > 1) ptx::tma_load_2d_pair(dst + kQPlaneBytes, &map_q1, bar_a_full, kb * kKBlock, qrow, ptx::kEvictFirst);
          }
        }
        __syncwarp();
        a_ph ^= 1;
      }
      // rendezvous group of this item: the items of the same chunk that run in the same round
      int sync_expect = 0;
      unsigned int* sync_row = nullptr;
      if (a.sync_cnt && rank == 0) {
        // (rounds are counted from the launch's first item; item0 is a multiple of n_qpairs, so first >= item0)
        const int chunk = item / a.n_qpairs, round = (item - a.item0) / n_clusters;
        const int first = chunk * a.n_qpairs, last = first + a.n_qpairs;                 // the chunk's items
        const int lo = max(first, a.item0 + round * n_clusters), hi = min(min(last, a.item0 + (round + 1) * n_clusters), a.item1);
        sync_expect = hi - lo;
        sync_row = a.sync_cnt + (static_cast<size_t>(chunk) * 2 + (round - (first - a.item0) / n_clusters)) * a.sync_blocks;
      }
      for (int step = 0; step < sp.len; ++step) {
        if (sync_expect > 1 && (step % a.sync_every) == 0 && step / a.sync_every < a.sync_blocks) {
          if (ptx::elect_one()) {
            unsigned int* c = sync_row + step / a.sync_every;
            atomicAdd(c, 1u);
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile unsigned int*>(c) < static_cast<unsigned int>(sync_expect) &&
                   clock64() - t0 < kSyncSpin) {
            }
          }
          __syncwarp();
        }
        const int row0 = tile_at(a, sp.tbase, span_offset(sp, step)) * kNTile + static_cast<int>(rank) * kHalfRows;
        for (int kb = 0; kb < a.n_kb; ++kb) {
          ptx::mbar_wait(&bar_empty[st], ph ^ 1);
          if (ptx::elect_one()) {
            uint8_t* stg = ring + static_cast<size_t>(st) * kStageBytes;
            if (rank == 0) ptx::mbar_arrive_expect_tx(&bar_full[st], 2 * kStageBytes);
            else ptx::mbar_arrive_leader(&bar_full[st]);
            ptx::tma_load_2d_pair(stg, &map_e0, &bar_full[st], kb * kKBlock, row0, e_policy);
            if (T::PE > 1) ptx::tma_load_2d_pair(stg + kEPlaneBytes, &map_e1, &bar_full[st], kb * kKBlock, row0, e_policy);
            if (!ARES) {
              ptx::tma_load_2d_pair(stg + kEBytes, &map_q0, &bar_full[st], kb * kKBlock, qrow, ptx::kEvictLast);
              if (T::PQ > 1)
                ptx::tma_load_2d_pair(stg + kEBytes + kQPlaneBytes, &map_q1, &bar_full[st], kb * kKBlock, qrow, ptx::kEvictLast);
            }
          }
          __syncwarp();
          if (++st == a.e_stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; the whole warp walks the loop, one elected lane issues) ==========
    if (rank == 0) {
      int st = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t ring_addr = ptx::smem_u32(ring);
      const uint32_t a_addr = ptx::smem_u32(a_buf);
      uint32_t a_full_ph = 0;
      for (int item = a.item0 + cluster_id; item < a.item1; item += n_clusters) {
        const ItemSpan sp = item_span(a, item, cluster_id, n_clusters);
        if (ARES && sp.len > 0) {
          ptx::mbar_wait(bar_a_full, a_full_ph);   // the item's row block (both CTAs' halves of M) has landed
          ptx::tc_fence_after();
          a_full_ph ^= 1;
        }
        for (int step = 0; step < sp.len; ++step) {
          ptx::mbar_wait(&bar_acc_empty[acc], acc_phase ^ 1);  // both CTAs' epilogues drained this accumulator
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kNTile);
          for (int kb = 0; kb < a.n_kb; ++kb) {
            ptx::mbar_wait(&bar_full[st], ph);  // both CTAs' TMA bytes have landed
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint32_t stg = ring_addr + static_cast<uint32_t>(st) * kStageBytes;
              const uint32_t qa = ARES ? a_addr + static_cast<uint32_t>(kb) * kQBytes : stg + kEBytes;
              const uint64_t de0 = ptx::make_desc_sw128(stg);
              const uint64_t de1 = ptx::make_desc_sw128(stg + kEPlaneBytes);
              const uint64_t dq0 = ptx::make_desc_sw128(qa);
              const uint64_t dq1 = ptx::make_desc_sw128(qa + kQPlaneBytes);
#pragma unroll
              for (int k = 0; k < kKBlock / 16; ++k) {
                const uint64_t adv = static_cast<uint64_t>(k * 2);  // 16 bf16 = 32 B = 2 x 16 B units
                ptx::mma_bf16_ss_pair(tmem_d, dq0 + adv, de0 + adv, kIdesc, (kb | k) != 0 ? 1u : 0u);
                if (MODE >= 1) ptx::mma_bf16_ss_pair(tmem_d, dq1 + adv, de0 + adv, kIdesc, 1u);
                if (MODE == 2) ptx::mma_bf16_ss_pair(tmem_d, dq0 + adv, de1 + adv, kIdesc, 1u);
              }
              ptx::mma_commit_pair(&bar_empty[st], 3);  // frees the slot in both CTAs
              if (kb == a.n_kb - 1) {
                ptx::mma_commit_pair(&bar_acc_full[acc], 3);
                if (ARES && step == sp.len - 1) ptx::mma_commit_pair(bar_a_empty, 3);  // row block free for the next item
              }
            }
            __syncwarp();
            if (++st == a.e_stages) { st = 0; ph ^= 1; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue: this CTA's 128 queries =====================
    const int quarter = warp & 3;
    const int qlane = quarter * 32 + lane;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int kc = a.kc;
    const int n_qb = a.n_qpairs * 2;
    LaneList l;
    l.s = list_s + qlane;
    l.i = list_i + qlane;
    l.ws = list_s + quarter * 32;
    l.wi = list_i + quarter * 32;
    l.stride = kQueryBlock;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = a.item0 + cluster_id; item < a.item1; item += n_clusters) {
      const int chunk = item / a.n_qpairs;
      const ItemSpan sp = item_span(a, item, cluster_id, n_clusters);
      const int qb = sp.qpair * 2 + static_cast<int>(rank);
      l.cnt = 0;
      l.thr = seed_threshold(a.seed, a.seed_stride, a.seed_off, a.n_queries, qb, qlane);
      float best = -INFINITY;
      JoinRow jr = {-INFINITY, -1, 0};
      const int i_row = qb * kQueryBlock + qlane;
      if (EPI == EPI_JOIN && i_row < a.m_rows) {
        // Start from the best this row has met so far (other items of the same rows, and -- symmetric join -- the
        // column direction of other blocks): the warp then takes the "new best" path only for real improvements
        // instead of ~32 ln(groups) times per item.
        jr.best = best_sim(__ldcg(a.row_best + (a.sym ? i_row + a.a_offset : i_row)));
      }
      const int col0 = ((warp - 2) >> 2) * kColsPerWarp;   // this warp's first column inside a tile
      unsigned long long gate_raw[kColsPerWarp / 32];  // symmetric join: current bests of the NEXT tile's columns
      float* col_cache = list_s + (warp - 2) * kColsPerWarp;  // the join keeps no candidate lists: their memory is free
      if (EPI == EPI_JOIN && a.sym && sp.len > 0)
        gate_prefetch<kColsPerWarp>(gate_raw, a, tile_at(a, sp.tbase, span_offset(sp, 0)) * kNTile + col0, lane);
      for (int step = 0; step < sp.len; ++step) {
        const int o = span_offset(sp, step);
        const int t = tile_at(a, sp.tbase, o);
        float lbv = INFINITY;
        if (EPI == EPI_JOIN && a.sym) {
          if (o > 0) lbv = gate_reduce<kColsPerWarp>(gate_raw, lane, col_cache);  // (the diagonal tile has no column direction)
          if (step + 1 < sp.len)
            gate_prefetch<kColsPerWarp>(gate_raw, a, tile_at(a, sp.tbase, span_offset(sp, step + 1)) * kNTile + col0, lane);
        }
        ptx::mbar_wait(&bar_acc_full[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t tcol = tmem_lane + static_cast<uint32_t>(acc * kNTile);
        if (EPI == EPI_TOPK && a.max_out && a.max_groups)
          max_groups_tile<kNTile>(a.max_out + static_cast<size_t>(qb) * kQueryBlock + qlane, static_cast<size_t>(n_qb) * kQueryBlock, t,
                                  tcol, t * kNTile, a.n_rows);
        else if (EPI == EPI_TOPK && a.max_out) best = max_tile<kNTile>(best, tcol, t * kNTile, a.n_rows);
        else if (EPI == EPI_TOPK) scan_tile<kNTile>(l, kc, tcol, t * kNTile, a.n_rows);
        else if (a.sym)
          join_scan_tile_sym<kColsPerWarp>(jr, a, i_row + a.a_offset, lane, tcol + col0, t * kNTile + col0, o == 0, lbv, col_cache);
        else join_scan_tile<kColsPerWarp>(jr, a, i_row, tcol + col0, t * kNTile + col0);
        if (EPI == EPI_JOIN && (step & 31) == 31 && jr.best_j >= 0 && i_row < a.m_rows) {
          // publish the row's running best every 32 tiles, not only at the end of a (possibly very long) item: the
          // column-direction gates of the OTHER blocks read row_best, and the sooner it is tight the fewer groups pass
          atomicMax(&a.row_best[a.sym ? i_row + a.a_offset : i_row], pack_best(jr.best, jr.best_j));
          jr.best_j = -1;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_leader(&bar_acc_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (EPI == EPI_TOPK && a.max_out) {
        if (!a.max_groups) a.max_out[(static_cast<size_t>(chunk) * n_qb + qb) * kQueryBlock + qlane] = best;
      } else if (EPI == EPI_TOPK) {
        const size_t slot = (static_cast<size_t>(chunk) * n_qb + qb) * kc * kQueryBlock + qlane;
        flush_item(l, kc, a.part_s + slot, a.part_i + slot);
      } else if (i_row < a.m_rows) {
        const long long i_out = a.sym ? i_row + a.a_offset : i_row;
        if (jr.best_j >= 0) atomicMax(&a.row_best[i_out], pack_best(jr.best, jr.best_j));
        if (jr.count) atomicAdd(&a.row_count[i_out], jr.count);
      }
    }
  }

  // Neither CTA may leave while the pair's MMAs can still read its shared memory or while remote
  // arrivals are in flight.
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

size_t stage_bytes(int mode, bool a_resident = false) {
  const int pe = (mode == 2) ? 2 : 1, pq = (mode == 0) ? 1 : 2;
  return static_cast<size_t>(pe) * kHalfRows * 128 + (a_resident ? 0 : static_cast<size_t>(pq) * kQueryBlock * 128);
}
size_t a_resident_bytes(int mode, int dim) {
  const int pq = (mode == 0) ? 1 : 2;
  return static_cast<size_t>(dim / kKBlock) * pq * kQueryBlock * 128;
}
size_t fixed_bytes(int kc) {
  return static_cast<size_t>(kc + kPending) * kQueryBlock * 8 + (2 * kMaxStages + 6) * 8 + 16 + 1024 /*alignment slack*/;
}

template <int MODE, int EPI, int ARES = 0>
int launch_one(const Tc2Plan& plan, const CUtensorMap& e0, const CUtensorMap& e1, const CUtensorMap& q0,
               const CUtensorMap& q1, const Tc2Args& args, cudaStream_t stream) {
  auto kern = search_tc2_kernel<MODE, EPI, ARES>;
  DEWI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.smem_bytes)));
  kern<<<plan.grid, EPI == EPI_JOIN ? kJoinThreads : kThreads, plan.smem_bytes, stream>>>(e0, e1, q0, q1, args);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int tc2_make_plan(int mode, int dim, int64_t n_rows, int n_qb, int kc, int sm_count, Tc2Plan* plan, int force_chunks,
                  int64_t tiles_per_item, int a_resident, int stage_first) {
  if (!tc_supported(dim, n_rows)) return fail("tcgen05 sweep needs dim % 64 == 0 and rows < 2^31");
  if (n_qb < 2 || (n_qb & 1)) return fail("the CTA-pair sweep needs an even number of query blocks");
  const size_t smem_max = 227 * 1024;
  const size_t fixed = fixed_bytes(kc) + (a_resident ? a_resident_bytes(mode, dim) : 0);
  int stages = fixed < smem_max ? static_cast<int>((smem_max - fixed) / stage_bytes(mode, a_resident != 0)) : 0;
  // The search streams the corpus from HBM: four 32 KB stages per CTA are the sweet spot (12.5M rows: B = 256 sweep 4.72 /
  // 4.19 / 4.26 ms with 3 / 4 / 5 stages, B = 4096 60.6 / 59.5 / 59.6 ms) -- a deeper ring only adds concurrent DRAM
  // streams.  The join's operands come from L2 and keep the deeper ring.
  if (!a_resident && kc > 0) stages = std::min(stages, 4);
  if (env_set("DEWI_TC2_STAGES")) stages = std::min(stages, std::max(2, env_int("DEWI_TC2_STAGES", stages)));
  if (stages < 2) return fail("candidate list capacity too large for the CTA-pair sweep's shared memory");
  stages = std::min(stages, kMaxStages);
  // tiles one query-block pair meets: the whole corpus, or the circulant half of a symmetric self-join
  const int64_t n_tiles = tiles_per_item > 0 ? tiles_per_item : ceil_div(n_rows, kNTile);
  const int n_qpairs = n_qb / 2;
  const int max_clusters = std::max(1, sm_count / 2);
  const int clusters = static_cast<int>(std::min<int64_t>(max_clusters, n_tiles * n_qpairs));
  // chunks x query-pairs work items, dealt round-robin to the clusters: choose the chunk count so that
  // the item count is a multiple of the cluster count (equal work) when the corpus is long enough
  const int64_t want = ceil_div(static_cast<int64_t>(clusters), n_qpairs);
  int64_t chunks = want;
  for (int64_t c = want; c <= want + clusters && c <= n_tiles; ++c) {
    if ((c * n_qpairs) % clusters == 0) { chunks = c; break; }
  }
  // STAGED sweep (many query pairs over a small corpus; api.cu): the first `a` chunks -- as many as one round of
  // clusters holds -- run in a launch of their own, and the kc-th best of each query's finished lists seeds the rest:
  // a bound drawn from a / chunks of the corpus instead of the pre-pass's 2-3 %, at no extra MMA work.  Chunks are
  // chosen so that the second launch is a whole number of rounds.
  plan->first_items = 0;
  if (stage_first && force_chunks <= 0 && tiles_per_item <= 0 && !a_resident) {
    const int64_t a = clusters / n_qpairs;
    // (the seed is the kc-th best of a * kc list entries per query, ranked in shared memory: long lists -- k above ~50 --
    // stay with the single launch)
    if (a >= 1 && a * kc <= kSeedWindow) {
      for (int64_t c = std::max<int64_t>(chunks, a + 1); c <= chunks + 2 * clusters && c <= n_tiles; ++c) {
        if (((c - a) * n_qpairs) % clusters == 0) {
          chunks = c;
          plan->first_items = static_cast<int>(a * n_qpairs);
          break;
        }
      }
    }
  }
  if (force_chunks > 0) chunks = force_chunks;
  chunks = std::max<int64_t>(1, std::min<int64_t>(chunks, n_tiles));
  plan->mode = mode;
  plan->n_stages = stages;
  plan->q_stages = 0;
  plan->n_chunks = static_cast<int>(chunks);
  plan->grid = 2 * static_cast<int>(std::min<int64_t>(clusters, chunks * n_qpairs));
  plan->smem_bytes = fixed + static_cast<size_t>(stages) * stage_bytes(mode, a_resident != 0);
  return 0;
}

int tc2_box_rows() { return kHalfRows; }

// Words of the rendezvous counters tc2_launch needs for this plan (0: the sweep does not rendezvous -- a single
// query pair, or items too short to drift apart).
static int sync_every() {
  if (env_set("DEWI_TC2_SYNC_EVERY")) return std::max(1, env_int("DEWI_TC2_SYNC_EVERY", kSyncEvery));  // experiments
  return kSyncEvery;
}

int64_t tc2_sync_words(const Tc2Plan& plan, int64_t n_rows, int n_qb) {
  if (n_qb < 4) return 0;
  if (env_int("DEWI_TC2_SYNC", 1) == 0) return 0;  // experiments
  const int64_t n_tiles = ceil_div(n_rows, kNTile);
  const int64_t per_item = ceil_div(n_tiles, plan.n_chunks);
  if (per_item < 4 * sync_every()) return 0;
  return 2 * static_cast<int64_t>(plan.n_chunks) * ceil_div(per_item + 1, sync_every());
}

int tc2_launch(const Tc2Plan& plan, const CUtensorMap& e0, const CUtensorMap& e1, const CUtensorMap& q0,
               const CUtensorMap& q1, int64_t n_rows, int dim, int n_qb, int kc, float* part_s, int* part_i,
               const SweepSeed& seed, cudaStream_t stream, unsigned int* sync_cnt, int fp16_planes, int item0, int item1) {
  Tc2Args a;
  a.fp16 = fp16_planes;
  a.n_rows = static_cast<int>(n_rows);
  a.n_tiles = static_cast<int>(ceil_div(n_rows, kNTile));
  a.n_kb = dim / kKBlock;
  a.n_qpairs = n_qb / 2;
  a.n_chunks = plan.n_chunks;
  a.n_items = plan.n_chunks * a.n_qpairs;
  a.item0 = item0;
  a.item1 = item1 < 0 ? a.n_items : std::min(item1, a.n_items);
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
  a.m_rows = 0;
  a.tau = 0.f;
  a.self_join = 0;
  a.a_offset = 0;
  a.sym = 0;
  a.blk0 = 0;
  a.rotate = 0;
  a.sync_cnt = nullptr;
  a.sync_blocks = 0;
  a.sync_every = sync_every();
  if (sync_cnt && a.n_qpairs > 1 && a.n_qpairs <= plan.grid / 2) {
    const int64_t words = tc2_sync_words(plan, n_rows, n_qb);
    if (words > 0) {
      if (item0 == 0) DEWI_CUDA(cudaMemsetAsync(sync_cnt, 0, static_cast<size_t>(words) * 4, stream));  // (rows are per chunk)
      a.sync_cnt = sync_cnt;
      a.sync_blocks = static_cast<int>(words / (2 * plan.n_chunks));
    }
  }
  a.row_best = nullptr;
  a.row_count = nullptr;
  a.pair_i = a.pair_j = nullptr;
  a.pair_sim = nullptr;
  a.pair_cap = 0;
  a.pair_count = nullptr;
  if (plan.mode == 0) return launch_one<0, EPI_TOPK>(plan, e0, e1, q0, q1, a, stream);
  if (plan.mode == 1) return launch_one<1, EPI_TOPK>(plan, e0, e1, q0, q1, a, stream);
  if (plan.mode == 2) return launch_one<2, EPI_TOPK>(plan, e0, e1, q0, q1, a, stream);
  return fail("unsupported CTA-pair sweep configuration");
}

// Thresholded join on the tensor cores: rows of A are the "queries" (128 per CTA, on TMEM lanes), rows of
// B stream through as the "corpus".  mode 0 = single bf16 plane (sims carry bf16 rounding, ~1e-3),
// mode 2 = hi/lo planes on both sides (three MMAs, ~1e-6).
int tc2_join_launch(int mode, const CUtensorMap& b0, const CUtensorMap& b1, const CUtensorMap& a0, const CUtensorMap& a1,
                    int64_t m_rows, int64_t m_pad, int64_t n_rows, int dim, int sm_count, float tau, int self_join, int64_t a_offset,
                    int sym, unsigned long long* row_best, int* row_count, int64_t* pair_i, int64_t* pair_j, float* pair_sim,
                    int64_t pair_cap, unsigned long long* pair_count, cudaStream_t stream) {
  Tc2Plan plan;
  const int n_qb = static_cast<int>(m_pad / kQueryBlock);
  const int64_t all_tiles = ceil_div(n_rows, kNTile);
  if (sym && (a_offset % kNTile) != 0) return fail("symmetric join: the row range must start on a multiple of 256");
  int sym_chunks = 0;
  if (sym) {
    // Items are dealt chunk-major (all row blocks walk their first chunk of tiles, then the second, ...), so a
    // handful of chunks per row block lets every row publish a tight running best early -- the column-direction
    // gates of all blocks read those -- while items stay long enough (>= 64 tiles) to hide their prologue.
    // Measured at 1M x 512 (B200, power-capped): 1 chunk 470 ms, 8 chunks 405 ms, 32 chunks 465 ms.
    const int64_t clusters = std::max(1, sm_count / 2), qpairs = n_qb / 2, per_item = all_tiles / 2 + 1;
    const int64_t want = std::max<int64_t>(1, std::min<int64_t>(8, per_item / 64));
    for (int64_t c = want; c <= per_item; ++c) {  // ... and at most 2 % of the cluster-rounds idle
      const int64_t items = c * qpairs, rounds = ceil_div(items, clusters);
      sym_chunks = static_cast<int>(c);
      if (static_cast<double>(items) >= 0.98 * static_cast<double>(rounds * clusters)) break;
    }
    if (env_set("DEWI_JOIN_CHUNKS")) sym_chunks = std::max(1, env_int("DEWI_JOIN_CHUNKS", sym_chunks));  // experiments
  }
  // row block resident in shared memory when it fits next to a ring of at least 4 stages
  bool a_res = fixed_bytes(0) + a_resident_bytes(mode, dim) + 4 * stage_bytes(mode, true) <= 227 * 1024;
  a_res = a_res && env_int("DEWI_JOIN_ARES", 1) != 0;  // experiments
  DEWI_TRY(tc2_make_plan(mode, dim, n_rows, n_qb, /*kc=*/0, sm_count, &plan, sym_chunks, sym ? all_tiles / 2 + 1 : 0, a_res ? 1 : 0));
  Tc2Args a;
  a.sym = sym ? 1 : 0;
  a.blk0 = sym ? static_cast<int>(a_offset / kNTile) : 0;
  a.rotate = 1;
  a.rotate = env_int("DEWI_JOIN_ROTATE", 1) != 0;  // experiments
  a.sync_cnt = nullptr;
  a.sync_blocks = 0;
  a.n_rows = static_cast<int>(n_rows);
  a.n_tiles = static_cast<int>(ceil_div(n_rows, kNTile));
  a.n_kb = dim / kKBlock;
  a.n_qpairs = n_qb / 2;
  a.n_chunks = plan.n_chunks;
  a.n_items = plan.n_chunks * a.n_qpairs;
  a.item0 = 0;
  a.item1 = a.n_items;
  a.kc = 0;
  a.e_stages = plan.n_stages;
  a.q_stages = plan.q_stages;
  a.part_s = nullptr;
  a.part_i = nullptr;
  a.seed = nullptr;
  a.seed_stride = a.seed_off = a.n_queries = 0;
  a.max_out = nullptr;
  a.max_groups = 0;
  a.fp16 = 0;
  a.m_rows = static_cast<int>(m_rows);
  a.tau = tau;
  a.self_join = self_join;
  a.a_offset = (self_join || sym) ? a_offset : 0;
  a.row_best = row_best;
  a.row_count = row_count;
  a.pair_i = reinterpret_cast<long long*>(pair_i);
  a.pair_j = reinterpret_cast<long long*>(pair_j);
  a.pair_sim = pair_sim;
  a.pair_cap = pair_cap;
  a.pair_count = pair_count;
  if (mode == 0 && a_res) return launch_one<0, EPI_JOIN, 1>(plan, b0, b1, a0, a1, a, stream);
  if (mode == 2 && a_res) return launch_one<2, EPI_JOIN, 1>(plan, b0, b1, a0, a1, a, stream);
  if (mode == 0) return launch_one<0, EPI_JOIN>(plan, b0, b1, a0, a1, a, stream);
  if (mode == 2) return launch_one<2, EPI_JOIN>(plan, b0, b1, a0, a1, a, stream);
  return fail("unsupported join precision mode");
}

}  // namespace dewi
