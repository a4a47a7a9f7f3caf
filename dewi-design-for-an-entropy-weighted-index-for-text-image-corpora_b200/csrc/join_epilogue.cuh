// Join epilogue of the CTA-pair sweep (search_tc2.cu, EPI_JOIN): thresholded similarity join on the tensor cores.
// Row direction (every tile): per row of A (TMEM lane) a running best match, a count of similarities >= tau and the
// (i, j, sim) pairs.  Column direction (off-diagonal tiles of the symmetric self-join): the same statistics for the
// rows of the OTHER block, behind a gate that rejects almost every 32 x 32 group.  The reference defines none of
// these reductions (src/dewi/signals/redundancy.py:36-38 stops at the dense product); see include/dewi_b200.h.
#pragma once
#include "internal.h"
#include "ptx.cuh"

namespace dewi {
namespace {

struct Tc2Args {
  int n_rows;
  int n_tiles;
  int n_kb;
  int n_qpairs;   // query-block pairs
  int n_chunks;
  int n_items;    // n_chunks * n_qpairs, dealt to clusters
  int item0, item1;  // the items this launch walks (a staged sweep runs the first chunks in a launch of their own)
  int kc;
  int e_stages;
  int q_stages;
  float* part_s;  // [chunk][qb][kc][128]
  int* part_i;
  const float* seed;  // optional admission thresholds from the sample pre-pass (see seed_threshold)
  int seed_stride, seed_off, n_queries;
  float* max_out;     // pre-pass mode: [chunk][qb][128] maximum score per query, no candidate lists
  int max_groups;     // pre-pass mode: one maximum per 32-row group of every tile instead (SweepSeed::max_groups)
  int fp16;           // operand planes hold fp16 (fp32 corpus) instead of bf16
  // EPI_JOIN: thresholded similarity join instead of candidate lists (dewi_join)
  int m_rows;                      // rows of A (the "query" side); rows >= m_rows are padding
  float tau;
  int self_join;                   // A's row i is B's row a_offset + i: exclude that column, emit only pairs beyond it
  long long a_offset;
  // Symmetric self-join (sym = 1): A is rows [a_offset, a_offset + m_rows) of B, a_offset % 256 == 0.  Row block
  // ig (256 rows) meets only the tiles ig, ig+1, ..., ig+L-1 (mod n_tiles) -- the circulant half, L ~ n_tiles / 2 --
  // and every off-diagonal tile also updates the statistics of its COLUMNS (rows of the other block), so each
  // unordered block pair is multiplied once.  row_best / row_count are then indexed by GLOBAL row.
  int sym;
  int blk0;                        // a_offset / 256
  int rotate;                      // align the clusters' tile walks (item_span)
  // Top-k sweep with several query-block pairs: the clusters that stream the SAME corpus chunk for different
  // query pairs only share its tiles through L2 while they stay within a few tiles of each other -- once they
  // drift, every cluster pulls its own copy from DRAM, the fill rate evicts tiles within ~20 us and the sweep
  // stays in that state (ncu: the corpus read 8.9x from DRAM at B = 4096).  The producers of a chunk's clusters
  // therefore meet every kSyncEvery tiles at a counter (bounded spin: a late cluster is never waited for longer
  // than kSyncSpin cycles, and it catches up because it never waits itself).
  unsigned int* sync_cnt;          // [n_chunks * 2][sync_blocks] arrival counters, zeroed before the launch (or null)
  int sync_blocks;
  int sync_every;
  unsigned long long* row_best;    // [m_rows] packed (orderable(sim) << 32 | ~j), atomicMax
  int* row_count;                  // [m_rows] sims >= tau
  long long* pair_i;
  long long* pair_j;
  float* pair_sim;
  long long pair_cap;
  unsigned long long* pair_count;
};


enum { EPI_TOPK = 0, EPI_JOIN = 1 };

__device__ __forceinline__ unsigned int orderable_u32(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Join epilogue state of one A row (TMEM lane): running best match and threshold count.
struct JoinRow {
  float best;
  int best_j;
  int count;
};

template <int N_TILE>
__device__ __forceinline__ void join_scan_tile(JoinRow& r, const Tc2Args& a, int i_row, uint32_t tcol, int col_base) {
#pragma unroll 1
  for (int c = 0; c < N_TILE / 32; ++c) {
    float v[32];
    ptx::tmem_ld_32x32(tcol + c * 32, v);
    const int j0 = col_base + c * 32;
    if (j0 + 32 > a.n_rows) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j0 + j >= a.n_rows) v[j] = -INFINITY;
    }
    const long long i_glob = i_row + a.a_offset;  // this row's index on the B side (self / slice joins)
    if (a.self_join && i_glob >= j0 && i_glob < j0 + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j0 + j == i_glob) v[j] = -INFINITY;  // the diagonal
    }
    float mx = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
    if (mx > r.best) {  // rare after the first tiles: the running maximum seldom improves
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (v[j] > r.best) { r.best = v[j]; r.best_j = j0 + j; }
    }
    if (mx >= a.tau && i_row < a.m_rows) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (v[j] >= a.tau) {
          ++r.count;
          if (!a.self_join || j0 + j > i_glob) {
            const unsigned long long slot = atomicAdd(a.pair_count, 1ull);
            if (static_cast<long long>(slot) < a.pair_cap) {
              a.pair_i[slot] = i_glob;
              a.pair_j[slot] = j0 + j;
              a.pair_sim[slot] = v[j];
            }
          }
        }
      }
    }
  }
}


__device__ __forceinline__ unsigned long long pack_best(float sim, long long idx) {
  return (static_cast<unsigned long long>(orderable_u32(sim)) << 32) |
         static_cast<unsigned int>(~static_cast<unsigned int>(idx));
}
__device__ __forceinline__ float best_sim(unsigned long long key) {
  if (key == 0ull) return -INFINITY;
  const unsigned int o = static_cast<unsigned int>(key >> 32);
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// Column direction of an off-diagonal tile of the symmetric join: column j0 + c is row j0 + c of the matrix,
// and the 32 rows this warp holds may contain its best match.  Almost every 32 x 32 group is rejected by the
// GATE -- no element beats the weakest current best of the group's 32 columns.  The gate values of a tile are
// fetched one tile ahead (row_best only ever grows, so a stale value merely lets a useless group through):
// lane c of `lbv` holds the minimum over group c's columns.
template <int N_TILE>
__device__ __forceinline__ void gate_prefetch(unsigned long long (&raw)[N_TILE / 32], const Tc2Args& a, int col_base, int lane) {
#pragma unroll
  for (int c = 0; c < N_TILE / 32; ++c) {
    const int j = col_base + c * 32 + lane;
    raw[c] = (j < a.n_rows) ? __ldcg(a.row_best + j) : ~0ull;  // ~0: padding column, never improved
  }
}
// `cache` (shared memory, one copy per epilogue warp) keeps the per-column values for the groups that pass.
template <int N_TILE>
__device__ __forceinline__ float gate_reduce(const unsigned long long (&raw)[N_TILE / 32], int lane, float* cache) {
  float lbv = INFINITY;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < N_TILE / 32; ++c) {
    float s = (raw[c] == ~0ull) ? INFINITY : best_sim(raw[c]);
    cache[c * 32 + lane] = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fminf(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (lane == c) lbv = s;
  }
  __syncwarp();
  return lbv;
}

// A group passed the gate: a transposed butterfly (31 shuffles) leaves lane c with column c's maximum over
// the warp's rows; the columns that beat their current best publish (sim, row) with one 64-bit atomicMax.
__device__ __forceinline__ void join_column_update(const float (&v)[32], bool row_ok, float cur_s, int j0, long long i_glob,
                                                   int lane, const Tc2Args& a) {
  float w16[16], w8[8], w4[4], w2[2];
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    // padding rows of the last block are zero vectors: they must not win any column
    const float lo = row_ok ? v[k] : -INFINITY, hi = row_ok ? v[k + 16] : -INFINITY;
    const float send = b4 ? lo : hi, keep = b4 ? hi : lo;
    w16[k] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float send = b3 ? w16[k] : w16[k + 8], keep = b3 ? w16[k + 8] : w16[k];
    w8[k] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = b2 ? w8[k] : w8[k + 4], keep = b2 ? w8[k + 4] : w8[k];
    w4[k] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 4));
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = b1 ? w4[k] : w4[k + 2], keep = b1 ? w4[k + 2] : w4[k];
    w2[k] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 2));
  }
  const float send = b0 ? w2[0] : w2[1], keep = b0 ? w2[1] : w2[0];
  const float m = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));  // column j0 + lane over the warp's 32 rows
  const unsigned int improved = __ballot_sync(0xffffffffu, m > cur_s);
  if (!improved) return;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (improved & (1u << c)) {  // warp-uniform
      const float mc = __shfl_sync(0xffffffffu, m, c);
      const unsigned int who = __ballot_sync(0xffffffffu, row_ok && v[c] == mc);
      if (who && lane == c) atomicMax(&a.row_best[j0 + c], pack_best(mc, i_glob - lane + (__ffs(who) - 1)));
    }
  }
}

// Symmetric self-join epilogue for one tile.  Diagonal tile (the block against itself): every row sees its
// whole 256-column neighbourhood, so only the row direction runs, own column excluded, pairs for j > i.
// Off-diagonal tile: row direction as usual, every sim >= tau also counts for row j and is emitted once as
// (min, max); the column direction updates the other block's best matches.
template <int N_TILE>
__device__ __forceinline__ void join_scan_tile_sym(JoinRow& r, const Tc2Args& a, long long i_glob, int lane, uint32_t tcol,
                                                   int col_base, bool diag, float lbv, const float* cache) {
  const bool row_ok = i_glob < a.n_rows;
#pragma unroll 1
  for (int c = 0; c < N_TILE / 32; ++c) {
    const int j0 = col_base + c * 32;
    float v[32];
    ptx::tmem_ld_32x32(tcol + c * 32, v);
    if (j0 + 32 > a.n_rows) {  // ragged last tile (warp-uniform)
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j0 + j >= a.n_rows) v[j] = -INFINITY;
    }
    if (diag && i_glob >= j0 && i_glob < j0 + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j0 + j == i_glob) v[j] = -INFINITY;
    }
    float mx = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
    if (!row_ok) mx = -INFINITY;  // padding rows of the last block take no part (their sims are all 0)
    if (mx > r.best) {  // rare: r.best starts from the row's best so far (see the item prologue)
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (v[j] > r.best) { r.best = v[j]; r.best_j = j0 + j; }
    }
    if (mx >= a.tau) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (v[j] >= a.tau) {
          ++r.count;
          const long long jj = j0 + j;
          if (!diag) atomicAdd(&a.row_count[jj], 1);
          if (!diag || jj > i_glob) {
            const unsigned long long slot = atomicAdd(a.pair_count, 1ull);
            if (static_cast<long long>(slot) < a.pair_cap) {
              a.pair_i[slot] = i_glob < jj ? i_glob : jj;
              a.pair_j[slot] = i_glob < jj ? jj : i_glob;
              a.pair_sim[slot] = v[j];
            }
          }
        }
      }
    }
    __syncwarp();
    if (!diag) {
      const float lb = __shfl_sync(0xffffffffu, lbv, c);
      // cache[]: the columns' bests as of one tile ago (+inf for padding columns) -- stale values only cost atomics
      if (__any_sync(0xffffffffu, mx > lb)) join_column_update(v, row_ok, cache[c * 32 + lane], j0, i_glob, lane, a);
    }
  }
}


}  // namespace
}  // namespace dewi
