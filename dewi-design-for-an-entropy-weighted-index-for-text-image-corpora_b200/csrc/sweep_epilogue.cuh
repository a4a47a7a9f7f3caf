// Pieces shared by the 1-CTA and 2-CTA tensor-core sweeps: work-item geometry and the epilogue's
// per-query candidate lists (TMEM lane == query).
#pragma once
#include "internal.h"
#include "ptx.cuh"

namespace dewi {
namespace sweep {

constexpr int kPending = 16;   // list slots beyond kc: candidates appended between two prunes
constexpr int kGroup = 8;      // columns examined between two overflow checks (kGroup <= kPending / 2)
constexpr int kQStagesMax = 8;

template <int MODE>
struct ModeTraits {
  static constexpr int PE = (MODE == 2) ? 2 : 1;  // corpus planes streamed
  static constexpr int PQ = (MODE == 0) ? 1 : 2;  // query planes streamed
};

__device__ __forceinline__ void tile_range(int chunk, int n_chunks, int n_tiles, int& t0, int& t1) {
  t0 = static_cast<int>((static_cast<long long>(chunk) * n_tiles) / n_chunks);
  t1 = static_cast<int>((static_cast<long long>(chunk + 1) * n_tiles) / n_chunks);
}

// Each epilogue thread owns one query (TMEM lane) and an UNSORTED list of up to kc + kPending
// (score, row) pairs in shared memory, laid out [slot][query] so that lanes never share a bank.
// Scores above the query's admission threshold are appended (two predicated stores); when any lane
// of the warp is about to overflow, all lanes prune together, in lockstep, down to their kc best
// and raise their thresholds.  Appending is cheap when candidates are rare (the steady state);
// pruning is lane-parallel when candidates are frequent (the first tiles of an item).
struct LaneList {
  float* s;    // this lane's column: &list_s[column]
  int* i;      // &list_i[column]
  float* ws;   // column of the warp's lane 0 (lanes of one warp own consecutive columns)
  int* wi;
  int stride;  // columns per slot row: 128, or 64 when only 64 queries are resident (M = 64 sweeps)
  int cnt;
  float thr;
};

__device__ __forceinline__ void lane_prune(LaneList& l, int kc) {
  // remove the smallest entries until kc remain; then thr = smallest kept score
  while (l.cnt > kc) {
    float m = l.s[0];
    int p = 0;
#pragma unroll 4
    for (int k = 1; k < l.cnt; ++k) {
      const float x = l.s[k * l.stride];
      if (x < m) { m = x; p = k; }
    }
    const int last = l.cnt - 1;
    l.s[p * l.stride] = l.s[last * l.stride];
    l.i[p * l.stride] = l.i[last * l.stride];
    l.cnt = last;
  }
  if (l.cnt == kc) {
    float m = l.s[0];
#pragma unroll 4
    for (int k = 1; k < kc; ++k) m = fminf(m, l.s[k * l.stride]);
    l.thr = m;
  }
}

// Admission threshold seeded from a sample pre-pass (api.cu): `seed[b * stride + off]` is the kc-th best
// score of query b over a sample of the corpus -- a lower bound of its kc-th best over the whole
// corpus, so nothing at or above it may be dropped: start just below it (admission is strict `>`).
// TMEM lane L of block qb holds query qb * 128 + (L % 32) * 4 + L / 32 (inverse of query_lane).
__device__ __forceinline__ float seed_threshold_query(const float* seed, int stride, int off, int n_queries, int b) {
  if (!seed) return -INFINITY;
  if (b >= n_queries) return -INFINITY;
  const float t = seed[static_cast<size_t>(b) * stride + off];
  if (!(t > -INFINITY)) return -INFINITY;  // fewer than kc sample rows (or NaN): no seed
  const unsigned int u = __float_as_uint(t);
  if (t > 0.f) return __uint_as_float(u - 1u);
  if (t < 0.f) return __uint_as_float(u + 1u);
  return __uint_as_float(0x80000001u);  // just below zero
}
__device__ __forceinline__ float seed_threshold(const float* seed, int stride, int off, int n_queries, int qb, int qlane) {
  return seed_threshold_query(seed, stride, off, n_queries, qb * kQueryBlock + (qlane & 31) * 4 + (qlane >> 5));
}

// Warp-cooperative prune of ONE lane's list (the steady state, where lanes overflow one at a time):
// every entry's rank among the `cnt` entries is counted by all 32 lanes together (each lane owns up to
// two entries and compares them with every entry, read as a shared-memory broadcast), the kc best
// are written back in rank order, and the kc-th becomes the threshold.  ~cnt broadcast loads instead
// of the (cnt - kc) x cnt dependent scans of lane_prune run by a single active lane.
// Requires kc + kPending <= 64.  `col_s` / `col_i` point at the target lane's column.
__device__ __forceinline__ float coop_prune(float* col_s, int* col_i, int cnt, int kc, int lane, int stride) {
  const int e0 = lane, e1 = lane + 32;
  const bool has0 = e0 < cnt, has1 = e1 < cnt;
  const float v0 = has0 ? col_s[e0 * stride] : 0.f, v1 = has1 ? col_s[e1 * stride] : 0.f;
  const int i0 = has0 ? col_i[e0 * stride] : 0, i1 = has1 ? col_i[e1 * stride] : 0;
  int r0 = 0, r1 = 0;
#pragma unroll 4
  for (int f = 0; f < cnt; ++f) {
    const float vf = col_s[f * stride];  // same address in every lane: broadcast
    r0 += (vf > v0 || (vf == v0 && f < e0)) ? 1 : 0;
    r1 += (vf > v1 || (vf == v1 && f < e1)) ? 1 : 0;
  }
  __syncwarp();  // all reads done before the in-place rewrite
  if (has0 && r0 < kc) { col_s[r0 * stride] = v0; col_i[r0 * stride] = i0; }
  if (has1 && r1 < kc) { col_s[r1 * stride] = v1; col_i[r1 * stride] = i1; }
  // the entry ranked kc-1 is the new admission threshold
  const bool mine0 = has0 && r0 == kc - 1, mine1 = has1 && r1 == kc - 1;
  const unsigned int who = __ballot_sync(0xffffffffu, mine0 || mine1);
  const float t = __shfl_sync(0xffffffffu, mine0 ? v0 : v1, who ? (__ffs(who) - 1) : 0);
  __syncwarp();
  return t;
}

// One accumulator tile: this thread's TMEM lane, N_TILE fp32 columns starting at `tcol`; column j is
// corpus row row_base + j.
template <int N_TILE>
__device__ __forceinline__ void scan_tile(LaneList& l, int kc, uint32_t tcol, int row_base, int n_rows) {
  const int lane = threadIdx.x & 31;
  const bool coop_ok = kc + kPending <= 64;
#pragma unroll 1
  for (int c = 0; c < N_TILE / 32; ++c) {
    float v[32];
    ptx::tmem_ld_32x32(tcol + c * 32, v);
    const int r0 = row_base + c * 32;
    if (r0 + 32 > n_rows) {  // ragged corpus tail: TMA zero-filled those rows
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (r0 + j >= n_rows) v[j] = -INFINITY;
    }
    // maxima of the four 8-column groups first: when hits are frequent but sparse (many query blocks over a small corpus:
    // a couple per warp and 32 columns) only the groups that hold one are walked
    float gm[32 / kGroup];
#pragma unroll
    for (int g = 0; g < 32 / kGroup; ++g) {
      gm[g] = v[g * kGroup];
#pragma unroll
      for (int jj = 1; jj < kGroup; ++jj) gm[g] = fmaxf(gm[g], v[g * kGroup + jj]);
    }
    float mx = gm[0];
#pragma unroll
    for (int g = 1; g < 32 / kGroup; ++g) mx = fmaxf(mx, gm[g]);
    if (__any_sync(0xffffffffu, mx > l.thr)) {
#pragma unroll
      for (int g = 0; g < 32 / kGroup; ++g) {
        if (!__any_sync(0xffffffffu, gm[g] > l.thr)) continue;
#pragma unroll
        for (int jj = 0; jj < kGroup; ++jj) {
          const int j = g * kGroup + jj;
          if (v[j] > l.thr) {
            l.s[l.cnt * l.stride] = v[j];
            l.i[l.cnt * l.stride] = r0 + j;
            ++l.cnt;
          }
        }
        // a lane enters a group with cnt <= kc + kPending - kGroup, so the appends above fit
        unsigned int over = __ballot_sync(0xffffffffu, l.cnt > kc + kPending - kGroup);
        if (over) {
          if (!coop_ok || __popc(over) >= 8) {
            lane_prune(l, kc);  // many lanes full at once (the first tiles of an item): lane-parallel
          } else {
            __syncwarp();
            while (over) {  // a few lanes: the warp prunes them one at a time, cooperatively
              const int tgt = __ffs(over) - 1;
              over &= over - 1;
              const int tcnt = __shfl_sync(0xffffffffu, l.cnt, tgt);
              const float t = coop_prune(l.ws + tgt, l.wi + tgt, tcnt, kc, lane, l.stride);
              if (lane == tgt) { l.cnt = kc; l.thr = t; }
            }
          }
        }
      }
    }
  }
}

// Pre-pass epilogue: only the running maximum of the lane's scores over the item (no lists).
template <int N_TILE>
__device__ __forceinline__ float max_tile(float best, uint32_t tcol, int row_base, int n_rows) {
#pragma unroll 1
  for (int c = 0; c < N_TILE / 32; ++c) {
    float v[32];
    ptx::tmem_ld_32x32(tcol + c * 32, v);
    const int r0 = row_base + c * 32;
    if (r0 + 32 > n_rows) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (r0 + j >= n_rows) v[j] = -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) best = fmaxf(best, v[j]);
  }
  return best;
}

// Pre-pass epilogue, fine granularity: the maximum of every 32-row group of the tile, written to
// out[(tile * G + group) * n_qb * 128] (G = N_TILE / 32; `out` already points at this query's column).
template <int N_TILE>
__device__ __forceinline__ void max_groups_tile(float* out, size_t group_stride, int tile, uint32_t tcol, int row_base, int n_rows) {
#pragma unroll 1
  for (int c = 0; c < N_TILE / 32; ++c) {
    float v[32];
    ptx::tmem_ld_32x32(tcol + c * 32, v);
    const int r0 = row_base + c * 32;
    if (r0 + 32 > n_rows) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (r0 + j >= n_rows) v[j] = -INFINITY;
    }
    float m = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
    out[(static_cast<size_t>(tile) * (N_TILE / 32) + c) * group_stride] = m;
  }
}

// End of a work item: keep the kc best and write them as [k][query] partials.
__device__ __forceinline__ void flush_item(LaneList& l, int kc, float* ps, int* pi) {
  lane_prune(l, kc);
  for (int k = 0; k < kc; ++k) {
    const bool live = k < l.cnt;
    ps[k * kQueryBlock] = live ? l.s[k * l.stride] : -INFINITY;   // (the partials keep 128 columns per slot)
    pi[k * kQueryBlock] = live ? l.i[k * l.stride] : -1;
  }
}

}  // namespace sweep
}  // namespace dewi
