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
  float* s;   // &list_s[qlane]
  int* i;     // &list_i[qlane]
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
      const float x = l.s[k * kQueryBlock];
      if (x < m) { m = x; p = k; }
    }
    const int last = l.cnt - 1;
    l.s[p * kQueryBlock] = l.s[last * kQueryBlock];
    l.i[p * kQueryBlock] = l.i[last * kQueryBlock];
    l.cnt = last;
  }
  if (l.cnt == kc) {
    float m = l.s[0];
#pragma unroll 4
    for (int k = 1; k < kc; ++k) m = fminf(m, l.s[k * kQueryBlock]);
    l.thr = m;
  }
}

// One accumulator tile: this thread's TMEM lane, N_TILE fp32 columns starting at `tcol`; column j is
// corpus row row_base + j.
template <int N_TILE>
__device__ __forceinline__ void scan_tile(LaneList& l, int kc, uint32_t tcol, int row_base, int n_rows) {
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
    float mx = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
    if (__any_sync(0xffffffffu, mx > l.thr)) {
#pragma unroll
      for (int g = 0; g < 32 / kGroup; ++g) {
#pragma unroll
        for (int jj = 0; jj < kGroup; ++jj) {
          const int j = g * kGroup + jj;
          if (v[j] > l.thr) {
            l.s[l.cnt * kQueryBlock] = v[j];
            l.i[l.cnt * kQueryBlock] = r0 + j;
            ++l.cnt;
          }
        }
        // a lane enters a group with cnt <= kc + kPending - kGroup, so the appends above fit
        if (__any_sync(0xffffffffu, l.cnt > kc + kPending - kGroup)) lane_prune(l, kc);
      }
    }
  }
}

// End of a work item: keep the kc best and write them as [k][query] partials.
__device__ __forceinline__ void flush_item(LaneList& l, int kc, float* ps, int* pi) {
  lane_prune(l, kc);
  for (int k = 0; k < kc; ++k) {
    const bool live = k < l.cnt;
    ps[k * kQueryBlock] = live ? l.s[k * kQueryBlock] : -INFINITY;
    pi[k * kQueryBlock] = live ? l.i[k * kQueryBlock] : -1;
  }
}

}  // namespace sweep
}  // namespace dewi
