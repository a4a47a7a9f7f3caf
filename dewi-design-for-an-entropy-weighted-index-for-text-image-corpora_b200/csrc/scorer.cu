// K3 / K4 -- robust statistics and the DEWI score.
//
// K3 fit_stats: exact median and MAD of each signal column, replacing RobustStats.fit
//   (reference src/dewi/scorer.py:18-26: np.median of a float32 column, then np.median of
//   |v - med| in float32, zero MAD -> 1e-8).  Exact order statistics by MSB-first radix selection
//   over order-preserving uint32 keys: three histogram passes (11 + 11 + 10 bits) per selection,
//   block-private shared-memory histograms flushed with one atomic per non-empty bin, the last block
//   of a pass picks the bins.  Large columns select inside a key window found from a sample, so each
//   statistic reads the column once (window_kernel).  All columns and both middle ranks (even n) are
//   selected in the same passes.  The windowed path is FOUR launches per statistic: the sample (with the
//   first histogram pass fused in), one more histogram pass over the sample (-> window bounds), the window
//   pass (with the histogram of the window-relative top digit and the pick of the two target bins fused in),
//   and a survivor pass that gathers the few hundred keys of those bins and sorts them in shared memory
//   (-> the exact middle values, and the state of the next statistic armed).
// K4 score: RobustStats.z + DewiScorer._components/score/score_conditional (scorer.py:28-31,49-89)
//   in float64: U as a weighted sum of (v - med) with the z-score scales folded into the weights, clip, sigmoid with a
//   polynomial exp -- within ~1e-11 of the reference's Python-float arithmetic (gate 1e-6); 7 fp32 loads and one store per row.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "internal.h"

namespace dewi {
namespace {

constexpr int kBins = 2048;
constexpr int kMaxCols = 32;
// windowed path (large n): sample -> key window around the target rank -> one full pass
constexpr long long kWindowMinRows = 4ll << 20;  // below this the plain radix passes are cheap enough
constexpr int kSampleRuns = 1024;                // evenly spaced runs of ...
constexpr int kSampleRunLen = 1024;              // ... contiguous elements (coalesced)
constexpr int kSample = kSampleRuns * kSampleRunLen;
constexpr int kSampleBlockRuns = 4;              // runs a sample_kernel block gathers per trip (16 loads per thread in flight)
constexpr int kSampleBlocks = 64;                // blocks per column (4 trips each)
// +- sample ranks around the middle: 4 sigma of a 2^20 sample's median rank (sigma = 512).  A window that misses its
// ranks (6e-5 per column) is detected and costs the fallback to the plain radix passes, never a wrong answer; a
// narrower window means fewer keys to collect in the full pass (the one-pass path collects ~5x the window's mass).
constexpr int kSampleMargin = 2048;
// Window pass with the in-window histogram fused in (DEWI_FIT_FUSED=1) or as a launch of its own (0).  Measured at
// 100M x 7 (CUDA events in-stream, warm): fused 483-498 us per pass, separate 445-451 + 20 us -- the extra live state
// costs the streaming loop more than the launch it saves, so the separate histogram is the default.
constexpr int kFitFusedDefault = 0;
constexpr int kWindowDigitBits = 11;             // window-relative digit of the in-window selection (2048 bins)
// One full pass for BOTH statistics (DEWI_FIT_ONEPASS): the sample also brackets the MAD -- as a window of |y - c0|
// around the centre c0 of the median's window -- so the pass over the column collects the median's keys and the
// MAD's keys together; the exact median then turns the second set into exact deviations.  Half the bytes of the
// "median, then a dependent pass for the MAD" scheme.
constexpr int kFitOnePassDefault = 1;
constexpr int kSurvCap = 8192;                   // keys of the target bins sorted in shared memory (typically a few hundred)

struct SelState {
  // radix selection, per column, two rank slots
  unsigned int prefix[kMaxCols][2];
  unsigned long long rank[kMaxCols][2];
  int same[kMaxCols];  // both slots still share one prefix -> one histogram serves both
  float med[kMaxCols];
  float mad[kMaxCols];
  // windowed path
  unsigned int lo[kMaxCols], hi[kMaxCols];  // inclusive key window
  unsigned long long below[kMaxCols];       // keys < lo
  unsigned int wcnt[kMaxCols];              // keys appended to the window buffer
  int miss[kMaxCols];                       // window overflowed or does not hold the target ranks
  // in-window selection: digit = (key - lo) >> wshift is the top <= 11 bits of the window's key span
  int wshift[kMaxCols];
  unsigned int wbin[kMaxCols][2];           // digits of the two middle ranks
  unsigned long long wrank[kMaxCols][2];    // their ranks inside those bins
  unsigned int tiles_done[kMaxCols];        // window_kernel: tiles of the column finished (self-resetting)
  unsigned int surv_n[kMaxCols];            // survivor_kernel: keys gathered from the target bins
  unsigned int surv_n0[kMaxCols];           //   ... of which in bin 0, when the two bins differ
  unsigned int surv_done[kMaxCols];         // survivor_kernel: block tickets (self-resetting)
  // one-pass path: the MAD's keys are collected in the SAME pass as the median's, around the window centre c0
  float c0[kMaxCols], w[kMaxCols];          // centre and (inflated) half-width of the median window, value space
  float A[kMaxCols], B[kMaxCols];           // collect y with A <= |y - c0| <= B  (deviation window around c0)
  float r0[kMaxCols];                       // radius of the median window around c0 as the pass tests it: |y - c0| <= r0
  float cd[kMaxCols], rd[kMaxCols];         // centre / radius of the deviation window as the pass tests it
  unsigned long long belowd[kMaxCols];      // #(|y - c0| < A)
  unsigned int cntd[kMaxCols];              // keys appended to the deviation window buffer
};

__device__ __forceinline__ unsigned int orderable(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

constexpr int kSamplePass1 = 3;   // pass id of the sample selection's second, coarser digit

__device__ __forceinline__ void pass_bits(int pass, int& shift, int& nbins, unsigned int& himask) {
  // pass 0: bits 31..21, pass 1: bits 20..10, pass 2: bits 9..0
  // pass 3 (sample selection only, after pass 0): bits 20..13 -- 256 bins instead of 2048, an eighth of the global
  // atomics when the blocks hand their histograms over; the window bounds round outwards by 2^13 keys instead of 2^10
  if (pass == 0) { shift = 21; nbins = 2048; himask = 0u; }
  else if (pass == 1) { shift = 10; nbins = 2048; himask = 0xFFE00000u; }
  else if (pass == 2) { shift = 0; nbins = 1024; himask = 0xFFFFFC00u; }
  else { shift = 13; nbins = 256; himask = 0xFFE00000u; }
}

// Key sources.  RAW: fp32 column value; DEV: |v - med| in fp32 (scorer.py:24); KEYS: a buffer of keys.
// SKEY_DEV: a buffer of keys of VALUES, re-keyed as |value - centre| (one-pass path: deviations of the sample around c0).
enum { SRC_RAW = 0, SRC_DEV = 1, SRC_KEYS = 2, SRC_SKEY_DEV = 3 };

template <int SRC>
__device__ __forceinline__ unsigned int load_key(const void* col, long long i, float med) {
  if (SRC == SRC_KEYS) return __ldg(static_cast<const unsigned int*>(col) + i);
  if (SRC == SRC_SKEY_DEV) {
    const unsigned int k = __ldg(static_cast<const unsigned int*>(col) + i);
    const float y = __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);   // from_orderable
    return orderable(fabsf(__fsub_rn(y, med)));
  }
  float v = __ldg(static_cast<const float*>(col) + i);
  if (SRC == SRC_DEV) v = fabsf(__fsub_rn(v, med));
  return orderable(v);
}

__device__ void pick_column(int pass, int c, SelState* st, unsigned int* ghist, unsigned long long* cum,
                            unsigned long long* wsum);

// What the last block of a selection pass does after it has picked its bins.
enum { OP_NONE = 0, OP_WINDOW_BOUNDS = 1, OP_WINDOW_BOUNDS_1P = 2, OP_DEV_BOUNDS = 3 };

// The sample selection resolved the top 19 bits (pass 0 and the coarse pass 3) of the keys at sample ranks
// s/2 -+ margin: rounding the first down and the second up to a multiple of 2^13 keeps the window conservative (finer
// digits would buy a few per cent fewer window keys for more atomics or another launch).  Also sizes the
// window-relative digit of the in-window selection.
__device__ __forceinline__ void arm_window(SelState* st, int c) {
  const unsigned int lo = st->prefix[c][0] & 0xFFFFE000u, hi = st->prefix[c][1] | 0x1FFFu;
  st->lo[c] = lo;
  st->hi[c] = hi;
  st->below[c] = 0ull;
  st->wcnt[c] = 0u;
  const int nbits = 32 - __clz(static_cast<int>(hi - lo));   // hi - lo >= 0x1FFF
  st->wshift[c] = max(nbits - kWindowDigitBits, 0);
}

// State of a fresh sample selection: ranks s/2 -+ margin of the kSample sample keys.
__device__ __forceinline__ void arm_sample(SelState* st, int c) {
  st->prefix[c][0] = st->prefix[c][1] = 0u;
  st->rank[c][0] = kSample / 2 - kSampleMargin;
  st->rank[c][1] = kSample / 2 + kSampleMargin;
  st->same[c] = 1;
}

__device__ __forceinline__ float key_to_float(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// One-pass path, after the median window [lo, hi] is known: its centre c0 and half-width w in value space (w inflated
// by 2^-10 and by a slack far above any fp32 rounding of the quantities below), and the sample selection of the
// deviations |s - c0| armed.  With med anywhere in the window, ||y - c0| - |y - med|| <= w for every y.
__device__ __forceinline__ void arm_dev_sample(SelState* st, int c) {
  const float lo_f = key_to_float(st->lo[c]), hi_f = key_to_float(st->hi[c]);
  const float c0 = 0.5f * lo_f + 0.5f * hi_f;
  float w = fmaxf(hi_f - c0, c0 - lo_f);
  w = w * 1.0009765625f + (fabsf(c0) + w) * 3.814697265625e-06f + 1e-30f;   // * (1 + 2^-10) + 2^-18 (|c0| + w)
  st->c0[c] = c0;
  st->w[c] = w;
  st->r0[c] = w;   // the pass collects |fl(y - c0)| <= r0: a superset of [lo, hi] (w is the inflated half-width)
  if (!(w < INFINITY) || !(c0 == c0)) st->miss[c] = 1;   // infinite / NaN bounds: not a usable window
  arm_sample(st, c);
}

// One-pass path, after the sample's deviations |s - c0| bracket the middle rank by [dlo, dhi] (19 key bits resolved,
// rounded outwards): the column's MAD -- deviations around the exact median, each within w of the deviation around
// c0 -- lies in [dlo - w, dhi + w], and every element with such a deviation has |y - c0| in [dlo - 2w, dhi + 2w] = [A, B].
__device__ __forceinline__ void arm_dev_window(SelState* st, int c) {
  const float dlo = key_to_float(st->prefix[c][0] & 0xFFFFE000u), dhi = key_to_float(st->prefix[c][1] | 0x1FFFu);
  const float w = st->w[c];
  // The pass tests | |fl(y - c0)| - cd | <= rd with cd, rd the centre and radius of [dlo - 2w, dhi + 2w]; A and B are
  // what the later exactness argument may assume of every uncollected element (|y - c0| < A or > B): the tested
  // interval shrunk by a slack far above the rounding of the two subtractions.
  const float a = dlo - 2.f * w, b = dhi + 2.f * w;
  const float cd = 0.5f * a + 0.5f * b;
  const float rd = fmaxf(b - cd, cd - a) * 1.0000038146972656f;
  const float slack = (fabsf(st->c0[c]) + fabsf(b)) * 3.814697265625e-06f;
  st->cd[c] = cd;
  st->rd[c] = rd;
  st->A[c] = fmaxf(cd - rd + slack, 0.f);
  st->B[c] = cd + rd - slack;    // (+inf / NaN collect everything: the buffer overflows and the column is flagged)
  if (!(rd < INFINITY) || !(cd == cd)) st->miss[c] = 1;
  st->belowd[c] = 0ull;
  st->cntd[c] = 0u;
}

// Histogram of the current digit over the keys whose higher digits match a slot's prefix: block-private
// shared-memory histograms (per-thread run-length combining in front of the shared atomics), one global
// atomic per non-empty bin at the end; the last block of a column then picks the bins (pick_column).
template <int SRC>
__global__ void __launch_bounds__(512)
hist_kernel(const void* __restrict__ src, long long n_host, const unsigned int* __restrict__ n_dev, long long ld, int pass,
            SelState* st, unsigned int* __restrict__ ghist, unsigned int* __restrict__ done, int op) {
  __shared__ __align__(8) unsigned int sh[2 * kBins];   // two slot histograms; reused as 2048 x u64 by pick_column
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned int ticket;
  const int c = blockIdx.y;
  int shift, nbins;
  unsigned int himask;
  pass_bits(pass, shift, nbins, himask);
  for (int i = threadIdx.x; i < 2 * kBins; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const unsigned int p0 = st->prefix[c][0] & himask, p1 = st->prefix[c][1] & himask;
  const bool same = st->same[c] != 0;
  const float med = (SRC == SRC_DEV) ? st->med[c] : ((SRC == SRC_SKEY_DEV) ? st->c0[c] : 0.f);
  const long long n = n_dev ? static_cast<long long>(min(static_cast<long long>(n_dev[c]), n_host)) : n_host;
  const char* col = static_cast<const char*>(src) + static_cast<size_t>(c) * ld * 4;
  const unsigned int binmask = static_cast<unsigned int>(nbins - 1);
  // kHistUnroll independent loads per thread are issued before any of them is consumed: the passes over the
  // small L2-resident key buffers (sample, window) are latency-bound otherwise (one dependent
  // load -> match -> atomic chain per key).  Whole warps iterate together (base is warp-uniform).
  constexpr int kHistUnroll = 8;
  unsigned int run_bin = 0xFFFFFFFFu, run_cnt = 0u;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * kHistUnroll;
  for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x * kHistUnroll; base < n; base += stride) {
    unsigned int key[kHistUnroll];
    bool ok[kHistUnroll];
#pragma unroll
    for (int u = 0; u < kHistUnroll; ++u) {
      const long long i = base + static_cast<long long>(u) * blockDim.x + threadIdx.x;
      ok[u] = i < n;
      key[u] = ok[u] ? load_key<SRC>(col, i, med) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kHistUnroll; ++u) {
      unsigned int tgt = 0xFFFFFFFFu;
      if (ok[u]) {
        const unsigned int hi = key[u] & himask;
        const unsigned int bin = (key[u] >> shift) & binmask;
        if (hi == p0) tgt = bin;
        else if (!same && hi == p1) tgt = kBins + bin;
      }
      // run-length combining per thread: real signal columns put most keys of a pass into a handful of bins
      // (pass 0 sees 11 leading bits), so consecutive keys of a thread usually share their bin and cost no atomic
      if (tgt != run_bin) {
        if (run_bin != 0xFFFFFFFFu) atomicAdd(&sh[run_bin], run_cnt);
        run_bin = tgt;
        run_cnt = 0u;
      }
      ++run_cnt;
    }
  }
  if (run_bin != 0xFFFFFFFFu) atomicAdd(&sh[run_bin], run_cnt);
  __syncthreads();
  unsigned int* g = ghist + static_cast<size_t>(c) * 2 * kBins;
  for (int i = threadIdx.x; i < 2 * kBins; i += blockDim.x) {
    const unsigned int v = sh[i];
    if (v) atomicAdd(&g[i], v);
  }
  // the last block to finish a column picks its bins (threadfence reduction pattern)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(&done[c], 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  pick_column(pass, c, st, ghist, reinterpret_cast<unsigned long long*>(sh), wsum);
  if (threadIdx.x == 0) {
    done[c] = 0u;
    if (op == OP_WINDOW_BOUNDS || op == OP_WINDOW_BOUNDS_1P) arm_window(st, c);
    if (op == OP_WINDOW_BOUNDS_1P) arm_dev_sample(st, c);
    if (op == OP_DEV_BOUNDS) arm_dev_window(st, c);
  }
}

// Locate the bin holding each slot's rank, extend the prefix, clear the histogram.  Runs in the LAST block
// of hist_kernel to finish a column (all of the block's threads; `cum` is the block's histogram memory
// reused, `wsum` 32 more words), so a selection pass is one launch.
__device__ void pick_column(int pass, int c, SelState* st, unsigned int* ghist, unsigned long long* cum,
                            unsigned long long* wsum) {
  int shift, nbins;
  unsigned int himask;
  pass_bits(pass, shift, nbins, himask);
  const int was_same = st->same[c];
  const int nslots = was_same ? 1 : 2;
  // ranks are read by every thread up front: the thread that finds the bin rewrites them below
  const unsigned long long rank_in[2] = {st->rank[c][0], st->rank[c][1]};
  const int T = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int per = kBins / T;  // consecutive bins per thread (T divides kBins)
  __syncthreads();
  for (int s = 0; s < nslots; ++s) {
    unsigned int* g = ghist + (static_cast<size_t>(c) * 2 + s) * kBins;
    unsigned long long mine[8];
    unsigned long long x = 0ull;
    for (int i = 0; i < per; ++i) {
      const int bin = t * per + i;
      mine[i] = (bin < nbins) ? __ldcg(g + bin) : 0u;
      x += mine[i];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      unsigned long long y = (lane < T / 32) ? wsum[lane] : 0ull;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long z = __shfl_up_sync(0xffffffffu, y, o);
        if (lane >= o) y += z;
      }
      wsum[lane] = y;
    }
    __syncthreads();
    unsigned long long run = ((w > 0) ? wsum[w - 1] : 0ull) + x;  // inclusive up to this thread's last bin
    for (int i = per - 1; i >= 0; --i) {
      cum[t * per + i] = run;
      run -= mine[i];
    }
    __syncthreads();
    // slots that share this histogram
    const int s_end = was_same ? 2 : s + 1;
    for (int ss = s; ss < s_end; ++ss) {
      const unsigned long long r = rank_in[ss];
      for (int bin = t; bin < nbins; bin += T) {
        const unsigned long long lo = (bin > 0) ? cum[bin - 1] : 0ull;
        if (r >= lo && r < cum[bin]) {
          st->prefix[c][ss] = (st->prefix[c][ss] & himask) | (static_cast<unsigned int>(bin) << shift);
          st->rank[c][ss] = r - lo;
        }
      }
    }
    __syncthreads();
    for (int i = t; i < kBins; i += T) g[i] = 0u;
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (t == 0 && was_same && st->prefix[c][0] != st->prefix[c][1]) st->same[c] = 0;
}

// The two middle ranks of n sorted values (equal when n is odd): np.median averages them in fp32.
__device__ __forceinline__ void middle_ranks(long long n, unsigned long long& r0, unsigned long long& r1) {
  r0 = (n & 1) ? static_cast<unsigned long long>((n - 1) / 2) : static_cast<unsigned long long>(n / 2 - 1);
  r1 = (n & 1) ? r0 : static_cast<unsigned long long>(n / 2);
}

// Glue between selection stages (one thread per column).
//   ARM_FULL   : select the middle ranks of all n keys
//   ARM_SAMPLE : select the window bounds in the sample: ranks s/2 -+ margin
//   ARM_WINDOW : the sample selection just finished: prefix[] are the window bounds -> lo/hi
//   ARM_INSIDE : the window pass just finished: select ranks (middle - below) inside the window
//   FINAL_MED / FINAL_MAD : the selection of the middle ranks finished: prefix[] -> med / mad
//   FINAL_MED_1P : one-pass path: prefix[] -> med, then the exact-deviation window of the MAD selection is armed
enum { ARM_FULL, ARM_SAMPLE, ARM_WINDOW, ARM_INSIDE, FINAL_MED, FINAL_MAD, FINAL_MED_1P };

__global__ void glue_kernel(int op, int f, long long n, unsigned int window_cap, SelState* st) {
  const int c = threadIdx.x;
  if (c >= f) return;
  unsigned long long r0, r1;
  middle_ranks(n, r0, r1);
  if (op == ARM_FULL) {
    st->prefix[c][0] = st->prefix[c][1] = 0u;
    st->rank[c][0] = r0;
    st->rank[c][1] = r1;
    st->same[c] = 1;
  } else if (op == ARM_SAMPLE) {
    arm_sample(st, c);
  } else if (op == ARM_WINDOW) {
    st->lo[c] = st->prefix[c][0];
    st->hi[c] = st->prefix[c][1];
    st->below[c] = 0ull;
    st->wcnt[c] = 0u;
  } else if (op == ARM_INSIDE) {
    const unsigned long long below = st->below[c], cnt = st->wcnt[c];
    const bool ok = cnt <= window_cap && below <= r0 && r1 < below + cnt;
    if (!ok) st->miss[c] = 1;
    st->prefix[c][0] = st->prefix[c][1] = 0u;
    st->rank[c][0] = ok ? r0 - below : 0ull;
    st->rank[c][1] = ok ? r1 - below : 0ull;
    st->same[c] = 1;
  } else {
    const float v0 = from_orderable(st->prefix[c][0]);
    const float v1 = from_orderable(st->prefix[c][1]);
    // np.median: mean of the middle element(s) in float32 -> (v0 + v1) / 2, or v0 when n is odd
    const float m = (n & 1) ? v0 : __fmul_rn(__fadd_rn(v0, v1), 0.5f);
    if (op == FINAL_MED) st->med[c] = m; else st->mad[c] = m;
  }
}

// Keys of kSampleRuns evenly spaced runs of kSampleRunLen contiguous elements -> skeys[c][kSample], with the FIRST
// selection pass over the sample fused in: the block histograms the top 11 key bits of what it gathers, and the last
// block of a column picks the bins of the two sample ranks (the state must be armed: arm_sample).
template <int SRC>
__global__ void __launch_bounds__(256)
sample_kernel(const float* __restrict__ cols, long long n, long long ld, SelState* st, unsigned int* __restrict__ skeys,
              unsigned int* __restrict__ ghist, unsigned int* __restrict__ done, long long run_step, int run_rem) {
  __shared__ __align__(8) unsigned int sh[2 * kBins];   // slot-0 histogram; reused as 2048 x u64 by pick_column
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned int ticket;
  const int c = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * kBins; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const float med = (SRC == SRC_DEV) ? st->med[c] : 0.f;
  const float* col = cols + static_cast<size_t>(c) * ld;
  unsigned int run_bin = 0xFFFFFFFFu, run_cnt = 0u;
  // kSampleBlockRuns runs per block and trip: all their loads (16 per thread) are in flight before the first is used --
  // the gather is a handful of dependent DRAM round trips otherwise
  constexpr int kPer = kSampleRunLen / 256;
  for (int run0 = blockIdx.x * kSampleBlockRuns; run0 < kSampleRuns; run0 += gridDim.x * kSampleBlockRuns) {
    unsigned int key[kSampleBlockRuns * kPer];
#pragma unroll
    for (int r = 0; r < kSampleBlockRuns; ++r) {
      const long long run = min(run0 + r, kSampleRuns - 1);
      // run * (n - L) / (R - 1) with the quotient and remainder of (n - L) / (R - 1) from the host (no wide division here)
      const long long start = run * run_step + (run * run_rem) / (kSampleRuns - 1);
#pragma unroll
      for (int u = 0; u < kPer; ++u) key[r * kPer + u] = load_key<SRC>(col, start + u * 256 + threadIdx.x, med);
    }
#pragma unroll
    for (int r = 0; r < kSampleBlockRuns; ++r) {
      if (run0 + r >= kSampleRuns) break;
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const unsigned int kk = key[r * kPer + u];
        skeys[static_cast<size_t>(c) * kSample + (run0 + r) * kSampleRunLen + u * 256 + threadIdx.x] = kk;
        const unsigned int bin = kk >> 21;   // pass 0: bits 31..21
        if (bin != run_bin) {
          if (run_bin != 0xFFFFFFFFu) atomicAdd(&sh[run_bin], run_cnt);
          run_bin = bin;
          run_cnt = 0u;
        }
        ++run_cnt;
      }
    }
  }
  if (run_bin != 0xFFFFFFFFu) atomicAdd(&sh[run_bin], run_cnt);
  __syncthreads();
  unsigned int* g = ghist + static_cast<size_t>(c) * 2 * kBins;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
    const unsigned int v = sh[i];
    if (v) atomicAdd(&g[i], v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(&done[c], 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  pick_column(0, c, st, ghist, reinterpret_cast<unsigned long long*>(sh), wsum);
  if (threadIdx.x == 0) done[c] = 0u;
}

// THE full pass of the windowed path: count keys below the window, collect the keys inside it.
//
// The pass is a pure HBM stream, so the per-element instruction count is what has to stay small:
// the window bounds are turned back into floats and every element costs four instructions --
// `p = x >= lo` (FSETP), `below += !p`, `q = p && x <= hi` (FSETP.AND), `mask |= q << e`.  Window hits
// are ~0.8 % of the keys, i.e. ~4 per warp per 16-element tile, so nearly every warp has SOME lane with
// a hit: the hit path must be short too.  Lanes with a non-zero mask walk its set bits, re-read the
// element (an L1 / L2 hit), and append its order-preserving key to a per-block shared buffer with one
// shared atomic each; the buffer goes to the global window with one global atomic when it is half full.
// Float order and key order differ only in (-0, +0) and NaNs: counting and collecting use the SAME float
// predicates and the selection inside the window uses key order (a refinement of float order), so the
// selected VALUE is exact.  The [column][row] space is flattened into 4096-element tiles and cut into
// one contiguous tile range per block, so a grid of exactly (SMs x resident blocks) is one balanced wave
// for any column count.
constexpr int kWinBuf = 3072;
constexpr int kWinThreads = 256;
#ifndef DEWI_WIN_PER_THREAD
#define DEWI_WIN_PER_THREAD 16
#endif
#ifndef DEWI_WIN_BLOCKS_PER_SM
#define DEWI_WIN_BLOCKS_PER_SM 8
#endif
constexpr int kWinPerThread = DEWI_WIN_PER_THREAD;
constexpr int kWinTile = kWinThreads * kWinPerThread;
constexpr int kFlushCheckEvery = 8;   // tiles between two looks at the buffer fill (8 tiles add ~260 keys)
constexpr int kWinBlocksPerSm = DEWI_WIN_BLOCKS_PER_SM;

// Inclusive prefix sums of g[0 .. kBins) into cum[] for the whole block (blockDim.x divides kBins, at most 8 bins
// per thread); `wsum` is 32 words of scratch.
__device__ void scan_bins(const unsigned int* g, unsigned long long* cum, unsigned long long* wsum) {
  const int T = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int per = kBins / T;
  unsigned long long mine[8];
  unsigned long long x = 0ull;
  for (int i = 0; i < per; ++i) {
    mine[i] = __ldcg(g + t * per + i);
    x += mine[i];
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  if (w == 0) {
    unsigned long long y = (lane < T / 32) ? wsum[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long z = __shfl_up_sync(0xffffffffu, y, o);
      if (lane >= o) y += z;
    }
    wsum[lane] = y;
  }
  __syncthreads();
  unsigned long long run = ((w > 0) ? wsum[w - 1] : 0ull) + x;
  for (int i = per - 1; i >= 0; --i) {
    cum[t * per + i] = run;
    run -= mine[i];
  }
  __syncthreads();
}

// Window-relative digit of a key the float predicates admitted: (key - lo) clamped into [0, hi - lo] (float order
// and key order disagree on -0 / +0 only), top <= 11 bits.
__device__ __forceinline__ unsigned int window_digit(unsigned int key, unsigned int lo, unsigned int span, int wshift) {
  const unsigned int rel = key >= lo ? min(key - lo, span) : 0u;
  return rel >> wshift;
}

// Shared state of a window_kernel block that its out-of-line helpers need.
struct WinShared {
  // [ buf : keys waiting for the next flush | hist : window-relative digit histogram of the current column ]; the
  // block that finishes a column last reuses the whole array as 2048 x u64 prefix sums (window_pick)
  unsigned int shm[kWinBuf + kBins];
  unsigned long long wsum[32];
  unsigned int ticket;
  // digit parameters of the current column (written by thread 0 at a column switch)
  unsigned int lo_key, span;
  int wshift;
};

// The block that completes a column (all its tiles accounted for) turns the column's counts into the two target
// bins of the in-window selection: ranks (middle - below) inside the collected keys, located in the histogram of
// the window-relative digit.  A window that overflowed or does not hold the ranks is flagged (miss).
__device__ __forceinline__ void window_pick(WinShared* ws, int c, long long n, unsigned int window_cap, SelState* st, unsigned int* g) {
  unsigned long long r0, r1;
  middle_ranks(n, r0, r1);
  const unsigned long long nbelow = __ldcg(&st->below[c]), cnt = __ldcg(&st->wcnt[c]);
  const bool ok = cnt <= window_cap && nbelow <= r0 && r1 < nbelow + cnt;
  if (ok) {
    unsigned long long* cum = reinterpret_cast<unsigned long long*>(ws->shm);
    scan_bins(g, cum, ws->wsum);
    const unsigned long long rr[2] = {r0 - nbelow, r1 - nbelow};
    for (int ss = 0; ss < 2; ++ss) {
      for (int bin = threadIdx.x; bin < kBins; bin += blockDim.x) {
        const unsigned long long before = bin ? cum[bin - 1] : 0ull;
        if (rr[ss] >= before && rr[ss] < cum[bin]) {
          st->wbin[c][ss] = static_cast<unsigned int>(bin);
          st->wrank[c][ss] = rr[ss] - before;
        }
      }
    }
    __syncthreads();
  } else if (threadIdx.x == 0) {
    st->miss[c] = 1;
  }
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) g[i] = 0u;                    // the histogram memory is reused
  for (int i = threadIdx.x; i < kWinBuf + kBins; i += blockDim.x) ws->shm[i] = 0u;   // (clobbered by the prefix sums)
  if (threadIdx.x == 0) st->tiles_done[c] = 0u;
  __syncthreads();
}

// End of a column for this block (block-uniform; the key buffer has been flushed): hand the digit histogram over,
// draw the column's tile ticket and, if that completes the column, pick the target bins.  Out of line: it runs once
// or twice per block and must not cost the streaming loop registers.
__device__ __forceinline__ void window_finish_column(WinShared* ws, int c, unsigned int tiles_mine, int tiles_per_col, long long n,
                                                  unsigned int window_cap, SelState* st, unsigned int* ghist) {
  unsigned int* g = ghist + static_cast<size_t>(c) * 2 * kBins;
  unsigned int* hist = ws->shm + kWinBuf;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
    const unsigned int v = hist[i];
    if (v) { atomicAdd(&g[i], v); hist[i] = 0u; }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ws->ticket = atomicAdd(&st->tiles_done[c], tiles_mine);
  __syncthreads();
  if (ws->ticket + tiles_mine == static_cast<unsigned int>(tiles_per_col)) {
    __threadfence();
    window_pick(ws, c, n, window_cap, st, g);
  }
  __syncthreads();
}

// Load the digit parameters of column c into the block's shared state (block-uniform).
__device__ __forceinline__ void window_enter_column(WinShared* ws, int c, const SelState* st) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ws->lo_key = st->lo[c];
    ws->span = st->hi[c] - st->lo[c];
    ws->wshift = st->wshift[c];
  }
  __syncthreads();
}

// FUSED = 1: the histogram of the window-relative digit and the pick of the target bins ride along (the digits are
// counted when the key buffer is flushed; the block that completes a column picks).  FUSED = 0: the pass only counts
// and collects, and window_hist_kernel histograms the collected keys in a launch of its own.
template <int SRC, int FUSED>
__global__ void __launch_bounds__(kWinThreads, kWinBlocksPerSm)
window_kernel(const float* __restrict__ cols, long long n, long long ld, int f, SelState* st,
              unsigned int* __restrict__ wkeys, unsigned int window_cap, unsigned int* __restrict__ ghist) {
  __shared__ __align__(8) WinShared ws;
  unsigned int* const buf = ws.shm;
  __shared__ unsigned int buf_n, flush_base;
  __shared__ unsigned long long below_blk;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) ws.shm[kWinBuf + i] = 0u;
  const int lane = threadIdx.x & 31;
  // (tile counts fit 32 bits: 2^31 rows x 32 columns / 4096)
  const int tiles_per_col = static_cast<int>((n + kWinTile - 1) / kWinTile);
  const long long total_tiles = static_cast<long long>(tiles_per_col) * f;
  const int t_begin = static_cast<int>((total_tiles * blockIdx.x) / gridDim.x);
  const int t_end = static_cast<int>((total_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (threadIdx.x == 0) { buf_n = 0u; below_blk = 0ull; }
  __syncthreads();

  int c = -1;
  float med = 0.f, lo_f = 0.f, hi_f = 0.f;
  const float* col = nullptr;
  unsigned int* out = nullptr;
  bool aligned = false;
  unsigned int below = 0u;
  int since_check = 0;

  auto flush = [&](unsigned int min_fill) {  // block-uniform
    __syncthreads();
    const unsigned int cnt = min(buf_n, static_cast<unsigned int>(kWinBuf));
    if (cnt > min_fill) {
      if (threadIdx.x == 0) flush_base = atomicAdd(&st->wcnt[c], cnt);
      __syncthreads();
      for (unsigned int k = threadIdx.x; k < cnt; k += blockDim.x) {
        const unsigned int key = buf[k];
        if (flush_base + k < window_cap) out[flush_base + k] = key;
        if (FUSED) atomicAdd(&ws.shm[kWinBuf + window_digit(key, ws.lo_key, ws.span, ws.wshift)], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) buf_n = 0u;
      __syncthreads();
    }
    since_check = 0;
  };
  auto commit_column = [&]() {  // block-uniform: hand this column's counts over before switching
    flush(0u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0 && below) atomicAdd(&below_blk, static_cast<unsigned long long>(below));
    __syncthreads();
    if (threadIdx.x == 0) {
      if (below_blk) atomicAdd(&st->below[c], below_blk);
      below_blk = 0ull;
    }
    below = 0u;
    __syncthreads();
    if (FUSED) {
      // tiles of column c inside this block's range [t_begin, t_end)
      // (the block's first tile is recomputed here rather than kept in a register across the streaming loop)
      const int c_lo = c * tiles_per_col;
      const int first = static_cast<int>((static_cast<long long>(tiles_per_col) * f * blockIdx.x) / gridDim.x);
      window_finish_column(&ws, c, static_cast<unsigned int>(min(t_end, c_lo + tiles_per_col) - max(first, c_lo)), tiles_per_col,
                           n, window_cap, st, ghist);
    }
  };
  auto push = [&](float v) {  // v lies in the window
    const unsigned int key = orderable(v);
    const unsigned int slot = atomicAdd(&buf_n, 1u);
    if (slot < kWinBuf) buf[slot] = key;   // (its digit is histogrammed when the buffer is flushed)
    else {  // shared buffer full (a huge tie group): straight to global; the caller will see wcnt > cap
      const unsigned int g = atomicAdd(&st->wcnt[c], 1u);
      if (g < window_cap) out[g] = key;
      if (FUSED) atomicAdd(&ws.shm[kWinBuf + window_digit(key, ws.lo_key, ws.span, ws.wshift)], 1u);
    }
  };

  for (int t = t_begin; t < t_end; ++t) {
    const int tc = t / tiles_per_col;
    if (tc != c) {
      if (c >= 0) commit_column();
      c = tc;
      if (FUSED) window_enter_column(&ws, c, st);
      med = (SRC == SRC_DEV) ? st->med[c] : 0.f;
      lo_f = from_orderable(st->lo[c]);
      hi_f = from_orderable(st->hi[c]);
      col = cols + static_cast<size_t>(c) * ld;
      out = wkeys + static_cast<size_t>(c) * window_cap;
      aligned = (reinterpret_cast<uintptr_t>(col) & 15) == 0;
    }
    const long long base = static_cast<long long>(t - c * tiles_per_col) * kWinTile;
    if (aligned && base + kWinTile <= n) {
      // element e of this thread: base + (e / 4) * (threads * 4) + tid * 4 + (e % 4)  (coalesced 128-bit loads)
      const float* p = col + base + threadIdx.x * 4;
      float x[kWinPerThread];
#pragma unroll
      for (int q = 0; q < kWinPerThread / 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + q * kWinThreads * 4));
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
      }
      unsigned int mask = 0u;
#pragma unroll
      for (int e = 0; e < kWinPerThread; ++e) {
        float y = x[e];
        if (SRC == SRC_DEV) y = fabsf(__fsub_rn(y, med));
        const bool ge = y >= lo_f;
        below += ge ? 0u : 1u;
        mask |= (ge && y <= hi_f) ? (1u << e) : 0u;
      }
      while (mask) {  // rare per lane; the element is re-read (cache hit) instead of indexing registers dynamically
        const int e = __ffs(mask) - 1;
        mask &= mask - 1u;
        float y = p[(e >> 2) * (kWinThreads * 4) + (e & 3)];
        if (SRC == SRC_DEV) y = fabsf(__fsub_rn(y, med));
        push(y);
      }
    } else {  // ragged last tile of a column, or a column that is not 16-byte aligned
      for (int e = 0; e < kWinPerThread; ++e) {
        const long long i = base + static_cast<long long>(e >> 2) * kWinThreads * 4 + threadIdx.x * 4 + (e & 3);
        if (i < n) {
          float y = __ldg(col + i);
          if (SRC == SRC_DEV) y = fabsf(__fsub_rn(y, med));
          const bool ge = y >= lo_f;
          below += ge ? 0u : 1u;
          if (ge && y <= hi_f) push(y);
        }
      }
    }
    if (++since_check == kFlushCheckEvery) flush(kWinBuf / 2);
  }
  if (c >= 0) commit_column();
}

// THE pass of the one-pass path: per element y, the median window [lo, hi] as in window_kernel (count below, collect
// inside) AND the deviation window A <= |y - c0| <= B around the centre of the median window (count below A, collect
// inside): one read of the column instead of two.  The MAD's keys are stored as keys of the VALUES: the exact median
// is not known yet.
//
// The pass is bound by instruction issue, not by HBM, unless the hits (~2.8 % of the elements, a dozen per warp and
// tile) stay off the streaming loop: ncu on the first version -- every hit re-read, re-keyed and appended on the spot in
// a divergent loop -- showed 681M warp instructions (500 per 512 elements) at 70 % issue utilisation and 40 % of the
// DRAM peak.  So (1) both windows are tested in centre / radius form on t = fl(y - c0), which they share: |t| <= r0
// for the median window, | |t| - cd | <= rd for the deviation window, t < -r0 and |t| - cd < -rd for the two "below"
// counts; the tests are monotone in y, so they define proper windows, and the counts and the collection partition
// the column consistently because both use the same rounded quantities.  (2) The loop only notes WHETHER any of a
// thread's 16 elements hit and queues the thread's element index -- one warp-aggregated shared atomic per warp and
// tile; the entries are re-tested and expanded into keys when the queue is flushed, every thread working on a
// different entry (dense, no divergence against the streaming loop).
constexpr int kWin2BlocksPerSm = 6;
constexpr int kWin2Queue = 2048;   // queued entries per block (one per thread and tile with a hit; ~80 per tile)

template <int DUMMY>
__global__ void __launch_bounds__(kWinThreads, kWin2BlocksPerSm)
window2_kernel(const float* __restrict__ cols, long long n, long long ld, int f, SelState* st, unsigned int* __restrict__ wkeys0,
               unsigned int cap0, unsigned int* __restrict__ wkeysd, unsigned int capd) {
  __shared__ unsigned int q_elem[kWin2Queue];   // index (within the column) of the queued thread's first element
  __shared__ unsigned int q_mask[kWin2Queue];   // flush: bit e = element e in the median window, bit 16 + e = in the deviation window
  __shared__ unsigned int warp_sum[2][kWinThreads / 32];
  __shared__ unsigned int q_n, base0_s, based_s;
  __shared__ unsigned long long below0_blk, belowd_blk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_per_col = static_cast<int>((n + kWinTile - 1) / kWinTile);
  const long long total_tiles = static_cast<long long>(tiles_per_col) * f;
  const int t_begin = static_cast<int>((total_tiles * blockIdx.x) / gridDim.x);
  const int t_end = static_cast<int>((total_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (threadIdx.x == 0) { q_n = 0u; below0_blk = 0ull; belowd_blk = 0ull; }
  __syncthreads();

  int c = -1;
  float c0 = 0.f, r0 = 0.f, cd = 0.f, rd = 0.f;
  const float* col = nullptr;
  bool aligned = false;
  unsigned int below0 = 0u, belowd = 0u;
  int since_check = 0;

  // The two window tests of one element (the streaming loop evaluates the same expressions).
  auto in0 = [&](float y) { return fabsf(__fsub_rn(y, c0)) <= r0; };
  auto ind = [&](float y) { return fabsf(__fsub_rn(fabsf(__fsub_rn(y, c0)), cd)) <= rd; };

  // Expand the queued entries into keys (block-uniform): re-test each entry's 16 elements, count the hits per window,
  // reserve both ranges with one global atomic each, then every thread writes the keys of its entries at its scanned offset.
  auto flush = [&]() {
    __syncthreads();
    const unsigned int cnt = min(q_n, static_cast<unsigned int>(kWin2Queue));
    unsigned int my0 = 0u, myd = 0u;
    for (unsigned int k = threadIdx.x; k < cnt; k += blockDim.x) {
      const float* p = col + q_elem[k];
      unsigned int m = 0u;
#pragma unroll
      for (int e = 0; e < kWinPerThread; ++e) {
        const float y = __ldg(p + (e >> 2) * (kWinThreads * 4) + (e & 3));   // (cache hit: read a few tiles ago)
        m |= in0(y) ? (1u << e) : 0u;
        m |= ind(y) ? (0x10000u << e) : 0u;
      }
      q_mask[k] = m;
      my0 += __popc(m & 0xFFFFu);
      myd += __popc(m >> 16);
    }
    unsigned int inc0 = my0, incd = myd;   // inclusive warp scans
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int y0 = __shfl_up_sync(0xffffffffu, inc0, o), yd = __shfl_up_sync(0xffffffffu, incd, o);
      if (lane >= o) { inc0 += y0; incd += yd; }
    }
    if (lane == 31) { warp_sum[0][warp] = inc0; warp_sum[1][warp] = incd; }
    __syncthreads();
    unsigned int off0 = inc0 - my0, offd = incd - myd, tot0 = 0u, totd = 0u;
#pragma unroll
    for (int w2 = 0; w2 < kWinThreads / 32; ++w2) {
      if (w2 < warp) { off0 += warp_sum[0][w2]; offd += warp_sum[1][w2]; }
      tot0 += warp_sum[0][w2];
      totd += warp_sum[1][w2];
    }
    if (threadIdx.x == 0) {
      base0_s = tot0 ? atomicAdd(&st->wcnt[c], tot0) : 0u;
      based_s = totd ? atomicAdd(&st->cntd[c], totd) : 0u;
    }
    __syncthreads();
    unsigned int* out0 = wkeys0 + static_cast<size_t>(c) * cap0;
    unsigned int* outd = wkeysd + static_cast<size_t>(c) * capd;
    unsigned int p0 = base0_s + off0, pd = based_s + offd;
    for (unsigned int k = threadIdx.x; k < cnt; k += blockDim.x) {
      unsigned int m = q_mask[k];
      const float* p = col + q_elem[k];
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1u;
        const int e = bit & 15;
        const unsigned int key = orderable(__ldg(p + (e >> 2) * (kWinThreads * 4) + (e & 3)));   // (L2 hit: read a few tiles ago)
        if (bit < 16) { if (p0 < cap0) out0[p0] = key; ++p0; }
        else { if (pd < capd) outd[pd] = key; ++pd; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) q_n = 0u;
    __syncthreads();
    since_check = 0;
  };
  auto commit_column = [&]() {  // block-uniform
    flush();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      below0 += __shfl_xor_sync(0xffffffffu, below0, o);
      belowd += __shfl_xor_sync(0xffffffffu, belowd, o);
    }
    if (lane == 0) {
      if (below0) atomicAdd(&below0_blk, static_cast<unsigned long long>(below0));
      if (belowd) atomicAdd(&belowd_blk, static_cast<unsigned long long>(belowd));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (below0_blk) atomicAdd(&st->below[c], below0_blk);
      if (belowd_blk) atomicAdd(&st->belowd[c], belowd_blk);
      below0_blk = 0ull;
      belowd_blk = 0ull;
    }
    below0 = 0u;
    belowd = 0u;
    __syncthreads();
  };
  // Ragged tile / unaligned column: per-element path, hits go straight to the global buffers (rare).
  auto slow_hit = [&](float y, bool dev) {
    const unsigned int g = atomicAdd(dev ? &st->cntd[c] : &st->wcnt[c], 1u);
    if (dev) { if (g < capd) wkeysd[static_cast<size_t>(c) * capd + g] = orderable(y); }
    else if (g < cap0) wkeys0[static_cast<size_t>(c) * cap0 + g] = orderable(y);
  };

  int next_switch = t_begin;   // first tile of the next column
  for (int t = t_begin; t < t_end; ++t) {
    if (t >= next_switch) {
      if (c >= 0) commit_column();
      c = t / tiles_per_col;
      next_switch = (c + 1) * tiles_per_col;
      c0 = st->c0[c];
      r0 = st->r0[c];
      cd = st->cd[c];
      rd = st->rd[c];
      col = cols + static_cast<size_t>(c) * ld;
      aligned = (reinterpret_cast<uintptr_t>(col) & 15) == 0;
    }
    const long long base = static_cast<long long>(t - (next_switch - tiles_per_col)) * kWinTile;
    if (aligned && base + kWinTile <= n) {
      const unsigned int first = static_cast<unsigned int>(base) + threadIdx.x * 4;
      const float* p = col + first;
      float x[kWinPerThread];
#pragma unroll
      for (int q = 0; q < kWinPerThread / 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + q * kWinThreads * 4));
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
      }
      bool any = false;
      const float nr0 = -r0, nrd = -rd;
#pragma unroll
      for (int e = 0; e < kWinPerThread; ++e) {
        const float t = __fsub_rn(x[e], c0);
        below0 += (t < nr0) ? 1u : 0u;
        any |= fabsf(t) <= r0;
        const float v = __fsub_rn(fabsf(t), cd);
        belowd += (v < nrd) ? 1u : 0u;
        any |= fabsf(v) <= rd;
      }
      const unsigned int hit = __ballot_sync(0xffffffffu, any);
      if (hit) {  // warp-uniform: one shared atomic reserves the warp's entries
        unsigned int slot = 0u;
        if (lane == 0) slot = atomicAdd(&q_n, static_cast<unsigned int>(__popc(hit)));
        slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(hit & ((1u << lane) - 1u));
        if (any) {
          if (slot < kWin2Queue) q_elem[slot] = first;
          else {  // queue full (cannot happen between two flush checks): expand here
            for (int e = 0; e < kWinPerThread; ++e) {
              const float y = p[(e >> 2) * (kWinThreads * 4) + (e & 3)];
              if (in0(y)) slow_hit(y, false);
              if (ind(y)) slow_hit(y, true);
            }
          }
        }
      }
    } else {  // ragged last tile of a column, or a column that is not 16-byte aligned
      for (int e = 0; e < kWinPerThread; ++e) {
        const long long i = base + static_cast<long long>(e >> 2) * kWinThreads * 4 + threadIdx.x * 4 + (e & 3);
        if (i < n) {
          const float y = __ldg(col + i);
          const float t = __fsub_rn(y, c0);
          below0 += (t < -r0) ? 1u : 0u;
          if (in0(y)) slow_hit(y, false);
          belowd += (__fsub_rn(fabsf(t), cd) < -rd) ? 1u : 0u;
          if (ind(y)) slow_hit(y, true);
        }
      }
    }
    // every thread may queue one entry per tile: flush while a whole round of tiles still fits
    if (++since_check == 4) {
      // (the barrier makes the decision block-uniform although the warps read the fill count at different moments)
      if (__syncthreads_or(q_n > kWin2Queue - 4 * kWinThreads)) flush(); else since_check = 0;
    }
  }
  if (c >= 0) commit_column();
}

// The histogram of the window-relative digit over the collected keys (L2-resident), one launch for all columns; the
// last block of a column picks the target bins (window_pick).  DEV = 1 (one-pass path, MAD): the buffer holds keys of
// VALUES collected around c0; each is re-keyed as the exact deviation |y - med|, counted as "below" when under the
// validity window [lo, hi] armed by the median's survivor pass, histogrammed when inside, ignored when above.
constexpr int kWhistThreads = 512;

template <int DEV>
__global__ void __launch_bounds__(kWhistThreads)
window_hist_kernel(const unsigned int* __restrict__ wkeys, unsigned int window_cap, long long n, SelState* st,
                   unsigned int* __restrict__ ghist) {
  __shared__ __align__(8) WinShared ws;
  __shared__ unsigned int below_s, in_s;
  const int c = blockIdx.y;
  if (DEV && st->miss[c]) return;   // (uniform over the column's blocks)
  unsigned int* hist = ws.shm + kWinBuf;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0u;
  if (threadIdx.x == 0) { below_s = 0u; in_s = 0u; }
  __syncthreads();
  const unsigned int lo = st->lo[c], hi = st->hi[c], span = hi - lo;
  const int wshift = st->wshift[c];
  const float med = DEV ? st->med[c] : 0.f;
  const unsigned int cnt = min(__ldcg(DEV ? &st->cntd[c] : &st->wcnt[c]), window_cap);
  unsigned int my_below = 0u, my_in = 0u;
  const unsigned int* src = wkeys + static_cast<size_t>(c) * window_cap;
  constexpr int kU = 8;   // independent loads in flight per thread (the pass is latency-bound otherwise)
  const unsigned int stride = gridDim.x * blockDim.x * kU;
  for (unsigned int base = blockIdx.x * blockDim.x * kU; base < cnt; base += stride) {
    unsigned int key[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const unsigned int i = base + u * blockDim.x + threadIdx.x;
      key[u] = i < cnt ? __ldg(src + i) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (base + u * blockDim.x + threadIdx.x >= cnt) continue;
      unsigned int k = key[u];
      if (DEV) {
        k = orderable(fabsf(__fsub_rn(key_to_float(k), med)));
        if (k < lo) { ++my_below; continue; }
        if (k > hi) continue;
        ++my_in;
      }
      atomicAdd(&hist[window_digit(k, lo, span, wshift)], 1u);
    }
  }
  if (DEV) {
    if (my_below) atomicAdd(&below_s, my_below);
    if (my_in) atomicAdd(&in_s, my_in);
  }
  __syncthreads();
  if (DEV && threadIdx.x == 0) {   // (before the ticket: the picking block must see every block's counts)
    if (below_s) atomicAdd(&st->below[c], static_cast<unsigned long long>(below_s));
    if (in_s) atomicAdd(&st->wcnt[c], in_s);
  }
  unsigned int* g = ghist + static_cast<size_t>(c) * 2 * kBins;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
    const unsigned int v = hist[i];
    if (v) atomicAdd(&g[i], v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ws.ticket = atomicAdd(&st->tiles_done[c], 1u);
  __syncthreads();
  if (ws.ticket != gridDim.x - 1) return;
  __threadfence();
  window_pick(&ws, c, n, window_cap, st, g);   // (also resets tiles_done[c] and the histogram)
}

// In-window selection, second half.  window_kernel left, per column, the digits (bins) of the two middle ranks and
// their ranks inside those bins; a bin of the window-relative top digit holds ~1/2048 of the window's keys -- a
// few hundred.  This pass re-reads the window keys (L2-resident), gathers the keys of the target bins, and the
// last block of a column sorts them in shared memory and reads off the two middle values: the median (then the
// sample selection of the MAD is armed) or the MAD.  More than kSurvCap survivors (a large tie group inside the
// bin) flags the column `miss` and the caller falls back to the plain radix passes.
constexpr int kSurvThreads = 512;

template <int DEV>
__global__ void __launch_bounds__(kSurvThreads)
survivor_kernel(const unsigned int* __restrict__ wkeys, unsigned int window_cap, long long n, SelState* st,
                unsigned int* __restrict__ surv, int final_op, unsigned int capd) {
  __shared__ unsigned int keys[kSurvCap];
  __shared__ unsigned int ticket;
  const int c = blockIdx.y;
  if (st->miss[c]) return;   // (uniform over the column's blocks: set by an earlier kernel)
  const unsigned int lo = st->lo[c], hi = st->hi[c], span = hi - lo;
  const int wshift = st->wshift[c];
  const float med = DEV ? st->med[c] : 0.f;
  const unsigned int b0 = st->wbin[c][0], b1 = st->wbin[c][1];
  const unsigned int cnt = min(DEV ? st->cntd[c] : st->wcnt[c], window_cap);
  const unsigned int* src = wkeys + static_cast<size_t>(c) * window_cap;
  unsigned int* dst = surv + static_cast<size_t>(c) * kSurvCap;
  constexpr int kU = 8;   // independent loads in flight per thread (the pass is latency-bound otherwise)
  const unsigned int stride = gridDim.x * blockDim.x * kU;
  for (unsigned int base = blockIdx.x * blockDim.x * kU; base < cnt; base += stride) {
    unsigned int key[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const unsigned int i = base + u * blockDim.x + threadIdx.x;
      key[u] = i < cnt ? __ldg(src + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (base + u * blockDim.x + threadIdx.x >= cnt) continue;
      if (DEV) {   // exact deviation of the collected value; only the validity window takes part
        key[u] = orderable(fabsf(__fsub_rn(key_to_float(key[u]), med)));
        if (key[u] < lo || key[u] > hi) continue;
      }
      const unsigned int d = window_digit(key[u], lo, span, wshift);
      if (d == b0 || d == b1) {
        const unsigned int slot = atomicAdd(&st->surv_n[c], 1u);
        if (slot < kSurvCap) dst[slot] = key[u];
        if (d == b0 && b0 != b1) atomicAdd(&st->surv_n0[c], 1u);
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(&st->surv_done[c], 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  const unsigned int m = __ldcg(&st->surv_n[c]), m0 = __ldcg(&st->surv_n0[c]);
  const unsigned long long i0 = st->wrank[c][0];
  const unsigned long long i1 = (b0 == b1) ? st->wrank[c][1] : static_cast<unsigned long long>(m0) + st->wrank[c][1];
  if (m > kSurvCap || i0 >= m || i1 >= m) {
    if (threadIdx.x == 0) st->miss[c] = 1;
  } else {
    int p = 1;
    while (p < static_cast<int>(m)) p <<= 1;
    for (int t = threadIdx.x; t < p; t += blockDim.x) keys[t] = t < static_cast<int>(m) ? __ldcg(dst + t) : 0xFFFFFFFFu;
    for (int size = 2; size <= p; size <<= 1) {       // ascending bitonic sort
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        __syncthreads();
        for (int t = threadIdx.x; t < (p >> 1); t += blockDim.x) {
          const int a = 2 * t - (t & (stride - 1)), b = a + stride;
          const bool asc = (a & size) == 0;
          const unsigned int ka = keys[a], kb = keys[b];
          if ((ka > kb) == asc) { keys[a] = kb; keys[b] = ka; }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float v0 = from_orderable(keys[i0]), v1 = from_orderable(keys[i1]);
      // np.median: mean of the middle element(s) in float32 -> (v0 + v1) / 2, or v0 when n is odd
      const float mid = (n & 1) ? v0 : __fmul_rn(__fadd_rn(v0, v1), 0.5f);
      if (final_op == FINAL_MED) {
        st->med[c] = mid;
        arm_sample(st, c);   // the MAD's sample selection follows
      } else if (final_op == FINAL_MED_1P) {
        // One-pass path: the MAD's keys are already collected (values with A <= |y - c0| <= B; `belowd` elements have
        // |y - c0| < A).  With delta = |c0 - med| every uncollected element has an exact deviation below L = A + delta
        // or above U = B - delta, so for x in [L, U]: #(deviation < x) = belowd + #(collected with deviation < x) --
        // the selection among the collected keys inside [L, U] is exact.  (Slack: 2^-18 of the magnitudes involved.)
        st->med[c] = mid;
        const float c0 = st->c0[c], a = st->A[c], bb = st->B[c];
        const float delta = fabsf(c0 - mid) * 1.0000038146972656f + (fabsf(c0) + fabsf(mid) + bb) * 3.814697265625e-06f;
        const float L = a + delta, U = bb - delta;
        const unsigned int lk = orderable(L), uk = orderable(U);
        if (!(U >= L) || !(U < INFINITY) || st->cntd[c] > capd) {
          st->miss[c] = 1;
        } else {
          st->lo[c] = lk;
          st->hi[c] = uk;
          const unsigned int sp = uk - lk;
          const int nbits = sp ? 32 - __clz(static_cast<int>(sp)) : 1;
          st->wshift[c] = max(nbits - kWindowDigitBits, 0);
          st->below[c] = st->belowd[c];
          st->wcnt[c] = 0u;   // recounted by the deviation histogram: keys inside [L, U]
        }
      } else {
        st->mad[c] = mid;
      }
    }
  }
  if (threadIdx.x == 0) st->surv_n[c] = st->surv_n0[c] = st->surv_done[c] = 0u;
}

struct ScoreParams {
  double med[7];
  double k[7];     // weight of (v_c - med_c) in U: the component weights, the 0.5 of Ht / Hi and 1 / (1.4826 * mad_c) folded
  double delta;
};

// exp(x) for -700 <= x <= 709.78 (U is clipped to +-delta first; the caller handles the tails): round-to-nearest
// range reduction and a degree-9 Taylor polynomial, relative error 1e-11 -- a third of the float64 instructions of
// the library exp.  The score kernel is otherwise bound by the float64 pipe, not by HBM: ~90 float64 instructions
// per row at 1.7e11 rows/s.  2^k is applied in two halves so that k = 1024 (x just below the overflow point) works.
__device__ __forceinline__ double exp_small(double x) {
  const double t = fma(x, 1.4426950408889634, 6755399441055744.0);  // round(x * log2 e) in the low mantissa bits
  const int k = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(-kf, 6.93147180369123816490e-01, x);
  r = fma(-kf, 1.90821492927058770002e-10, r);
  double p = 1.0 / 362880.0;
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int k1 = k >> 1, k2 = k - k1;
  return p * __hiloint2double((k1 + 1023) << 20, 0) * __hiloint2double((k2 + 1023) << 20, 0);  // * 2^k
}

// U = sum_c k_c (v_c - med_c) is scorer.py:53-57,67-73 (or :80-87) with the constant factors folded; it differs
// from the reference's operation order by a few ulp of float64 -- the gate is 1e-6 relative (measured 8e-12 for
// float64 output, 9e-8 for float32 output, whose final division runs in float32).
template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
score_kernel(const InT* __restrict__ cols, long long n, long long ld, const ScoreParams p, OutT* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double U = 0.0;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      const double v = static_cast<double>(__ldg(cols + static_cast<size_t>(c) * ld + i));
      U = fma(p.k[c], v - p.med[c], U);  // scorer.py:31 `(val - med) / (1.4826 * mad)`, weighted
    }
    // scorer.py:74 np.clip == minimum(maximum(U, -delta), delta).  CUDA's fmin / fmax drop a NaN operand where
    // numpy propagates it: a row with a NaN signal (or a NaN delta) must score NaN, not sigmoid(-delta).
    const bool is_nan = (U != U) || (p.delta != p.delta);
    U = fmin(fmax(U, -p.delta), p.delta);
    // scorer.py:62: 1 / (1 + exp(-U)); any delta is accepted, as in the reference: exp overflows to +inf beyond
    // 709.78 (score 0) and 1 + exp(x) == 1 below -700 (score 1)
    const double x = -U;
    double e = (x > 709.782712893384) ? INFINITY : ((x < -700.0) ? 0.0 : exp_small(x));
    if (is_nan) e = nan("");
    if (sizeof(OutT) == 4) out[i] = static_cast<OutT>(__fdiv_rn(1.f, static_cast<float>(1.0 + e)));
    else out[i] = static_cast<OutT>(__ddiv_rn(1.0, 1.0 + e));
  }
}

// Work buffers are kept per device between calls (grow-only): cudaMalloc / cudaFree would cost more
// than the kernels.  Calls are serialised by the mutex (the C ABI is one-thread-per-handle anyway).
struct FitWork {
  SelState* st = nullptr;
  unsigned int* ghist = nullptr;
  unsigned int* skeys = nullptr;
  unsigned int* wkeys = nullptr;
  unsigned int* wkeysd = nullptr; // one-pass path: keys of the values collected around c0 for the MAD
  size_t wkeysd_bytes = 0;
  unsigned int* surv = nullptr;   // [columns][kSurvCap] keys of the target bins
  unsigned int* done = nullptr;   // per-column block tickets of hist_kernel / sample_kernel (self-resetting)
  size_t ghist_bytes = 0, skeys_bytes = 0, wkeys_bytes = 0, surv_bytes = 0;
  int sm_count = 148;
};
std::mutex g_fit_mu;
FitWork g_fit_work[64];

int ensure_buf(unsigned int** p, size_t* have, size_t need) {
  if (need <= *have) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  DEWI_CUDA(cudaMalloc(p, need));
  *have = need;
  return 0;
}

template <int SRC>
void select3(const void* src, long long n_host, const unsigned int* n_dev, long long ld, int f, SelState* st,
             unsigned int* ghist, unsigned int* done, cudaStream_t stream) {
  const int threads = 512;
  const int per_thread = 8;  // (64 keys per thread on the small key buffers measured slower: 43 vs 30 us per pass)
  int bx = static_cast<int>(std::min<int64_t>(ceil_div(n_host, threads * per_thread), std::max(1, current_sm_count() * 4 / f)));
  dim3 grid(std::max(bx, 1), f);
  for (int pass = 0; pass < 3; ++pass)
    hist_kernel<SRC><<<grid, threads, 0, stream>>>(src, n_host, n_dev, ld, pass, st, ghist, done, OP_NONE);
}

// Exact radix selection over the whole column: 3 full passes per statistic.
void fit_full(const float* cols, long long n, int f, long long ld, FitWork& w, cudaStream_t stream) {
  glue_kernel<<<1, kMaxCols, 0, stream>>>(ARM_FULL, f, n, 0u, w.st);
  select3<SRC_RAW>(cols, n, nullptr, ld, f, w.st, w.ghist, w.done, stream);
  glue_kernel<<<1, kMaxCols, 0, stream>>>(FINAL_MED, f, n, 0u, w.st);
  glue_kernel<<<1, kMaxCols, 0, stream>>>(ARM_FULL, f, n, 0u, w.st);
  select3<SRC_DEV>(cols, n, nullptr, ld, f, w.st, w.ghist, w.done, stream);
  glue_kernel<<<1, kMaxCols, 0, stream>>>(FINAL_MAD, f, n, 0u, w.st);
}

// One full pass per statistic: window bounds from a sample, keys inside the window collected, the
// middle ranks selected among them.  st->miss[c] reports a window that did not hold the ranks.
// DEWI_FIT_TIMING=1: CUDA-event time of every stage of a windowed fit on stderr (warm caches, in-stream -- what ncu's
// serialised cold-cache replays cannot show for the small kernels).
struct StageTimer {
  bool on = false;
  cudaEvent_t ev[32] = {};
  const char* name[32] = {};
  int n = 0;
  void mark(cudaStream_t stream, const char* what) {
    if (!on || n >= 32) return;
    if (!ev[n]) cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], stream);
    name[n++] = what;
  }
  void report() {
    if (!on || n < 2) return;
    cudaEventSynchronize(ev[n - 1]);
    float total = 0.f;
    cudaEventElapsedTime(&total, ev[0], ev[n - 1]);
    fprintf(stderr, "[dewi_fit_stats] %.1f us on the device:", total * 1e3f);
    for (int i = 1; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, " %s %.1f", name[i], ms * 1e3f);
    }
    fprintf(stderr, "\n");
    n = 0;
  }
};
StageTimer g_fit_timer;

// Four launches (the state must be armed for a sample selection: glue ARM_SAMPLE, or the previous statistic's
// survivor pass):
//   sample_kernel    gathers the sample keys, histograms their top 11 bits, picks the bins of ranks s/2 -+ margin
//   hist_kernel      pass 1 over the sample keys (next 11 bits)  ->  window bounds [lo, hi]   (OP_WINDOW_BOUNDS)
//   window_kernel    THE pass over the column: counts keys below the window, collects the keys inside it, histograms
//                    their window-relative top digit; the block finishing a column picks the two target bins
//   survivor_kernel  gathers the keys of those bins, sorts them in shared memory  ->  med / mad
template <int SRC>
void fit_windowed_stat(const float* cols, long long n, int f, long long ld, unsigned int cap, FitWork& w, int final_op,
                       cudaStream_t stream) {
  sample_kernel<SRC><<<dim3(kSampleBlocks, f), 256, 0, stream>>>(cols, n, ld, w.st, w.skeys, w.ghist, w.done,
                                                                   (n - kSampleRunLen) / (kSampleRuns - 1),
                                                                   static_cast<int>((n - kSampleRunLen) % (kSampleRuns - 1)));
  g_fit_timer.mark(stream, "sample");
  {
    const int threads = 512, per_thread = 8;
    const int bx = static_cast<int>(std::min<int64_t>(ceil_div(kSample, threads * per_thread), std::max(1, w.sm_count * 4 / f)));
    hist_kernel<SRC_KEYS><<<dim3(std::max(bx, 1), f), threads, 0, stream>>>(w.skeys, kSample, nullptr, kSample, kSamplePass1, w.st,
                                                                             w.ghist, w.done, OP_WINDOW_BOUNDS);
  }
  g_fit_timer.mark(stream, "hist");
  if (env_int("DEWI_FIT_FUSED", kFitFusedDefault)) {
    window_kernel<SRC, 1><<<w.sm_count * kWinBlocksPerSm, kWinThreads, 0, stream>>>(cols, n, ld, f, w.st, w.wkeys, cap, w.ghist);
    g_fit_timer.mark(stream, "window+hist");
  } else {
    window_kernel<SRC, 0><<<w.sm_count * kWinBlocksPerSm, kWinThreads, 0, stream>>>(cols, n, ld, f, w.st, w.wkeys, cap, w.ghist);
    g_fit_timer.mark(stream, "window");
    window_hist_kernel<0><<<dim3(std::max(1, w.sm_count * 4 / f), f), kWhistThreads, 0, stream>>>(w.wkeys, cap, n, w.st, w.ghist);
    g_fit_timer.mark(stream, "whist");
  }
  survivor_kernel<0><<<dim3(std::max(1, w.sm_count * 4 / f), f), kSurvThreads, 0, stream>>>(w.wkeys, cap, n, w.st, w.surv, final_op, 0u);
  g_fit_timer.mark(stream, "survivor");
}

// Both statistics from ONE pass over the columns (the state must be armed for a sample selection):
//   sample_kernel           sample keys + first histogram pass
//   hist_kernel             coarse pass  ->  median window [lo, hi], its centre c0 and half-width w
//   hist_kernel x 2         the same two passes over the sample's deviations |s - c0|  ->  deviation window [A, B]
//   window2_kernel          THE pass: median-window keys and deviation-window keys collected together
//   window_hist + survivor  exact median (as in the two-pass path), then the validity window [L, U] of exact deviations
//   window_hist + survivor  (DEV) exact MAD among the collected values, re-keyed as |y - med|
void fit_onepass(const float* cols, long long n, int f, long long ld, unsigned int cap0, unsigned int capd, FitWork& w,
                 cudaStream_t stream) {
  const dim3 small(std::max(1, w.sm_count * 4 / f), f);
  const int threads = 512, per_thread = 8;
  const int bx = static_cast<int>(std::min<int64_t>(ceil_div(kSample, threads * per_thread), std::max(1, w.sm_count * 4 / f)));
  const dim3 hgrid(std::max(bx, 1), f);
  sample_kernel<SRC_RAW><<<dim3(kSampleBlocks, f), 256, 0, stream>>>(cols, n, ld, w.st, w.skeys, w.ghist, w.done,
                                                                       (n - kSampleRunLen) / (kSampleRuns - 1),
                                                                       static_cast<int>((n - kSampleRunLen) % (kSampleRuns - 1)));
  g_fit_timer.mark(stream, "sample");
  hist_kernel<SRC_KEYS><<<hgrid, threads, 0, stream>>>(w.skeys, kSample, nullptr, kSample, kSamplePass1, w.st, w.ghist, w.done,
                                                       OP_WINDOW_BOUNDS_1P);
  g_fit_timer.mark(stream, "hist");
  hist_kernel<SRC_SKEY_DEV><<<hgrid, threads, 0, stream>>>(w.skeys, kSample, nullptr, kSample, 0, w.st, w.ghist, w.done, OP_NONE);
  g_fit_timer.mark(stream, "dev-hist0");
  hist_kernel<SRC_SKEY_DEV><<<hgrid, threads, 0, stream>>>(w.skeys, kSample, nullptr, kSample, kSamplePass1, w.st, w.ghist, w.done,
                                                           OP_DEV_BOUNDS);
  g_fit_timer.mark(stream, "dev-hist1");
  window2_kernel<0><<<w.sm_count * kWin2BlocksPerSm, kWinThreads, 0, stream>>>(cols, n, ld, f, w.st, w.wkeys, cap0, w.wkeysd, capd);
  g_fit_timer.mark(stream, "window2");
  window_hist_kernel<0><<<small, kWhistThreads, 0, stream>>>(w.wkeys, cap0, n, w.st, w.ghist);
  g_fit_timer.mark(stream, "whist");
  survivor_kernel<0><<<small, kSurvThreads, 0, stream>>>(w.wkeys, cap0, n, w.st, w.surv, FINAL_MED_1P, capd);
  g_fit_timer.mark(stream, "survivor");
  window_hist_kernel<1><<<small, kWhistThreads, 0, stream>>>(w.wkeysd, capd, n, w.st, w.ghist);
  g_fit_timer.mark(stream, "dev-whist");
  survivor_kernel<1><<<small, kSurvThreads, 0, stream>>>(w.wkeysd, capd, n, w.st, w.surv, FINAL_MAD, capd);
  g_fit_timer.mark(stream, "dev-survivor");
}

}  // namespace
}  // namespace dewi

using namespace dewi;

extern "C" int dewi_fit_stats(const float* cols, int64_t n, int f, int64_t ld, double* med_host, double* mad_host,
                              int device, void* stream_) {
  if (!cols || !med_host || !mad_host) return fail("null argument");
  if (n <= 0) return fail("fit_stats needs at least one row");
  if (f <= 0 || f > kMaxCols) return fail("fit_stats supports 1..32 columns");
  if (ld < n) return fail("ld must be >= n");
  int sms = 148;
  DEWI_TRY(dewi_device_check(device, &sms, nullptr, nullptr));
  DEWI_CUDA(cudaSetDevice(device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool windowed = n >= kWindowMinRows;
  // expected window population is 2 * margin / sample = 0.5 % of n (+ the outward rounding of its bounds); allow 1.6 %
  const unsigned int cap = windowed ? static_cast<unsigned int>(std::min<int64_t>(n / 64 + 65536, 1ll << 30)) : 0u;
  // one-pass path: the deviation window holds its own 0.5 % plus 4 half-widths of the median window on either side
  // (~2.5 % on smooth densities); allow 6.25 %
  const bool onepass = windowed && env_int("DEWI_FIT_ONEPASS", kFitOnePassDefault) != 0;
  const unsigned int capd = onepass ? static_cast<unsigned int>(std::min<int64_t>(n / 16 + 65536, 1ll << 30)) : 0u;
  if (device >= 64) return fail("device ordinal out of range");
  std::lock_guard<std::mutex> lock(g_fit_mu);
  FitWork& w = g_fit_work[device];
  w.sm_count = sms;
  if (!w.st) DEWI_CUDA(cudaMalloc(&w.st, sizeof(SelState)));
  if (!w.done) {
    DEWI_CUDA(cudaMalloc(&w.done, kMaxCols * sizeof(unsigned int)));
    DEWI_CUDA(cudaMemsetAsync(w.done, 0, kMaxCols * sizeof(unsigned int), stream));
  }
  DEWI_TRY(ensure_buf(&w.ghist, &w.ghist_bytes, static_cast<size_t>(kMaxCols) * 2 * kBins * 4));
  if (windowed) {
    DEWI_TRY(ensure_buf(&w.skeys, &w.skeys_bytes, static_cast<size_t>(f) * kSample * 4));
    DEWI_TRY(ensure_buf(&w.wkeys, &w.wkeys_bytes, static_cast<size_t>(f) * cap * 4));
    DEWI_TRY(ensure_buf(&w.surv, &w.surv_bytes, static_cast<size_t>(kMaxCols) * kSurvCap * 4));
    if (onepass) DEWI_TRY(ensure_buf(&w.wkeysd, &w.wkeysd_bytes, static_cast<size_t>(f) * capd * 4));
  }
  DEWI_CUDA(cudaMemsetAsync(w.st, 0, sizeof(SelState), stream));
  DEWI_CUDA(cudaMemsetAsync(w.ghist, 0, static_cast<size_t>(f) * 2 * kBins * 4, stream));
  SelState res;
  bool need_full = !windowed;
  if (windowed) {
    g_fit_timer.on = env_set("DEWI_FIT_TIMING");
    g_fit_timer.mark(stream, "start");
    glue_kernel<<<1, kMaxCols, 0, stream>>>(ARM_SAMPLE, f, n, cap, w.st);
    g_fit_timer.mark(stream, "arm");
    if (onepass) {
      fit_onepass(cols, n, f, ld, cap, capd, w, stream);
    } else {
      fit_windowed_stat<SRC_RAW>(cols, n, f, ld, cap, w, FINAL_MED, stream);
      fit_windowed_stat<SRC_DEV>(cols, n, f, ld, cap, w, FINAL_MAD, stream);
    }
    DEWI_CUDA(cudaGetLastError());
    DEWI_CUDA(cudaMemcpyAsync(&res, w.st, sizeof(res), cudaMemcpyDeviceToHost, stream));
    DEWI_CUDA(cudaStreamSynchronize(stream));
    g_fit_timer.report();
    for (int c = 0; c < f; ++c) need_full = need_full || res.miss[c] != 0;  // e.g. a huge tie group at the median
    if (need_full) DEWI_CUDA(cudaMemsetAsync(w.ghist, 0, static_cast<size_t>(f) * 2 * kBins * 4, stream));
  }
  if (need_full) {
    fit_full(cols, n, f, ld, w, stream);
    DEWI_CUDA(cudaGetLastError());
    DEWI_CUDA(cudaMemcpyAsync(&res, w.st, sizeof(res), cudaMemcpyDeviceToHost, stream));
    DEWI_CUDA(cudaStreamSynchronize(stream));
  }
  for (int c = 0; c < f; ++c) {
    med_host[c] = static_cast<double>(res.med[c]);
    const double m = static_cast<double>(res.mad[c]);
    mad_host[c] = (m == 0.0) ? 1e-8 : m;  // scorer.py:24: `... or 1e-8`
  }
  return 0;
}

extern "C" int dewi_score(const void* cols, int in_f64, int64_t n, int64_t ld, const double* med7, const double* mad7,
                          const double* w6, int conditional, void* out, int out_f64, int device, void* stream_) {
  if (!cols || !med7 || !mad7 || !w6 || !out) return fail("null argument");
  if (n <= 0) return fail("score needs at least one row");
  if (ld < n) return fail("ld must be >= n");
  DEWI_TRY(dewi_device_check(device, nullptr, nullptr, nullptr));
  DEWI_CUDA(cudaSetDevice(device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ScoreParams p;
  const double a_t = w6[0], a_i = w6[1], a_m = w6[2], a_r = w6[3], a_n = w6[4];
  // scorer.py:53-54 (Ht, Hi are means of two z-scores), :67-73 standard, :80-87 conditional (Ht - I, Hi - I)
  const double wk[7] = {0.5 * a_t, 0.5 * a_t, 0.5 * a_i, 0.5 * a_i, conditional ? -(a_t + a_i) : -a_m, -a_r, -a_n};
  for (int c = 0; c < 7; ++c) {
    p.med[c] = med7[c];
    p.k[c] = wk[c] / (1.4826 * mad7[c]);
  }
  p.delta = w6[5];
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n, threads), static_cast<int64_t>(current_sm_count()) * 16));
  const float* c32 = static_cast<const float*>(cols);
  const double* c64 = static_cast<const double*>(cols);
  if (in_f64 && out_f64)
    score_kernel<double, double><<<blocks, threads, 0, stream>>>(c64, n, ld, p, static_cast<double*>(out));
  else if (in_f64)
    score_kernel<double, float><<<blocks, threads, 0, stream>>>(c64, n, ld, p, static_cast<float*>(out));
  else if (out_f64)
    score_kernel<float, double><<<blocks, threads, 0, stream>>>(c32, n, ld, p, static_cast<double*>(out));
  else
    score_kernel<float, float><<<blocks, threads, 0, stream>>>(c32, n, ld, p, static_cast<float*>(out));
  DEWI_CUDA(cudaGetLastError());
  return 0;
}
