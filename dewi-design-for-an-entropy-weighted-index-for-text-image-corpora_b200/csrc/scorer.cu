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
// narrower window means fewer keys for the full pass to collect.
constexpr int kSampleMargin = 2048;
constexpr int kWindowDigitBits = 11;             // window-relative digit of the in-window selection (2048 bins)
constexpr int kSurvCap = 8192;                   // keys of the target bins sorted in shared memory (typically a few hundred)

struct SelState {
  // results first: the host reads back exactly this head (FitResult) into pinned memory
  float med[kMaxCols];
  float mad[kMaxCols];
  int miss[kMaxCols];                       // window overflowed or does not hold the target ranks
  // radix selection, per column, two rank slots
  unsigned int prefix[kMaxCols][2];
  unsigned long long rank[kMaxCols][2];
  int same[kMaxCols];  // both slots still share one prefix -> one histogram serves both
  // windowed path
  unsigned int lo[kMaxCols], hi[kMaxCols];  // inclusive key window
  unsigned long long below[kMaxCols];       // keys < lo
  unsigned int wcnt[kMaxCols];              // keys appended to the window buffer
  // in-window selection: digit = (key - lo) >> wshift is the top <= 11 bits of the window's key span
  int wshift[kMaxCols];
  unsigned int wbin[kMaxCols][2];           // digits of the two middle ranks
  unsigned long long wrank[kMaxCols][2];    // their ranks inside those bins
  unsigned int tiles_done[kMaxCols];        // window_kernel: tiles of the column finished (self-resetting)
  unsigned int surv_n[kMaxCols];            // survivor_kernel: keys gathered from the target bins
  unsigned int surv_n0[kMaxCols];           //   ... of which in bin 0, when the two bins differ
  unsigned int surv_done[kMaxCols];         // survivor_kernel: block tickets (self-resetting)
  float c0[kMaxCols], r0[kMaxCols];         // the window as the full pass tests it: |fl(x - c0)| <= r0 (a superset of [lo, hi])
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
enum { SRC_RAW = 0, SRC_DEV = 1, SRC_KEYS = 2 };

template <int SRC>
__device__ __forceinline__ unsigned int load_key(const void* col, long long i, float med) {
  if (SRC == SRC_KEYS) return __ldg(static_cast<const unsigned int*>(col) + i);
  float v = __ldg(static_cast<const float*>(col) + i);
  if (SRC == SRC_DEV) v = fabsf(__fsub_rn(v, med));
  return orderable(v);
}

__device__ void pick_column(int pass, int c, SelState* st, unsigned int* ghist, unsigned long long* cum,
                            unsigned long long* wsum);

// What the last block of a selection pass does after it has picked its bins.
enum { OP_NONE = 0, OP_WINDOW_BOUNDS = 1 };

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
  // The full pass tests the window in centre / radius form, |fl(x - c0)| <= r0 (one subtraction shared by the
  // "below" count and the membership test): r0 is the half-width inflated by 2^-10 and by a slack far above the
  // rounding of the subtraction, so the tested window is a superset of [lo, hi].
  const float lo_f = from_orderable(lo), hi_f = from_orderable(hi);
  const float c0 = 0.5f * lo_f + 0.5f * hi_f;
  float r0 = fmaxf(hi_f - c0, c0 - lo_f);
  r0 = r0 * 1.0009765625f + (fabsf(c0) + r0) * 3.814697265625e-06f + 1e-30f;
  st->c0[c] = c0;
  st->r0[c] = r0;
  if (!(r0 < INFINITY) || !(c0 == c0)) st->miss[c] = 1;   // infinite / NaN bounds: not a usable window
}

// State of a fresh sample selection: ranks s/2 -+ margin of the kSample sample keys.
__device__ __forceinline__ void arm_sample(SelState* st, int c) {
  st->prefix[c][0] = st->prefix[c][1] = 0u;
  st->rank[c][0] = kSample / 2 - kSampleMargin;
  st->rank[c][1] = kSample / 2 + kSampleMargin;
  st->same[c] = 1;
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
  const float med = (SRC == SRC_DEV) ? st->med[c] : 0.f;
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
    if (op == OP_WINDOW_BOUNDS) arm_window(st, c);
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
enum { ARM_FULL, ARM_SAMPLE, ARM_WINDOW, ARM_INSIDE, FINAL_MED, FINAL_MAD };

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
// The pass is an HBM stream that the instruction issue rate must not hold back: the first version -- `x >= lo`,
// `below += !p`, `p && x <= hi`, a hit mask, and every hit re-read, re-keyed and appended on the spot -- ran 292M warp
// instructions for 700M elements (ncu: issue slots 62 % busy, DRAM at 80 % of its peak, 425 us per pass).  Now
// (1) the window is tested in centre / radius form on t = fl(x - c0): `t < -r0` feeds the "below" count and
// `|t| <= r0` the membership, four instructions per element; the tests are monotone in x, so they define a proper
// window (a superset of the [lo, hi] the sample bracketed), and count and collection partition the column
// consistently because both use the same rounded t.  (2) Hits (~0.5 % of the elements) set a bit of a per-thread mask;
// lanes with a non-zero mask walk its set bits, re-read the element (an L1 hit: it was loaded a few instructions ago)
// and append its order-preserving key to a per-block shared buffer with one shared atomic each; the buffer goes to the
// global window with one global atomic when it is half full.  Variants that were measured and lost: queueing the
// thread's element index and re-testing at flush time (the re-reads miss L2 -- ncu: 4.99 GB read from DRAM for a
// 2.8 GB pass, 731 us); appending the hits from the registers (x[16] stays live across the hit path, 32 registers
// spill: 571-614 us); fusing the window histogram into the pass (483-498 us).  Float order and key order differ
// only in (-0, +0) and NaNs, and the selection inside the window uses key order (a refinement of float order), so
// the selected VALUE is exact.  The [column][row] space is flattened into 4096-element tiles and cut into one
// contiguous tile range per block, so a grid of exactly (SMs x resident blocks) is one balanced wave for any column count.
constexpr int kWinBuf = 3072;
constexpr int kWinThreads = 256;
constexpr int kWinPerThread = 16;
constexpr int kWinTile = kWinThreads * kWinPerThread;
constexpr int kWinBlocksPerSm = 8;
constexpr int kFlushCheckEvery = 8;  // tiles between two looks at the buffer fill (8 tiles add ~160 keys)

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

// Shared memory of the block that picks a column's target bins (window_pick): 2048 x u64 prefix sums + scratch.
struct WinShared {
  unsigned int shm[kWinBuf + kBins];   // [ (unused here) | histogram of the window-relative digit ], reused as the prefix sums
  unsigned long long wsum[32];
  unsigned int ticket;
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

template <int SRC>
__global__ void __launch_bounds__(kWinThreads, kWinBlocksPerSm)
window_kernel(const float* __restrict__ cols, long long n, long long ld, int f, SelState* st,
              unsigned int* __restrict__ wkeys, unsigned int window_cap) {
  __shared__ unsigned int buf[kWinBuf];   // keys waiting for the next flush
  __shared__ unsigned int buf_n, flush_base;
  __shared__ unsigned long long below_blk;
  const int lane = threadIdx.x & 31;
  // (tile counts fit 32 bits: 2^31 rows x 32 columns / 4096)
  const int tiles_per_col = static_cast<int>((n + kWinTile - 1) / kWinTile);
  const long long total_tiles = static_cast<long long>(tiles_per_col) * f;
  const int t_begin = static_cast<int>((total_tiles * blockIdx.x) / gridDim.x);
  const int t_end = static_cast<int>((total_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (threadIdx.x == 0) { buf_n = 0u; below_blk = 0ull; }
  __syncthreads();

  int c = -1;
  float med = 0.f, c0 = 0.f, r0 = 0.f;
  const float* col = nullptr;
  unsigned int* out = nullptr;
  bool aligned = false;
  unsigned int below = 0u;
  int since_check = 0;

  // the key of an element: RAW the value, DEV |v - med| in fp32 (scorer.py:24)
  auto keyed = [&](float y) { return (SRC == SRC_DEV) ? fabsf(__fsub_rn(y, med)) : y; };

  auto flush = [&](unsigned int min_fill) {  // block-uniform
    __syncthreads();
    const unsigned int cnt = min(buf_n, static_cast<unsigned int>(kWinBuf));
    if (cnt > min_fill) {
      if (threadIdx.x == 0) flush_base = atomicAdd(&st->wcnt[c], cnt);
      __syncthreads();
      for (unsigned int k = threadIdx.x; k < cnt; k += blockDim.x)
        if (flush_base + k < window_cap) out[flush_base + k] = buf[k];
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
  };
  auto push = [&](float x) {  // x lies in the window
    const unsigned int key = orderable(x);
    const unsigned int slot = atomicAdd(&buf_n, 1u);
    if (slot < kWinBuf) buf[slot] = key;
    else {  // shared buffer full (a huge tie group): straight to global; the caller will see wcnt > cap
      const unsigned int g = atomicAdd(&st->wcnt[c], 1u);
      if (g < window_cap) out[g] = key;
    }
  };

  int next_switch = t_begin;   // first tile of the next column
  for (int t = t_begin; t < t_end; ++t) {
    if (t >= next_switch) {
      if (c >= 0) commit_column();
      c = t / tiles_per_col;
      next_switch = (c + 1) * tiles_per_col;
      med = (SRC == SRC_DEV) ? st->med[c] : 0.f;
      c0 = st->c0[c];
      r0 = st->r0[c];
      col = cols + static_cast<size_t>(c) * ld;
      out = wkeys + static_cast<size_t>(c) * window_cap;
      aligned = (reinterpret_cast<uintptr_t>(col) & 15) == 0;
    }
    const long long base = static_cast<long long>(t - (next_switch - tiles_per_col)) * kWinTile;
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
      const float nr0 = -r0;
#pragma unroll
      for (int e = 0; e < kWinPerThread; ++e) {
        const float tt = __fsub_rn(keyed(x[e]), c0);
        below += (tt < nr0) ? 1u : 0u;
        mask |= (fabsf(tt) <= r0) ? (1u << e) : 0u;
      }
      while (mask) {  // rare per lane; the element is re-read (an L1 hit) instead of indexing registers dynamically
        const int e = __ffs(mask) - 1;
        mask &= mask - 1u;
        push(keyed(p[(e >> 2) * (kWinThreads * 4) + (e & 3)]));
      }
    } else {  // ragged last tile of a column, or a column that is not 16-byte aligned
      for (int e = 0; e < kWinPerThread; ++e) {
        const long long i = base + static_cast<long long>(e >> 2) * kWinThreads * 4 + threadIdx.x * 4 + (e & 3);
        if (i < n) {
          const float xx = keyed(__ldg(col + i));
          const float tt = __fsub_rn(xx, c0);
          below += (tt < -r0) ? 1u : 0u;
          if (fabsf(tt) <= r0) push(xx);
        }
      }
    }
    if (++since_check == kFlushCheckEvery) flush(kWinBuf / 2);
  }
  if (c >= 0) commit_column();
}

// The histogram of the window-relative digit over the collected keys (L2-resident), one launch for all columns; the
// last block of a column picks the target bins (window_pick).  (Fusing this histogram into the window pass was tried:
// 483-498 us per pass instead of 445 + 20 -- the extra live state cost the streaming loop more than the launch.)
constexpr int kWhistThreads = 512;

__global__ void __launch_bounds__(kWhistThreads)
window_hist_kernel(const unsigned int* __restrict__ wkeys, unsigned int window_cap, long long n, SelState* st,
                   unsigned int* __restrict__ ghist) {
  __shared__ __align__(8) WinShared ws;
  const int c = blockIdx.y;
  unsigned int* hist = ws.shm + kWinBuf;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const unsigned int lo = st->lo[c], span = st->hi[c] - lo;
  const int wshift = st->wshift[c];
  const unsigned int cnt = min(__ldcg(&st->wcnt[c]), window_cap);
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
    for (int u = 0; u < kU; ++u)
      if (base + u * blockDim.x + threadIdx.x < cnt) atomicAdd(&hist[window_digit(key[u], lo, span, wshift)], 1u);
  }
  __syncthreads();
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

__global__ void __launch_bounds__(kSurvThreads)
survivor_kernel(const unsigned int* __restrict__ wkeys, unsigned int window_cap, long long n, SelState* st,
                unsigned int* __restrict__ surv, int final_op) {
  __shared__ unsigned int keys[kSurvCap];
  __shared__ unsigned int ticket;
  const int c = blockIdx.y;
  if (st->miss[c]) return;   // (uniform over the column's blocks: set by an earlier kernel)
  const unsigned int lo = st->lo[c], span = st->hi[c] - lo;
  const int wshift = st->wshift[c];
  const unsigned int b0 = st->wbin[c][0], b1 = st->wbin[c][1];
  const unsigned int cnt = min(st->wcnt[c], window_cap);
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
struct FitResult {   // == the head of SelState
  float med[kMaxCols];
  float mad[kMaxCols];
  int miss[kMaxCols];
};

struct FitWork {
  SelState* st = nullptr;
  FitResult* res_host = nullptr;   // pinned: the read-back is one small asynchronous copy
  unsigned int* ghist = nullptr;
  unsigned int* skeys = nullptr;
  unsigned int* wkeys = nullptr;
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
  window_kernel<SRC><<<w.sm_count * kWinBlocksPerSm, kWinThreads, 0, stream>>>(cols, n, ld, f, w.st, w.wkeys, cap);
  g_fit_timer.mark(stream, "window");
  window_hist_kernel<<<dim3(std::max(1, w.sm_count * 4 / f), f), kWhistThreads, 0, stream>>>(w.wkeys, cap, n, w.st, w.ghist);
  g_fit_timer.mark(stream, "whist");
  survivor_kernel<<<dim3(std::max(1, w.sm_count * 4 / f), f), kSurvThreads, 0, stream>>>(w.wkeys, cap, n, w.st, w.surv, final_op);
  g_fit_timer.mark(stream, "survivor");
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
  struct DeviceGuard {   // the caller's current device is left as it was (the Python wrapper no longer switches it)
    int prev = -1, want;
    explicit DeviceGuard(int d) : want(d) {
      cudaGetDevice(&prev);
      if (prev != want) cudaSetDevice(want);
    }
    ~DeviceGuard() {
      if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
  } guard(device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool windowed = n >= kWindowMinRows;
  // expected window population is 2 * margin / sample = 0.8 % of n; allow twice that
  const unsigned int cap = windowed ? static_cast<unsigned int>(std::min<int64_t>(n / 64 + 65536, 1ll << 30)) : 0u;
  if (device >= 64) return fail("device ordinal out of range");
  std::lock_guard<std::mutex> lock(g_fit_mu);
  FitWork& w = g_fit_work[device];
  w.sm_count = sms;
  if (!w.st) DEWI_CUDA(cudaMalloc(&w.st, sizeof(SelState)));
  if (!w.res_host) DEWI_CUDA(cudaHostAlloc(&w.res_host, sizeof(FitResult), cudaHostAllocDefault));
  const bool fresh_hist = w.ghist_bytes == 0;
  if (!w.done) {
    DEWI_CUDA(cudaMalloc(&w.done, kMaxCols * sizeof(unsigned int)));
    DEWI_CUDA(cudaMemsetAsync(w.done, 0, kMaxCols * sizeof(unsigned int), stream));
  }
  DEWI_TRY(ensure_buf(&w.ghist, &w.ghist_bytes, static_cast<size_t>(kMaxCols) * 2 * kBins * 4));
  if (windowed) {
    DEWI_TRY(ensure_buf(&w.skeys, &w.skeys_bytes, static_cast<size_t>(f) * kSample * 4));
    DEWI_TRY(ensure_buf(&w.wkeys, &w.wkeys_bytes, static_cast<size_t>(f) * cap * 4));
    DEWI_TRY(ensure_buf(&w.surv, &w.surv_bytes, static_cast<size_t>(kMaxCols) * kSurvCap * 4));
  }
  DEWI_CUDA(cudaMemsetAsync(w.st, 0, sizeof(SelState), stream));
  // (the histograms are zeroed once: every pass leaves them clean behind its pick)
  if (fresh_hist) DEWI_CUDA(cudaMemsetAsync(w.ghist, 0, static_cast<size_t>(kMaxCols) * 2 * kBins * 4, stream));
  FitResult& res = *w.res_host;
  bool need_full = !windowed;
  if (windowed) {
    g_fit_timer.on = env_set("DEWI_FIT_TIMING");
    g_fit_timer.mark(stream, "start");
    glue_kernel<<<1, kMaxCols, 0, stream>>>(ARM_SAMPLE, f, n, cap, w.st);
    g_fit_timer.mark(stream, "arm");
    fit_windowed_stat<SRC_RAW>(cols, n, f, ld, cap, w, FINAL_MED, stream);
    fit_windowed_stat<SRC_DEV>(cols, n, f, ld, cap, w, FINAL_MAD, stream);
    DEWI_CUDA(cudaGetLastError());
    DEWI_CUDA(cudaMemcpyAsync(&res, w.st, sizeof(FitResult), cudaMemcpyDeviceToHost, stream));
    DEWI_CUDA(cudaStreamSynchronize(stream));
    g_fit_timer.report();
    for (int c = 0; c < f; ++c) need_full = need_full || res.miss[c] != 0;  // e.g. a huge tie group at the median
    if (need_full) DEWI_CUDA(cudaMemsetAsync(w.ghist, 0, static_cast<size_t>(f) * 2 * kBins * 4, stream));
  }
  if (need_full) {
    fit_full(cols, n, f, ld, w, stream);
    DEWI_CUDA(cudaGetLastError());
    DEWI_CUDA(cudaMemcpyAsync(&res, w.st, sizeof(FitResult), cudaMemcpyDeviceToHost, stream));
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
