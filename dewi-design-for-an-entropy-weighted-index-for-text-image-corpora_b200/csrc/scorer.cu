// K3 / K4 -- robust statistics and the DEWI score.
//
// K3 fit_stats: exact median and MAD of each signal column, replacing RobustStats.fit
//   (reference src/dewi/scorer.py:18-26: np.median of a float32 column, then np.median of
//   |v - med| in float32, zero MAD -> 1e-8).  Exact order statistics by MSB-first radix selection
//   over order-preserving uint32 keys: three histogram passes (11 + 11 + 10 bits) per statistic,
//   block-private shared-memory histograms flushed with one atomic per non-empty bin.  All columns
//   and both middle ranks (even n) are selected in the same passes.
// K4 score: RobustStats.z + DewiScorer._components/score/score_conditional (scorer.py:28-31,49-89)
//   in float64 with one rounding per operation (no FMA contraction), so results match the reference's
//   Python-float arithmetic; HBM-bound: 7 fp32 loads and one store per row.
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.h"

namespace dewi {
namespace {

constexpr int kBins = 2048;
constexpr int kMaxCols = 32;

struct SelState {
  // per column, per slot (two middle ranks)
  unsigned int prefix[kMaxCols][2];
  unsigned long long rank[kMaxCols][2];
  float value[kMaxCols][2];
  float med[kMaxCols];
  float mad[kMaxCols];
  int same[kMaxCols];  // both slots still share one prefix -> one histogram serves both
};

__device__ __forceinline__ unsigned int orderable(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ __forceinline__ void pass_bits(int pass, int& shift, int& nbins, unsigned int& himask) {
  // pass 0: bits 31..21, pass 1: bits 20..10, pass 2: bits 9..0
  if (pass == 0) { shift = 21; nbins = 2048; himask = 0u; }
  else if (pass == 1) { shift = 10; nbins = 2048; himask = 0xFFE00000u; }
  else { shift = 0; nbins = 1024; himask = 0xFFFFFC00u; }
}

template <bool MAD>
__global__ void __launch_bounds__(512)
hist_kernel(const float* __restrict__ cols, long long n, long long ld, int pass, const SelState* __restrict__ st,
            unsigned int* __restrict__ ghist) {
  __shared__ unsigned int sh[2][kBins];
  const int c = blockIdx.y;
  int shift, nbins;
  unsigned int himask;
  pass_bits(pass, shift, nbins, himask);
  for (int i = threadIdx.x; i < 2 * kBins; i += blockDim.x) (&sh[0][0])[i] = 0u;
  __syncthreads();
  const unsigned int p0 = st->prefix[c][0] & himask, p1 = st->prefix[c][1] & himask;
  const bool same = st->same[c] != 0;
  const float med = MAD ? st->med[c] : 0.f;
  const float* col = cols + static_cast<size_t>(c) * ld;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const unsigned int binmask = static_cast<unsigned int>(nbins - 1);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = __ldg(col + i);
    if (MAD) v = fabsf(__fsub_rn(v, med));
    const unsigned int key = orderable(v);
    const unsigned int hi = key & himask;
    const unsigned int bin = (key >> shift) & binmask;
    if (hi == p0) atomicAdd(&sh[0][bin], 1u);
    if (!same && hi == p1) atomicAdd(&sh[1][bin], 1u);
  }
  __syncthreads();
  unsigned int* g = ghist + static_cast<size_t>(c) * 2 * kBins;
  for (int i = threadIdx.x; i < 2 * kBins; i += blockDim.x) {
    const unsigned int v = (&sh[0][0])[i];
    if (v) atomicAdd(&g[i], v);
  }
}

// One block per column: locate the bin holding each slot's rank, extend the prefix, clear the histogram.
template <bool MAD>
__global__ void __launch_bounds__(1024)
pick_kernel(int pass, long long n, SelState* st, unsigned int* ghist) {
  __shared__ unsigned long long cum[kBins];
  __shared__ unsigned long long wsum[32];
  const int c = blockIdx.x;
  int shift, nbins;
  unsigned int himask;
  pass_bits(pass, shift, nbins, himask);
  const int was_same = st->same[c];
  const int nslots = was_same ? 1 : 2;
  // ranks are read by every thread up front: the thread that finds the bin rewrites them below
  const unsigned long long rank_in[2] = {st->rank[c][0], st->rank[c][1]};
  __syncthreads();
  for (int s = 0; s < nslots; ++s) {
    unsigned int* g = ghist + (static_cast<size_t>(c) * 2 + s) * kBins;
    // inclusive scan of nbins (<= 2048) counts with 1024 threads, two bins per thread
    const int t = threadIdx.x;
    const unsigned long long a = (2 * t < nbins) ? g[2 * t] : 0ull;
    const unsigned long long b = (2 * t + 1 < nbins) ? g[2 * t + 1] : 0ull;
    unsigned long long x = a + b;
    const int lane = t & 31, w = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      unsigned long long y = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long z = __shfl_up_sync(0xffffffffu, y, o);
        if (lane >= o) y += z;
      }
      wsum[lane] = y;
    }
    __syncthreads();
    const unsigned long long base = (w > 0) ? wsum[w - 1] : 0ull;
    cum[2 * t] = base + x - b;   // inclusive up to bin 2t
    cum[2 * t + 1] = base + x;   // inclusive up to bin 2t+1
    __syncthreads();
    // slots that share this histogram
    const int s_end = was_same ? 2 : s + 1;
    for (int ss = s; ss < s_end; ++ss) {
      const unsigned long long r = rank_in[ss];
      for (int bin = threadIdx.x; bin < nbins; bin += blockDim.x) {
        const unsigned long long lo = (bin > 0) ? cum[bin - 1] : 0ull;
        if (r >= lo && r < cum[bin]) {
          st->prefix[c][ss] = (st->prefix[c][ss] & himask) | (static_cast<unsigned int>(bin) << shift);
          st->rank[c][ss] = r - lo;
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) g[i] = 0u;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (was_same && st->prefix[c][0] != st->prefix[c][1]) st->same[c] = 0;
    if (pass == 2) {
      const float v0 = from_orderable(st->prefix[c][0]);
      const float v1 = from_orderable(st->prefix[c][1]);
      // np.median: mean of the middle element(s) in float32 -> (v0 + v1) / 2, or v0 when n is odd
      const float m = (n & 1) ? v0 : __fmul_rn(__fadd_rn(v0, v1), 0.5f);
      if (MAD) st->mad[c] = m; else st->med[c] = m;
      // re-arm for the next statistic
      const unsigned long long r0 = (n & 1) ? static_cast<unsigned long long>((n - 1) / 2) : static_cast<unsigned long long>(n / 2 - 1);
      const unsigned long long r1 = (n & 1) ? r0 : static_cast<unsigned long long>(n / 2);
      st->prefix[c][0] = st->prefix[c][1] = 0u;
      st->rank[c][0] = r0;
      st->rank[c][1] = r1;
      st->same[c] = 1;
    }
  }
}

struct ScoreParams {
  double med[7];
  double den[7];  // 1.4826 * mad
  double w[6];    // alpha_t, alpha_i, alpha_m, alpha_r, alpha_n, delta
  int conditional;
};

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
score_kernel(const InT* __restrict__ cols, long long n, long long ld, const ScoreParams p, OutT* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double z[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      const double v = static_cast<double>(__ldg(cols + static_cast<size_t>(c) * ld + i));
      z[c] = __ddiv_rn(__dsub_rn(v, p.med[c]), p.den[c]);  // scorer.py:31
    }
    const double Ht = __dmul_rn(0.5, __dadd_rn(z[0], z[1]));  // scorer.py:53
    const double Hi = __dmul_rn(0.5, __dadd_rn(z[2], z[3]));  // scorer.py:54
    const double I = z[4], R = z[5], N = z[6];
    double U;
    if (!p.conditional) {  // scorer.py:67-73, evaluated left to right
      U = __dadd_rn(__dmul_rn(p.w[0], Ht), __dmul_rn(p.w[1], Hi));
      U = __dsub_rn(U, __dmul_rn(p.w[2], I));
      U = __dsub_rn(U, __dmul_rn(p.w[3], R));
      U = __dsub_rn(U, __dmul_rn(p.w[4], N));
    } else {  // scorer.py:80-87
      const double HtI = __dsub_rn(Ht, I), HiI = __dsub_rn(Hi, I);
      U = __dadd_rn(__dmul_rn(p.w[0], HtI), __dmul_rn(p.w[1], HiI));
      U = __dsub_rn(U, __dmul_rn(p.w[3], R));
      U = __dsub_rn(U, __dmul_rn(p.w[4], N));
    }
    U = fmin(fmax(U, -p.w[5]), p.w[5]);                              // scorer.py:74
    const double s = __ddiv_rn(1.0, __dadd_rn(1.0, exp(-U)));        // scorer.py:62
    out[i] = static_cast<OutT>(s);
  }
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
  DEWI_TRY(dewi_device_check(device, nullptr, nullptr, nullptr));
  DEWI_CUDA(cudaSetDevice(device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SelState init;
  memset(&init, 0, sizeof(init));
  const unsigned long long r0 = (n & 1) ? (n - 1) / 2 : n / 2 - 1;
  const unsigned long long r1 = (n & 1) ? r0 : n / 2;
  for (int c = 0; c < f; ++c) {
    init.rank[c][0] = r0;
    init.rank[c][1] = r1;
    init.same[c] = 1;
  }
  SelState* st = nullptr;
  unsigned int* ghist = nullptr;
  DEWI_CUDA(cudaMalloc(&st, sizeof(SelState)));
  DEWI_CUDA(cudaMalloc(&ghist, static_cast<size_t>(f) * 2 * kBins * 4));
  int rc = 0;
  do {
    if (cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaMemsetAsync(ghist, 0, static_cast<size_t>(f) * 2 * kBins * 4, stream) != cudaSuccess) {
      rc = fail("fit_stats: state upload failed");
      break;
    }
    const int threads = 512;
    int bx = static_cast<int>(std::min<int64_t>(ceil_div(n, threads * 8), std::max(1, 148 * 4 / f)));
    bx = std::max(bx, 1);
    dim3 grid(bx, f);
    for (int phase = 0; phase < 2 && rc == 0; ++phase) {
      for (int pass = 0; pass < 3; ++pass) {
        if (phase == 0) {
          hist_kernel<false><<<grid, threads, 0, stream>>>(cols, n, ld, pass, st, ghist);
          pick_kernel<false><<<f, 1024, 0, stream>>>(pass, n, st, ghist);
        } else {
          hist_kernel<true><<<grid, threads, 0, stream>>>(cols, n, ld, pass, st, ghist);
          pick_kernel<true><<<f, 1024, 0, stream>>>(pass, n, st, ghist);
        }
      }
      if (cudaGetLastError() != cudaSuccess) rc = fail("fit_stats: kernel launch failed");
    }
    if (rc) break;
    SelState res;
    if (cudaMemcpyAsync(&res, st, sizeof(res), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
      rc = fail(std::string("fit_stats: ") + cudaGetErrorString(cudaGetLastError()));
      break;
    }
    for (int c = 0; c < f; ++c) {
      med_host[c] = static_cast<double>(res.med[c]);
      const double m = static_cast<double>(res.mad[c]);
      mad_host[c] = (m == 0.0) ? 1e-8 : m;  // scorer.py:24: `... or 1e-8`
    }
  } while (0);
  cudaFree(st);
  cudaFree(ghist);
  return rc;
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
  for (int c = 0; c < 7; ++c) {
    p.med[c] = med7[c];
    p.den[c] = 1.4826 * mad7[c];
  }
  for (int i = 0; i < 6; ++i) p.w[i] = w6[i];
  p.conditional = conditional ? 1 : 0;
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n, threads), 148 * 16));
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
