// Operand preparation: row normalisation (ExactIndex.add, reference src/dewi/backends.py:403-405;
// query normalisation, :420-424) fused with the split of fp32 values into bf16 planes that the
// tensor-core sweep streams.  One warp per row, coalesced, fp32 arithmetic with IEEE division.
#include <algorithm>

#include "internal.h"

namespace dewi {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// hi = bf16(x); lo = bf16(x - hi): hi + lo carries 16 significand bits of x.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(__fsub_rn(x, __bfloat162float(hi)));
}
// The same split in fp16 (22 significand bits where lo stays normal); the bit patterns travel in the 16-bit plane
// storage.  hf / lf return the values the planes hold.
__device__ __forceinline__ void split_fp16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo, float& hf, float& lf) {
  const __half h = __float2half_rn(x);
  hf = __half2float(h);
  const __half l = __float2half_rn(__fsub_rn(x, hf));
  lf = __half2float(l);
  hi = __ushort_as_bfloat16(__half_as_ushort(h));
  lo = __ushort_as_bfloat16(__half_as_ushort(l));
}

// Rounding bookkeeping for the CERTIFIED single-plane sweep (api.cu): how far the bf16 planes are from the fp32 values.
//   row_stats[row] = { ||x||, ||x - hi||, ||x - hi - lo||, ||hi|| }   (queries: one record per query)
//   plane_max      = { max_r ||x_r - hi_r||, max_r ||hi_r|| }          (corpus: running maxima, float bits via atomicMax)
// Norms are evaluated in fp32 and inflated by 2^-10 (they bound an error, so they may only err upwards).
__device__ __forceinline__ float up(float v) { return v * 1.0009765625f + 1e-12f; }

__global__ void prep_rows_kernel(const float* __restrict__ src, long long n, long long n_out, int dim, int normalize,
                                 int guard_zero, int lane_order, float* __restrict__ dst_f32,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                 int* __restrict__ bad_flag, float* __restrict__ row_stats, unsigned int* __restrict__ plane_max,
                                 int fp16_planes) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  float warp_max_err = 0.f, warp_max_hi = 0.f;
  for (long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n_out; row += warps) {
    const size_t off = static_cast<size_t>(row) * dim;
    // bf16 planes may use the TMEM-lane order of the sweep (internal.h: query_lane)
    long long prow = row;
    if (lane_order == 1) prow = (row / kQueryBlock) * kQueryBlock + query_lane(static_cast<int>(row % kQueryBlock));
    if (lane_order == 2 && row < 64) prow = (row & 3) * 16 + (row >> 2);
    const size_t poff = static_cast<size_t>(prow) * dim;
    if (row >= n) {  // zero padding rows (query block padding)
      for (int d = lane; d < dim; d += 32) {
        if (dst_f32) dst_f32[off + d] = 0.f;
        if (hi) hi[poff + d] = __float2bfloat16_rn(0.f);
        if (lo) lo[poff + d] = __float2bfloat16_rn(0.f);
      }
      if (row_stats && lane < 4) row_stats[row * 4 + lane] = 0.f;
      continue;
    }
    const float* s = src + off;
    float scale_div = 1.f;
    bool do_div = false;
    if (normalize) {
      float ss = 0.f;
      for (int d = lane; d < dim; d += 32) {
        const float x = s[d];
        ss = fmaf(x, x, ss);
      }
      ss = warp_sum(ss);
      const float nrm = sqrtf(ss);
      if (nrm > 0.f) {
        scale_div = nrm;
        do_div = true;
      } else if (guard_zero && lane == 0 && bad_flag) {
        atomicExch(bad_flag, 1);  // zero-norm corpus row: the reference would store NaNs
      }
    }
    float sx = 0.f, s1 = 0.f, s2 = 0.f, sh = 0.f;
    for (int d = lane; d < dim; d += 32) {
      float x = s[d];
      if (do_div) x = __fdiv_rn(x, scale_div);
      if (dst_f32) dst_f32[off + d] = x;
      if (hi) {
        __nv_bfloat16 h, l;
        float hf, lf;
        if (fp16_planes) {
          split_fp16(x, h, l, hf, lf);
        } else {
          split_bf16(x, h, l);
          hf = __bfloat162float(h);
          lf = __bfloat162float(l);
        }
        hi[poff + d] = h;
        if (lo) lo[poff + d] = l;
        if (row_stats || plane_max) {
          const float r1 = x - hf, r2 = r1 - lf;   // (both differences are exact)
          sx = fmaf(x, x, sx);
          s1 = fmaf(r1, r1, s1);
          s2 = fmaf(r2, r2, s2);
          sh = fmaf(hf, hf, sh);
        }
      }
    }
    if (hi && (row_stats || plane_max)) {
      sx = up(sqrtf(warp_sum(sx)));
      s1 = up(sqrtf(warp_sum(s1)));
      s2 = up(sqrtf(warp_sum(s2)));
      sh = up(sqrtf(warp_sum(sh)));
      if (row_stats && lane == 0) {
        row_stats[row * 4 + 0] = sx;
        row_stats[row * 4 + 1] = s1;
        row_stats[row * 4 + 2] = s2;
        row_stats[row * 4 + 3] = sh;
      }
      warp_max_err = fmaxf(warp_max_err, s1);
      warp_max_hi = fmaxf(warp_max_hi, sh);
    }
  }
  if (plane_max && lane == 0) {  // one pair of atomics per warp (positive floats order like their bit patterns)
    atomicMax(plane_max + 0, __float_as_uint(warp_max_err));
    atomicMax(plane_max + 1, __float_as_uint(warp_max_hi));
  }
}

int launch(const float* src, int64_t n, int64_t n_out, int dim, int normalize, int guard_zero, int lane_order,
           float* dst_f32, __nv_bfloat16* hi, __nv_bfloat16* lo, int* bad_flag, cudaStream_t stream, float* row_stats = nullptr,
           unsigned int* plane_max = nullptr, int fp16_planes = 0) {
  if (n_out <= 0) return 0;
  const int threads = 256;
  const int64_t blocks = std::min<int64_t>(ceil_div(n_out * 32, threads), static_cast<int64_t>(current_sm_count()) * 16);
  prep_rows_kernel<<<static_cast<int>(blocks), threads, 0, stream>>>(src, n, n_out, dim, normalize, guard_zero,
                                                                    lane_order, dst_f32, hi, lo, bad_flag, row_stats, plane_max, fp16_planes);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

namespace {
__global__ void widen_bf16_kernel(const uint16_t* __restrict__ src, long long count, float* __restrict__ dst) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride)
    dst[i] = __uint_as_float(static_cast<uint32_t>(src[i]) << 16);
}
}  // namespace

int launch_widen_bf16(const __nv_bfloat16* src, int64_t count, float* dst, cudaStream_t stream) {
  if (count <= 0) return 0;
  const int threads = 256;
  const int64_t blocks = std::min<int64_t>(ceil_div(count, threads), static_cast<int64_t>(current_sm_count()) * 16);
  widen_bf16_kernel<<<static_cast<int>(blocks), threads, 0, stream>>>(reinterpret_cast<const uint16_t*>(src), count, dst);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

int launch_prep_corpus(const float* src, int64_t n, int dim, int normalize, float* dst_f32, __nv_bfloat16* hi,
                       __nv_bfloat16* lo, int* bad_flag, cudaStream_t stream, unsigned int* plane_max, int fp16_planes) {
  return launch(src, n, n, dim, normalize, 1, 0, dst_f32, hi, lo, bad_flag, stream, nullptr, plane_max, fp16_planes);
}

int launch_prep_queries(const float* q, int B, int b_pad, int dim, int normalize, float* qn, __nv_bfloat16* hi,
                        __nv_bfloat16* lo, cudaStream_t stream, int lane_order, float* q_stats, int fp16_planes) {
  // a zero query stays zero (backends.py:422-424: divide only when the norm is positive)
  return launch(q, B, b_pad, dim, normalize, 0, lane_order, qn, hi, lo, nullptr, stream, q_stats, nullptr, fp16_planes);
}

}  // namespace dewi
