// K1' -- exact fp32 sweep on the CUDA cores with the candidate selection fused in.
//
// Same contract as the tcgen05 sweep (search_tc.cu) for the cases tensor cores do not fit:
// `space="l2"` (scores = -sum((E - q)^2), reference src/dewi/backends.py:434-436), dims that are
// not a multiple of 64, and tiny corpora.  Scores are computed directly from the stored rows in
// fp32, so the lists it emits are already exact (no re-score pass needed).
//
// One warp owns a contiguous row range; lanes stride over the row (coalesced 128-bit loads), two
// rows in flight per iteration; each warp keeps a top-kc list per query in shared memory guarded by
// a register threshold; the block merges its warps' lists before writing [block][k][query] partials.
#include <algorithm>

#include "internal.h"

namespace dewi {
namespace {

constexpr int kSimtThreads = 512;
constexpr int kSimtWarps = kSimtThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename RowT, int VEC>
struct RowLoad;
template <>
struct RowLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&x)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  }
};
template <>
struct RowLoad<float, 1> {
  static __device__ __forceinline__ void ld(const float* p, float (&x)[1]) { x[0] = __ldg(p); }
};
template <>
struct RowLoad<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&x)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[2 * i] = __uint_as_float(w[i] << 16);
      x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <>
struct RowLoad<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&x)[1]) { x[0] = __bfloat162float(*p); }
};

template <int VEC>
__device__ __forceinline__ void q_load(const float* p, float (&x)[VEC]) {
  if (VEC == 1) {
    x[0] = __ldg(p);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
      x[i] = v.x; x[i + 1] = v.y; x[i + 2] = v.z; x[i + 3] = v.w;
    }
  }
}

// Warp-cooperative insertion into an unsorted list of kc entries (all lanes hold the same v).
__device__ __forceinline__ void list_insert(float* ls, int* li, int kc, int lane, float v, int idx, float& thr) {
  float m = INFINITY;
  int p = 0x7fffffff;
  for (int k = lane; k < kc; k += 32) {
    const float x = ls[k];
    if (x < m) { m = x; p = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int op = __shfl_xor_sync(0xffffffffu, p, o);
    if (om < m || (om == m && op < p)) { m = om; p = op; }
  }
  if (lane == 0) { ls[p] = v; li[p] = idx; }
  __syncwarp();
  m = INFINITY;
  for (int k = lane; k < kc; k += 32) m = fminf(m, ls[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  thr = m;
}

template <typename RowT, int VEC, bool IS_L2, int QB>
__global__ void __launch_bounds__(kSimtThreads)
search_simt_kernel(const RowT* __restrict__ rows, long long n_rows, int dim, const float* __restrict__ qn, int B,
                   int kc, int n_qb, float* __restrict__ part_s, int* __restrict__ part_i, const SweepBlend blend) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ls_all = sm;                                                   // [warps][QB][kc]
  int* li_all = reinterpret_cast<int*>(sm + kSimtWarps * QB * kc);      // [warps][QB][kc]
  float* ls = ls_all + (warp * QB) * kc;
  int* li = li_all + (warp * QB) * kc;

  const long long total_warps = static_cast<long long>(gridDim.x) * kSimtWarps;
  const long long gw = static_cast<long long>(blockIdx.x) * kSimtWarps + warp;
  const long long r0 = gw * n_rows / total_warps;
  const long long r1 = (gw + 1) * n_rows / total_warps;

  for (int g0 = 0; g0 < B; g0 += QB) {
    const int nq = min(QB, B - g0);
    for (int k = lane; k < QB * kc; k += 32) { ls[k] = -INFINITY; li[k] = -1; }
    __syncwarp();
    float thr[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) thr[q] = -INFINITY;

    for (long long r = r0; r < r1; r += 2) {
      const bool two = (r + 1 < r1);
      const RowT* pa = rows + static_cast<size_t>(r) * dim;
      const RowT* pb = rows + static_cast<size_t>(two ? r + 1 : r) * dim;
      float acc_a[QB], acc_b[QB];
#pragma unroll
      for (int q = 0; q < QB; ++q) { acc_a[q] = 0.f; acc_b[q] = 0.f; }
      // rerank_scope = "full": the lists are kept by the blended key w_sim * sim + w_dewi * dewi (+ pref * ent)
      float bias_a = 0.f, bias_b = 0.f;
      if (blend.enabled) {
        bias_a = blend.w_dewi * __ldg(blend.dewi + r);
        bias_b = blend.w_dewi * __ldg(blend.dewi + (two ? r + 1 : r));
        if (blend.use_pref) {
          bias_a = fmaf(blend.pref, __ldg(blend.ent + r), bias_a);
          bias_b = fmaf(blend.pref, __ldg(blend.ent + (two ? r + 1 : r)), bias_b);
        }
      }
      for (int d = lane * VEC; d < dim; d += 32 * VEC) {
        float xa[VEC], xb[VEC];
        RowLoad<RowT, VEC>::ld(pa + d, xa);
        RowLoad<RowT, VEC>::ld(pb + d, xb);
#pragma unroll
        for (int q = 0; q < QB; ++q) {
          if (q < nq) {
            float qv[VEC];
            q_load<VEC>(qn + static_cast<size_t>(g0 + q) * dim + d, qv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              if (IS_L2) {
                const float da = xa[i] - qv[i], db = xb[i] - qv[i];
                acc_a[q] = fmaf(da, da, acc_a[q]);
                acc_b[q] = fmaf(db, db, acc_b[q]);
              } else {
                acc_a[q] = fmaf(xa[i], qv[i], acc_a[q]);
                acc_b[q] = fmaf(xb[i], qv[i], acc_b[q]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        if (q < nq) {
          float va = warp_sum(acc_a[q]), vb = warp_sum(acc_b[q]);
          if (IS_L2) { va = -va; vb = -vb; }
          if (blend.enabled) { va = fmaf(blend.w_sim, va, bias_a); vb = fmaf(blend.w_sim, vb, bias_b); }
          if (va > thr[q]) list_insert(ls + q * kc, li + q * kc, kc, lane, va, static_cast<int>(r), thr[q]);
          if (two && vb > thr[q]) list_insert(ls + q * kc, li + q * kc, kc, lane, vb, static_cast<int>(r + 1), thr[q]);
        }
      }
    }
    __syncthreads();
    // Block merge: warp w reduces the kSimtWarps lists of query g0 + w to one top-kc list.
    if (warp < nq) {
      const int q = warp;
      const int b = g0 + q;
      const int qb = b / kQueryBlock, ql = query_lane(b % kQueryBlock);
      float* ps = part_s + (static_cast<size_t>(blockIdx.x) * n_qb + qb) * kc * kQueryBlock;
      int* pi = part_i + (static_cast<size_t>(blockIdx.x) * n_qb + qb) * kc * kQueryBlock;
      const int total = kSimtWarps * kc;
      for (int k = 0; k < kc; ++k) {
        float m = -INFINITY;
        int p = -1;
        for (int e = lane; e < total; e += 32) {
          const int w = e / kc, kk = e - w * kc;
          const float x = ls_all[(w * QB + q) * kc + kk];
          if (x > m) { m = x; p = e; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float om = __shfl_xor_sync(0xffffffffu, m, o);
          const int op = __shfl_xor_sync(0xffffffffu, p, o);
          if (om > m || (om == m && op >= 0 && (p < 0 || op < p))) { m = om; p = op; }
        }
        int idx = -1;
        if (p >= 0) {
          const int w = p / kc, kk = p - w * kc;
          idx = li_all[(w * QB + q) * kc + kk];
          __syncwarp();
          if (lane == 0) ls_all[(w * QB + q) * kc + kk] = -INFINITY;  // consumed
        }
        if (lane == 0) {
          ps[k * kQueryBlock + ql] = (p >= 0) ? m : -INFINITY;
          pi[k * kQueryBlock + ql] = idx;
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

template <typename RowT, int VEC, bool IS_L2>
int launch_t(const RowT* rows, int64_t n_rows, int dim, const float* qn, int B, int kc, int n_chunks, int n_qb,
             float* part_s, int* part_i, cudaStream_t stream, const SweepBlend& blend) {
  if (B >= 4) {
    constexpr int QB = 4;
    const size_t smem = static_cast<size_t>(kSimtWarps) * QB * kc * 8;
    auto kern = search_simt_kernel<RowT, VEC, IS_L2, QB>;
    DEWI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<n_chunks, kSimtThreads, smem, stream>>>(rows, n_rows, dim, qn, B, kc, n_qb, part_s, part_i, blend);
  } else {
    constexpr int QB = 1;
    const size_t smem = static_cast<size_t>(kSimtWarps) * QB * kc * 8;
    auto kern = search_simt_kernel<RowT, VEC, IS_L2, QB>;
    DEWI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<n_chunks, kSimtThreads, smem, stream>>>(rows, n_rows, dim, qn, B, kc, n_qb, part_s, part_i, blend);
  }
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int simt_plan(int64_t n_rows, int B, int sm_count, int* n_chunks) {
  (void)B;
  // one chunk == one block of 16 warps; keep >= ~8 rows per warp, at most 2 blocks per SM
  const int64_t by_rows = ceil_div(n_rows, static_cast<int64_t>(kSimtWarps) * 8);
  *n_chunks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(by_rows, static_cast<int64_t>(sm_count) * 2)));
  return 0;
}

int simt_launch(const void* rows, int rows_are_bf16, int64_t n_rows, int dim, int space, const float* qn, int B, int kc,
                int n_chunks, float* part_s, int* part_i, cudaStream_t stream, const SweepBlend* blend_) {
  if (static_cast<size_t>(kSimtWarps) * 4 * kc * 8 > 200 * 1024)
    return fail("k too large: a search keeps at most 400 candidates (min(2k, N), backends.py:440) per query, i.e. k <= 200 on "
                "a corpus of more than 400 rows (documented limit; the reference accepts any k <= N)");
  const SweepBlend blend = blend_ ? *blend_ : SweepBlend();
  const int n_qb = static_cast<int>(ceil_div(B, kQueryBlock));
  const bool l2 = (space == DEWI_SPACE_L2);
  if (rows_are_bf16) {
    const auto* r = static_cast<const __nv_bfloat16*>(rows);
    if (dim % 8 == 0)
      return l2 ? launch_t<__nv_bfloat16, 8, true>(r, n_rows, dim, qn, B, kc, n_chunks, n_qb, part_s, part_i, stream, blend)
                : launch_t<__nv_bfloat16, 8, false>(r, n_rows, dim, qn, B, kc, n_chunks, n_qb, part_s, part_i, stream, blend);
    return l2 ? launch_t<__nv_bfloat16, 1, true>(r, n_rows, dim, qn, B, kc, n_chunks, n_qb, part_s, part_i, stream, blend)
              : launch_t<__nv_bfloat16, 1, false>(r, n_rows, dim, qn, B, kc, n_chunks, n_qb, part_s, part_i, stream, blend);
  }
  const auto* r = static_cast<const float*>(rows);
  if (dim % 4 == 0)
    return l2 ? launch_t<float, 4, true>(r, n_rows, dim, qn, B, kc, n_chunks, n_qb, part_s, part_i, stream, blend)
              : launch_t<float, 4, false>(r, n_rows, dim, qn, B, kc, n_chunks, n_qb, part_s, part_i, stream, blend);
  return l2 ? launch_t<float, 1, true>(r, n_rows, dim, qn, B, kc, n_chunks, n_qb, part_s, part_i, stream, blend)
            : launch_t<float, 1, false>(r, n_rows, dim, qn, B, kc, n_chunks, n_qb, part_s, part_i, stream, blend);
}

}  // namespace dewi
