// K5 -- redundancy similarity pass.
//
// dewi_similarity_dense: `F.normalize(t) @ F.normalize(i).T` exactly as
//   RedundancyEstimator.compute_cross_modal_similarity does after the CLIP forward
//   (reference src/dewi/signals/redundancy.py:36-38), dense fp32 output for small inputs.
// dewi_join: the same product with the [M, N] matrix never materialised: a fused epilogue keeps
//   per-row max / argmax / count(sim >= tau) and appends (i, j, sim) pairs.  The reference defines
//   no such reduction; the definition lives in include/dewi_b200.h.
//
// This translation unit is the fp32 CUDA-core implementation (64x64 tiles, 4x4 register blocking,
// shared-memory staged).  It is exact to fp32 rounding and serves as the small-size path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "internal.h"

namespace dewi {
namespace {

__global__ void normalize_rows_kernel(const float* __restrict__ src, long long n, int d, float eps,
                                      float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    const float* s = src + static_cast<size_t>(row) * d;
    float ss = 0.f;
    for (int k = lane; k < d; k += 32) ss = fmaf(s[k], s[k], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float den = fmaxf(sqrtf(ss), eps);  // torch F.normalize: x / max(||x||, eps)
    for (int k = lane; k < d; k += 32) dst[static_cast<size_t>(row) * d + k] = __fdiv_rn(s[k], den);
  }
}

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ unsigned int orderable(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// MODE 0: dense store.  MODE 1: thresholded join epilogue.
template <int MODE>
__global__ void __launch_bounds__(256)
sim_tile_kernel(const float* __restrict__ a, long long m, const float* __restrict__ b, long long n, int d,
                float* __restrict__ out, float tau, int self_join, long long a_offset, unsigned long long* __restrict__ row_best,
                int* __restrict__ row_count, long long* __restrict__ pair_i, long long* __restrict__ pair_j,
                float* __restrict__ pair_sim, long long pair_cap, unsigned long long* __restrict__ pair_count) {
  __shared__ float sa[TK][TM + 1];
  __shared__ float sb[TK][TN + 1];
  const long long i0 = static_cast<long long>(blockIdx.y) * TM;
  const long long j0 = static_cast<long long>(blockIdx.x) * TN;
  if (MODE == 1 && self_join && j0 + TN <= i0) return;  // strictly-lower tiles: covered by symmetry
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int k0 = 0; k0 < d; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int r = e / TK, k = e % TK;
      sa[k][r] = (i0 + r < m && k0 + k < d) ? a[static_cast<size_t>(i0 + r) * d + k0 + k] : 0.f;
      sb[k][r] = (j0 + r < n && k0 + k < d) ? b[static_cast<size_t>(j0 + r) * d + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) av[r] = sa[k][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = sb[k][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long i = i0 + ty * 4 + r;
    if (i >= m) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long j = j0 + tx * 4 + c;
      if (j >= n) continue;
      const float s = acc[r][c];
      if (MODE == 0) {
        out[static_cast<size_t>(i) * n + j] = s;
      } else {
        if (self_join && i == j) continue;
        const unsigned long long key =
            (static_cast<unsigned long long>(orderable(s)) << 32) | static_cast<unsigned int>(~static_cast<unsigned int>(j));
        if (!self_join) {
          // a_offset >= 0: A is the slice of B starting there -- row i is B's row a_offset + i
          if (a_offset >= 0 && j == i + a_offset) continue;
          atomicMax(&row_best[i], key);
          if (s >= tau) {
            atomicAdd(&row_count[i], 1);
            if (a_offset < 0 || j > i + a_offset) {
              const unsigned long long slot = atomicAdd(pair_count, 1ull);
              if (static_cast<long long>(slot) < pair_cap) {
                pair_i[slot] = i + (a_offset > 0 ? a_offset : 0);
                pair_j[slot] = j;
                pair_sim[slot] = s;
              }
            }
          }
        } else {
          // upper-triangle tiles update both rows; diagonal tiles hold (i, j) and (j, i) themselves
          const bool diag_tile = (j0 < i0 + TM) && (i0 < j0 + TN);
          atomicMax(&row_best[i], key);
          if (s >= tau) atomicAdd(&row_count[i], 1);
          if (!diag_tile) {
            const unsigned long long key_t =
                (static_cast<unsigned long long>(orderable(s)) << 32) | static_cast<unsigned int>(~static_cast<unsigned int>(i));
            atomicMax(&row_best[j], key_t);
            if (s >= tau) atomicAdd(&row_count[j], 1);
          }
          if (s >= tau && j > i) {
            const unsigned long long slot = atomicAdd(pair_count, 1ull);
            if (static_cast<long long>(slot) < pair_cap) { pair_i[slot] = i; pair_j[slot] = j; pair_sim[slot] = s; }
          }
        }
      }
    }
  }
}

__global__ void join_finish_kernel(const unsigned long long* __restrict__ row_best, long long m,
                                   float* __restrict__ row_max, long long* __restrict__ row_argmax) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long key = row_best[i];
  if (key == 0ull) {
    row_max[i] = -INFINITY;
    row_argmax[i] = -1;
    return;
  }
  const unsigned int o = static_cast<unsigned int>(key >> 32);
  row_max[i] = __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
  row_argmax[i] = static_cast<long long>(~static_cast<unsigned int>(key & 0xffffffffull));
}

int normalize_into(const float* src, int64_t n, int d, float* dst, cudaStream_t stream) {
  const int threads = 256;
  const int64_t blocks = std::min<int64_t>(ceil_div(n * 32, threads), static_cast<int64_t>(current_sm_count()) * 16);
  normalize_rows_kernel<<<static_cast<int>(std::max<int64_t>(blocks, 1)), threads, 0, stream>>>(src, n, d, 1e-12f, dst);
  DEWI_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace dewi

using namespace dewi;

extern "C" int dewi_similarity_dense(const float* a, int64_t m, const float* b, int64_t n, int d, float* out,
                                     int device, void* stream_) {
  if (!a || !b || !out) return fail("null argument");
  if (m <= 0 || n <= 0 || d <= 0) return fail("similarity needs positive sizes");
  DEWI_TRY(dewi_device_check(device, nullptr, nullptr, nullptr));
  DEWI_CUDA(cudaSetDevice(device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float *an = nullptr, *bn = nullptr;
  DEWI_CUDA(cudaMalloc(&an, static_cast<size_t>(m) * d * 4));
  if (cudaMalloc(&bn, static_cast<size_t>(n) * d * 4) != cudaSuccess) {
    cudaFree(an);
    return fail("cudaMalloc failed");
  }
  int rc = normalize_into(a, m, d, an, stream);
  if (!rc) rc = normalize_into(b, n, d, bn, stream);
  if (!rc) {
    dim3 grid(static_cast<unsigned>(ceil_div(n, TN)), static_cast<unsigned>(ceil_div(m, TM)));
    sim_tile_kernel<0><<<grid, 256, 0, stream>>>(an, m, bn, n, d, out, 0.f, 0, -1, nullptr, nullptr, nullptr, nullptr,
                                                 nullptr, 0, nullptr);
    if (cudaGetLastError() != cudaSuccess) rc = fail("similarity kernel launch failed");
  }
  cudaStreamSynchronize(stream);
  cudaFree(an);
  cudaFree(bn);
  return rc;
}

namespace dewi {
namespace {

// Scratch memory of one join call: carved out of the caller's workspace (the Python wrapper hands in a
// torch tensor, so repeated calls cost no cudaMalloc / cudaFree -- those dominated the wall time of a
// join by far: a cudaFree of the GB-sized operand planes synchronises and returns the pages to the
// driver), or, with no workspace, allocated here and freed on return.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  bool dry = false;             // only add up the sizes (dewi_join_workspace_bytes)
  void* owned[8] = {};
  int n_owned = 0;
  void* take(size_t bytes) {
    bytes = static_cast<size_t>(round_up(static_cast<int64_t>(std::max<size_t>(bytes, 1)), 256));
    if (dry) { used += bytes; return reinterpret_cast<void*>(uintptr_t(256)); }
    if (base) {
      if (used + bytes > cap) return nullptr;
      void* p = base + used;
      used += bytes;
      return p;
    }
    void* p = nullptr;
    if (n_owned >= 8 || cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    owned[n_owned++] = p;
    return p;
  }
  ~Arena() {
    for (int i = 0; i < n_owned; ++i) cudaFree(owned[i]);
  }
};
#define DEWI_TAKE(var, type, arena, bytes)                                              \
  type var = static_cast<type>((arena).take(bytes));                                    \
  if (!var) return fail("join: workspace too small or cudaMalloc failed")

// Tensor-core path: rows normalised into bf16 planes, A on the query side of the CTA-pair sweep.
// sym_lo >= 0: the symmetric self-join of `a` (= b, m = n) restricted to the row blocks [sym_lo, sym_hi): the
// outputs then cover ALL n rows (partial statistics when the range is not the whole matrix).
int join_tensor(const float* a, int64_t m, const float* b, int64_t n, int d, float tau, int self_join, int64_t a_offset,
                int bf16_only, int64_t sym_lo, int64_t sym_hi,
                float* row_max, int64_t* row_argmax, int32_t* row_count, int64_t* pair_i, int64_t* pair_j,
                float* pair_sim, int64_t pair_cap, int64_t* pair_count_host, int device, cudaStream_t stream, Arena& w) {
  int sms = 148;
  if (!w.dry) DEWI_TRY(dewi_device_check(device, &sms, nullptr, nullptr));
  const int64_t m_pad = round_up(m, 2 * kQueryBlock);
  const size_t plane_a = static_cast<size_t>(m_pad) * d * 2, plane_b = static_cast<size_t>(n) * d * 2;
  DEWI_TAKE(a_hi, __nv_bfloat16*, w, plane_a);
  __nv_bfloat16 *a_lo = nullptr, *b_hi = a_hi, *b_lo = nullptr;
  if (!bf16_only) {
    DEWI_TAKE(p, __nv_bfloat16*, w, plane_a);
    a_lo = b_lo = p;
  }
  if (!self_join) {
    DEWI_TAKE(p, __nv_bfloat16*, w, plane_b);
    b_hi = p;
    if (!bf16_only) {
      DEWI_TAKE(q, __nv_bfloat16*, w, plane_b);
      b_lo = q;
    }
  }
  DEWI_TAKE(best, unsigned long long*, w, static_cast<size_t>(m) * 8);
  DEWI_TAKE(count, unsigned long long*, w, 8);
  if (w.dry) return 0;
  DEWI_TRY(launch_prep_queries(a, static_cast<int>(m), static_cast<int>(m_pad), d, 1, nullptr, a_hi, a_lo, stream));
  if (!self_join)
    DEWI_TRY(launch_prep_queries(b, static_cast<int>(n), static_cast<int>(n), d, 1, nullptr, b_hi, b_lo, stream));
  DEWI_CUDA(cudaMemsetAsync(best, 0, static_cast<size_t>(m) * 8, stream));
  DEWI_CUDA(cudaMemsetAsync(count, 0, 8, stream));
  DEWI_CUDA(cudaMemsetAsync(row_count, 0, static_cast<size_t>(m) * 4, stream));
  CUtensorMap ma0, ma1, mb0, mb1;
  DEWI_TRY(tc_encode_rows_map(&ma0, a_hi, m_pad, d, kQueryBlock));
  DEWI_TRY(tc_encode_rows_map(&mb0, b_hi, n, d, tc2_box_rows()));
  ma1 = ma0;
  mb1 = mb0;
  if (!bf16_only) {
    DEWI_TRY(tc_encode_rows_map(&ma1, a_lo, m_pad, d, kQueryBlock));
    DEWI_TRY(tc_encode_rows_map(&mb1, b_lo, n, d, tc2_box_rows()));
  }
  // DEWI_JOIN_TIMING=1: CUDA-event time of the join kernel alone on stderr (profiling / bench_paths.py)
  const bool timing = env_set("DEWI_JOIN_TIMING");
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (timing) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  if (sym_lo >= 0) {
    // A is the row range of the same planes: the kernel offsets its query rows by sym_lo
    const int64_t rows = sym_hi - sym_lo;
    if (rows > 0)
      DEWI_TRY(tc2_join_launch(bf16_only ? 0 : 2, mb0, mb1, mb0, mb1, rows, round_up(rows, 2 * kQueryBlock), n, d, sms, tau, 1,
                               sym_lo, 1, best, row_count, pair_i, pair_j, pair_sim, pair_cap, count, stream));
  } else {
    DEWI_TRY(tc2_join_launch(bf16_only ? 0 : 2, mb0, mb1, ma0, ma1, m, m_pad, n, d, sms, tau,
                             (self_join || a_offset >= 0) ? 1 : 0, a_offset > 0 ? a_offset : 0, 0, best, row_count,
                             pair_i, pair_j, pair_sim, pair_cap, count, stream));
  }
  if (timing) cudaEventRecord(ev1, stream);
  join_finish_kernel<<<static_cast<int>(ceil_div(m, 256)), 256, 0, stream>>>(best, m, row_max,
                                                                             reinterpret_cast<long long*>(row_argmax));
  DEWI_CUDA(cudaGetLastError());
  unsigned long long cnt = 0;
  DEWI_CUDA(cudaMemcpyAsync(&cnt, count, 8, cudaMemcpyDeviceToHost, stream));
  DEWI_CUDA(cudaStreamSynchronize(stream));
  *pair_count_host = static_cast<int64_t>(cnt);
  if (timing) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    fprintf(stderr, "[dewi_join] kernel %.3f ms (m=%lld n=%lld d=%d %s%s)\n", ms, static_cast<long long>(m),
            static_cast<long long>(n), d, bf16_only ? "bf16" : "hi/lo", sym_lo >= 0 ? " symmetric" : "");
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  }
  return 0;
}

}  // namespace
}  // namespace dewi

namespace dewi {
namespace {

bool join_uses_tensor_cores(int64_t m, int64_t n, int d, int flags) {
  const bool tensor_ok = tc_supported(d, n) && !(flags & DEWI_JOIN_FORCE_SIMT);
  const bool big = static_cast<double>(m) * static_cast<double>(n) >= 4.0e6;
  return tensor_ok && (big || (flags & DEWI_JOIN_FORCE_TC));
}

// fp32 CUDA-core join (small inputs, d % 64 != 0).
int join_simt(const float* a, int64_t m, const float* b, int64_t n, int d, float tau, int self_join, int64_t a_offset,
              float* row_max, int64_t* row_argmax, int32_t* row_count, int64_t* pair_i, int64_t* pair_j, float* pair_sim,
              int64_t pair_cap, int64_t* pair_count_host, cudaStream_t stream, Arena& w) {
  DEWI_TAKE(an, float*, w, static_cast<size_t>(m) * d * 4);
  float* bn = an;
  if (!self_join) {
    DEWI_TAKE(p, float*, w, static_cast<size_t>(n) * d * 4);
    bn = p;
  }
  DEWI_TAKE(best, unsigned long long*, w, static_cast<size_t>(m) * 8);
  DEWI_TAKE(count, unsigned long long*, w, 8);
  if (w.dry) return 0;
  DEWI_CUDA(cudaMemsetAsync(best, 0, static_cast<size_t>(m) * 8, stream));
  DEWI_CUDA(cudaMemsetAsync(count, 0, 8, stream));
  DEWI_CUDA(cudaMemsetAsync(row_count, 0, static_cast<size_t>(m) * 4, stream));
  DEWI_TRY(normalize_into(a, m, d, an, stream));
  if (!self_join) DEWI_TRY(normalize_into(b, n, d, bn, stream));
  dim3 grid(static_cast<unsigned>(ceil_div(n, TN)), static_cast<unsigned>(ceil_div(m, TM)));
  sim_tile_kernel<1><<<grid, 256, 0, stream>>>(an, m, bn, n, d, nullptr, tau, self_join, a_offset, best, row_count,
                                               reinterpret_cast<long long*>(pair_i), reinterpret_cast<long long*>(pair_j),
                                               pair_sim, pair_cap, count);
  join_finish_kernel<<<static_cast<int>(ceil_div(m, 256)), 256, 0, stream>>>(best, m, row_max,
                                                                             reinterpret_cast<long long*>(row_argmax));
  DEWI_CUDA(cudaGetLastError());
  unsigned long long cnt = 0;
  DEWI_CUDA(cudaMemcpyAsync(&cnt, count, 8, cudaMemcpyDeviceToHost, stream));
  DEWI_CUDA(cudaStreamSynchronize(stream));
  *pair_count_host = static_cast<int64_t>(cnt);
  return 0;
}

int join_dispatch(const float* a, int64_t m, const float* b, int64_t n, int d, float tau, int self_join, int64_t a_offset,
                  int flags, float* row_max, int64_t* row_argmax, int32_t* row_count, int64_t* pair_i, int64_t* pair_j,
                  float* pair_sim, int64_t pair_cap, int64_t* pair_count_host, int device, cudaStream_t stream, Arena& w) {
  if (join_uses_tensor_cores(m, n, d, flags))
    return join_tensor(a, m, b, n, d, tau, self_join, a_offset, (flags & DEWI_JOIN_BF16) ? 1 : 0,
                       (self_join && !(flags & DEWI_JOIN_NO_SYMMETRY)) ? 0 : -1, m, row_max, row_argmax, row_count,
                       pair_i, pair_j, pair_sim, pair_cap, pair_count_host, device, stream, w);
  if (flags & DEWI_JOIN_FORCE_TC) return fail("tensor-core join needs d % 64 == 0");
  return join_simt(a, m, b, n, d, tau, self_join, a_offset, row_max, row_argmax, row_count, pair_i, pair_j, pair_sim, pair_cap,
                   pair_count_host, stream, w);
}

}  // namespace
}  // namespace dewi

extern "C" int64_t dewi_join_workspace_bytes(int64_t m, int64_t n, int d, int self_join, int flags) {
  if (self_join) n = m;
  if (m <= 0 || n <= 0 || d <= 0) return 0;
  Arena w;
  w.dry = true;
  int64_t dummy = 0;
  if (join_dispatch(nullptr, m, nullptr, n, d, 0.f, self_join, -1, flags, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                    &dummy, 0, nullptr, w) != 0)
    return 0;
  return static_cast<int64_t>(w.used);
}

extern "C" int dewi_join(const float* a, int64_t m, const float* b, int64_t n, int d, float tau, int self_join,
                         int64_t a_offset, int flags, float* row_max, int64_t* row_argmax, int32_t* row_count, int64_t* pair_i, int64_t* pair_j,
                         float* pair_sim, int64_t pair_cap, int64_t* pair_count_host, void* workspace, int64_t workspace_bytes,
                         int device, void* stream_) {
  if (!a || !row_max || !row_argmax || !row_count || !pair_count_host) return fail("null argument");
  if (self_join) { b = a; n = m; a_offset = -1; }
  if (!b) return fail("null argument");
  if (a_offset >= 0 && a_offset + m > n) return fail("a_offset + m exceeds the rows of b");
  if (m <= 0 || n <= 0 || d <= 0) return fail("join needs positive sizes");
  if (n >= (int64_t(1) << 31) || m >= (int64_t(1) << 31)) return fail("join supports fewer than 2^31 rows per side");
  if (pair_cap > 0 && (!pair_i || !pair_j || !pair_sim)) return fail("pair buffers missing");
  DEWI_TRY(dewi_device_check(device, nullptr, nullptr, nullptr));
  DEWI_CUDA(cudaSetDevice(device));
  Arena w;
  w.base = static_cast<char*>(workspace);
  w.cap = workspace ? static_cast<size_t>(std::max<int64_t>(workspace_bytes, 0)) : 0;
  return join_dispatch(a, m, b, n, d, tau, self_join, a_offset, flags, row_max, row_argmax, row_count, pair_i, pair_j, pair_sim,
                       pair_cap, pair_count_host, device, static_cast<cudaStream_t>(stream_), w);
}

extern "C" int dewi_self_join_range(const float* x, int64_t n, int d, float tau, int64_t row_lo, int64_t row_hi, int flags,
                                    float* row_max, int64_t* row_argmax, int32_t* row_count, int64_t* pair_i, int64_t* pair_j,
                                    float* pair_sim, int64_t pair_cap, int64_t* pair_count_host, void* workspace,
                                    int64_t workspace_bytes, int device, void* stream_) {
  if (!x || !row_max || !row_argmax || !row_count || !pair_count_host) return fail("null argument");
  if (n <= 0 || d <= 0) return fail("join needs positive sizes");
  if (n >= (int64_t(1) << 31)) return fail("join supports fewer than 2^31 rows per side");
  if (row_lo < 0 || row_hi < row_lo || row_hi > n) return fail("row range outside the matrix");
  if (row_lo % 256 != 0 || (row_hi % 256 != 0 && row_hi != n))
    return fail("row range must start and end on multiples of 256 (or at the last row)");
  if (pair_cap > 0 && (!pair_i || !pair_j || !pair_sim)) return fail("pair buffers missing");
  if (!tc_supported(d, n)) return fail("the symmetric range join runs on the tensor cores: needs d % 64 == 0");
  DEWI_TRY(dewi_device_check(device, nullptr, nullptr, nullptr));
  DEWI_CUDA(cudaSetDevice(device));
  Arena w;
  w.base = static_cast<char*>(workspace);
  w.cap = workspace ? static_cast<size_t>(std::max<int64_t>(workspace_bytes, 0)) : 0;
  return join_tensor(x, n, x, n, d, tau, 1, -1, (flags & DEWI_JOIN_BF16) ? 1 : 0, row_lo, row_hi, row_max, row_argmax, row_count,
                     pair_i, pair_j, pair_sim, pair_cap, pair_count_host, device, static_cast<cudaStream_t>(stream_), w);
}
