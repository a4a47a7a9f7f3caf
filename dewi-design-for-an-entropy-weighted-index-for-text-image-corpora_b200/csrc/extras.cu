// "Next" rows of SURVEY.md section 8f that reuse the path's kernels:
//   dewi_local_weights  -- local_weights_from_surprisal (reference src/dewi/local_weights.py:5-26): robust
//                          z-score of a surprisal array (median / MAD from K3), clip to +-5, softplus; fp32.
//   dewi_cluster_pairs  -- connected components of the near-duplicate pair list the join emits (the
//                          clusters consumed by metrics.duplicate_rate / cluster_coverage, metrics.py:173-212).
#include <algorithm>

#include "internal.h"

namespace dewi {
namespace {

__global__ void local_weights_kernel(const float* __restrict__ s, long long n, float med, float den,
                                     float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float z = __fdiv_rn(__fsub_rn(__ldg(s + i), med), den);   // (s - med) / (1.4826 * mad)
    const bool is_nan = z != z;                               // (fminf / fmaxf drop a NaN, np.clip propagates it)
    z = fminf(fmaxf(z, -5.f), 5.f);                           // np.clip(z, -5, 5)
    out[i] = is_nan ? nanf("") : log1pf(expf(z));             // np.log1p(np.exp(z))
  }
}

__device__ __forceinline__ int find_root(const int* __restrict__ parent, int x) {
  int p = parent[x];
  while (p != x) {
    x = p;
    p = parent[x];
  }
  return x;
}

__global__ void cluster_init_kernel(int* parent, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = static_cast<int>(i);
}

// Hooks the larger root under the smaller one (parents only ever decrease, so the forest stays acyclic).
__global__ void cluster_link_kernel(const long long* __restrict__ pi, const long long* __restrict__ pj, long long n_pairs,
                                    int* parent, int* changed) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_pairs) return;
  int a = find_root(parent, static_cast<int>(pi[e]));
  int b = find_root(parent, static_cast<int>(pj[e]));
  while (a != b) {
    if (a < b) { const int t = a; a = b; b = t; }   // a > b: hook a under b
    const int old = atomicMin(&parent[a], b);
    *changed = 1;
    if (old == a) break;                             // a was still a root: done
    a = find_root(parent, old);                      // somebody hooked a meanwhile: retry from the new roots
    b = find_root(parent, b);
  }
}

__global__ void cluster_compress_kernel(int* parent, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = find_root(parent, static_cast<int>(i));
}

}  // namespace
}  // namespace dewi

using namespace dewi;

extern "C" int dewi_local_weights(const float* s, int64_t n, float* out, int device, void* stream_) {
  if (!s || !out) return fail("null argument");
  if (n <= 0) return fail("local_weights needs at least one value");
  double med = 0, mad = 0;
  // K3 returns a zero MAD as 1e-8 (scorer.py:24); here the reference ADDS 1e-8 to the float32 MAD
  // (local_weights.py:21), which gives the same float32 value in that case
  DEWI_TRY(dewi_fit_stats(s, n, 1, n, &med, &mad, device, stream_));
  const float mad32 = (mad == 1e-8) ? 1e-8f : (static_cast<float>(mad) + 1e-8f);
  const float den = 1.4826f * mad32;  // float32(1.4826) * float32 mad, as numpy's weak-scalar product
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n, threads), static_cast<int64_t>(current_sm_count()) * 16));
  local_weights_kernel<<<blocks, threads, 0, stream>>>(s, n, static_cast<float>(med), den, out);
  DEWI_CUDA(cudaGetLastError());
  DEWI_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

extern "C" int dewi_cluster_pairs(const int64_t* pair_i, const int64_t* pair_j, int64_t n_pairs, int64_t n, int32_t* labels,
                                  int device, void* stream_) {
  if (!labels || (n_pairs > 0 && (!pair_i || !pair_j))) return fail("null argument");
  if (n <= 0 || n >= (int64_t(1) << 31)) return fail("cluster_pairs supports 1 .. 2^31-1 documents");
  DEWI_TRY(dewi_device_check(device, nullptr, nullptr, nullptr));
  DEWI_CUDA(cudaSetDevice(device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int threads = 256;
  cluster_init_kernel<<<static_cast<int>(ceil_div(n, threads)), threads, 0, stream>>>(labels, n);
  if (n_pairs > 0) {
    int* changed = nullptr;
    DEWI_CUDA(cudaMalloc(&changed, sizeof(int)));
    int rc = 0;
    for (int iter = 0; iter < 64; ++iter) {
      int h = 0;
      cudaMemsetAsync(changed, 0, sizeof(int), stream);
      cluster_link_kernel<<<static_cast<int>(ceil_div(n_pairs, threads)), threads, 0, stream>>>(
          reinterpret_cast<const long long*>(pair_i), reinterpret_cast<const long long*>(pair_j), n_pairs, labels, changed);
      cluster_compress_kernel<<<static_cast<int>(ceil_div(n, threads)), threads, 0, stream>>>(labels, n);
      if (cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
          cudaStreamSynchronize(stream) != cudaSuccess) {
        rc = fail(std::string("cluster_pairs: ") + cudaGetErrorString(cudaGetLastError()));
        break;
      }
      if (!h) break;
    }
    cudaFree(changed);
    if (rc) return rc;
  }
  DEWI_CUDA(cudaGetLastError());
  DEWI_CUDA(cudaStreamSynchronize(stream));
  return 0;
}
