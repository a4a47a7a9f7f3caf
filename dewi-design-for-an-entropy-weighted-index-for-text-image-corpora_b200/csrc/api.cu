// C ABI of libdewi_b200.so: index handle, search orchestration, error plumbing.
// Entry points are documented in include/dewi_b200.h next to the reference code they replace.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "internal.h"

namespace dewi {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}

namespace {
std::mutex g_env_mu;
std::map<std::string, std::pair<bool, int>> g_env_cache;  // name -> (set, value)
std::pair<bool, int> env_lookup(const char* name) {
  std::lock_guard<std::mutex> lock(g_env_mu);
  auto it = g_env_cache.find(name);
  if (it != g_env_cache.end()) return it->second;
  const char* v = getenv(name);
  const std::pair<bool, int> r = v ? std::make_pair(true, atoi(v)) : std::make_pair(false, 0);
  g_env_cache.emplace(name, r);
  return r;
}
}  // namespace

int env_int(const char* name, int dflt) {
  const auto r = env_lookup(name);
  return r.first ? r.second : dflt;
}
bool env_set(const char* name) { return env_lookup(name).first; }

int current_sm_count() {
  static std::mutex mu;
  static int sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    sms[dev] = v;
  }
  return sms[dev];
}

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    const size_t want = need + need / 4;
    DEWI_CUDA(cudaMalloc(&p, want));
    bytes = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

// Minimum corpus size for which the tensor-core sweep is selected automatically.
constexpr int64_t kTcMinRows = 2048;
// Certified single-plane sweep of an fp32 corpus: worth its host synchronisation (the certificate's verdict is read
// back before the candidates are used) once a plane is tens of megabytes.
constexpr int64_t kCertMinElems = int64_t(32) << 20;

}  // namespace
}  // namespace dewi

using namespace dewi;

struct dewi_index {
  int dim = 0, space = 0, dtype = 0, device = 0, sm_count = 0;
  int64_t n = 0, cap = 0, id_base = 0;
  float* rows_f32 = nullptr;          // [cap, dim]  (fp32 mode)
  __nv_bfloat16* plane0 = nullptr;    // [cap, dim]  bf16 rows / hi plane
  __nv_bfloat16* plane1 = nullptr;    // [cap, dim]  lo plane (fp32 mode)
  float* dewi_col = nullptr;          // [cap]
  float* ent_col = nullptr;           // [cap]
  int* bad_flag = nullptr;
  // certified single-plane sweep (fp32 corpus): running maxima of ||row - hi|| and ||hi|| (float bits), a fail counter
  unsigned int* plane_max = nullptr;   // [2] plane maxima, [2] = certificate fail counter
  // host I/O (DEWI_FLAG_HOST_IO): pinned staging so the query upload and the result download are one small truly
  // asynchronous copy each (pageable copies go through the driver's own staging and cost ~10-20 us apiece -- more than
  // the kernels of a single-query search over a small corpus)
  void* pin_q = nullptr;
  size_t pin_q_bytes = 0;
  void* pin_out = nullptr;
  size_t pin_out_bytes = 0;
  long long cert_used = 0, cert_failed = 0;   // searches answered by the certified sweep / re-run with the full product
  bool plane_max_stale = false;               // plane_max changed on the device since plane_hi_max was read
  float plane_hi_max = 0.f;                   // host copy of max_r ||hi_r|| (fp16 planes: range guard)
  // cached tensor maps of the corpus planes
  CUtensorMap map_e0, map_e1;
  int64_t map_rows = -1;
  int map_box = 0;
  // workspaces
  DevBuf stage, stage2, qraw, qn, q0, q1, qstats, qbar, part_s, part_i, seed_max, seed_sim, cand_idx, cand_sim, loc_sim, loc_id, loc_dewi, loc_ent, out_id,
      out_score;
  int last_launches = 0;
  // host ingest pipeline: copies run on `copy_stream`, two staging buffers, one "kernel done with buffer" event each
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t stage_free[2] = {}, stage_full[2] = {};
  DevBuf push_ticket;               // block counter of the peer-push finalize kernel
  DevBuf sync_cnt;                  // rendezvous counters of the CTA-pair sweep
  bool push_ticket_zeroed = false;
  // optional CUDA-event bracket around the sweep kernel (bench.py's roofline figure)
  int profile = 0;
  int last_sweep_kind = 0;  // 1 = tcgen05 sweep, 2 = CUDA-core sweep, 3 = CTA-pair sweep, 4 = rows-on-M sweep
  static constexpr int kEvRing = 64;
  cudaEvent_t ev0[kEvRing] = {}, ev1[kEvRing] = {};
  long long searches = 0;  // sweeps bracketed since profiling was enabled
  // rerank_scope = "full" (DEWI_FLAG_SCOPE_FULL): blend weights of the key the sweep selects by (dewi_index_set_blend)
  SweepBlend blend;
  bool blend_set = false;
};

namespace {

int set_device(const dewi_index* h) {
  DEWI_CUDA(cudaSetDevice(h->device));
  return 0;
}

int grow(dewi_index* h, int64_t need_rows, cudaStream_t stream) {
  if (need_rows <= h->cap) return 0;
  const int64_t newcap = std::max<int64_t>(need_rows, h->cap + h->cap / 2);
  const size_t d = static_cast<size_t>(h->dim);
  auto regrow = [&](void** p, size_t elem_bytes_per_row) -> int {
    void* np = nullptr;
    DEWI_CUDA(cudaMalloc(&np, static_cast<size_t>(newcap) * elem_bytes_per_row));
    if (*p && h->n > 0)
      DEWI_CUDA(cudaMemcpyAsync(np, *p, static_cast<size_t>(h->n) * elem_bytes_per_row, cudaMemcpyDeviceToDevice, stream));
    if (*p) {
      DEWI_CUDA(cudaStreamSynchronize(stream));
      DEWI_CUDA(cudaFree(*p));
    }
    *p = np;
    return 0;
  };
  if (h->dtype == DEWI_DTYPE_FP32) {
    DEWI_TRY(regrow(reinterpret_cast<void**>(&h->rows_f32), d * 4));
    if (h->space == DEWI_SPACE_COSINE) DEWI_TRY(regrow(reinterpret_cast<void**>(&h->plane1), d * 2));
  }
  if (h->dtype == DEWI_DTYPE_BF16 || h->space == DEWI_SPACE_COSINE)
    DEWI_TRY(regrow(reinterpret_cast<void**>(&h->plane0), d * 2));
  {
    // payload columns default to zero (Payload() defaults, types.py:11-18)
    float* nd = nullptr;
    float* ne = nullptr;
    DEWI_CUDA(cudaMalloc(&nd, static_cast<size_t>(newcap) * 4));
    DEWI_CUDA(cudaMalloc(&ne, static_cast<size_t>(newcap) * 4));
    DEWI_CUDA(cudaMemsetAsync(nd, 0, static_cast<size_t>(newcap) * 4, stream));
    DEWI_CUDA(cudaMemsetAsync(ne, 0, static_cast<size_t>(newcap) * 4, stream));
    if (h->dewi_col && h->n > 0) {
      DEWI_CUDA(cudaMemcpyAsync(nd, h->dewi_col, static_cast<size_t>(h->n) * 4, cudaMemcpyDeviceToDevice, stream));
      DEWI_CUDA(cudaMemcpyAsync(ne, h->ent_col, static_cast<size_t>(h->n) * 4, cudaMemcpyDeviceToDevice, stream));
    }
    DEWI_CUDA(cudaStreamSynchronize(stream));
    if (h->dewi_col) cudaFree(h->dewi_col);
    if (h->ent_col) cudaFree(h->ent_col);
    h->dewi_col = nd;
    h->ent_col = ne;
  }
  h->cap = newcap;
  h->map_rows = -1;
  return 0;
}

int ensure_corpus_maps(dewi_index* h, int box_rows) {
  if (h->map_rows == h->n && h->map_box == box_rows) return 0;
  DEWI_TRY(tc_encode_rows_map(&h->map_e0, h->plane0, h->n, h->dim, box_rows));
  if (h->plane1)
    DEWI_TRY(tc_encode_rows_map(&h->map_e1, h->plane1, h->n, h->dim, box_rows));
  else
    h->map_e1 = h->map_e0;
  h->map_rows = h->n;
  h->map_box = box_rows;
  return 0;
}

}  // namespace

extern "C" {

int dewi_abi_version(void) { return DEWI_B200_ABI_VERSION; }

const char* dewi_last_error(void) { return g_last_error.c_str(); }

int dewi_device_check(int device, int* sm_count, size_t* free_bytes, size_t* total_bytes) {
  // cudaGetDeviceProperties costs milliseconds: look each device up once
  static std::mutex mu;
  static int n_devices = -1;
  static int cc_major[64], cc_minor[64], sms[64];
  static char names[64][96];
  {
    std::lock_guard<std::mutex> lock(mu);
    if (n_devices < 0) {
      int count = 0;
      cudaError_t e = cudaGetDeviceCount(&count);
      if (e != cudaSuccess || count <= 0)
        return fail(std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                    " -- this library has no CPU fallback");
      count = std::min(count, 64);
      for (int d = 0; d < count; ++d) {
        cudaDeviceProp prop;
        DEWI_CUDA(cudaGetDeviceProperties(&prop, d));
        cc_major[d] = prop.major;
        cc_minor[d] = prop.minor;
        sms[d] = prop.multiProcessorCount;
        strncpy(names[d], prop.name, sizeof(names[d]) - 1);
        names[d][sizeof(names[d]) - 1] = 0;
      }
      n_devices = count;
    }
  }
  if (device < 0 || device >= n_devices) return fail("device ordinal out of range");
  if (cc_major[device] != 10)
    return fail(std::string("device '") + names[device] + "' is compute capability " + std::to_string(cc_major[device]) +
                "." + std::to_string(cc_minor[device]) + "; libdewi_b200 is built for sm_100a (B200) only");
  if (sm_count) *sm_count = sms[device];
  if (free_bytes || total_bytes) {
    size_t f = 0, t = 0;
    DEWI_CUDA(cudaSetDevice(device));
    DEWI_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
  }
  return 0;
}

int dewi_index_create(int dim, int space, int dtype, int device, dewi_index_t** out) {
  if (!out) return fail("out is NULL");
  if (dim <= 0) return fail("dim must be positive");
  if (space != DEWI_SPACE_COSINE && space != DEWI_SPACE_L2) return fail("unknown space");
  if (dtype != DEWI_DTYPE_FP32 && dtype != DEWI_DTYPE_BF16) return fail("unknown dtype");
  int sms = 0;
  DEWI_TRY(dewi_device_check(device, &sms, nullptr, nullptr));
  DEWI_CUDA(cudaSetDevice(device));
  dewi_index* h = new dewi_index();
  h->dim = dim;
  h->space = space;
  h->dtype = dtype;
  h->device = device;
  h->sm_count = sms;
  if (cudaMalloc(&h->bad_flag, sizeof(int)) != cudaSuccess || cudaMalloc(&h->plane_max, 4 * sizeof(unsigned int)) != cudaSuccess) {
    cudaFree(h->bad_flag);
    delete h;
    return fail("cudaMalloc failed");
  }
  cudaMemset(h->bad_flag, 0, sizeof(int));
  cudaMemset(h->plane_max, 0, 4 * sizeof(unsigned int));
  *out = h;
  return 0;
}

int dewi_index_destroy(dewi_index_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->rows_f32);
  cudaFree(h->plane0);
  cudaFree(h->plane1);
  cudaFree(h->dewi_col);
  cudaFree(h->ent_col);
  cudaFree(h->bad_flag);
  cudaFree(h->plane_max);
  if (h->pin_q) cudaFreeHost(h->pin_q);
  if (h->pin_out) cudaFreeHost(h->pin_out);
  for (int i = 0; i < dewi_index::kEvRing; ++i) {
    if (h->ev0[i]) cudaEventDestroy(h->ev0[i]);
    if (h->ev1[i]) cudaEventDestroy(h->ev1[i]);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < 2; ++i) {
    if (h->stage_free[i]) cudaEventDestroy(h->stage_free[i]);
    if (h->stage_full[i]) cudaEventDestroy(h->stage_full[i]);
  }
  for (DevBuf* b : {&h->stage, &h->stage2, &h->qraw, &h->qn, &h->q0, &h->q1, &h->qstats, &h->qbar, &h->part_s, &h->part_i, &h->seed_max, &h->seed_sim, &h->cand_idx, &h->cand_sim,
                    &h->loc_sim, &h->loc_id, &h->loc_dewi, &h->loc_ent, &h->out_id, &h->out_score, &h->push_ticket, &h->sync_cnt})
    b->release();
  delete h;
  return 0;
}

int dewi_index_reserve(dewi_index_t* h, int64_t rows) {
  if (!h) return fail("null handle");
  DEWI_TRY(set_device(h));
  return grow(h, rows, nullptr);
}

int dewi_index_set_id_base(dewi_index_t* h, int64_t id_base) {
  if (!h) return fail("null handle");
  h->id_base = id_base;
  return 0;
}

int dewi_index_append(dewi_index_t* h, const float* rows, int64_t n, int normalized, int src_is_host, void* stream_) {
  if (!h) return fail("null handle");
  if (n < 0) return fail("negative row count");
  if (n == 0) return 0;
  if (!rows) return fail("rows is NULL");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEWI_TRY(set_device(h));
  if (h->n + n >= (int64_t(1) << 31)) return fail("a shard holds at most 2^31-1 rows");
  DEWI_TRY(grow(h, h->n + n, stream));
  const int do_norm = (h->space == DEWI_SPACE_COSINE && !normalized) ? 1 : 0;
  const size_t d = static_cast<size_t>(h->dim);
  // Host rows go through two 64 MB staging buffers: the copy of chunk i+1 (on the handle's copy stream) overlaps
  // the normalise / plane-split kernel of chunk i (on the caller's stream); events order the two per buffer.
  const int64_t chunk = src_is_host ? std::max<int64_t>(1, (int64_t(64) << 20) / static_cast<int64_t>(d * 4)) : n;
  if (src_is_host && !h->copy_stream) {
    DEWI_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      DEWI_CUDA(cudaEventCreateWithFlags(&h->stage_free[i], cudaEventDisableTiming));
      DEWI_CUDA(cudaEventCreateWithFlags(&h->stage_full[i], cudaEventDisableTiming));
    }
  }
  int buf = 0;
  int used[2] = {0, 0};
  for (int64_t done = 0; done < n; done += chunk, buf ^= 1) {
    const int64_t m = std::min(chunk, n - done);
    const float* src = rows + static_cast<size_t>(done) * d;
    if (src_is_host) {
      DevBuf& st = buf ? h->stage2 : h->stage;
      if (used[buf]) DEWI_CUDA(cudaStreamWaitEvent(h->copy_stream, h->stage_free[buf], 0));  // kernel of chunk i-2 is done
      else DEWI_TRY(st.ensure(static_cast<size_t>(std::min(chunk, n)) * d * 4));
      DEWI_CUDA(cudaMemcpyAsync(st.p, src, static_cast<size_t>(m) * d * 4, cudaMemcpyHostToDevice, h->copy_stream));
      DEWI_CUDA(cudaEventRecord(h->stage_full[buf], h->copy_stream));
      DEWI_CUDA(cudaStreamWaitEvent(stream, h->stage_full[buf], 0));
      src = st.as<float>();
    }
    const size_t off = static_cast<size_t>(h->n + done) * d;
    DEWI_TRY(launch_prep_corpus(src, m, h->dim, do_norm, h->rows_f32 ? h->rows_f32 + off : nullptr,
                                h->plane0 ? h->plane0 + off : nullptr, h->plane1 ? h->plane1 + off : nullptr,
                                h->bad_flag, stream, h->dtype == DEWI_DTYPE_FP32 ? h->plane_max : nullptr,
                                h->dtype == DEWI_DTYPE_FP32 ? 1 : 0));
    if (h->dtype == DEWI_DTYPE_FP32) h->plane_max_stale = true;
    if (src_is_host) {
      DEWI_CUDA(cudaEventRecord(h->stage_free[buf], stream));
      used[buf] = 1;
    }
  }
  if (src_is_host) DEWI_CUDA(cudaStreamSynchronize(stream));  // the caller's host rows may be released on return
  if (do_norm) {
    int bad = 0;
    DEWI_CUDA(cudaMemcpyAsync(&bad, h->bad_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    DEWI_CUDA(cudaStreamSynchronize(stream));
    if (bad) {
      DEWI_CUDA(cudaMemsetAsync(h->bad_flag, 0, sizeof(int), stream));
      return fail("zero-norm embedding in cosine space (the reference would store a NaN row, backends.py:405)");
    }
  }
  h->n += n;
  h->map_rows = -1;
  return 0;
}

int dewi_index_set_payload(dewi_index_t* h, const float* dewi_v, const float* ent_v, int64_t offset, int64_t n,
                           int src_is_host, void* stream_) {
  if (!h) return fail("null handle");
  if (offset < 0 || n < 0 || offset + n > h->n) return fail("payload range outside the corpus");
  if (n == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEWI_TRY(set_device(h));
  const cudaMemcpyKind kind = src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  DEWI_CUDA(cudaMemcpyAsync(h->dewi_col + offset, dewi_v, static_cast<size_t>(n) * 4, kind, stream));
  DEWI_CUDA(cudaMemcpyAsync(h->ent_col + offset, ent_v, static_cast<size_t>(n) * 4, kind, stream));
  if (src_is_host) DEWI_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

int dewi_index_size(const dewi_index_t* h, int64_t* rows) {
  if (!h || !rows) return fail("null argument");
  *rows = h->n;
  return 0;
}

int dewi_index_get_row(dewi_index_t* h, int64_t row, float* out_host) {
  if (!h || !out_host) return fail("null argument");
  if (row < 0 || row >= h->n) return fail("row out of range");
  DEWI_TRY(set_device(h));
  const size_t d = static_cast<size_t>(h->dim);
  if (h->rows_f32) {
    DEWI_CUDA(cudaMemcpy(out_host, h->rows_f32 + static_cast<size_t>(row) * d, d * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<uint16_t> tmp(d);
    DEWI_CUDA(cudaMemcpy(tmp.data(), h->plane0 + static_cast<size_t>(row) * d, d * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < d; ++i) {
      const uint32_t u = static_cast<uint32_t>(tmp[i]) << 16;
      std::memcpy(&out_host[i], &u, 4);
    }
  }
  return 0;
}

int dewi_index_export_rows(dewi_index_t* h, int64_t row0, int64_t n, float* out, int dst_is_host, void* stream_) {
  if (!h || !out) return fail("null argument");
  if (row0 < 0 || n < 0 || row0 + n > h->n) return fail("row range outside the corpus");
  if (n == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEWI_TRY(set_device(h));
  const size_t d = static_cast<size_t>(h->dim);
  const cudaMemcpyKind kind = dst_is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (h->rows_f32) {  // fp32 mode keeps the exact rows: one copy (the driver stages a pageable destination itself)
    DEWI_CUDA(cudaMemcpyAsync(out, h->rows_f32 + static_cast<size_t>(row0) * d, static_cast<size_t>(n) * d * 4, kind, stream));
    if (dst_is_host) DEWI_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  // bf16 mode: widen on the device (exact), in >= 64 MB chunks through the staging buffer when the destination is host memory
  const int64_t chunk = dst_is_host ? std::max<int64_t>(1, (int64_t(128) << 20) / static_cast<int64_t>(d * 4)) : n;
  for (int64_t done = 0; done < n; done += chunk) {
    const int64_t m = std::min(chunk, n - done);
    const __nv_bfloat16* src = h->plane0 + static_cast<size_t>(row0 + done) * d;
    float* dst = out + static_cast<size_t>(done) * d;
    if (dst_is_host) {
      DEWI_TRY(h->stage.ensure(static_cast<size_t>(std::min(chunk, n)) * d * 4));
      DEWI_TRY(launch_widen_bf16(src, static_cast<int64_t>(m) * h->dim, h->stage.as<float>(), stream));
      DEWI_CUDA(cudaMemcpyAsync(dst, h->stage.p, static_cast<size_t>(m) * d * 4, cudaMemcpyDeviceToHost, stream));
      DEWI_CUDA(cudaStreamSynchronize(stream));  // staging buffer is reused
    } else {
      DEWI_TRY(launch_widen_bf16(src, static_cast<int64_t>(m) * h->dim, dst, stream));
    }
  }
  return 0;
}

int dewi_index_export_bf16(dewi_index_t* h, int64_t row0, int64_t n, uint16_t* out, int dst_is_host, void* stream_) {
  if (!h || !out) return fail("null argument");
  if (row0 < 0 || n < 0 || row0 + n > h->n) return fail("row range outside the corpus");
  if (h->dtype != DEWI_DTYPE_BF16 || !h->plane0) return fail("export_bf16 needs a bf16-storage index");
  if (n == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEWI_TRY(set_device(h));
  const size_t d = static_cast<size_t>(h->dim);
  DEWI_CUDA(cudaMemcpyAsync(out, h->plane0 + static_cast<size_t>(row0) * d, static_cast<size_t>(n) * d * 2,
                            dst_is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, stream));
  if (dst_is_host) DEWI_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

int dewi_index_append_bf16(dewi_index_t* h, const uint16_t* rows, int64_t n, int src_is_host, void* stream_) {
  if (!h) return fail("null handle");
  if (n < 0) return fail("negative row count");
  if (n == 0) return 0;
  if (!rows) return fail("rows is NULL");
  if (h->dtype != DEWI_DTYPE_BF16) return fail("append_bf16 needs a bf16-storage index");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEWI_TRY(set_device(h));
  if (h->n + n >= (int64_t(1) << 31)) return fail("a shard holds at most 2^31-1 rows");
  DEWI_TRY(grow(h, h->n + n, stream));
  const size_t d = static_cast<size_t>(h->dim);
  DEWI_CUDA(cudaMemcpyAsync(h->plane0 + static_cast<size_t>(h->n) * d, rows, static_cast<size_t>(n) * d * 2,
                            src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, stream));
  if (src_is_host) DEWI_CUDA(cudaStreamSynchronize(stream));
  h->n += n;
  h->map_rows = -1;
  return 0;
}

int dewi_index_get_payload(dewi_index_t* h, int64_t offset, int64_t n, float* dewi_out, float* ent_out, int dst_is_host,
                           void* stream_) {
  if (!h || !dewi_out || !ent_out) return fail("null argument");
  if (offset < 0 || n < 0 || offset + n > h->n) return fail("payload range outside the corpus");
  if (n == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEWI_TRY(set_device(h));
  const cudaMemcpyKind kind = dst_is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  DEWI_CUDA(cudaMemcpyAsync(dewi_out, h->dewi_col + offset, static_cast<size_t>(n) * 4, kind, stream));
  DEWI_CUDA(cudaMemcpyAsync(ent_out, h->ent_col + offset, static_cast<size_t>(n) * 4, kind, stream));
  if (dst_is_host) DEWI_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

// `rerank` (single shard only): when the fused tail runs, the DEWI blend and the final top-k happen in the same launch
// and *rerank_done is set; otherwise the caller re-ranks the local outputs itself.
static int search_local_impl(dewi_index_t* h, const float* queries, int B, int kcand, int flags, float* out_sim,
                             int64_t* out_id, float* out_dewi, float* out_ent, void* stream_, const PeerPush* push,
                             const TailRerank* rerank = nullptr, bool* rerank_done = nullptr);

int dewi_index_search_local(dewi_index_t* h, const float* queries, int B, int kcand, int flags, float* out_sim,
                            int64_t* out_id, float* out_dewi, float* out_ent, void* stream_) {
  if (!out_sim || !out_id || !out_dewi || !out_ent) return fail("null output");
  return search_local_impl(h, queries, B, kcand, flags, out_sim, out_id, out_dewi, out_ent, stream_, nullptr);
}

int dewi_index_search_local_push(dewi_index_t* h, const float* queries, int B, int kcand, int flags, int world, int my_rank,
                                 const uint64_t* peer_bases, const uint64_t* peer_flags, int64_t block_stride_bytes,
                                 uint32_t seq, void* stream_) {
  if (!h) return fail("null handle");
  if (!peer_bases || !peer_flags) return fail("null peer table");
  if (world < 1 || world > kMaxPeers || my_rank < 0 || my_rank >= world) return fail("invalid world / rank");
  if (block_stride_bytes < static_cast<int64_t>(B) * kcand * 20 || block_stride_bytes % 8 != 0)
    return fail("block stride too small for [id i64 | sim | dewi | ent] x B x kcand, or not a multiple of 8");
  DEWI_TRY(set_device(h));
  DEWI_TRY(h->push_ticket.ensure(4));
  if (!h->push_ticket_zeroed) {
    DEWI_CUDA(cudaMemsetAsync(h->push_ticket.p, 0, 4, static_cast<cudaStream_t>(stream_)));
    h->push_ticket_zeroed = true;
  }
  PeerPush push;
  push.world = world;
  push.my_rank = my_rank;
  push.seq = seq;
  push.block_stride = block_stride_bytes;
  for (int r = 0; r < world; ++r) {
    if (!peer_bases[r] || !peer_flags[r]) return fail("null peer pointer");
    push.base[r] = peer_bases[r];
    push.flags[r] = peer_flags[r];
  }
  push.ticket = h->push_ticket.as<unsigned int>();
  return search_local_impl(h, queries, B, kcand, flags, nullptr, nullptr, nullptr, nullptr, stream_, &push);
}

static int search_local_impl(dewi_index_t* h, const float* queries, int B, int kcand, int flags, float* out_sim,
                             int64_t* out_id, float* out_dewi, float* out_ent, void* stream_, const PeerPush* push,
                             const TailRerank* rerank, bool* rerank_done) {
  if (!h) return fail("null handle");
  if (rerank_done) *rerank_done = false;
  if (B <= 0 || kcand <= 0) return fail("B and kcand must be positive");
  if (h->n <= 0) return fail("index is empty");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEWI_TRY(set_device(h));
  h->last_launches = 0;
  const int dim = h->dim;
  // rerank_scope = "full" (opt-in, not the reference's semantics): the sweep selects by the blended key over the whole
  // corpus -- rows-on-M tensor-core sweep (bf16 corpus; the key is evaluated per row in its epilogue) or the exact
  // CUDA-core sweep (fp32 corpus, l2 space, odd dims) -- and the kcand best by the exactly re-evaluated blend come back.
  const bool full = (flags & DEWI_FLAG_SCOPE_FULL) != 0;
  SweepBlend blend;
  if (full) {
    if (!h->blend_set) return fail("full-corpus blend: call dewi_index_set_blend (or dewi_index_search) first");
    if (push && push->world > 0) return fail("full-corpus blend: use the all-gather exchange, not the fused peer push");
    if (!h->dewi_col || !h->ent_col) return fail("full-corpus blend: payload columns are not set");
    blend = h->blend;
    blend.enabled = 1;
    blend.dewi = h->dewi_col;
    blend.ent = h->ent_col;
  }
  int b_pad = static_cast<int>(round_up(B, kQueryBlock));
  int n_qb = b_pad / kQueryBlock;
  const int kc_valid = static_cast<int>(std::min<int64_t>(kcand, h->n));

  bool use_tc = h->space == DEWI_SPACE_COSINE && h->plane0 && tc_supported(dim, h->n) && !(flags & DEWI_FLAG_FORCE_SIMT) &&
                (h->n >= kTcMinRows || (flags & DEWI_FLAG_FORCE_TC));
  const int fp16_planes = h->dtype == DEWI_DTYPE_FP32 ? 1 : 0;
  if (use_tc && fp16_planes) {
    // The planes of an fp32 corpus are fp16: rows handed in as "already normalised" that are far from unit norm
    // could leave its range (products would overflow) -- such a corpus is swept on the CUDA cores instead.
    if (h->plane_max_stale) {
      unsigned int bits = 0;
      DEWI_CUDA(cudaMemcpyAsync(&bits, h->plane_max + 1, sizeof(bits), cudaMemcpyDeviceToHost, stream));
      DEWI_CUDA(cudaStreamSynchronize(stream));
      std::memcpy(&h->plane_hi_max, &bits, 4);
      h->plane_max_stale = false;
    }
    if (!(h->plane_hi_max <= 64.f)) {
      if (flags & DEWI_FLAG_FORCE_TC) return fail("rows too far from unit norm for the fp16 planes of the tensor-core sweep");
      use_tc = false;
    }
  }
  int mode = (h->dtype == DEWI_DTYPE_FP32) ? 2 : ((flags & DEWI_FLAG_PRECISE_QUERY) ? 1 : 0);
  if (full && use_tc) {
    // only the rows-on-M sweep evaluates the key (one plane, at most 64 queries per launch).  An fp32 corpus goes to the
    // exact CUDA-core sweep unless the tensor cores are forced (then: fp16 hi plane, over-fetch, exact re-score).
    if (mode == 2 && (flags & DEWI_FLAG_FORCE_TC)) mode = 0;
    if (mode != 0 || (flags & (DEWI_FLAG_NO_M64 | DEWI_FLAG_NO_ROWS_ON_M))) {
      if (flags & DEWI_FLAG_FORCE_TC) return fail("full-corpus blend: the tensor-core path needs the rows-on-M sweep (one plane)");
      use_tc = false;
    } else if (B > 64) {
      for (int b0 = 0; b0 < B; b0 += 64) {
        const int nb = std::min(64, B - b0);
        DEWI_TRY(search_local_impl(h, queries + static_cast<size_t>(b0) * dim, nb, kcand, flags, out_sim + static_cast<size_t>(b0) * kcand,
                                   out_id + static_cast<size_t>(b0) * kcand, out_dewi + static_cast<size_t>(b0) * kcand,
                                   out_ent + static_cast<size_t>(b0) * kcand, stream_, nullptr, nullptr, nullptr));
      }
      return 0;
    }
  }
  // list capacity: over-fetch so that rounding in the bf16-plane sweep cannot push a true top-2k row out
  int kc = (mode == 0) ? std::max(32, kc_valid + 16) : kc_valid + 8;
  TcPlan plan;
  Tc2Plan plan2;
  // at most 64 queries: M = 64 MMAs (half the tensor work / power and half the query re-stream)
  const int q_rows = (B <= 64 && !(flags & DEWI_FLAG_NO_M64)) ? 64 : kQueryBlock;
  // ... and, where it applies, with the corpus rows on M and the queries on N = 16 / 32 / 64 (search_tcr.cu)
  const int swap_b = (q_rows == 64 && !(flags & DEWI_FLAG_NO_ROWS_ON_M)) ? B : 0;
  plan.rows_on_m = 0;
  const int n_qb_in = n_qb;
  // more than one query block: the CTA-pair sweep (two query blocks share every corpus tile)
  bool use_pair = false;
  // rows the seeding pre-pass of the CTA-pair sweep samples, in tiles of 256 (see the seeding block below)
  auto pair_sample_tiles = [&](int n_qb_) -> int64_t {
    const int64_t tiles_total = ceil_div(h->n, 2 * tc2_box_rows());
    int64_t t = std::max<int64_t>(std::max(1, h->sm_count / 2), tiles_total / 128);
    if (n_qb_ >= 8) t = std::max<int64_t>(t, std::min<int64_t>(128, tiles_total / 16));
    return t;
  };
  auto make_plans = [&](int mode_, int kc_) -> bool {   // false: no tensor-core plan for this (mode, list capacity)
    n_qb = n_qb_in;
    b_pad = n_qb * kQueryBlock;
    use_pair = n_qb >= 2 && !(flags & DEWI_FLAG_NO_PAIR);
    if (use_pair) {
      const int n_qb2 = static_cast<int>(round_up(n_qb, 2));
      // Many query pairs over a small corpus: the pre-pass sample is too small for a tight seed (see the seeding block
      // below: ~1024 kc / sample_rows hits per warp and 32-column group), so the sweep is STAGED -- the first round of
      // chunks runs as a launch of its own and its finished lists seed the rest (search_tc2.cu: tc2_make_plan).
      const int64_t sample_rows = pair_sample_tiles(n_qb2) * 2 * tc2_box_rows();
      const int staged = (n_qb2 >= 8 && !(flags & DEWI_FLAG_NO_SEED) && sample_rows * 3 < static_cast<int64_t>(4096) * kc_ &&
                          env_int("DEWI_TC2_STAGED", 1) != 0) ? 1 : 0;
      if (tc2_make_plan(mode_, dim, h->n, n_qb2, kc_, h->sm_count, &plan2, 0, 0, 0, staged) == 0) {
        n_qb = n_qb2;
        b_pad = n_qb * kQueryBlock;
        return true;
      }
      use_pair = false;
    }
    return tc_make_plan(mode_, dim, h->n, n_qb, kc_, h->sm_count, &plan, 0, q_rows, swap_b) == 0;
  };
  // CERTIFIED single-plane sweep (fp32 corpus).  The fp16 hi plane alone -- half the bytes of the hi/lo stream below
  // the ridge, a third of its MMAs above -- is swept with a longer candidate list; a certificate (select.cu) then
  // proves, from rigorous bounds on the two fp16 roundings and the fp32 accumulation, that the list holds the exact
  // top-2k, and the exact fp32 re-score orders it.  When the certificate cannot be given for some query (scores packed
  // more densely than the bound resolves) the batch is re-run with the full hi/lo product: exact either way.
  int cert_planes = 0;
  if (use_tc && !full && mode == 2 && !(flags & DEWI_FLAG_NO_CERT) && env_int("DEWI_CERT", 1) != 0 &&
      (static_cast<int64_t>(h->n) * dim >= kCertMinElems || (flags & DEWI_FLAG_FORCE_CERT))) {
    const int planes = 1;
    // List capacity.  With fp16 planes the bound is eps ~ 6e-4 at dim 768 (2.1e-4 per rounded operand, 1.8e-4 for the
    // accumulation): ~23 rows are expected inside the proof's margin at k = 10 on 1M Gaussian rows, with a heavier
    // than Poisson tail across queries (the 2k-th best score itself fluctuates); 48 slots keep the re-run rare and
    // are what fits beside five ring stages (CTA pairs) / the resident query block (B <= 64).  (bf16 planes would need
    // ~49 expected, 82 seen in 1024 queries -- and every 16 more slots cost the CTA-pair sweep ~9 % at B = 4096.)
    int cert_kc = std::max(48, kc_valid * 2 + 8);
    if (env_set("DEWI_CERT_KC")) cert_kc = std::max(kc_valid + 1, env_int("DEWI_CERT_KC", cert_kc));   // experiments
    if (make_plans(planes == 1 ? 0 : 1, cert_kc) && (use_pair || plan.n_stages >= 3)) {
      cert_planes = planes;
      mode = planes == 1 ? 0 : 1;
      kc = cert_kc;
    }
  }
  if (use_tc && !cert_planes) {
    if (!make_plans(mode, kc)) {
      if (flags & DEWI_FLAG_FORCE_TC) return 1;
      use_tc = false;
      use_pair = false;
      n_qb = n_qb_in;
      b_pad = n_qb * kQueryBlock;
    }
  } else if (!use_tc && (flags & DEWI_FLAG_FORCE_TC)) {
    return fail("tcgen05 sweep not applicable (needs cosine space and dim % 64 == 0)");
  }
  if (full && use_tc && !plan.rows_on_m) {   // (query block too large to stay resident, very long lists)
    if (flags & DEWI_FLAG_FORCE_TC) return fail("full-corpus blend: no rows-on-M plan for this shape");
    use_tc = false;
    use_pair = false;
    n_qb = n_qb_in;
    b_pad = n_qb * kQueryBlock;
  }

  // queries -> normalised fp32 + bf16 planes
  DEWI_TRY(h->qn.ensure(static_cast<size_t>(b_pad) * dim * 4));
  DEWI_TRY(h->q0.ensure(static_cast<size_t>(b_pad) * dim * 2));
  DEWI_TRY(h->q1.ensure(static_cast<size_t>(b_pad) * dim * 2));
  const int qnorm = (h->space == DEWI_SPACE_COSINE && !(flags & DEWI_FLAG_QUERY_NORMALIZED)) ? 1 : 0;
  if (cert_planes) DEWI_TRY(h->qstats.ensure(static_cast<size_t>(b_pad) * 16));
  DEWI_TRY(launch_prep_queries(queries, B, b_pad, dim, qnorm, h->qn.as<float>(), h->q0.as<__nv_bfloat16>(),
                               h->q1.as<__nv_bfloat16>(), stream,
                               /*lane_order=*/(use_tc && !use_pair && plan.rows_on_m) ? 0 : ((use_tc && !use_pair && q_rows == 64) ? 2 : 1),
                               cert_planes ? h->qstats.as<float>() : nullptr, fp16_planes));
  h->last_launches++;

  const void* exact_rows = h->rows_f32 ? static_cast<const void*>(h->rows_f32) : static_cast<const void*>(h->plane0);
  const int exact_is_bf16 = h->rows_f32 ? 0 : 1;
  Partials parts;
  const int ev_slot = static_cast<int>(h->searches % dewi_index::kEvRing);
  if (h->profile) DEWI_CUDA(cudaEventRecord(h->ev0[ev_slot], stream));
  h->last_sweep_kind = use_tc ? (use_pair ? 3 : (plan.rows_on_m ? 4 : 1)) : 2;
  if (use_tc) {
    DEWI_TRY(ensure_corpus_maps(h, use_pair ? tc2_box_rows() : plan.n_tile));
    CUtensorMap mq0, mq1;
    const int q_box = use_pair ? kQueryBlock : plan.q_rows;
    DEWI_TRY(tc_encode_rows_map(&mq0, h->q0.p, b_pad, dim, q_box));
    DEWI_TRY(tc_encode_rows_map(&mq1, h->q1.p, b_pad, dim, q_box));
    {  // size the partial-list buffers for the main sweep up front: growing them later would cudaFree
       // (a device-wide synchronisation) between the pre-pass and the main sweep
      const size_t main_items = static_cast<size_t>(use_pair ? plan2.n_chunks : plan.n_chunks) * n_qb;
      DEWI_TRY(h->part_s.ensure(main_items * kc * kQueryBlock * 4));
      DEWI_TRY(h->part_i.ensure(main_items * kc * kQueryBlock * 4));
    }
    auto sweep = [&](const TcPlan& p1, const Tc2Plan& p2, int64_t rows, const SweepSeed& sd) -> int {
      const int chunks = use_pair ? p2.n_chunks : p1.n_chunks;
      const size_t items = static_cast<size_t>(chunks) * n_qb;
      DEWI_TRY(h->part_s.ensure(items * kc * kQueryBlock * 4));
      DEWI_TRY(h->part_i.ensure(items * kc * kQueryBlock * 4));
      if (use_pair) {
        const int64_t sync_words = tc2_sync_words(p2, rows, n_qb);
        if (sync_words > 0) DEWI_TRY(h->sync_cnt.ensure(static_cast<size_t>(sync_words) * 4));
        unsigned int* sc = sync_words > 0 ? h->sync_cnt.as<unsigned int>() : nullptr;
        if (p2.first_items > 0 && !sd.max_out) {
          // staged: the first round of chunks, then the kc-th best of every query's finished lists as the seed of the rest
          DEWI_TRY(tc2_launch(p2, h->map_e0, h->map_e1, mq0, mq1, rows, dim, n_qb, kc, h->part_s.as<float>(), h->part_i.as<int>(), sd,
                              stream, sc, fp16_planes, 0, p2.first_items));
          DEWI_TRY(h->seed_sim.ensure(static_cast<size_t>(B) * 4));
          DEWI_TRY(launch_seed_from_partials(h->part_s.as<float>(), p2.first_items / (n_qb / 2), n_qb, B, kc, sd.values,
                                             h->seed_sim.as<float>(), stream));
          SweepSeed sd2;
          sd2.values = h->seed_sim.as<float>();
          sd2.stride = 1;
          sd2.off = 0;
          sd2.n_queries = B;
          DEWI_TRY(tc2_launch(p2, h->map_e0, h->map_e1, mq0, mq1, rows, dim, n_qb, kc, h->part_s.as<float>(), h->part_i.as<int>(), sd2,
                              stream, sc, fp16_planes, p2.first_items, -1));
          h->last_launches += 2;
        } else {
          DEWI_TRY(tc2_launch(p2, h->map_e0, h->map_e1, mq0, mq1, rows, dim, n_qb, kc, h->part_s.as<float>(),
                              h->part_i.as<int>(), sd, stream, sc, fp16_planes));
        }
      }
      else
        DEWI_TRY(tc_launch(p1, h->map_e0, h->map_e1, mq0, mq1, rows, dim, n_qb, kc, h->part_s.as<float>(),
                           h->part_i.as<int>(), sd, stream, fp16_planes, B, full ? &blend : nullptr));
      h->last_launches++;
      return 0;
    };
    // Threshold seeding: sweep a sample (the first 1/128 of the rows, at least one tile per CTA) first,
    // recording only each work item's best score per query; the kc-th largest of those maxima (they
    // belong to distinct rows) is a lower bound of the query's kc-th best score overall, so the main
    // sweep starts with that admission threshold instead of -inf and its candidate lists see ~kc
    // entries per CTA instead of ~kc * ln(rows per CTA), with no expensive list warm-up anywhere.
    // Exactness is unaffected (sweep_epilogue.cuh: seed_threshold).
    SweepSeed seed;
    {
      const int n_tile = use_pair ? 2 * tc2_box_rows() : plan.n_tile;
      const int64_t tiles_total = ceil_div(h->n, n_tile);
      const int64_t workers = use_pair ? std::max(1, h->sm_count / 2) : h->sm_count;
      if (!(flags & DEWI_FLAG_NO_SEED) && tiles_total >= 16 * workers) {
        int64_t sample_tiles = use_pair ? pair_sample_tiles(n_qb) : std::max<int64_t>(workers, tiles_total / 128);
        // (many query blocks over a SMALL corpus: every (chunk, query pair) work item restarts its lists at the seed, and a
        // seed drawn from 2 % of 1M rows leaves ~70 rows per query above it in every item -- the epilogue then walks its
        // slow path in 92 % of the 32-column groups and sets the pace (ncu: tensor pipe 43 % active at 1M rows, B = 4096).
        // pair_sample_tiles takes 128 tiles (3 % of 1M rows) there: B = 4096 sweep 6.91 -> 6.22 ms, B = 1024 1.77 -> 1.65 ms;
        // 256 tiles cost more than they return (7.70 ms).  The staged sweep then tightens the seed further.)
        if (env_set("DEWI_SEED_TILES")) sample_tiles = std::min<int64_t>(tiles_total / 2, std::max<int64_t>(workers, env_int("DEWI_SEED_TILES", 0)));  // experiments
        const int64_t sample_rows = sample_tiles * n_tile;
        TcPlan s1{};
        Tc2Plan s2{};
        // Granularity of the recorded maxima.  A small sample (a small corpus, or a long candidate list next to it)
        // records one maximum per 32-row group of every tile -- up to 2048 of them, all of distinct rows; a large one
        // is cut into ~4 kc chunks per query block and records one maximum per chunk.  Either way the kc-th largest
        // maximum is a lower bound of the query's kc-th best score.
        const int groups = n_tile / 32;
        const bool by_group = sample_tiles * groups <= 2048;
        const int want = by_group ? 0 : static_cast<int>(std::min<int64_t>(std::max<int64_t>(workers, 4 * kc), 2048));
        const int rc = use_pair ? tc2_make_plan(mode, dim, sample_rows, n_qb, kc, h->sm_count, &s2, want)
                                : tc_make_plan(mode, dim, sample_rows, n_qb, kc, h->sm_count, &s1, want, q_rows,
                                               plan.rows_on_m ? B : 0);
        const int s_chunks = by_group ? static_cast<int>(sample_tiles * groups) : (use_pair ? s2.n_chunks : s1.n_chunks);
        if (rc == 0 && s_chunks >= kc && s_chunks <= 2048) {
          DEWI_TRY(h->seed_max.ensure(static_cast<size_t>(s_chunks) * n_qb * kQueryBlock * 4));
          DEWI_TRY(h->seed_sim.ensure(static_cast<size_t>(B) * 4));
          SweepSeed pre;
          pre.max_out = h->seed_max.as<float>();
          pre.max_groups = by_group ? 1 : 0;
          DEWI_TRY(sweep(s1, s2, sample_rows, pre));
          DEWI_TRY(launch_seed_from_maxima(h->seed_max.as<float>(), s_chunks, n_qb, B, kc, h->seed_sim.as<float>(), stream));
          h->last_launches++;
          seed.values = h->seed_sim.as<float>();
          seed.stride = 1;
          seed.off = 0;
          seed.n_queries = B;
        }
      }
    }
    const int n_chunks = use_pair ? plan2.n_chunks : plan.n_chunks;
    DEWI_TRY(sweep(plan, plan2, h->n, seed));
    h->last_launches--;  // counted once more just below
    h->last_launches++;
    parts.n_chunks = n_chunks;
  } else {
    // (full scope: the sweep's fused-multiply-add key and the re-rank's separately rounded blend can order near-ties
    // differently; a few extra slots make the kcand best by the exact blend certain to be among the candidates)
    kc = full ? static_cast<int>(std::min<int64_t>(static_cast<int64_t>(kc_valid) + 8, h->n)) : kc_valid;
    int n_chunks = 1;
    DEWI_TRY(simt_plan(h->n, B, h->sm_count, &n_chunks));
    const size_t items = static_cast<size_t>(n_chunks) * n_qb;
    DEWI_TRY(h->part_s.ensure(items * kc * kQueryBlock * 4));
    DEWI_TRY(h->part_i.ensure(items * kc * kQueryBlock * 4));
    DEWI_TRY(simt_launch(exact_rows, exact_is_bf16, h->n, dim, h->space, h->qn.as<float>(), B, kc, n_chunks,
                         h->part_s.as<float>(), h->part_i.as<int>(), stream, full ? &blend : nullptr));
    h->last_launches++;
    parts.n_chunks = n_chunks;
  }
  if (h->profile) {
    DEWI_CUDA(cudaEventRecord(h->ev1[ev_slot], stream));
    h->searches++;
  }
  parts.s = h->part_s.as<float>();
  parts.i = h->part_i.as<int>();
  parts.n_qb = n_qb;
  parts.kc = kc;

  int* fails_dev = reinterpret_cast<int*>(h->plane_max + 2);
  if (!full && tail_supported(kc, kcand, rerank ? rerank->k : 1) && env_int("DEWI_FUSED_TAIL", 1) != 0) {
    // ONE launch for everything after the sweep (select.cu: tail_kernel)
    TailCert cert{h->qstats.as<float>(), cert_planes, h->plane_max, kc_valid, fails_dev};
    if (cert_planes) DEWI_CUDA(cudaMemsetAsync(fails_dev, 0, sizeof(int), stream));
    DEWI_TRY(launch_tail(parts, B, cert_planes ? &cert : nullptr, use_tc ? exact_rows : nullptr, exact_is_bf16, dim, h->qn.as<float>(),
                         kcand, h->id_base, h->dewi_col, h->ent_col, out_sim, out_id, out_dewi, out_ent, push, rerank, stream));
    h->last_launches++;
    if (cert_planes) {
      int fails = 0;
      DEWI_CUDA(cudaMemcpyAsync(&fails, fails_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
      DEWI_CUDA(cudaStreamSynchronize(stream));
      h->cert_used++;
      if (fails > 0 && env_int("DEWI_CERT_IGNORE", 0)) fails = 0;   // experiments (timing only: results may be inexact)
      if (fails > 0) {   // not provable for `fails` queries: the whole batch again with the full hi/lo product
        h->cert_failed++;  // (a peer push was withheld by the kernel: the re-run publishes)
        return search_local_impl(h, queries, B, kcand, flags | DEWI_FLAG_NO_CERT, out_sim, out_id, out_dewi, out_ent, stream_, push,
                                 rerank, rerank_done);
      }
    }
    if (rerank && rerank_done) *rerank_done = true;
    return 0;
  }
  DEWI_TRY(h->cand_idx.ensure(static_cast<size_t>(B) * kc * 4));
  DEWI_TRY(h->cand_sim.ensure(static_cast<size_t>(B) * kc * 4));
  DEWI_TRY(launch_merge_select(parts, B, kc, h->cand_idx.as<int>(), h->cand_sim.as<float>(), stream));
  h->last_launches++;
  if (cert_planes) {
    DEWI_CUDA(cudaMemsetAsync(fails_dev, 0, sizeof(int), stream));
    DEWI_TRY(h->qbar.ensure(static_cast<size_t>(B) * 4));
    DEWI_TRY(launch_certificate(h->cand_sim.as<float>(), h->cand_idx.as<int>(), B, kc, kc_valid, h->qstats.as<float>(), cert_planes,
                                h->plane_max, dim, fails_dev, h->qbar.as<float>(), stream));
    h->last_launches++;
    int fails = 0;
    DEWI_CUDA(cudaMemcpyAsync(&fails, fails_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
    DEWI_CUDA(cudaStreamSynchronize(stream));
    h->cert_used++;
    if (fails > 0 && env_int("DEWI_CERT_IGNORE", 0)) fails = 0;   // experiments (timing only: results may be inexact)
    if (fails > 0) {   // not provable for `fails` queries: the whole batch again with the full hi/lo product
      h->cert_failed++;
      return search_local_impl(h, queries, B, kcand, flags | DEWI_FLAG_NO_CERT, out_sim, out_id, out_dewi, out_ent, stream_, push,
                               rerank, rerank_done);
    }
  }
  if (use_tc) {
    DEWI_TRY(launch_rescore(exact_rows, exact_is_bf16, dim, h->qn.as<float>(), h->cand_idx.as<int>(), B, kc,
                            h->cand_sim.as<float>(), stream, cert_planes ? h->qbar.as<float>() : nullptr));
    h->last_launches++;
  } else if (full) {   // the CUDA-core sweep's lists hold blended keys: back to the exact scores
    DEWI_TRY(launch_rescore_any(exact_rows, exact_is_bf16, dim, h->space == DEWI_SPACE_L2 ? 1 : 0, h->qn.as<float>(),
                                h->cand_idx.as<int>(), B, kc, h->cand_sim.as<float>(), stream));
    h->last_launches++;
  }
  DEWI_TRY(launch_finalize_local(h->cand_idx.as<int>(), h->cand_sim.as<float>(), B, kc, kcand, h->id_base, h->dewi_col,
                                 h->ent_col, out_sim, out_id, out_dewi, out_ent, stream, push, full ? &blend : nullptr));
  h->last_launches++;
  return 0;
}

int dewi_rerank_gathered(const float* sim, const int64_t* id, const float* dewi_v, const float* ent_v, int B, int n_shards,
                         int kcand, int64_t shard_stride_bytes, int cand_count, int k, double eta, double entropy_pref,
                         int64_t* out_id, float* out_score, const uint32_t* ready_flags, uint32_t seq, uint32_t* status_word,
                         double timeout_s, int device, void* stream_) {
  if (!sim || !id || !dewi_v || !ent_v || !out_id || !out_score || !ready_flags) return fail("null argument");
  if (!(timeout_s > 0.0)) timeout_s = static_cast<double>(env_int("DEWI_PUSH_TIMEOUT_S", 120));
  if (B <= 0 || n_shards <= 0 || kcand <= 0 || k <= 0) return fail("B, n_shards, kcand and k must be positive");
  if (shard_stride_bytes % 8 != 0) return fail("shard stride must be a multiple of 8 bytes");
  const int64_t ncand = static_cast<int64_t>(n_shards) * kcand;
  if (cand_count > ncand) cand_count = static_cast<int>(ncand);
  if (k > cand_count) return fail("k exceeds the number of candidates (k > N)");
  DEWI_CUDA(cudaSetDevice(device));
  return launch_rerank(sim, id, dewi_v, ent_v, B, n_shards, kcand, shard_stride_bytes, cand_count, k,
                       static_cast<float>(1.0 - eta), static_cast<float>(eta), static_cast<float>(entropy_pref),
                       entropy_pref != 0.0 ? 1 : 0, out_id, out_score, static_cast<cudaStream_t>(stream_), ready_flags, seq,
                       status_word, timeout_s);
}

int dewi_rerank(const float* sim, const int64_t* id, const float* dewi_v, const float* ent_v, int B, int n_shards,
                int kcand, int64_t shard_stride_bytes, int cand_count, int k, double eta, double entropy_pref,
                int64_t* out_id, float* out_score, int device, void* stream_) {
  if (!sim || !id || !dewi_v || !ent_v || !out_id || !out_score) return fail("null argument");
  if (B <= 0 || n_shards <= 0 || kcand <= 0 || k <= 0) return fail("B, n_shards, kcand and k must be positive");
  if (shard_stride_bytes % 8 != 0) return fail("shard stride must be a multiple of 8 bytes");
  const int64_t ncand = static_cast<int64_t>(n_shards) * kcand;
  if (cand_count > ncand) cand_count = static_cast<int>(ncand);
  if (k > cand_count) return fail("k exceeds the number of candidates (k > N)");
  DEWI_CUDA(cudaSetDevice(device));
  // numpy evaluates (1 - eta) in float64 and rounds the weak scalar to float32 (backends.py:461)
  return launch_rerank(sim, id, dewi_v, ent_v, B, n_shards, kcand, shard_stride_bytes, cand_count, k,
                       static_cast<float>(1.0 - eta), static_cast<float>(eta), static_cast<float>(entropy_pref),
                       entropy_pref != 0.0 ? 1 : 0, out_id, out_score, static_cast<cudaStream_t>(stream_));
}

int dewi_index_set_blend(dewi_index_t* h, double eta, double entropy_pref) {
  if (!h) return fail("null handle");
  // the same weak-scalar float32 weights as the re-rank (backends.py:461-465)
  h->blend.w_sim = static_cast<float>(1.0 - eta);
  h->blend.w_dewi = static_cast<float>(eta);
  h->blend.pref = static_cast<float>(entropy_pref);
  h->blend.use_pref = entropy_pref != 0.0 ? 1 : 0;
  h->blend_set = true;
  return 0;
}

int dewi_index_search(dewi_index_t* h, const float* queries, int B, int k, double eta, double entropy_pref, int flags,
                      int64_t* out_id, float* out_score, void* stream_) {
  if (!h) return fail("null handle");
  if (B <= 0 || k <= 0) return fail("B and k must be positive");
  if (k > h->n) return fail("k exceeds the number of indexed rows (ExactIndex raises ValueError, backends.py:468)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // (the caller's current device is left as it was: the Python wrapper does not switch it around the call)
  struct DeviceGuard {
    int prev = -1, want;
    explicit DeviceGuard(int d) : want(d) {
      cudaGetDevice(&prev);
      if (prev != want) cudaSetDevice(want);
    }
    ~DeviceGuard() {
      if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
  } guard(h->device);
  const int kcand = static_cast<int>(std::min<int64_t>(2 * static_cast<int64_t>(k), h->n));  // backends.py:440
  const bool host_io = (flags & DEWI_FLAG_HOST_IO) != 0;
  const float* q_dev = queries;
  auto pinned = [](void** p, size_t* have, size_t need) -> int {
    if (need <= *have) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    DEWI_CUDA(cudaHostAlloc(p, need + need / 2, cudaHostAllocDefault));
    *have = need + need / 2;
    return 0;
  };
  const size_t q_bytes = static_cast<size_t>(B) * h->dim * 4;
  // small batches go through pinned staging; large ones are copied straight from the caller's memory
  const bool stage_io = host_io && q_bytes <= (size_t(4) << 20);
  if (host_io) {
    DEWI_TRY(h->qraw.ensure(q_bytes));
    const void* src = queries;
    if (stage_io) {
      DEWI_TRY(pinned(&h->pin_q, &h->pin_q_bytes, q_bytes));
      std::memcpy(h->pin_q, queries, q_bytes);
      src = h->pin_q;
    }
    DEWI_CUDA(cudaMemcpyAsync(h->qraw.p, src, q_bytes, cudaMemcpyHostToDevice, stream));
    q_dev = h->qraw.as<float>();
  }
  const size_t nc = static_cast<size_t>(B) * kcand;
  DEWI_TRY(h->loc_sim.ensure(nc * 4));
  DEWI_TRY(h->loc_id.ensure(nc * 8));
  DEWI_TRY(h->loc_dewi.ensure(nc * 4));
  DEWI_TRY(h->loc_ent.ensure(nc * 4));
  int64_t* d_id = out_id;
  float* d_sc = out_score;
  const size_t id_bytes = static_cast<size_t>(B) * k * 8, sc_bytes = static_cast<size_t>(B) * k * 4;
  if (host_io) {   // ids and scores share one device buffer, so the download is one copy
    DEWI_TRY(h->out_id.ensure(id_bytes + sc_bytes));
    d_id = h->out_id.as<int64_t>();
    d_sc = reinterpret_cast<float*>(h->out_id.as<char>() + id_bytes);
  }
  // numpy evaluates (1 - eta) in float64 and rounds the weak scalar to float32 (backends.py:461)
  const float w_sim = static_cast<float>(1.0 - eta);
  const float w_dewi = static_cast<float>(eta);
  const float pref = static_cast<float>(entropy_pref);
  const TailRerank rr{k, w_sim, w_dewi, pref, entropy_pref != 0.0 ? 1 : 0, d_id, d_sc};
  if (flags & DEWI_FLAG_SCOPE_FULL) DEWI_TRY(dewi_index_set_blend(h, eta, entropy_pref));
  bool reranked = false;
  DEWI_TRY(search_local_impl(h, q_dev, B, kcand, flags, h->loc_sim.as<float>(), h->loc_id.as<int64_t>(), h->loc_dewi.as<float>(),
                             h->loc_ent.as<float>(), stream_, nullptr, &rr, &reranked));
  if (!reranked) {
    DEWI_TRY(launch_rerank(h->loc_sim.as<float>(), h->loc_id.as<int64_t>(), h->loc_dewi.as<float>(), h->loc_ent.as<float>(),
                           B, 1, kcand, 0, kcand, k, w_sim, w_dewi, pref, entropy_pref != 0.0 ? 1 : 0, d_id, d_sc, stream));
    h->last_launches++;
  }
  if (host_io && stage_io) {
    DEWI_TRY(pinned(&h->pin_out, &h->pin_out_bytes, id_bytes + sc_bytes));
    DEWI_CUDA(cudaMemcpyAsync(h->pin_out, d_id, id_bytes + sc_bytes, cudaMemcpyDeviceToHost, stream));
    DEWI_CUDA(cudaStreamSynchronize(stream));
    std::memcpy(out_id, h->pin_out, id_bytes);
    std::memcpy(out_score, static_cast<char*>(h->pin_out) + id_bytes, sc_bytes);
  } else if (host_io) {
    DEWI_CUDA(cudaMemcpyAsync(out_id, d_id, id_bytes, cudaMemcpyDeviceToHost, stream));
    DEWI_CUDA(cudaMemcpyAsync(out_score, d_sc, sc_bytes, cudaMemcpyDeviceToHost, stream));
    DEWI_CUDA(cudaStreamSynchronize(stream));
  }
  return 0;
}

int dewi_index_set_profiling(dewi_index_t* h, int enable) {
  if (!h) return fail("null handle");
  DEWI_TRY(set_device(h));
  if (enable && !h->ev0[0]) {
    for (int i = 0; i < dewi_index::kEvRing; ++i) {
      DEWI_CUDA(cudaEventCreate(&h->ev0[i]));
      DEWI_CUDA(cudaEventCreate(&h->ev1[i]));
    }
  }
  h->profile = enable ? 1 : 0;
  h->searches = 0;
  return 0;
}

int dewi_index_sweep_ms(dewi_index_t* h, int back, float* ms, int* kind) {
  if (!h || !ms) return fail("null argument");
  if (!h->ev0[0]) return fail("profiling was never enabled on this handle");
  if (back < 0 || back >= dewi_index::kEvRing || back >= h->searches) return fail("no such profiled search");
  DEWI_TRY(set_device(h));
  const int slot = static_cast<int>((h->searches - 1 - back) % dewi_index::kEvRing);
  DEWI_CUDA(cudaEventSynchronize(h->ev1[slot]));
  DEWI_CUDA(cudaEventElapsedTime(ms, h->ev0[slot], h->ev1[slot]));
  if (kind) *kind = h->last_sweep_kind;
  return 0;
}

int dewi_index_cert_stats(const dewi_index_t* h, int64_t* used, int64_t* failed) {
  if (!h || !used || !failed) return fail("null argument");
  *used = h->cert_used;
  *failed = h->cert_failed;
  return 0;
}

int dewi_plan_probe(int which, int mode, int dim, int64_t n_rows, int n_qb, int kc, int sm_count, int q_rows, int opt,
                    int* out) {
  if (!out) return fail("null argument");
  for (int i = 0; i < 10; ++i) out[i] = 0;
  if (which == 0) {   // single-CTA sweeps: queries on M (search_tc.cu) or, with opt = batch size <= 64, rows on M (search_tcr.cu)
    TcPlan p{};
    const int rc = tc_make_plan(mode, dim, n_rows, n_qb, kc, sm_count, &p, 0, q_rows, opt);
    if (rc != 0) return rc;
    const int v[10] = {p.rows_on_m, p.q_rows, p.n_tile, p.n_stages, p.q_stages, p.q_resident, p.n_chunks, p.grid,
                       static_cast<int>(p.smem_bytes), p.mode};
    for (int i = 0; i < 10; ++i) out[i] = v[i];
    return 0;
  }
  if (which == 1) {   // CTA-pair sweep (search_tc2.cu); opt = 1 asks for the staged form
    Tc2Plan p{};
    const int rc = tc2_make_plan(mode, dim, n_rows, n_qb, kc, sm_count, &p, 0, 0, 0, opt);
    if (rc != 0) return rc;
    const int v[10] = {p.mode, p.n_stages, p.q_stages, p.n_chunks, p.grid, p.first_items, static_cast<int>(p.smem_bytes),
                       2 * tc2_box_rows(), 0, 0};
    for (int i = 0; i < 10; ++i) out[i] = v[i];
    return 0;
  }
  return fail("plan probe: which must be 0 (single-CTA sweeps) or 1 (CTA-pair sweep)");
}

int dewi_index_last_launches(const dewi_index_t* h, int* launches) {
  if (!h || !launches) return fail("null argument");
  *launches = h->last_launches;
  return 0;
}

}  // extern "C"
