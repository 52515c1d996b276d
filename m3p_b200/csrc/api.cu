// Host-side plumbing shared by every entry point: thread-local error text, the cached TMA
// descriptor factory (cuTensorMapEncodeTiled resolved through the runtime so the library has no
// link-time dependency on libcuda and still loads on a machine without a driver), device checks.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace m3p {

static thread_local char tls_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_last_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return M3P_ERR_CUDA;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static const uint64_t* g_seed_mix = nullptr;
const uint64_t* seed_mix_ptr() { return g_seed_mix; }

float* scratch_f32(size_t n_floats) {
  static float* buf = nullptr;
  static size_t cap = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (n_floats > cap) {
    size_t want = n_floats < (size_t(4) << 20) ? (size_t(4) << 20) : n_floats * 2;
    if (buf != nullptr) {
      cudaDeviceSynchronize();
      cudaFree(buf);
      buf = nullptr;
      cap = 0;
    }
    if (cudaMalloc(&buf, want * sizeof(float)) != cudaSuccess) {
      set_last_error("scratch allocation of %zu bytes failed", want * sizeof(float));
      buf = nullptr;
      return nullptr;
    }
    cap = want;
  }
  return buf;
}

// ---- tensor-map cache ------------------------------------------------------------------------
struct TmapKey {
  uint64_t v[10];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available (cuda error %d)", (int)e);
    return nullptr;
  }
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

static int encode_cached(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                         const uint64_t* pitches_elems, const uint32_t* box, int elem_bytes = 2) {
  TmapKey key{};
  key.v[0] = reinterpret_cast<uint64_t>(ptr);
  key.v[1] = (uint64_t)rank;
  key.v[9] = (uint64_t)elem_bytes;
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[5 + i] = (i + 1 < rank) ? pitches_elems[i] : 0;
    key.v[8] = key.v[8] * 1024 + box[i];
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *out = it->second; return M3P_OK; }
  }
  auto fn = get_encode_fn();
  if (!fn) return M3P_ERR_CUDA;
  cuuint64_t gdim[3];
  cuuint64_t gstride[2];
  cuuint32_t bdim[3];
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstride[i] = pitches_elems[i] * (uint64_t)elem_bytes;  // bytes
  for (int i = 0; i + 1 < rank; ++i)
    if (gstride[i] % 16 != 0) {
      set_last_error("TMA: row pitch %llu bytes is not a multiple of 16", (unsigned long long)gstride[i]);
      return M3P_ERR_INVALID_ARGUMENT;
    }
  // the box's inner extent picks the swizzle: 64 bf16 = 128-byte rows (operand tiles), 32 bf16 = 64-byte rows
  // (the GEMM epilogue's staging tiles)
  const CUtensorMapSwizzle swz = (box[0] * elem_bytes == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim,
                  gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return M3P_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
  }
  return M3P_OK;
}

int get_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                     uint64_t pitch_elems, uint32_t box0, uint32_t box1) {
  const uint64_t dims[2] = {dim0, dim1};
  const uint64_t pitches[1] = {pitch_elems};
  const uint32_t box[2] = {box0, box1};
  return encode_cached(out, ptr, 2, dims, pitches, box);
}

int get_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t pitch_elems,
                    uint32_t box0, uint32_t box1) {
  const uint64_t dims[2] = {dim0, dim1};
  const uint64_t pitches[1] = {pitch_elems};
  const uint32_t box[2] = {box0, box1};
  return encode_cached(out, ptr, 2, dims, pitches, box, 4);
}

int get_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                     uint64_t pitch1, uint64_t pitch2, uint32_t box0, uint32_t box1, uint32_t box2) {
  const uint64_t dims[3] = {dim0, dim1, dim2};
  const uint64_t pitches[2] = {pitch1, pitch2};
  const uint32_t box[3] = {box0, box1, box2};
  return encode_cached(out, ptr, 3, dims, pitches, box);
}

}  // namespace m3p

extern "C" int m3p_version(void) { return 101; }

extern "C" int m3p_set_seed_mix(const uint64_t* device_word) {
  m3p::g_seed_mix = device_word;
  return M3P_OK;
}

extern "C" const char* m3p_last_error(void) { return m3p::tls_error; }

extern "C" int m3p_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return m3p::cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return m3p::cuda_fail(e, "cudaDeviceGetAttribute", __FILE__, __LINE__);
  if (major != 10) {
    m3p::set_last_error("m3p_b200 requires an sm_100 device (B200); found compute capability %d.x", major);
    return M3P_ERR_UNSUPPORTED;
  }
  return M3P_OK;
}
