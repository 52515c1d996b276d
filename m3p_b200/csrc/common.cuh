// Shared host/device helpers: error reporting across the C ABI, the counter-based dropout
// generator (identical in every kernel that needs the same mask in forward and backward),
// erf-form GELU and its derivative, warp reductions.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/m3p_b200.h"

namespace m3p {

// ---- error plumbing (host) ------------------------------------------------------------------
void set_last_error(const char* fmt, ...);  // api.cu
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define M3P_CUDA_OK(expr)                                                     \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return m3p::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define M3P_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      m3p::set_last_error(__VA_ARGS__);     \
      return M3P_ERR_INVALID_ARGUMENT;      \
    }                                       \
  } while (0)

// ---- TMA descriptor construction (host, cached) ----------------------------------------------
// 2-D bf16 tensor: dim0 (contiguous) x dim1, row pitch in elements, box0 x box1, 128B swizzle.
int get_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                     uint64_t pitch_elems, uint32_t box0, uint32_t box1);
// 2-D fp32 tensor (the GEMM epilogue's fp32 residual tiles: box0 = 16 floats = 64-byte rows, 64B swizzle).
int get_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t pitch_elems,
                    uint32_t box0, uint32_t box1);
// 3-D bf16 tensor: dim0 (contiguous) x dim1 x dim2 with element pitches pitch1, pitch2.
int get_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                     uint64_t dim2, uint64_t pitch1, uint64_t pitch2, uint32_t box0, uint32_t box1,
                     uint32_t box2);
int sm_count();
// Library-owned scratch (per-row losses of the cross-entropy kernels): one buffer per process (= per GPU),
// grown on demand, reused by every call on the stream.
float* scratch_f32(size_t n_floats);

// Device word XOR-ed into every dropout seed at kernel start (m3p_set_seed_mix): lets a captured CUDA graph
// replay with fresh masks — the caller bumps the word between replays, the launch parameters stay constant.
const uint64_t* seed_mix_ptr();  // api.cu; NULL when unset
__device__ __forceinline__ void mix_seed(const uint64_t* mix, uint32_t& lo, uint32_t& hi) {
  if (mix != nullptr) {
    const uint64_t m = *mix;
    lo ^= static_cast<uint32_t>(m);
    hi ^= static_cast<uint32_t>(m >> 32);
  }
}

// ---- dropout generator -------------------------------------------------------------------------
// One 32-bit hash per PAIR of consecutive elements; each 16-bit half decides one element:
// keep iff half >= thr16, thr16 = round(p * 65536).  Element index is the row-major linear index
// of the tensor the mask applies to, taken modulo 2^32.  tests/test_dropout_generator.py restates it in numpy.
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t pair_idx, uint32_t seed_lo,
                                                       uint32_t seed_hi) {
  // two multiply / xor-shift rounds (7 integer instructions per element PAIR): plenty for a Bernoulli
  // mask over sequential indices, and cheap enough to regenerate in every backward kernel
  uint32_t x = (pair_idx ^ seed_lo) * 0x9E3779B1u + seed_hi;
  x ^= x >> 15;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_thr16(float p) {
  return static_cast<uint32_t>(p * 65536.0f + 0.5f);
}
// keep-flag for a single element (slow path; vector paths hash once per pair)
__device__ __forceinline__ bool drop_keep(uint32_t elem_idx, uint32_t seed_lo, uint32_t seed_hi,
                                          uint32_t thr16) {
  uint32_t h = drop_hash(elem_idx >> 1, seed_lo, seed_hi);
  uint32_t half = (elem_idx & 1) ? (h >> 16) : (h & 0xffffu);
  return half >= thr16;
}

// ---- GELU (erf form, reference transformer.py:48-56) -------------------------------------------
// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16 rounding of the result);
// also returns e = exp(-z^2) which the derivative reuses.
__device__ __forceinline__ float erf_as(float z, float* e_out) {
  const float a = fabsf(z);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, a, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = exp2f(-a * a * 1.4426950408889634f);
  *e_out = e;
  return copysignf(fmaf(-poly, e, 1.0f), z);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float er = erf_as(x * 0.70710678118654752f, &e);
  return 0.5f * x * (1.0f + er);
}
// d/dx [0.5 x (1 + erf(x/sqrt2))] = 0.5 (1 + erf(x/sqrt2)) + x exp(-x^2/2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float er = erf_as(x * 0.70710678118654752f, &e);
  return fmaf(x * e, 0.3989422804014327f, 0.5f * (1.0f + er));
}

// GELU and its derivative from ONE erf / exp evaluation, on raw MUFU approximations (no denormal /
// range fix-up code: the arguments are bounded, |erf err| stays ~1e-6, far below bf16 rounding).
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& gp) {
  const float a = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752f, a, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));  // exp(-x^2 / 2)
  const float er = copysignf(fmaf(-poly, e, 1.0f), x);
  const float cdf = fmaf(0.5f, er, 0.5f);
  g = x * cdf;
  gp = fmaf(x * e, 0.3989422804014327f, cdf);
}

// ---- warp reductions ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace m3p
