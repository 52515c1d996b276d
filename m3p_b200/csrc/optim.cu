// Optimizer step of the M3P trainer as two HBM-bound multi-tensor passes over the FLAT parameter / gradient
// buffers (reference: Trainer.optimize, M3P/src/xtrainer.py:205-243 — clip_grad_norm_ over every parameter,
// then optim.Adam.step, M3P/src/optim.py:45-86 — a Python loop over ~400 tensors x ~10 elementwise kernels).
//
//   pass 1  m3p_sumsq_f32   : sum of squares of a gradient buffer -> one device scalar (atomic across CTAs)
//   pass 2  m3p_adam_step   : per element, reading the scalar on the device (no host sync):
//               g'   = g * min(1, max_norm / (sqrt(sumsq) + 1e-6))            torch clip_grad_norm_
//               m    = b1 m + (1 - b1) g' ;  v = b2 v + (1 - b2) g'^2          optim.py:72-73
//               p   -= wd * lr * p                                             optim.py:80-81
//               p   -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)  optim.py:74-83
//             and, fused into the same pass because the bytes are already in registers: the bf16 tensor-core
//             copy of the new parameter (replaces the per-step cast kernel) and the zeroing of the gradient
//             (replaces zero_grad's memset).
// 16-byte vector accesses, grid = 8 CTAs per SM, grid-stride.  Traffic per parameter: 16 B read (p, g, m, v) +
// 12 B written (p, m, v) + 2 B (bf16 copy) + 4 B (zeroed gradient) = 34 B.
#include "common.cuh"
#include "ptx.cuh"

namespace m3p {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS)
sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    a0 = fmaf(v.x, v.x, a0); a1 = fmaf(v.y, v.y, a1); a2 = fmaf(v.z, v.z, a2); a3 = fmaf(v.w, v.w, a3);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail (n not a multiple of 4)
    const float t = x[(n4 << 2) + threadIdx.x];
    a0 = fmaf(t, t, a0);
  }
  float s = warp_sum((a0 + a1) + (a2 + a3));
  __shared__ float part[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < OPT_THREADS / 32 ? part[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);  // one atomic per CTA (<= 8 x SMs in total)
  }
}

struct AdamParams {
  float* p; float* g; float* m; float* v; __nv_bfloat16* p16;
  long long n;
  float lr, beta1, beta2, eps, weight_decay, step_size;  // step_size = lr * sqrt(1 - b2^t) / (1 - b1^t)
  const float* sumsq; float max_norm;
  int zero_grad;
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamParams& a, float coef) {
  g *= coef;
  m = fmaf(a.beta1, m, (1.0f - a.beta1) * g);
  v = fmaf(a.beta2, v, (1.0f - a.beta2) * g * g);
  const float denom = sqrtf(v) + a.eps;
  if (a.weight_decay != 0.f) p = fmaf(-a.weight_decay * a.lr, p, p);
  p = fmaf(-a.step_size, m / denom, p);
  return p;
}

__global__ void __launch_bounds__(OPT_THREADS)
adam_step_kernel(const AdamParams a) {
  float coef = 1.0f;
  if (a.sumsq != nullptr && a.max_norm > 0.f) {
    coef = a.max_norm / (sqrtf(*a.sumsq) + 1e-6f);  // torch.nn.utils.clip_grad_norm_
    coef = coef < 1.0f ? coef : 1.0f;
  }
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g = reinterpret_cast<const float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    adam_one(p.x, g.x, m.x, v.x, a, coef);
    adam_one(p.y, g.y, m.y, v.y, a, coef);
    adam_one(p.z, g.z, m.z, v.z, a, coef);
    adam_one(p.w, g.w, m.w, v.w, a, coef);
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
    if (a.p16 != nullptr)
      reinterpret_cast<uint2*>(a.p16)[i] = make_uint2(pack_bf16x2(p.x, p.y), pack_bf16x2(p.z, p.w));
    if (a.zero_grad) reinterpret_cast<float4*>(a.g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {  // tail
    const long long i = (n4 << 2) + threadIdx.x;
    float p = a.p[i], m = a.m[i], v = a.v[i];
    adam_one(p, a.g[i], m, v, a, coef);
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
    if (a.p16 != nullptr) a.p16[i] = __float2bfloat16_rn(p);
    if (a.zero_grad) a.g[i] = 0.f;
  }
}

static int opt_grid(long long n4) {
  long long need = (n4 + OPT_THREADS - 1) / OPT_THREADS;
  long long cap = (long long)sm_count() * 8;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace m3p

using namespace m3p;

extern "C" int m3p_sumsq_f32(const float* x, int64_t n, float* out, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(x && out, "m3p_sumsq_f32: null pointer");
  M3P_REQUIRE(n > 0, "m3p_sumsq_f32: empty buffer");
  M3P_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "m3p_sumsq_f32: buffer must be 16-byte aligned");
  sumsq_kernel<<<opt_grid(n >> 2), OPT_THREADS, 0, stream>>>(x, n, out);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_adam_step(const m3p_adam_args* a, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(a && a->param && a->grad && a->exp_avg && a->exp_avg_sq, "m3p_adam_step: null pointer");
  M3P_REQUIRE(a->n > 0, "m3p_adam_step: empty buffer");
  M3P_REQUIRE(a->step >= 1, "m3p_adam_step: step counts from 1 (optim.py:68)");
  M3P_REQUIRE(a->beta1 >= 0.f && a->beta1 < 1.f && a->beta2 >= 0.f && a->beta2 < 1.f && a->lr >= 0.f && a->eps >= 0.f,
              "m3p_adam_step: invalid hyper-parameter");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->param) | reinterpret_cast<uintptr_t>(a->grad) |
                       reinterpret_cast<uintptr_t>(a->exp_avg) | reinterpret_cast<uintptr_t>(a->exp_avg_sq);
  M3P_REQUIRE((al & 15) == 0, "m3p_adam_step: buffers must be 16-byte aligned");
  M3P_REQUIRE(a->param_bf16 == nullptr || (reinterpret_cast<uintptr_t>(a->param_bf16) & 7) == 0,
              "m3p_adam_step: bf16 copy must be 8-byte aligned");
  AdamParams p{};
  p.p = a->param; p.g = a->grad; p.m = a->exp_avg; p.v = a->exp_avg_sq;
  p.p16 = reinterpret_cast<__nv_bfloat16*>(a->param_bf16);
  p.n = a->n;
  p.lr = a->lr; p.beta1 = a->beta1; p.beta2 = a->beta2; p.eps = a->eps; p.weight_decay = a->weight_decay;
  // bias corrections in double on the host, exactly as the reference's Python floats (optim.py:76-78)
  const double bc1 = 1.0 - pow((double)a->beta1, (double)a->step);
  const double bc2 = 1.0 - pow((double)a->beta2, (double)a->step);
  p.step_size = (float)((double)a->lr * sqrt(bc2) / bc1);
  p.sumsq = a->grad_sumsq; p.max_norm = a->max_grad_norm;
  p.zero_grad = a->zero_grad;
  adam_step_kernel<<<opt_grid(a->n >> 2), OPT_THREADS, 0, stream>>>(p);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}
