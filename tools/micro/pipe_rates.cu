// Issue-rate micro-benchmark of the instructions the attention softmax / dS math is made of (sm_100a):
// lane-operations per clock per SM of MUFU.EX2, F2FP.BF16.PACK_AB, an integer bf16 pack (IADD + PRMT), FFMA, and mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_rates tools/micro/pipe_rates.cu && /tmp/pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t f2fp(float lo, float hi) {
  uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ uint32_t ipack(float lo, float hi) {  // round-half-up bf16 pair on the integer pipe
  uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u, r;
  asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  float v[ILP];
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = seed + threadIdx.x * 1e-3f + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]);                                    // 1 MUFU
      if (MODE == 1) { acc ^= f2fp(v[i], v[(i + 1) % ILP]); v[i] += 1.0f; }   // 1 F2FP (+1 FADD)
      if (MODE == 2) { acc ^= ipack(v[i], v[(i + 1) % ILP]); v[i] += 1.0f; }  // 2 IADD + 1 PRMT (+1 FADD)
      if (MODE == 3) v[i] = fmaf(v[i], 1.0001f, 0.5f);                    // 1 FFMA
      if (MODE == 4) { v[i] = ex2(v[i]); acc ^= f2fp(v[i], v[(i + 1) % ILP]); }   // MUFU + F2FP
      if (MODE == 5) { v[i] = ex2(v[i]); acc ^= ipack(v[i], v[(i + 1) % ILP]); }  // MUFU + integer pack
      if (MODE == 6) {  // exp2 on the FMA pipe: magic-number floor + degree-3 polynomial + exponent insert
        const float x = fmaxf(v[i], -126.f);
        const float tt = x + 12582912.f;   // 1.5 * 2^23: the integer part of x lands in the low mantissa bits
        const float f = x - (tt - 12582912.f);  // [-0.5, 0.5]
        float pl = fmaf(f, 0.0555041f, 0.2402265f);
        pl = fmaf(pl, f, 0.6931472f);
        pl = fmaf(pl, f, 1.0f);
        v[i] = __uint_as_float(__float_as_uint(pl) + (__float_as_uint(tt) << 23));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_elem, int warps) {
  float* out; long long* cyc;
  const int blocks = 148, threads = warps * 32;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, threads>>>(out, cyc, 0.25f);
  k<MODE><<<blocks, threads>>>(out, cyc, 0.25f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; ++i) c += h[i]; c /= blocks;
  const double elems = double(ITERS) * ILP * threads;
  printf("%-44s warps/SM %2d  %7.2f elements/clk/SM  (%d counted op(s) per element -> %7.2f lane-ops/clk/SM)\n", name, warps,
         elems / c, ops_per_elem, elems * ops_per_elem / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {8, 16, 32}) {
    run<0>("MUFU.EX2", 1, w);
    run<1>("F2FP.BF16.PACK_AB (+FADD)", 1, w);
    run<2>("integer bf16 pack: 2 IADD + PRMT (+FADD)", 3, w);
    run<3>("FFMA", 1, w);
    run<4>("MUFU.EX2 + F2FP", 2, w);
    run<5>("MUFU.EX2 + integer pack", 4, w);
    run<6>("exp2 on the FMA pipe (poly3)", 1, w);
  }
  return 0;
}
