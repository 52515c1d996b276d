// HBM-bound kernels of the M3P path: LayerNorm forward/backward (with the residual-stream row mask
// and the dropout of the preceding linear folded into the backward), bias-gradient column sums,
// fp32->bf16 casts, gather / scatter of token rows, cross-entropy forward+backward, tiny linears.
// One warp owns one row; 16-byte vector loads; warp-shuffle reductions; no shared-memory staging
// (every element is touched once).  Grids are sized as multiples of the SM count.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace m3p {

// ---- 8-wide vector access ---------------------------------------------------------------------
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* v) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
  v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
}
__device__ __forceinline__ void load8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                            pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void store8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// 8 dropout keep-flags for elements [e0, e0+8), e0 even
__device__ __forceinline__ void drop8(uint32_t e0, uint32_t seed_lo, uint32_t seed_hi, uint32_t thr16,
                                      float scale, float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = drop_hash((e0 >> 1) + j, seed_lo, seed_hi);
    v[2 * j] = ((h & 0xffffu) >= thr16) ? v[2 * j] * scale : 0.f;
    v[2 * j + 1] = ((h >> 16) >= thr16) ? v[2 * j + 1] * scale : 0.f;
  }
}

// 8 elements kept as loaded (16 bytes of bf16 / 32 bytes of fp32) so that several rows can be in flight per thread
// without holding their unpacked floats in registers
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float* f) const {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void unpack(float* f) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

constexpr int EW_THREADS = 256;
constexpr int EW_WARPS = EW_THREADS / 32;

static int ew_grid(long long rows) {
  long long need = (rows + EW_WARPS - 1) / EW_WARPS;
  long long cap = (long long)sm_count() * 8;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// =================================================================================================
// LayerNorm forward:  y = mask * ((x - mean) * rstd * gamma + beta)      transformer.py:953,957-958
// =================================================================================================
template <typename XT, int MAXC>
__global__ void __launch_bounds__(EW_THREADS)
ln_fwd_kernel(const XT* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, const int32_t* __restrict__ seqlen, long long S,
              __nv_bfloat16* __restrict__ y, float* __restrict__ y32, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, long long rows, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  const int nchunks = d >> 3;
  // (two rows in flight per warp were measured slower: 430 vs 370 us per step — more registers, 3 CTAs/SM instead of 4)
  for (long long row = warp0; row < rows; row += nwarps) {
    float v[MAXC][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        load8(x + row * d + ch * 8, v[c]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[c][j];
      }
    }
    const float mean = warp_sum(sum) / d;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float t = v[c][j] - mean; sq += t * t; }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / d + eps);
    bool valid = true;
    if (seqlen != nullptr) valid = (row % S) < seqlen[row / S];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        float g[8], b[8], o[8];
        load8(gamma + ch * 8, g);
        load8(beta + ch * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = valid ? fmaf((v[c][j] - mean) * rstd, g[j], b[j]) : 0.f;
        store8(y + row * d + ch * 8, o);
        if (y32 != nullptr) store8(y32 + row * d + ch * 8, o);  // the fp32 copy the residual add of the next linear reads
      }
    }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  }
}

// Row-pipelined LayerNorm forward for the encoder's big launches: one CTA of 12 warps per SM, every warp streams its
// rows through a private ring of shared-memory stages filled by 1-D bulk copies (cp.async.bulk, completion on the
// warp's own mbarriers).  A warp-per-row kernel that loads straight into registers has one row per warp in flight
// and alternates between waiting and computing; here up to LNP_STAGES rows per warp (>= 100 KB per SM) are in flight
// while the previous row is being normalised, which is what a ~25 us HBM-bound launch needs on a B200.
constexpr int LNP_WARPS = 12;
constexpr int LNP_THREADS = LNP_WARPS * 32;
constexpr int LNP_MAX_STAGES = 4;
constexpr uint32_t LNP_SMEM_BUDGET = 216 * 1024;

// 4-wide accessors: lane l owns columns [128 i + 4 l, +4) of vector i, so a warp's shared-memory access is 32
// consecutive 16-byte (fp32) / 8-byte (bf16) words — conflict-free — and its global stores are fully coalesced
__device__ __forceinline__ void load4(const float* p, float* v) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float* v) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
}
__device__ __forceinline__ void store4(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float* v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}
// 4 dropout keep-flags for elements [e0, e0+4), e0 a multiple of 4 (same generator as drop8)
__device__ __forceinline__ void drop4(uint32_t e0, uint32_t seed_lo, uint32_t seed_hi, uint32_t thr16, float scale,
                                      float* v) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint32_t h = drop_hash((e0 >> 1) + j, seed_lo, seed_hi);
    v[2 * j] = ((h & 0xffffu) >= thr16) ? v[2 * j] * scale : 0.f;
    v[2 * j + 1] = ((h >> 16) >= thr16) ? v[2 * j + 1] * scale : 0.f;
  }
}

// NV = d / 128 vectors of 4 columns per lane (d a multiple of 128)
template <typename XT, int NV>
__global__ void __launch_bounds__(LNP_THREADS, 1)
ln_fwd_pipe_kernel(const XT* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const int32_t* __restrict__ seqlen, long long S, __nv_bfloat16* __restrict__ y,
                   float* __restrict__ y32, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows,
                   int d, float eps, int stages) {
  extern __shared__ __align__(128) uint8_t lnp_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)d * sizeof(XT);
  uint8_t* ring = lnp_smem + (size_t)warp * stages * row_bytes;
  float* sgamma = reinterpret_cast<float*>(lnp_smem + (size_t)LNP_WARPS * stages * row_bytes);
  float* sbeta = sgamma + d;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbeta + d) + warp * LNP_MAX_STAGES;
  for (int c = threadIdx.x; c < d; c += LNP_THREADS) { sgamma[c] = gamma[c]; sbeta[c] = beta[c]; }
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const long long warp0 = (long long)blockIdx.x * LNP_WARPS + warp;
  const long long nwarps = (long long)gridDim.x * LNP_WARPS;
  const long long n_my = warp0 < rows ? (rows - warp0 + nwarps - 1) / nwarps : 0;
  auto issue = [&](long long k) {  // lane 0: request this warp's k-th row into ring slot k % stages
    const int s = (int)(k % stages);
    mbar_arrive_expect_tx(&bars[s], row_bytes);
    bulk_load_1d(ring + (size_t)s * row_bytes, x + (warp0 + k * nwarps) * d, row_bytes, &bars[s]);
  };
  if (lane == 0)
    for (long long k = 0; k < n_my && k < stages; ++k) issue(k);
  // this lane's gamma / beta never change: keep them in registers (2 * 4 * NV)
  float gm[NV][4], bt[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) { load4(sgamma + i * 128 + lane * 4, gm[i]); load4(sbeta + i * 128 + lane * 4, bt[i]); }
  const float inv_d = 1.0f / d;
  for (long long k = 0; k < n_my; ++k) {
    const long long row = warp0 + k * nwarps;
    const int s = (int)(k % stages);
    bool valid = true;
    if (seqlen != nullptr) valid = (row % S) < seqlen[row / S];
    mbar_wait(&bars[s], (uint32_t)((k / stages) & 1));
    const XT* xr = reinterpret_cast<const XT*>(ring + (size_t)s * row_bytes);
    float v[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) load4(xr + i * 128 + lane * 4, v[i]);
    __syncwarp();  // every lane has its vectors in registers: the slot can take the row after next
    if (lane == 0 && k + stages < n_my) issue(k + stages);
    float p4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains instead of one of 4 NV dependent adds
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) p4[j] += v[i][j];
    const float mean = warp_sum((p4[0] + p4[1]) + (p4[2] + p4[3])) * inv_d;
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[i][j] -= mean; q4[j] = fmaf(v[i][j], v[i][j], q4[j]); }
    const float rstd = rsqrtf(warp_sum((q4[0] + q4[1]) + (q4[2] + q4[3])) * inv_d + eps);
    const float rs = valid ? rstd : 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = valid ? fmaf(v[i][j] * rs, gm[i][j], bt[i][j]) : 0.f;
      store4(y + row * d + i * 128 + lane * 4, o);
      if (y32 != nullptr) store4(y32 + row * d + i * 128 + lane * 4, o);
    }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  }
}

// =================================================================================================
// LayerNorm backward.
//   dy_eff = mask * dropout_dy(dy)                 (dropout_dy: a dropout that followed the LN)
//   xhat = (x - mean) * rstd ;  g = dy_eff * gamma
//   dx = rstd * (g - mean_j(g) - xhat * mean_j(g * xhat))
//   dgamma += sum_rows dy_eff * xhat ; dbeta += sum_rows dy_eff
//   dx_drop = dropout_dx(dx)  (the dropout of the linear whose output fed this LN through the
//             residual add: transformer.py:951,226) ; dbias += sum_rows dx_drop   (that linear's bias)
// =================================================================================================
struct LnBwdParams {
  const void* dy; const void* x; const float* mean; const float* rstd; const float* gamma;
  const int32_t* seqlen; long long S;
  void* dx; __nv_bfloat16* dx_drop;
  uint32_t dx_thr16, dx_seed_lo, dx_seed_hi; float dx_scale;
  uint32_t dy_thr16, dy_seed_lo, dy_seed_hi; float dy_scale;
  const uint64_t* seed_mix;
  float* dgamma; float* dbeta; float* dbias;
  long long rows; int d;
  float* col_scratch;  // optional [LN_PARTS_MAX][3][d]: per-CTA column partials of the fused row+column pass
};

// Pass 1 — one warp per row, no per-thread column accumulators (keeps registers low and occupancy high:
// the kernel is latency-bound on its two row loads).
template <typename XT, typename DYT, typename DXT, int MAXC>
__global__ void __launch_bounds__(EW_THREADS, (MAXC <= 3) ? 3 : 2)
ln_bwd_dx_kernel(const LnBwdParams p) {
  const int d = p.d;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  const int nchunks = d >> 3;
  const XT* x = reinterpret_cast<const XT*>(p.x);
  const DYT* dy = reinterpret_cast<const DYT*>(p.dy);
  DXT* dx = reinterpret_cast<DXT*>(p.dx);
  uint32_t dy_lo = p.dy_seed_lo, dy_hi = p.dy_seed_hi, dx_lo = p.dx_seed_lo, dx_hi = p.dx_seed_hi;
  if (p.dy_thr16 != 0) mix_seed(p.seed_mix, dy_lo, dy_hi);
  if (p.dx_thr16 != 0) mix_seed(p.seed_mix, dx_lo, dx_hi);
  for (long long row = warp0; row < p.rows; row += nwarps) {
    bool valid = true;
    if (p.seqlen != nullptr) valid = (row % p.S) < p.seqlen[row / p.S];
    const float mean = p.mean[row], rstd = p.rstd[row];
    float xh[MAXC][8], g[MAXC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        float xv[8], dv[8], gm[8];
        load8(x + row * d + ch * 8, xv);
        load8(dy + row * d + ch * 8, dv);
        load8(p.gamma + ch * 8, gm);
        if (p.dy_thr16 != 0)
          drop8((uint32_t)row * (uint32_t)d + ch * 8, dy_lo, dy_hi, p.dy_thr16, p.dy_scale, dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[c][j] = (xv[j] - mean) * rstd;
          g[c][j] = valid ? dv[j] * gm[j] : 0.f;
          s1 += g[c][j];
          s2 = fmaf(g[c][j], xh[c][j], s2);
        }
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[c][j] - s1 - xh[c][j] * s2);
        store8(dx + row * d + ch * 8, o);
        if (p.dx_drop != nullptr) {
          if (p.dx_thr16 != 0)
            drop8((uint32_t)row * (uint32_t)d + ch * 8, dx_lo, dx_hi, p.dx_thr16, p.dx_scale, o);
          store8(p.dx_drop + row * d + ch * 8, o);
        }
      }
    }
  }
}

// Fused pass — the row pass above PLUS the column sums (dgamma, dbeta, dbias) accumulated in registers while the row
// is in flight, so dy / x / dx_drop are read once instead of twice (the fp32 residual stream doubled their size).
// Every lane owns the same 8-column chunks of every row it visits; a CTA folds its warps through shared memory and
// stores ONE partial per column into its slot of `col_scratch` (no atomics here: thousands of CTAs adding into
// the same 2 304 addresses serialise in L2); ln_colpart_reduce_kernel adds the slots.
constexpr int LNF_THREADS = 128;
constexpr int LNF_WARPS = LNF_THREADS / 32;
constexpr int LN_PARTS_MAX = 512;
template <typename XT, typename DYT, typename DXT, int MAXC>
__global__ void __launch_bounds__(LNF_THREADS, 3)
ln_bwd_fused_kernel(const LnBwdParams p) {
  __shared__ float sfold[LNF_WARPS][MAXC * 256];
  const int d = p.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * LNF_WARPS + warp;
  const long long nwarps = (long long)gridDim.x * LNF_WARPS;
  const int nchunks = d >> 3;
  const XT* x = reinterpret_cast<const XT*>(p.x);
  const DYT* dy = reinterpret_cast<const DYT*>(p.dy);
  DXT* dx = reinterpret_cast<DXT*>(p.dx);
  uint32_t dy_lo = p.dy_seed_lo, dy_hi = p.dy_seed_hi, dx_lo = p.dx_seed_lo, dx_hi = p.dx_seed_hi;
  if (p.dy_thr16 != 0) mix_seed(p.seed_mix, dy_lo, dy_hi);
  if (p.dx_thr16 != 0) mix_seed(p.seed_mix, dx_lo, dx_hi);
  float ag[MAXC][8], ab[MAXC][8], abias[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[c][j] = 0.f; ab[c][j] = 0.f; abias[c][j] = 0.f; }
  for (long long row = warp0; row < p.rows; row += nwarps) {
    bool valid = true;
    if (p.seqlen != nullptr) valid = (row % p.S) < p.seqlen[row / p.S];
    const float mean = p.mean[row], rstd = p.rstd[row];
    float xh[MAXC][8], g[MAXC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        float xv[8], dv[8], gm[8];
        load8(x + row * d + ch * 8, xv);
        load8(dy + row * d + ch * 8, dv);
        load8(p.gamma + ch * 8, gm);
        if (p.dy_thr16 != 0)
          drop8((uint32_t)row * (uint32_t)d + ch * 8, dy_lo, dy_hi, p.dy_thr16, p.dy_scale, dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[c][j] = (xv[j] - mean) * rstd;
          const float de = valid ? dv[j] : 0.f;  // dy_eff
          ag[c][j] = fmaf(de, xh[c][j], ag[c][j]);
          ab[c][j] += de;
          g[c][j] = de * gm[j];
          s1 += g[c][j];
          s2 = fmaf(g[c][j], xh[c][j], s2);
        }
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[c][j] - s1 - xh[c][j] * s2);
        store8(dx + row * d + ch * 8, o);
        if (p.dx_drop != nullptr) {
          if (p.dx_thr16 != 0)
            drop8((uint32_t)row * (uint32_t)d + ch * 8, dx_lo, dx_hi, p.dx_thr16, p.dx_scale, o);
          store8(p.dx_drop + row * d + ch * 8, o);
        }
        // bias gradient of the preceding linear: column sum of its output gradient, taken from the fp32 values
#pragma unroll
        for (int j = 0; j < 8; ++j) abias[c][j] += o[j];
      }
    }
  }
  // fold the CTA's warps, one accumulator kind at a time; partial slot layout [blockIdx.x][kind][d]
  float* slot = p.col_scratch + (long long)blockIdx.x * 3 * d;
#pragma unroll
  for (int kind = 0; kind < 3; ++kind) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sfold[warp][(lane + 32 * c) * 8 + j] = kind == 0 ? ag[c][j] : (kind == 1 ? ab[c][j] : abias[c][j]);
    __syncthreads();
    for (int col = threadIdx.x; col < d; col += LNF_THREADS) {
      float sum = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < LNF_WARPS; ++w2) sum += sfold[w2][col];
      slot[kind * d + col] = sum;
    }
  }
}

// The same fused pass with the rows streamed through per-warp shared-memory rings by bulk copies (see
// ln_fwd_pipe_kernel): the 72 column accumulators per lane leave no registers to prefetch the next row the classic
// way, and with one row per warp in flight the register version reaches 68 % of the HBM peak.  One CTA of 12 warps
// per SM => 148 partial slots instead of 444 for the reduction that follows.
template <typename XT, typename DYT, typename DXT, int NV>
__global__ void __launch_bounds__(LNP_THREADS, 1)
ln_bwd_pipe_kernel(const LnBwdParams p, int stages) {
  extern __shared__ __align__(128) uint8_t lnp_smem[];
  const int d = p.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t xb = (uint32_t)d * sizeof(XT), yb = (uint32_t)d * sizeof(DYT), stage_bytes = xb + yb;
  uint8_t* ring = lnp_smem + (size_t)warp * stages * stage_bytes;
  float* sgamma = reinterpret_cast<float*>(lnp_smem + (size_t)LNP_WARPS * stages * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sgamma + d) + warp * LNP_MAX_STAGES;
  for (int c = threadIdx.x; c < d; c += LNP_THREADS) sgamma[c] = p.gamma[c];
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const XT* x = reinterpret_cast<const XT*>(p.x);
  const DYT* dy = reinterpret_cast<const DYT*>(p.dy);
  DXT* dx = reinterpret_cast<DXT*>(p.dx);
  const long long warp0 = (long long)blockIdx.x * LNP_WARPS + warp;
  const long long nwarps = (long long)gridDim.x * LNP_WARPS;
  const long long n_my = warp0 < p.rows ? (p.rows - warp0 + nwarps - 1) / nwarps : 0;
  auto issue = [&](long long k) {
    const int s = (int)(k % stages);
    const long long row = warp0 + k * nwarps;
    mbar_arrive_expect_tx(&bars[s], stage_bytes);
    bulk_load_1d(ring + (size_t)s * stage_bytes, x + row * d, xb, &bars[s]);
    bulk_load_1d(ring + (size_t)s * stage_bytes + xb, dy + row * d, yb, &bars[s]);
  };
  if (lane == 0)
    for (long long k = 0; k < n_my && k < stages; ++k) issue(k);
  uint32_t dy_lo = p.dy_seed_lo, dy_hi = p.dy_seed_hi, dx_lo = p.dx_seed_lo, dx_hi = p.dx_seed_hi;
  if (p.dy_thr16 != 0) mix_seed(p.seed_mix, dy_lo, dy_hi);
  if (p.dx_thr16 != 0) mix_seed(p.seed_mix, dx_lo, dx_hi);
  const float inv_d = 1.0f / d;
  float ag[NV][4], ab[NV][4], abias[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { ag[i][j] = 0.f; ab[i][j] = 0.f; abias[i][j] = 0.f; }
  for (long long k = 0; k < n_my; ++k) {
    const long long row = warp0 + k * nwarps;
    const int s = (int)(k % stages);
    bool valid = true;
    if (p.seqlen != nullptr) valid = (row % p.S) < p.seqlen[row / p.S];
    const float mean = p.mean[row], rstd = p.rstd[row];
    mbar_wait(&bars[s], (uint32_t)((k / stages) & 1));
    const XT* xr = reinterpret_cast<const XT*>(ring + (size_t)s * stage_bytes);
    const DYT* dr = reinterpret_cast<const DYT*>(ring + (size_t)s * stage_bytes + xb);
    float xh[NV][4], g[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) { load4(xr + i * 128 + lane * 4, xh[i]); load4(dr + i * 128 + lane * 4, g[i]); }
    __syncwarp();  // the row is in registers: refill the slot
    if (lane == 0 && k + stages < n_my) issue(k + stages);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const float nmr = -mean * rstd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float gm[4];
      load4(sgamma + i * 128 + lane * 4, gm);
      if (p.dy_thr16 != 0)
        drop4((uint32_t)row * (uint32_t)d + i * 128 + lane * 4, dy_lo, dy_hi, p.dy_thr16, p.dy_scale, g[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xh[i][j] = fmaf(xh[i][j], rstd, nmr);
        const float de = valid ? g[i][j] : 0.f;  // dy_eff
        ag[i][j] = fmaf(de, xh[i][j], ag[i][j]);
        ab[i][j] += de;
        g[i][j] = de * gm[j];
        s1[j] += g[i][j];
        s2[j] = fmaf(g[i][j], xh[i][j], s2[j]);
      }
    }
    const float m1 = warp_sum((s1[0] + s1[1]) + (s1[2] + s1[3])) * inv_d;
    const float m2 = warp_sum((s2[0] + s2[1]) + (s2[2] + s2[3])) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = rstd * (g[i][j] - m1 - xh[i][j] * m2);
      store4(dx + row * d + i * 128 + lane * 4, o);
      if (p.dx_drop != nullptr) {
        if (p.dx_thr16 != 0)
          drop4((uint32_t)row * (uint32_t)d + i * 128 + lane * 4, dx_lo, dx_hi, p.dx_thr16, p.dx_scale, o);
        store4(p.dx_drop + row * d + i * 128 + lane * 4, o);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) abias[i][j] += o[j];
    }
  }
  // every bulk copy this CTA issued has been waited for: the rings are free to hold the cross-warp fold
  float* sfold = reinterpret_cast<float*>(lnp_smem);  // [LNP_WARPS][d]
  float* slot = p.col_scratch + (long long)blockIdx.x * 3 * d;
#pragma unroll
  for (int kind = 0; kind < 3; ++kind) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float* src = kind == 0 ? ag[i] : (kind == 1 ? ab[i] : abias[i]);
      store4(sfold + warp * d + i * 128 + lane * 4, src);
    }
    __syncthreads();
    for (int col = threadIdx.x; col < d; col += LNP_THREADS) {
      float sum = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < LNP_WARPS; ++w2) sum += sfold[w2 * d + col];
      slot[kind * d + col] = sum;
    }
  }
}

// dgamma / dbeta / dbias += sum over the n_part slots of col_scratch.  grid (ceil(3d / 256), LNR_GROUPS): every thread
// adds its share of the slots for one (kind, column) and issues one atomic (LNR_GROUPS-way contention per address).
constexpr int LNR_GROUPS = 8;
__global__ void __launch_bounds__(EW_THREADS)
ln_colpart_reduce_kernel(const float* __restrict__ part, int n_part, int d, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, float* __restrict__ dbias) {
  const int idx = blockIdx.x * EW_THREADS + threadIdx.x;  // kind * d + col
  if (idx >= 3 * d) return;
  const int kind = idx / d, col = idx - kind * d;
  float* dst = kind == 0 ? dgamma : (kind == 1 ? dbeta : dbias);
  if (dst == nullptr) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int s = blockIdx.y;
  for (; s + 3 * LNR_GROUPS < n_part; s += 4 * LNR_GROUPS) {
    a0 += part[(long long)s * 3 * d + idx];
    a1 += part[(long long)(s + LNR_GROUPS) * 3 * d + idx];
    a2 += part[(long long)(s + 2 * LNR_GROUPS) * 3 * d + idx];
    a3 += part[(long long)(s + 3 * LNR_GROUPS) * 3 * d + idx];
  }
  for (; s < n_part; s += LNR_GROUPS) a0 += part[(long long)s * 3 * d + idx];
  atomicAdd(dst + col, (a0 + a1) + (a2 + a3));
}

// M3P_LN_PIPE=0 falls back to the register-staged kernels (A/B measurements)
static bool use_ln_pipe() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("M3P_LN_PIPE");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

static int ln_fused_grid(long long rows) {
  long long need = (rows + LNF_WARPS - 1) / LNF_WARPS;
  long long cap = (long long)sm_count() * 3;
  if (cap > LN_PARTS_MAX) cap = LN_PARTS_MAX;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// Column-reduction geometry shared by the LayerNorm column pass and the bias-gradient column sums:
// a CTA owns a stripe of TX*8 columns (TX threads x one 16-byte vector) and a contiguous block of rows,
// walked by TY = 256/TX row lanes; the TY partials are folded in shared memory and the CTA issues ONE
// atomic per column.  The grid is (stripes, row blocks) with few row blocks (<= ~24), so at most that
// many CTAs ever add into the same address: same-address atomics serialise in L2 at ~250 cycles per
// op on B200, which made the naive "one atomic per CTA per column" version 5x slower than the loads.
struct ColGeom { int tx, ty, stripes, row_blocks; long long rows_per_block; };
static ColGeom col_geom(long long rows, int n) {
  ColGeom g;
  const int nvec = n / 8;
  // widest stripe that still yields >= 2 CTAs per SM with <= 24 row blocks
  g.tx = 8;
  for (int t = 32; t >= 8; t >>= 1) {
    if (((nvec + t - 1) / t) * 24 >= 2 * sm_count()) { g.tx = t; break; }
  }
  g.ty = EW_THREADS / g.tx;
  g.stripes = (nvec + g.tx - 1) / g.tx;
  int rb = (2 * sm_count() + g.stripes - 1) / g.stripes;
  if (rb > 24) rb = 24;
  if (rb < 1) rb = 1;
  long long rpb = (rows + rb - 1) / rb;
  rpb = (rpb + g.ty - 1) / g.ty * g.ty;
  g.rows_per_block = rpb;
  g.row_blocks = (int)((rows + rpb - 1) / rpb);
  return g;
}

// Pass 2 — column sums over rows: dgamma = sum dy_eff * xhat, dbeta = sum dy_eff, dbias = sum dx_drop.
// dy / x were just read by pass 1 and dx_drop just written, so this pass mostly hits L2.
constexpr int LNC_U = 4;  // rows in flight per thread (3 x 16-byte loads each): the pass is latency-bound
template <typename XT, typename DYT, typename BT>
__global__ void __launch_bounds__(EW_THREADS, 2)
ln_bwd_cols_kernel(const LnBwdParams p, const BT* __restrict__ bias_src, int tx, long long rows_per_block) {
  extern __shared__ float sred[];  // [ty][tx * 8]
  const int d = p.d;
  const int ty = EW_THREADS / tx;
  const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
  const int col0 = (blockIdx.x * tx + cx) * 8;
  const bool col_ok = col0 < d;
  const XT* x = reinterpret_cast<const XT*>(p.x);
  const DYT* dy = reinterpret_cast<const DYT*>(p.dy);
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < p.rows ? r0 + rows_per_block : p.rows;
  float ag[8], ab[8], abias[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ag[j] = 0.f; ab[j] = 0.f; abias[j] = 0.f; }
  const bool want_gb = (p.dgamma != nullptr) || (p.dbeta != nullptr);
  uint32_t dy_lo = p.dy_seed_lo, dy_hi = p.dy_seed_hi;
  if (p.dy_thr16 != 0) mix_seed(p.seed_mix, dy_lo, dy_hi);
  if (col_ok) {
    for (long long r = r0 + ry; r < r1; r += LNC_U * ty) {
      Raw8<XT> xr[LNC_U];
      Raw8<DYT> dr[LNC_U];
      Raw8<BT> br[LNC_U];
      float mean[LNC_U], rstd[LNC_U];
      bool ok[LNC_U];
#pragma unroll
      for (int i = 0; i < LNC_U; ++i) {
        const long long row = r + (long long)i * ty;
        ok[i] = row < r1;
        const long long rr = ok[i] ? row : r0;
        if (want_gb) {
          xr[i].load(x + rr * d + col0);
          dr[i].load(dy + rr * d + col0);
          mean[i] = p.mean[rr];
          rstd[i] = p.rstd[rr];
        }
        if (bias_src != nullptr) br[i].load(bias_src + rr * d + col0);
      }
#pragma unroll
      for (int i = 0; i < LNC_U; ++i) {
        const long long row = r + (long long)i * ty;
        if (!ok[i]) continue;
        if (want_gb) {
          bool valid = true;
          if (p.seqlen != nullptr) valid = (row % p.S) < p.seqlen[row / p.S];
          if (valid) {
            float xv[8], dv[8];
            xr[i].unpack(xv);
            dr[i].unpack(dv);
            if (p.dy_thr16 != 0)
              drop8((uint32_t)row * (uint32_t)d + col0, dy_lo, dy_hi, p.dy_thr16, p.dy_scale, dv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              ag[j] = fmaf(dv[j], (xv[j] - mean[i]) * rstd[i], ag[j]);
              ab[j] += dv[j];
            }
          }
        }
        if (bias_src != nullptr) {
          float bv[8];
          br[i].unpack(bv);
#pragma unroll
          for (int j = 0; j < 8; ++j) abias[j] += bv[j];
        }
      }
    }
  }
  // fold the ty row lanes, then one atomic per column per CTA
  auto fold = [&](const float* acc, float* dst) {
    if (dst == nullptr) return;  // uniform
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sred[ry * (tx * 8) + cx * 8 + j] = acc[j];
    __syncthreads();
    for (int c = threadIdx.x; c < tx * 8; c += EW_THREADS) {
      float sum = 0.f;
      for (int y = 0; y < ty; ++y) sum += sred[y * (tx * 8) + c];
      const int col = blockIdx.x * tx * 8 + c;
      if (col < d) atomicAdd(dst + col, sum);
    }
  };
  fold(ag, p.dgamma);
  fold(ab, p.dbeta);
  fold(abias, p.dbias);
}

template <typename XT, typename DYT, typename DXT>
static int launch_ln_bwd(LnBwdParams p, cudaStream_t stream, int phases) {
  if (p.col_scratch != nullptr && p.d <= 1024 && (p.dgamma || p.dbeta || p.dbias)) {
    // fused row + column pass into per-CTA partials (phase 1), partial reduction (phase 2)
    const uint32_t stage_bytes = (uint32_t)p.d * (sizeof(XT) + sizeof(DYT));
    const uint32_t tail = (uint32_t)p.d * 4 + LNP_WARPS * LNP_MAX_STAGES * 8;
    int stages = (int)((LNP_SMEM_BUDGET - tail) / (LNP_WARPS * stage_bytes));
    if (stages > LNP_MAX_STAGES) stages = LNP_MAX_STAGES;
    const bool pipe = use_ln_pipe() && stages >= 2 && p.rows >= 4096 && p.d % 128 == 0 &&
                      (p.d == 256 || p.d == 512 || p.d == 768 || p.d == 1024);
    if (pipe) {
      const int pgrid = sm_count() < LN_PARTS_MAX ? sm_count() : LN_PARTS_MAX;
      const size_t smem = (size_t)LNP_WARPS * stages * stage_bytes + tail;
      if (phases & 1) {
#define M3P_LN_BWD_PIPE(NV)                                                                               \
  do {                                                                                                    \
    auto kfn = ln_bwd_pipe_kernel<XT, DYT, DXT, NV>;                                                    \
    M3P_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LNP_SMEM_BUDGET + 4096)); \
    kfn<<<pgrid, LNP_THREADS, smem, stream>>>(p, stages);                                                 \
  } while (0)
        if (p.d == 256) M3P_LN_BWD_PIPE(2);
        else if (p.d == 512) M3P_LN_BWD_PIPE(4);
        else if (p.d == 768) M3P_LN_BWD_PIPE(6);
        else M3P_LN_BWD_PIPE(8);
#undef M3P_LN_BWD_PIPE
        M3P_CUDA_OK(cudaGetLastError());
      }
      if (phases & 2) {
        const dim3 rgrid((3 * p.d + EW_THREADS - 1) / EW_THREADS, LNR_GROUPS);
        ln_colpart_reduce_kernel<<<rgrid, EW_THREADS, 0, stream>>>(p.col_scratch, pgrid, p.d, p.dgamma, p.dbeta, p.dbias);
        M3P_CUDA_OK(cudaGetLastError());
      }
      return M3P_OK;
    }
    const int grid = ln_fused_grid(p.rows);
    if (phases & 1) {
      if (p.d <= 256) ln_bwd_fused_kernel<XT, DYT, DXT, 1><<<grid, LNF_THREADS, 0, stream>>>(p);
      else if (p.d <= 768) ln_bwd_fused_kernel<XT, DYT, DXT, 3><<<grid, LNF_THREADS, 0, stream>>>(p);
      else ln_bwd_fused_kernel<XT, DYT, DXT, 4><<<grid, LNF_THREADS, 0, stream>>>(p);
      M3P_CUDA_OK(cudaGetLastError());
    }
    if (phases & 2) {
      const dim3 rgrid((3 * p.d + EW_THREADS - 1) / EW_THREADS, LNR_GROUPS);
      ln_colpart_reduce_kernel<<<rgrid, EW_THREADS, 0, stream>>>(p.col_scratch, grid, p.d, p.dgamma, p.dbeta, p.dbias);
      M3P_CUDA_OK(cudaGetLastError());
    }
    return M3P_OK;
  }
  if (phases & 1) {
    const int grid = ew_grid(p.rows);
    if (p.d <= 256) ln_bwd_dx_kernel<XT, DYT, DXT, 1><<<grid, EW_THREADS, 0, stream>>>(p);
    else if (p.d <= 768) ln_bwd_dx_kernel<XT, DYT, DXT, 3><<<grid, EW_THREADS, 0, stream>>>(p);
    else ln_bwd_dx_kernel<XT, DYT, DXT, 4><<<grid, EW_THREADS, 0, stream>>>(p);
    M3P_CUDA_OK(cudaGetLastError());
  }
  if (!(phases & 2) || !(p.dgamma || p.dbeta || p.dbias)) return M3P_OK;
  const ColGeom g = col_geom(p.rows, p.d);
  const dim3 cgrid(g.stripes, g.row_blocks);
  const size_t smem = (size_t)EW_THREADS * 8 * sizeof(float);
  if (p.dbias == nullptr) {
    ln_bwd_cols_kernel<XT, DYT, DXT><<<cgrid, EW_THREADS, smem, stream>>>(p, static_cast<const DXT*>(nullptr), g.tx,
                                                                         g.rows_per_block);
  } else if (p.dx_drop != nullptr) {
    ln_bwd_cols_kernel<XT, DYT, __nv_bfloat16><<<cgrid, EW_THREADS, smem, stream>>>(p, p.dx_drop, g.tx, g.rows_per_block);
  } else {
    ln_bwd_cols_kernel<XT, DYT, DXT><<<cgrid, EW_THREADS, smem, stream>>>(p, reinterpret_cast<const DXT*>(p.dx), g.tx,
                                                                         g.rows_per_block);
  }
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

// =================================================================================================
// column sums (bias gradients): out[j] += sum_rows x[row][j]
// =================================================================================================
__global__ void __launch_bounds__(EW_THREADS)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out, long long rows, int n,
              int tx, long long rows_per_block) {
  extern __shared__ float sred[];  // [ty][tx * 8]
  const int ty = EW_THREADS / tx;
  const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
  const int col0 = (blockIdx.x * tx + cx) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col0 < n) {
    long long r = r0 + ry;
    // 8 independent 16-byte loads in flight per thread
    for (; r + 7LL * ty < r1; r += 8LL * ty) {
      uint4 t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const uint4*>(x + (r + (long long)i * ty) * ld + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0] += bf16_lo(t[i].x); acc[1] += bf16_hi(t[i].x); acc[2] += bf16_lo(t[i].y); acc[3] += bf16_hi(t[i].y);
        acc[4] += bf16_lo(t[i].z); acc[5] += bf16_hi(t[i].z); acc[6] += bf16_lo(t[i].w); acc[7] += bf16_hi(t[i].w);
      }
    }
    for (; r < r1; r += ty) {
      float v[8];
      load8(x + r * ld + col0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sred[ry * (tx * 8) + cx * 8 + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < tx * 8; c += EW_THREADS) {
    float sum = 0.f;
    for (int y = 0; y < ty; ++y) sum += sred[y * (tx * 8) + c];
    const int col = blockIdx.x * tx * 8 + c;
    if (col < n) atomicAdd(out + col, sum);  // <= row_blocks-way contention per address
  }
}

// =================================================================================================
// casts / permutes
// =================================================================================================
__global__ void __launch_bounds__(EW_THREADS)
cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n8, float scale) {
  for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < n8;
       i += (long long)gridDim.x * EW_THREADS) {
    float v[8];
    load8(in + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= scale;
    store8(out + i * 8, v);
  }
}
// out = scale * float(in): the way back from a bf16 gradient all-reduce into the fp32 gradient buffer
__global__ void __launch_bounds__(EW_THREADS)
cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n8, float scale) {
  for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < n8;
       i += (long long)gridDim.x * EW_THREADS) {
    float v[8];
    load8(in + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= scale;
    store8(out + i * 8, v);
  }
}
// out = bf16(sum_s in[s * stride + i]), slabs added in index order: the deterministic reduction of split-K partials
// (m3p_gemm_args.split_stride) — no atomics, so a bf16 result cannot flip with the arrival order of the partial sums
__global__ void __launch_bounds__(EW_THREADS)
sum_slabs_bf16_kernel(const float* __restrict__ in, int n_slabs, long long stride, __nv_bfloat16* __restrict__ out,
                      long long n8) {
  for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < n8;
       i += (long long)gridDim.x * EW_THREADS) {
    float acc[8];
    load8(in + i * 8, acc);
    for (int s = 1; s < n_slabs; ++s) {
      float v[8];
      load8(in + (long long)s * stride + i * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    store8(out + i * 8, acc);
  }
}
// (A, B, F) fp32 -> (B, A, F) bf16     (seq-first reference inputs -> batch-major rows)
__global__ void __launch_bounds__(EW_THREADS)
permute_cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int A, int B, int F8) {
  const long long total = (long long)A * B * F8;
  for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * EW_THREADS) {
    const int f = (int)(i % F8);
    const long long ab = i / F8;
    const int a = (int)(ab % A);
    const int b = (int)(ab / A);  // output-major order: (b, a, f)
    float v[8];
    load8(in + (((long long)a * B + b) * F8 + f) * 8, v);
    store8(out + i * 8, v);
  }
}

// Device-side half of the region pipeline the reference runs on the host per sample (dataset_pretrain.py:258-292,379):
// zero the regions picked for masking, L2-normalise every 2048-d feature row (F.normalize, eps 1e-12), and hand the
// encoder its bf16 batch-major operand — one pass over the raw features instead of host loops + an fp32 round trip.
// One warp per (b, r) row.  in (R, B, F) fp32; out (B, R, F) bf16; ori (B, R, F) fp32 = the normalised UNMASKED
// features (the MRFR regression target, xtrainer.py:2340), optional.
__global__ void __launch_bounds__(EW_THREADS)
region_prep_kernel(const float* __restrict__ in, const uint8_t* __restrict__ zero_mask, int normalize,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ ori, int R, int B, int F) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  const long long rows = (long long)R * B;
  const int nv = F >> 3;
  for (long long o = warp0; o < rows; o += nwarps) {  // o = b * R + r (output-major)
    const int b = (int)(o / R), r = (int)(o % R);
    const float* src = in + ((long long)r * B + b) * F;
    float ss = 0.f;
    if (normalize) {
      for (int c = lane; c < nv; c += 32) {
        float v[8];
        load8(src + c * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
      }
      ss = warp_sum(ss);
    }
    const float sc = normalize ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
    const bool zero = zero_mask != nullptr && zero_mask[o] != 0;
    for (int c = lane; c < nv; c += 32) {  // second read of the row hits L1/L2
      float v[8];
      load8(src + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= sc;
      if (ori != nullptr) store8(ori + o * F + c * 8, v);
      if (zero) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      store8(out + o * F + c * 8, v);
    }
  }
}

// du = dg * gp (gp = gelu_erf'(u) stashed by the forward GEMM epilogue): backward of the GELU inside
// BertPredictionHeadTransform (transformer.py:603-604), where the LayerNorm backward sits between the
// next linear's dgrad and this activation.
__global__ void __launch_bounds__(EW_THREADS)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dg, const __nv_bfloat16* __restrict__ u,
                __nv_bfloat16* __restrict__ du, long long n8) {
  for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < n8;
       i += (long long)gridDim.x * EW_THREADS) {
    float a[8], b[8];
    load8(dg + i * 8, a);
    load8(u + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= b[j];
    store8(du + i * 8, a);
  }
}

// =================================================================================================
// row gather / scatter (prediction heads, FreeLB input grads)
//   gather : dst[i][:] = src[rows[i]][:]            (src addressed as base + row_t*stride_t + row_b*stride_b)
// =================================================================================================
__global__ void __launch_bounds__(EW_THREADS)
gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const int64_t* __restrict__ flat_idx, long long n_inner,
                   long long stride_outer, long long stride_inner, __nv_bfloat16* __restrict__ dst, long long n, int d) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  for (long long i = warp0; i < n; i += nwarps) {
    const long long f = flat_idx[i];
    const __nv_bfloat16* s = src + (f / n_inner) * stride_outer + (f % n_inner) * stride_inner;
    for (int ch = lane; ch < (d >> 3); ch += 32)
      *reinterpret_cast<uint4*>(dst + i * d + ch * 8) = *reinterpret_cast<const uint4*>(s + ch * 8);
  }
}
// scatter-add of bf16 rows into a (possibly strided) bf16 tensor is done by zero-fill + row copy
// because masked positions are unique:  dst[rows[i]][:] = src[i][:]
__global__ void __launch_bounds__(EW_THREADS)
scatter_rows_kernel(const __nv_bfloat16* __restrict__ src, const int64_t* __restrict__ flat_idx, long long n_inner,
                    long long stride_outer, long long stride_inner, __nv_bfloat16* __restrict__ dst, long long n, int d) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  for (long long i = warp0; i < n; i += nwarps) {
    const long long f = flat_idx[i];
    __nv_bfloat16* t = dst + (f / n_inner) * stride_outer + (f % n_inner) * stride_inner;
    for (int ch = lane; ch < (d >> 3); ch += 32)
      *reinterpret_cast<uint4*>(t + ch * 8) = *reinterpret_cast<const uint4*>(src + i * d + ch * 8);
  }
}

// =================================================================================================
// cross-entropy forward + backward over bf16 logits (F.cross_entropy(reduction='mean', ignore_index))
//   transformer.py:112 (MLM), :581 (MRM).  One CTA per row.
//   loss += -(logit[y] - lse) / n_valid ;  dlogits = (softmax - onehot) / n_valid   (0 for ignored rows)
// =================================================================================================
// *loss = mean over non-ignored rows of row_loss ; *inv_count = 1 / #non-ignored rows   (one CTA: a
// thousand CTAs adding into one address would serialise in L2)
__global__ void ce_finish_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ y, long long n,
                                 long long ignore_index, float* __restrict__ loss, float* __restrict__ inv_count) {
  __shared__ float s_sum[32];
  __shared__ int s_cnt[32];
  float sum = 0.f;
  int cnt = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    if (y[i] != ignore_index) { cnt += 1; sum += row_loss[i]; }
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    int c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += s_sum[w]; c += s_cnt[w]; }
    // 0 valid rows: torch gives nan; we give loss 0 and zero gradients
    *inv_count = c > 0 ? 1.0f / (float)c : 0.f;
    *loss = c > 0 ? t / (float)c : 0.f;
  }
}

// (m, s) <- combine((m, s), (m2, s2)) for the online log-sum-exp
__device__ __forceinline__ void lse_combine(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
  m = mn;
}

__global__ void __launch_bounds__(EW_THREADS)
ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, const int64_t* __restrict__ y, int V,
              long long ignore_index, float* __restrict__ row_loss, float* __restrict__ lse_out) {
  __shared__ float s_m[EW_WARPS], s_s[EW_WARPS];
  const long long row = blockIdx.x;
  const __nv_bfloat16* lp = logits + row * ld;
  const long long target = y[row];
  if (target == ignore_index) {
    if (threadIdx.x == 0) lse_out[row] = 0.f;
    return;
  }
  const int nvec = V >> 3;
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < nvec; i += EW_THREADS) {
    float v[8];
    load8(lp + i * 8, v);
    float mx = v[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) mx = fmaxf(mx, v[j]);
    if (mx > m) { s *= __expf(m - mx); m = mx; }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf(v[j] - m);
  }
  for (int i = nvec * 8 + threadIdx.x; i < V; i += EW_THREADS) {
    const float v = __bfloat162float(lp[i]);
    if (v > m) { s *= __expf(m - v); m = v; }
    s += __expf(v - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_combine(m, s, m2, s2);
  }
  if ((threadIdx.x & 31) == 0) { s_m[threadIdx.x >> 5] = m; s_s[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    m = s_m[0]; s = s_s[0];
#pragma unroll
    for (int w = 1; w < EW_WARPS; ++w) lse_combine(m, s, s_m[w], s_s[w]);
    const float lse = m + logf(s);
    lse_out[row] = lse;
    row_loss[row] = lse - __bfloat162float(lp[target]);
  }
}

__global__ void __launch_bounds__(EW_THREADS)
ce_bwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, const int64_t* __restrict__ y, int V,
              long long ignore_index, const float* __restrict__ lse_in, const float* __restrict__ inv_count,
              const float* __restrict__ grad_scale, __nv_bfloat16* __restrict__ dlogits, long long ldd) {
  const long long row = blockIdx.x;
  const __nv_bfloat16* lp = logits + row * ld;
  __nv_bfloat16* dp = dlogits + row * ldd;
  const long long target = y[row];
  const int nvec = V >> 3;
  if (target == ignore_index) {
    for (int i = threadIdx.x; i < nvec; i += EW_THREADS) *reinterpret_cast<uint4*>(dp + i * 8) = make_uint4(0, 0, 0, 0);
    for (int i = nvec * 8 + threadIdx.x; i < V; i += EW_THREADS) dp[i] = __float2bfloat16_rn(0.f);
    return;
  }
  const float lse = lse_in[row];
  const float g = (*inv_count) * (grad_scale != nullptr ? *grad_scale : 1.0f);
  for (int i = threadIdx.x; i < nvec; i += EW_THREADS) {
    float v[8];
    load8(lp + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float pj = __expf(v[j] - lse);
      if (i * 8 + j == target) pj -= 1.0f;
      v[j] = pj * g;
    }
    store8(dp + i * 8, v);
  }
  for (int i = nvec * 8 + threadIdx.x; i < V; i += EW_THREADS) {
    float pj = __expf(__bfloat162float(lp[i]) - lse);
    if (i == target) pj -= 1.0f;
    dp[i] = __float2bfloat16_rn(pj * g);
  }
}

// =================================================================================================
// masked MSE of the MRFR objective (xtrainer.py:2333-2348): F.mse_loss(pred[sel], target[sel]) over the masked
// regions, written as sum_rows w[row] * sum_f (pred - target)^2 with w[row] = sel[row] / (n_sel * d) prepared with
// the batch — no boolean gather, no host sync.  Rows with w == 0 are never read.
// =================================================================================================
__global__ void __launch_bounds__(EW_THREADS)
mse_rows_kernel(const __nv_bfloat16* __restrict__ pred, long long ld, const float* __restrict__ target,
                const float* __restrict__ w, float* __restrict__ row_loss, long long n, int d) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  for (long long row = warp0; row < n; row += nwarps) {
    const float wr = w[row];
    float acc = 0.f;
    if (wr != 0.f) {
      for (int c = lane; c < (d >> 3); c += 32) {
        float a[8], b[8];
        load8(pred + row * ld + c * 8, a);
        load8(target + row * (long long)d + c * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float t = a[j] - b[j]; acc = fmaf(t, t, acc); }
      }
      acc = warp_sum(acc) * wr;
    }
    if (lane == 0) row_loss[row] = acc;
  }
}
// *out = sum_i x[i], one CTA, fixed order (deterministic)
__global__ void sum_rows_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float s_sum[32];
  float sum = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) sum += x[i];
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) t += s_sum[wi];
    *out = t;
  }
}
// dpred[row][:] = bf16(2 * w[row] * g * (pred - target)); rows with w == 0 get zeros
__global__ void __launch_bounds__(EW_THREADS)
mse_bwd_kernel(const __nv_bfloat16* __restrict__ pred, long long ld, const float* __restrict__ target,
               const float* __restrict__ w, const float* __restrict__ grad_scale, __nv_bfloat16* __restrict__ dpred,
               long long ldd, long long n, int d) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  const float g = grad_scale != nullptr ? *grad_scale : 1.0f;
  for (long long row = warp0; row < n; row += nwarps) {
    const float wr = 2.0f * w[row] * g;
    for (int c = lane; c < (d >> 3); c += 32) {
      float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (wr != 0.f) {
        float b[8];
        load8(pred + row * ld + c * 8, a);
        load8(target + row * (long long)d + c * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = wr * (a[j] - b[j]);
      }
      store8(dpred + row * ldd + c * 8, a);
    }
  }
}

// =================================================================================================
// ITM loss of pretrain_under_step / t2i_step (xtrainer.py:2359-2372, 1917-1942) on the B = n_groups * sample_n
// matching scores, forward AND gradient in one launch (the eager version is ~30 one-microsecond kernels in the
// middle of the step):
//   loss = w_multi * mean_g( logsumexp(s_g) - s_g[pos_g] ) + w_bin * mean_i( BCEWithLogits(s_i, onehot_i) )
//   dscores_i = w_multi * (softmax_g(s)_i - onehot_i) / n_groups + w_bin * (sigmoid(s_i) - onehot_i) / B
// One CTA; thread t owns group t, t + blockDim, ...; the two means are reduced in a fixed order.
// =================================================================================================
__global__ void relation_loss_kernel(const float* __restrict__ scores, const int64_t* __restrict__ pos, int n_groups,
                                     int sample_n, float w_multi, float w_bin, float* __restrict__ loss,
                                     float* __restrict__ dscores) {
  __shared__ float s_ce[32], s_bce[32];
  float ce = 0.f, bce = 0.f;
  const float inv_g = 1.0f / (float)n_groups, inv_b = 1.0f / ((float)n_groups * (float)sample_n);
  for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
    const float* sg = scores + (long long)g * sample_n;
    const int p = (int)pos[g];
    float mx = -INFINITY;
    for (int j = 0; j < sample_n; ++j) mx = fmaxf(mx, sg[j]);
    float sum = 0.f;
    for (int j = 0; j < sample_n; ++j) sum += expf(sg[j] - mx);
    const float lse = mx + logf(sum);
    ce += lse - sg[p];
    for (int j = 0; j < sample_n; ++j) {
      const float x = sg[j], y = (j == p) ? 1.f : 0.f;
      bce += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
      const float sig = 1.0f / (1.0f + expf(-x));
      dscores[(long long)g * sample_n + j] = w_multi * (expf(x - lse) - y) * inv_g + w_bin * (sig - y) * inv_b;
    }
  }
  ce = warp_sum(ce);
  bce = warp_sum(bce);
  if ((threadIdx.x & 31) == 0) { s_ce[threadIdx.x >> 5] = ce; s_bce[threadIdx.x >> 5] = bce; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_ce[w]; b += s_bce[w]; }
    *loss = w_multi * a * inv_g + w_bin * b * inv_b;
  }
}

// =================================================================================================
// tiny linear d -> 1 (seq_relationship, transformer.py:713,1196) and its backward
// =================================================================================================
__global__ void __launch_bounds__(EW_THREADS)
rowdot_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, long long rows, int d) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  float acc = 0.f;
  for (int j = lane; j < d; j += 32) acc += __bfloat162float(x[row * d + j]) * w[j];
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc + bias[0];
}
// dx[row][j] = dout[row] * w[j] (* (1 - x^2) when x is a tanh output whose pre-activation gradient is
// wanted: BertPooler, transformer.py:556-557) ; dw[j] += sum_rows dout[row] * x[row][j] ; db += sum dout
__global__ void __launch_bounds__(EW_THREADS)
rowdot_bwd_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                  __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long rows, int d,
                  int tanh_grad) {
  // grid (column blocks, row blocks of 16): a dependent load per row made the single-pass loop latency-bound
  const int j = blockIdx.x * EW_THREADS + threadIdx.x;
  if (j >= d) return;
  float acc = 0.f, accb = 0.f;
  const float wj = w[j];
  const long long r_lo = (long long)blockIdx.y * 16, r_hi = r_lo + 16 < rows ? r_lo + 16 : rows;
#pragma unroll 4
  for (long long r = r_lo; r < r_hi; ++r) {
    const float g = dout[r];
    const float xv = __bfloat162float(x[r * d + j]);
    acc += g * xv;
    accb += g;
    dx[r * d + j] = __float2bfloat16_rn(tanh_grad ? g * wj * (1.0f - xv * xv) : g * wj);
  }
  atomicAdd(dw + j, acc);
  if (j == 0) atomicAdd(db, accb);
}

// dst[i][:] = table[idx[i]][:] (fp32): nn.Embedding lookup kept in fp32 (FreeLB's embeds_init, xtrainer.py:2700-2705)
__global__ void __launch_bounds__(EW_THREADS)
gather_rows_f32_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, float* __restrict__ dst,
                       long long n, int d) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  for (long long i = warp0; i < n; i += nwarps) {
    const float* s = table + idx[i] * d;
    for (int c = lane; c < (d >> 2); c += 32)
      *reinterpret_cast<float4*>(dst + i * d + c * 4) = *reinterpret_cast<const float4*>(s + c * 4);
  }
}

// same with bf16 source rows (the all-gathered embedding-gradient rows of the data-parallel exchange)
__global__ void __launch_bounds__(EW_THREADS)
scatter_add_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, const int64_t* __restrict__ idx, long long skip_index,
                             float* __restrict__ dst, long long n, int d) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  for (long long i = warp0; i < n; i += nwarps) {
    const long long r = idx[i];
    if (r == skip_index) continue;
    for (int c = lane; c < (d >> 3); c += 32) {
      float v[8];
      load8(src + i * d + c * 8, v);
      atomicAdd(reinterpret_cast<float4*>(dst + r * d + c * 8), make_float4(v[0], v[1], v[2], v[3]));
      atomicAdd(reinterpret_cast<float4*>(dst + r * d + c * 8 + 4), make_float4(v[4], v[5], v[6], v[7]));
    }
  }
}

// bf16 -> fp32 elementwise add into (embedding-style) fp32 rows:  dst[idx[i]][:] += src[i][:]
__global__ void __launch_bounds__(EW_THREADS)
scatter_add_rows_f32_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, long long skip_index,
                            float* __restrict__ dst, long long n, int d) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EW_WARPS;
  for (long long i = warp0; i < n; i += nwarps) {
    const long long r = idx[i];
    if (r == skip_index) continue;
    for (int c = lane; c < (d >> 2); c += 32) {
      const float4 v = *reinterpret_cast<const float4*>(src + i * d + c * 4);
      atomicAdd(reinterpret_cast<float4*>(dst + r * d + c * 4), v);
    }
  }
}

}  // namespace m3p

// =================================================================================================
// C ABI
// =================================================================================================
using namespace m3p;

template <typename XT>
static int launch_ln_fwd(const m3p_ln_fwd_args* a, cudaStream_t stream) {
  auto X = reinterpret_cast<const XT*>(a->x);
  auto Y = reinterpret_cast<__nv_bfloat16*>(a->y);
  const int d = (int)a->d;
  if (use_ln_pipe() && a->rows >= 4096 && (d == 256 || d == 512 || d == 768 || d == 1024)) {
    const uint32_t row_bytes = (uint32_t)d * sizeof(XT);
    const uint32_t tail = (uint32_t)d * 8 + LNP_WARPS * LNP_MAX_STAGES * 8;
    int stages = (int)((LNP_SMEM_BUDGET - tail) / (LNP_WARPS * row_bytes));
    if (stages > LNP_MAX_STAGES) stages = LNP_MAX_STAGES;
    const size_t smem = (size_t)LNP_WARPS * stages * row_bytes + tail;
    const int pgrid = sm_count();
#define M3P_LN_FWD_PIPE(NV)                                                                               \
  do {                                                                                                    \
    auto kfn = ln_fwd_pipe_kernel<XT, NV>;                                                              \
    M3P_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LNP_SMEM_BUDGET + 4096)); \
    kfn<<<pgrid, LNP_THREADS, smem, stream>>>(X, a->gamma, a->beta, a->seqlen, a->S, Y, a->y_f32, a->mean, a->rstd, \
                                              a->rows, d, a->eps, stages);                                \
  } while (0)
    if (d == 256) M3P_LN_FWD_PIPE(2);
    else if (d == 512) M3P_LN_FWD_PIPE(4);
    else if (d == 768) M3P_LN_FWD_PIPE(6);
    else M3P_LN_FWD_PIPE(8);
#undef M3P_LN_FWD_PIPE
    return M3P_OK;
  }
  const int grid = ew_grid(a->rows);
#define M3P_LN_FWD(MAXC) ln_fwd_kernel<XT, MAXC><<<grid, EW_THREADS, 0, stream>>>( \
      X, a->gamma, a->beta, a->seqlen, a->S, Y, a->y_f32, a->mean, a->rstd, a->rows, d, a->eps)
  if (d <= 256) M3P_LN_FWD(1);
  else if (d <= 768) M3P_LN_FWD(3);
  else if (d <= 1024) M3P_LN_FWD(4);
  else M3P_LN_FWD(8);
#undef M3P_LN_FWD
  return M3P_OK;
}

extern "C" int m3p_layernorm_fwd(const m3p_ln_fwd_args* a, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(a && a->x && a->gamma && a->beta && a->y && a->mean && a->rstd, "m3p_layernorm_fwd: null pointer");
  M3P_REQUIRE(a->rows > 0 && a->d > 0 && a->d % 8 == 0 && a->d <= 2048,
              "m3p_layernorm_fwd: d=%lld must be a multiple of 8, <= 2048", (long long)a->d);
  M3P_REQUIRE(a->seqlen == nullptr || a->S > 0, "m3p_layernorm_fwd: S must be > 0 with a row mask");
  const int rc = a->x_f32 ? launch_ln_fwd<float>(a, stream) : launch_ln_fwd<__nv_bfloat16>(a, stream);
  if (rc) return rc;
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

// phases: bit 0 = the row pass (dx, dx_drop), bit 1 = the column pass (dgamma, dbeta, dbias)
static int layernorm_bwd_impl(const m3p_ln_bwd_args* a, m3p_stream_t stream_, int phases) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(a && a->dy && a->x && a->mean && a->rstd && a->gamma && a->dx, "m3p_layernorm_bwd: null pointer");
  M3P_REQUIRE(a->rows > 0 && a->d > 0 && a->d % 8 == 0 && a->d <= 1024,
              "m3p_layernorm_bwd: d=%lld must be a multiple of 8, <= 1024", (long long)a->d);
  M3P_REQUIRE(a->seqlen == nullptr || a->S > 0, "m3p_layernorm_bwd: S must be > 0 with a row mask");
  M3P_REQUIRE(a->dx_drop_p >= 0.f && a->dx_drop_p < 1.f && a->dy_drop_p >= 0.f && a->dy_drop_p < 1.f,
              "m3p_layernorm_bwd: dropout probability out of range");
  LnBwdParams p{};
  p.dy = a->dy; p.x = a->x; p.mean = a->mean; p.rstd = a->rstd; p.gamma = a->gamma;
  p.seqlen = a->seqlen; p.S = a->S;
  p.dx = a->dx; p.dx_drop = reinterpret_cast<__nv_bfloat16*>(a->dx_drop);
  p.dx_thr16 = a->dx_drop_p > 0.f ? drop_thr16(a->dx_drop_p) : 0;
  p.dx_scale = 1.0f / (1.0f - a->dx_drop_p);
  p.dx_seed_lo = (uint32_t)(a->dx_seed & 0xffffffffu); p.dx_seed_hi = (uint32_t)(a->dx_seed >> 32);
  p.dy_thr16 = a->dy_drop_p > 0.f ? drop_thr16(a->dy_drop_p) : 0;
  p.dy_scale = 1.0f / (1.0f - a->dy_drop_p);
  p.dy_seed_lo = (uint32_t)(a->dy_seed & 0xffffffffu); p.dy_seed_hi = (uint32_t)(a->dy_seed >> 32);
  p.seed_mix = seed_mix_ptr();
  p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.dbias = a->dbias;
  p.rows = a->rows; p.d = (int)a->d;
  p.col_scratch = a->col_scratch;
  const int key = (a->x_f32 ? 4 : 0) | (a->dy_f32 ? 2 : 0) | (a->dx_f32 ? 1 : 0);
  switch (key) {
    case 0: return launch_ln_bwd<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(p, stream, phases);
    case 5: return launch_ln_bwd<float, __nv_bfloat16, float>(p, stream, phases);
    case 6: return launch_ln_bwd<float, float, __nv_bfloat16>(p, stream, phases);
    case 4: return launch_ln_bwd<float, __nv_bfloat16, __nv_bfloat16>(p, stream, phases);
    case 7: return launch_ln_bwd<float, float, float>(p, stream, phases);
    default:
      set_last_error("m3p_layernorm_bwd: unsupported dtype combination x_f32=%d dy_f32=%d dx_f32=%d", a->x_f32,
                     a->dy_f32, a->dx_f32);
      return M3P_ERR_UNSUPPORTED;
  }
}

extern "C" int m3p_layernorm_bwd(const m3p_ln_bwd_args* a, m3p_stream_t stream) {
  return layernorm_bwd_impl(a, stream, 3);
}

extern "C" int m3p_layernorm_bwd_rows(const m3p_ln_bwd_args* a, m3p_stream_t stream) {
  return layernorm_bwd_impl(a, stream, 1);
}

extern "C" int m3p_layernorm_bwd_cols(const m3p_ln_bwd_args* a, m3p_stream_t stream) {
  return layernorm_bwd_impl(a, stream, 2);
}

extern "C" int m3p_colsum_bf16(const void* x, int64_t ld, float* out, int64_t rows, int64_t n, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(x && out, "m3p_colsum_bf16: null pointer");
  // n itself may be ragged (V = 250 002): the kernel reads whole 8-column vectors inside the row pitch and only adds
  // the columns < n
  M3P_REQUIRE(rows > 0 && n > 0 && ld % 8 == 0 && (n + 7) / 8 * 8 <= ld,
              "m3p_colsum_bf16: ld must be a multiple of 8 and cover n rounded up to 8");
  const ColGeom g = col_geom(rows, (int)((n + 7) / 8 * 8));
  colsum_kernel<<<dim3(g.stripes, g.row_blocks), EW_THREADS, (size_t)EW_THREADS * 8 * sizeof(float), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ld, out, rows, (int)n, g.tx, g.rows_per_block);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_cast_f32_bf16(const float* in, void* out, int64_t n, float scale, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(in && out, "m3p_cast_f32_bf16: null pointer");
  M3P_REQUIRE(n > 0 && n % 8 == 0, "m3p_cast_f32_bf16: n must be a positive multiple of 8");
  const long long n8 = n / 8;
  long long g = (n8 + EW_THREADS - 1) / EW_THREADS;
  const long long cap = (long long)sm_count() * 16;
  cast_f32_bf16_kernel<<<(int)(g < cap ? g : cap), EW_THREADS, 0, stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out), n8, scale);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_cast_bf16_f32(const void* in, float* out, int64_t n, float scale, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(in && out, "m3p_cast_bf16_f32: null pointer");
  M3P_REQUIRE(n > 0 && n % 8 == 0, "m3p_cast_bf16_f32: n must be a positive multiple of 8");
  const long long n8 = n / 8;
  long long g = (n8 + EW_THREADS - 1) / EW_THREADS;
  const long long cap = (long long)sm_count() * 16;
  cast_bf16_f32_kernel<<<(int)(g < cap ? g : cap), EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, n8, scale);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_sum_slabs_bf16(const float* in, int64_t n_slabs, int64_t slab_stride, void* out, int64_t n,
                                  m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(in && out, "m3p_sum_slabs_bf16: null pointer");
  M3P_REQUIRE(n > 0 && n % 8 == 0 && n_slabs >= 1 && n_slabs < (1 << 20) && slab_stride % 4 == 0 && slab_stride >= n,
              "m3p_sum_slabs_bf16: n must be a positive multiple of 8, slab_stride a multiple of 4 and >= n");
  const long long n8 = n / 8;
  long long g = (n8 + EW_THREADS - 1) / EW_THREADS;
  const long long cap = (long long)sm_count() * 16;
  sum_slabs_bf16_kernel<<<(int)(g < cap ? g : cap), EW_THREADS, 0, stream>>>(in, (int)n_slabs, slab_stride,
                                                                           reinterpret_cast<__nv_bfloat16*>(out), n8);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_gelu_bwd(const void* dg, const void* u, void* du, int64_t n, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(dg && u && du, "m3p_gelu_bwd: null pointer");
  M3P_REQUIRE(n > 0 && n % 8 == 0, "m3p_gelu_bwd: n must be a positive multiple of 8");
  const long long n8 = n / 8;
  long long g = (n8 + EW_THREADS - 1) / EW_THREADS;
  const long long cap = (long long)sm_count() * 16;
  gelu_bwd_kernel<<<(int)(g < cap ? g : cap), EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dg),
                                                                     reinterpret_cast<const __nv_bfloat16*>(u),
                                                                     reinterpret_cast<__nv_bfloat16*>(du), n8);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_permute_cast_f32_bf16(const float* in, void* out, int64_t A, int64_t B, int64_t F,
                                         m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(in && out, "m3p_permute_cast_f32_bf16: null pointer");
  M3P_REQUIRE(A > 0 && B > 0 && F > 0 && F % 8 == 0, "m3p_permute_cast_f32_bf16: F must be a multiple of 8");
  const long long total = A * B * (F / 8);
  long long g = (total + EW_THREADS - 1) / EW_THREADS;
  const long long cap = (long long)sm_count() * 16;
  permute_cast_kernel<<<(int)(g < cap ? g : cap), EW_THREADS, 0, stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out),
                                                                         (int)A, (int)B, (int)(F / 8));
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_region_prep(const float* in, const uint8_t* zero_mask, int32_t normalize, void* out, float* ori,
                               int64_t R, int64_t B, int64_t F, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(in && out, "m3p_region_prep: null pointer");
  M3P_REQUIRE(R > 0 && B > 0 && F > 0 && F % 8 == 0, "m3p_region_prep: F must be a multiple of 8");
  region_prep_kernel<<<ew_grid(R * B), EW_THREADS, 0, stream>>>(in, zero_mask, (int)normalize,
                                                               reinterpret_cast<__nv_bfloat16*>(out), ori, (int)R, (int)B, (int)F);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_gather_rows_bf16(const void* src, const int64_t* flat_idx, int64_t n_inner, int64_t stride_outer,
                                    int64_t stride_inner, void* dst, int64_t n, int64_t d, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(src && flat_idx && dst, "m3p_gather_rows_bf16: null pointer");
  M3P_REQUIRE(n > 0 && d % 8 == 0 && stride_outer % 8 == 0 && stride_inner % 8 == 0 && n_inner > 0,
              "m3p_gather_rows_bf16: d and strides must be multiples of 8");
  gather_rows_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), flat_idx, n_inner,
                                                            stride_outer, stride_inner,
                                                            reinterpret_cast<__nv_bfloat16*>(dst), n, (int)d);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_scatter_rows_bf16(const void* src, const int64_t* flat_idx, int64_t n_inner, int64_t stride_outer,
                                     int64_t stride_inner, void* dst, int64_t n, int64_t d, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(src && flat_idx && dst, "m3p_scatter_rows_bf16: null pointer");
  M3P_REQUIRE(n > 0 && d % 8 == 0 && stride_outer % 8 == 0 && stride_inner % 8 == 0 && n_inner > 0,
              "m3p_scatter_rows_bf16: d and strides must be multiples of 8");
  scatter_rows_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), flat_idx, n_inner,
                                                             stride_outer, stride_inner,
                                                             reinterpret_cast<__nv_bfloat16*>(dst), n, (int)d);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_cross_entropy_fwd(const void* logits, int64_t ld, const int64_t* y, int64_t n, int64_t V,
                                     int64_t ignore_index, float* loss, float* lse, float* inv_count,
                                     m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(logits && y && loss && lse && inv_count, "m3p_cross_entropy_fwd: null pointer");
  M3P_REQUIRE(n > 0 && V > 0 && ld % 8 == 0 && ld >= V, "m3p_cross_entropy_fwd: pitch must be a multiple of 8 and >= V");
  float* row_loss = scratch_f32((size_t)n);
  if (row_loss == nullptr) return M3P_ERR_CUDA;
  ce_fwd_kernel<<<(unsigned)n, EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), ld, y, (int)V,
                                                        ignore_index, row_loss, lse);
  ce_finish_kernel<<<1, 1024, 0, stream>>>(row_loss, y, n, ignore_index, loss, inv_count);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_cross_entropy_bwd(const void* logits, int64_t ld, const int64_t* y, int64_t n, int64_t V,
                                     int64_t ignore_index, const float* lse, const float* inv_count,
                                     const float* grad_scale, void* dlogits, int64_t ldd, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(logits && y && lse && inv_count && dlogits, "m3p_cross_entropy_bwd: null pointer");
  M3P_REQUIRE(n > 0 && V > 0 && ld % 8 == 0 && ldd % 8 == 0 && ld >= V && ldd >= V,
              "m3p_cross_entropy_bwd: pitches must be multiples of 8 and >= V");
  ce_bwd_kernel<<<(unsigned)n, EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), ld, y, (int)V,
                                                        ignore_index, lse, inv_count, grad_scale,
                                                        reinterpret_cast<__nv_bfloat16*>(dlogits), ldd);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_masked_mse_fwd(const void* pred, int64_t ld, const float* target, const float* weight, int64_t n,
                                  int64_t d, float* loss, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(pred && target && weight && loss, "m3p_masked_mse_fwd: null pointer");
  M3P_REQUIRE(n > 0 && d > 0 && d % 8 == 0 && ld % 8 == 0 && ld >= d, "m3p_masked_mse_fwd: d and ld must be multiples of 8");
  float* row_loss = scratch_f32((size_t)n);
  if (row_loss == nullptr) return M3P_ERR_CUDA;
  mse_rows_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(pred), ld, target, weight,
                                                         row_loss, n, (int)d);
  sum_rows_kernel<<<1, 1024, 0, stream>>>(row_loss, n, loss);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_masked_mse_bwd(const void* pred, int64_t ld, const float* target, const float* weight,
                                  const float* grad_scale, void* dpred, int64_t ldd, int64_t n, int64_t d,
                                  m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(pred && target && weight && dpred, "m3p_masked_mse_bwd: null pointer");
  M3P_REQUIRE(n > 0 && d > 0 && d % 8 == 0 && ld % 8 == 0 && ldd % 8 == 0 && ld >= d && ldd >= d,
              "m3p_masked_mse_bwd: d and pitches must be multiples of 8");
  mse_bwd_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(pred), ld, target, weight,
                                                        grad_scale, reinterpret_cast<__nv_bfloat16*>(dpred), ldd, n, (int)d);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_relation_loss(const float* scores, const int64_t* pos_labels, int64_t n_groups, int64_t sample_n,
                                 float w_multi, float w_bin, float* loss, float* dscores, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(scores && pos_labels && loss && dscores, "m3p_relation_loss: null pointer");
  M3P_REQUIRE(n_groups > 0 && sample_n > 0 && n_groups < (1 << 24) && sample_n <= 4096, "m3p_relation_loss: bad shape");
  const int threads = n_groups >= 1024 ? 1024 : (int)((n_groups + 31) / 32 * 32);
  relation_loss_kernel<<<1, threads, 0, stream>>>(scores, pos_labels, (int)n_groups, (int)sample_n, w_multi, w_bin, loss,
                                                  dscores);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_rowdot_fwd(const void* x, const float* w, const float* bias, float* out, int64_t rows, int64_t d,
                              m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(x && w && bias && out && rows > 0 && d > 0, "m3p_rowdot_fwd: bad arguments");
  rowdot_fwd_kernel<<<(unsigned)((rows + EW_WARPS - 1) / EW_WARPS), EW_THREADS, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), w, bias, out, rows, (int)d);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_rowdot_bwd(const float* dout, const void* x, const float* w, void* dx, float* dw, float* db,
                              int64_t rows, int64_t d, int32_t tanh_grad, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(dout && x && w && dx && dw && db && rows > 0 && d > 0, "m3p_rowdot_bwd: bad arguments");
  rowdot_bwd_kernel<<<dim3((unsigned)((d + EW_THREADS - 1) / EW_THREADS), (unsigned)((rows + 15) / 16)), EW_THREADS, 0, stream>>>(
      dout, reinterpret_cast<const __nv_bfloat16*>(x), w, reinterpret_cast<__nv_bfloat16*>(dx), dw, db, rows, (int)d,
      (int)tanh_grad);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_gather_rows_f32(const float* table, const int64_t* idx, float* dst, int64_t n, int64_t d,
                                   m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(table && idx && dst && n > 0 && d > 0 && d % 4 == 0, "m3p_gather_rows_f32: bad arguments");
  gather_rows_f32_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(table, idx, dst, n, (int)d);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_scatter_add_rows_bf16(const void* src, const int64_t* idx, int64_t skip_index, float* dst, int64_t n,
                                         int64_t d, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(src && idx && dst && n > 0 && d > 0 && d % 8 == 0, "m3p_scatter_add_rows_bf16: bad arguments");
  scatter_add_rows_bf16_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), idx,
                                                                      skip_index, dst, n, (int)d);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_scatter_add_rows_f32(const float* src, const int64_t* idx, int64_t skip_index, float* dst, int64_t n,
                                        int64_t d, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(src && idx && dst && n > 0 && d > 0 && d % 4 == 0, "m3p_scatter_add_rows_f32: bad arguments");
  scatter_add_rows_f32_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(src, idx, skip_index, dst, n, (int)d);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}
