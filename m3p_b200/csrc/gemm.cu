// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   C[m][n] = sum_k A(m,k) * B(n,k)      bf16 operands, fp32 accumulation in TMEM
//
// Roles (384 threads / CTA, one CTA per SM):
//   warp 0   TMA producer: cp.async.bulk.tensor tiles into a STAGES-deep 128B-swizzled smem ring
//   warp 1   MMA issuer  : one elected thread issues tcgen05.mma (128 x BN x 16) and commits
//   warp 2   TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4-11 epilogue  : tcgen05.ld 16 columns at a time (the next chunk in flight), fused math, results staged
//                          in per-warp swizzled shared-memory tiles and written by TMA bulk stores (aux operands
//                          arrive the same way, two staging groups ahead); two warps per TMEM lane quarter, each
//                          taking half of the tile's columns.  The staging-group loop is NOT fully unrolled: the
//                          epilogue has to stay inside the instruction cache (see epilogue_math).
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), and the
// static persistent tile loop, so the epilogue of tile i overlaps the main loop of tile i+1.  The last partial wave
// of tiles may be issued as half-width units (GemmKernelParams::wide_units).
//
// Operand layouts: either operand may be K-major (reduction dim contiguous: forward "TN" GEMMs)
// or MN-major (output dim contiguous: dgrad uses an MN-major B, wgrad MN-major A and B), so no
// transposed copies of weights or activations are ever made.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace m3p {

// Phase timing for kernel bring-up (build with M3P_NVCC_EXTRA=-DM3P_GEMM_TRACE): SM-clock stamps of the producer,
// MMA and one epilogue warp for the first tiles of a few CTAs.  Never compiled into the product build.
#ifdef M3P_GEMM_TRACE
#define GT_DECL long long _gt[24]; int _gn = 0; const long long _g0 = clock64();
#define GT_MARK() do { if (_gn < 24) _gt[_gn++] = clock64() - _g0; } while (0)
#define GT_DUMP(role) do { if (blockIdx.x == 0 || blockIdx.x == 2 || blockIdx.x == 100) { \
    for (int _i = _gn; _i < 24; ++_i) _gt[_i] = 0; \
    printf(role " cta %d n=%d: %lld %lld %lld %lld | %lld %lld %lld %lld | %lld %lld %lld %lld | %lld %lld %lld %lld | %lld %lld %lld %lld | %lld %lld %lld %lld\n", (int)blockIdx.x, _gn, \
      _gt[0], _gt[1], _gt[2], _gt[3], _gt[4], _gt[5], _gt[6], _gt[7], _gt[8], _gt[9], _gt[10], _gt[11], _gt[12], _gt[13], _gt[14], _gt[15], \
      _gt[16], _gt[17], _gt[18], _gt[19], _gt[20], _gt[21], _gt[22], _gt[23]); } } while (0)
#else
#define GT_DECL
#define GT_MARK()
#define GT_DUMP(role)
#endif

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr uint32_t SLAB_BYTES = 64 * BLOCK_K * 2;  // one 64-wide MN slab of an MN-major tile

struct GemmKernelParams {
  int M, N, K;
  int num_n_tiles, num_m_tiles, m_fastest, split_k, num_units;
  // Tail balancing: units [wide_units, num_units) are HALF-width tiles (BN / 2 columns).  When the last wave of full
  // tiles would occupy at most half of the CTA pairs (N = 768: 171 tiles on 74 pairs = 2 waves + 23), those tiles are
  // cut in two so the wave costs half a tile time (3 -> 2.5 tile times).  wide_units == num_units: uniform tiling.
  int wide_units;
  uint32_t idesc_half;
  int kblocks_total, kblocks_per_split;
  int a_mn, b_mn;
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;
  uint32_t idesc;
  float alpha;
  const float* bias;
  void* out;
  long long ldo;
  __nv_bfloat16* out2;
  long long ldo2;
  const __nv_bfloat16* aux;
  long long ldaux;
  float drop_scale;
  uint32_t thr16, seed_lo, seed_hi;
  const uint64_t* seed_mix;
  int accumulate;
  int vec_ok;  // all pitches / bases allow 16-byte vector access
  int tma_store;  // bf16 outputs leave through smem staging + cp.async.bulk.tensor stores
  float* colsum;  // optional [N]: += column sums of the stored (bf16-rounded) output
  long long split_stride;  // fp32 output, split_k > 1, !accumulate: K split ks stores its partial at out + ks * split_stride
  // fp32 residual epilogue, optional: the aux tile holds the PRE-LayerNorm sum of the previous sub-layer and the residual
  // is recomputed here as rowmask * ((aux - mean[row]) * rstd[row] * gamma[col] + beta[col]) — the LayerNorm kernel then
  // only writes its bf16 operand copy, never an fp32 copy of its output
  const float* ln_mean; const float* ln_rstd; const float* ln_gamma; const float* ln_beta;
  const int32_t* ln_seqlen; long long ln_S;
};

// Unit order: K split fastest, then the tile dimension with FEWER tiles.  Units that are adjacent in this order
// run concurrently on neighbouring clusters, so the operand tile they share is fetched from HBM once and hit in
// L2 by the others; the operand indexed by the slower dimension is re-read once per wave.  Making the short
// dimension fast keeps the sharers of the LARGE operand together (FFN lin2 wgrad, 3 x 12 tiles: 214 -> ~134 MB of
// DRAM reads per launch).
// n0 / bn: first column and width of the unit's tile (bn = BN, or BN / 2 for the half-width tail units).
template <int BN>
__device__ __forceinline__ void decode_unit(const GemmKernelParams& p, int u, int& m_tile, int& n0, int& bn, int& ks) {
  if (u >= p.wide_units) {
    // half-width tail (host guarantees split_k == 1 and the N-fastest order): count in half-tile columns
    const int g = 2 * p.wide_units + (u - p.wide_units);
    const int per_row = 2 * p.num_n_tiles;
    m_tile = g / per_row;
    n0 = (g % per_row) * (BN / 2);
    bn = BN / 2;
    ks = 0;
    return;
  }
  ks = u % p.split_k;
  const int t = u / p.split_k;
  int n_tile;
  if (p.m_fastest) {
    m_tile = t % p.num_m_tiles;
    n_tile = t / p.num_m_tiles;
  } else {
    n_tile = t % p.num_n_tiles;
    m_tile = t / p.num_n_tiles;
  }
  n0 = n_tile * BN;
  bn = BN;
}

// ---- epilogue math on one 16-column chunk held by one thread (one output row) -------------------
constexpr int EW = 16;  // epilogue chunk width (columns): one tcgen05.ld.32x32b.x16
constexpr int GW = 32;  // staging group width (columns): one [32 rows][64 B] SWIZZLE_64B tile per warp = one TMA box
constexpr uint32_t STG_TILE = 32 * 64;

__device__ __forceinline__ void unpack_bf16x16(const uint4* t, float* x) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    x[8 * j + 0] = bf16_lo(t[j].x); x[8 * j + 1] = bf16_hi(t[j].x);
    x[8 * j + 2] = bf16_lo(t[j].y); x[8 * j + 3] = bf16_hi(t[j].y);
    x[8 * j + 4] = bf16_lo(t[j].z); x[8 * j + 5] = bf16_hi(t[j].z);
    x[8 * j + 6] = bf16_lo(t[j].w); x[8 * j + 7] = bf16_hi(t[j].w);
  }
}
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* p, const float* v) {
  uint4* o4 = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int j = 0; j < 2; ++j)
    o4[j] = make_uint4(pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}

template <int EPI>
constexpr bool epi_has_aux() { return EPI == M3P_EPI_DROP_RES || EPI == M3P_EPI_DGELU || EPI == M3P_EPI_DTANH; }

// Staging tiles are [32 rows][64 B] in the SWIZZLE_64B layout TMA expects: the 16-byte chunk c16 (0..3) of row r
// lives at r*64 + ((c16 ^ ((r >> 1) & 3)) << 4) — a warp's row-per-lane 16-byte accesses are conflict-free.
__device__ __forceinline__ uint32_t stg_off(int r, int c16) {
  return static_cast<uint32_t>(r) * 64u + (static_cast<uint32_t>(c16 ^ ((r >> 1) & 3)) << 4);
}
__device__ __forceinline__ void stage_bf16x16(uint8_t* stg, int r, int chunk_in_group, const float* v) {
#pragma unroll
  for (int h = 0; h < 2; ++h)
    *reinterpret_cast<uint4*>(stg + stg_off(r, chunk_in_group * 2 + h)) =
        make_uint4(pack_bf16x2(v[8 * h + 0], v[8 * h + 1]), pack_bf16x2(v[8 * h + 2], v[8 * h + 3]),
                   pack_bf16x2(v[8 * h + 4], v[8 * h + 5]), pack_bf16x2(v[8 * h + 6], v[8 * h + 7]));
}
__device__ __forceinline__ void unstage_bf16x16(const uint8_t* stg, int r, int chunk_in_group, float* x) {
  uint4 t[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) t[h] = *reinterpret_cast<const uint4*>(stg + stg_off(r, chunk_in_group * 2 + h));
  unpack_bf16x16(t, x);
}

// fp32 staging tiles ([32 rows][16 floats] = the same 64-byte rows, same swizzle): 4 chunks of 16 bytes per row
__device__ __forceinline__ void stage_f32x16(uint8_t* stg, int r, const float* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<float4*>(stg + stg_off(r, c)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
__device__ __forceinline__ void unstage_f32x16(const uint8_t* stg, int r, float* x) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 t = *reinterpret_cast<const float4*>(stg + stg_off(r, c));
    x[4 * c] = t.x; x[4 * c + 1] = t.y; x[4 * c + 2] = t.z; x[4 * c + 3] = t.w;
  }
}
// fp32 residual stream: M3P_EPI_DROP_RES with an fp32 output takes an fp32 aux (residual) as well and runs through
// the same TMA ring as the bf16 epilogues, 16 columns per staging group instead of 32
template <int EPI, bool OUT_F32>
constexpr bool f32_ring() { return OUT_F32 && EPI == M3P_EPI_DROP_RES; }

// v = epilogue(alpha * acc + bias [, x]) for 16 columns of one row; gq = gelu(.) for M3P_EPI_GELU (v = gelu').
// x: the chunk's aux values (residual / stashed gelu' / tanh output); e0: row-major element index of the chunk's
// first element (dropout counter).
// residual = LayerNorm(aux) recomputed on the fly (see GemmKernelParams::ln_mean): a = rstd, b = -mean * rstd of this
// thread's row (a = 0 for rows removed by the mask, which also kills beta through `on`)
__device__ __forceinline__ void ln_residual16(float* x, const float* sgamma, const float* sbeta, float a, float b, float on) {
#pragma unroll
  for (int j = 0; j < EW / 4; ++j) {
    const float4 g = *reinterpret_cast<const float4*>(sgamma + 4 * j);
    const float4 be = *reinterpret_cast<const float4*>(sbeta + 4 * j);
    x[4 * j + 0] = fmaf(fmaf(x[4 * j + 0], a, b), g.x, be.x * on);
    x[4 * j + 1] = fmaf(fmaf(x[4 * j + 1], a, b), g.y, be.y * on);
    x[4 * j + 2] = fmaf(fmaf(x[4 * j + 2], a, b), g.z, be.z * on);
    x[4 * j + 3] = fmaf(fmaf(x[4 * j + 3], a, b), g.w, be.w * on);
  }
}

// EVEN_E0: the caller guarantees an even element index e0 (true on the TMA path, whose 16-byte pitch rule makes N a
// multiple of 4): the per-element fallback of the dropout generator is then not even compiled — the epilogue of the fp32
// residual GEMM was 139 KB of SASS and lost a third of its issue slots to instruction fetch (ncu: stall_no_inst 33 %).
template <int EPI, bool EVEN_E0 = false>
__device__ __forceinline__ void epilogue_math(const GemmKernelParams& p, const uint32_t* acc, const float* sbias,
                                              const float* x, uint32_t e0, uint32_t seed_lo, uint32_t seed_hi,
                                              float* v, float* gq) {
#pragma unroll
  for (int j = 0; j < EW / 4; ++j) {
    const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * j);  // broadcast LDS.128
    v[4 * j + 0] = fmaf(p.alpha, __uint_as_float(acc[4 * j + 0]), b.x);
    v[4 * j + 1] = fmaf(p.alpha, __uint_as_float(acc[4 * j + 1]), b.y);
    v[4 * j + 2] = fmaf(p.alpha, __uint_as_float(acc[4 * j + 2]), b.z);
    v[4 * j + 3] = fmaf(p.alpha, __uint_as_float(acc[4 * j + 3]), b.w);
  }
  if constexpr (EPI == M3P_EPI_DROP_RES) {
    if (p.thr16 != 0) {
      if (EVEN_E0 || (e0 & 1u) == 0) {
#pragma unroll
        for (int j = 0; j < EW / 2; ++j) {
          const uint32_t h = drop_hash((e0 >> 1) + j, seed_lo, seed_hi);
          v[2 * j] = ((h & 0xffffu) >= p.thr16) ? v[2 * j] * p.drop_scale : 0.f;
          v[2 * j + 1] = ((h >> 16) >= p.thr16) ? v[2 * j + 1] * p.drop_scale : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < EW; ++j)
          v[j] = drop_keep(e0 + j, seed_lo, seed_hi, p.thr16) ? v[j] * p.drop_scale : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] += x[j];
  } else if constexpr (EPI == M3P_EPI_DGELU) {
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] *= x[j];  // x = gelu'(u) stashed by the forward epilogue
  } else if constexpr (EPI == M3P_EPI_DTANH) {
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] *= (1.0f - x[j] * x[j]);
  } else if constexpr (EPI == M3P_EPI_TANH) {
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] = tanhf(v[j]);
  } else if constexpr (EPI == M3P_EPI_GELU) {
    // gq = gelu(v), v = gelu'(v): one erf / exp evaluation serves both, and the backward (M3P_EPI_DGELU)
    // becomes a plain multiply by the stashed derivative.
#pragma unroll
    for (int j = 0; j < EW; ++j) gelu_and_grad(v[j], gq[j], v[j]);
  }
}

// Direct-to-global epilogue of one chunk: fp32 outputs (plain or atomic accumulate) and the bf16 fallback for
// operands whose pitches / bases rule out TMA.
template <int EPI, bool OUT_F32>
__device__ __forceinline__ void epilogue_chunk_direct(const GemmKernelParams& p, const uint32_t* acc, const float* sbias,
                                                      long long row, int col0, int ncols, uint32_t seed_lo,
                                                      uint32_t seed_hi, long long slab_off) {
  float v[EW], gq[EW], x[EW];
  const bool full = (ncols == EW) && p.vec_ok;
  if constexpr (f32_ring<EPI, OUT_F32>()) {
    const float* ap = reinterpret_cast<const float*>(p.aux) + row * p.ldaux + col0;
#pragma unroll
    for (int j = 0; j < EW; ++j) x[j] = (j < ncols) ? ap[j] : 0.f;
    if (p.ln_mean != nullptr) {
      bool on = true;
      if (p.ln_seqlen != nullptr) on = (row % p.ln_S) < p.ln_seqlen[row / p.ln_S];
      const float a = on ? p.ln_rstd[row] : 0.f, b = on ? -p.ln_mean[row] * a : 0.f;
#pragma unroll
      for (int j = 0; j < EW; ++j)
        if (j < ncols) x[j] = on ? fmaf(fmaf(x[j], a, b), p.ln_gamma[col0 + j], p.ln_beta[col0 + j]) : 0.f;
    }
  } else if constexpr (epi_has_aux<EPI>()) {
    const __nv_bfloat16* ap = p.aux + row * p.ldaux + col0;
    if (full) {
      uint4 t[2];
      t[0] = __ldg(reinterpret_cast<const uint4*>(ap));
      t[1] = __ldg(reinterpret_cast<const uint4*>(ap) + 1);
      unpack_bf16x16(t, x);
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j) x[j] = (j < ncols) ? __bfloat162float(ap[j]) : 0.f;
    }
  }
  const uint32_t e0 = static_cast<uint32_t>(row) * static_cast<uint32_t>(p.N) + static_cast<uint32_t>(col0);
  epilogue_math<EPI>(p, acc, sbias, x, e0, seed_lo, seed_hi, v, gq);
  if constexpr (EPI == M3P_EPI_GELU) {
    __nv_bfloat16* gp = p.out2 + row * p.ldo2 + col0;
    if (full) {
      store_bf16x16(gp, gq);
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j)
        if (j < ncols) gp[j] = __float2bfloat16_rn(gq[j]);
    }
  }
  if constexpr (OUT_F32) {
    float* op = reinterpret_cast<float*>(p.out) + slab_off + row * p.ldo + col0;
    if (full) {
      float4* o4 = reinterpret_cast<float4*>(op);
      if (p.accumulate) {
#pragma unroll
        for (int j = 0; j < EW / 4; ++j)
          atomicAdd(o4 + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < EW / 4; ++j)
          o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j) {
        if (j < ncols) {
          if (p.accumulate) atomicAdd(op + j, v[j]);
          else op[j] = v[j];
        }
      }
    }
  } else {
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + col0;
    if (full) {
      store_bf16x16(op, v);
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j)
        if (j < ncols) op[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// CTA2 = true: the kernel runs as clusters of two CTAs issuing cta_group::2 MMAs on a 256 x BN tile;
// each CTA stages its own 128 rows of A and its own BN/2 rows of B (32 KB / k-block instead of 48 KB:
// the 128 x 256 single-CTA tile is bound by the ~64 B/clk L2 -> SM path at ~2/3 of tensor peak).
#ifndef M3P_AUX_RING
#define M3P_AUX_RING 3
#endif
constexpr int AUX_RING_MAX = 6;

template <int BN, bool CTA2, int EPI, bool OUT_F32>
struct GemmCfg {
  static constexpr int BN_LOAD = CTA2 ? BN / 2 : BN;  // B rows this CTA stages per k-block
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BN_LOAD * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
  // per epilogue warp: its slice of the bias (+ gamma and beta of the recomputed residual LayerNorm)
  static constexpr uint32_t BIAS_BYTES = EPI_WARPS * (BN / 2) * 4 * (f32_ring<EPI, OUT_F32>() ? 3 : 1);
  // Staging for the TMA epilogue: per epilogue warp a ring of [32 rows][32 cols] bf16 tiles per output.  Epilogues
  // with an aux operand TMA-LOAD the aux tile into the ring slot two groups ahead, overwrite it in place with the
  // result and TMA-STORE it (ring of 3); the others only store (ring of 2).
  static constexpr bool F32R = f32_ring<EPI, OUT_F32>();
  static constexpr int N_OUT = OUT_F32 ? (F32R ? 1 : 0) : (EPI == M3P_EPI_GELU ? 2 : 1);
  // Ring depth of the aux epilogues: after staging group gg a warp may only refill a slot whose TMA store has
  // finished READING it.  With RING = 3 that is the store issued one group earlier, so lane 0 sat in wait_group.read
  // for most of a store latency at every group and the refill ran just two groups ahead of its use — the epilogue of
  // the residual GEMMs was latency-bound (25 k cycles per 256 x 256 tile against a 6 k-cycle K = 768 main loop).
  // AUX_WAIT = 2 waits for the store issued TWO groups earlier (done long ago) and the deeper ring keeps
  // RING - AUX_WAIT aux tiles in flight per warp.
  static constexpr int RING = epi_has_aux<EPI>() ? M3P_AUX_RING : 2;
  static constexpr int AUX_WAIT = RING >= 4 ? 2 : 1;
  static constexpr int AUX_LEAD = RING - AUX_WAIT;
  static constexpr uint32_t WARP_STG = N_OUT * RING * STG_TILE;
  static constexpr uint32_t STG_BYTES = EPI_WARPS * WARP_STG;
  static constexpr uint32_t BAR_BYTES = 640;
  static constexpr uint32_t BUDGET = 220 * 1024;
  static constexpr int STAGES_FIT = (BUDGET - STG_BYTES - BIAS_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BIAS_BYTES + 1024 /*align*/ + BAR_BYTES;
};

template <int BN, int EPI, bool OUT_F32, bool CTA2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_bh, const __grid_constant__ CUtensorMap tmap_o,
            const __grid_constant__ CUtensorMap tmap_o2, const __grid_constant__ CUtensorMap tmap_aux,
            const GemmKernelParams p) {
  using Cfg = GemmCfg<BN, CTA2, EPI, OUT_F32>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NCTA = CTA2 ? 2 : 1;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array, so every derived pointer keeps its shared-memory
  // provenance and the compiler emits LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stg_smem = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-aligned: STAGE_BYTES is a multiple of 1024
  float* bias_smem = reinterpret_cast<float*>(stg_smem + Cfg::STG_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::STG_BYTES + Cfg::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_bar = tmem_empty + 2;  // [EPI_WARPS][AUX_RING_MAX]: aux tile of a ring slot has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS * AUX_RING_MAX);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  // persistent loop over work units: one unit = one (BLOCK_M * NCTA) x BN output tile (x one K split)
  const int unit0 = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;
  const int unit_stride = CTA2 ? (gridDim.x >> 1) : gridDim.x;

  if (warp_idx == 0 && elect_one()) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (p.wide_units < p.num_units && !p.b_mn) prefetch_tmap(&tmap_bh);
    if (p.tma_store) {
      prefetch_tmap(&tmap_o);
      if constexpr (EPI == M3P_EPI_GELU) prefetch_tmap(&tmap_o2);
      if constexpr (epi_has_aux<EPI>()) prefetch_tmap(&tmap_aux);
    }
  }
  if (warp_idx == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_WARPS * NCTA);  // one arrival per epilogue warp of every CTA of the pair
    }
    for (int i = 0; i < EPI_WARPS * AUX_RING_MAX; ++i) mbar_init(&aux_bar[i], 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    if constexpr (CTA2) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ===================== TMA producer (every CTA stages its own halves) =====================
    if (elect_one()) {
      GT_DECL
      int stage = 0;
      uint32_t phase = 0;
      for (int u = unit0; u < p.num_units; u += unit_stride) {
        GT_MARK();  // producer: tile start
        int m_tile, n_base, bn, ks;
        decode_unit<BN>(p, u, m_tile, n_base, bn, ks);
        const int bn_load = CTA2 ? bn / 2 : bn;  // B rows this CTA stages for this unit
        const int m0 = (m_tile * NCTA + (int)cta_rank) * BLOCK_M;
        const int n0 = n_base + (int)cta_rank * bn_load;
        const uint32_t stage_tx = (Cfg::A_BYTES + static_cast<uint32_t>(bn_load) * BLOCK_K * 2) * NCTA;
        const int kb0 = ks * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // the pair's bytes are all counted on the leader's barrier
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = kb * BLOCK_K;
          auto load = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
            if constexpr (CTA2) tma_load_2d_2sm(dst, tm, &full_bar[stage], c0, c1);
            else tma_load_2d(dst, tm, &full_bar[stage], c0, c1);
          };
          if (!p.a_mn) {
            load(sa, &tmap_a, k0, m0);
          } else {
#pragma unroll
            for (int s = 0; s < BLOCK_M / 64; ++s) load(sa + s * SLAB_BYTES, &tmap_a, m0 + s * 64, k0);
          }
          if (!p.b_mn) {
            load(sb, bn == BN ? &tmap_b : &tmap_bh, k0, n0);  // the half-width box is half as tall
          } else {
#pragma unroll
            for (int s = 0; s < Cfg::BN_LOAD / 64; ++s)
              if (s * 64 < bn_load) load(sb + s * SLAB_BYTES, &tmap_b, n0 + s * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        GT_MARK();  // producer: last k-block of the tile issued
      }
      GT_DUMP("prod");
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (the leader CTA's elected thread) =====================
    if (leader && elect_one()) {
      GT_DECL
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      const uint64_t adesc_base = make_smem_desc(smem_u32(smem), p.a_lbo, p.a_sbo);
      const uint64_t bdesc_base = make_smem_desc(smem_u32(smem) + Cfg::A_BYTES, p.b_lbo, p.b_sbo);
      const uint64_t a_kstep4 = p.a_kstep >> 4, b_kstep4 = p.b_kstep >> 4;
      for (int u = unit0; u < p.num_units; u += unit_stride) {
        int m_tile, n_base, bn, ks;
        decode_unit<BN>(p, u, m_tile, n_base, bn, ks);
        const uint32_t idesc = bn == BN ? p.idesc : p.idesc_half;
        const int kb0 = ks * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        GT_MARK();  // mma: tile start
        mbar_wait(&tmem_empty[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        GT_MARK();  // mma: accumulator stage free
        const uint32_t d_tmem = tmem_base + acc_stage * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // descriptors differ from the stage-0 / k-0 ones only in the (address >> 4) field: one 64-bit
          // add per MMA instead of rebuilding them (the single issuing thread is latency-bound)
          const uint64_t adesc0 = adesc_base + static_cast<uint64_t>((stage * Cfg::STAGE_BYTES) >> 4);
          const uint64_t bdesc0 = bdesc_base + static_cast<uint64_t>((stage * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = adesc0 + static_cast<uint64_t>(k) * a_kstep4;
            const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(k) * b_kstep4;
            if constexpr (CTA2) umma_ss_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // smem slot reusable (in both CTAs) once these MMAs retire
          if constexpr (CTA2) umma_commit_2sm(&empty_bar[stage], 3);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (each CTA drains its own 128 rows)
        if constexpr (CTA2) umma_commit_2sm(&tmem_full[acc_stage], 3);
        else umma_commit(&tmem_full[acc_stage]);
        GT_MARK();  // mma: all MMAs of the tile issued
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
      GT_DUMP("mma ");
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    const int ew = warp_idx - 4;
    const int q = warp_idx & 3;   // TMEM lane quarter this warp may access
    const int half = ew >> 2;     // which half of the tile's columns this warp drains
    constexpr bool F32R = Cfg::F32R;  // fp32 aux + fp32 output through the ring (16 columns per 64-byte tile row)
    constexpr int GWC = F32R ? EW : GW;  // columns per staging group
    constexpr int NG = BN / 2 / GWC;     // staging groups per warp per full-width tile (half of that for a half-width one)
    constexpr int CPG = GWC / EW;        // 16-column chunks per staging group
    constexpr bool HAS_AUX = epi_has_aux<EPI>();
    constexpr int RING = Cfg::RING;
    float* sbias = bias_smem + ew * (BN / 2) * (F32R ? 3 : 1);
    float* sgamma = sbias + BN / 2;   // F32R only
    float* sbeta = sgamma + BN / 2;
    uint8_t* ring = stg_smem + ew * Cfg::WARP_STG;  // slot b, output o at ring + (b * N_OUT + o) * STG_TILE
    uint64_t* my_aux_bar = aux_bar + ew * AUX_RING_MAX;
    const bool use_tma = (!OUT_F32 || F32R) && p.tma_store;
    uint32_t seed_lo = p.seed_lo, seed_hi = p.seed_hi;
    if constexpr (EPI == M3P_EPI_DROP_RES) mix_seed(p.seed_mix, seed_lo, seed_hi);

    // Staging groups are numbered gg = 0, 1, ... over ALL tiles of this warp (NG per tile); the static tile
    // schedule makes the coordinates of any future group known in advance, so the aux tile of group gg + 2
    // (possibly the next tile's) is requested while group gg + 1 is being computed.  The request cursor below
    // walks the same sequence two groups ahead (one unit decode per tile, no per-group divisions).
    long long ax_u = unit0;      // unit of the group the cursor points at
    int ax_g = 0, ax_slot = 0;   // its group index inside the tile and its ring slot
    int ax_col = 0, ax_row0 = 0; // first column of the tile's slice for this warp / first row of its sub-tile
    int ax_ng = NG;              // staging groups of the cursor's tile
    auto ax_decode = [&]() {
      if (ax_u < p.num_units) {
        int m_tile, n_base, bn, ks;
        decode_unit<BN>(p, static_cast<int>(ax_u), m_tile, n_base, bn, ks);
        ax_col = n_base + half * (bn / 2);
        ax_row0 = (m_tile * NCTA + (int)cta_rank) * BLOCK_M + q * 32;
        ax_ng = bn / 2 / GWC;
      }
    };
    auto issue_aux = [&]() {  // lane 0: request the cursor's group, then advance the cursor
      const int col0 = ax_col + ax_g * GWC;
      if (ax_u < p.num_units && col0 < p.N && ax_row0 < p.M) {  // a box entirely outside the tensor is skipped
        mbar_arrive_expect_tx(&my_aux_bar[ax_slot], STG_TILE);
        tma_load_2d(ring + ax_slot * STG_TILE, &tmap_aux, &my_aux_bar[ax_slot], col0, ax_row0);
      }
      if (++ax_slot == RING) ax_slot = 0;
      if (++ax_g == ax_ng) { ax_g = 0; ax_u += unit_stride; ax_decode(); }
    };
    if constexpr (HAS_AUX) {
      if (use_tma && lane == 0) {
        ax_decode();
#pragma unroll
        for (int i = 0; i < Cfg::AUX_LEAD; ++i) issue_aux();
      }
    }
    uint32_t aux_phase = 0;  // bit b: parity the next wait on ring slot b expects
    int gg = 0;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    GT_DECL
    for (int u = unit0; u < p.num_units; u += unit_stride) {
      GT_MARK();  // epi: tile start
      int m_tile, n0, bn, ks;
      decode_unit<BN>(p, u, m_tile, n0, bn, ks);
      const int cbase = half * (bn / 2);  // this warp's half of the tile's columns
      const int nch = bn / 2 / EW;        // 16-column chunks per warp for this tile
      const int ng = bn / 2 / GWC;        // staging groups per warp for this tile
      const int row0 = (m_tile * NCTA + (int)cta_rank) * BLOCK_M + q * 32;  // first row of this warp's sub-tile
      const long long row = static_cast<long long>(row0) + lane;
      const bool row_ok = row < p.M;
      // stage this warp's bias slice (the global loads overlap the wait for the accumulator)
      for (int c = lane * 4; c < bn / 2; c += 128) {
        const int col = n0 + cbase + c;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr) {
          if (p.vec_ok && col + 3 < p.N) {
            b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          } else {
            if (col + 0 < p.N) b.x = __ldg(p.bias + col + 0);
            if (col + 1 < p.N) b.y = __ldg(p.bias + col + 1);
            if (col + 2 < p.N) b.z = __ldg(p.bias + col + 2);
            if (col + 3 < p.N) b.w = __ldg(p.bias + col + 3);
          }
        }
        *reinterpret_cast<float4*>(sbias + c) = b;
        if constexpr (F32R) {
          if (p.ln_mean != nullptr) {  // N is a multiple of 4 and col < N whenever this tile has work (checked on the host)
            const bool in = col + 3 < p.N;
            *reinterpret_cast<float4*>(sgamma + c) = in ? __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(sbeta + c) = in ? __ldg(reinterpret_cast<const float4*>(p.ln_beta + col))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      float ln_a = 0.f, ln_b = 0.f, ln_on = 0.f;  // this thread's row: rstd, -mean * rstd, 1 (all 0 for masked rows)
      if constexpr (F32R) {
        if (p.ln_mean != nullptr && row_ok) {
          bool on = true;
          if (p.ln_seqlen != nullptr) on = (row % p.ln_S) < p.ln_seqlen[row / p.ln_S];
          if (on) {
            ln_a = __ldg(p.ln_rstd + row);
            ln_b = -__ldg(p.ln_mean + row) * ln_a;
            ln_on = 1.f;
          }
        }
      }
      __syncwarp();
      GT_MARK();  // epi: bias staged
      mbar_wait(&tmem_full[acc_stage], acc_phase);
      tc_fence_after();
      GT_MARK();  // epi: accumulator ready
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_stage * BN;
      // software-pipelined drain: the tcgen05.ld of chunk i+1 is in flight while chunk i is processed
      uint32_t acc[2][EW];
      tmem_ld_32x32b_x16(t_base + cbase, acc[0]);
      if (use_tma) {
        // GU groups per loop trip: just enough unrolling to keep the accumulator double-buffer index static
        // (fully unrolled, the eight-group body no longer fitted the instruction cache)
        constexpr int GU = (CPG & 1) ? 2 : 1;
#pragma unroll 1
        for (int gb = 0; gb < ng; gb += GU) {
#pragma unroll
        for (int gu = 0; gu < GU; ++gu) {
          const int g = gb + gu;
          if (g >= ng) break;  // (ng is a multiple of GU; kept for safety)
          const int b = gg % RING;
          const int gcol0 = n0 + cbase + g * GWC;
          const bool valid = gcol0 < p.N && row0 < p.M;
          uint8_t* slot = ring + b * Cfg::N_OUT * STG_TILE;
          if constexpr (HAS_AUX) {
            if (valid) {  // this group's aux tile (requested two groups ago) must have landed in the slot
              mbar_wait(&my_aux_bar[b], (aux_phase >> b) & 1u);
              aux_phase ^= 1u << b;
            }
#ifdef M3P_GEMM_TRACE
            if (gg >= NG && gg < 2 * NG) GT_MARK();  // epi (2nd tile): aux landed
#endif
          } else {
            if (lane == 0) tma_store_wait_read<RING - 1>();  // the slot's previous store (group gg - RING) was read
            __syncwarp();
          }
#pragma unroll
          for (int ci = 0; ci < CPG; ++ci) {
            const int i = g * CPG + ci;
            const int ii = (gu * CPG + ci) & 1;  // == i & 1: gb * CPG is even
            tmem_ld_wait16(acc[ii]);
            if (i + 1 < nch) tmem_ld_32x32b_x16(t_base + cbase + (i + 1) * EW, acc[ii ^ 1]);
            float v[EW], gq[EW], x[EW];
            if constexpr (F32R) {
              unstage_f32x16(slot, lane, x);
              if (p.ln_mean != nullptr) ln_residual16(x, sgamma + i * EW, sbeta + i * EW, ln_a, ln_b, ln_on);
            } else if constexpr (HAS_AUX) {
              unstage_bf16x16(slot, lane, ci, x);
            }
            const uint32_t e0 = static_cast<uint32_t>(row) * static_cast<uint32_t>(p.N) +
                                static_cast<uint32_t>(gcol0 + ci * EW);
            epilogue_math<EPI, true>(p, acc[ii], sbias + i * EW, x, e0, seed_lo, seed_hi, v, gq);
            if constexpr (F32R) stage_f32x16(slot, lane, v);
            else stage_bf16x16(slot, lane, ci, v);  // in place over the aux values this thread just consumed
            if constexpr (EPI == M3P_EPI_GELU) stage_bf16x16(slot + STG_TILE, lane, ci, gq);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (!F32R && p.colsum != nullptr && valid) {
            // column sums of the staged [32 rows][32 cols] tile: lane = (row parity, column pair); the two half-warps
            // read rows of opposite parity (64-byte rows: two consecutive rows cover all 32 banks)
            const int cp = lane & 15, par = lane >> 4;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int r2 = 0; r2 < 16; ++r2) {
              const int r = 2 * r2 + par;
              const uint32_t w = *reinterpret_cast<const uint32_t*>(slot + stg_off(r, cp >> 2) + (cp & 3) * 4);
              if (row0 + r < p.M) { s0 += bf16_lo(w); s1 += bf16_hi(w); }
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            const int col = gcol0 + 2 * cp;
            if (par == 0 && col < p.N) atomicAdd(p.colsum + col, s0);
            if (par == 1 && col + 1 < p.N) atomicAdd(p.colsum + col + 1, s1);
          }
#ifdef M3P_GEMM_TRACE
          if (gg >= NG && gg < 2 * NG) GT_MARK();  // epi (2nd tile): group computed and staged
#endif
          if (lane == 0) {
            if (valid) {
              tma_store_2d(&tmap_o, slot, gcol0, row0);  // rows / columns past the problem are clipped
              if constexpr (EPI == M3P_EPI_GELU) tma_store_2d(&tmap_o2, slot + STG_TILE, gcol0, row0);
            }
            tma_store_commit();  // one bulk group per staging group, valid or not: keeps the wait counts uniform
            if constexpr (HAS_AUX) {
              // all stores but the AUX_WAIT most recent have been read: slot (gg - AUX_WAIT + 1) % RING is free and
              // takes the aux tile of group gg + AUX_LEAD
              tma_store_wait_read<Cfg::AUX_WAIT>();
              issue_aux();
            }
          }
#ifdef M3P_GEMM_TRACE
          if (gg >= NG && gg < 2 * NG) GT_MARK();  // epi (2nd tile): store issued, next aux requested
#endif
          ++gg;
        }
        }
      } else {
#pragma unroll 1
        for (int i0 = 0; i0 < nch; i0 += 2) {
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            const int i = i0 + ii;
            tmem_ld_wait16(acc[ii]);
            if (i + 1 < nch) tmem_ld_32x32b_x16(t_base + cbase + (i + 1) * EW, acc[ii ^ 1]);
            const int col0 = n0 + cbase + i * EW;
            if (row_ok && col0 < p.N)
              epilogue_chunk_direct<EPI, OUT_F32>(p, acc[ii], sbias + i * EW, row, col0, min(EW, p.N - col0), seed_lo,
                                                  seed_hi, static_cast<long long>(ks) * p.split_stride);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      GT_MARK();  // epi: tile drained
      if (lane == 0) {
        if constexpr (CTA2) mbar_arrive_cluster(&tmem_empty[acc_stage], 0);  // the leader's MMA thread waits on it
        else mbar_arrive(&tmem_empty[acc_stage]);
      }
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_read<0>();  // smem must outlive the bulk stores' reads (the writes drain after exit)
#ifdef M3P_GEMM_TRACE
    if (warp_idx == 4 && lane == 0) { GT_MARK(); GT_DUMP("epi "); }
#endif
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  if (warp_idx == 2) {
    if constexpr (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
template <int BN, int EPI, bool OUT_F32, bool CTA2>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbh, const CUtensorMap& to,
                       const CUtensorMap& to2, const CUtensorMap& tx, const GemmKernelParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTA2, EPI, OUT_F32>;
  auto kfn = gemm_kernel<BN, EPI, OUT_F32, CTA2>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    M3P_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(Cfg::SMEM_BYTES)));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  if (CTA2) {
    const int clusters = p.num_units < sm_count() / 2 ? p.num_units : sm_count() / 2;
    cfg.gridDim = dim3(2 * clusters);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(p.num_units < sm_count() ? p.num_units : sm_count());
  }
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  M3P_CUDA_OK(cudaLaunchKernelEx(&cfg, kfn, ta, tb, tbh, to, to2, tx, p));
  return M3P_OK;
}

template <int BN, bool CTA2>
static int dispatch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbh, const CUtensorMap& to,
                        const CUtensorMap& to2, const CUtensorMap& tx, const GemmKernelParams& p, int epi, bool out_f32,
                        cudaStream_t s) {
  if (out_f32) {
    if (epi == M3P_EPI_DROP_RES) return launch_gemm<BN, M3P_EPI_DROP_RES, true, CTA2>(ta, tb, tbh, to, to2, tx, p, s);
    if (epi != M3P_EPI_LINEAR) {
      set_last_error("m3p_gemm_bf16: fp32 output only with M3P_EPI_LINEAR or M3P_EPI_DROP_RES");
      return M3P_ERR_UNSUPPORTED;
    }
    return launch_gemm<BN, M3P_EPI_LINEAR, true, CTA2>(ta, tb, tbh, to, to2, tx, p, s);
  }
  switch (epi) {
    case M3P_EPI_LINEAR: return launch_gemm<BN, M3P_EPI_LINEAR, false, CTA2>(ta, tb, tbh, to, to2, tx, p, s);
    case M3P_EPI_GELU: return launch_gemm<BN, M3P_EPI_GELU, false, CTA2>(ta, tb, tbh, to, to2, tx, p, s);
    case M3P_EPI_DROP_RES: return launch_gemm<BN, M3P_EPI_DROP_RES, false, CTA2>(ta, tb, tbh, to, to2, tx, p, s);
    case M3P_EPI_DGELU: return launch_gemm<BN, M3P_EPI_DGELU, false, CTA2>(ta, tb, tbh, to, to2, tx, p, s);
    case M3P_EPI_TANH: return launch_gemm<BN, M3P_EPI_TANH, false, CTA2>(ta, tb, tbh, to, to2, tx, p, s);
    case M3P_EPI_DTANH: return launch_gemm<BN, M3P_EPI_DTANH, false, CTA2>(ta, tb, tbh, to, to2, tx, p, s);
    default:
      set_last_error("m3p_gemm_bf16: unknown epilogue %d", epi);
      return M3P_ERR_INVALID_ARGUMENT;
  }
}

// CTA pairs (cta_group::2) are the default whenever the problem has more than one 128-row tile;
// M3P_GEMM_2CTA=0 forces the single-CTA kernel (A/B measurements, bring-up).
static bool use_tma_store() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("M3P_GEMM_TMA_STORE");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("M3P_GEMM_2CTA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// M3P_GEMM_TAIL_BALANCE=0 keeps uniform tiles (A/B measurements)
static bool use_tail_balance() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("M3P_GEMM_TAIL_BALANCE");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int gemm_impl(const m3p_gemm_args* a, int a_lbo, int a_sbo, int a_kstep, int b_lbo, int b_sbo,
              int b_kstep, cudaStream_t stream) {
  M3P_REQUIRE(a != nullptr, "m3p_gemm_bf16: null args");
  M3P_REQUIRE(a->a && a->b && a->out, "m3p_gemm_bf16: null operand pointer");
  M3P_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, "m3p_gemm_bf16: empty problem m=%lld n=%lld k=%lld",
              (long long)a->m, (long long)a->n, (long long)a->k);
  M3P_REQUIRE(a->m < (1ll << 31) && a->n < (1ll << 31) && a->k < (1ll << 31),
              "m3p_gemm_bf16: dimension overflow");
  M3P_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "m3p_gemm_bf16: lda/ldb must be multiples of 8");
  M3P_REQUIRE(aligned16(a->a) && aligned16(a->b), "m3p_gemm_bf16: operands must be 16-byte aligned");
  M3P_REQUIRE(a->split_k >= 1, "m3p_gemm_bf16: split_k must be >= 1");
  M3P_REQUIRE(a->split_k == 1 || (a->out_f32 && (a->accumulate || a->split_stride > 0)),
              "m3p_gemm_bf16: split_k > 1 needs an fp32 output that accumulates or per-split slabs (split_stride)");
  M3P_REQUIRE(a->split_stride == 0 || (a->out_f32 && !a->accumulate && a->split_stride % 4 == 0),
              "m3p_gemm_bf16: split_stride needs a non-accumulating fp32 output and a multiple of 4");
  M3P_REQUIRE(!a->accumulate || a->out_f32, "m3p_gemm_bf16: accumulate needs fp32 output");
  M3P_REQUIRE(a->split_k == 1 || a->bias == nullptr, "m3p_gemm_bf16: split_k > 1 cannot take a bias");
  if (a->epilogue == M3P_EPI_GELU) M3P_REQUIRE(a->out2 != nullptr, "m3p_gemm_bf16: GELU needs out2");
  if (a->epilogue == M3P_EPI_DROP_RES || a->epilogue == M3P_EPI_DGELU || a->epilogue == M3P_EPI_DTANH)
    M3P_REQUIRE(a->aux != nullptr, "m3p_gemm_bf16: epilogue %d needs aux", a->epilogue);
  M3P_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "m3p_gemm_bf16: drop_p out of range");
  const bool f32r = a->out_f32 && a->epilogue == M3P_EPI_DROP_RES;  // fp32 residual stream: fp32 aux -> fp32 out
  M3P_REQUIRE((a->aux_f32 != 0) == f32r, "m3p_gemm_bf16: aux_f32 goes with (and only with) M3P_EPI_DROP_RES + out_f32");
  M3P_REQUIRE(!f32r || (!a->accumulate && a->split_k == 1 && a->colsum == nullptr),
              "m3p_gemm_bf16: the fp32 residual epilogue does not accumulate, split K or sum columns");

  const int BN = (a->n > 128) ? 256 : 128;
  const bool cta2 = use_cta_pairs() && a->m > BLOCK_M;
  const int tile_m = cta2 ? 2 * BLOCK_M : BLOCK_M;
  GemmKernelParams p{};
  p.M = (int)a->m; p.N = (int)a->n; p.K = (int)a->k;
  const int num_m_tiles = (p.M + tile_m - 1) / tile_m;
  p.num_n_tiles = (p.N + BN - 1) / BN;
  p.num_m_tiles = num_m_tiles;
  p.m_fastest = num_m_tiles < p.num_n_tiles ? 1 : 0;
  p.kblocks_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  int split = a->split_k < p.kblocks_total ? a->split_k : p.kblocks_total;
  p.kblocks_per_split = (p.kblocks_total + split - 1) / split;
  split = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;  // no empty split
  p.split_k = split;
  const long long units = (long long)num_m_tiles * p.num_n_tiles * split;
  M3P_REQUIRE(units < (1ll << 31), "m3p_gemm_bf16: too many tiles");
  p.num_units = (int)units;
  p.wide_units = p.num_units;
  // Tail balancing (see GemmKernelParams::wide_units): full 256 x 256 tiles on CTA pairs, N-fastest order, no K split,
  // and a last wave that would leave at least half of the pairs idle -> its tiles are issued as two half-width units.
  if (cta2 && BN == 256 && split == 1 && !p.m_fastest && p.N % BN == 0 && use_tail_balance()) {
    const int workers = sm_count() / 2;
    const int rem = p.num_units % workers;
    if (p.num_units > workers && rem > 0 && 2 * rem <= workers) {
      p.wide_units = p.num_units - rem;
      p.num_units = p.wide_units + 2 * rem;
    }
  }
  p.a_mn = a->a_mn_major ? 1 : 0;
  p.b_mn = a->b_mn_major ? 1 : 0;
  // descriptor constants (see ptx.cuh make_smem_desc)
  p.a_lbo = p.a_mn ? BLOCK_K * 128 : 0;
  p.a_sbo = 1024;
  p.a_kstep = p.a_mn ? UMMA_K * 128 : UMMA_K * 2;
  p.b_lbo = p.b_mn ? BLOCK_K * 128 : 0;
  p.b_sbo = 1024;
  p.b_kstep = p.b_mn ? UMMA_K * 128 : UMMA_K * 2;
  if (a_lbo >= 0) p.a_lbo = a_lbo;
  if (a_sbo >= 0) p.a_sbo = a_sbo;
  if (a_kstep >= 0) p.a_kstep = a_kstep;
  if (b_lbo >= 0) p.b_lbo = b_lbo;
  if (b_sbo >= 0) p.b_sbo = b_sbo;
  if (b_kstep >= 0) p.b_kstep = b_kstep;
  p.idesc = make_idesc_bf16(tile_m, BN, p.a_mn, p.b_mn);
  p.idesc_half = make_idesc_bf16(tile_m, BN / 2, p.a_mn, p.b_mn);
  p.alpha = a->alpha;
  p.bias = a->bias;
  p.out = a->out; p.ldo = a->ldo;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2); p.ldo2 = a->ldo2;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); p.ldaux = a->ldaux;
  p.thr16 = (a->epilogue == M3P_EPI_DROP_RES && a->drop_p > 0.f) ? drop_thr16(a->drop_p) : 0;
  p.drop_scale = 1.0f / (1.0f - a->drop_p);
  p.seed_lo = (uint32_t)(a->seed & 0xffffffffu);
  p.seed_hi = (uint32_t)(a->seed >> 32);
  p.seed_mix = seed_mix_ptr();
  p.accumulate = a->accumulate;
  p.colsum = a->colsum;
  p.split_stride = a->split_stride;
  M3P_REQUIRE(a->aux_ln_mean == nullptr || (f32r && a->aux_ln_rstd && a->aux_ln_gamma && a->aux_ln_beta && a->n % 4 == 0 &&
                                            aligned16(a->aux_ln_gamma) && aligned16(a->aux_ln_beta) &&
                                            (a->aux_ln_seqlen == nullptr || a->aux_ln_S > 0)),
              "m3p_gemm_bf16: aux_ln_* needs the fp32 residual epilogue, all four arrays (16-byte aligned gamma / beta), "
              "n %% 4 == 0 and S > 0 with a row mask");
  p.ln_mean = a->aux_ln_mean; p.ln_rstd = a->aux_ln_rstd; p.ln_gamma = a->aux_ln_gamma; p.ln_beta = a->aux_ln_beta;
  p.ln_seqlen = a->aux_ln_seqlen; p.ln_S = a->aux_ln_S;
  {
    const int osz = a->out_f32 ? 4 : 2;
    bool ok = aligned16(a->out) && ((a->ldo * osz) % 16 == 0);
    if (a->bias) ok = ok && aligned16(a->bias);
    if (a->out2) ok = ok && aligned16(a->out2) && (a->ldo2 % 8 == 0);
    if (a->aux) ok = ok && aligned16(a->aux) && (a->ldaux % (f32r ? 4 : 8) == 0);
    p.vec_ok = ok ? 1 : 0;
  }

  CUtensorMap ta, tb;
  int rc;
  if (!p.a_mn) rc = get_tmap_2d_bf16(&ta, a->a, (uint64_t)a->k, (uint64_t)a->m, (uint64_t)a->lda, BLOCK_K, BLOCK_M);
  else         rc = get_tmap_2d_bf16(&ta, a->a, (uint64_t)a->m, (uint64_t)a->k, (uint64_t)a->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!p.b_mn) rc = get_tmap_2d_bf16(&tb, a->b, (uint64_t)a->k, (uint64_t)a->n, (uint64_t)a->ldb, BLOCK_K,
                                     cta2 ? BN / 2 : BN);
  else         rc = get_tmap_2d_bf16(&tb, a->b, (uint64_t)a->n, (uint64_t)a->k, (uint64_t)a->ldb, 64, BLOCK_K);
  if (rc) return rc;
  CUtensorMap tbh = tb;  // K-major B box of a half-width unit (an MN-major B just loads half as many 64-wide slabs)
  if (p.wide_units < p.num_units && !p.b_mn) {
    rc = get_tmap_2d_bf16(&tbh, a->b, (uint64_t)a->k, (uint64_t)a->n, (uint64_t)a->ldb, BLOCK_K, BN / 4);
    if (rc) return rc;
  }

  // bf16 outputs (and the aux operand): [32 rows][32 cols] SWIZZLE_64B boxes for the TMA epilogue
  CUtensorMap to = ta, to2 = ta, tx = ta;
  p.tma_store = ((!a->out_f32 || f32r) && p.vec_ok && use_tma_store()) ? 1 : 0;
  if (p.tma_store && f32r) {
    rc = get_tmap_2d_f32(&to, a->out, (uint64_t)a->n, (uint64_t)a->m, (uint64_t)a->ldo, EW, 32);
    if (rc) return rc;
    rc = get_tmap_2d_f32(&tx, a->aux, (uint64_t)a->n, (uint64_t)a->m, (uint64_t)a->ldaux, EW, 32);
    if (rc) return rc;
  } else if (p.tma_store) {
    rc = get_tmap_2d_bf16(&to, a->out, (uint64_t)a->n, (uint64_t)a->m, (uint64_t)a->ldo, GW, 32);
    if (rc) return rc;
    if (a->epilogue == M3P_EPI_GELU) {
      rc = get_tmap_2d_bf16(&to2, a->out2, (uint64_t)a->n, (uint64_t)a->m, (uint64_t)a->ldo2, GW, 32);
      if (rc) return rc;
    }
    if (a->aux != nullptr) {
      rc = get_tmap_2d_bf16(&tx, a->aux, (uint64_t)a->n, (uint64_t)a->m, (uint64_t)a->ldaux, GW, 32);
      if (rc) return rc;
    }
  }
  if (a->colsum != nullptr && !p.tma_store) {
    set_last_error("m3p_gemm_bf16: colsum needs a bf16 output on the TMA path (16-byte aligned bases and pitches)");
    return M3P_ERR_UNSUPPORTED;
  }
  if (cta2) {
    if (BN == 256) return dispatch_epi<256, true>(ta, tb, tbh, to, to2, tx, p, a->epilogue, a->out_f32 != 0, stream);
    return dispatch_epi<128, true>(ta, tb, tbh, to, to2, tx, p, a->epilogue, a->out_f32 != 0, stream);
  }
  if (BN == 256) return dispatch_epi<256, false>(ta, tb, tbh, to, to2, tx, p, a->epilogue, a->out_f32 != 0, stream);
  return dispatch_epi<128, false>(ta, tb, tbh, to, to2, tx, p, a->epilogue, a->out_f32 != 0, stream);
}

}  // namespace m3p

extern "C" int m3p_gemm_bf16(const m3p_gemm_args* args, m3p_stream_t stream) {
  return m3p::gemm_impl(args, -1, -1, -1, -1, -1, -1, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int m3p_gemm_bf16_debug(const m3p_gemm_args* args, int32_t a_lbo, int32_t a_sbo,
                                   int32_t a_kstep, int32_t b_lbo, int32_t b_sbo, int32_t b_kstep,
                                   m3p_stream_t stream) {
  return m3p::gemm_impl(args, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep,
                        reinterpret_cast<cudaStream_t>(stream));
}
