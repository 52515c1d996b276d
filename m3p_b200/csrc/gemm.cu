// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   C[m][n] = sum_k A(m,k) * B(n,k)      bf16 operands, fp32 accumulation in TMEM
//
// Roles (384 threads / CTA, one CTA per SM):
//   warp 0   TMA producer: cp.async.bulk.tensor tiles into a STAGES-deep 128B-swizzled smem ring
//   warp 1   MMA issuer  : one elected thread issues tcgen05.mma (128 x BN x 16) and commits
//   warp 2   TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4-11 epilogue  : tcgen05.ld 32 columns at a time, fused math, vectorised global stores;
//                          two warps per TMEM lane quarter, each taking half of the tile's columns
//                          (the erf-GELU / dropout epilogues are issue-bound with one warp per quarter)
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), and the
// static persistent tile loop, so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Operand layouts: either operand may be K-major (reduction dim contiguous: forward "TN" GEMMs)
// or MN-major (output dim contiguous: dgrad uses an MN-major B, wgrad MN-major A and B), so no
// transposed copies of weights or activations are ever made.
#include "common.cuh"
#include "ptx.cuh"

namespace m3p {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr uint32_t SLAB_BYTES = 64 * BLOCK_K * 2;  // one 64-wide MN slab of an MN-major tile

struct GemmKernelParams {
  int M, N, K;
  int num_n_tiles, split_k, num_units;
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
  int accumulate;
  int vec_ok;  // all pitches / bases allow 16-byte vector access
};

__device__ __forceinline__ void decode_unit(const GemmKernelParams& p, int u, int& m_tile,
                                            int& n_tile, int& ks) {
  ks = u % p.split_k;
  const int t = u / p.split_k;
  n_tile = t % p.num_n_tiles;
  m_tile = t / p.num_n_tiles;
}

// ---- epilogue math on one 16-column chunk held by one thread (one output row) -------------------
// 16 columns = 32 bytes of bf16 per row: every global access of a thread is one full 32-byte sector.
constexpr int EW = 16;  // epilogue chunk width (columns)

__device__ __forceinline__ void load_bf16x16(const __nv_bfloat16* p, float* x) {
  const uint4* a4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint4 t = __ldg(a4 + j);
    x[8 * j + 0] = bf16_lo(t.x); x[8 * j + 1] = bf16_hi(t.x);
    x[8 * j + 2] = bf16_lo(t.y); x[8 * j + 3] = bf16_hi(t.y);
    x[8 * j + 4] = bf16_lo(t.z); x[8 * j + 5] = bf16_hi(t.z);
    x[8 * j + 6] = bf16_lo(t.w); x[8 * j + 7] = bf16_hi(t.w);
  }
}
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* p, const float* v) {
  uint4* o4 = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int j = 0; j < 2; ++j)
    o4[j] = make_uint4(pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}

template <int EPI, bool OUT_F32>
__device__ __forceinline__ void epilogue_chunk(const GemmKernelParams& p, const uint32_t* acc,
                                               long long row, int col0, int ncols) {
  float v[EW];
  const bool full = (ncols == EW) && p.vec_ok;
  // v = alpha * acc + bias
  if (p.bias != nullptr) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int j = 0; j < EW / 4; ++j) {
        const float4 b = __ldg(b4 + j);
        v[4 * j + 0] = fmaf(p.alpha, __uint_as_float(acc[4 * j + 0]), b.x);
        v[4 * j + 1] = fmaf(p.alpha, __uint_as_float(acc[4 * j + 1]), b.y);
        v[4 * j + 2] = fmaf(p.alpha, __uint_as_float(acc[4 * j + 2]), b.z);
        v[4 * j + 3] = fmaf(p.alpha, __uint_as_float(acc[4 * j + 3]), b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j)
        v[j] = fmaf(p.alpha, __uint_as_float(acc[j]), (j < ncols) ? __ldg(p.bias + col0 + j) : 0.f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] = p.alpha * __uint_as_float(acc[j]);
  }

  // auxiliary operand (residual / stashed gelu' / tanh output)
  if constexpr (EPI == M3P_EPI_DROP_RES || EPI == M3P_EPI_DGELU || EPI == M3P_EPI_DTANH) {
    float x[EW];
    const __nv_bfloat16* ap = p.aux + row * p.ldaux + col0;
    if (full) {
      load_bf16x16(ap, x);
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j) x[j] = (j < ncols) ? __bfloat162float(ap[j]) : 0.f;
    }
    if constexpr (EPI == M3P_EPI_DROP_RES) {
      if (p.thr16 != 0) {
        const uint32_t e0 = static_cast<uint32_t>(row) * static_cast<uint32_t>(p.N) +
                            static_cast<uint32_t>(col0);
        if ((e0 & 1u) == 0) {
#pragma unroll
          for (int j = 0; j < EW / 2; ++j) {
            const uint32_t h = drop_hash((e0 >> 1) + j, p.seed_lo, p.seed_hi);
            v[2 * j] = ((h & 0xffffu) >= p.thr16) ? v[2 * j] * p.drop_scale : 0.f;
            v[2 * j + 1] = ((h >> 16) >= p.thr16) ? v[2 * j + 1] * p.drop_scale : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < EW; ++j)
            v[j] = drop_keep(e0 + j, p.seed_lo, p.seed_hi, p.thr16) ? v[j] * p.drop_scale : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < EW; ++j) v[j] += x[j];
    } else if constexpr (EPI == M3P_EPI_DGELU) {
#pragma unroll
      for (int j = 0; j < EW; ++j) v[j] *= x[j];  // x = gelu'(u) stashed by the forward epilogue
    } else {  // DTANH
#pragma unroll
      for (int j = 0; j < EW; ++j) v[j] *= (1.0f - x[j] * x[j]);
    }
  }
  if constexpr (EPI == M3P_EPI_TANH) {
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] = tanhf(v[j]);
  }
  if constexpr (EPI == M3P_EPI_GELU) {
    // out2 = gelu(v), out = gelu'(v): one erf / exp evaluation serves both, and the backward
    // (M3P_EPI_DGELU) becomes a plain multiply by the stashed derivative.
    float gq[EW];
#pragma unroll
    for (int j = 0; j < EW; ++j) {
      float e;
      const float er = erf_as(v[j] * 0.70710678118654752f, &e);
      const float cdf = fmaf(0.5f, er, 0.5f);
      gq[j] = v[j] * cdf;
      v[j] = fmaf(v[j] * e, 0.3989422804014327f, cdf);
    }
    __nv_bfloat16* gp = p.out2 + row * p.ldo2 + col0;
    if (full) {
      store_bf16x16(gp, gq);
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j)
        if (j < ncols) gp[j] = __float2bfloat16_rn(gq[j]);
    }
  }

  // ---- stores ----
  if constexpr (OUT_F32) {
    float* op = reinterpret_cast<float*>(p.out) + row * p.ldo + col0;
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

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BN * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int EPI, bool OUT_F32>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmKernelParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one()) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp_idx == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_WARPS);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        int m_tile, n_tile, ks;
        decode_unit(p, u, m_tile, n_tile, ks);
        const int m0 = m_tile * BLOCK_M, n0 = n_tile * BN;
        const int kb0 = ks * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int s = 0; s < BLOCK_M / 64; ++s)
              tma_load_2d(sa + s * SLAB_BYTES, &tmap_a, &full_bar[stage], m0 + s * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int s = 0; s < BN / 64; ++s)
              tma_load_2d(sb + s * SLAB_BYTES, &tmap_b, &full_bar[stage], n0 + s * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        int m_tile, n_tile, ks;
        decode_unit(p, u, m_tile, n_tile, ks);
        const int kb0 = ks * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        mbar_wait(&tmem_empty[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_stage * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc(a_addr + k * p.a_kstep, p.a_lbo, p.a_sbo);
            const uint64_t bdesc = make_smem_desc(b_addr + k * p.b_kstep, p.b_lbo, p.b_sbo);
            umma_ss(d_tmem, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc_stage]);  // accumulator complete
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    const int q = warp_idx & 3;            // TMEM lane quarter this warp may access
    const int half = (warp_idx - 4) >> 2;  // which half of the tile's columns this warp drains
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      int m_tile, n_tile, ks;
      decode_unit(p, u, m_tile, n_tile, ks);
      const int n0 = n_tile * BN;
      const long long row = static_cast<long long>(m_tile) * BLOCK_M + q * 32 + lane;
      mbar_wait(&tmem_full[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_stage * BN;
      // software-pipelined drain: the tcgen05.ld of chunk i+1 is in flight while chunk i is processed
      constexpr int NCH = BN / 2 / EW;  // chunks per warp (this warp's half of the tile)
      const int cbase = half * (BN / 2);
      uint32_t acc[2][EW];
      tmem_ld_32x32b_x16(t_base + cbase, acc[0]);
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        tmem_ld_wait16(acc[i & 1]);
        if (i + 1 < NCH) tmem_ld_32x32b_x16(t_base + cbase + (i + 1) * EW, acc[(i + 1) & 1]);
        const int col0 = n0 + cbase + i * EW;
        if (row < p.M && col0 < p.N)
          epilogue_chunk<EPI, OUT_F32>(p, acc[i & 1], row, col0, min(EW, p.N - col0));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc_stage]);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
template <int BN, int EPI, bool OUT_F32>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kfn = gemm_kernel<BN, EPI, OUT_F32>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    M3P_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(Cfg::SMEM_BYTES)));
    attr_set = true;
  }
  const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
  kfn<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

template <int BN>
static int dispatch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelParams& p,
                        int epi, bool out_f32, cudaStream_t s) {
  if (out_f32) {
    if (epi != M3P_EPI_LINEAR) {
      set_last_error("m3p_gemm_bf16: fp32 output only with M3P_EPI_LINEAR");
      return M3P_ERR_UNSUPPORTED;
    }
    return launch_gemm<BN, M3P_EPI_LINEAR, true>(ta, tb, p, s);
  }
  switch (epi) {
    case M3P_EPI_LINEAR: return launch_gemm<BN, M3P_EPI_LINEAR, false>(ta, tb, p, s);
    case M3P_EPI_GELU: return launch_gemm<BN, M3P_EPI_GELU, false>(ta, tb, p, s);
    case M3P_EPI_DROP_RES: return launch_gemm<BN, M3P_EPI_DROP_RES, false>(ta, tb, p, s);
    case M3P_EPI_DGELU: return launch_gemm<BN, M3P_EPI_DGELU, false>(ta, tb, p, s);
    case M3P_EPI_TANH: return launch_gemm<BN, M3P_EPI_TANH, false>(ta, tb, p, s);
    case M3P_EPI_DTANH: return launch_gemm<BN, M3P_EPI_DTANH, false>(ta, tb, p, s);
    default:
      set_last_error("m3p_gemm_bf16: unknown epilogue %d", epi);
      return M3P_ERR_INVALID_ARGUMENT;
  }
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
  M3P_REQUIRE(a->split_k == 1 || (a->out_f32 && a->accumulate),
              "m3p_gemm_bf16: split_k > 1 needs fp32 accumulate output");
  M3P_REQUIRE(!a->accumulate || a->out_f32, "m3p_gemm_bf16: accumulate needs fp32 output");
  M3P_REQUIRE(a->split_k == 1 || a->bias == nullptr, "m3p_gemm_bf16: split_k > 1 cannot take a bias");
  if (a->epilogue == M3P_EPI_GELU) M3P_REQUIRE(a->out2 != nullptr, "m3p_gemm_bf16: GELU needs out2");
  if (a->epilogue == M3P_EPI_DROP_RES || a->epilogue == M3P_EPI_DGELU || a->epilogue == M3P_EPI_DTANH)
    M3P_REQUIRE(a->aux != nullptr, "m3p_gemm_bf16: epilogue %d needs aux", a->epilogue);
  M3P_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "m3p_gemm_bf16: drop_p out of range");

  const int BN = (a->n > 128) ? 256 : 128;
  GemmKernelParams p{};
  p.M = (int)a->m; p.N = (int)a->n; p.K = (int)a->k;
  const int num_m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_tiles = (p.N + BN - 1) / BN;
  p.kblocks_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  int split = a->split_k < p.kblocks_total ? a->split_k : p.kblocks_total;
  p.kblocks_per_split = (p.kblocks_total + split - 1) / split;
  split = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;  // no empty split
  p.split_k = split;
  const long long units = (long long)num_m_tiles * p.num_n_tiles * split;
  M3P_REQUIRE(units < (1ll << 31), "m3p_gemm_bf16: too many tiles");
  p.num_units = (int)units;
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
  p.idesc = make_idesc_bf16(BLOCK_M, BN, p.a_mn, p.b_mn);
  p.alpha = a->alpha;
  p.bias = a->bias;
  p.out = a->out; p.ldo = a->ldo;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2); p.ldo2 = a->ldo2;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); p.ldaux = a->ldaux;
  p.thr16 = (a->epilogue == M3P_EPI_DROP_RES && a->drop_p > 0.f) ? drop_thr16(a->drop_p) : 0;
  p.drop_scale = 1.0f / (1.0f - a->drop_p);
  p.seed_lo = (uint32_t)(a->seed & 0xffffffffu);
  p.seed_hi = (uint32_t)(a->seed >> 32);
  p.accumulate = a->accumulate;
  {
    const int osz = a->out_f32 ? 4 : 2;
    bool ok = aligned16(a->out) && ((a->ldo * osz) % 16 == 0);
    if (a->bias) ok = ok && aligned16(a->bias);
    if (a->out2) ok = ok && aligned16(a->out2) && (a->ldo2 % 8 == 0);
    if (a->aux) ok = ok && aligned16(a->aux) && (a->ldaux % 8 == 0);
    p.vec_ok = ok ? 1 : 0;
  }

  CUtensorMap ta, tb;
  int rc;
  if (!p.a_mn) rc = get_tmap_2d_bf16(&ta, a->a, (uint64_t)a->k, (uint64_t)a->m, (uint64_t)a->lda, BLOCK_K, BLOCK_M);
  else         rc = get_tmap_2d_bf16(&ta, a->a, (uint64_t)a->m, (uint64_t)a->k, (uint64_t)a->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!p.b_mn) rc = get_tmap_2d_bf16(&tb, a->b, (uint64_t)a->k, (uint64_t)a->n, (uint64_t)a->ldb, BLOCK_K, BN);
  else         rc = get_tmap_2d_bf16(&tb, a->b, (uint64_t)a->n, (uint64_t)a->k, (uint64_t)a->ldb, 64, BLOCK_K);
  if (rc) return rc;

  if (BN == 256) return dispatch_epi<256>(ta, tb, p, a->epilogue, a->out_f32 != 0, stream);
  return dispatch_epi<128>(ta, tb, p, a->epilogue, a->out_f32 != 0, stream);
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
