// Fused self-attention of MultiHeadAttention.forward (reference transformer.py:149-210, self-attn
// branch) and its backward, on tcgen05 tensor cores.
//
//   scores = (q / sqrt(dh)) k^T ; key-padding mask -> -inf ; softmax in fp32 ; dropout ; ctx = w v
//
// The whole joint sequence (S = 100 regions + 128 tokens = 228 <= 256 keys) fits one CTA, so there
// is no online-softmax rescaling: S = Q K^T lands in TMEM in one shot (128 query rows x <=256 key
// columns of fp32), each thread owns one query row for max / exp2 / sum, P is written as bf16 into
// 128B-swizzled shared memory (exactly the layout TMA would have produced) and fed back to the
// tensor core as the A operand of O = P V.  The score matrix never touches HBM (the reference
// materialises it >= 4 times per layer).
//
// Operands come straight out of the packed QKV projection [B*S][3d] through 3-D TMA maps
// (dim0 = feature, dim1 = position in sequence, dim2 = sequence), so rows beyond S are zero-filled
// by the TMA unit and no padding copies exist.  Head dim is 64 (= one 128-byte swizzle row), which
// holds for M3P-base (768/12) and M3P-large (1024/16).
//
// Forward : grid = B*H*ceil(S/128) CTAs of 128 threads, 96 KB smem, 256 TMEM columns -> 2 CTAs/SM,
//           so one CTA's softmax overlaps the other's MMAs.
// Backward: grid = B*H CTAs of 256 threads; 128x128 (query x key) blocks; S and dP recomputed into
//           TMEM, dQ/dK/dV accumulate in TMEM (512 columns used), P and dS staged through smem
//           and consumed both K-major (dQ) and MN-major (dK, dV) without a transpose.
#include "common.cuh"
#include "ptx.cuh"

namespace m3p {

// Phase timing for kernel bring-up: build with M3P_NVCC_EXTRA=-DM3P_ATTN_TRACE and a few CTAs print the
// SM-clock stamps of their phases (never compiled into the product build).
#ifdef M3P_ATTN_TRACE
#define TRACE_DECL long long _t[16]; int _nt = 0;
#define TRACE_MARK() do { if (threadIdx.x == 0 && _nt < 16) _t[_nt++] = clock64(); } while (0)
#define TRACE_DUMP(name) do { if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 3 || blockIdx.x == 300 || blockIdx.x == 700)) { \
    for (int _i = _nt; _i < 16; ++_i) _t[_i] = _t[_nt - 1]; \
    printf(name " cta %d: %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", (int)blockIdx.x, \
      _t[1]-_t[0], _t[2]-_t[1], _t[3]-_t[2], _t[4]-_t[3], _t[5]-_t[4], _t[6]-_t[5], _t[7]-_t[6], _t[8]-_t[7], _t[9]-_t[8], \
      _t[10]-_t[9], _t[11]-_t[10], _t[12]-_t[11], _t[13]-_t[12], _t[14]-_t[13], _t[15]-_t[14]); } } while (0)
#else
#define TRACE_DECL
#define TRACE_MARK()
#define TRACE_DUMP(name)
#endif

constexpr int ATT_DH = 64;
constexpr int ATT_MAX_S = 256;
constexpr uint32_t TILE16K = 128 * 128;  // [128 rows][128 B]

struct AttnKernelParams {
  int B, S, H, d;
  int n_kv;  // S rounded up to 16: N of the score MMA and K of the PV MMA
  int MT;    // ceil(S / 128)
  float scale, scale_log2;
  const int32_t* seqlen;
  __nv_bfloat16* ctx;
  float* lse;  // [B][H][S], log2 domain: max*scale*log2e + log2(sum)
  const __nv_bfloat16* ctx_in;   // backward: forward output (for delta)
  const __nv_bfloat16* dctx;     // backward
  __nv_bfloat16* dqkv;           // backward
  uint32_t thr16, seed_lo, seed_hi;
  const uint64_t* seed_mix;
  float drop_scale;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of the 16-byte chunk c16 (0..7) of row `row` inside a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz_off(int row, int c16) {
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(c16 ^ (row & 7)) << 4);
}

// Attention-probability dropout: ONE full hash per (sequence, head, query row, 32-key chunk) seeds a 32-bit LCG that
// is advanced once per key pair (1 IMAD) and folded (x ^ x >> 16) into two 16-bit draws — 3 integer ops per pair
// instead of the 7 of a full hash per pair; the softmax / dS math warps are bound by the half-rate integer pipe.
// Forward and backward walk the same chunk in the same order, so they regenerate the same mask.  (Keep rate,
// per-position rates, neighbour / chunk correlations and the drops-per-chunk distribution of this generator were
// checked against Binomial(32, p) on 6.4 M draws: rate 0.90005 for p = 0.1, all correlations < 7e-4.)
__device__ __forceinline__ uint32_t attn_drop_next(uint32_t& x) {
  const uint32_t h = x ^ (x >> 16);
  x = x * 0x2C9277B5u + 0xAC564B05u;
  return h;
}

// dropout keep-scales for 32 consecutive keys starting at key0 (multiple of 32) of query row q
__device__ __forceinline__ uint32_t attn_drop_base(int bh, int q, int key0) {
  return (static_cast<uint32_t>(bh) * ATT_MAX_S + static_cast<uint32_t>(q)) * ATT_MAX_S +
         static_cast<uint32_t>(key0);
}

// =================================================================================================
// forward
// =================================================================================================
constexpr uint32_t FWD_SMEM_TILES = 96 * 1024;  // sQ 16K | sK 32K | pad 16K | sV 32K ; sP aliases first 64K
constexpr uint32_t FWD_SMEM_BYTES = FWD_SMEM_TILES + 2 * 2 * 128 * 4 + 64 + 1024;
constexpr int FWD_MATH_THREADS = 256;
constexpr int FWD_THREADS = FWD_MATH_THREADS + 32;
constexpr uint32_t NBF_MATH = 1, NBF_KBLOCK0 = 2;  // named barriers: math-only sync; P k-block kb ready = 2 + kb

// Warp roles (288 threads): warps 0-7 = softmax, two threads per query row, each owning every other
// 32-key chunk (thread half h takes chunks 2*kb + h); warp 8 = control (TMA loads, MMA issue).  The
// P V product is issued per 64-key block as soon as both threads of every row have written that block
// of P, so it runs on the tensor pipe underneath the exponentials of the later blocks.  O accumulates
// in the TMEM columns of the first 64 scores, which every thread has consumed before block 0 is ready.
__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                const __grid_constant__ CUtensorMap tmap_ctx, const AttnKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array, so every derived pointer keeps its shared-memory
  // provenance and the compiler emits LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16 * 1024;
  uint8_t* sV = smem + 64 * 1024;
  uint8_t* sP = smem;  // written only after the score MMA has consumed sQ / sK
  float* s_max = reinterpret_cast<float*>(smem + FWD_SMEM_TILES);  // [2][128]
  float* s_sum = s_max + 2 * 128;                                  // [2][128]
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(s_sum + 2 * 128);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_o = bar_qk + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 4);

  TRACE_DECL
  TRACE_MARK();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool control = warp == 8;
  const int u = blockIdx.x;
  const int mt = u % p.MT;
  const int h = (u / p.MT) % p.H;
  const int b = u / (p.MT * p.H);

  if (control) {
    if (elect_one()) {
      prefetch_tmap(&tmap_q);
      prefetch_tmap(&tmap_kv);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      fence_barrier_init();
      // the operand loads only need the barriers: they fly during the TMEM allocation and the CTA-wide sync
      mbar_arrive_expect_tx(bar_qk, 48 * 1024);
      tma_load_3d(sQ, &tmap_q, bar_qk, h * ATT_DH, mt * 128, b);
      tma_load_3d(sK, &tmap_kv, bar_qk, p.d + h * ATT_DH, 0, b);
      mbar_arrive_expect_tx(bar_v, 32 * 1024);
      tma_load_3d(sV, &tmap_kv, bar_v, 2 * p.d + h * ATT_DH, 0, b);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nchunk = (p.n_kv + 31) >> 5;
  const int nkb = (p.n_kv + 63) >> 6;  // 64-key blocks of P / V

  if (control) {
    // ================================ control warp ================================
    if (lane == 0) {
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, p.n_kv, 0, 0);
      const uint64_t qd = make_smem_desc(smem_u32(sQ), 0, 1024), kd = make_smem_desc(smem_u32(sK), 0, 1024);
#pragma unroll
      for (int k = 0; k < ATT_DH / 16; ++k) umma_ss(tmem_base, qd + k * 2, kd + k * 2, idesc, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
      mbar_wait(bar_v, 0);
    }
    __syncwarp();
    const uint32_t idesc_pv = make_idesc_bf16(128, ATT_DH, 0, 1);
    const uint64_t pd0 = make_smem_desc(smem_u32(sP), 0, 1024), vd0 = make_smem_desc(smem_u32(sV), 256 * 128, 1024);
    const int nk = p.n_kv >> 4;
    for (int kb = 0; kb < nkb; ++kb) {
      named_bar_sync(NBF_KBLOCK0 + kb, FWD_THREADS);  // P[:, kb*64 .. +64) is in shared memory
      if (lane == 0) {
        tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int kk = kb * 4 + k4;
          if (kk < nk)
            umma_ss(tmem_base, pd0 + static_cast<uint64_t>(kb) * (TILE16K >> 4) + k4 * 2,
                    vd0 + static_cast<uint64_t>(kk) * (2048 >> 4), idesc_pv, kk > 0 ? 1u : 0u);
        }
        if (kb == nkb - 1) umma_commit(bar_o);
      }
      __syncwarp();
    }
  } else {
    // ================================ softmax warps ================================
    const int w4 = warp & 3, half = warp >> 2;
    const int row = w4 * 32 + lane;
    const int q_idx = mt * 128 + row;
    int L = p.seqlen[b];
    L = L < 0 ? 0 : (L > p.S ? p.S : L);
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(w4 * 32) << 16);
    const int nfull = L >> 5;  // chunks with every key valid
    uint32_t seed_lo = p.seed_lo, seed_hi = p.seed_hi;
    if (p.thr16 != 0) mix_seed(p.seed_mix, seed_lo, seed_hi);
    TRACE_MARK();  // setup
    mbar_wait(bar_s, 0);
    __syncwarp();
    tc_fence_after();
    TRACE_MARK();  // S ready

    // Interior 32-key chunks (entirely below the key length L) take a path without per-element masking;
    // only the chunk that straddles L pays for the compares.  The dropout scale 1/(1-p) is folded into
    // the final 1/sum normalisation, so a dropped weight is a plain select-to-zero.
    float mx0 = -INFINITY, mx1 = -INFINITY;
    for (int c = half; c < nchunk; c += 2) {
      if (c * 32 >= L) break;
      uint32_t acc[32];
      tmem_ld_32x32b_x32(t_row + c * 32, acc);
      tmem_ld_wait();
      if (c < nfull) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(acc[j]));
          mx1 = fmaxf(mx1, __uint_as_float(acc[j + 1]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < L) mx0 = fmaxf(mx0, __uint_as_float(acc[j]));
      }
    }
    s_max[half * 128 + row] = fmaxf(mx0, mx1);
    named_bar_sync(NBF_MATH, FWD_MATH_THREADS);
    TRACE_MARK();  // pass 1 (max)
    const float mx = fmaxf(s_max[row], s_max[128 + row]);
    const float mxs = (L > 0) ? mx * p.scale_log2 : 0.f;
    float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
    const int bh = b * p.H + h;
    for (int kb = 0; kb < nkb; ++kb) {
      const int c = 2 * kb + half;
      if (c < nchunk) {
        float pv[32];
        if (c * 32 < L) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(t_row + c * 32, acc);
          tmem_ld_wait();
          if (c < nfull) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              pv[j + 0] = ex2(fmaf(__uint_as_float(acc[j + 0]), p.scale_log2, -mxs));
              pv[j + 1] = ex2(fmaf(__uint_as_float(acc[j + 1]), p.scale_log2, -mxs));
              pv[j + 2] = ex2(fmaf(__uint_as_float(acc[j + 2]), p.scale_log2, -mxs));
              pv[j + 3] = ex2(fmaf(__uint_as_float(acc[j + 3]), p.scale_log2, -mxs));
              sum0 += pv[j + 0]; sum1 += pv[j + 1]; sum2 += pv[j + 2]; sum3 += pv[j + 3];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = ex2(fmaf(__uint_as_float(acc[j]), p.scale_log2, -mxs));
              pv[j] = (c * 32 + j < L) ? e : 0.f;
              sum0 += pv[j];
            }
          }
          if (p.thr16 != 0) {
            uint32_t lcg = drop_hash(attn_drop_base(bh, q_idx, c * 32) >> 1, seed_lo, seed_hi);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t hsh = attn_drop_next(lcg);
              pv[2 * j] = ((hsh & 0xffffu) >= p.thr16) ? pv[2 * j] : 0.f;
              pv[2 * j + 1] = ((hsh >> 16) >= p.thr16) ? pv[2 * j + 1] : 0.f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = 0.f;
        }
        // keys [c*32, c*32+32) live in k-block c/2, 16-byte chunks (c&1)*4 .. +3 of this row
        uint8_t* blk = sP + kb * TILE16K;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 v = make_uint4(pack_bf16x2(pv[8 * g + 0], pv[8 * g + 1]), pack_bf16x2(pv[8 * g + 2], pv[8 * g + 3]),
                                     pack_bf16x2(pv[8 * g + 4], pv[8 * g + 5]), pack_bf16x2(pv[8 * g + 6], pv[8 * g + 7]));
          *reinterpret_cast<uint4*>(blk + swz_off(row, (c & 1) * 4 + g)) = v;
        }
        fence_proxy_async_smem();
      } else {
        // n_kv ends inside this block's first 32 keys: the other half of the block must read as zeros
        uint8_t* blk = sP + kb * TILE16K;
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(blk + swz_off(row, (c & 1) * 4 + g)) = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
      }
      tc_fence_before();
      named_bar_arrive(NBF_KBLOCK0 + kb, FWD_THREADS);
    }
    s_sum[half * 128 + row] = (sum0 + sum1) + (sum2 + sum3);
    named_bar_sync(NBF_MATH, FWD_MATH_THREADS);
    TRACE_MARK();  // pass 2 (exp, P -> smem)
    const float sum = s_sum[row] + s_sum[128 + row];
    mbar_wait(bar_o, 0);
    __syncwarp();
    tc_fence_after();
    TRACE_MARK();  // PV MMA done
    const float inv = sum > 0.f ? p.drop_scale / sum : 0.f;  // dropout's 1/(1-p) folded in here
    uint32_t o0[32];
    tmem_ld_32x32b_x32(t_row + half * 32, o0);  // this thread's 32 of the 64 output columns
    tmem_ld_wait();
    // O tile -> the (now dead) V staging area in the swizzled layout -> one TMA bulk store; rows >= S are clipped
#pragma unroll
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<uint4*>(sV + swz_off(row, half * 4 + g)) =
          make_uint4(pack_bf16x2(__uint_as_float(o0[8 * g + 0]) * inv, __uint_as_float(o0[8 * g + 1]) * inv),
                     pack_bf16x2(__uint_as_float(o0[8 * g + 2]) * inv, __uint_as_float(o0[8 * g + 3]) * inv),
                     pack_bf16x2(__uint_as_float(o0[8 * g + 4]) * inv, __uint_as_float(o0[8 * g + 5]) * inv),
                     pack_bf16x2(__uint_as_float(o0[8 * g + 6]) * inv, __uint_as_float(o0[8 * g + 7]) * inv));
    if (half == 0 && q_idx < p.S)
      p.lse[(static_cast<long long>(bh)) * p.S + q_idx] = (sum > 0.f) ? mxs + log2f(sum) : INFINITY;
    fence_proxy_async_smem();
    named_bar_sync(NBF_MATH, FWD_MATH_THREADS);
    if (threadIdx.x == 0) {
      tma_store_3d(&tmap_ctx, sV, h * ATT_DH, mt * 128, b);
      tma_store_commit();
      tma_store_wait_read<0>();  // shared memory must outlive the bulk store's READS (the writes drain after exit)
    }
  }
  tc_fence_before();
  __syncthreads();
  TRACE_MARK();  // epilogue
  TRACE_DUMP("fwd");
  if (control) tmem_dealloc(tmem_base, 256);
}

// =================================================================================================
// backward
// =================================================================================================
// smem: sQ | sK | sV | sdO (each [256 rows][128 B]) | sP | sdS (each 2 k-blocks of [128 rows][128 B])
constexpr uint32_t BWD_SMEM_TILES = 6 * 32 * 1024;
constexpr uint32_t BWD_SMEM_BYTES = BWD_SMEM_TILES + 2 * ATT_MAX_S * 4 + 64 + 1024;
constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DQ = 256, TM_DK = 384, TM_DV = 448;

__device__ __forceinline__ void store_acc32(__nv_bfloat16* dst, const uint32_t* acc, float mul) {
  uint4* o4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int g = 0; g < 4; ++g)
    o4[g] = make_uint4(pack_bf16x2(__uint_as_float(acc[8 * g + 0]) * mul, __uint_as_float(acc[8 * g + 1]) * mul),
                       pack_bf16x2(__uint_as_float(acc[8 * g + 2]) * mul, __uint_as_float(acc[8 * g + 3]) * mul),
                       pack_bf16x2(__uint_as_float(acc[8 * g + 4]) * mul, __uint_as_float(acc[8 * g + 5]) * mul),
                       pack_bf16x2(__uint_as_float(acc[8 * g + 6]) * mul, __uint_as_float(acc[8 * g + 7]) * mul));
}

// Warp roles (MW * 32 + 32 threads, MW = 8 or 16 math warps): the math warps do the softmax / dS math and the
// epilogues — MW / 4 threads per query row, each owning 128 / (MW / 4) keys of a block; the last warp is the control
// warp — one lane issues every TMA load and every tcgen05.mma, so the math warps never stall behind MMA issue and the
// tensor pipe works on block n's dV / dK / dQ while the math warps are already on block n+1.  Hand-offs:
//   bar_sd (mbarrier, tcgen05.commit)  control -> math : S and dP of the block are in TMEM
//   named barrier 1 (math arrive, control sync)        : S / dP have been read out of TMEM
//   named barrier 2 (math arrive, control sync)        : P / dS of the block are in shared memory
//   bar_g  (mbarrier, tcgen05.commit)  control -> math : the block's dV / dK / dQ MMAs have retired
// MW = 8 is what runs: with 16 math warps (four per scheduler, 96 registers per thread) the block math — ~10 instructions
// per score element, half of them on the half-rate integer pipe — was NOT faster (pipe-throughput-bound, not latency-
// bound) and the kernel lost 10 % to spills and the wider barriers.
constexpr uint32_t NB_TMEM_FREE = 1, NB_SMEM_READY = 2, NB_MATH = 3;

// P and dS of one 32-key chunk (row q, keys key0 .. key0 + 31) from the raw scores / dP values, packed to bf16 pairs
__device__ __forceinline__ void bwd_chunk_math(const AttnKernelParams& p, const uint32_t* sacc, const uint32_t* dacc,
                                               float lse2, float delta, int key0, int L, int bh, int q, uint32_t seed_lo,
                                               uint32_t seed_hi, uint32_t* ppk, uint32_t* dpk) {
  if (key0 >= L) {
#pragma unroll
    for (int g = 0; g < 16; ++g) { ppk[g] = 0u; dpk[g] = 0u; }
    return;
  }
  float pd[32];
  // P (normalised: the forward's log2-sum-exp is subtracted inside the exponent)
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) pd[jj] = ex2(fmaf(__uint_as_float(sacc[jj]), p.scale_log2, -lse2));
  if (key0 + 32 > L) {  // the one chunk that straddles the key length
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) pd[jj] = (key0 + jj < L) ? pd[jj] : 0.f;
  }
  if (p.thr16 != 0) {
    uint32_t lcg = drop_hash(attn_drop_base(bh, q, key0) >> 1, seed_lo, seed_hi);
#pragma unroll
    for (int jp = 0; jp < 16; ++jp) {
      const uint32_t hsh = attn_drop_next(lcg);
      const float k0 = ((hsh & 0xffffu) >= p.thr16) ? p.drop_scale : 0.f;
      const float k1 = ((hsh >> 16) >= p.thr16) ? p.drop_scale : 0.f;
      dpk[jp] = pack_bf16x2(pd[2 * jp] * fmaf(__uint_as_float(dacc[2 * jp]), k0, -delta),
                            pd[2 * jp + 1] * fmaf(__uint_as_float(dacc[2 * jp + 1]), k1, -delta));
      ppk[jp] = pack_bf16x2(pd[2 * jp] * k0, pd[2 * jp + 1] * k1);
    }
  } else {
#pragma unroll
    for (int jp = 0; jp < 16; ++jp) {
      dpk[jp] = pack_bf16x2(pd[2 * jp] * (__uint_as_float(dacc[2 * jp]) - delta),
                            pd[2 * jp + 1] * (__uint_as_float(dacc[2 * jp + 1]) - delta));
      ppk[jp] = pack_bf16x2(pd[2 * jp], pd[2 * jp + 1]);
    }
  }
}

template <int MW>
__global__ void __launch_bounds__(MW * 32 + 32, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_dqkv,
                const AttnKernelParams p) {
  constexpr int MATH_T = MW * 32;       // math threads
  constexpr int THREADS = MATH_T + 32;  // + control warp
  constexpr int CG = MW / 4;            // column groups: threads per query row
  constexpr int NC = 4 / CG;            // 32-key chunks per thread per 128-key block (2 or 1)
  constexpr int OC = ATT_DH / CG;       // columns of a [128][64] gradient tile per thread (32 or 16)
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array, so every derived pointer keeps its shared-memory
  // provenance and the compiler emits LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 32 * 1024;
  uint8_t* sV = smem + 64 * 1024;
  uint8_t* sdO = smem + 96 * 1024;
  uint8_t* sP = smem + 128 * 1024;
  uint8_t* sdS = smem + 160 * 1024;
  float* s_delta = reinterpret_cast<float*>(smem + BWD_SMEM_TILES);
  float* s_lse = s_delta + ATT_MAX_S;
  uint64_t* bar_ld = reinterpret_cast<uint64_t*>(s_lse + ATT_MAX_S);
  uint64_t* bar_sd = bar_ld + 1;
  uint64_t* bar_g = bar_ld + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ld + 3);

  TRACE_DECL
  TRACE_MARK();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H;
  const int b = blockIdx.x / p.H;
  const int bh = blockIdx.x;
  const bool control = warp == MW;

  if (control) {
    if (elect_one()) {
      prefetch_tmap(&tmap_qkv);
      prefetch_tmap(&tmap_do);
      prefetch_tmap(&tmap_o);
      prefetch_tmap(&tmap_dqkv);
      mbar_init(bar_ld, 1);
      mbar_init(bar_sd, 1);
      mbar_init(bar_g, 1);
      fence_barrier_init();
      // Q, K, V, dO and (temporarily, in the P staging buffer) the forward output O: the loads only need the barrier,
      // so they fly during the TMEM allocation and the CTA-wide sync
      mbar_arrive_expect_tx(bar_ld, 5 * 32 * 1024);
      tma_load_3d(sQ, &tmap_qkv, bar_ld, h * ATT_DH, 0, b);
      tma_load_3d(sK, &tmap_qkv, bar_ld, p.d + h * ATT_DH, 0, b);
      tma_load_3d(sV, &tmap_qkv, bar_ld, 2 * p.d + h * ATT_DH, 0, b);
      tma_load_3d(sdO, &tmap_do, bar_ld, h * ATT_DH, 0, b);
      tma_load_3d(sP, &tmap_o, bar_ld, h * ATT_DH, 0, b);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int NT = p.MT;
  const int nblocks = NT * NT;
  int L = p.seqlen[b];
  L = L < 0 ? 0 : (L > p.S ? p.S : L);

  if (control) {
    // ================================ control warp ================================
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);     // S, dP : K-major A and B
    const uint32_t idesc_t = make_idesc_bf16(128, ATT_DH, 1, 1);  // dV, dK: MN-major A and B
    const uint32_t idesc_q = make_idesc_bf16(128, ATT_DH, 0, 1);  // dQ    : K-major A, MN-major B
    // Descriptor bases (address field = tile 0): a tile / k-step offset is one 64-bit add of (bytes >> 4),
    // so the issuing thread spends a few instructions per MMA instead of rebuilding both descriptors.
    constexpr uint64_t T16 = TILE16K >> 4, KS = 2048 >> 4, K32 = 32 >> 4;
    const uint64_t dQk = make_smem_desc(smem_u32(sQ), 0, 1024), dKk = make_smem_desc(smem_u32(sK), 0, 1024);
    const uint64_t dVk = make_smem_desc(smem_u32(sV), 0, 1024), dOk = make_smem_desc(smem_u32(sdO), 0, 1024);
    const uint64_t dSk = make_smem_desc(smem_u32(sdS), 0, 1024);                  // dS as K-major A (dQ)
    const uint64_t dPm = make_smem_desc(smem_u32(sP), TILE16K, 1024);             // P^T  as MN-major A (dV)
    const uint64_t dSm = make_smem_desc(smem_u32(sdS), TILE16K, 1024);            // dS^T as MN-major A (dK)
    const uint64_t dOm = make_smem_desc(smem_u32(sdO), TILE16K, 1024);            // dO as MN-major B (dV)
    const uint64_t dQm = make_smem_desc(smem_u32(sQ), TILE16K, 1024);             // Q  as MN-major B (dK)
    const uint64_t dKm = make_smem_desc(smem_u32(sK), TILE16K, 1024);             // K  as MN-major B (dQ)
    auto issue_s_dp = [&](int i, int j) {
      tc_fence_after();
      const uint64_t qa = dQk + i * T16, ka = dKk + j * T16, da = dOk + i * T16, va = dVk + j * T16;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tmem_base + TM_S, qa + k * K32, ka + k * K32, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tmem_base + TM_DP, da + k * K32, va + k * K32, idesc_s, k > 0 ? 1u : 0u);
      umma_commit(bar_sd);
    };
    if (lane == 0) {
      mbar_wait(bar_ld, 0);
      issue_s_dp(0, 0);
    }
    __syncwarp();
    for (int n = 0; n < nblocks; ++n) {
      const int j = n / NT, i = n % NT;
      named_bar_sync(NB_TMEM_FREE, THREADS);
      if (lane == 0 && n + 1 < nblocks) issue_s_dp((n + 1) % NT, (n + 1) / NT);
      __syncwarp();
      named_bar_sync(NB_SMEM_READY, THREADS);
      if (lane == 0) {
        tc_fence_after();
        const uint64_t doa = dOm + i * T16, qa = dQm + i * T16, ka = dKm + j * T16;
        const uint32_t acc_kv = i > 0 ? 1u : 0u, acc_q = j > 0 ? 1u : 0u;
        // dV_j += Pd^T dO_i ; dK_j += dS^T Q_i      (M = 128 keys, N = 64, K = 128 queries)
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ss(tmem_base + TM_DV, dPm + k * KS, doa + k * KS, idesc_t, k > 0 ? 1u : acc_kv);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ss(tmem_base + TM_DK, dSm + k * KS, qa + k * KS, idesc_t, k > 0 ? 1u : acc_kv);
        // dQ_i += dS K_j                             (M = 128 queries, N = 64, K = 128 keys)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem_base + TM_DQ + i * ATT_DH, dSk + (k >> 2) * T16 + (k & 3) * K32, ka + k * KS, idesc_q,
                  k > 0 ? 1u : acc_q);
        umma_commit(bar_g);
      }
      __syncwarp();
    }
  } else {
    // ================================ math warps ================================
    const int t = threadIdx.x;
    const int w4 = warp & 3, cg = warp >> 2;  // TMEM lane quarter (rows); column group inside a block / gradient tile
    const int row = w4 * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(w4 * 32) << 16);
    uint32_t seed_lo = p.seed_lo, seed_hi = p.seed_hi;
    if (p.thr16 != 0) mix_seed(p.seed_mix, seed_lo, seed_hi);
    if (t < ATT_MAX_S)
      s_lse[t] = (t < p.S) ? p.lse[static_cast<long long>(bh) * p.S + t] : INFINITY;  // phantom rows: exp2(-inf) = 0
    TRACE_MARK();  // setup
    mbar_wait(bar_ld, 0);
    __syncwarp();
    TRACE_MARK();  // loads landed
    // ---- delta = rowsum(dO * O) from the TMA-staged (128B-swizzled) tiles: MATH_T / 256 threads per query row ----
    {
      constexpr int TPR = MATH_T / ATT_MAX_S;  // 1 or 2
      const int r = t / TPR, part = t % TPR;
      const uint8_t* dob = sdO + (r >> 7) * TILE16K;
      const uint8_t* ob = sP + (r >> 7) * TILE16K;
      float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
      for (int g = 0; g < 8 / TPR; ++g) {
        const uint32_t off = swz_off(r & 127, part * (8 / TPR) + g);
        const uint4 a = *reinterpret_cast<const uint4*>(dob + off);
        const uint4 o = *reinterpret_cast<const uint4*>(ob + off);
        dl0 = fmaf(bf16_lo(a.x), bf16_lo(o.x), dl0); dl1 = fmaf(bf16_hi(a.x), bf16_hi(o.x), dl1);
        dl0 = fmaf(bf16_lo(a.y), bf16_lo(o.y), dl0); dl1 = fmaf(bf16_hi(a.y), bf16_hi(o.y), dl1);
        dl0 = fmaf(bf16_lo(a.z), bf16_lo(o.z), dl0); dl1 = fmaf(bf16_hi(a.z), bf16_hi(o.z), dl1);
        dl0 = fmaf(bf16_lo(a.w), bf16_lo(o.w), dl0); dl1 = fmaf(bf16_hi(a.w), bf16_hi(o.w), dl1);
      }
      float dl = dl0 + dl1;
      if constexpr (TPR == 2) dl += __shfl_xor_sync(0xffffffffu, dl, 1);
      if (part == 0) s_delta[r] = dl;
    }
    named_bar_sync(NB_MATH, MATH_T);  // s_delta / s_lse visible; O consumed before P overwrites it
    TRACE_MARK();  // delta
    uint32_t ph_sd = 0, ph_g = 0;
    int g_pending = 0, epi_j = -1;
    // Gradient tiles leave through shared memory + one TMA bulk store per [128 rows][64 cols] tile (a
    // row-per-thread 16-byte global store costs 32 LSU cycles per warp instruction).  Staging space: the K / V
    // tiles of a finished key tile, and at the very end the Q tiles — all dead once the MMAs that read them
    // have retired.  The softmax scale is applied here, not per element.
    auto stage_cols = [&](uint8_t* tile, const uint32_t* acc, float mul) {  // this thread's OC of the tile's 64 columns
#pragma unroll
      for (int g = 0; g < OC / 8; ++g)
        *reinterpret_cast<uint4*>(tile + swz_off(row, cg * (OC / 8) + g)) =
            make_uint4(pack_bf16x2(__uint_as_float(acc[8 * g + 0]) * mul, __uint_as_float(acc[8 * g + 1]) * mul),
                       pack_bf16x2(__uint_as_float(acc[8 * g + 2]) * mul, __uint_as_float(acc[8 * g + 3]) * mul),
                       pack_bf16x2(__uint_as_float(acc[8 * g + 4]) * mul, __uint_as_float(acc[8 * g + 5]) * mul),
                       pack_bf16x2(__uint_as_float(acc[8 * g + 6]) * mul, __uint_as_float(acc[8 * g + 7]) * mul));
    };
    auto drain_tile = [&](uint32_t tcol, uint8_t* tile, float mul) {  // TMEM [128][64] fp32 -> bf16 staging tile
      uint32_t a[OC];
      if constexpr (OC == 32) {
        tmem_ld_32x32b_x32(t_lane + tcol + cg * OC, a);
        tmem_ld_wait32(a);
      } else {
        tmem_ld_32x32b_x16(t_lane + tcol + cg * OC, a);
        tmem_ld_wait16(a);
      }
      stage_cols(tile, a, mul);
    };
    auto store_dk_dv = [&](int j) {
      drain_tile(TM_DK, sK + j * TILE16K, p.scale);
      drain_tile(TM_DV, sV + j * TILE16K, 1.0f);
      fence_proxy_async_smem();
      named_bar_sync(NB_MATH, MATH_T);
      if (t == 0) {
        tma_store_3d(&tmap_dqkv, sK + j * TILE16K, p.d + h * ATT_DH, j * 128, b);
        tma_store_3d(&tmap_dqkv, sV + j * TILE16K, 2 * p.d + h * ATT_DH, j * 128, b);
        tma_store_commit();
      }
    };
    for (int n = 0; n < nblocks; ++n) {
      const int j = n / NT, i = n % NT;
      mbar_wait(bar_sd, ph_sd);
      ph_sd ^= 1;
      __syncwarp();
      tc_fence_after();
      TRACE_MARK();  // S, dP ready

      // ---- P and dS for this thread's row x (128 / CG) keys, kept packed in registers ----
      const int q = i * 128 + row;
      const float lse2 = s_lse[q], delta = s_delta[q];
      uint32_t ppk[NC][16], dpk[NC][16];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col0 = (cg * NC + c) * 32;  // column inside the 128-key tile
        uint32_t sacc[32], dacc[32];
        tmem_ld_32x32b_x32(t_lane + TM_S + col0, sacc);
        tmem_ld_32x32b_x32(t_lane + TM_DP + col0, dacc);
        tmem_ld_wait32(sacc);
        tmem_ld_wait32(dacc);
        if (c == NC - 1) {
          // everything this thread needs from S / dP is in registers: the control warp may start the
          // next block's score MMAs while this chunk's math runs
          tc_fence_before();
          named_bar_arrive(NB_TMEM_FREE, THREADS);
        }
        bwd_chunk_math(p, sacc, dacc, lse2, delta, j * 128 + col0, L, bh, q, seed_lo, seed_hi, ppk[c], dpk[c]);
      }
      TRACE_MARK();  // softmax / dS math
      // the previous block's dV/dK/dQ MMAs still read sP / sdS: wait before overwriting them.  If that block
      // closed a key tile, its dK / dV are final now: drain them here, i.e. AFTER this block's math, so the
      // MMA batch ran underneath the math instead of being waited for (the MMAs that reuse the dK / dV
      // columns are only issued after this thread's next SMEM_READY arrive below).
      if (g_pending) {
        mbar_wait(bar_g, ph_g);
        ph_g ^= 1;
        g_pending = 0;
        __syncwarp();
        if (epi_j >= 0) {
          tc_fence_after();
          store_dk_dv(epi_j);
          epi_j = -1;
          tc_fence_before();
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col0 = (cg * NC + c) * 32;
        uint8_t* pblk = sP + (col0 >> 6) * TILE16K;   // 64-key k-block of the staged P / dS
        uint8_t* dblk = sdS + (col0 >> 6) * TILE16K;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = swz_off(row, ((col0 & 63) >> 3) + g);
          *reinterpret_cast<uint4*>(pblk + off) = make_uint4(ppk[c][4 * g], ppk[c][4 * g + 1], ppk[c][4 * g + 2], ppk[c][4 * g + 3]);
          *reinterpret_cast<uint4*>(dblk + off) = make_uint4(dpk[c][4 * g], dpk[c][4 * g + 1], dpk[c][4 * g + 2], dpk[c][4 * g + 3]);
        }
      }
      fence_proxy_async_smem();
      named_bar_arrive(NB_SMEM_READY, THREADS);
      g_pending = 1;
      if (i == NT - 1) epi_j = j;  // dK_j / dV_j complete once this block's MMAs retire
    }
    // ---- last key tile's dK / dV, then dQ (this bar_g wait covers every MMA) ----
    mbar_wait(bar_g, ph_g);
    __syncwarp();
    tc_fence_after();
    store_dk_dv(epi_j);
    for (int i = 0; i < NT; ++i) drain_tile(TM_DQ + i * ATT_DH, sQ + i * TILE16K, p.scale);
    fence_proxy_async_smem();
    named_bar_sync(NB_MATH, MATH_T);
    if (t == 0) {
      for (int i = 0; i < NT; ++i) tma_store_3d(&tmap_dqkv, sQ + i * TILE16K, h * ATT_DH, i * 128, b);
      tma_store_commit();
      tma_store_wait_read<0>();  // shared memory must outlive the bulk stores' READS (the writes drain after exit)
    }
  }
  tc_fence_before();
  __syncthreads();
  TRACE_MARK();
  TRACE_DUMP("bwd");
  if (control) tmem_dealloc(tmem_base, 512);
}

static int fill_params(const m3p_attn_args* a, AttnKernelParams& p, const char* who) {
  M3P_REQUIRE(a != nullptr, "%s: null args", who);
  M3P_REQUIRE(a->qkv && a->seqlen && a->ctx && a->lse, "%s: null pointer", who);
  M3P_REQUIRE(a->B > 0 && a->H > 0 && a->S > 0, "%s: empty problem", who);
  M3P_REQUIRE(a->S <= ATT_MAX_S, "%s: sequence length %lld > %d is not supported by the single-tile kernel", who,
              (long long)a->S, ATT_MAX_S);
  M3P_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "%s: drop_p out of range", who);
  M3P_REQUIRE(a->B * a->H < (1ll << 22), "%s: too many (sequence, head) pairs", who);
  p.B = (int)a->B; p.S = (int)a->S; p.H = (int)a->H; p.d = (int)a->H * ATT_DH;
  p.n_kv = (p.S + 15) & ~15;
  p.MT = (p.S + 127) / 128;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.seqlen = a->seqlen;
  p.ctx = reinterpret_cast<__nv_bfloat16*>(a->ctx);
  p.ctx_in = reinterpret_cast<const __nv_bfloat16*>(a->ctx);
  p.lse = a->lse;
  p.dctx = reinterpret_cast<const __nv_bfloat16*>(a->dctx);
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(a->dqkv);
  p.thr16 = a->drop_p > 0.f ? drop_thr16(a->drop_p) : 0;
  p.drop_scale = 1.0f / (1.0f - a->drop_p);
  p.seed_lo = (uint32_t)(a->seed & 0xffffffffu);
  p.seed_hi = (uint32_t)(a->seed >> 32);
  p.seed_mix = seed_mix_ptr();
  return M3P_OK;
}

}  // namespace m3p

using namespace m3p;

extern "C" int m3p_attention_fwd(const m3p_attn_args* a, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  AttnKernelParams p{};
  int rc = fill_params(a, p, "m3p_attention_fwd");
  if (rc) return rc;
  CUtensorMap tq, tkv, tctx;
  const uint64_t d3 = 3ull * p.d;
  rc = get_tmap_3d_bf16(&tq, a->qkv, d3, (uint64_t)p.S, (uint64_t)p.B, d3, d3 * p.S, ATT_DH, 128, 1);
  if (rc) return rc;
  rc = get_tmap_3d_bf16(&tkv, a->qkv, d3, (uint64_t)p.S, (uint64_t)p.B, d3, d3 * p.S, ATT_DH, 256, 1);
  if (rc) return rc;
  rc = get_tmap_3d_bf16(&tctx, a->ctx, (uint64_t)p.d, (uint64_t)p.S, (uint64_t)p.B, (uint64_t)p.d, (uint64_t)p.d * p.S,
                        ATT_DH, 128, 1);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    M3P_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM_BYTES));
    attr_set = true;
  }
  attn_fwd_kernel<<<p.B * p.H * p.MT, FWD_THREADS, FWD_SMEM_BYTES, stream>>>(tq, tkv, tctx, p);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_attention_bwd(const m3p_attn_args* a, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  AttnKernelParams p{};
  int rc = fill_params(a, p, "m3p_attention_bwd");
  if (rc) return rc;
  M3P_REQUIRE(a->dctx && a->dqkv, "m3p_attention_bwd: dctx / dqkv missing");
  CUtensorMap tqkv, tdo, to, tdq;
  const uint64_t d3 = 3ull * p.d, d1 = (uint64_t)p.d;
  rc = get_tmap_3d_bf16(&tqkv, a->qkv, d3, (uint64_t)p.S, (uint64_t)p.B, d3, d3 * p.S, ATT_DH, 256, 1);
  if (rc) return rc;
  rc = get_tmap_3d_bf16(&tdo, a->dctx, d1, (uint64_t)p.S, (uint64_t)p.B, d1, d1 * p.S, ATT_DH, 256, 1);
  if (rc) return rc;
  rc = get_tmap_3d_bf16(&to, a->ctx, d1, (uint64_t)p.S, (uint64_t)p.B, d1, d1 * p.S, ATT_DH, 256, 1);
  if (rc) return rc;
  rc = get_tmap_3d_bf16(&tdq, a->dqkv, d3, (uint64_t)p.S, (uint64_t)p.B, d3, d3 * p.S, ATT_DH, 128, 1);
  if (rc) return rc;
  // 8 math warps: the 16-warp instantiation (four threads per query row, 96 registers) was measured slower
  // (105 vs 95 us per launch at B = 64: the block math is bound by pipe throughput, not by dependent-issue latency)
  constexpr int MW = 8;
  static bool attr_set = false;
  if (!attr_set) {
    M3P_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<MW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM_BYTES));
    attr_set = true;
  }
  attn_bwd_kernel<MW><<<p.B * p.H, MW * 32 + 32, BWD_SMEM_BYTES, stream>>>(tqkv, tdo, to, tdq, p);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

// bring-up aid (not part of the public header): resident CTAs per SM of the attention kernels
extern "C" __attribute__((visibility("default"))) int m3p_debug_attn_occupancy(int bwd) {
  int n = -1;
  if (bwd) {
    cudaFuncSetAttribute(attn_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM_BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, attn_bwd_kernel<8>, 8 * 32 + 32, BWD_SMEM_BYTES);
  } else {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM_BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, attn_fwd_kernel, FWD_THREADS, FWD_SMEM_BYTES);
  }
  return n;
}
