// Embedding stage of TransformerModel.jointfwd / fwd / crossfwd (transformer.py:897-943, 820-831,
// 1044-1062) as one fused HBM-bound kernel: image rows take the region projection (GEMM output) +
// location projection (K = 5, done as FMAs) -> LayerNorm_img -> dropout; text rows gather the token
// embedding (or a supplied text_embed); both add the position (and language) embedding, apply the
// length mask, layer_norm_emb and the second dropout, and land in the batch-major [B*S][d] bf16
// residual stream.  One warp per row, 16/32-byte vector loads, warp-shuffle mean/variance.
#include "common.cuh"
#include "ptx.cuh"

namespace m3p {

__device__ __forceinline__ void ld8f(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8b(__nv_bfloat16* p, const float* v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                            pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void drop8e(uint32_t e0, uint32_t seed_lo, uint32_t seed_hi, uint32_t thr16, float scale,
                                       float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = drop_hash((e0 >> 1) + j, seed_lo, seed_hi);
    v[2 * j] = ((h & 0xffffu) >= thr16) ? v[2 * j] * scale : 0.f;
    v[2 * j + 1] = ((h >> 16) >= thr16) ? v[2 * j + 1] * scale : 0.f;
  }
}

constexpr int EMB_THREADS = 256;
constexpr int EMB_WARPS = EMB_THREADS / 32;

template <int MAXC>
__global__ void __launch_bounds__(EMB_THREADS)
embed_fwd_kernel(const m3p_embed_args a, const uint32_t thr16, const float scale, const uint64_t* seed_mix) {
  const int lane = threadIdx.x & 31;
  const int d = (int)a.d;
  const int nchunks = d >> 3;
  const long long S = a.R + a.T;
  const long long rows = a.B * S;
  const long long warp0 = (long long)blockIdx.x * EMB_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EMB_WARPS;
  uint32_t si_lo = (uint32_t)(a.seed_img & 0xffffffffu), si_hi = (uint32_t)(a.seed_img >> 32);
  uint32_t se_lo = (uint32_t)(a.seed_emb & 0xffffffffu), se_hi = (uint32_t)(a.seed_emb >> 32);
  if (thr16 != 0) {
    mix_seed(seed_mix, si_lo, si_hi);
    mix_seed(seed_mix, se_lo, se_hi);
  }
  __nv_bfloat16* h0 = reinterpret_cast<__nv_bfloat16*>(a.h0);

  for (long long row = warp0; row < rows; row += nwarps) {
    const long long b = row / S, s = row % S;
    const bool valid = s < a.seqlen[b];
    float y[MAXC][8];
    long long pidx = s;
    if (s < a.R) {
      // ---------------- image row: BertImageEmbeddings (transformer.py:247-269) ----------------
      const long long ir = b * a.R + s;
      float loc[5];
#pragma unroll
      for (int c = 0; c < 5; ++c) loc[c] = a.image_loc[(s * a.B + b) * 5 + c];
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
          float bl[8];
          ld8f(a.e_img + ir * d + ch * 8, y[c]);
          ld8f(a.b_loc + ch * 8, bl);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float* w = a.w_loc + (ch * 8 + j) * 5;
            float t = bl[j];
#pragma unroll
            for (int q = 0; q < 5; ++q) t = fmaf(loc[q], w[q], t);
            y[c][j] += t;
            sum += y[c][j];
          }
          st8f(a.e_img + ir * d + ch * 8, y[c]);  // stash: input of LayerNorm_img
        }
      }
      const float mean = warp_sum(sum) / d;
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float t = y[c][j] - mean; sq += t * t; }
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) / d + a.eps);
      if (lane == 0) { a.img_mean[ir] = mean; a.img_rstd[ir] = rstd; }
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
          float g[8], be[8];
          ld8f(a.ln_img_g + ch * 8, g);
          ld8f(a.ln_img_b + ch * 8, be);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[c][j] = fmaf((y[c][j] - mean) * rstd, g[j], be[j]);
          if (thr16 != 0) drop8e((uint32_t)ir * (uint32_t)d + ch * 8, si_lo, si_hi, thr16, scale, y[c]);
        }
      }
    } else {
      // ---------------- text row: token gather (transformer.py:913) or FreeLB text_embed ----------
      const long long t = s - a.R;
      const float* src;
      if (a.text_embed != nullptr) src = a.text_embed + (b * a.T + t) * d;
      else src = a.tok_emb + a.x[t * a.B + b] * d;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) ld8f(src + ch * 8, y[c]);
      }
      if (a.positions != nullptr) pidx = a.positions[t * a.B + b];
      if (a.langs != nullptr) {
        const float* lsrc = a.lang_emb + a.langs[t * a.B + b] * d;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int ch = lane + 32 * c;
          if (ch < nchunks) {
            float l[8];
            ld8f(lsrc + ch * 8, l);
#pragma unroll
            for (int j = 0; j < 8; ++j) y[c][j] += l[j];
          }
        }
      }
    }
    if (a.flags & M3P_EMB_POS) {
      const float* psrc = a.pos_emb + pidx * d;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
          float pe[8];
          ld8f(psrc + ch * 8, pe);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[c][j] += pe[j];
        }
      }
    }
    if ((a.flags & M3P_EMB_MASK_PRE) && !valid) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) y[c][j] = 0.f;
    }
    if (a.flags & M3P_EMB_LN) {
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
          st8f(a.y_pre + row * d + ch * 8, y[c]);  // stash: input of layer_norm_emb
#pragma unroll
          for (int j = 0; j < 8; ++j) sum += y[c][j];
        }
      }
      const float mean = warp_sum(sum) / d;
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float t = y[c][j] - mean; sq += t * t; }
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) / d + a.eps);
      if (lane == 0) { a.emb_mean[row] = mean; a.emb_rstd[row] = rstd; }
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
          float g[8], be[8];
          ld8f(a.ln_emb_g + ch * 8, g);
          ld8f(a.ln_emb_b + ch * 8, be);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[c][j] = fmaf((y[c][j] - mean) * rstd, g[j], be[j]);
        }
      }
    }
    const bool zero_out = (a.flags & M3P_EMB_MASK_POST) && !valid;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        if ((a.flags & M3P_EMB_DROP2) && thr16 != 0)
          drop8e((uint32_t)row * (uint32_t)d + ch * 8, se_lo, se_hi, thr16, scale, y[c]);
        if (zero_out) {
#pragma unroll
          for (int j = 0; j < 8; ++j) y[c][j] = 0.f;
        }
        st8b(h0 + row * d + ch * 8, y[c]);
        if (a.h0_f32 != nullptr) st8f(a.h0_f32 + row * d + ch * 8, y[c]);
      }
    }
  }
}

// Backward routing of d(y_pre) [B*S][d] fp32 (output of the layer_norm_emb backward):
//   every row : d_pos_emb[pos] += g
//   text row  : d_tok_emb[x] += g (skipping padding_idx, transformer.py:658)  or  d_text_embed = g;
//               d_lang_emb[lang] += g
//   image row : dy_img[b*R + s] = g      (input of the LayerNorm_img backward)
__global__ void __launch_bounds__(EMB_THREADS)
embed_bwd_route_kernel(const m3p_embed_bwd_args a, const uint32_t thr16, const float scale, const uint64_t* seed_mix) {
  const int lane = threadIdx.x & 31;
  uint32_t se_lo = (uint32_t)(a.seed_emb & 0xffffffffu), se_hi = (uint32_t)(a.seed_emb >> 32);
  if (thr16 != 0) mix_seed(seed_mix, se_lo, se_hi);
  const int d = (int)a.d;
  const long long S = a.R + a.T;
  const long long rows = a.B * S;
  const long long warp0 = (long long)blockIdx.x * EMB_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * EMB_WARPS;
  for (long long row = warp0; row < rows; row += nwarps) {
    const long long b = row / S, s = row % S;
    const bool valid = s < a.seqlen[b];
    const float* g = a.dy_pre + row * d;
    float* dtok = nullptr;
    float* dlang = nullptr;
    float* dcopy = nullptr;
    long long pidx = s;
    if (s < a.R) {
      dcopy = a.dy_img + (b * a.R + s) * d;
    } else {
      const long long t = s - a.R;
      if (a.d_text_embed != nullptr) dcopy = a.d_text_embed + (b * a.T + t) * d;
      else if (a.d_tok_emb != nullptr) {
        const long long tok = a.x[t * a.B + b];
        if (tok != a.pad_index) dtok = a.d_tok_emb + tok * d;
      }
      if (a.positions != nullptr) pidx = a.positions[t * a.B + b];
      if (a.langs != nullptr && a.d_lang_emb != nullptr) dlang = a.d_lang_emb + a.langs[t * a.B + b] * d;
    }
    // rows removed by the mask (before the LayerNorm, transformer.py:940, or after the dropout, :831 / :1062)
    // carry no gradient to the embeddings
    const bool live = !((a.flags & (M3P_EMB_MASK_PRE | M3P_EMB_MASK_POST)) && !valid);
    float* dpos = ((a.flags & M3P_EMB_POS) && a.d_pos_emb != nullptr) ? a.d_pos_emb + pidx * d : nullptr;
    for (int c = lane; c < (d >> 2); c += 32) {
      float4 v = *reinterpret_cast<const float4*>(g + c * 4);
      if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (thr16 != 0) {  // the forward's second dropout (same element counter as embed_fwd_kernel)
        const uint32_t e0 = (uint32_t)row * (uint32_t)d + (uint32_t)c * 4u;
        const uint32_t h0 = drop_hash(e0 >> 1, se_lo, se_hi), h1 = drop_hash((e0 >> 1) + 1, se_lo, se_hi);
        v.x = ((h0 & 0xffffu) >= thr16) ? v.x * scale : 0.f;
        v.y = ((h0 >> 16) >= thr16) ? v.y * scale : 0.f;
        v.z = ((h1 & 0xffffu) >= thr16) ? v.z * scale : 0.f;
        v.w = ((h1 >> 16) >= thr16) ? v.w * scale : 0.f;
      }
      if (dcopy) *reinterpret_cast<float4*>(dcopy + c * 4) = v;
      if (live) {
        if (dpos) atomicAdd(reinterpret_cast<float4*>(dpos + c * 4), v);
        if (dtok) atomicAdd(reinterpret_cast<float4*>(dtok + c * 4), v);
        if (dlang) atomicAdd(reinterpret_cast<float4*>(dlang + c * 4), v);
      }
    }
  }
}

// d w_loc[j][c] += sum_rows de[row][j] * loc[row][c]     (image_location_embeddings weight, K = 5)
__global__ void __launch_bounds__(EMB_THREADS)
loc_wgrad_kernel(const __nv_bfloat16* __restrict__ de, const float* __restrict__ image_loc, float* __restrict__ dw,
                 long long B, long long R, int d, long long rows_per_cta) {
  const int j = blockIdx.y * EMB_THREADS + threadIdx.x;
  if (j >= d) return;
  const long long rows = B * R;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  float acc[5] = {0, 0, 0, 0, 0};
#pragma unroll 8
  for (long long ir = r0; ir < r1; ++ir) {  // unrolled: 8 rows' loads in flight (the loop is latency-bound)
    const long long b = ir / R, s = ir % R;
    const float g = __bfloat162float(de[ir * d + j]);
    const float* l = image_loc + (s * B + b) * 5;
#pragma unroll
    for (int c = 0; c < 5; ++c) acc[c] = fmaf(g, l[c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 5; ++c) atomicAdd(dw + j * 5 + c, acc[c]);
}

}  // namespace m3p

using namespace m3p;

extern "C" int m3p_embed_fwd(const m3p_embed_args* a, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(a != nullptr, "m3p_embed_fwd: null args");
  M3P_REQUIRE(a->B > 0 && a->R >= 0 && a->T >= 0 && a->R + a->T > 0, "m3p_embed_fwd: bad shape");
  M3P_REQUIRE(a->d > 0 && a->d % 8 == 0 && a->d <= 1024, "m3p_embed_fwd: d must be a multiple of 8, <= 1024");
  M3P_REQUIRE(a->seqlen && a->h0, "m3p_embed_fwd: seqlen / h0 missing");
  if (a->R > 0)
    M3P_REQUIRE(a->e_img && a->image_loc && a->w_loc && a->b_loc && a->ln_img_g && a->ln_img_b && a->img_mean && a->img_rstd,
                "m3p_embed_fwd: image-stream pointers missing");
  if (a->T > 0) M3P_REQUIRE((a->x && a->tok_emb) || a->text_embed, "m3p_embed_fwd: text-stream pointers missing");
  if (a->flags & M3P_EMB_POS) M3P_REQUIRE(a->pos_emb, "m3p_embed_fwd: pos_emb missing");
  if (a->flags & M3P_EMB_LN)
    M3P_REQUIRE(a->ln_emb_g && a->ln_emb_b && a->y_pre && a->emb_mean && a->emb_rstd, "m3p_embed_fwd: LN pointers missing");
  if (a->langs) M3P_REQUIRE(a->lang_emb, "m3p_embed_fwd: lang_emb missing");
  M3P_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "m3p_embed_fwd: drop_p out of range");
  const uint32_t thr16 = a->drop_p > 0.f ? drop_thr16(a->drop_p) : 0;
  const float scale = 1.0f / (1.0f - a->drop_p);
  const long long rows = a->B * (a->R + a->T);
  long long g = (rows + EMB_WARPS - 1) / EMB_WARPS;
  const long long cap = (long long)sm_count() * 8;
  const int grid = (int)(g < cap ? g : cap);
  if (a->d <= 256) embed_fwd_kernel<1><<<grid, EMB_THREADS, 0, stream>>>(*a, thr16, scale, seed_mix_ptr());
  else if (a->d <= 768) embed_fwd_kernel<3><<<grid, EMB_THREADS, 0, stream>>>(*a, thr16, scale, seed_mix_ptr());
  else embed_fwd_kernel<4><<<grid, EMB_THREADS, 0, stream>>>(*a, thr16, scale, seed_mix_ptr());
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_embed_bwd_route(const m3p_embed_bwd_args* a, m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(a != nullptr && a->dy_pre && a->seqlen, "m3p_embed_bwd_route: null args");
  M3P_REQUIRE(a->B > 0 && a->R >= 0 && a->T >= 0 && a->d > 0 && a->d % 4 == 0, "m3p_embed_bwd_route: bad shape");
  if (a->R > 0) M3P_REQUIRE(a->dy_img, "m3p_embed_bwd_route: dy_img missing");
  if (a->T > 0 && a->d_text_embed == nullptr && a->d_tok_emb != nullptr)
    M3P_REQUIRE(a->x, "m3p_embed_bwd_route: token ids missing");
  const long long rows = a->B * (a->R + a->T);
  long long g = (rows + EMB_WARPS - 1) / EMB_WARPS;
  const long long cap = (long long)sm_count() * 8;
  M3P_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "m3p_embed_bwd_route: drop_p out of range");
  const bool drop = (a->flags & M3P_EMB_DROP2) && !(a->flags & M3P_EMB_LN) && a->drop_p > 0.f;
  embed_bwd_route_kernel<<<(int)(g < cap ? g : cap), EMB_THREADS, 0, stream>>>(*a, drop ? drop_thr16(a->drop_p) : 0u,
                                                                          1.0f / (1.0f - a->drop_p), seed_mix_ptr());
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}

extern "C" int m3p_loc_wgrad(const void* de, const float* image_loc, float* dw_loc, int64_t B, int64_t R, int64_t d,
                             m3p_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3P_REQUIRE(de && image_loc && dw_loc && B > 0 && R > 0 && d > 0, "m3p_loc_wgrad: bad arguments");
  const int gy = (int)((d + EMB_THREADS - 1) / EMB_THREADS);
  const long long rows = B * R;
  int gx = sm_count() * 4 / gy;
  if (gx < 1) gx = 1;
  if (gx > rows) gx = (int)rows;
  const long long rpc = (rows + gx - 1) / gx;
  gx = (int)((rows + rpc - 1) / rpc);
  loc_wgrad_kernel<<<dim3(gx, gy), EMB_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(de), image_loc, dw_loc,
                                                            B, R, (int)d, rpc);
  M3P_CUDA_OK(cudaGetLastError());
  return M3P_OK;
}
