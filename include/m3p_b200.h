/*
 * m3p_b200 — C ABI of the B200 (sm_100a) kernels behind the M3P encoder training path.
 *
 * The reference (microsoft/M3P) has no native code and therefore no FFI: its hot path is the
 * PyTorch module M3P/src/model/transformer.py.  Each entry point below replaces the library calls
 * of one block of that module; the file:line it replaces is cited on every declaration.  The
 * Python host (m3p_b200/transformer.py) binds these through ctypes — see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes; every pointer is a DEVICE pointer unless its name ends in _host.
 *   - everything is enqueued on `stream` (a cudaStream_t passed as void*); no implicit sync,
 *     no allocation (the caller owns outputs and stashes).
 *   - returns 0 (M3P_OK) or an error code; m3p_last_error() gives the thread-local message.
 *   - activations are bf16 row-major [rows][features]; parameters enter as bf16 copies for the
 *     tensor-core operands and fp32 for biases / LayerNorm affine; parameter gradients are fp32
 *     and ACCUMULATED (+=) into caller-zeroed buffers.
 *   - token rows are batch-major: row = b * S + s (the reference's (bs, slen, dim) tensor,
 *     transformer.py:929-943).
 *   - threading: one host thread per process drives one GPU (the one-process-per-GPU model of the path).  Entry points
 *     may be called from different threads, but the seed word registered by m3p_set_seed_mix(), the library-owned scratch
 *     buffer and the tensor-map cache are PROCESS-global and unsynchronised: concurrent callers must serialise themselves.
 */
#ifndef M3P_B200_H_
#define M3P_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* m3p_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define M3P_API __attribute__((visibility("default")))
#else
#define M3P_API
#endif

enum {
  M3P_OK = 0,
  M3P_ERR_INVALID_ARGUMENT = 1,
  M3P_ERR_CUDA = 2,
  M3P_ERR_UNSUPPORTED = 3
};

M3P_API int m3p_version(void);
M3P_API const char* m3p_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x (the only target; no fallback). */
M3P_API int m3p_device_check(void);
/* Dropout seeds: every entry point with a `seed` uses seed ^ *device_word (read on the device at kernel
 * start) once a word has been registered here (NULL unregisters).  The launch parameters of a step can
 * then stay constant — e.g. inside a captured CUDA graph — while the caller advances the word between
 * steps to draw fresh masks; forward and backward of one step must see the same value. */
M3P_API int m3p_set_seed_mix(const uint64_t* device_word);

/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05.mma, TMA-staged, fp32 accumulation in TMEM) with fused epilogue.
 *   C[m][n] = sum_k A(m,k) * B(n,k)
 *   a_mn_major = 0: A stored [m][k] (pitch lda);  1: A stored [k][m]
 *   b_mn_major = 0: B stored [n][k] (nn.Linear weight layout, pitch ldb);  1: B stored [k][n]
 * Replaces every nn.Linear / F.linear on the path and their autograd dgrad/wgrad:
 *   q/k/v/out_lin transformer.py:178-181,208; lin1/lin2 :223-225; image_embeddings :259;
 *   heads :104-117, 576-584, 602-606, 1195-1204.
 * Epilogues (v = alpha * acc + bias[n]):
 *   M3P_EPI_LINEAR   out = v                                   (bf16 or fp32; fp32 may accumulate)
 *   M3P_EPI_GELU     out2 = gelu_erf(v), out = gelu_erf'(v) (stash for bwd)   transformer.py:48-56,224
 *   M3P_EPI_DROP_RES out = aux + dropout(v)        (aux = residual; bf16 -> bf16 or, with out_f32 and
 *                    aux_f32, fp32 -> fp32: the residual stream of the encoder)  transformer.py:951-952,956
 *   M3P_EPI_DGELU    out = v * aux                 (aux = the stashed gelu_erf'; backward of :224)
 *   M3P_EPI_TANH     out = tanh(v)                                          transformer.py:556-557
 *   M3P_EPI_DTANH    out = v * (1 - aux^2)         (aux = tanh output; backward of :557)
 * split_k > 1 requires out_f32 = 1 and either accumulate = 1 (partial sums are reduced with red.add: the
 * fp32 sum depends on arrival order) or split_stride > 0 (K split s stores its partial at out + s * split_stride
 * elements; m3p_sum_slabs_bf16 adds the slabs in index order — deterministic, used wherever the sum is rounded
 * to bf16 afterwards).
 * Pitches are in elements; lda, ldb must be multiples of 8 (TMA 16-byte rule).
 * ------------------------------------------------------------------------------------------ */
enum {
  M3P_EPI_LINEAR = 0,
  M3P_EPI_GELU = 1,
  M3P_EPI_DROP_RES = 2,
  M3P_EPI_DGELU = 3,
  M3P_EPI_TANH = 4,
  M3P_EPI_DTANH = 5
};

typedef struct m3p_gemm_args {
  const void* a; /* bf16 */
  const void* b; /* bf16 */
  int64_t m, n, k;
  int64_t lda, ldb;
  int32_t a_mn_major, b_mn_major;
  int32_t epilogue;
  int32_t out_f32;    /* 0: out is bf16, 1: out is fp32 */
  int32_t accumulate; /* fp32 only: out += result (atomic) */
  int32_t split_k;    /* >= 1 */
  float alpha;
  const float* bias; /* [n] fp32 or NULL */
  void* out;
  int64_t ldo;
  void* out2; /* bf16, M3P_EPI_GELU only */
  int64_t ldo2;
  const void* aux; /* bf16 */
  int64_t ldaux;
  float drop_p; /* M3P_EPI_DROP_RES: 0 disables */
  uint64_t seed;
  /* optional [n] fp32: colsum[j] += sum_rows out[row][j] (the bf16-rounded values that are stored), accumulated
   * atomically from the epilogue's staging tiles — the bias gradient of the layer whose output gradient this GEMM
   * produces (e.g. d b1 of the FFN from the lin2 dgrad, transformer.py:223), saving a pass over `out`.
   * bf16 outputs on the TMA path only (16-byte aligned bases and pitches); otherwise M3P_ERR_UNSUPPORTED. */
  float* colsum;
  int64_t split_stride; /* elements between the per-split output slabs (fp32, !accumulate); 0 = none */
  /* fp32 residual stream: M3P_EPI_DROP_RES with out_f32 = 1 reads an fp32 residual (aux_f32 = 1, ldaux in floats)
   * and stores the pre-LayerNorm sum in fp32, so the residual stream is never rounded to bf16
   * (transformer.py:951-952,956 run in fp32 in the reference). */
  int32_t aux_f32;
  /* optional, fp32 residual epilogue only: aux then holds the PRE-LayerNorm sum x of the previous sub-layer and the
   * residual is recomputed in the epilogue as rowmask * ((x - mean[row]) * rstd[row] * gamma[col] + beta[col])
   * (rowmask(row) = (row % S) < seqlen[row / S]; seqlen = NULL: no mask) — the LayerNorm output never has to exist
   * in fp32 (transformer.py:951-953, 956-958: `tensor = layer_norm(tensor + ...)` feeding the next residual add). */
  const float* aux_ln_mean;
  const float* aux_ln_rstd;
  const float* aux_ln_gamma;
  const float* aux_ln_beta;
  const int32_t* aux_ln_seqlen;
  int64_t aux_ln_S;
} m3p_gemm_args;

M3P_API int m3p_gemm_bf16(const m3p_gemm_args* args, m3p_stream_t stream);

/* Bring-up variant: same as m3p_gemm_bf16 but with the UMMA smem-descriptor constants overridden
 * (tests/tools only; lets one GPU session sweep descriptor hypotheses).  Any value < 0 keeps the
 * built-in constant. */
M3P_API int m3p_gemm_bf16_debug(const m3p_gemm_args* args, int32_t a_lbo, int32_t a_sbo, int32_t a_kstep,
                        int32_t b_lbo, int32_t b_sbo, int32_t b_kstep, m3p_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * Fused self-attention (tcgen05): MultiHeadAttention.forward self-attn branch, transformer.py:149-210
 * minus the four projections (those are m3p_gemm_bf16 calls):
 *   scores = scale * q k^T ; scores[:, key >= seqlen[b]] = -inf (:199-200) ; w = softmax_fp32 (:202) ;
 *   w = dropout(w, drop_p) (:203) ; ctx = w v (:204), heads re-interleaved as `unshape` (:176,205).
 * qkv is the packed projection output [B*S][3*d] = [q | k | v], d = H*64 (head dim must be 64),
 * head h in columns h*64..h*64+63 of each third.  S <= 256.  The score matrix never leaves the SM.
 * lse [B][H][S] (fp32, log2 domain) is stashed for the backward, which recomputes the weights and
 * returns dqkv in the same packed layout (dq already includes `scale`).
 * ------------------------------------------------------------------------------------------ */
typedef struct m3p_attn_args {
  const void* qkv;       /* bf16 [B*S][3*d] */
  const int32_t* seqlen; /* [B] number of valid keys of each sequence */
  int64_t B, S, H;
  float scale; /* 1/sqrt(head dim), applied to the scores (== q / sqrt(dh), :197) */
  float drop_p;
  uint64_t seed;
  void* ctx;  /* bf16 [B*S][d]: output of fwd, input of bwd */
  float* lse; /* [B*H*S]: output of fwd, input of bwd */
  const void* dctx; /* bwd: bf16 [B*S][d] */
  void* dqkv;       /* bwd: bf16 [B*S][3*d] */
} m3p_attn_args;
M3P_API int m3p_attention_fwd(const m3p_attn_args* args, m3p_stream_t stream);
M3P_API int m3p_attention_bwd(const m3p_attn_args* args, m3p_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm (eps 1e-12, biased variance, affine) over the last dim of a [rows][d] tensor (bf16, or fp32 with
 * x_f32 = 1: the encoder's pre-LayerNorm sums are kept in fp32).
 * Replaces layer_norm1 / layer_norm2 + the `tensor *= mask` that follows layer_norm2
 * (transformer.py:953,957-958) and BertPredictionHeadTransform.LayerNorm (:605).
 * Row mask: row = b*S + s is valid iff s < seqlen[b]; invalid rows are written as exact zeros.
 * seqlen = NULL disables the mask.  y is the bf16 tensor-core operand copy of the result; y_f32 (optional) the
 * fp32 copy that the next residual add reads.  mean/rstd [rows] are stashed for the backward.
 * ------------------------------------------------------------------------------------------ */
typedef struct m3p_ln_fwd_args {
  const void* x;
  int32_t x_f32;
  const float* gamma;
  const float* beta;
  const int32_t* seqlen;
  int64_t S;
  void* y;      /* bf16 [rows][d] */
  float* y_f32; /* optional fp32 [rows][d] */
  float* mean;
  float* rstd;
  int64_t rows, d;
  float eps;
} m3p_ln_fwd_args;
M3P_API int m3p_layernorm_fwd(const m3p_ln_fwd_args* args, m3p_stream_t stream);

/* Backward of LayerNorm, fused with the pieces that surround it on the path:
 *   dy_eff  = rowmask * dropout_dy(dy)             dropout that FOLLOWED the LN (embeddings :266,943)
 *   dx      = LN backward of dy_eff                (x, mean, rstd from the forward)
 *   dx_drop = dropout_dx(dx)                       dropout that PRECEDED the residual add (:951, :226)
 *   dgamma += sum dy_eff*xhat ; dbeta += sum dy_eff ; dbias += sum_rows dx_drop (bias of the linear
 *   whose output fed the residual add).  Any of dx_drop / dgamma / dbeta / dbias may be NULL.
 * dtype flags select fp32 (1) or bf16 (0) for x, dy, dx; supported: (0,0,0) (1,0,1) (1,1,0) (1,0,0) (1,1,1).
 * The encoder layers run (1,1,1): fp32 pre-LN sums, fp32 residual-gradient chain; dx_drop is the bf16 operand
 * copy of dx (after the dropout mask) that the following dgrad / wgrad GEMMs read. */
typedef struct m3p_ln_bwd_args {
  const void* dy;
  const void* x;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const int32_t* seqlen;
  int64_t S;
  void* dx;
  void* dx_drop; /* bf16 */
  float dx_drop_p;
  uint64_t dx_seed;
  float dy_drop_p;
  uint64_t dy_seed;
  float* dgamma;
  float* dbeta;
  float* dbias;
  int64_t rows, d;
  int32_t x_f32, dy_f32, dx_f32;
  /* optional fp32 scratch of M3P_LN_BWD_SCRATCH_FLOATS(d) floats: with it the row pass also accumulates the column
   * sums (dy, x and dx_drop are read once instead of twice) into per-CTA partials stored here, and the column pass
   * (m3p_layernorm_bwd_cols, possibly on another stream) only adds those partials into dgamma / dbeta / dbias.
   * The buffer must stay untouched between the two calls.  NULL keeps the two independent passes. */
  float* col_scratch;
} m3p_ln_bwd_args;
#define M3P_LN_BWD_SCRATCH_FLOATS(d) (512 * 3 * (d))
M3P_API int m3p_layernorm_bwd(const m3p_ln_bwd_args* args, m3p_stream_t stream);
/* The two passes of m3p_layernorm_bwd on their own, so a caller can put the parameter-gradient column
 * pass (dgamma / dbeta / dbias; reads dy, x, mean, rstd and the dx / dx_drop the row pass wrote) on a
 * second stream underneath the GEMMs that continue the activation-gradient chain. */
M3P_API int m3p_layernorm_bwd_rows(const m3p_ln_bwd_args* args, m3p_stream_t stream);
M3P_API int m3p_layernorm_bwd_cols(const m3p_ln_bwd_args* args, m3p_stream_t stream);

/* out[j] += sum_rows x[row][j], j < n   (bias gradients of q/k/v, lin1 and the MLM projection: autograd of
 * :178-181,223,112).  ld a multiple of 8 that covers n rounded up to 8 (n itself may be ragged, e.g. V = 250 002). */
M3P_API int m3p_colsum_bf16(const void* x, int64_t ld, float* out, int64_t rows, int64_t n, m3p_stream_t stream);

/* out = bf16(scale * in): refreshes the bf16 tensor-core copies of the fp32 master parameters. */
M3P_API int m3p_cast_f32_bf16(const float* in, void* out, int64_t n, float scale, m3p_stream_t stream);
/* out = scale * float(in): a bf16 buffer back into an fp32 one (the data-parallel gradient exchange reduces bf16
 * copies of the flat gradient slices — Apex DDP all-reduces the AMP half-precision gradients, xtrainer.py:77-83). */
M3P_API int m3p_cast_bf16_f32(const void* in, float* out, int64_t n, float scale, m3p_stream_t stream);
/* out[i] = bf16(sum_{s < n_slabs} in[s * slab_stride + i]), slabs added in index order (deterministic reduction
 * of the split-K partials written with m3p_gemm_args.split_stride; d rows of the MLM head, transformer.py:104-117). */
M3P_API int m3p_sum_slabs_bf16(const float* in, int64_t n_slabs, int64_t slab_stride, void* out, int64_t n,
                               m3p_stream_t stream);
/* du = dg * gp, bf16, gp = gelu_erf'(u) as stashed by the M3P_EPI_GELU epilogue: backward of the
 * activation of BertPredictionHeadTransform (transformer.py:603-604; the FFN's GELU backward is fused
 * into a GEMM epilogue instead). */
M3P_API int m3p_gelu_bwd(const void* dg, const void* gp, void* du, int64_t n, m3p_stream_t stream);
/* (A,B,F) fp32 -> (B,A,F) bf16: the reference's sequence-first inputs (x_img (R,bs,2048),
 * transformer.py:895) to batch-major GEMM rows. */
M3P_API int m3p_permute_cast_f32_bf16(const float* in, void* out, int64_t A, int64_t B, int64_t F,
                                      m3p_stream_t stream);

/* Device-side region pipeline (the reference does this on the host per sample,
 * dataset_pretrain.py:258-292,379): regions flagged in zero_mask (B,R) are zeroed, every 2048-d row is L2-normalised
 * (F.normalize, eps 1e-12; normalize = 0 skips it), and the result lands as the encoder's bf16 batch-major operand
 * out (B,R,F).  in is (R,B,F) fp32 raw features.  ori (B,R,F) fp32, optional: the normalised UNMASKED rows — the
 * MRFR regression target (xtrainer.py:2340). */
M3P_API int m3p_region_prep(const float* in, const uint8_t* zero_mask, int32_t normalize, void* out, float* ori,
                            int64_t R, int64_t B, int64_t F, m3p_stream_t stream);

/* dst[i][:] = src[(f / n_inner) * stride_outer + (f % n_inner) * stride_inner ...], f = flat_idx[i]:
 * the boolean-mask row gather of predict() (transformer.py:1206) on a strided (slen, bs, d) view.
 * m3p_scatter_rows_bf16 is its adjoint for unique indices (dst pre-zeroed by the caller). */
M3P_API int m3p_gather_rows_bf16(const void* src, const int64_t* flat_idx, int64_t n_inner, int64_t stride_outer,
                                 int64_t stride_inner, void* dst, int64_t n, int64_t d, m3p_stream_t stream);
M3P_API int m3p_scatter_rows_bf16(const void* src, const int64_t* flat_idx, int64_t n_inner, int64_t stride_outer,
                                  int64_t stride_inner, void* dst, int64_t n, int64_t d, m3p_stream_t stream);

/* F.cross_entropy(logits, y, reduction='mean', ignore_index) over bf16 logits (transformer.py:112 MLM,
 * :581 MRM).  Forward: *loss = mean over non-ignored rows, lse[n] (natural log) and *inv_count =
 * 1/#valid rows are stashed.  Backward: dlogits = g * (softmax - onehot) * inv_count (0 on ignored
 * rows), g = *grad_scale (device scalar: the upstream dloss, read on the device so the host never
 * syncs) or 1 when NULL; dlogits may alias logits (same pitch). */
M3P_API int m3p_cross_entropy_fwd(const void* logits, int64_t ld, const int64_t* y, int64_t n, int64_t V,
                                  int64_t ignore_index, float* loss, float* lse, float* inv_count,
                                  m3p_stream_t stream);
M3P_API int m3p_cross_entropy_bwd(const void* logits, int64_t ld, const int64_t* y, int64_t n, int64_t V,
                                  int64_t ignore_index, const float* lse, const float* inv_count,
                                  const float* grad_scale, void* dlogits, int64_t ldd, m3p_stream_t stream);

/* Masked MSE of the MRFR objective (xtrainer.py:2333-2348): *loss = sum_rows weight[row] * sum_f (pred - target)^2
 * with weight[row] = selected[row] / (n_selected * d) (== F.mse_loss(pred[sel], target[sel]); prepared with the batch,
 * so nothing syncs).  pred bf16 [n][ld], target fp32 [n][d].  Backward: dpred = 2 * weight * g * (pred - target),
 * g = *grad_scale (device scalar) or 1; unselected rows are written as zeros. */
M3P_API int m3p_masked_mse_fwd(const void* pred, int64_t ld, const float* target, const float* weight, int64_t n,
                               int64_t d, float* loss, m3p_stream_t stream);
M3P_API int m3p_masked_mse_bwd(const void* pred, int64_t ld, const float* target, const float* weight,
                               const float* grad_scale, void* dpred, int64_t ldd, int64_t n, int64_t d,
                               m3p_stream_t stream);

/* ITM loss of pretrain_under_step / t2i_step / i2t_step (xtrainer.py:2359-2372, 1917-1942) on the
 * n_groups * sample_n matching scores, value and gradient in one launch:
 *   *loss = w_multi * CE(scores.view(-1, sample_n), pos_labels) + w_bin * BCEWithLogits(scores.view(-1), onehot(pos))
 *   dscores = d loss / d scores (the caller scales it by the upstream gradient). */
M3P_API int m3p_relation_loss(const float* scores, const int64_t* pos_labels, int64_t n_groups, int64_t sample_n,
                              float w_multi, float w_bin, float* loss, float* dscores, m3p_stream_t stream);

/* seq_relationship / seq_relationship2: Linear(d, 1) (transformer.py:713,716,1196,1200) and backward.
 * tanh_grad != 0: x is the BertPooler tanh output (:556-557) and dx is returned w.r.t. the
 * pre-activation, dx = dout * w * (1 - x^2). */
M3P_API int m3p_rowdot_fwd(const void* x, const float* w, const float* bias, float* out, int64_t rows, int64_t d,
                           m3p_stream_t stream);
M3P_API int m3p_rowdot_bwd(const float* dout, const void* x, const float* w, void* dx, float* dw, float* db,
                           int64_t rows, int64_t d, int32_t tanh_grad, m3p_stream_t stream);

/* dst[i][:] = table[idx[i]][:] (fp32): the nn.Embedding lookup on its own (model.embeddings(ids), used by FreeLB's
 * embeds_init, xtrainer.py:2700-2705); m3p_scatter_add_rows_f32 with skip_index = padding_idx is its backward. */
M3P_API int m3p_gather_rows_f32(const float* table, const int64_t* idx, float* dst, int64_t n, int64_t d,
                                m3p_stream_t stream);
/* dst[idx[i]][:] += src[i][:] (fp32, atomic), rows with idx == skip_index skipped. */
M3P_API int m3p_scatter_add_rows_f32(const float* src, const int64_t* idx, int64_t skip_index, float* dst, int64_t n,
                                     int64_t d, m3p_stream_t stream);
/* the same with bf16 source rows (all-gathered embedding-gradient rows of the data-parallel exchange). */
M3P_API int m3p_scatter_add_rows_bf16(const void* src, const int64_t* idx, int64_t skip_index, float* dst, int64_t n,
                                      int64_t d, m3p_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Embedding stage of jointfwd / fwd / crossfwd (transformer.py:897-943, 820-831, 1044-1062).
 * Rows are batch-major, row = b*(R+T) + s, image rows first (s < R) as torch.cat at :929.
 *   image row: e = e_img (x_img W_img^T + b_img, from m3p_gemm_bf16) + loc W_loc^T + b_loc ;
 *              y = dropout(LN_img(e))                                         (:259-268)
 *   text row : y = tok_emb[x] or text_embed                                    (:910-913)
 *   then     : y += pos_emb[pos] (M3P_EMB_POS) ; += lang_emb[lang] (langs != NULL, text rows)
 *              y *= mask (M3P_EMB_MASK_PRE, jointfwd order :940) ; y_pre = y
 *              y = LN_emb(y) (M3P_EMB_LN) ; y = dropout(y) (M3P_EMB_DROP2) ;
 *              y *= mask (M3P_EMB_MASK_POST, fwd/crossfwd order :831,1062) ; h0 = bf16(y)
 * mask(b,s) = s < seqlen[b].  e_img is updated in place to the LN_img input (stash).
 * ------------------------------------------------------------------------------------------ */
enum { M3P_EMB_POS = 1, M3P_EMB_LN = 2, M3P_EMB_MASK_PRE = 4, M3P_EMB_MASK_POST = 8, M3P_EMB_DROP2 = 16 };

typedef struct m3p_embed_args {
  int64_t B, R, T, d;
  int32_t flags;
  float eps;
  float drop_p;
  uint64_t seed_img, seed_emb;
  /* image stream */
  float* e_img;           /* [B*R][d] fp32 in/out */
  const float* image_loc; /* (R,B,5) fp32 */
  const float* w_loc;     /* [d][5] */
  const float* b_loc;     /* [d] */
  const float* ln_img_g;
  const float* ln_img_b;
  float* img_mean;
  float* img_rstd;
  /* text stream */
  const int64_t* x;        /* (T,B) token ids */
  const float* tok_emb;    /* [V][d] fp32 */
  const float* text_embed; /* optional [B][T][d] fp32 */
  const int64_t* positions; /* optional (T,B) */
  const float* pos_emb;
  const int64_t* langs; /* optional (T,B) */
  const float* lang_emb;
  const int32_t* seqlen; /* [B] */
  const float* ln_emb_g;
  const float* ln_emb_b;
  /* outputs */
  float* y_pre; /* [B*S][d] fp32 (M3P_EMB_LN) */
  float* emb_mean;
  float* emb_rstd;
  void* h0; /* [B*S][d] bf16 */
  float* h0_f32; /* optional [B*S][d] fp32 copy of h0 before rounding: the residual the first layer adds to */
} m3p_embed_args;
M3P_API int m3p_embed_fwd(const m3p_embed_args* args, m3p_stream_t stream);

/* Routing half of the embedding backward: dy_pre [B*S][d] fp32 (from m3p_layernorm_bwd on
 * layer_norm_emb) is scattered to d_pos_emb / d_tok_emb (padding_idx skipped, :658) / d_lang_emb
 * (+=, atomic) or copied to d_text_embed (FreeLB) and, for image rows, to dy_img [B*R][d] fp32. */
typedef struct m3p_embed_bwd_args {
  int64_t B, R, T, d;
  int32_t flags;
  const float* dy_pre;
  const int32_t* seqlen;
  const int64_t* x;
  const int64_t* positions;
  const int64_t* langs;
  int64_t pad_index;
  float* d_tok_emb;
  float* d_text_embed;
  float* d_pos_emb;
  float* d_lang_emb;
  float* dy_img;
  /* M3P_EMB_DROP2 without M3P_EMB_LN (crossfwd image stream, transformer.py:1044-1049: a second dropout straight
   * after BertImageEmbeddings, no layer_norm_emb): its mask is re-applied to dy_pre here.  With M3P_EMB_LN the
   * dropout sits behind the LayerNorm and m3p_layernorm_bwd(dy_drop_p) handles it. */
  float drop_p;
  uint64_t seed_emb;
} m3p_embed_bwd_args;
M3P_API int m3p_embed_bwd_route(const m3p_embed_bwd_args* args, m3p_stream_t stream);
/* d w_loc[j][c] += sum_rows de[row][j] * image_loc[row][c] (autograd of :261, K = 5). */
M3P_API int m3p_loc_wgrad(const void* de, const float* image_loc, float* dw_loc, int64_t B, int64_t R, int64_t d,
                          m3p_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step: Trainer.optimize (xtrainer.py:205-243) = clip_grad_norm_ over all
 * parameters + optim.Adam.step (optim.py:45-86), on the flat fp32 parameter / gradient buffers.
 * m3p_sumsq_f32: *out += sum_i x[i]^2 (one device scalar; call once per gradient buffer, zero it first).
 * m3p_adam_step, per element (grad_sumsq read on the device, no host sync):
 *   g' = g * min(1, max_grad_norm / (sqrt(*grad_sumsq) + 1e-6))        (skipped when grad_sumsq == NULL
 *                                                                       or max_grad_norm <= 0)
 *   m = beta1 m + (1-beta1) g' ; v = beta2 v + (1-beta2) g'^2 ; p -= weight_decay*lr*p ;
 *   p -= lr * sqrt(1-beta2^step)/(1-beta1^step) * m / (sqrt(v) + eps)
 *   param_bf16[i] = bf16(p) when given (the tensor-core operand copy); grad[i] = 0 when zero_grad. */
M3P_API int m3p_sumsq_f32(const float* x, int64_t n, float* out, m3p_stream_t stream);
typedef struct m3p_adam_args {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  void* param_bf16; /* optional */
  int64_t n;
  int64_t step; /* 1-based update count (optim.py:68) */
  float lr, beta1, beta2, eps, weight_decay;
  const float* grad_sumsq; /* optional device scalar */
  float max_grad_norm;
  int32_t zero_grad;
} m3p_adam_args;
M3P_API int m3p_adam_step(const m3p_adam_args* args, m3p_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* M3P_B200_H_ */
