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
 *   M3P_EPI_GELU     out = v (pre-activation), out2 = gelu_erf(v)          transformer.py:48-56,224
 *   M3P_EPI_DROP_RES out = aux + dropout(v)        (aux = residual)        transformer.py:951-952,956
 *   M3P_EPI_DGELU    out = v * gelu_erf'(aux)      (aux = pre-activation; backward of :224)
 *   M3P_EPI_TANH     out = tanh(v)                                          transformer.py:556-557
 *   M3P_EPI_DTANH    out = v * (1 - aux^2)         (aux = tanh output; backward of :557)
 * split_k > 1 requires out_f32 = 1 and accumulate = 1 (partial sums are reduced with red.add).
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
} m3p_gemm_args;

M3P_API int m3p_gemm_bf16(const m3p_gemm_args* args, m3p_stream_t stream);

/* Bring-up variant: same as m3p_gemm_bf16 but with the UMMA smem-descriptor constants overridden
 * (tests/tools only; lets one GPU session sweep descriptor hypotheses).  Any value < 0 keeps the
 * built-in constant. */
M3P_API int m3p_gemm_bf16_debug(const m3p_gemm_args* args, int32_t a_lbo, int32_t a_sbo, int32_t a_kstep,
                        int32_t b_lbo, int32_t b_sbo, int32_t b_kstep, m3p_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* M3P_B200_H_ */
