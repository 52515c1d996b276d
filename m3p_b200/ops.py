"""Thin torch-tensor -> C-ABI call layer (one function per entry point of include/m3p_b200.h).

Every function enqueues on torch's current CUDA stream and returns nothing but its outputs; there is
no CPU path and no PyTorch fallback.  Tensors are only used as typed device pointers.
"""
import ctypes

import torch

from . import lib as L

_byref = ctypes.byref
_vp = ctypes.c_void_p


LAUNCHES = 0  # kernels of libm3p_sm100.so enqueued by this process (every entry point is one kernel launch,
              # m3p_cross_entropy_fwd two); bench.py reports the per-step delta as `gpu_launches`


_STREAM = None  # cached c_void_p of the stream to enqueue on (torch.cuda.current_stream() costs ~15 us per call)


def use_current_stream():
    """Latch torch's current CUDA stream for the calls that follow.  Every entry point of the host layer
    (encoder forward / backward, each head) calls this once, so stream contexts and CUDA-graph capture
    are honoured without paying the lookup on each of the ~300 kernel launches of a step."""
    global _STREAM
    _STREAM = _vp(torch.cuda.current_stream().cuda_stream)
    return _STREAM


class on_stream:
    """`with ops.on_stream(torch_stream):` — enqueue the calls inside on another stream (the side stream
    that carries parameter-gradient kernels underneath the activation-gradient chain)."""

    def __init__(self, stream):
        self.handle = _vp(stream.cuda_stream)

    def __enter__(self):
        global _STREAM
        self.prev, _STREAM = _STREAM, self.handle

    def __exit__(self, *exc):
        global _STREAM
        _STREAM = self.prev


def _stream():
    global LAUNCHES
    LAUNCHES += 1
    if _STREAM is None:
        return use_current_stream()
    return _STREAM


def _p(t):
    return None if t is None else t.data_ptr()


def _lib():
    return L.load()


def device_check():
    L.check(_lib().m3p_device_check(), "m3p_device_check")


def split_k_for(m, n, k):
    """Split-K factor for the weight-gradient GEMMs (small M x N, very long K): aim at ~2 waves of
    128 x 256 tiles over the 148 SMs, keep >= 8 k-blocks of 64 per split."""
    bn = 256 if n > 128 else 128
    tiles = ((m + 127) // 128) * ((n + bn - 1) // bn)
    s = max(1, min(8, round(296 / tiles)))
    kb = (k + 63) // 64
    return max(1, min(s, kb // 8 if kb >= 8 else 1))


def gemm(a, b, m, n, k, out, *, lda=None, ldb=None, ldo=None, a_mn=False, b_mn=False, epi=L.M3P_EPI_LINEAR,
         out_f32=False, accumulate=False, split_k=1, alpha=1.0, bias=None, out2=None, ldo2=None, aux=None,
         ldaux=None, drop_p=0.0, seed=0, colsum=None, split_stride=0, aux_ln=None):
    """C[m][n] = sum_k A(m,k) B(n,k) with a fused epilogue (see m3p_gemm_bf16 in the header)."""
    g = L.GemmArgs()
    g.a, g.b = a.data_ptr(), b.data_ptr()
    g.m, g.n, g.k = m, n, k
    g.lda = a.stride(0) if lda is None else lda
    g.ldb = b.stride(0) if ldb is None else ldb
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.epilogue = epi
    g.out_f32, g.accumulate, g.split_k = int(out_f32), int(accumulate), split_k
    g.alpha = alpha
    g.bias = _p(bias)
    g.out = out.data_ptr()
    g.ldo = out.stride(0) if ldo is None else ldo
    if out2 is not None:
        g.out2 = out2.data_ptr()
        g.ldo2 = out2.stride(0) if ldo2 is None else ldo2
    if aux is not None:
        g.aux = aux.data_ptr()
        g.ldaux = aux.stride(0) if ldaux is None else ldaux
    g.drop_p, g.seed = drop_p, seed
    g.colsum = _p(colsum)
    g.split_stride = split_stride
    g.aux_f32 = int(aux is not None and aux.dtype == torch.float32)
    if aux_ln is not None:  # (mean, rstd, gamma, beta, seqlen or None, S): the residual is LayerNorm(aux), recomputed
        g.aux_ln_mean, g.aux_ln_rstd = aux_ln[0].data_ptr(), aux_ln[1].data_ptr()
        g.aux_ln_gamma, g.aux_ln_beta = aux_ln[2].data_ptr(), aux_ln[3].data_ptr()
        g.aux_ln_seqlen, g.aux_ln_S = _p(aux_ln[4]), aux_ln[5]
    L.check(_lib().m3p_gemm_bf16(_byref(g), _stream()), "m3p_gemm_bf16")
    return out


def linear(x, w, bias, out, **kw):
    """out[rows][n] = x[rows][k] w[n][k]^T (+ bias): nn.Linear forward."""
    return gemm(x, w, x.shape[0], w.shape[0], w.shape[1], out, bias=bias, **kw)


def dgrad(dy, w, out, **kw):
    """out[rows][k] = dy[rows][n] w[n][k]: input gradient of nn.Linear (B operand MN-major, no transpose)."""
    return gemm(dy, w, dy.shape[0], w.shape[1], w.shape[0], out, b_mn=True, **kw)


def wgrad(dy, x, dw, alpha=1.0):
    """dw[n][k] += alpha * dy[rows][n]^T x[rows][k]: weight gradient (both operands MN-major, split-K, fp32 +=)."""
    n, kk, rows = dy.shape[1], x.shape[1], dy.shape[0]
    return gemm(dy, x, n, kk, rows, dw, a_mn=True, b_mn=True, out_f32=True, accumulate=True,
                split_k=split_k_for(n, kk, rows), alpha=alpha, ldo=kk)


def attention_fwd(qkv, seqlen, B, S, H, scale, drop_p, seed, ctx, lse):
    a = L.AttnArgs()
    a.qkv, a.seqlen = qkv.data_ptr(), seqlen.data_ptr()
    a.B, a.S, a.H = B, S, H
    a.scale, a.drop_p, a.seed = scale, drop_p, seed
    a.ctx, a.lse = ctx.data_ptr(), lse.data_ptr()
    L.check(_lib().m3p_attention_fwd(_byref(a), _stream()), "m3p_attention_fwd")


def attention_bwd(qkv, seqlen, B, S, H, scale, drop_p, seed, ctx, lse, dctx, dqkv):
    a = L.AttnArgs()
    a.qkv, a.seqlen = qkv.data_ptr(), seqlen.data_ptr()
    a.B, a.S, a.H = B, S, H
    a.scale, a.drop_p, a.seed = scale, drop_p, seed
    a.ctx, a.lse = ctx.data_ptr(), lse.data_ptr()
    a.dctx, a.dqkv = dctx.data_ptr(), dqkv.data_ptr()
    L.check(_lib().m3p_attention_bwd(_byref(a), _stream()), "m3p_attention_bwd")


def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps, seqlen=None, S=0, y32=None):
    """y (bf16) [, y32 (fp32)] = rowmask * LayerNorm(x); x is bf16 or fp32."""
    a = L.LnFwdArgs()
    a.x, a.x_f32 = x.data_ptr(), int(x.dtype == torch.float32)
    a.gamma, a.beta, a.seqlen, a.S = gamma.data_ptr(), beta.data_ptr(), _p(seqlen), S
    a.y, a.y_f32, a.mean, a.rstd = y.data_ptr(), _p(y32), mean.data_ptr(), rstd.data_ptr()
    a.rows, a.d = x.shape
    a.eps = eps
    L.check(_lib().m3p_layernorm_fwd(_byref(a), _stream()), "m3p_layernorm_fwd")


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, *, seqlen=None, S=0, dx_drop=None, dx_drop_p=0.0, dx_seed=0,
                  dy_drop_p=0.0, dy_seed=0, dgamma=None, dbeta=None, dbias=None, phase="both", col_scratch=None):
    """phase: "both" (row pass then column pass), "rows" (dx / dx_drop only) or "cols" (dgamma / dbeta / dbias
    only, after the row pass of the same arguments has been enqueued)."""
    a = L.LnBwdArgs()
    a.dy, a.x, a.mean, a.rstd, a.gamma = dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr()
    a.seqlen, a.S = _p(seqlen), S
    a.dx, a.dx_drop = dx.data_ptr(), _p(dx_drop)
    a.dx_drop_p, a.dx_seed, a.dy_drop_p, a.dy_seed = dx_drop_p, dx_seed, dy_drop_p, dy_seed
    a.dgamma, a.dbeta, a.dbias = _p(dgamma), _p(dbeta), _p(dbias)
    a.rows, a.d = x.shape
    a.col_scratch = _p(col_scratch)
    if col_scratch is not None:
        assert col_scratch.dtype == torch.float32 and col_scratch.numel() >= 512 * 3 * x.shape[1]
    a.x_f32 = int(x.dtype == torch.float32)
    a.dy_f32 = int(dy.dtype == torch.float32)
    a.dx_f32 = int(dx.dtype == torch.float32)
    fn = {"both": _lib().m3p_layernorm_bwd, "rows": _lib().m3p_layernorm_bwd_rows,
          "cols": _lib().m3p_layernorm_bwd_cols}[phase]
    L.check(fn(_byref(a), _stream()), "m3p_layernorm_bwd")


def colsum(x, out, rows=None, n=None, ld=None):
    rows = x.shape[0] if rows is None else rows
    n = x.shape[1] if n is None else n
    ld = x.stride(0) if ld is None else ld
    L.check(_lib().m3p_colsum_bf16(x.data_ptr(), ld, out.data_ptr(), rows, n, _stream()), "m3p_colsum_bf16")


def cast_f32_bf16(src, dst, n=None, scale=1.0):
    n = src.numel() if n is None else n
    L.check(_lib().m3p_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), n, scale, _stream()), "m3p_cast_f32_bf16")


def cast_bf16_f32(src, dst, n=None, scale=1.0):
    n = src.numel() if n is None else n
    L.check(_lib().m3p_cast_bf16_f32(src.data_ptr(), dst.data_ptr(), n, scale, _stream()), "m3p_cast_bf16_f32")


def sum_slabs_bf16(src, n_slabs, slab_stride, dst, n):
    """dst = bf16(sum of the n_slabs fp32 slabs of src, in index order): deterministic split-K reduction."""
    L.check(_lib().m3p_sum_slabs_bf16(src.data_ptr(), n_slabs, slab_stride, dst.data_ptr(), n, _stream()),
            "m3p_sum_slabs_bf16")


def gelu_bwd(dg, gp, du):
    L.check(_lib().m3p_gelu_bwd(dg.data_ptr(), gp.data_ptr(), du.data_ptr(), gp.numel(), _stream()), "m3p_gelu_bwd")


def permute_cast(src, dst, A, B, F):
    L.check(_lib().m3p_permute_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), A, B, F, _stream()),
            "m3p_permute_cast_f32_bf16")


def region_prep(raw, zero_mask, normalize, out, ori, R, B, F):
    L.check(_lib().m3p_region_prep(raw.data_ptr(), _p(zero_mask), int(normalize), out.data_ptr(), _p(ori), R, B, F,
                                   _stream()), "m3p_region_prep")


def gather_rows(src, flat_idx, n_inner, stride_outer, stride_inner, dst, n, d):
    L.check(_lib().m3p_gather_rows_bf16(src.data_ptr(), flat_idx.data_ptr(), n_inner, stride_outer, stride_inner,
                                        dst.data_ptr(), n, d, _stream()), "m3p_gather_rows_bf16")


def scatter_rows(src, flat_idx, n_inner, stride_outer, stride_inner, dst, n, d):
    L.check(_lib().m3p_scatter_rows_bf16(src.data_ptr(), flat_idx.data_ptr(), n_inner, stride_outer, stride_inner,
                                         dst.data_ptr(), n, d, _stream()), "m3p_scatter_rows_bf16")


def cross_entropy_fwd(logits, y, V, ignore_index, loss, lse, inv_count):
    global LAUNCHES
    LAUNCHES += 1
    L.check(_lib().m3p_cross_entropy_fwd(logits.data_ptr(), logits.stride(0), y.data_ptr(), logits.shape[0], V,
                                         ignore_index, loss.data_ptr(), lse.data_ptr(), inv_count.data_ptr(),
                                         _stream()), "m3p_cross_entropy_fwd")


def cross_entropy_bwd(logits, y, V, ignore_index, lse, inv_count, grad_scale, dlogits):
    L.check(_lib().m3p_cross_entropy_bwd(logits.data_ptr(), logits.stride(0), y.data_ptr(), logits.shape[0], V,
                                         ignore_index, lse.data_ptr(), inv_count.data_ptr(), _p(grad_scale),
                                         dlogits.data_ptr(), dlogits.stride(0), _stream()), "m3p_cross_entropy_bwd")


def masked_mse_fwd(pred, target, weight, loss):
    global LAUNCHES
    LAUNCHES += 1
    n, d = pred.shape
    L.check(_lib().m3p_masked_mse_fwd(pred.data_ptr(), pred.stride(0), target.data_ptr(), weight.data_ptr(), n, d,
                                      loss.data_ptr(), _stream()), "m3p_masked_mse_fwd")


def masked_mse_bwd(pred, target, weight, grad_scale, dpred):
    n, d = pred.shape
    L.check(_lib().m3p_masked_mse_bwd(pred.data_ptr(), pred.stride(0), target.data_ptr(), weight.data_ptr(),
                                      _p(grad_scale), dpred.data_ptr(), dpred.stride(0), n, d, _stream()),
            "m3p_masked_mse_bwd")


def relation_loss(scores, pos_labels, sample_n, w_multi, w_bin, loss, dscores):
    L.check(_lib().m3p_relation_loss(scores.data_ptr(), pos_labels.data_ptr(), pos_labels.numel(), sample_n, w_multi, w_bin,
                                     loss.data_ptr(), dscores.data_ptr(), _stream()), "m3p_relation_loss")


def rowdot_fwd(x, w, bias, out):
    L.check(_lib().m3p_rowdot_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                                  _stream()), "m3p_rowdot_fwd")


def rowdot_bwd(dout, x, w, dx, dw, db, tanh_grad=False):
    L.check(_lib().m3p_rowdot_bwd(dout.data_ptr(), x.data_ptr(), w.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                  db.data_ptr(), x.shape[0], x.shape[1], int(tanh_grad), _stream()), "m3p_rowdot_bwd")


def gather_rows_f32(table, idx, dst, n, d):
    L.check(_lib().m3p_gather_rows_f32(table.data_ptr(), idx.data_ptr(), dst.data_ptr(), n, d, _stream()),
            "m3p_gather_rows_f32")


def scatter_add_rows_f32(src, idx, skip_index, dst, n, d):
    if src.dtype == torch.bfloat16:
        L.check(_lib().m3p_scatter_add_rows_bf16(src.data_ptr(), idx.data_ptr(), skip_index, dst.data_ptr(), n, d,
                                                 _stream()), "m3p_scatter_add_rows_bf16")
        return
    L.check(_lib().m3p_scatter_add_rows_f32(src.data_ptr(), idx.data_ptr(), skip_index, dst.data_ptr(), n, d,
                                            _stream()), "m3p_scatter_add_rows_f32")


def embed_fwd(args):
    L.check(_lib().m3p_embed_fwd(_byref(args), _stream()), "m3p_embed_fwd")


def embed_bwd_route(args):
    L.check(_lib().m3p_embed_bwd_route(_byref(args), _stream()), "m3p_embed_bwd_route")


def loc_wgrad(de, image_loc, dw_loc, B, R, d):
    L.check(_lib().m3p_loc_wgrad(de.data_ptr(), image_loc.data_ptr(), dw_loc.data_ptr(), B, R, d, _stream()),
            "m3p_loc_wgrad")


def sumsq(x, out):
    """out[0] += sum(x^2) (fp32, 16-byte aligned buffer)."""
    L.check(_lib().m3p_sumsq_f32(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "m3p_sumsq_f32")


def adam_step(p, g, m, v, step, lr, beta1, beta2, eps, weight_decay=0.0, p16=None, grad_sumsq=None, max_grad_norm=0.0,
              zero_grad=False):
    a = L.AdamArgs()
    a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
    a.param_bf16 = _p(p16)
    a.n, a.step = p.numel(), step
    a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = lr, beta1, beta2, eps, weight_decay
    a.grad_sumsq, a.max_grad_norm, a.zero_grad = _p(grad_sumsq), max_grad_norm, int(zero_grad)
    L.check(_lib().m3p_adam_step(_byref(a), _stream()), "m3p_adam_step")
