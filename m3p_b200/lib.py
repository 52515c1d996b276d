"""ctypes binding of libm3p_sm100.so (the C ABI declared in include/m3p_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os
from ctypes import POINTER, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libm3p_sm100.so")

M3P_EPI_LINEAR, M3P_EPI_GELU, M3P_EPI_DROP_RES, M3P_EPI_DGELU, M3P_EPI_TANH, M3P_EPI_DTANH = range(6)
M3P_EMB_POS, M3P_EMB_LN, M3P_EMB_MASK_PRE, M3P_EMB_MASK_POST, M3P_EMB_DROP2 = 1, 2, 4, 8, 16


class M3PError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", c_void_p), ("b", c_void_p),
        ("m", c_int64), ("n", c_int64), ("k", c_int64),
        ("lda", c_int64), ("ldb", c_int64),
        ("a_mn_major", c_int32), ("b_mn_major", c_int32),
        ("epilogue", c_int32), ("out_f32", c_int32), ("accumulate", c_int32), ("split_k", c_int32),
        ("alpha", c_float),
        ("bias", c_void_p),
        ("out", c_void_p), ("ldo", c_int64),
        ("out2", c_void_p), ("ldo2", c_int64),
        ("aux", c_void_p), ("ldaux", c_int64),
        ("drop_p", c_float), ("seed", c_uint64),
        ("colsum", c_void_p),
        ("split_stride", c_int64),
        ("aux_f32", c_int32),
        ("aux_ln_mean", c_void_p), ("aux_ln_rstd", c_void_p), ("aux_ln_gamma", c_void_p), ("aux_ln_beta", c_void_p),
        ("aux_ln_seqlen", c_void_p), ("aux_ln_S", c_int64),
    ]


class AttnArgs(ctypes.Structure):
    _fields_ = [
        ("qkv", c_void_p), ("seqlen", c_void_p),
        ("B", c_int64), ("S", c_int64), ("H", c_int64),
        ("scale", c_float), ("drop_p", c_float), ("seed", c_uint64),
        ("ctx", c_void_p), ("lse", c_void_p), ("dctx", c_void_p), ("dqkv", c_void_p),
    ]


class LnFwdArgs(ctypes.Structure):
    _fields_ = [
        ("x", c_void_p), ("x_f32", c_int32), ("gamma", c_void_p), ("beta", c_void_p), ("seqlen", c_void_p), ("S", c_int64),
        ("y", c_void_p), ("y_f32", c_void_p), ("mean", c_void_p), ("rstd", c_void_p), ("rows", c_int64), ("d", c_int64),
        ("eps", c_float),
    ]


class LnBwdArgs(ctypes.Structure):
    _fields_ = [
        ("dy", c_void_p), ("x", c_void_p), ("mean", c_void_p), ("rstd", c_void_p), ("gamma", c_void_p),
        ("seqlen", c_void_p), ("S", c_int64),
        ("dx", c_void_p), ("dx_drop", c_void_p),
        ("dx_drop_p", c_float), ("dx_seed", c_uint64),
        ("dy_drop_p", c_float), ("dy_seed", c_uint64),
        ("dgamma", c_void_p), ("dbeta", c_void_p), ("dbias", c_void_p),
        ("rows", c_int64), ("d", c_int64),
        ("x_f32", c_int32), ("dy_f32", c_int32), ("dx_f32", c_int32),
        ("col_scratch", c_void_p),
    ]


class EmbedArgs(ctypes.Structure):
    _fields_ = [
        ("B", c_int64), ("R", c_int64), ("T", c_int64), ("d", c_int64),
        ("flags", c_int32), ("eps", c_float), ("drop_p", c_float),
        ("seed_img", c_uint64), ("seed_emb", c_uint64),
        ("e_img", c_void_p), ("image_loc", c_void_p), ("w_loc", c_void_p), ("b_loc", c_void_p),
        ("ln_img_g", c_void_p), ("ln_img_b", c_void_p), ("img_mean", c_void_p), ("img_rstd", c_void_p),
        ("x", c_void_p), ("tok_emb", c_void_p), ("text_embed", c_void_p), ("positions", c_void_p),
        ("pos_emb", c_void_p), ("langs", c_void_p), ("lang_emb", c_void_p), ("seqlen", c_void_p),
        ("ln_emb_g", c_void_p), ("ln_emb_b", c_void_p),
        ("y_pre", c_void_p), ("emb_mean", c_void_p), ("emb_rstd", c_void_p), ("h0", c_void_p),
        ("h0_f32", c_void_p),
    ]


class EmbedBwdArgs(ctypes.Structure):
    _fields_ = [
        ("B", c_int64), ("R", c_int64), ("T", c_int64), ("d", c_int64),
        ("flags", c_int32),
        ("dy_pre", c_void_p), ("seqlen", c_void_p), ("x", c_void_p), ("positions", c_void_p), ("langs", c_void_p),
        ("pad_index", c_int64),
        ("d_tok_emb", c_void_p), ("d_text_embed", c_void_p), ("d_pos_emb", c_void_p), ("d_lang_emb", c_void_p),
        ("dy_img", c_void_p),
        ("drop_p", c_float), ("seed_emb", c_uint64),
    ]


class AdamArgs(ctypes.Structure):
    _fields_ = [
        ("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
        ("param_bf16", c_void_p),
        ("n", c_int64), ("step", c_int64),
        ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("weight_decay", c_float),
        ("grad_sumsq", c_void_p), ("max_grad_norm", c_float), ("zero_grad", c_int32),
    ]


# name -> argtypes (all return int).  Must list every M3P_API symbol of include/m3p_b200.h
# (tests/test_abi.py checks this table against the header).
PROTOTYPES = {
    "m3p_device_check": [],
    "m3p_set_seed_mix": [c_void_p],
    "m3p_gemm_bf16": [POINTER(GemmArgs), c_void_p],
    "m3p_gemm_bf16_debug": [POINTER(GemmArgs)] + [c_int32] * 6 + [c_void_p],
    "m3p_attention_fwd": [POINTER(AttnArgs), c_void_p],
    "m3p_attention_bwd": [POINTER(AttnArgs), c_void_p],
    "m3p_layernorm_fwd": [POINTER(LnFwdArgs), c_void_p],
    "m3p_layernorm_bwd": [POINTER(LnBwdArgs), c_void_p],
    "m3p_layernorm_bwd_rows": [POINTER(LnBwdArgs), c_void_p],
    "m3p_layernorm_bwd_cols": [POINTER(LnBwdArgs), c_void_p],
    "m3p_colsum_bf16": [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p],
    "m3p_cast_f32_bf16": [c_void_p, c_void_p, c_int64, c_float, c_void_p],
    "m3p_cast_bf16_f32": [c_void_p, c_void_p, c_int64, c_float, c_void_p],
    "m3p_sum_slabs_bf16": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p],
    "m3p_gelu_bwd": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p],
    "m3p_permute_cast_f32_bf16": [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p],
    "m3p_region_prep": [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p],
    "m3p_gather_rows_bf16": [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p],
    "m3p_scatter_rows_bf16": [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p],
    "m3p_cross_entropy_fwd": [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                              c_void_p],
    "m3p_cross_entropy_bwd": [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int64, c_void_p],
    "m3p_masked_mse_fwd": [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p],
    "m3p_masked_mse_bwd": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p],
    "m3p_relation_loss": [c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_void_p, c_void_p, c_void_p],
    "m3p_rowdot_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p],
    "m3p_rowdot_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                       c_void_p],
    "m3p_gather_rows_f32": [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p],
    "m3p_scatter_add_rows_bf16": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p],
    "m3p_scatter_add_rows_f32": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p],
    "m3p_embed_fwd": [POINTER(EmbedArgs), c_void_p],
    "m3p_embed_bwd_route": [POINTER(EmbedBwdArgs), c_void_p],
    "m3p_loc_wgrad": [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p],
    "m3p_sumsq_f32": [c_void_p, c_int64, c_void_p, c_void_p],
    "m3p_adam_step": [POINTER(AdamArgs), c_void_p],
}

_lib = None


def load():
    """Load the shared library (once). Raises M3PError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # a fresh checkout (the .so is git-ignored): build it in-tree if the toolchain is here — still no fallback
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # noqa: BLE001 — reported below with the original cause
            raise M3PError("libm3p_sm100.so is missing and could not be built (%s); build it with "
                           "`python -m m3p_b200.build` (there is no CPU or PyTorch fallback for the M3P hot path)" % e)
    if not os.path.exists(LIB_PATH):
        raise M3PError(
            "libm3p_sm100.so not found at %s — build it with `python -m m3p_b200.build` "
            "(there is no CPU or PyTorch fallback for the M3P hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.m3p_version.restype = c_int
    lib.m3p_version.argtypes = []
    lib.m3p_last_error.restype = ctypes.c_char_p
    lib.m3p_last_error.argtypes = []
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().m3p_last_error()
        raise M3PError("%s failed (code %d): %s" % (what or "m3p call", rc, msg.decode() if msg else "?"))
