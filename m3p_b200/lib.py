"""ctypes binding of libm3p_sm100.so (the C ABI declared in include/m3p_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libm3p_sm100.so")

M3P_EPI_LINEAR, M3P_EPI_GELU, M3P_EPI_DROP_RES, M3P_EPI_DGELU, M3P_EPI_TANH, M3P_EPI_DTANH = range(6)


class M3PError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p),
        ("b", ctypes.c_void_p),
        ("m", ctypes.c_int64),
        ("n", ctypes.c_int64),
        ("k", ctypes.c_int64),
        ("lda", ctypes.c_int64),
        ("ldb", ctypes.c_int64),
        ("a_mn_major", ctypes.c_int32),
        ("b_mn_major", ctypes.c_int32),
        ("epilogue", ctypes.c_int32),
        ("out_f32", ctypes.c_int32),
        ("accumulate", ctypes.c_int32),
        ("split_k", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("bias", ctypes.c_void_p),
        ("out", ctypes.c_void_p),
        ("ldo", ctypes.c_int64),
        ("out2", ctypes.c_void_p),
        ("ldo2", ctypes.c_int64),
        ("aux", ctypes.c_void_p),
        ("ldaux", ctypes.c_int64),
        ("drop_p", ctypes.c_float),
        ("seed", ctypes.c_uint64),
    ]


_lib = None


def load():
    """Load the shared library (once). Raises M3PError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise M3PError(
            "libm3p_sm100.so not found at %s — build it with `python -m m3p_b200.build` "
            "(there is no CPU or PyTorch fallback for the M3P hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.m3p_version.restype = ctypes.c_int
    lib.m3p_last_error.restype = ctypes.c_char_p
    lib.m3p_device_check.restype = ctypes.c_int
    lib.m3p_gemm_bf16.restype = ctypes.c_int
    lib.m3p_gemm_bf16.argtypes = [ctypes.POINTER(GemmArgs), ctypes.c_void_p]
    lib.m3p_gemm_bf16_debug.restype = ctypes.c_int
    lib.m3p_gemm_bf16_debug.argtypes = [ctypes.POINTER(GemmArgs)] + [ctypes.c_int32] * 6 + [ctypes.c_void_p]
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().m3p_last_error()
        raise M3PError("%s failed (code %d): %s" % (what or "m3p call", rc, msg.decode() if msg else "?"))
