"""ctypes binding of the C-ABI in include/vgqa_b200.h.  No CPU fallback: a missing library is an error."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvgqa_b200.so")

_lib = None

c_void_p, c_int, c_float, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_char_p


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m vgqa_b200.build` (nvcc, sm_100a). "
                "vgqa_b200 has no CPU or PyTorch fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.vgqa_last_error.restype = c_char_p
        _declare(_lib)
    return _lib


def _declare(L):
    L.vgqa_gemm_bf16.restype = c_int
    L.vgqa_gemm_bf16.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                 c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_float, c_void_p]
    L.vgqa_ffn_fused.restype = c_int
    L.vgqa_ffn_fused.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
    L.vgqa_mha32.restype = c_int
    L.vgqa_mha32.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p, c_float, c_void_p]
    L.vgqa_enc_attn.restype = c_int
    L.vgqa_enc_attn.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_int, c_void_p]
    L.vgqa_xattn1.restype = c_int
    L.vgqa_xattn1.argtypes = [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, ctypes.c_longlong,
                              c_void_p, c_void_p, c_int, ctypes.c_longlong, c_void_p, c_int, c_float, c_void_p,
                              c_void_p, c_void_p]


    L.vgqa_xattn1_bias.restype = c_int
    L.vgqa_xattn1_bias.argtypes = [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                   c_float, c_void_p, c_void_p, c_void_p]


def check(status: int):
    if status != 0:
        raise RuntimeError("vgqa_b200: " + lib().vgqa_last_error().decode())


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())
