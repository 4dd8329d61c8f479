"""ctypes binding of libgillb200.so (the C ABI declared in include/gillb200.h).

The library is the only compute path of this package: if it is missing or a call fails, we raise. There is no
PyTorch / CPU fallback anywhere behind these functions.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgillb200.so")

BF16, F16, F32 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU, ACT_GEGLU, ACT_QUICK_GELU = 0, 1, 2, 3, 4, 5

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_ll = ctypes.c_longlong
_c_float = ctypes.c_float


class GemmArgs(ctypes.Structure):
    """Mirror of `gillb200_gemm_args` (include/gillb200.h)."""

    _fields_ = [
        ("a", _c_void_p), ("lda", _c_ll),
        ("a2", _c_void_p), ("lda2", _c_ll), ("k2", _c_int), ("a2_mode", _c_int),
        ("b", _c_void_p), ("ldb", _c_ll),
        ("M", _c_int), ("N", _c_int), ("K", _c_int),
        ("in_dtype", _c_int),
        ("conv3x3", _c_int),
        ("conv_B", _c_int), ("conv_H", _c_int), ("conv_W", _c_int), ("conv_C", _c_int),
        ("out", _c_void_p), ("ldo", _c_ll), ("out_dtype", _c_int),
        ("out_lo", _c_void_p),
        ("bias", _c_void_p), ("bias_along_m", _c_int),
        ("rowbias", _c_void_p), ("ld_rowbias", _c_ll), ("rows_per_group", _c_int),
        ("residual", _c_void_p), ("ldr", _c_ll), ("res_dtype", _c_int),
        ("act", _c_int), ("alpha", _c_float),
        ("block_n", _c_int),
        ("tile_order", _c_int),
        ("cta_pair", _c_int),
        ("sk_workspace", _c_void_p),
        ("stream_k", _c_int),
        ("stats_out", _c_void_p),
        ("rowstats_out", _c_void_p), ("ln_stats", _c_void_p), ("ln_cs", _c_void_p), ("ln_C", _c_int), ("ln_eps", _c_float),
        ("conv_stride", _c_int),
        ("conv_phase", _c_int),
        ("gn_scale_shift", _c_void_p), ("gn_silu", _c_int), ("a_cat", _c_void_p), ("a_cat_C", _c_int),
    ]


class GillB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the ctypes handle. Raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GillB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). gill_b200 has no fallback compute path."
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.gillb200_last_error.restype = ctypes.c_char_p
        _declare(_lib)
    return _lib


def _declare(L):
    L.gillb200_version.restype = _c_int
    L.gillb200_num_sms.restype = _c_int
    L.gillb200_gemm.argtypes = [ctypes.POINTER(GemmArgs), _c_void_p]
    L.gillb200_gemm.restype = _c_int
    L.gillb200_gemm_streamk_workspace_bytes.restype = _c_ll
    from . import _lib_decl

    _lib_decl.declare(L)


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().gillb200_last_error().decode("utf-8", "replace")
        raise GillB200Error(f"{what} failed (rc={rc}): {msg}")


# every extern "C" symbol include/gillb200.h declares (checked by tests/test_abi.py against the header)
def exported_symbols(header_path=None):
    import re

    header_path = header_path or os.path.join(os.path.dirname(_HERE), "include", "gillb200.h")
    txt = open(header_path).read()
    return sorted(set(re.findall(r"\b(gillb200_[a-z0-9_]+)\s*\(", txt)))


class AttnArgs(ctypes.Structure):
    """Mirror of `gillb200_attn_args` (include/gillb200.h)."""

    _fields_ = [
        ("q", _c_void_p), ("k", _c_void_p), ("v", _c_void_p), ("out", _c_void_p),
        ("ldq", _c_ll), ("ldk", _c_ll), ("ldv", _c_ll), ("ldo", _c_ll),
        ("q_bstride", _c_ll), ("k_bstride", _c_ll), ("v_bstride", _c_ll), ("o_bstride", _c_ll),
        ("kv_lens", _c_void_p),
        ("B", _c_int), ("H", _c_int), ("Lq", _c_int), ("Lk", _c_int), ("hd_pad", _c_int),
        ("causal", _c_int), ("causal_offset", _c_int),
        ("dtype", _c_int),
        ("scale", _c_float),
        ("ones_col", _c_int),
        ("head_stride", _c_int),
    ]
