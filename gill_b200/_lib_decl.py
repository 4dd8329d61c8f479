"""argtypes/restype declarations for the non-GEMM entry points of libgillb200.so."""
import ctypes

vp, ci, cll, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def declare(L):
    from ._lib import AttnArgs

    def d(name, *argtypes, restype=ci):
        fn = getattr(L, name)
        fn.argtypes = list(argtypes)
        fn.restype = restype

    d("gillb200_topk_workspace_bytes", ci, cll, restype=cll)
    d("gillb200_topk_scores", vp, cll, ci, cll, vp, ci, cll, ci, cll, vp, ci, vp, vp, vp, vp)
    d("gillb200_topk_merge", vp, vp, ci, ci, ci, ci, vp, vp, vp)
    d("gillb200_attention", ctypes.POINTER(AttnArgs), vp)
