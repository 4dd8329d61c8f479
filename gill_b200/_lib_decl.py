"""argtypes/restype declarations for the non-GEMM entry points of libgillb200.so."""
import ctypes

vp, ci, cll, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def declare(L):
    def d(name, *argtypes):
        fn = getattr(L, name)
        fn.argtypes = list(argtypes)
        fn.restype = ci

    return d
