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
    d("gillb200_topk_scores", vp, cll, ci, cll, vp, ci, cll, ci, cll, vp, ci, cll, vp, vp, vp, vp)
    d("gillb200_topk_merge", vp, vp, ci, ci, ci, ci, vp, vp, vp)
    d("gillb200_topk_merge_strided", vp, cll, vp, cll, cll, ci, ci, ci, ci, vp, vp, vp)
    d("gillb200_attention", ctypes.POINTER(AttnArgs), vp)
    d("gillb200_layernorm", vp, cll, ci, vp, vp, cf, ci, ci, vp, cll, ci, vp, vp)
    d("gillb200_groupnorm_workspace_bytes", ci, ci, restype=cll)
    d("gillb200_groupnorm", vp, ci, vp, ci, ci, ci, ci, ci, vp, vp, cf, ci, vp, ci, vp, vp)
    d("gillb200_groupnorm_from_stats", vp, ci, vp, vp, ci, vp, ci, ci, ci, ci, vp, vp, cf, ci, vp, ci, vp, vp)
    d("gillb200_groupnorm_scale_shift", ci, vp, ci, vp, ci, ci, ci, vp, vp, cf, vp, vp)
    d("gillb200_softmax_rows", vp, cll, ci, cf, cll, ci, vp, cll, ci, vp)
    d("gillb200_gather_add_rows", vp, vp, vp, cll, cll, ci, ci, vp, vp)
    d("gillb200_upsample2x", vp, ci, ci, ci, ci, vp, vp)
    d("gillb200_im2col3x3", vp, ci, ci, ci, ci, ci, vp, cll, vp)
    d("gillb200_plms_step", vp, ci, cf, vp, ci, ci, cf, cf, vp, vp, vp, ci, cll, vp)
    d("gillb200_image_to_u8", vp, ci, cll, ci, ci, vp, vp)
    d("gillb200_tap_sum3x3", vp, cll, ci, ci, ci, ci, vp, vp, ci, cll, vp)
    d("gillb200_l2norm_rows", vp, cll, ci, ci, vp, cll, ci, vp)
    d("gillb200_cast_add", vp, ci, vp, ci, cll, vp, ci, vp, cll, vp)
    d("gillb200_attn_small_f32", vp, cll, cll, vp, cll, cll, vp, cll, cll, ci, ci, ci, ci, ci, cf, vp, cll, cll, ci, vp, vp)
    d("gillb200_launch_count", restype=cll)
    d("gillb200_channel_mix", vp, ci, vp, vp, ci, cll, vp, ci, vp)
    d("gillb200_clip_preprocess_u8", vp, ci, ci, ci, ci, ctypes.POINTER(cf), ctypes.POINTER(cf), vp, ci, vp, vp)
    d("gillb200_clip_preprocess_u8_crop", vp, ci, ci, ci, ci, ci, ci, ci, ci, ctypes.POINTER(cf), ctypes.POINTER(cf), vp, ci,
      vp, vp)
