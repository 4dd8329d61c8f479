"""Thin Python wrappers over the libgillb200.so primitives. torch is used only for device buffers and streams."""
import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_RELU, ACT_SILU, BF16, F16, F32, GemmArgs, check, lib

_DT = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}
_ACT = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "gelu": ACT_GELU, "silu": ACT_SILU, "geglu": ACT_GEGLU}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk2d(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor")
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be 2-D with a contiguous last dim, got shape {tuple(t.shape)} stride {t.stride()}")


def gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    out: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    bias: Optional[torch.Tensor] = None,
    bias_along_m: bool = False,
    rowbias: Optional[torch.Tensor] = None,
    rows_per_group: int = 1,
    residual: Optional[torch.Tensor] = None,
    act: Optional[str] = None,
    alpha: float = 1.0,
    a2: Optional[torch.Tensor] = None,
    a2_mode: int = 0,
    out_lo: Optional[torch.Tensor] = None,
    block_n: int = 0,
) -> torch.Tensor:
    """out = act(alpha * a @ b.T + bias + rowbias) + residual     (a: [M,K], b: [N,K], 16-bit; fp32 accumulate).

    a2_mode=1: K-concatenation, out = [a | a2] @ b.T.   a2_mode=2: split precision, out = (a + a2) @ b.T.
    act='geglu': b rows are interleaved (value, gate) pairs and the output has N/2 columns.
    """
    _chk2d(a, "a")
    _chk2d(b, "b")
    M, K = a.shape
    N = b.shape[0]
    n_out = N // 2 if act == "geglu" else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=out_dtype or a.dtype)
    _chk2d(out, "out")
    g = GemmArgs()
    g.a, g.lda = a.data_ptr(), a.stride(0)
    if a2 is not None:
        _chk2d(a2, "a2")
        g.a2, g.lda2, g.k2, g.a2_mode = a2.data_ptr(), a2.stride(0), a2.shape[1], a2_mode
    g.b, g.ldb = b.data_ptr(), b.stride(0)
    g.M, g.N, g.K = M, N, K
    g.in_dtype = _DT[a.dtype]
    g.out, g.ldo, g.out_dtype = out.data_ptr(), out.stride(0), _DT[out.dtype]
    if out_lo is not None:
        g.out_lo = out_lo.data_ptr()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        g.bias, g.bias_along_m = bias.data_ptr(), int(bias_along_m)
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.stride(1) == 1
        g.rowbias, g.ld_rowbias, g.rows_per_group = rowbias.data_ptr(), rowbias.stride(0), rows_per_group
    if residual is not None:
        _chk2d(residual, "residual")
        g.residual, g.ldr, g.res_dtype = residual.data_ptr(), residual.stride(0), _DT[residual.dtype]
    g.act, g.alpha, g.block_n = _ACT[act], alpha, block_n
    check(lib().gillb200_gemm(ctypes.byref(g), _stream()), "gillb200_gemm")
    return out


def conv3x3(
    x: torch.Tensor,
    w: torch.Tensor,
    *,
    out: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    bias: Optional[torch.Tensor] = None,
    rowbias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    act: Optional[str] = None,
    a2: Optional[torch.Tensor] = None,
    block_n: int = 0,
) -> torch.Tensor:
    """3x3 / stride 1 / pad 1 convolution as an implicit GEMM.

    x: NHWC [B,H,W,C] (C % 64 == 0); w: [Cout, 9*C (+k2)] with k = (ky*3+kx)*C + c; returns NHWC [B,H,W,Cout].
    rowbias [B, Cout] is added per sample (time embedding); a2 [B*H*W, k2] K-concatenates a 1x1 shortcut input whose
    weights occupy the trailing k2 columns of w.
    """
    B, H, W, C = x.shape
    assert x.is_contiguous()
    _chk2d(w, "w")
    N = w.shape[0]
    if out is None:
        out = torch.empty((B, H, W, N), device=x.device, dtype=out_dtype or x.dtype)
    g = GemmArgs()
    g.a, g.lda = x.data_ptr(), C
    g.conv3x3, g.conv_B, g.conv_H, g.conv_W, g.conv_C = 1, B, H, W, C
    if a2 is not None:
        _chk2d(a2, "a2")
        g.a2, g.lda2, g.k2, g.a2_mode = a2.data_ptr(), a2.stride(0), a2.shape[1], 1
    g.b, g.ldb = w.data_ptr(), w.stride(0)
    g.M, g.N, g.K = B * H * W, N, 9 * C
    g.in_dtype = _DT[x.dtype]
    o2 = out.view(B * H * W, N)
    g.out, g.ldo, g.out_dtype = o2.data_ptr(), o2.stride(0), _DT[out.dtype]
    if bias is not None:
        assert bias.dtype == torch.float32
        g.bias = bias.data_ptr()
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.shape[0] == B
        g.rowbias, g.ld_rowbias, g.rows_per_group = rowbias.data_ptr(), rowbias.stride(0), H * W
    if residual is not None:
        r2 = residual.view(B * H * W, N)
        g.residual, g.ldr, g.res_dtype = r2.data_ptr(), r2.stride(0), _DT[residual.dtype]
    g.act, g.alpha, g.block_n = _ACT[act], 1.0, block_n
    check(lib().gillb200_gemm(ctypes.byref(g), _stream()), "gillb200_gemm(conv3x3)")
    return out
