"""Thin Python wrappers over the libgillb200.so primitives. torch is used only for device buffers and streams."""
import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_QUICK_GELU, ACT_RELU, ACT_SILU, BF16, F16, F32, GemmArgs, check, lib

_DT = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}
_ACT = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "gelu": ACT_GELU, "silu": ACT_SILU, "geglu": ACT_GEGLU,
        "quick_gelu": ACT_QUICK_GELU}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# Optional per-launch timing (bench.py / tools): set ops.PROFILE = [] and every wrapper appends
# (kernel family, algorithmic flops, algorithmic bytes, start event, end event) around its launch.
PROFILE = None


class _P:
    __slots__ = ("name", "flops", "bytes", "e0", "sig")

    def __init__(self, name, flops=0.0, nbytes=0.0, sig=""):
        self.name, self.flops, self.bytes, self.sig = name, flops, nbytes, sig

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.name, self.flops, self.bytes, self.e0, e1, self.sig))
        return False


def profile_summary(records, by_shape=False):
    """Aggregate PROFILE records per kernel family (or per family+shape) -> {name: dict(ms, launches, flops, bytes)}.
    Call after a synchronize."""
    out = {}
    for name, fl, by, e0, e1, sig in records:
        d = out.setdefault(name + (" " + sig if by_shape and sig else ""), dict(ms=0.0, launches=0, flops=0.0, bytes=0.0))
        d["ms"] += e0.elapsed_time(e1)
        d["launches"] += 1
        d["flops"] += fl
        d["bytes"] += by
    return out


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk2d(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor")
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be 2-D with a contiguous last dim, got shape {tuple(t.shape)} stride {t.stride()}")


_sk_ws = {}
_capture_streams = {}


def ensure_streamk_ws(device: torch.device, stream: Optional[torch.cuda.Stream] = None) -> int:
    """Allocate (eagerly, zeroed once: the kernel's arrival flags re-arm themselves) the stream-K scratch of
    (device, stream). One buffer per stream because two GEMMs running concurrently must not share it."""
    device = torch.device(device)
    stream = stream or torch.cuda.current_stream(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream.cuda_stream)
    ws = _sk_ws.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            # a buffer allocated during capture would live in that graph's private pool and dangle once the graph is
            # freed while later captures keep using it
            raise _lib.GillB200Error("stream-K workspace requested for the first time inside a CUDA-graph capture; "
                                     "capture through gill_b200.ops.graph_capture() (it allocates the workspace first)")
        with torch.cuda.stream(stream):
            ws = torch.zeros(lib().gillb200_gemm_streamk_workspace_bytes(), device=device, dtype=torch.uint8)
        stream.synchronize()
        _sk_ws[key] = ws
    return ws.data_ptr()


def _streamk_ws(device: torch.device) -> int:
    return ensure_streamk_ws(device, torch.cuda.current_stream(device))


_zero_bias = {}


def _zero_bias_ptr(device: torch.device, n: int) -> int:
    """A shared all-zero fp32 bias row. The staged epilogue's compile-time variants (bias / bias + residual / ...) all
    take a bias; a linear WITHOUT one fell through to the all-runtime generic variant, which is about twice as slow on
    the short-K shapes (UNet cross-attention to_q, M65536 N384 K320: 55.8 us in the per-shape profile against 28 us
    with a bias). Allocated eagerly (never inside a graph capture: the buffer must outlive every graph)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    z = _zero_bias.get(idx)
    if z is None or z.numel() < n:
        if torch.cuda.is_current_stream_capturing():
            return 0
        z = torch.zeros(max(n, 16384), device=device, dtype=torch.float32)
        torch.cuda.current_stream(device).synchronize()
        _zero_bias[idx] = z
    return z.data_ptr()


def graph_capture(graph: "torch.cuda.CUDAGraph", device=None):
    """`torch.cuda.graph(graph)` on a long-lived per-device capture stream whose stream-K workspace exists BEFORE the
    capture starts (eager allocation, outside every graph's private memory pool)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _capture_streams.get(idx)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _capture_streams[idx] = st
    ensure_streamk_ws(device, st)
    _zero_bias_ptr(device, 16384)
    return torch.cuda.graph(graph, stream=st)


class GraphedCall:
    """A fixed-shape device function replayed from CUDA graphs: one captured graph per input-shape key, static input
    buffers that the caller's tensors are copied into, static outputs (valid until the next call with the same key).
    Used where a stage is a few hundred short launches whose eager submission costs more CPU time than the GPU needs
    (OPT prefill: ~400 launches, GILLMapper: 88) -- the UNet has always run this way. `fn(*tensors)` must be free of host
    synchronisation and data-dependent control flow. At most `max_graphs` shapes are kept (oldest dropped)."""

    def __init__(self, fn, device, max_graphs: int = 4):
        self.fn, self.device, self.max_graphs = fn, torch.device(device), max_graphs
        self._g = {}

    def __call__(self, *tensors, key=()):
        k = (tuple((tuple(t.shape), t.dtype) for t in tensors), key)
        ent = self._g.get(k)
        if ent is None:
            if torch.cuda.is_current_stream_capturing():
                return self.fn(*tensors)
            while len(self._g) >= self.max_graphs:
                self._g.pop(next(iter(self._g)))
            static = [t.detach().clone() for t in tensors]
            self.fn(*static)                                  # eager warm-up: lazy workspaces exist before the capture
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with graph_capture(g, self.device):
                out = self.fn(*static)
            ent = (g, static, out)
            self._g[k] = ent
        g, static, out = ent
        for s_, t in zip(static, tensors):
            s_.copy_(t)
        g.replay()
        return out


def _attach_stats(g: GemmArgs, M: int, n_out: int, device) -> torch.Tensor:
    """GroupNorm statistics of the output, produced by the GEMM epilogue: float [M/32, N, 2] = per 32-row slab and column
    {sum, sum of squares}. Returned tensor is hung on the output as `.gn_stats` (views must carry it over by hand)."""
    st = torch.empty((M // 32, n_out, 2), device=device, dtype=torch.float32)
    g.stats_out = st.data_ptr()
    return st


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
    cta_pair: int = 0,
    tile_order: int = 0,
    stream_k: int = 0,
    stats: bool = False,
    rowstats: bool = False,
    ln=None,
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
    elif out.dtype == torch.float16 and alpha == 1.0 and out_lo is None:
        g.bias = _zero_bias_ptr(a.device, N) or None  # x + 0.0f is exact: same bits as the bias-free epilogue
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.stride(1) == 1
        g.rowbias, g.ld_rowbias, g.rows_per_group = rowbias.data_ptr(), rowbias.stride(0), rows_per_group
    if residual is not None:
        _chk2d(residual, "residual")
        g.residual, g.ldr, g.res_dtype = residual.data_ptr(), residual.stride(0), _DT[residual.dtype]
    g.act, g.alpha, g.block_n, g.cta_pair, g.tile_order = _ACT[act], alpha, block_n, cta_pair, tile_order
    g.stream_k = stream_k
    if stream_k != 1:
        g.sk_workspace = _streamk_ws(a.device)
    if stats:
        out.gn_stats = _attach_stats(g, M, n_out, a.device)
    if rowstats:
        # per 32-column panel and row {sum, sumsq} of the output: the LayerNorm statistics of the NEXT (ln=...) GEMM
        out.ln_stats = torch.empty((n_out // 32, M, 2), device=a.device, dtype=torch.float32)
        g.rowstats_out = out.ln_stats.data_ptr()
    if ln is not None:
        # LayerNorm folded into this GEMM: ln = (row statistics of `a` from its producer, colsum(b), eps); `b` carries the
        # LayerNorm scale and `bias` the shifted bias (see gillb200.h)
        st, cs, eps = ln
        assert st.shape == (K // 32, M, 2) and cs.dtype == torch.float32 and cs.numel() == N
        g.ln_stats, g.ln_cs, g.ln_C, g.ln_eps = st.data_ptr(), cs.data_ptr(), K, eps
    ktot = K + (g.k2 if a2_mode == 1 else 0)
    with _P("gemm" if a2_mode != 2 else "gemm_split", 2.0 * M * N * ktot * (2 if a2_mode == 2 else 1),
            2.0 * (M * ktot + N * ktot) + out.element_size() * M * n_out, f"M{M} N{N} K{ktot} {act or ''}"):
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
    cta_pair: int = 0,
    tile_order: int = 0,
    stream_k: int = 0,
    stride: int = 1,
    stats: bool = False,
    gn=None,
    x2: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """3x3 / pad 1 convolution (stride 1 or 2) as an implicit GEMM.

    gn = (scale_shift [B,2,C] fp32 from groupnorm_scale_shift, silu): GroupNorm(+SiLU) of the INPUT applied inside the conv
    (x is the raw activation; the normalised tensor is never written); x2: second NHWC source holding the last x2.shape[3]
    input channels (cat([x, x2], -1) without the concat; only with gn). Both need conv3x3_gn_supported(x, w).

    x: NHWC [B,H,W,C] (C % 64 == 0); w: [Cout, 9*C (+k2)] with k = (ky*3+kx)*C + c; returns NHWC [B,H/s,W/s,Cout].
    rowbias [B, Cout] is added per sample (time embedding); a2 [B*H*W, k2] K-concatenates a 1x1 shortcut input whose
    weights occupy the trailing k2 columns of w.
    """
    B, H, W, C = x.shape
    assert x.is_contiguous()
    _chk2d(w, "w")
    N = w.shape[0]
    assert stride in (1, 2) and H % stride == 0 and W % stride == 0
    C0 = C
    if x2 is not None:
        assert gn is not None and x2.is_contiguous() and x2.shape[:3] == x.shape[:3] and x2.dtype == x.dtype
        C = C0 + x2.shape[3]
    Hi, Wi = H, W
    H, W = H // stride, W // stride
    if out is None:
        out = torch.empty((B, H, W, N), device=x.device, dtype=out_dtype or x.dtype)
    g = GemmArgs()
    g.a, g.lda = x.data_ptr(), C0
    g.conv3x3, g.conv_B, g.conv_H, g.conv_W, g.conv_C, g.conv_stride = 1, B, Hi, Wi, C, stride
    if gn is not None:
        ss, silu = gn
        assert ss.dtype == torch.float32 and ss.is_contiguous() and ss.shape == (B, 2, C) and stride == 1 and a2 is None
        g.gn_scale_shift, g.gn_silu = ss.data_ptr(), int(silu)
        if x2 is not None:
            g.a_cat, g.a_cat_C = x2.data_ptr(), x2.shape[3]
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
        assert rowbias.dtype == torch.float32 and rowbias.shape[0] in (1, B) and rowbias.stride(1) == 1
        # one row per sample (per-sample timestep) or a single row shared by the whole batch
        g.rowbias, g.ld_rowbias = rowbias.data_ptr(), rowbias.stride(0)
        g.rows_per_group = H * W if rowbias.shape[0] == B and B > 1 else B * H * W
    if residual is not None:
        r2 = residual.view(B * H * W, N)
        g.residual, g.ldr, g.res_dtype = r2.data_ptr(), r2.stride(0), _DT[residual.dtype]
    g.act, g.alpha, g.block_n, g.cta_pair, g.tile_order = _ACT[act], 1.0, block_n, cta_pair, tile_order
    g.stream_k = stream_k
    if stream_k != 1:
        g.sk_workspace = _streamk_ws(x.device)
    if stats:
        out.gn_stats = _attach_stats(g, B * H * W, N, x.device)
    with _P("conv3x3", 2.0 * B * H * W * N * (9 * C + (g.k2 or 0)), 2.0 * (B * H * W * C + N * 9 * C + B * H * W * N),
            f"B{B} {H}x{W} C{C}->{N}"):
        check(lib().gillb200_gemm(ctypes.byref(g), _stream()), "gillb200_gemm(conv3x3)")
    return out


def topk_scores(bank: torch.Tensor, q: torch.Tensor, k: int, *, index_base: int = 0,
                exclude_idx: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None, out=None):
    """Fused (bank @ q.T) + top-k. bank [N,D] bf16, q [Q,D] bf16 (row stride free) -> (values [Q,k] fp32, global
    indices [Q,k] int64). Ties resolve to the lowest row index. exclude_idx: int64 device tensor of global rows to
    down-weight by 1000 -- [n] / [1,n] (shared by all queries) or [Q,n] (one list per query; entries < 0 are unused)."""
    _chk2d(bank, "bank")
    _chk2d(q, "q")
    assert bank.dtype == torch.bfloat16 and q.dtype == torch.bfloat16 and bank.shape[1] == q.shape[1]
    Q, N = q.shape[0], bank.shape[0]
    need = lib().gillb200_topk_workspace_bytes(Q, N)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, device=bank.device, dtype=torch.uint8)
    if out is None:
        vals = torch.empty((Q, k), device=bank.device, dtype=torch.float32)
        idx = torch.empty((Q, k), device=bank.device, dtype=torch.int64)
    else:
        vals, idx = out
        assert vals.dtype == torch.float32 and idx.dtype == torch.int64 and vals.is_contiguous() and idx.is_contiguous()
        assert tuple(vals.shape) == (Q, k) and tuple(idx.shape) == (Q, k)
    n_ex, ex_ld = 0, 0
    if exclude_idx is not None and exclude_idx.numel() > 0:
        ex = exclude_idx if exclude_idx.dim() == 2 else exclude_idx.view(1, -1)
        assert ex.dtype == torch.int64 and ex.is_cuda and ex.stride(1) == 1 and ex.shape[0] in (1, Q)
        n_ex = ex.shape[1]
        ex_ld = ex.stride(0) if ex.shape[0] == Q and Q > 1 else 0
        exclude_idx = ex
    with _P("topk_scores", 2.0 * N * bank.shape[1] * Q, 2.0 * N * bank.shape[1], f"N{N} D{bank.shape[1]} Q{Q} K{k}"):
        check(lib().gillb200_topk_scores(bank.data_ptr(), N, bank.shape[1], bank.stride(0), q.data_ptr(), Q,
                                         q.stride(0), k, index_base, _ptr(exclude_idx) if n_ex else None, n_ex, ex_ld,
                                         workspace.data_ptr(), vals.data_ptr(), idx.data_ptr(), _stream()),
              "gillb200_topk_scores")
    return vals, idx


def topk_merge(cand_val: torch.Tensor, cand_idx: torch.Tensor, k: int):
    """Merge candidate lists [R,Q,Kc] -> top-k per query ordered by (value desc, index asc)."""
    assert cand_val.dtype == torch.float32 and cand_idx.dtype == torch.int64
    assert cand_val.is_contiguous() and cand_idx.is_contiguous() and cand_val.shape == cand_idx.shape
    R, Q, Kc = cand_val.shape
    vals = torch.empty((Q, k), device=cand_val.device, dtype=torch.float32)
    idx = torch.empty((Q, k), device=cand_val.device, dtype=torch.int64)
    with _P("topk_merge"):
        check(lib().gillb200_topk_merge(cand_val.data_ptr(), cand_idx.data_ptr(), R, Q, Kc, k, vals.data_ptr(),
                                        idx.data_ptr(), _stream()), "gillb200_topk_merge")
    return vals, idx


def topk_merge_packed(recv: torch.Tensor, R: int, list_bytes: int, idx_offset_bytes: int, q_first: int, Q: int, Kc: int,
                      k: int, out=None):
    """Merge straight out of a packed exchange buffer (ShardedBank): `recv` holds R lists of `list_bytes` bytes, each
    [fp32 values [Qall,Kc] | padding | int64 indices [Qall,Kc] at idx_offset_bytes]; merges queries q_first..q_first+Q."""
    assert recv.dtype == torch.uint8 and recv.is_contiguous() and list_bytes % 8 == 0 and idx_offset_bytes % 8 == 0
    if out is None:
        vals = torch.empty((Q, k), device=recv.device, dtype=torch.float32)
        idx = torch.empty((Q, k), device=recv.device, dtype=torch.int64)
    else:
        vals, idx = out
    base = recv.data_ptr()
    with _P("topk_merge"):
        check(lib().gillb200_topk_merge_strided(base + q_first * Kc * 4, list_bytes // 4,
                                                base + idx_offset_bytes + q_first * Kc * 8, list_bytes // 8, Kc, R, Q, Kc,
                                                k, vals.data_ptr(), idx.data_ptr(), _stream()),
              "gillb200_topk_merge_strided")
    return vals, idx


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, hd_pad: int, scale: float, *,
              out: Optional[torch.Tensor] = None, causal: bool = False, causal_offset: int = 0,
              kv_lens: Optional[torch.Tensor] = None, ones_col: int = 0, head_dim: int = 0,
              head_stride: int = 0) -> torch.Tensor:
    """q [B,Lq,>=H*stride], k/v [B,Lk,>=H*stride] (views into fused QKV buffers allowed; last dim contiguous), stride =
    head_stride or hd_pad (column pitch of the heads; hd_pad selects the kernel tile width). Returns out [B,Lq,H*stride]. Pad columns of each head must be zero in q, k, v -- except that v may carry 1.0 in
    pad column `ones_col` (> 0) of every head, which moves the softmax row sum onto the tensor core."""
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    for t in (q, k, v):
        assert t.stride(2) == 1 and t.dtype == q.dtype
    if out is None:
        out = torch.empty((B, Lq, heads * (head_stride or hd_pad)), device=q.device, dtype=q.dtype)
    a = _lib.AttnArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.ldq, a.ldk, a.ldv, a.ldo = q.stride(1), k.stride(1), v.stride(1), out.stride(1)
    a.q_bstride, a.k_bstride, a.v_bstride, a.o_bstride = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    a.kv_lens = _ptr(kv_lens)
    a.B, a.H, a.Lq, a.Lk, a.hd_pad = B, heads, Lq, Lk, hd_pad
    a.causal, a.causal_offset = int(causal), causal_offset
    a.dtype, a.scale, a.ones_col, a.head_stride = _DT[q.dtype], scale, ones_col, head_stride
    # profile records count ALGORITHMIC work: the true head dim (head_dim, when the caller states it), not the padding
    hd = head_dim or hd_pad
    with _P("attention", 4.0 * B * heads * Lq * Lk * hd * (0.5 if causal and Lq == Lk else 1.0),
            2.0 * B * heads * hd * (2 * Lq + 2 * Lk), f"B{B} H{heads} Lq{Lq} Lk{Lk} hd{hd} hp{hd_pad}"):
        check(lib().gillb200_attention(ctypes.byref(a), _stream()), "gillb200_attention")
    return out


# ---------------------------------------------------------------------------------------------------------------
# norms / softmax
# ---------------------------------------------------------------------------------------------------------------
def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5, *,
              out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None,
              out_lo: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over the last dim of a [rows, C] tensor (fp32 weights). out_lo receives the bf16 residue."""
    _chk2d(x, "x")
    rows, C = x.shape
    assert w.dtype == torch.float32 and b.dtype == torch.float32
    if out is None:
        out = torch.empty((rows, C), device=x.device, dtype=out_dtype or x.dtype)
    with _P("layernorm", 0.0, (x.element_size() + out.element_size()) * rows * C, f"rows{rows} C{C}"):
        check(lib().gillb200_layernorm(x.data_ptr(), x.stride(0), _DT[x.dtype], w.data_ptr(), b.data_ptr(), eps, rows, C,
                                       out.data_ptr(), out.stride(0), _DT[out.dtype], _ptr(out_lo), _stream()), "gillb200_layernorm")
    return out


_gn_ws = {}
USE_EPILOGUE_GN_STATS = True   # groupnorm() uses `.gn_stats` left by the producing GEMM when present (A/B: set False)


def groupnorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, groups: int, eps: float, *, silu: bool = False,
              x2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """GroupNorm(+SiLU) over NHWC x [B,H,W,C0] (optionally channel-concatenated with x2 [B,H,W,C1])."""
    B, H, W, C0 = x.shape
    C1 = 0 if x2 is None else x2.shape[3]
    assert x.is_contiguous() and (x2 is None or (x2.is_contiguous() and x2.dtype == x.dtype))
    if out is None:
        out = torch.empty((B, H, W, C0 + C1), device=x.device, dtype=out_dtype or x.dtype)
    key = (x.device.index, B, groups)
    ws = _gn_ws.get(key)
    if ws is None:
        # zero-initialised once: the kernel's per-sample arrival counters reset themselves after every launch
        ws = torch.zeros(lib().gillb200_groupnorm_workspace_bytes(B, groups), device=x.device, dtype=torch.uint8)
        _gn_ws[key] = ws
    st0, st1 = getattr(x, "gn_stats", None), (getattr(x2, "gn_stats", None) if x2 is not None else None)
    if USE_EPILOGUE_GN_STATS and st0 is not None and (x2 is None or st1 is not None) and (H * W) % 32 == 0 \
            and x.dtype != torch.float32:
        # the producing GEMMs already left per-slab channel sums: only the (sample, group) reduction + elementwise pass
        with _P("groupnorm", 0.0, (x.element_size() + out.element_size()) * B * H * W * (C0 + C1), f"B{B} {H}x{W} C{C0 + C1} (stats from epilogue)"):
            check(lib().gillb200_groupnorm_from_stats(x.data_ptr(), C0, st0.data_ptr(), _ptr(x2), C1, _ptr(st1), _DT[x.dtype],
                                                      B, H * W, groups, w.data_ptr(), b.data_ptr(), eps, int(silu),
                                                      out.data_ptr(), _DT[out.dtype], ws.data_ptr(), _stream()),
                  "gillb200_groupnorm_from_stats")
        return out
    with _P("groupnorm", 0.0, (2 * x.element_size() + out.element_size()) * B * H * W * (C0 + C1), f"B{B} {H}x{W} C{C0 + C1}"):
        check(lib().gillb200_groupnorm(x.data_ptr(), C0, _ptr(x2), C1, _DT[x.dtype], B, H * W, groups, w.data_ptr(),
                                       b.data_ptr(), eps, int(silu), out.data_ptr(), _DT[out.dtype], ws.data_ptr(),
                                       _stream()), "gillb200_groupnorm")
    return out


def conv3x3_gn_supported(x: torch.Tensor, n_out: int, x2: Optional[torch.Tensor] = None, min_tiles: int = 48) -> bool:
    """Shapes for which conv3x3(..., gn=...) can apply the GroupNorm inside the conv (halo-tile CTA-pair kernel) and whose
    producers left GroupNorm statistics."""
    B, H, W, C0 = x.shape
    C1 = 0 if x2 is None else x2.shape[3]
    tiles = (B * H * W // 256) * ((n_out + 319) // 320 if n_out % 320 == 0 else (n_out + 255) // 256)
    return (USE_EPILOGUE_GN_STATS and x.dtype != torch.float32 and H % 16 == 0 and W % 16 == 0 and (B * H * W) % 256 == 0
            and C0 % 64 == 0 and C1 % 64 == 0 and n_out % 32 == 0 and tiles >= min_tiles
            and getattr(x, "gn_stats", None) is not None and (x2 is None or getattr(x2, "gn_stats", None) is not None))


def groupnorm_scale_shift(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, groups: int, eps: float,
                          x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The (sample, group) reduction of GroupNorm only, from the producers' `.gn_stats`: float [B, 2, C] = per sample and
    channel (scale, shift) with y = x * scale + shift -- for a consumer that applies the normalisation itself
    (conv3x3(..., gn=...))."""
    B, H, W, C0 = x.shape
    C1 = 0 if x2 is None else x2.shape[3]
    st0, st1 = x.gn_stats, (x2.gn_stats if x2 is not None else None)
    ss = torch.empty((B, 2, C0 + C1), device=x.device, dtype=torch.float32)
    with _P("groupnorm", 0.0, 0.0, f"B{B} {H}x{W} C{C0 + C1} (scale/shift only)"):
        check(lib().gillb200_groupnorm_scale_shift(C0, st0.data_ptr(), C1, _ptr(st1), B, H * W, groups, w.data_ptr(), b.data_ptr(),
                                                   eps, ss.data_ptr(), _stream()), "gillb200_groupnorm_scale_shift")
    return ss


def softmax_rows(x: torch.Tensor, scale: float, out_dtype: torch.dtype, out: Optional[torch.Tensor] = None):
    _chk2d(x, "x")
    rows, n = x.shape
    if out is None:
        out = torch.empty((rows, n), device=x.device, dtype=out_dtype)
    with _P("softmax_rows"):
        check(lib().gillb200_softmax_rows(x.data_ptr(), x.stride(0), _DT[x.dtype], scale, rows, n, out.data_ptr(),
                                          out.stride(0), _DT[out.dtype], _stream()), "gillb200_softmax_rows")
    return out


# ---------------------------------------------------------------------------------------------------------------
# small kernels
# ---------------------------------------------------------------------------------------------------------------
def gather_add_rows(table: torch.Tensor, idx: torch.Tensor, x: Optional[torch.Tensor] = None, idx_offset: int = 0,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = (x[i] if x is given else 0) + table[idx[i] + idx_offset]; table [V,D] 16-bit, idx int64 [rows]."""
    assert idx.dtype == torch.int64 and idx.is_contiguous() and table.is_contiguous()
    rows, D = idx.numel(), table.shape[1]
    if out is None:
        out = torch.empty((rows, D), device=table.device, dtype=table.dtype)
    if x is not None:
        assert x.is_contiguous() and x.dtype == table.dtype and x.numel() == rows * D
    with _P("gather_add_rows"):
        check(lib().gillb200_gather_add_rows(_ptr(x), table.data_ptr(), idx.data_ptr(), idx_offset, rows, D,
                                             _DT[table.dtype], out.data_ptr(), _stream()), "gillb200_gather_add_rows")
    return out


def conv3x3_up2_weights(w: torch.Tensor) -> torch.Tensor:
    """[Cout, 9*C] conv weight (k = (ky*3+kx)*C + c) -> [4, Cout, 4*C]: the pre-summed 2 x 2 kernels of the four output
    phases of `conv3x3(upsample2x(x))` (see conv3x3_up2). Phase (a, b), low-res offset (a - 1 + u, b - 1 + v):
    rows ky in {0} / {1, 2} (a = 0; u = 0 / 1) or {0, 1} / {2} (a = 1), columns likewise; summed in fp32, rounded once."""
    cout, C = w.shape[0], w.shape[1] // 9
    acc = torch.float32 if w.dtype in (torch.float16, torch.bfloat16) else w.dtype
    w9 = w.to(acc).view(cout, 3, 3, C)
    sel = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}
    out = torch.zeros((4, cout, 4, C), dtype=acc, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            for u in (0, 1):
                for v in (0, 1):
                    out[a * 2 + b, :, u * 2 + v] = sum(w9[:, ky, kx] for ky in sel[(a, u)] for kx in sel[(b, v)])
    return out.view(4, cout, 4 * C).to(w.dtype).contiguous()


def conv3x3_up2_supported(x: torch.Tensor) -> bool:
    """Shapes the halo-tile phase launches cover (gillb200_gemm_args::conv_phase)."""
    B, H, W, C = x.shape
    return x.is_cuda and x.dtype != torch.float32 and H % 16 == 0 and W % 16 == 0 and C % 64 == 0 and (B * H * W) % 256 == 0 \
        and B * H * W >= 4096


def conv3x3_up2(x: torch.Tensor, w_up2: torch.Tensor, *, bias: Optional[torch.Tensor] = None, stats: bool = False,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`conv3x3(upsample2x(x), w)` (nearest 2x, then 3x3 / pad 1) WITHOUT the upsampled tensor and with 4/9 of the flops:
    output pixel (2i+a, 2j+b) only sees the 2 x 2 low-res pixels (i+a-1 .. i+a, j+b-1 .. j+b), so each of the four output
    phases is a 2 x 2 convolution of the low-res input with pre-summed weights (conv3x3_up2_weights). Four launches of the
    halo-tile implicit-GEMM kernel, each writing its quarter of the NHWC output through a strided tensor map.
    x NHWC [B,H,W,C] -> NHWC [B,2H,2W,Cout]; with stats=True the output carries `.gn_stats` like conv3x3."""
    B, H, W, C = x.shape
    assert x.is_contiguous() and w_up2.dim() == 3 and w_up2.shape[0] == 4 and w_up2.shape[2] == 4 * C and w_up2.is_contiguous()
    N = w_up2.shape[1]
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, N), device=x.device, dtype=x.dtype)
    assert out.is_contiguous() and out.shape == (B, 2 * H, 2 * W, N)
    st = torch.empty((4 * B * H * W // 32, N, 2), device=x.device, dtype=torch.float32) if stats else None
    with _P("conv3x3_up2", 2.0 * 4 * B * H * W * N * 4 * C, 2.0 * (B * H * W * C + 4 * N * 4 * C + 4 * B * H * W * N),
            f"B{B} {H}x{W}->{2 * H}x{2 * W} C{C}->{N}"):
        for ph in range(4):
            g = GemmArgs()
            g.a, g.lda = x.data_ptr(), C
            g.conv3x3, g.conv_B, g.conv_H, g.conv_W, g.conv_C, g.conv_stride, g.conv_phase = 1, B, H, W, C, 1, ph + 1
            g.b, g.ldb = w_up2[ph].data_ptr(), w_up2.stride(1)
            g.M, g.N, g.K = B * H * W, N, 4 * C
            g.in_dtype = _DT[x.dtype]
            g.out, g.ldo, g.out_dtype = out.data_ptr(), N, _DT[out.dtype]
            g.bias = bias.data_ptr() if bias is not None else (_zero_bias_ptr(x.device, N) or None)
            g.act, g.alpha, g.stream_k = ACT_NONE, 1.0, 1
            if st is not None:
                g.stats_out = st.data_ptr()
            check(lib().gillb200_gemm(ctypes.byref(g), _stream()), "gillb200_gemm(conv3x3 up2 phase)")
    if st is not None:
        out.gn_stats = st
    return out


def conv3x3_taps_weight(w: torch.Tensor, cout: int, pad_to: int = 16) -> torch.Tensor:
    """[Cout(+pad), 9*C] conv weight (k = tap*C + c) -> per-tap GEMM weight [9*cout (padded to a multiple of pad_to), C] with
    row tap*cout + o = W[o, tap*C : (tap+1)*C] (see conv3x3_narrow)."""
    C = w.shape[1] // 9
    wt = w[:cout].reshape(cout, 9, C).permute(1, 0, 2).reshape(9 * cout, C)
    rows = (9 * cout + pad_to - 1) // pad_to * pad_to
    out = torch.zeros((rows, C), dtype=w.dtype, device=w.device)
    out[: 9 * cout] = wt
    return out.contiguous()


def conv3x3_narrow(x: torch.Tensor, w_taps: torch.Tensor, cout: int, bias: Optional[torch.Tensor] = None,
                   out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """3x3 / pad 1 convolution with very few output channels (UNet conv_out 320 -> 4, VAE conv_out 128 -> 3) as ONE plain
    GEMM [B*H*W, C] x [C, 9*cout] (fp32 per-tap products; the activation is read once, not nine shifted times through a
    32-column tile that is 7/8 padding) + a nine-tap shifted sum. x NHWC [B,H,W,C] 16-bit -> NHWC [B,H,W,cout]."""
    B, H, W, C = x.shape
    assert x.is_contiguous() and w_taps.shape[1] == C and w_taps.shape[0] >= 9 * cout
    y = gemm(x.view(B * H * W, C), w_taps, out_dtype=torch.float32)
    out = torch.empty((B, H, W, cout), device=x.device, dtype=out_dtype or x.dtype)
    with _P("tap_sum3x3", 0.0, 4.0 * y.numel() + out.element_size() * out.numel(), f"B{B} {H}x{W} Cout{cout}"):
        check(lib().gillb200_tap_sum3x3(y.data_ptr(), y.stride(0), B, H, W, cout, _ptr(bias), out.data_ptr(), _DT[out.dtype],
                                        cout, _stream()), "gillb200_tap_sum3x3")
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    B, H, W, C = x.shape
    assert x.is_contiguous()
    out = torch.empty((B, 2 * H, 2 * W, C), device=x.device, dtype=x.dtype)
    with _P("upsample2x"):
        check(lib().gillb200_upsample2x(x.data_ptr(), B, H, W, C, out.data_ptr(), _stream()), "gillb200_upsample2x")
    return out


def im2col3x3(x: torch.Tensor, stride: int, ld_out: Optional[int] = None) -> torch.Tensor:
    """NHWC x -> [B*Ho*Wo, ld_out] patches (k = (ky*3+kx)*C + c), zero padded to ld_out columns."""
    B, H, W, C = x.shape
    assert x.is_contiguous()
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    ld_out = ld_out or 9 * C
    out = torch.empty((B * Ho * Wo, ld_out), device=x.device, dtype=x.dtype)
    with _P("im2col3x3"):
        check(lib().gillb200_im2col3x3(x.data_ptr(), B, H, W, C, stride, out.data_ptr(), ld_out, _stream()), "gillb200_im2col3x3")
    return out


def plms_step(eps_pair: torch.Tensor, guidance: float, ets: torch.Tensor, head: int, mode: int, c_sample: float,
              c_eps: float, latents: torch.Tensor, cur_sample: torch.Tensor, lat16_pair: Optional[torch.Tensor]):
    """In-place fused CFG + PLMS update (see gillb200_plms_step). eps_pair [2,n...]; ets [4,n] fp32; latents fp32."""
    n = latents.numel()
    assert eps_pair.numel() == 2 * n and ets.numel() == 4 * n and cur_sample.numel() == n
    assert latents.dtype == torch.float32 and ets.dtype == torch.float32 and cur_sample.dtype == torch.float32
    with _P("plms_step"):
        check(lib().gillb200_plms_step(eps_pair.data_ptr(), _DT[eps_pair.dtype], guidance, ets.data_ptr(), head, mode,
                                       c_sample, c_eps, latents.data_ptr(), cur_sample.data_ptr(), _ptr(lat16_pair),
                                       _DT[lat16_pair.dtype] if lat16_pair is not None else 0, n, _stream()), "gillb200_plms_step")


def image_to_u8(x: torch.Tensor, channels: int = 3) -> torch.Tensor:
    """x NHWC [B,H,W,ld>=channels] -> uint8 [B,H,W,channels] = round(clamp(x/2+0.5,0,1)*255)."""
    B, H, W, ld = x.shape
    assert x.is_contiguous()
    out = torch.empty((B, H, W, channels), device=x.device, dtype=torch.uint8)
    with _P("image_to_u8"):
        check(lib().gillb200_image_to_u8(x.data_ptr(), _DT[x.dtype], B * H * W, ld, channels, out.data_ptr(), _stream()), "gillb200_image_to_u8")
    return out


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # openai/clip-vit-large-patch14 preprocessor_config.json
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def hf_clip_geometry(H: int, W: int, size: int = 224):
    """(RH, RW, top, left) of the HF CLIP feature extractor (transformers 4.30.2 CLIPImageProcessor with the
    openai/clip-vit-large-patch14 config): shortest edge -> `size` keeping the aspect (long edge = int(size * long / short)),
    then the centred size x size crop (top = (RH - size) // 2, left = (RW - size) // 2)."""
    if H <= W:
        RH, RW = size, int(size * W / H)
    else:
        RH, RW = int(size * H / W), size
    return RH, RW, (RH - size) // 2, (RW - size) // 2


def clip_preprocess_u8(img_u8: torch.Tensor, size: int = 224, out_dtype: torch.dtype = torch.bfloat16,
                       mean=CLIP_MEAN, std=CLIP_STD, return_resized: bool = False, mode: str = "square"):
    """uint8 NHWC [B,H,W,3] on the device -> CLIP pixel_values NCHW [B,3,size,size]: PIL-exact 8-bit bicubic resample +
    rescale + normalise.
      mode="square"       : `img.resize((size, size))` then the feature extractor on the already square image -- the
                            re-rank of generated images (gill/models.py:733-737);
      mode="feature_extractor": resize shortest edge to `size`, centre crop -- what the HF feature extractor does to image
                            PROMPTS and bank images (gill/utils.py:117-119, models.py:608, scripts/extract_img_embs.py:37)."""
    assert img_u8.dtype == torch.uint8 and img_u8.is_cuda and img_u8.is_contiguous() and img_u8.shape[-1] == 3
    B, H, W, _ = img_u8.shape
    if mode == "square":
        RH, RW, top, left = size, size, 0, 0
    elif mode == "feature_extractor":
        RH, RW, top, left = hf_clip_geometry(H, W, size)
    else:
        raise ValueError(f"mode must be 'square' or 'feature_extractor', got {mode!r}")
    out = torch.empty((B, 3, size, size), device=img_u8.device, dtype=out_dtype)
    rz = torch.empty((B, size, size, 3), device=img_u8.device, dtype=torch.uint8) if return_resized else None
    m3 = (ctypes.c_float * 3)(*mean)
    s3 = (ctypes.c_float * 3)(*std)
    with _P("clip_preprocess_u8"):
        check(lib().gillb200_clip_preprocess_u8_crop(img_u8.data_ptr(), B, H, W, RH, RW, top, left, size, m3, s3,
                                                     out.data_ptr(), _DT[out_dtype], _ptr(rz), _stream()),
              "gillb200_clip_preprocess_u8_crop")
    return (out, rz) if return_resized else out


def l2norm_rows(x: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    _chk2d(x, "x")
    assert x.dtype == torch.float32
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    with _P("l2norm_rows"):
        check(lib().gillb200_l2norm_rows(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(), out.stride(0),
                                         _DT[out_dtype], _stream()), "gillb200_l2norm_rows")
    return out


def cast_add(x: torch.Tensor, y: Optional[torch.Tensor], out_dtype: torch.dtype, *, y_period: int = 0,
             out_lo: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = cast(x + y[i % y_period]) elementwise over contiguous tensors; optional bf16 residue."""
    assert x.is_contiguous() and (y is None or y.is_contiguous())
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    with _P("cast_add"):
        check(lib().gillb200_cast_add(x.data_ptr(), _DT[x.dtype], _ptr(y), _DT[y.dtype] if y is not None else 0, y_period,
                                      out.data_ptr(), _DT[out.dtype], _ptr(out_lo), x.numel(), _stream()), "gillb200_cast_add")
    return out


def attn_small_f32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float, *,
                   out: torch.Tensor, out_lo: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 attention for short sequences: q [B,Lq,*], k/v [B,Lk,*] fp32 views (head h at columns h*128)."""
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    with _P("attn_small_f32"):
        check(lib().gillb200_attn_small_f32(q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
                                            v.data_ptr(), v.stride(1), v.stride(0), B, heads, 128, Lq, Lk, scale,
                                            out.data_ptr(), out.stride(1), out.stride(0), _DT[out.dtype], _ptr(out_lo),
                                            _stream()), "gillb200_attn_small_f32")
    return out


def channel_mix(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """out[..., :] = w @ x[..., :] + b over the last (tiny, <= 8) channel dim; x fp32 contiguous."""
    assert x.dtype == torch.float32 and x.is_contiguous() and w.dtype == torch.float32 and b.dtype == torch.float32
    cin, cout = x.shape[-1], w.shape[0]
    out = torch.empty(x.shape[:-1] + (cout,), device=x.device, dtype=out_dtype)
    with _P("channel_mix"):
        check(lib().gillb200_channel_mix(x.data_ptr(), cin, w.data_ptr(), b.data_ptr(), cout, x.numel() // cin,
                                         out.data_ptr(), _DT[out_dtype], _stream()), "gillb200_channel_mix")
    return out
