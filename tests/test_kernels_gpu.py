"""GPU parity tests of the primitive kernels, called through the C ABI (gill_b200.ops -> libgillb200.so)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


def rel(a, b):
    a, b = a.float().cpu().double(), b.float().cpu().double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from gill_b200 import ops as o

    return o


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bn", [0, 32, 64, 128, 160, 256])
def test_gemm_block_n_variants(ops, dt, bn):
    torch.manual_seed(0)
    a = torch.randn(300, 328, device=dev).to(dt)
    b = torch.randn(520, 328, device=dev).to(dt)
    got = ops.gemm(a, b, block_n=bn, out_dtype=torch.float32)
    assert rel(got, a.float() @ b.float().T) < 1e-5


@pytest.mark.parametrize("shape", [(1, 8, 64), (77, 768, 512), (8, 512, 4096), (1000, 328, 72), (4096, 4096, 1024)])
def test_gemm_ragged_shapes(ops, shape):
    M, N, K = shape
    torch.manual_seed(1)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    assert rel(ops.gemm(a, b, out_dtype=torch.float32), a.float() @ b.float().T) < 1e-5


def test_gemm_epilogues(ops):
    torch.manual_seed(2)
    M, N, K = 384, 640, 320
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    base = a.float() @ b.float().T
    got = ops.gemm(a, b, bias=bias, residual=res, act="relu", alpha=0.5, out_dtype=torch.float32)
    assert rel(got, torch.relu(0.5 * base + bias) + res) < 1e-5
    assert rel(ops.gemm(a, b, bias=bias, act="gelu", out_dtype=torch.float32), F.gelu(base + bias)) < 1e-5
    assert rel(ops.gemm(a, b, bias=bias, act="silu", out_dtype=torch.float32), F.silu(base + bias)) < 1e-5
    # GEGLU with interleaved (value, gate) rows
    val, gate = b[: N // 2], b[N // 2:]
    bi = torch.stack([val, gate], 1).reshape(N, K).contiguous()
    bias_i = torch.stack([bias[: N // 2], bias[N // 2:]], 1).reshape(N).contiguous()
    ref = (a.float() @ val.float().T + bias[: N // 2]) * F.gelu(a.float() @ gate.float().T + bias[N // 2:])
    assert rel(ops.gemm(a, bi, bias=bias_i, act="geglu", out_dtype=torch.float32), ref) < 1e-5
    # bias along M, per-row-group bias
    bm = torch.randn(M, device=dev)
    rb = torch.randn(M // 128, N, device=dev)
    got = ops.gemm(a, b, bias=bm, bias_along_m=True, rowbias=rb, rows_per_group=128, out_dtype=torch.float32)
    assert rel(got, base + bm[:, None] + rb.repeat_interleave(128, 0)) < 1e-5
    # 16-bit output with residual
    got = ops.gemm(a, b, bias=bias, residual=res.bfloat16())
    assert got.dtype == torch.bfloat16 and rel(got, base + bias + res.bfloat16().float()) < 5e-3


def test_gemm_second_a_source(ops):
    torch.manual_seed(3)
    M, N, K = 256, 192, 320
    a = torch.randn(M, K, device=dev).bfloat16()
    a2 = torch.randn(M, 128, device=dev).bfloat16()
    b = torch.randn(N, K + 128, device=dev).bfloat16()
    assert rel(ops.gemm(a, b, a2=a2, a2_mode=1, out_dtype=torch.float32), torch.cat([a, a2], 1).float() @ b.float().T) < 1e-5
    x = torch.randn(M, K, device=dev)
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    got = ops.gemm(hi, w, a2=lo, a2_mode=2, out_dtype=torch.float32)
    assert rel(got, x.double() @ w.double().T) < 2e-5          # split precision: ~16 mantissa bits of the activations
    assert rel(ops.gemm(hi, w, out_dtype=torch.float32), x.double() @ w.double().T) > 5e-4   # plain bf16 is much worse
    o_hi = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    o_lo = torch.empty_like(o_hi)
    ops.gemm(hi, w, out=o_hi, out_lo=o_lo)
    assert rel(o_hi.float() + o_lo.float(), hi.float() @ w.float().T) < 2e-5


@pytest.mark.parametrize("case", [(2, 64, 64, 64, 64), (2, 32, 32, 128, 320), (2, 16, 16, 64, 96), (4, 8, 8, 128, 64),
                                  (3, 8, 8, 64, 64), (1, 128, 128, 64, 32), (1, 256, 256, 64, 16)])
def test_conv3x3_implicit_gemm(ops, case):
    B, H, W, C, Co = case
    torch.manual_seed(4)
    x = torch.randn(B, H, W, C, device=dev).half()
    w = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half()
    bias = torch.randn(Co, device=dev)
    rb = torch.randn(B, Co, device=dev)
    res = torch.randn(B, H, W, Co, device=dev).half()
    wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    got = ops.conv3x3(x, wk, bias=bias, rowbias=rb, residual=res, out_dtype=torch.float32)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1) + rb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1) + res.float()
    assert rel(got, ref) < 1e-4


@pytest.mark.parametrize("case", [(2, 64, 64, 64, 320, "res"), (3, 32, 32, 128, 640, "rowbias"), (16, 16, 16, 64, 320, "res"),
                                  (75, 16, 16, 64, 320, "stats"), (38, 32, 32, 64, 640, "res"), (5, 8, 8, 64, 320, "res")])
def test_conv3x3_wide_pair_tile(ops, case):
    """block_n = 320: the CTA-pair wide tile (two N = 160 MMAs per K-step into one 320-column accumulator, no TMEM double
    buffering), including tile counts that exercise whole rounds, a split tail (both N halves on different pairs) and an
    unsplit tail, the residual / per-sample row bias / GroupNorm-statistics epilogues and an M tail (5*64 = 320 rows)."""
    B, H, W, C, Co, mode = case
    torch.manual_seed(5)
    x = torch.randn(B, H, W, C, device=dev).half()
    w = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half()
    bias = torch.randn(Co, device=dev)
    wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    kw = {}
    if mode == "res":
        res = torch.randn(B, H, W, Co, device=dev).half()
        kw["residual"] = res
        ref = ref + res.float()
    elif mode == "rowbias":
        rb = torch.randn(B, Co, device=dev)
        kw["rowbias"] = rb
        ref = ref + rb[:, None, None, :]
    got = ops.conv3x3(x, wk, bias=bias, block_n=320, stats=(mode == "stats"), **kw)
    assert got.dtype == torch.float16 and rel(got, ref) < 2e-3
    auto = ops.conv3x3(x, wk, bias=bias, block_n=160, **kw)
    assert rel(got, auto) < 1e-3
    if mode == "stats":
        # slabs are 32 rows of ONE sample (consecutive rows, or a 4 x 8 pixel patch under the halo-tile conv): the consumer
        # only ever adds all slabs of a sample, so that is what is pinned
        st = got.gn_stats.view(B, H * W // 32, Co, 2).sum(1)
        g = got.float().view(B, H * W, Co)
        assert rel(st[..., 0], g.sum(1)) < 1e-3 and rel(st[..., 1], (g * g).sum(1)) < 1e-3


@pytest.mark.parametrize("case", [(16, 8, 8, 320, 640, "res"), (16, 8, 8, 320, 1280, "rowbias"), (8, 16, 16, 320, 640, "stats"),
                                  (16, 8, 8, 640, 1280, "stats_res")])
def test_conv3x3_split_k_pair(ops, case):
    """Small-M convs (UNet 8x8 / 16x16 levels), stream_k=3: K cut into slices over the wide CTA-pair tile, fp32 partials
    through the scratch buffer, second pass adds bias / row bias / residual, rounds and leaves the GroupNorm statistics.
    (Opt-in: measured slower than the default 1-CTA stream-K on B200, see gemm.cu.)"""
    B, H, W, C, Co, mode = case
    torch.manual_seed(6)
    x = torch.randn(B, H, W, C, device=dev).half()
    w = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half()
    bias = torch.randn(Co, device=dev)
    wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    kw = {}
    if "res" in mode:
        res = torch.randn(B, H, W, Co, device=dev).half()
        kw["residual"] = res
        ref = ref + res.float()
    if mode == "rowbias":
        rb = torch.randn(B, Co, device=dev)
        kw["rowbias"] = rb
        ref = ref + rb[:, None, None, :]
    n0 = ops.lib().gillb200_launch_count()
    got = ops.conv3x3(x, wk, bias=bias, stats="stats" in mode, stream_k=3, **kw)
    assert ops.lib().gillb200_launch_count() - n0 == 2                      # main launch + split-K reduce
    assert got.dtype == torch.float16 and rel(got, ref) < 2e-3
    plain = ops.conv3x3(x, wk, bias=bias, block_n=160, stream_k=1, **kw)
    assert rel(got, plain) < 1e-3
    if "stats" in mode:
        st = got.gn_stats.view(B, H * W // 32, Co, 2)
        g32 = got.float().view(B, H * W // 32, 32, Co)
        assert rel(st[..., 0], g32.sum(2)) < 1e-3 and rel(st[..., 1], (g32 * g32).sum(2)) < 1e-3
    assert torch.equal(ops.conv3x3(x, wk, bias=bias, stream_k=3, **kw), got)   # deterministic (fixed slice order)


@pytest.mark.parametrize("case", [(1024, 1280, 5120, 0), (4096, 1280, 1536, 0), (520, 384, 2048, 128), (1024, 10240, 1280, 256)])
def test_gemm_stream_k(ops, case):
    """stream_k=2 forces the K range of every tile to be shared between CTAs (fp32 partials through the workspace,
    finished in a fixed CTA order): same result as whole-tile scheduling up to fp32 summation order, bit-stable."""
    M, N, K, bn = case
    torch.manual_seed(20)
    a = torch.randn(M, K, device=dev).half()
    b = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).half()
    ref = a.float() @ b.float().T + bias + res.float()
    plain = ops.gemm(a, b, bias=bias, residual=res, stream_k=1, block_n=bn)
    sk1 = ops.gemm(a, b, bias=bias, residual=res, stream_k=2, block_n=bn)
    sk2 = ops.gemm(a, b, bias=bias, residual=res, stream_k=2, block_n=bn)
    assert rel(sk1, ref) < 1e-3 and rel(plain, ref) < 1e-3 and rel(sk1, plain) < 5e-4
    assert torch.equal(sk1, sk2)
    # fp32 output goes through the generic staged epilogue
    o32 = ops.gemm(a, b, bias=bias, stream_k=2, block_n=bn, out_dtype=torch.float32)
    assert rel(o32, a.float() @ b.float().T + bias) < 1e-5
    # ragged M (rows beyond M in the last tile) must not leave arrival flags behind for the next launch
    from gill_b200 import ops as _o
    for ws in _o._sk_ws.values():
        torch.cuda.synchronize()
        assert int(ws[:16384].view(torch.int32).abs().sum().item()) == 0


@pytest.mark.parametrize("case", [(16, 8, 1280, 1280), (4, 16, 640, 1280), (2, 8, 2560, 320)])
def test_conv3x3_stream_k(ops, case):
    B, HW, C, Co = case
    torch.manual_seed(21)
    x = torch.randn(B, HW, HW, C, device=dev).half()
    w = (torch.randn(Co, C, 3, 3, device=dev) * 0.02).half()
    wk = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    bias = torch.randn(Co, device=dev)
    rb = torch.randn(B, Co, device=dev)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1) + rb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1)
    for sk in (1, 2, 0):
        got = ops.conv3x3(x, wk, bias=bias, rowbias=rb, stream_k=sk)
        assert rel(got, ref) < 1e-3, sk
    res = torch.randn(B, HW, HW, Co, device=dev).half()
    got = ops.conv3x3(x, wk, bias=bias, residual=res, stream_k=2)
    ref2 = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1).permute(0, 2, 3, 1) + res.float()
    assert rel(got, ref2) < 1e-3
    assert torch.equal(got, ops.conv3x3(x, wk, bias=bias, residual=res, stream_k=2))


@pytest.mark.parametrize("case", [(2, 64, 320, 320), (2, 32, 640, 640), (4, 16, 1280, 1280), (1, 16, 64, 96)])
def test_conv3x3_stride2_implicit(ops, case):
    """Stride-2 downsampler through TMA element strides == F.conv2d(stride=2, padding=1)."""
    B, HW, C, Co = case
    torch.manual_seed(22)
    x = torch.randn(B, HW, HW, C, device=dev).half()
    w = (torch.randn(Co, C, 3, 3, device=dev) * 0.03).half()
    wk = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    bias = torch.randn(Co, device=dev)
    got = ops.conv3x3(x, wk, bias=bias, stride=2, out_dtype=torch.float32)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, stride=2, padding=1).permute(0, 2, 3, 1)
    assert got.shape == ref.shape and rel(got, ref) < 1e-5


def test_geglu_fast_epilogue_fp16(ops):
    """fp16-output GEGLU takes the specialised staged epilogue with the polynomial GELU (|err| <= 7e-5 abs)."""
    torch.manual_seed(23)
    M, N, K = 512, 1280, 320
    a = torch.randn(M, K, device=dev).half()
    b = (torch.randn(N, K, device=dev) * 0.08).half()
    bias = torch.randn(N, device=dev)
    val, gate = b[0::2], b[1::2]
    ref = (a.float() @ val.float().T + bias[0::2]) * F.gelu(a.float() @ gate.float().T + bias[1::2])
    got = ops.gemm(a, b, bias=bias, act="geglu")
    assert got.dtype == torch.float16 and rel(got, ref) < 1e-3
    assert (got.float() - ref).abs().max() < 2e-3 * ref.abs().max()


def test_conv_helpers(ops):
    torch.manual_seed(5)
    x = torch.randn(2, 16, 16, 64, device=dev).half()
    w = (torch.randn(96, 64, 3, 3, device=dev) * 0.05).half()
    cols = ops.im2col3x3(x, 2)
    got = ops.gemm(cols, w.permute(0, 2, 3, 1).reshape(96, -1).contiguous(), out_dtype=torch.float32).view(2, 8, 8, 96)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    assert rel(got, ref) < 1e-5
    x4 = torch.randn(2, 8, 8, 4, device=dev).half()
    cols = ops.im2col3x3(x4, 1, ld_out=64)
    assert cols.shape == (128, 64) and (cols[:, 36:] == 0).all()
    up = ops.upsample2x(x)
    assert torch.equal(up, F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1))


@pytest.mark.parametrize("C,dt", [(320, torch.float16), (200, torch.float16), (512, torch.float32), (640, torch.float16),
                                  (1280, torch.float16), (4096, torch.float32)])
def test_layernorm(ops, C, dt):
    torch.manual_seed(6)
    x = torch.randn(301, C, device=dev).to(dt)
    w, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    assert rel(ops.layernorm(x, w, b, 1e-5, out_dtype=torch.float32), F.layer_norm(x.float(), (C,), w, b, 1e-5)) < 1e-5
    if dt == torch.float32:
        hi = torch.empty(301, C, device=dev, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        ops.layernorm(x, w, b, 1e-5, out=hi, out_lo=lo)
        assert rel(hi.float() + lo.float(), F.layer_norm(x, (C,), w, b, 1e-5)) < 2e-5


@pytest.mark.parametrize("rows,C,dt", [(65536, 320, torch.float16), (16387, 640, torch.float16), (4099, 1280, torch.float16),
                                       (2051, 256, torch.bfloat16), (40000, 1000 - 8, torch.float16)])
def test_layernorm_persistent_16bit(ops, rows, C, dt):
    """The persistent 16-bit LayerNorm (grid-stride row blocks, next block prefetched): the UNet's shapes, ragged row counts,
    a channel count that leaves lanes without a vector, and a strided input view."""
    torch.manual_seed(61)
    big = (torch.randn(rows, C + 64, device=dev) * 1.5 + 0.3).to(dt)
    x = big[:, :C]                                                    # row stride C + 64
    w, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    got = ops.layernorm(x, w, b, 1e-5)
    assert got.dtype == dt and got.shape == (rows, C)
    ref = F.layer_norm(x.float(), (C,), w, b, 1e-5)
    assert rel(got, ref) < (6e-3 if dt == torch.bfloat16 else 8e-4)


@pytest.mark.parametrize("case", [(2, 64, 64, 320, 0, True), (2, 32, 32, 640, 320, True), (2, 8, 8, 1280, 1280, False),
                                  (1, 128, 128, 256, 0, True), (2, 16, 16, 1280, 640, True)])
def test_groupnorm_nhwc_with_concat(ops, case):
    B, H, W, C0, C1, silu = case
    torch.manual_seed(7)
    x0 = torch.randn(B, H, W, C0, device=dev).half() * 2 + 0.5
    x1 = torch.randn(B, H, W, C1, device=dev).half() if C1 else None
    C = C0 + C1
    w, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    got = ops.groupnorm(x0, w, b, 32, 1e-5, silu=silu, x2=x1, out_dtype=torch.float32)
    xx = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.group_norm(xx.float().permute(0, 3, 1, 2), 32, w, b, 1e-5)
    ref = F.silu(ref) if silu else ref
    assert rel(got, ref.permute(0, 2, 3, 1)) < 2e-5


def test_small_kernels(ops):
    torch.manual_seed(8)
    s = torch.randn(200, 4096, device=dev)
    assert rel(ops.softmax_rows(s, 0.3, torch.float32), torch.softmax(s * 0.3, -1)) < 1e-5
    tab = torch.randn(100, 256, device=dev).bfloat16()
    idx = torch.randint(0, 90, (37,), device=dev)
    xx = torch.randn(37, 256, device=dev).bfloat16()
    assert torch.equal(ops.gather_add_rows(tab, idx, x=xx, idx_offset=2), (xx.float() + tab[idx + 2].float()).bfloat16())
    assert torch.equal(ops.gather_add_rows(tab, idx), tab[idx])
    x = torch.randn(9, 256, device=dev)
    assert rel(ops.l2norm_rows(x, torch.float32), x / x.norm(dim=-1, keepdim=True)) < 1e-6
    img = torch.randn(2, 16, 16, 8, device=dev).half()
    ref = ((img[..., :3].float() / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
    assert torch.equal(ops.image_to_u8(img, 3), ref)
    y = torch.randn(8, 64, device=dev)
    xb = torch.randn(5, 8, 64, device=dev)
    lo = torch.empty(5, 8, 64, device=dev, dtype=torch.bfloat16)
    hi = ops.cast_add(xb, y, torch.bfloat16, y_period=y.numel(), out_lo=lo)
    assert rel(hi.float() + lo.float(), xb + y) < 2e-5


def test_plms_step_matches_oracle_scheduler(ops):
    from gill_b200 import sd as psd
    from oracle import sd15 as osd

    torch.manual_seed(9)
    table = psd.plms_table(50)
    sched = osd.PNDM()
    sched.set_timesteps(50)
    n = 4 * 8 * 8 * 4
    lat = torch.randn(n, device=dev)
    lat_ref = lat.clone().cpu()
    ets, cur = torch.zeros(4, n, device=dev), torch.zeros(n, device=dev)
    pair = torch.zeros(2, n, device=dev, dtype=torch.float16)
    head = 0
    for i, (t, cs, ce, mode) in enumerate(table):
        eps = torch.randn(2, n, device=dev)
        ops.plms_step(eps, 7.5, ets, head, mode, cs, ce, lat, cur, pair)
        if mode != 1:
            head = (head + 1) & 3
        e = eps.cpu()
        lat_ref = sched.step(e[0] + 7.5 * (e[1] - e[0]), t, lat_ref)
        assert rel(lat, lat_ref) < 1e-5, i
    assert rel(pair[0], lat) < 1e-3 and torch.equal(pair[0], pair[1])


CASES = [  # B, H, Lq, Lk, hd, hd_pad, dtype, causal
    (2, 2, 128, 128, 40, 64, torch.float16, False), (2, 8, 1024, 1024, 40, 64, torch.float16, False),
    (2, 8, 1024, 77, 40, 64, torch.float16, False), (2, 8, 256, 256, 80, 128, torch.float16, False),
    (2, 8, 64, 64, 160, 192, torch.float16, False), (2, 8, 256, 77, 160, 192, torch.float16, False),
    (3, 32, 81, 81, 128, 128, torch.bfloat16, True), (2, 4, 300, 300, 128, 128, torch.bfloat16, True),
    # cross-attention shapes (Lk <= 128), ragged Lq
    (2, 4, 1000, 128, 40, 64, torch.float16, False), (1, 4, 512, 50, 80, 128, torch.float16, False),
    (2, 8, 4096, 77, 40, 64, torch.float16, False), (2, 2, 640, 77, 64, 64, torch.bfloat16, False),
]


def attn_ref(q, k, v, H, hp, scale, causal=False, kv_lens=None):
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    qh, kh, vh = (t.float().view(B, -1, H, hp).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) * scale
    if causal:
        i, j = torch.arange(Lq, device=dev)[:, None], torch.arange(Lk, device=dev)[None]
        s = s.masked_fill(j > i, float("-inf"))
    if kv_lens is not None:
        j = torch.arange(Lk, device=dev)[None, None, None]
        s = s.masked_fill(j >= kv_lens.view(B, 1, 1, 1), float("-inf"))
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Lq, H * hp)


@pytest.mark.parametrize("case", CASES)
def test_fused_attention(ops, case):
    B, H, Lq, Lk, hd, hp, dt, causal = case
    torch.manual_seed(10)

    def mk(L):
        t = torch.zeros(B, L, H, hp, device=dev)
        t[..., :hd] = torch.randn(B, L, H, hd, device=dev)
        return t.view(B, L, H * hp).to(dt)

    q, k, v = mk(Lq), mk(Lk), mk(Lk)
    got = ops.attention(q, k, v, H, hp, hd ** -0.5, causal=causal)
    tol = 1e-2 if dt == torch.bfloat16 else 3e-3   # P is rounded to the operand dtype before P.V
    assert torch.isfinite(got.float()).all() and rel(got, attn_ref(q, k, v, H, hp, hd ** -0.5, causal)) < tol
    if dt == torch.float16 and hp > hd:
        # ones-column mode: V[..., hd] = 1 moves the softmax denominator onto the tensor core (packed f16x2 exp2)
        v1 = v.view(B, Lk, H, hp).clone()
        v1[..., hd] = 1.0
        v1 = v1.view(B, Lk, H * hp)
        got1 = ops.attention(q, k, v1, H, hp, hd ** -0.5, causal=causal, ones_col=hd)
        ref1 = attn_ref(q, k, v1, H, hp, hd ** -0.5, causal)
        assert torch.isfinite(got1.float()).all() and rel(got1, ref1) < tol
        assert (got1.view(B, Lq, H, hp)[..., hd].float() - 1).abs().max() < 2e-3


PITCH_CASES = [  # B, H, Lq, Lk, hd, pitch, tile, dtype, causal     (SD-1.5 head dims 40 / 80 / 160 at pitch 48 / 96 / 176)
    (2, 8, 1024, 1024, 40, 48, 64, torch.float16, False), (2, 8, 384, 77, 40, 48, 64, torch.float16, False),
    (2, 8, 512, 512, 80, 96, 128, torch.float16, False), (2, 8, 300, 77, 80, 96, 128, torch.float16, False),
    (2, 8, 256, 256, 160, 176, 192, torch.float16, False), (1, 8, 64, 77, 160, 176, 192, torch.float16, False),
    (2, 3, 200, 200, 32, 48, 64, torch.bfloat16, True), (2, 8, 2048, 2048, 40, 48, 64, torch.float16, False),
    # short key sequences at the UNet's real batch: attn4q walks several query tiles per CTA with K / V resident
    (16, 8, 4096, 77, 40, 48, 64, torch.float16, False), (16, 8, 1024, 77, 80, 96, 128, torch.float16, False),
    (3, 4, 700, 96, 32, 48, 64, torch.bfloat16, False), (2, 2, 1000, 128, 80, 96, 128, torch.float16, False),
    (5, 3, 2000, 40, 40, 48, 64, torch.float16, False),
]


@pytest.mark.parametrize("case", PITCH_CASES)
def test_fused_attention_head_pitch(ops, case):
    """Heads stored at a column pitch below the kernel's tile width (head_stride): fewer K-steps, narrower P.V, the TMA
    boxes read into the neighbouring head / past the last head (zero filled) without ever multiplying those columns.
    q/k/v are views into one fused buffer, like the UNet's qkv projection output."""
    B, H, Lq, Lk, hd, pt, tile, dt, causal = case
    torch.manual_seed(12)
    L = max(Lq, Lk)
    buf = torch.zeros(B, L, 3, H, pt, device=dev)
    buf[..., :hd] = torch.randn(B, L, 3, H, hd, device=dev)
    use_ones = dt == torch.float16
    if use_ones:
        buf[:, :, 2, :, hd] = 1.0
    buf = buf.view(B, L, 3 * H * pt).to(dt)
    q, k, v = buf[:, :Lq, : H * pt], buf[:, :Lk, H * pt: 2 * H * pt], buf[:, :Lk, 2 * H * pt:]
    got = ops.attention(q, k, v, H, tile, hd ** -0.5, causal=causal, ones_col=hd if use_ones else 0, head_stride=pt)
    ref = attn_ref(q, k, v, H, pt, hd ** -0.5, causal)
    assert got.shape == (B, Lq, H * pt)
    tol = 1e-2 if dt == torch.bfloat16 else 3e-3
    assert torch.isfinite(got.float()).all() and rel(got, ref) < tol
    if use_ones:
        assert (got.view(B, Lq, H, pt)[..., hd].float() - 1).abs().max() < 2e-3
    assert got.view(B, Lq, H, pt)[..., hd + 1:].abs().max() == 0            # padding columns stay exactly zero


@pytest.mark.parametrize("hd,hp,dt", [(40, 64, torch.float16), (80, 128, torch.float16), (128, 128, torch.bfloat16)])
def test_fused_attention_growing_max(ops, hd, hp, dt):
    """Key norms grow along the sequence, so the running row max rises by far more than the lazy-rescale threshold
    (2^8) several times: exercises the single-pass softmax's redo + O-rescale path."""
    torch.manual_seed(13)
    B, H, L = 2, 4, 1024

    def mk(scale_rows=None):
        t = torch.zeros(B, L, H, hp, device=dev)
        x = torch.randn(B, L, H, hd, device=dev)
        if scale_rows is not None:
            x = x * scale_rows.view(1, L, 1, 1)
        t[..., :hd] = x
        return t

    q, k, v = mk(), mk(torch.linspace(0.05, 12.0, L, device=dev)), mk()
    oc = 0
    if hp > hd and dt == torch.float16:
        v[..., hd] = 1.0
        oc = hd
    q, k, v = (t.view(B, L, H * hp).to(dt) for t in (q, k, v))
    got = ops.attention(q, k, v, H, hp, hd ** -0.5, ones_col=oc)
    ref = attn_ref(q, k, v, H, hp, hd ** -0.5)
    assert torch.isfinite(got.float()).all() and rel(got, ref) < (1e-2 if dt == torch.bfloat16 else 3e-3)


@pytest.mark.parametrize("hd,pt,tile,dt", [(40, 48, 64, torch.float16), (80, 96, 128, torch.float16), (32, 48, 64, torch.bfloat16)])
def test_attn4_growing_max_kv_lens_and_ragged_tiles(ops, hd, pt, tile, dt):
    """The decoupled-S kernels (48-key tiles at pitch 48, 64-key tiles at pitch 96): rescale path (row max rising far beyond
    2^8 along the keys), per-sample key counts that end inside a tile, Lq not a multiple of 128."""
    torch.manual_seed(14)
    B, H, Lq, Lk = 3, 4, 333, 500
    buf = torch.zeros(B, Lk, 3, H, pt, device=dev)
    x = torch.randn(B, Lk, 3, H, hd, device=dev)
    x[:, :, 1] *= torch.linspace(0.05, 12.0, Lk, device=dev).view(1, Lk, 1, 1)      # key norms grow along the sequence
    buf[..., :hd] = x
    use_ones = dt == torch.float16
    if use_ones:
        buf[:, :, 2, :, hd] = 1.0
    buf = buf.view(B, Lk, 3 * H * pt).to(dt)
    q, k, v = buf[:, :Lq, : H * pt], buf[:, :, H * pt: 2 * H * pt], buf[:, :, 2 * H * pt:]
    kvl = torch.tensor([500, 49, 257], device=dev, dtype=torch.int32)
    got = ops.attention(q, k, v, H, tile, hd ** -0.5, ones_col=hd if use_ones else 0, head_stride=pt, kv_lens=kvl)
    ref = attn_ref(q, k, v, H, pt, hd ** -0.5, kv_lens=kvl)
    tol = 1e-2 if dt == torch.bfloat16 else 3e-3
    assert torch.isfinite(got.float()).all() and rel(got, ref) < tol
    got2 = ops.attention(q, k, v, H, tile, hd ** -0.5, ones_col=hd if use_ones else 0, head_stride=pt)
    assert rel(got2, attn_ref(q, k, v, H, pt, hd ** -0.5)) < tol


@pytest.mark.parametrize("hd,pt,tile,Lk", [(40, 48, 64, 96), (80, 96, 128, 128)])
def test_attn4q_per_sample_key_counts(ops, hd, pt, tile, Lk):
    """attn4q (several query tiles per CTA, K / V resident) with per-sample key counts: one or two key tiles per sample."""
    torch.manual_seed(15)
    B, H, Lq = 4, 4, 1500
    L = max(Lq, Lk)
    buf = torch.zeros(B, L, 3, H, pt, device=dev)
    buf[..., :hd] = torch.randn(B, L, 3, H, hd, device=dev)
    buf[:, :, 2, :, hd] = 1.0
    buf = buf.view(B, L, 3 * H * pt).half()
    q, k, v = buf[:, :Lq, : H * pt], buf[:, :Lk, H * pt: 2 * H * pt], buf[:, :Lk, 2 * H * pt:]
    kvl = torch.tensor([Lk, 77, 20, Lk // 2 + 1], device=dev, dtype=torch.int32)
    got = ops.attention(q, k, v, H, tile, hd ** -0.5, ones_col=hd, head_stride=pt, kv_lens=kvl)
    ref = attn_ref(q, k, v, H, pt, hd ** -0.5, kv_lens=kvl)
    assert torch.isfinite(got.float()).all() and rel(got, ref) < 3e-3


def test_fused_attention_kv_lens_and_views(ops):
    torch.manual_seed(11)
    B, H, L, hp = 3, 4, 200, 128
    qkv = torch.randn(B, L, 3 * H * hp, device=dev).bfloat16()
    q, k, v = qkv[:, :, : H * hp], qkv[:, :, H * hp: 2 * H * hp], qkv[:, :, 2 * H * hp:]
    kvl = torch.tensor([200, 77, 130], device=dev, dtype=torch.int32)
    got = ops.attention(q, k, v, H, hp, hp ** -0.5, kv_lens=kvl)
    assert rel(got, attn_ref(q, k, v, H, hp, hp ** -0.5, kv_lens=kvl)) < 1e-2


def test_attn_small_f32(ops):
    torch.manual_seed(12)
    B = 3
    for Lq, Lk in ((77, 8), (77, 77), (8, 8)):
        q = torch.randn(B, Lq, 512, device=dev)
        kv = torch.randn(B, Lk, 1024, device=dev)
        out = torch.empty(B, Lq, 512, device=dev)
        ops.attn_small_f32(q, kv[:, :, :512], kv[:, :, 512:], 4, 128 ** -0.5, out=out)
        qh = q.view(B, Lq, 4, 128).transpose(1, 2)
        kh = kv[:, :, :512].reshape(B, Lk, 4, 128).transpose(1, 2)
        vh = kv[:, :, 512:].reshape(B, Lk, 4, 128).transpose(1, 2)
        ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 128 ** -0.5, -1) @ vh).transpose(1, 2).reshape(B, Lq, 512)
        assert rel(out, ref) < 1e-5


def test_bad_arguments_raise(ops):
    from gill_b200._lib import GillB200Error

    a = torch.randn(16, 30, device=dev).bfloat16()      # K stride not a multiple of 8 elements
    b = torch.randn(16, 30, device=dev).bfloat16()
    with pytest.raises(GillB200Error):
        ops.gemm(a, b)
    with pytest.raises(GillB200Error):
        ops.topk_scores(torch.randn(64, 64, device=dev).bfloat16(), torch.randn(2, 64, device=dev).bfloat16(), 17)


@pytest.mark.parametrize("hw", [(512, 512), (300, 400), (96, 128)])
def test_clip_preprocess_u8_is_pil_exact(ops, hw):
    """Device pre-processing of generated images for the re-rank step == PIL `img.resize((224, 224))` bit for bit, and
    == the HF CLIP normalisation to fp32 rounding (gill/models.py:733-737, gill/utils.py:117-119)."""
    import numpy as np
    from PIL import Image
    from oracle import clip as oclip

    h, w = hw
    rng = np.random.default_rng(1)
    imgs = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
    imgs[:, : h // 3] = (np.linspace(0, 255, w)[None, None, :, None]).astype(np.uint8)
    pv, rz = ops.clip_preprocess_u8(torch.from_numpy(imgs).to(dev), 224, out_dtype=torch.float32, return_resized=True)
    for b in range(3):
        ref = np.asarray(Image.fromarray(imgs[b]).resize((224, 224)).convert("RGB"))
        assert np.array_equal(rz[b].cpu().numpy(), ref)
        assert (pv[b].cpu() - oclip.clip_preprocess(imgs[b], 224)).abs().max() < 1e-6
    pv16 = ops.clip_preprocess_u8(torch.from_numpy(imgs).to(dev), 224)
    assert pv16.dtype == torch.bfloat16 and torch.equal(pv16, pv.bfloat16())


def _slab_stats(out2d):
    o = out2d.float().view(out2d.shape[0] // 32, 32, out2d.shape[1])
    return torch.stack([o.sum(1), (o * o).sum(1)], -1)


@pytest.mark.parametrize("hw", [(300, 400), (500, 333), (224, 224), (640, 480), (231, 517)])
def test_clip_feature_extractor_on_device_is_pil_exact(ops, hw):
    """Image PROMPTS / bank images (gill/utils.py:117-119): resize shortest edge to 224 (PIL 8-bit bicubic), centre crop,
    rescale, normalise -- the device kernel's cropped uint8 image is bit-identical to PIL, pixel_values match the oracle."""
    from PIL import Image
    from oracle import clip as oclip

    h, w = hw
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
    img[:, : h // 3] = (np.linspace(0, 255, w)[None, None, :, None]).astype(np.uint8)
    pv, rz = ops.clip_preprocess_u8(torch.from_numpy(img).to(dev), 224, out_dtype=torch.float32, return_resized=True,
                                    mode="feature_extractor")
    RH, RW, top, left = oclip.hf_clip_geometry(h, w)
    for b in range(2):
        ref = np.asarray(Image.fromarray(img[b]).resize((RW, RH), resample=Image.BICUBIC))[top:top + 224, left:left + 224]
        assert np.array_equal(rz[b].cpu().numpy(), ref)
        opv, _ = oclip.clip_feature_extractor(img[b])
        assert torch.allclose(pv[b].cpu(), opv, atol=1e-6)


def test_epilogue_groupnorm_statistics(ops):
    """stats=True: the GEMM / conv epilogue leaves {sum, sumsq} per 32-row slab and column of the ROUNDED output -- every
    staged-epilogue variant, the CTA-pair kernel and the stream-K finisher."""
    torch.manual_seed(30)
    M, N, K = 2048, 640, 320
    a = torch.randn(M, K, device=dev).half()
    b = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).half()
    rb = torch.randn(M // 256, N, device=dev)
    for kw in (dict(), dict(residual=res), dict(rowbias=rb, rows_per_group=256), dict(act="silu"),
               dict(residual=res, cta_pair=2, block_n=128), dict(stream_k=2)):
        out = ops.gemm(a, b, bias=bias, stats=True, **kw)
        assert out.gn_stats.shape == (M // 32, N, 2)
        ref = _slab_stats(out)
        assert rel(out.gn_stats, ref) < 1e-5, kw
    ob = ops.gemm(a.bfloat16(), b.bfloat16(), bias=bias, stats=True)
    assert rel(ob.gn_stats, _slab_stats(ob)) < 1e-5
    x = torch.randn(2, 16, 16, 1280, device=dev).half()                       # 20 tiles x 180 k-blocks: stream-K finisher
    w = (torch.randn(1280, 9 * 1280, device=dev) * 0.01).half()
    oc = ops.conv3x3(x, w, bias=torch.randn(1280, device=dev), stats=True, stream_k=2)
    assert rel(oc.gn_stats, _slab_stats(oc.view(-1, 1280))) < 1e-5
    with pytest.raises(Exception):
        ops.gemm(a, b, bias=bias, stats=True, out_dtype=torch.float32)


@pytest.mark.parametrize("case", [(2, 64, 64, 320, 0, True), (2, 32, 32, 640, 320, True), (2, 8, 8, 1280, 1280, False),
                                  (2, 16, 16, 1280, 640, True), (2, 16, 16, 640, 0, True), (3, 8, 8, 1280, 0, True),
                                  (2, 32, 32, 320, 0, False), (16, 32, 32, 640, 0, True)])
def test_groupnorm_from_epilogue_statistics_equals_two_pass(ops, case):
    B, H, W, C0, C1, silu = case
    torch.manual_seed(31)

    def produce(C):
        a = torch.randn(B * H * W, 64, device=dev).half()
        wt = (torch.randn(C, 64, device=dev) * 0.2).half()
        o = ops.gemm(a, wt, bias=torch.randn(C, device=dev), stats=True)
        o4 = o.view(B, H, W, C)
        o4.gn_stats = o.gn_stats
        return o4

    x0 = produce(C0)
    x1 = produce(C1) if C1 else None
    C = C0 + C1
    w, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    fast = ops.groupnorm(x0, w, b, 32, 1e-5, silu=silu, x2=x1)
    ops.USE_EPILOGUE_GN_STATS = False
    try:
        slow = ops.groupnorm(x0, w, b, 32, 1e-5, silu=silu, x2=x1)
    finally:
        ops.USE_EPILOGUE_GN_STATS = True
    assert rel(fast, slow) < 2e-3 and (fast.float() - slow.float()).abs().max() < 2e-2
    xc = x0 if x1 is None else torch.cat([x0, x1], -1)
    ref = F.group_norm(xc.permute(0, 3, 1, 2).float(), 32, w, b, 1e-5)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
    assert rel(fast, ref) < 2e-3


@pytest.mark.parametrize("C,N,geglu", [(320, 1536, False), (640, 512, False), (1280, 2560, True), (320, 2560, True)])
def test_layernorm_folded_into_the_consuming_gemm(ops, C, N, geglu):
    """Producer leaves per-row partial sums (rowstats); the consumer GEMM with ln=... computes LayerNorm(x) W^T + b from
    the raw x: rstd * (x W'^T - mean * colsum(W')) + (b + W beta)."""
    torch.manual_seed(40)
    M = 1000                                                          # ragged last tile
    a0 = torch.randn(M, 64, device=dev).half()
    w0 = (torch.randn(C, 64, device=dev) * 0.3).half()
    res = (torch.randn(M, C, device=dev) * 2 + 0.7).half()            # non-zero mean: exercises the mean * colsum term
    x = ops.gemm(a0, w0, bias=torch.randn(C, device=dev), residual=res, rowstats=True)
    ref_stats = torch.stack([x.float().view(M, C // 32, 32).sum(-1), (x.float() ** 2).view(M, C // 32, 32).sum(-1)], -1)
    assert rel(x.ln_stats.permute(1, 0, 2), ref_stats) < 2e-3         # sums of the unrounded values vs the fp16 tensor
    gam, bet = torch.randn(C, device=dev), torch.randn(C, device=dev)
    W = (torch.randn(N, C, device=dev) * 0.05).half()
    b = torch.randn(N, device=dev)
    Wf = (W.float() * gam[None]).half()
    out = ops.gemm(x, Wf, bias=b + W.float() @ bet, act="geglu" if geglu else None,
                   ln=(x.ln_stats, Wf.float().sum(1).contiguous(), 1e-5))
    y = F.layer_norm(x.float(), (C,), gam, bet, 1e-5) @ W.float().T + b
    ref = y[:, 0::2] * F.gelu(y[:, 1::2]) if geglu else y
    assert out.shape == ref.shape and rel(out, ref) < 3e-3
    # and equals the unfolded two-kernel path to fp16 noise
    n = ops.layernorm(x, gam, bet, 1e-5)
    two = ops.gemm(n, W, bias=b, act="geglu" if geglu else None)
    assert rel(out, two) < 3e-3


@pytest.mark.parametrize("B,H,W,C,cout", [(2, 64, 64, 320, 4), (1, 32, 48, 128, 3), (2, 8, 8, 64, 4)])
def test_conv3x3_narrow_equals_conv2d(ops, B, H, W, C, cout):
    """conv_out as one plain GEMM onto per-tap products + nine-tap shifted sum == F.conv2d(padding=1)."""
    torch.manual_seed(77)
    x = torch.randn(B, H, W, C, device=dev).half()
    w4 = (torch.randn(cout, C, 3, 3, device=dev) * 0.05).half()
    bias = torch.randn(cout, device=dev)
    wk = w4.permute(0, 2, 3, 1).reshape(cout, 9 * C).contiguous()        # k = (ky*3+kx)*C + c, as sd._conv_w lays it out
    got = ops.conv3x3_narrow(x, ops.conv3x3_taps_weight(wk, cout), cout, bias=bias)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w4.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert got.shape == ref.shape and got.dtype == torch.float16
    assert rel(got, ref) < 1e-3
    assert (got.float() - ref).abs().max() < 2e-2


@pytest.mark.parametrize("case", [(2, 64, 64, 64, 320, 320, "res"), (3, 32, 32, 128, 640, 320, "rowbias"), (2, 32, 32, 128, 320, 256, "res"),
                                  (2, 32, 32, 128, 256, 256, "stats"), (5, 16, 16, 192, 320, 160, "res"), (2, 48, 32, 128, 128, 128, "stats"),
                                  (16, 16, 16, 640, 640, 0, "res")])
def test_conv3x3_halo_tile(ops, case):
    """A_CONV3X3_HALO: one 18 x 10 pixel halo tile per 64-channel block feeds all nine taps through shifted descriptors;
    M tiles are 16 x 8 pixel blocks stored through 3-D tensor maps. Against F.conv2d and against the one-box-per-tap form
    (separate process-wide switch is read once, so the comparison launch forces the 1-CTA kernel instead)."""
    B, H, W, C, Co, bn, mode = case
    torch.manual_seed(6)
    x = torch.randn(B, H, W, C, device=dev).half()
    w = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half()
    bias = torch.randn(Co, device=dev)
    wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    kw = {}
    if mode == "res":
        res = torch.randn(B, H, W, Co, device=dev).half()
        kw["residual"] = res
        ref = ref + res.float()
    elif mode == "rowbias":
        rb = torch.randn(B, Co, device=dev)
        kw["rowbias"] = rb
        ref = ref + rb[:, None, None, :]
    got = ops.conv3x3(x, wk, bias=bias, block_n=bn, cta_pair=2 if bn else 0, stats=(mode == "stats"), **kw)
    assert got.dtype == torch.float16 and rel(got, ref) < 2e-3
    one_cta = ops.conv3x3(x, wk, bias=bias, block_n=128, cta_pair=1, stream_k=1, **kw)     # per-tap boxes, 1-CTA kernel
    assert rel(got, one_cta) < 1e-3
    if mode == "stats":
        st = got.gn_stats.view(B, H * W // 32, Co, 2).sum(1)
        g = got.float().view(B, H * W, Co)
        assert rel(st[..., 0], g.sum(1)) < 1e-3 and rel(st[..., 1], (g * g).sum(1)) < 1e-3
        w2, b2 = torch.randn(Co, device=dev), torch.randn(Co, device=dev)
        gn = ops.groupnorm(got, w2, b2, 32, 1e-5, silu=True)
        gref = F.silu(F.group_norm(got.permute(0, 3, 1, 2).float(), 32, w2, b2, 1e-5)).permute(0, 2, 3, 1)
        assert rel(gn, gref) < 2e-3


@pytest.mark.parametrize("case", [(16, 16, 16, 128, 320, True), (4, 48, 32, 64, 128, False), (4, 32, 32, 192, 640, True),
                                  (1, 64, 64, 64, 256, True)])
def test_conv3x3_up2_equals_upsample_then_conv(ops, case):
    """Nearest 2x upsample + 3x3 conv as four 2 x 2 phase convolutions of the low-res tensor (pre-summed weights, halo-tile
    kernel, strided output maps, remapped GroupNorm-statistics slabs) == F.conv2d(F.interpolate(x, 2x nearest))."""
    B, H, W, C, Co, stats = case
    torch.manual_seed(8)
    x = torch.randn(B, H, W, C, device=dev).half()
    w4 = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half()
    bias = torch.randn(Co, device=dev)
    wk = w4.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    assert ops.conv3x3_up2_supported(x)
    got = ops.conv3x3_up2(x, ops.conv3x3_up2_weights(wk), bias=bias, stats=stats)
    up = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, w4.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert got.shape == ref.shape and got.dtype == torch.float16
    assert rel(got, ref) < 2e-3
    two = ops.conv3x3(ops.upsample2x(x), wk, bias=bias)
    assert rel(got, two) < 2e-3
    if stats:
        st = got.gn_stats.view(B, 4 * H * W // 32, Co, 2).sum(1)
        g = got.float().view(B, 4 * H * W, Co)
        assert rel(st[..., 0], g.sum(1)) < 1e-3 and rel(st[..., 1], (g * g).sum(1)) < 1e-3
        w2, b2 = torch.randn(Co, device=dev), torch.randn(Co, device=dev)
        gn = ops.groupnorm(got, w2, b2, 32, 1e-5, silu=True)
        gref = F.silu(F.group_norm(got.permute(0, 3, 1, 2).float(), 32, w2, b2, 1e-5)).permute(0, 2, 3, 1)
        assert rel(gn, gref) < 2e-3


@pytest.mark.parametrize("case", [(2, 64, 64, 128, 0, 320, "rowbias"), (4, 32, 32, 128, 64, 640, "res"), (16, 16, 16, 192, 128, 320, "none"),
                                  (2, 32, 32, 64, 0, 256, "res")])
def test_conv3x3_with_groupnorm_fused_into_the_input(ops, case):
    """norm -> SiLU -> conv3x3 with the GroupNorm applied on the halo tile inside the conv (scale / shift from the producers'
    statistics, zero padding after the normalisation, optional channel-concat second source) == GroupNorm kernel + conv."""
    B, H, W, C0, C1, Co, mode = case
    torch.manual_seed(9)

    def produce(C):
        a = torch.randn(B * H * W, 64, device=dev).half()
        wt = (torch.randn(C, 64, device=dev) * 0.2).half()
        o = ops.gemm(a, wt, bias=torch.randn(C, device=dev), stats=True)
        o4 = o.view(B, H, W, C)
        o4.gn_stats = o.gn_stats
        return o4

    x0 = produce(C0)
    x1 = produce(C1) if C1 else None
    C = C0 + C1
    gw, gb = torch.randn(C, device=dev), torch.randn(C, device=dev)
    w4 = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half()
    bias = torch.randn(Co, device=dev)
    wk = w4.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    kw = {}
    if mode == "res":
        kw["residual"] = torch.randn(B, H, W, Co, device=dev).half()
    elif mode == "rowbias":
        kw["rowbias"] = torch.randn(B, Co, device=dev)
    assert ops.conv3x3_gn_supported(x0, Co, x1, min_tiles=0)
    ss = ops.groupnorm_scale_shift(x0, gw, gb, 32, 1e-5, x2=x1)
    got = ops.conv3x3(x0, wk, bias=bias, gn=(ss, True), x2=x1, stats=True, **kw)
    xc = x0 if x1 is None else torch.cat([x0, x1], -1)
    n = F.silu(F.group_norm(xc.permute(0, 3, 1, 2).float(), 32, gw, gb, 1e-5))
    ref = F.conv2d(n, w4.float(), bias, padding=1).permute(0, 2, 3, 1)
    if mode == "res":
        ref = ref + kw["residual"].float()
    elif mode == "rowbias":
        ref = ref + kw["rowbias"][:, None, None, :]
    assert got.shape == ref.shape and rel(got, ref) < 3e-3
    two = ops.conv3x3(ops.groupnorm(x0, gw, gb, 32, 1e-5, silu=True, x2=x1), wk, bias=bias, **kw)
    assert rel(got, two) < 2e-3
    st = got.gn_stats.view(B, H * W // 32, Co, 2).sum(1)
    g = got.float().view(B, H * W, Co)
    assert rel(st[..., 0], g.sum(1)) < 1e-3
