"""Small driver for ncu captures: a few launches of the hot kernels at their UNet shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev = "cuda"
which = sys.argv[1]
torch.manual_seed(0)
if which == "attn":
    B, H, L, hd, hp = 16, 8, 4096, 40, 64
    q = torch.randn(B, L, H * hp, device=dev).half(); k = torch.randn_like(q); v = torch.randn_like(q)
    for _ in range(2): ops.attention(q, k, v, H, hp, hd ** -0.5)
    v.view(B, L, H, hp)[..., hd] = 1.0
    for _ in range(2): ops.attention(q, k, v, H, hp, hd ** -0.5, ones_col=hd)
elif which == "geglu":
    a = torch.randn(65536, 320, device=dev).half(); b = torch.randn(2560, 320, device=dev).half(); bias = torch.randn(2560, device=dev)
    for _ in range(3): ops.gemm(a, b, bias=bias, act="geglu")
elif which == "gemm320":
    a = torch.randn(65536, 320, device=dev).half(); b = torch.randn(320, 320, device=dev).half(); bias = torch.randn(320, device=dev)
    r = torch.randn(65536, 320, device=dev).half()
    for _ in range(3): ops.gemm(a, b, bias=bias, residual=r)
elif which == "conv":
    x = torch.randn(16, 64, 64, 320, device=dev).half(); w = torch.randn(320, 2880, device=dev).half() * 0.02; bias = torch.randn(320, device=dev)
    for _ in range(3): ops.conv3x3(x, w, bias=bias)
elif which == "qkv":
    a = torch.randn(65536, 320, device=dev).half(); b = torch.randn(1536, 320, device=dev).half()
    for _ in range(3): ops.gemm(a, b)
elif which == "conv8":
    x = torch.randn(16, 8, 8, 1280, device=dev).half(); w = torch.randn(1280, 11520, device=dev).half() * 0.01; bias = torch.randn(1280, device=dev)
    for _ in range(3): ops.conv3x3(x, w, bias=bias)
elif which == "big":
    a = torch.randn(8192, 8192, device=dev).bfloat16(); b = torch.randn(8192, 8192, device=dev).bfloat16()
    for _ in range(3): ops.gemm(a, b, block_n=256)
elif which == "gn":
    x = torch.randn(16, 64, 64, 320, device=dev).half(); w = torch.randn(320, device=dev); b = torch.randn(320, device=dev)
    for _ in range(3): ops.groupnorm(x, w, b, 32, 1e-5, silu=True)
elif which in ("attn48", "xattn48"):
    # round-2 layout: head pitch 48 (hd 40), V's ones column, P through tensor memory
    B, H, L, hd, pt = 16, 8, 4096, 40, 48
    Lk = L if which == "attn48" else 77
    buf = torch.zeros(B, L, 3, H, pt, device=dev)
    buf[..., :hd] = torch.randn(B, L, 3, H, hd, device=dev)
    buf[:, :, 2, :, hd] = 1.0
    buf = buf.view(B, L, 3 * H * pt).half()
    q, k, v = buf[:, :, : H * pt], buf[:, :Lk, H * pt: 2 * H * pt], buf[:, :Lk, 2 * H * pt:]
    for _ in range(4): ops.attention(q, k, v, H, 64, hd ** -0.5, ones_col=hd, head_stride=pt)
elif which == "topk":
    from gill_b200 import retrieval
    bank = torch.randn(1_000_000, 768, device=dev).bfloat16(); q = torch.randn(1024, 768, device=dev).bfloat16()
    for _ in range(3): retrieval.retrieval_topk(bank, q, 16)
elif which == "topk1":
    from gill_b200 import retrieval
    bank = torch.randn(3_000_000, 256, device=dev).bfloat16(); q = torch.randn(1, 256, device=dev).bfloat16()
    for _ in range(3): retrieval.retrieval_topk(bank, q, 3, exclude_idx=[5, 9])
elif which == "convwide":
    x = torch.randn(16, 64, 64, 320, device=dev).half(); w = torch.randn(320, 2880, device=dev).half() * 0.02; bias = torch.randn(320, device=dev)
    for _ in range(3): ops.conv3x3(x, w, bias=bias, block_n=320)
elif which == "convgn":
    # resnet conv1 of an up block: GroupNorm + SiLU of cat([h, skip]) applied inside the halo-tile conv
    def produce(C):
        a = torch.randn(16 * 64 * 64, 64, device=dev).half(); wt = (torch.randn(C, 64, device=dev) * 0.2).half()
        o = ops.gemm(a, wt, stats=True); o4 = o.view(16, 64, 64, C); o4.gn_stats = o.gn_stats
        return o4
    x0, x1 = produce(320), produce(320)
    w = torch.randn(320, 9 * 640, device=dev).half() * 0.02; bias = torch.randn(320, device=dev)
    ss = ops.groupnorm_scale_shift(x0, torch.randn(640, device=dev), torch.randn(640, device=dev), 32, 1e-5, x2=x1)
    for _ in range(3): ops.conv3x3(x0, w, bias=bias, gn=(ss, True), x2=x1, stats=True)
elif which == "convup2":
    x = torch.randn(16, 32, 32, 640, device=dev).half(); w = torch.randn(640, 9 * 640, device=dev).half() * 0.02
    wph = ops.conv3x3_up2_weights(w); bias = torch.randn(640, device=dev)
    for _ in range(2): ops.conv3x3_up2(x, wph, bias=bias, stats=True)
torch.cuda.synchronize()
