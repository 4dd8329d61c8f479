"""Microbenchmark of the UNet attention shapes (B=16, 8 heads). Usage: python tools/gpu_attn_bench.py [tag]
Environment toggles are read by the library once per process: run once per variant."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops

dev = "cuda"
tag = sys.argv[1] if len(sys.argv) > 1 else ""


def bench(fn, n=20, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def ref(q, k, v, H, pt, scale):
    B, Lq, _ = q.shape
    qh, kh, vh = (t.float().view(B, -1, H, pt).transpose(1, 2) for t in (q, k, v))
    return torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, scale=scale).transpose(1, 2).reshape(B, Lq, H * pt)


for (L, Lk, hd, pt, tile) in [(4096, 4096, 40, 64, 64), (4096, 4096, 40, 48, 64), (4096, 77, 40, 64, 64), (4096, 77, 40, 48, 64),
                              (1024, 1024, 80, 128, 128), (1024, 1024, 80, 96, 128), (1024, 77, 80, 96, 128),
                              (256, 256, 160, 192, 192), (256, 256, 160, 176, 192), (256, 77, 160, 176, 192)]:
    B, H = 16, 8
    torch.manual_seed(0)
    LL = max(L, Lk)
    buf = torch.zeros(B, LL, 3, H, pt, device=dev)
    buf[..., :hd] = torch.randn(B, LL, 3, H, hd, device=dev)
    buf[:, :, 2, :, hd] = 1.0
    buf = buf.view(B, LL, 3 * H * pt).half()
    q, k, v = buf[:, :L, : H * pt], buf[:, :Lk, H * pt: 2 * H * pt], buf[:, :Lk, 2 * H * pt:]
    out = torch.empty(B, L, H * pt, device=dev, dtype=torch.float16)
    f = lambda: ops.attention(q, k, v, H, tile, hd ** -0.5, ones_col=hd, head_stride=pt, out=out)
    us = bench(f)
    r = ref(q[:2], k[:2], v[:2], H, pt, hd ** -0.5)
    err = ((out[:2].float() - r).norm() / r.norm()).item()
    fl = 4.0 * B * H * L * Lk * hd
    print(f"{tag:>10s} L{L} Lk{Lk} hd{hd} pitch{pt} tile{tile}: {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s (algorithmic)  rel {err:.2e}", flush=True)
