"""Microbenchmark of the UNet's GroupNorm(+SiLU) / LayerNorm shapes at B=16. Usage: python tools/gpu_norm_bench.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gill_b200 import ops

dev, tag = "cuda", (sys.argv[1] if len(sys.argv) > 1 else "")


def bench(fn, n=30, w=5):
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


flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for (H, C0, C1, silu) in [(64, 320, 0, True), (64, 320, 0, False), (64, 320, 320, True), (64, 640, 320, True), (32, 640, 0, True),
                          (32, 640, 640, True), (16, 1280, 0, True), (16, 1280, 1280, True), (8, 1280, 1280, True)]:
    B, G = 16, 32
    torch.manual_seed(0)
    C = C0 + C1
    # producer GEMM leaves the statistics: emulate with a real GEMM epilogue (identity weights would be slow): use the two-pass entry
    x = torch.randn(B, H, H, C0, device=dev).half()
    x2 = torch.randn(B, H, H, C1, device=dev).half() if C1 else None
    w, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = torch.empty(B, H, H, C, device=dev, dtype=torch.float16)
    f = lambda: ops.groupnorm(x, w, b, G, 1e-5, silu=silu, x2=x2, out=out)
    us = bench(f)
    xx = torch.cat([x, x2], -1) if C1 else x
    ref = F.group_norm(xx.float().permute(0, 3, 1, 2), G, w, b, 1e-5)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    mb = 2 * B * H * H * C * 2 / 1e6
    print(f"{tag:>8s} groupnorm(two-pass) {H}x{H} C{C0}+{C1} silu={int(silu)}: {us:7.1f} us ({mb:6.1f} MB r+w; the two-pass entry reads x twice) rel {err:.1e}", flush=True)
for rows, C in [(65536, 320), (16384, 640), (4096, 1280)]:
    x = torch.randn(rows, C, device=dev).half()
    w, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = torch.empty_like(x)
    us = bench(lambda: ops.layernorm(x, w, b, 1e-5, out=out))
    ref = F.layer_norm(x.float(), (C,), w, b, 1e-5)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"{tag:>8s} layernorm rows{rows} C{C}: {us:7.1f} us  {2 * rows * C * 2 / us / 1e6:6.2f} TB/s rel {err:.1e}", flush=True)
