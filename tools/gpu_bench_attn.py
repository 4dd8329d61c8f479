"""Fused attention at the UNet / OPT shapes. GILLB200_ATTN=1 forces the warp-specialised kernel, =2 the 4-warp attn2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev = "cuda"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, H, Lq, Lk, hd, hp, dt, causal) in [(16, 8, 4096, 4096, 40, 64, torch.float16, False), (16, 8, 4096, 77, 40, 64, torch.float16, False),
                                            (16, 8, 1024, 1024, 80, 128, torch.float16, False), (16, 8, 1024, 77, 80, 128, torch.float16, False),
                                            (16, 8, 256, 256, 160, 192, torch.float16, False), (16, 8, 256, 77, 160, 192, torch.float16, False),
                                            (16, 8, 64, 64, 160, 192, torch.float16, False), (8, 32, 81, 81, 128, 128, torch.bfloat16, True)]:
    q = torch.randn(B, Lq, H * hp, device=dev).to(dt); k = torch.randn(B, Lk, H * hp, device=dev).to(dt); v = torch.randn(B, Lk, H * hp, device=dev).to(dt)
    oc = 0
    if hp > hd:
        for t in (q, k, v): t.view(B, -1, H, hp)[..., hd:] = 0
        v.view(B, Lk, H, hp)[..., hd] = 1.0
        oc = hd
    t = timeit(lambda: ops.attention(q, k, v, H, hp, hd ** -0.5, causal=causal, ones_col=oc))
    fl = 4.0 * B * H * Lq * Lk * hd * (0.5 if causal else 1.0)
    print(f"attn B{B} H{H} Lq{Lq} Lk{Lk} hd{hd}/{hp}: {t:.1f} us  {fl/t/1e6:.0f} TF/s (true head dim)")
