"""LayerNorm / GroupNorm at the UNet's shapes: achieved GB/s (algorithmic bytes: one read + one write of the activation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for rows, C in ((65536, 320), (16384, 640), (4096, 1280), (1024, 1280)):
    x = torch.randn(rows, C, device=dev).half(); w = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    out = torch.empty_like(x)
    t = timeit(lambda: ops.layernorm(x, w, b, 1e-5, out=out))
    print(f"layernorm rows{rows} C{C}: {t:.1f} us  {4*rows*C/t/1e3:.0f} GB/s")
for B, HW, C0, C1 in ((16, 64, 320, 0), (16, 64, 320, 320), (16, 64, 640, 320), (16, 32, 640, 0), (16, 32, 640, 640), (16, 32, 1280, 640),
                      (16, 16, 1280, 0), (16, 16, 1280, 1280), (16, 8, 1280, 0), (16, 8, 1280, 1280), (8, 512, 128, 0), (8, 256, 256, 0)):
    x = torch.randn(B, HW, HW, C0, device=dev).half(); x2 = torch.randn(B, HW, HW, C1, device=dev).half() if C1 else None
    C = C0 + C1
    w = torch.randn(C, device=dev); b = torch.randn(C, device=dev); out = torch.empty(B, HW, HW, C, device=dev, dtype=torch.float16)
    t = timeit(lambda: ops.groupnorm(x, w, b, 32, 1e-5, silu=True, x2=x2, out=out))
    print(f"groupnorm B{B} {HW}x{HW} C{C0}+{C1}: {t:.1f} us  {4*B*HW*HW*C/t/1e3:.0f} GB/s (1R+1W)")
