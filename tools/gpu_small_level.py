"""8x8 / 16x16-level shapes (M = 1024 / 4096): timing under GILLB200_GEMM_DEBUG modes. Usage: python tools/gpu_small_level.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev, tag = "cuda", (sys.argv[1] if len(sys.argv) > 1 else "")
def timeit(fn, n=10):
    """GPU time per launch with the launches replayed from a CUDA graph (no CPU launch floor: an eager ctypes call costs ~15 us)."""
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with ops.graph_capture(g, dev):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3
for (B, HW, C, Co) in [(16, 8, 1280, 1280), (16, 8, 2560, 1280), (16, 16, 1280, 1280), (16, 16, 640, 640)]:
    x = torch.randn(B, HW, HW, C, device=dev).half(); w = torch.randn(Co, 9 * C, device=dev).half() * 0.02; bias = torch.randn(Co, device=dev)
    res = torch.randn(B, HW, HW, Co, device=dev).half(); out = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
    fl = 2.0 * B * HW * HW * Co * 9 * C
    row = []
    for name, kw in (("auto", {}), ("nosk256", dict(block_n=256, cta_pair=1, stream_k=1)), ("nosk128", dict(block_n=128, cta_pair=1, stream_k=1)),
                     ("sk128", dict(block_n=128, stream_k=2)), ("pair256", dict(block_n=256, cta_pair=2)), ("splitk", dict(stream_k=3))):
        try:
            t = timeit(lambda: ops.conv3x3(x, w, out=out, bias=bias, residual=res, stats=True, **kw))
            row.append(f"{name} {t:6.1f} ({fl / t / 1e6:4.0f})")
        except Exception as e:
            row.append(f"{name} ERR {str(e)[:30]}")
    print(f"{tag} conv B{B} {HW}x{HW} C{C}->{Co}: " + " | ".join(row), flush=True)
for (M, N, K, res) in [(1024, 1280, 5120, True), (1024, 1280, 2560, False), (1024, 1280, 1280, True), (4096, 1280, 1280, True), (4096, 1408, 1280, False)]:
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half() * 0.05; bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float16); r = torch.randn(M, N, device=dev).half() if res else None
    fl = 2.0 * M * N * K
    row = []
    for name, kw in (("auto", {}), ("bn256", dict(block_n=256, cta_pair=1)), ("bn128", dict(block_n=128, cta_pair=1)), ("bn64", dict(block_n=64, cta_pair=1)),
                     ("sk256", dict(block_n=256, stream_k=2)), ("sk128", dict(block_n=128, stream_k=2))):
        try:
            t = timeit(lambda: ops.gemm(a, b, out=out, bias=bias, residual=r, **kw))
            row.append(f"{name} {t:6.1f} ({fl / t / 1e6:4.0f})")
        except Exception as e:
            row.append(f"{name} ERR {str(e)[:30]}")
    print(f"{tag} gemm M{M} N{N} K{K}{' +res' if res else ''}: " + " | ".join(row), flush=True)
# launch floor: a trivially small GEMM and the norms at the 8x8 level
a = torch.randn(128, 64, device=dev).half(); b = torch.randn(64, 64, device=dev).half(); out = torch.empty(128, 64, device=dev, dtype=torch.float16)
print(tag, "tiny gemm 128x64x64: %.1f us" % timeit(lambda: ops.gemm(a, b, out=out)))
x = torch.randn(16, 8, 8, 1280, device=dev).half(); w = torch.randn(1280, device=dev); bb = torch.randn(1280, device=dev)
print(tag, "groupnorm 8x8 C1280 (stats kernel + apply): %.1f us" % timeit(lambda: ops.groupnorm(x, w, bb, 32, 1e-5, silu=True)))
x2 = torch.randn(1024, 1280, device=dev).half()
print(tag, "layernorm rows1024 C1280: %.1f us" % timeit(lambda: ops.layernorm(x2, w, bb, 1e-5)))

# GroupNorm(+SiLU) from the producer's statistics at every UNet level (one fused launch for HW <= 1024, else finalize + apply)
for (HW, C0, C1) in [(8, 1280, 0), (8, 1280, 1280), (16, 1280, 0), (16, 1280, 1280), (16, 640, 0), (32, 640, 0), (32, 640, 640), (32, 320, 0), (64, 320, 0), (64, 320, 320)]:
    def produce(C):
        a = torch.randn(16 * HW * HW, 64, device=dev).half(); wt = (torch.randn(C, 64, device=dev) * 0.2).half()
        o = ops.gemm(a, wt, stats=True); o4 = o.view(16, HW, HW, C); o4.gn_stats = o.gn_stats
        return o4
    x0 = produce(C0); x1 = produce(C1) if C1 else None
    C = C0 + C1
    w = torch.randn(C, device=dev); bb = torch.randn(C, device=dev); out = torch.empty(16, HW, HW, C, device=dev, dtype=torch.float16)
    t = timeit(lambda: ops.groupnorm(x0, w, bb, 32, 1e-5, silu=True, x2=x1, out=out))
    print(f"{tag} groupnorm(from stats) B16 {HW}x{HW} C{C0}+{C1}: {t:6.1f} us  {4.0 * 16 * HW * HW * C / t / 1e3:6.0f} GB/s", flush=True)

for (rows, C) in [(65536, 320), (16384, 640), (4096, 1280), (1024, 1280)]:
    xl = torch.randn(rows, C, device=dev).half(); wl = torch.randn(C, device=dev); bl = torch.randn(C, device=dev); ol = torch.empty_like(xl)
    t = timeit(lambda: ops.layernorm(xl, wl, bl, 1e-5, out=ol))
    print(f"{tag} layernorm rows{rows} C{C}: {t:6.1f} us  {4.0 * rows * C / t / 1e3:6.0f} GB/s", flush=True)
