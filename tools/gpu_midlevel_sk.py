"""32x32 / 16x16-level + residual linears: auto vs forced stream-K / tile widths (graph-timed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev, tag = "cuda", (sys.argv[1] if len(sys.argv) > 1 else "")
def timeit(fn, n=10):
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
for (M, N, K, res) in [(16384, 640, 640, True), (16384, 640, 768, True), (16384, 768, 640, False), (4096, 1280, 1280, True), (4096, 1280, 1408, True), (4096, 1408, 1280, False),
                       (16384, 640, 2560, True), (4096, 1280, 5120, True), (65536, 320, 1280, True), (65536, 320, 320, True), (65536, 320, 384, True), (65536, 384, 320, False)]:
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half() * 0.05; bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float16); r = torch.randn(M, N, device=dev).half() if res else None
    fl = 2.0 * M * N * K
    row = []
    for name, kw in (("auto", {}), ("sk160", dict(block_n=160, stream_k=2)), ("sk256", dict(block_n=256, stream_k=2)), ("sk128", dict(block_n=128, stream_k=2)),
                     ("bn160", dict(block_n=160, cta_pair=1, stream_k=1)), ("bn128", dict(block_n=128, cta_pair=1, stream_k=1)), ("pair160", dict(block_n=160, cta_pair=2)),
                     ("pair128", dict(block_n=128, cta_pair=2)), ("wide320", dict(block_n=320))):
        if name == "wide320" and N % 320: continue
        try:
            t = timeit(lambda: ops.gemm(a, b, out=out, bias=bias, residual=r, **kw))
            row.append(f"{name} {t:5.1f}")
        except Exception as e:
            row.append(f"{name} ERR")
    print(f"{tag} M{M} N{N} K{K}{' +res' if res else ''}: " + " | ".join(row), flush=True)
