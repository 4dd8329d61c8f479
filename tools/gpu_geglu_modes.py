"""The 64x64-level short-K linears under GILLB200_GEMM_DEBUG (0 normal, 1 no TMA = MMA + epilogue ceiling, 2 no MMA = fill + epilogue ceiling)."""
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
for (M, N, K, act, res) in [(65536, 2560, 320, "geglu", False), (65536, 2560, 320, None, False), (65536, 1152, 320, None, False), (65536, 320, 320, None, True),
                            (65536, 320, 1280, None, True), (16384, 5120, 640, "geglu", False)]:
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half() * 0.05; bias = torch.randn(N, device=dev)
    n_out = N // 2 if act == "geglu" else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float16); r = torch.randn(M, n_out, device=dev).half() if res else None
    row = []
    for name, kw in (("auto", {}), ("bn256", dict(block_n=256, cta_pair=1)), ("bn256pair", dict(block_n=256, cta_pair=2)), ("bn128", dict(block_n=128, cta_pair=1))):
        try:
            t = timeit(lambda: ops.gemm(a, b, out=out, bias=bias, residual=r, act=act, **kw))
            row.append(f"{name} {t:6.1f}")
        except Exception as e:
            row.append(f"{name} ERR")
    print(f"{tag} M{M} N{N} K{K} {act or ''}{' +res' if res else ''}: " + " | ".join(row), flush=True)
