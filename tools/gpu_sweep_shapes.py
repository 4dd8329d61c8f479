"""Sweep tile width x (1-CTA | CTA pair) x tile order over the SD-1.5 UNet's GEMM / conv shapes at batch 16."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev = "cuda"
def timeit(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
gemms = [(65536, 2560, 320, "geglu", False), (65536, 1536, 320, None, False), (65536, 320, 512, None, True), (65536, 320, 320, None, True),
         (65536, 320, 1280, None, True), (16384, 5120, 640, "geglu", False), (16384, 3072, 640, None, False), (16384, 640, 1024, None, True),
         (16384, 640, 640, None, True), (16384, 640, 2560, None, True), (4096, 10240, 1280, "geglu", False), (4096, 1280, 1536, None, True),
         (4096, 1280, 1280, None, True), (4096, 1280, 5120, None, True), (4096, 4608, 1280, None, False), (1024, 1280, 1536, None, True),
         (1024, 10240, 1280, "geglu", False), (1024, 1280, 5120, None, True)]
convs = [(16, 64, 320, 320), (16, 64, 640, 320), (16, 64, 960, 320), (16, 32, 640, 640), (16, 32, 1280, 640), (16, 32, 320, 640), (16, 32, 1920, 640),
         (16, 16, 1280, 1280), (16, 16, 2560, 1280), (16, 16, 640, 1280), (16, 8, 1280, 1280), (16, 8, 2560, 1280)]
cfgs = [(bn, pair, order) for bn in (64, 128, 160, 256) for pair in (1, 2) for order in (1, 2)]
print("GEMM shapes: best configs (us) [bn, pair(1=single,2=pair), order(1=m-fastest,2=n-inner)]")
for (M, N, K, act, res) in gemms:
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half() * 0.05; bias = torch.randn(N, device=dev)
    n_out = N // 2 if act == "geglu" else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float16); r = torch.randn(M, n_out, device=dev).half() if res else None
    rs = []
    for (bn, pair, order) in cfgs:
        if N % bn and bn > N: continue
        try:
            t = timeit(lambda: ops.gemm(a, b, out=out, bias=bias, residual=r, act=act, block_n=bn, cta_pair=pair, tile_order=order))
            rs.append((t, bn, pair, order))
        except Exception as e:
            pass
    t0 = timeit(lambda: ops.gemm(a, b, out=out, bias=bias, residual=r, act=act))
    rs.sort()
    print(f"M{M} N{N} K{K} {act or ''}: auto {t0:.1f} | " + "  ".join(f"{t:.1f}[{bn},{pr},{o}]" for t, bn, pr, o in rs[:4]) + f" | worst {rs[-1][0]:.1f}", flush=True)
print("CONV shapes")
for (B, HW, C, Co) in convs:
    x = torch.randn(B, HW, HW, C, device=dev).half(); w = torch.randn(Co, 9 * C, device=dev).half() * 0.02; bias = torch.randn(Co, device=dev)
    out = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
    rs = []
    for (bn, pair, order) in cfgs:
        try:
            t = timeit(lambda: ops.conv3x3(x, w, out=out, bias=bias, block_n=bn, cta_pair=pair, tile_order=order))
            rs.append((t, bn, pair, order))
        except Exception as e:
            pass
    t0 = timeit(lambda: ops.conv3x3(x, w, out=out, bias=bias))
    rs.sort()
    fl = 2.0 * B * HW * HW * Co * 9 * C
    print(f"conv B{B} {HW}x{HW} C{C}->{Co}: auto {t0:.1f} ({fl/t0/1e6:.0f} TF/s) | " + "  ".join(f"{t:.1f}[{bn},{pr},{o}]" for t, bn, pr, o in rs[:4]) + f" | worst {rs[-1][0]:.1f}", flush=True)
