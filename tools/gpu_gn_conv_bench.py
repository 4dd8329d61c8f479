"""conv3x3 with / without the GroupNorm applied inside (graph-timed). GILLB200_GEMM_DEBUG=4: transform math skipped (sync cost only)."""
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
for (B, HW, C, Co) in [(16, 64, 320, 320), (16, 64, 640, 320), (16, 32, 640, 640), (16, 32, 1280, 640), (16, 16, 1280, 1280), (16, 16, 2560, 1280)]:
    a = torch.randn(B * HW * HW, 64, device=dev).half(); wt = (torch.randn(C, 64, device=dev) * 0.2).half()
    o = ops.gemm(a, wt, stats=True); x = o.view(B, HW, HW, C); x.gn_stats = o.gn_stats
    w = torch.randn(Co, 9 * C, device=dev).half() * 0.02; bias = torch.randn(Co, device=dev)
    gw, gb = torch.randn(C, device=dev), torch.randn(C, device=dev)
    res = torch.randn(B, HW, HW, Co, device=dev).half(); out = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
    ss = ops.groupnorm_scale_shift(x, gw, gb, 32, 1e-5)
    row = []
    t_plain = timeit(lambda: ops.conv3x3(x, w, out=out, bias=bias, residual=res, stats=True))
    t_gn = timeit(lambda: ops.groupnorm(x, gw, gb, 32, 1e-5, silu=True))
    for name, kw in (("fused auto", {}), ("fused wide320", dict(block_n=320)), ("fused bn160", dict(block_n=160)), ("fused bn256", dict(block_n=256))):
        try:
            t = timeit(lambda: ops.conv3x3(x, w, out=out, bias=bias, residual=res, stats=True, gn=(ss, True), **kw))
            row.append(f"{name} {t:6.1f}")
        except Exception as e:
            row.append(f"{name} ERR")
    print(f"{tag} B{B} {HW}x{HW} C{C}->{Co}: plain conv {t_plain:6.1f} + groupnorm {t_gn:5.1f} = {t_plain + t_gn:6.1f} us | " + " | ".join(row), flush=True)
