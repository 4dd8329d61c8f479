"""UNet implicit-conv shapes at B=16: auto dispatch vs forced tile configs. Usage: python tools/gpu_conv_bench.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev, tag = "cuda", (sys.argv[1] if len(sys.argv) > 1 else "")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
convs = [(16, 64, 320, 320, 7), (16, 64, 640, 320, 2), (16, 64, 960, 320, 1), (16, 64, 320, 4, 1), (16, 32, 640, 640, 6), (16, 32, 1280, 640, 1),
         (16, 32, 320, 640, 1), (16, 32, 1920, 640, 1), (16, 32, 960, 640, 1), (16, 32, 320, 320, 1), (16, 16, 1280, 1280, 7), (16, 16, 2560, 1280, 2),
         (16, 16, 640, 1280, 1), (16, 16, 1920, 1280, 1), (16, 16, 640, 640, 1), (16, 8, 1280, 1280, 12), (16, 8, 2560, 1280, 3)]
tot = {}
for (B, HW, C, Co, cnt) in convs:
    x = torch.randn(B, HW, HW, C, device=dev).half(); w = torch.randn(Co, 9 * C, device=dev).half() * 0.02; bias = torch.randn(Co, device=dev)
    res = torch.randn(B, HW, HW, Co, device=dev).half()
    out = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
    fl = 2.0 * B * HW * HW * Co * 9 * C
    row = []
    for name, kw in (("auto", {}), ("bn160pair", dict(block_n=160, cta_pair=2)), ("bn256pair", dict(block_n=256, cta_pair=2)), ("wide320", dict(block_n=320))):
        if Co % 320 and name == "wide320": continue
        if Co < 160 and name != "auto": continue
        try:
            t = timeit(lambda: ops.conv3x3(x, w, out=out, bias=bias, residual=res, stats=True, **kw))
            row.append(f"{name} {t:7.1f} us ({fl / t / 1e6:5.0f} TF/s)")
            tot[name] = tot.get(name, 0.0) + t * cnt
        except Exception as e:
            row.append(f"{name} ERR {str(e)[:40]}")
    print(f"{tag} conv B{B} {HW}x{HW} C{C}->{Co} x{cnt}: " + " | ".join(row), flush=True)
print(tag, "weighted totals per UNet eval (ms):", {k: round(v / 1e3, 3) for k, v in tot.items()})
