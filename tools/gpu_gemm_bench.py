"""UNet linear shapes at B=16 (head pitch 48/96/176 layout): auto dispatch vs forced configs. Usage: python tools/gpu_gemm_bench.py [tag]"""
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
# (M, N, K, act, residual, count per eval)
gemms = [(65536, 2560, 320, "geglu", False, 5), (65536, 1152, 320, None, False, 5), (65536, 384, 320, None, False, 5), (65536, 320, 384, None, True, 10),
         (65536, 320, 320, None, True, 10), (65536, 320, 1280, None, True, 5), (65536, 320, 640, None, False, 2), (65536, 320, 960, None, False, 1),
         (16384, 5120, 640, "geglu", False, 5), (16384, 2304, 640, None, False, 5), (16384, 768, 640, None, False, 5), (16384, 640, 768, None, True, 10),
         (16384, 640, 640, None, True, 10), (16384, 640, 2560, None, True, 5), (16384, 640, 1280, None, False, 1), (16384, 640, 1920, None, False, 1),
         (4096, 10240, 1280, "geglu", False, 5), (4096, 4224, 1280, None, False, 5), (4096, 1408, 1280, None, False, 5), (4096, 1280, 1408, None, True, 10),
         (4096, 1280, 1280, None, True, 10), (4096, 1280, 5120, None, True, 5), (4096, 1280, 2560, None, False, 2), (1024, 10240, 1280, "geglu", False, 1),
         (1024, 1280, 5120, None, True, 1), (1024, 1280, 2560, None, False, 6)]
tot = {}
for (M, N, K, act, res, cnt) in gemms:
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half() * 0.05; bias = torch.randn(N, device=dev)
    n_out = N // 2 if act == "geglu" else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float16); r = torch.randn(M, n_out, device=dev).half() if res else None
    fl = 2.0 * M * N * K
    row = []
    for name, kw in (("auto", {}), ("wide320", dict(block_n=320)), ("bn256pair", dict(block_n=256, cta_pair=2)), ("bn256", dict(block_n=256, cta_pair=1)),
                     ("bn128", dict(block_n=128, cta_pair=1))):
        if name == "wide320" and (N % 320 or act): continue
        try:
            t = timeit(lambda: ops.gemm(a, b, out=out, bias=bias, residual=r, act=act, **kw))
            row.append(f"{name} {t:6.1f} ({fl / t / 1e6:4.0f})")
            tot[name] = tot.get(name, 0.0) + t * cnt
            if name == "auto": tot["auto_on_wide_shapes"] = tot.get("auto_on_wide_shapes", 0.0) + (t * cnt if (N % 320 == 0 and not act) else 0)
        except Exception as e:
            row.append(f"{name} ERR")
    print(f"{tag} M{M} N{N} K{K} {act or ''}{' +res' if res else ''} x{cnt}: " + " | ".join(row), flush=True)
print(tag, "weighted totals per UNet eval (ms) [wide320 covers only the N%320==0 shapes]:", {k: round(v / 1e3, 3) for k, v in tot.items()})
