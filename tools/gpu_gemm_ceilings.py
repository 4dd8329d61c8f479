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
    return e0.elapsed_time(e1) / n
mode = os.environ.get("GILLB200_GEMM_DEBUG", "0")
for (M, N, K, bn) in [(8192, 8192, 8192, 256), (8192, 8192, 8192, 128), (65536, 320, 2880, 160), (65536, 1536, 320, 256)]:
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16(); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.gemm(a, b, out=out, block_n=bn, cta_pair=1))
    print(f"debug={mode} {M}x{N}x{K} bn{bn}: {t*1e3:.1f} us  ({2*M*N*K/t/1e9:.0f} TF/s-equivalent)", flush=True)
