"""One GEMM / conv shape, a few launches: the target of `ncu --set full` captures.
   python tools/gpu_one_gemm.py M N K [geglu|none] [res] [bn] [pair] [order] [reps]
   python tools/gpu_one_gemm.py conv B HW C Co [bn] [pair] [order] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev = "cuda"
a = sys.argv[1:]
def iarg(i, d=0):
    return int(a[i]) if len(a) > i else d
if a[0] == "conv":
    B, HW, C, Co = map(int, a[1:5])
    bn, pair, order, reps = iarg(5), iarg(6), iarg(7), iarg(8, 5)
    x = torch.randn(B, HW, HW, C, device=dev).half(); w = torch.randn(Co, 9 * C, device=dev).half() * 0.02
    bias = torch.randn(Co, device=dev); out = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
    fn = lambda: ops.conv3x3(x, w, out=out, bias=bias, block_n=bn, cta_pair=pair, tile_order=order)
    fl = 2.0 * B * HW * HW * Co * 9 * C
else:
    M, N, K = map(int, a[0:3])
    act = a[3] if len(a) > 3 and a[3] != "none" else None
    res = len(a) > 4 and a[4] == "res"
    bn, pair, order, reps = iarg(5), iarg(6), iarg(7), iarg(8, 5)
    A = torch.randn(M, K, device=dev).half(); Bm = torch.randn(N, K, device=dev).half() * 0.05; bias = torch.randn(N, device=dev)
    n_out = N // 2 if act == "geglu" else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float16); r = torch.randn(M, n_out, device=dev).half() if res else None
    fn = lambda: ops.gemm(A, Bm, out=out, bias=bias, residual=r, act=act, block_n=bn, cta_pair=pair, tile_order=order)
    fl = 2.0 * M * N * K
for _ in range(2): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): fn()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / reps
print(f"{' '.join(a)}: {t*1e3:.1f} us  {fl/t/1e9:.0f} TF/s")
