"""GPU check of the CTA-pair (cta_group::2) GEMM / conv kernel against torch, plus timing vs the 1-CTA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from gill_b200 import ops
dev = "cuda"; torch.manual_seed(0)
def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
ok = True
def rep(name, r, tol=1e-5):
    global ok
    good = r < tol; ok &= good
    print(f"[{'OK ' if good else 'BAD'}] {name}: rel={r:.3e}", flush=True)
for (M, N, K, bn) in [(512, 512, 256, 128), (512, 512, 256, 256), (512, 640, 256, 160), (512, 256, 128, 64), (300, 200, 200, 128),
                      (100, 512, 320, 256), (4096, 4096, 1024, 256), (65536, 320, 320, 160), (1000, 328, 72, 160)]:
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    got = ops.gemm(a, b, block_n=bn, cta_pair=2, out_dtype=torch.float32); torch.cuda.synchronize()
    rep(f"pair gemm {M}x{N}x{K} bn{bn}", rel(got, a.float() @ b.float().T))
M, N, K = 1024, 640, 320
a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half(); bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev).half()
got = ops.gemm(a, b, bias=bias, residual=res, act="silu", cta_pair=2, block_n=160, out_dtype=torch.float32); torch.cuda.synchronize()
rep("pair epilogue bias+silu+res", rel(got, F.silu(a.float() @ b.float().T + bias) + res.float()))
a2 = torch.randn(M, 128, device=dev).half(); bc = torch.randn(N, K + 128, device=dev).half()
got = ops.gemm(a, bc, a2=a2, a2_mode=1, cta_pair=2, block_n=128, out_dtype=torch.float32); torch.cuda.synchronize()
rep("pair a2 concat", rel(got, torch.cat([a, a2], 1).float() @ bc.float().T))
for (B, H, W, C, Co, bn) in [(2, 64, 64, 64, 160, 160), (2, 32, 32, 128, 320, 160), (4, 8, 8, 128, 256, 256), (2, 16, 16, 64, 128, 128), (1, 128, 128, 64, 128, 128)]:
    x = torch.randn(B, H, W, C, device=dev).half(); w = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half(); bias = torch.randn(Co, device=dev)
    wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    got = ops.conv3x3(x, wk, bias=bias, out_dtype=torch.float32, block_n=bn, cta_pair=2); torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    rep(f"pair conv3x3 B{B} {H}x{W} C{C}->{Co}", rel(got.reshape(-1, Co), ref.reshape(-1, Co)), 1e-4)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K, bn) in [(8192, 8192, 8192, 256), (65536, 320, 2880, 160), (65536, 1536, 320, 256), (65536, 2560, 320, 256), (16384, 640, 5760, 160), (640, 16384, 4096, 256)]:
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16(); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t1 = timeit(lambda: ops.gemm(a, b, out=out, block_n=bn, cta_pair=1)); t2 = timeit(lambda: ops.gemm(a, b, out=out, block_n=bn, cta_pair=2))
    print(f"perf {M}x{N}x{K} bn{bn}: 1-CTA {t1*1e3:.1f} us ({2*M*N*K/t1/1e9:.0f} TF/s)   pair {t2*1e3:.1f} us ({2*M*N*K/t2/1e9:.0f} TF/s)", flush=True)
x = torch.randn(16, 64, 64, 320, device=dev).half(); w = torch.randn(320, 2880, device=dev).half() * 0.02
t1 = timeit(lambda: ops.conv3x3(x, w, block_n=160, cta_pair=1)); t2 = timeit(lambda: ops.conv3x3(x, w, block_n=160, cta_pair=2))
fl = 2 * 65536 * 320 * 2880
print(f"perf conv 16x64x64 320->320: 1-CTA {t1*1e3:.1f} us ({fl/t1/1e9:.0f} TF/s)   pair {t2*1e3:.1f} us ({fl/t2/1e9:.0f} TF/s)", flush=True)
print("ALL OK" if ok else "SOME BAD")
