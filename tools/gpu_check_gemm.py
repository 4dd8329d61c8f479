"""First-contact GPU check of the tcgen05 GEMM / implicit conv (run under gpurun; not a pytest)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from gill_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def report(name, got, ref, tol):
    r = rel(got, ref)
    mx = (got.float() - ref.float()).abs().max().item()
    ok = r < tol
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel={r:.3e} maxabs={mx:.3e}", flush=True)
    if not ok:
        d = (got.float() - ref.float()).abs()
        bad_rows = (d.max(dim=1).values > 10 * tol * ref.float().abs().max()).nonzero().flatten()[:16].tolist()
        bad_cols = (d.max(dim=0).values > 10 * tol * ref.float().abs().max()).nonzero().flatten()[:16].tolist()
        print("     first bad rows", bad_rows, "first bad cols", bad_cols, flush=True)
        print("     got[0,:8]", got[0, :8].float().tolist(), flush=True)
        print("     ref[0,:8]", ref[0, :8].float().tolist(), flush=True)
    return ok


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    allok = True
    for dt in (torch.bfloat16, torch.float16):
        for bn in (128, 32, 64, 160, 256):
            M, N, K = 256, 512, 256
            a = torch.randn(M, K, device=dev).to(dt)
            b = torch.randn(N, K, device=dev).to(dt)
            got = ops.gemm(a, b, block_n=bn, out_dtype=torch.float32)
            torch.cuda.synchronize()
            allok &= report(f"gemm {dt} bn={bn} {M}x{N}x{K}", got, a.float() @ b.float().T, 1e-5)
    # ragged sizes, K tail, multi-tile persistence
    for (M, N, K) in [(300, 200, 200), (1000, 328, 72), (4096, 4096, 1024), (77, 768, 512), (8, 512, 4096)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        got = ops.gemm(a, b, out_dtype=torch.float32)
        torch.cuda.synchronize()
        allok &= report(f"gemm ragged {M}x{N}x{K}", got, a.float() @ b.float().T, 1e-5)
    # epilogues
    M, N, K = 384, 640, 320
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16()
    ref = torch.relu(0.5 * (a.float() @ b.float().T) + bias) + res.float()
    got = ops.gemm(a, b, bias=bias, residual=res, act="relu", alpha=0.5)
    torch.cuda.synchronize()
    allok &= report("epilogue bias+relu+res bf16 out", got, ref, 5e-3)
    got = ops.gemm(a, b, bias=bias, residual=res.float(), act="relu", alpha=0.5, out_dtype=torch.float32)
    torch.cuda.synchronize()
    allok &= report("epilogue bias+relu+res f32 out", got, ref, 1e-5)
    for act, fn in (("gelu", F.gelu), ("silu", F.silu)):
        got = ops.gemm(a, b, bias=bias, act=act, out_dtype=torch.float32)
        torch.cuda.synchronize()
        allok &= report(f"epilogue {act}", got, fn(a.float() @ b.float().T + bias), 1e-5)
    # geglu: interleaved (value, gate) rows
    val, gate = torch.randn(N // 2, K, device=dev).bfloat16(), torch.randn(N // 2, K, device=dev).bfloat16()
    bi = torch.stack([val, gate], 1).reshape(N, K).contiguous()
    bv, bg = torch.randn(N // 2, device=dev), torch.randn(N // 2, device=dev)
    bias_i = torch.stack([bv, bg], 1).reshape(N).contiguous()
    ref = (a.float() @ val.float().T + bv) * F.gelu(a.float() @ gate.float().T + bg)
    got = ops.gemm(a, bi, bias=bias_i, act="geglu", out_dtype=torch.float32)
    torch.cuda.synchronize()
    allok &= report("epilogue geglu", got, ref, 1e-5)
    # bias along m + rowbias
    bm = torch.randn(M, device=dev)
    rb = torch.randn(M // 128, N, device=dev)
    got = ops.gemm(a, b, bias=bm, bias_along_m=True, rowbias=rb, rows_per_group=128, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T + bm[:, None] + rb.repeat_interleave(128, 0)
    allok &= report("epilogue bias_m+rowbias", got, ref, 1e-5)
    # second A source
    a2 = torch.randn(M, 128, device=dev).bfloat16()
    bcat = torch.randn(N, K + 128, device=dev).bfloat16()
    got = ops.gemm(a, bcat, a2=a2, a2_mode=1, out_dtype=torch.float32)
    torch.cuda.synchronize()
    allok &= report("a2 concat", got, torch.cat([a, a2], 1).float() @ bcat.float().T, 1e-5)
    x32 = torch.randn(M, K, device=dev)
    hi = x32.bfloat16()
    lo = (x32 - hi.float()).bfloat16()
    got = ops.gemm(hi, b, a2=lo, a2_mode=2, out_dtype=torch.float32)
    torch.cuda.synchronize()
    allok &= report("a2 split-precision", got, x32.double() @ b.double().T, 2e-5)
    # hi/lo output
    o_hi = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    o_lo = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, b, out=o_hi, out_lo=o_lo)
    torch.cuda.synchronize()
    allok &= report("hi+lo output", o_hi.float() + o_lo.float(), a.float() @ b.float().T, 2e-5)
    # conv3x3
    for (B, H, W, C, Co) in [(2, 64, 64, 64, 64), (2, 32, 32, 128, 320), (2, 16, 16, 64, 96), (4, 8, 8, 128, 64),
                             (3, 8, 8, 64, 64), (1, 128, 128, 64, 32), (1, 256, 256, 64, 16)]:
        x = torch.randn(B, H, W, C, device=dev).half()
        w = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half()
        bias = torch.randn(Co, device=dev)
        wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
        got = ops.conv3x3(x, wk, bias=bias, out_dtype=torch.float32)
        torch.cuda.synchronize()
        ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1).permute(0, 2, 3, 1)
        allok &= report(f"conv3x3 B{B} {H}x{W} C{C}->{Co}", got.reshape(-1, Co), ref.reshape(-1, Co), 1e-4)
    # timing
    for (M, N, K, bn) in [(8192, 8192, 8192, 256), (8192, 8192, 8192, 128), (65536, 320, 2880, 160), (640, 16384, 4096, 0)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, b, out=out, block_n=bn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, b, out=out, block_n=bn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"perf gemm {M}x{N}x{K} bn={bn}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        t0 = time.time()
        for _ in range(3):
            torch.matmul(a, b.T)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            torch.matmul(a, b.T)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"     cublas: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    print("ALL OK" if allok else "SOME BAD", flush=True)


if __name__ == "__main__":
    main()
