"""GPU check: fused attention and retrieval top-k (run under gpurun; not a pytest)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
from oracle import retrieval as orc

dev = "cuda"
torch.manual_seed(0)

def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

def attn_ref(q, k, v, H, hd, hd_pad, scale, causal=False, off=0, kv_lens=None):
    B, Lq, _ = q.shape; Lk = k.shape[1]
    qh = q.float().view(B, Lq, H, hd_pad).transpose(1, 2)
    kh = k.float().view(B, Lk, H, hd_pad).transpose(1, 2)
    vh = v.float().view(B, Lk, H, hd_pad).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if causal:
        i = torch.arange(Lq, device=dev)[:, None]; j = torch.arange(Lk, device=dev)[None]
        s = s.masked_fill(j > i + off, float("-inf"))
    if kv_lens is not None:
        j = torch.arange(Lk, device=dev)[None, None, None]
        s = s.masked_fill(j >= kv_lens.view(B, 1, 1, 1), float("-inf"))
    o = torch.softmax(s, -1) @ vh
    return o.transpose(1, 2).reshape(B, Lq, H * hd_pad)

def check_attn():
    ok = True
    cases = [  # B, H, Lq, Lk, hd, hd_pad, dtype, causal
        (2, 2, 128, 128, 40, 64, torch.float16, False),
        (2, 8, 1024, 1024, 40, 64, torch.float16, False),
        (1, 8, 4096, 4096, 40, 64, torch.float16, False),
        (2, 8, 1024, 77, 40, 64, torch.float16, False),
        (2, 8, 256, 256, 80, 128, torch.float16, False),
        (2, 8, 64, 64, 160, 192, torch.float16, False),
        (2, 8, 256, 77, 160, 192, torch.float16, False),
        (3, 32, 81, 81, 128, 128, torch.bfloat16, True),
        (2, 4, 300, 300, 128, 128, torch.bfloat16, True),
    ]
    for (B, H, Lq, Lk, hd, hp, dt, causal) in cases:
        def mk(L):
            t = torch.zeros(B, L, H, hp, device=dev)
            t[..., :hd] = torch.randn(B, L, H, hd, device=dev)
            return t.view(B, L, H * hp).to(dt)
        q, k, v = mk(Lq), mk(Lk), mk(Lk)
        scale = hd ** -0.5
        try:
            got = ops.attention(q, k, v, H, hp, scale, causal=causal)
            torch.cuda.synchronize()
        except Exception as e:
            print("attention EXC", (B, H, Lq, Lk, hd, hp), e, flush=True); ok = False; continue
        ref = attn_ref(q, k, v, H, hd, hp, scale, causal)
        r = rel(got, ref)
        good = r < (1e-2 if dt == torch.bfloat16 else 3e-3) and torch.isfinite(got.float()).all().item()
        ok &= good
        print(f"[{'OK ' if good else 'BAD'}] attn B{B} H{H} Lq{Lq} Lk{Lk} hd{hd}/{hp} {dt} causal={causal}: rel={r:.3e}", flush=True)
        if not good:
            d = (got.float() - ref).abs()
            print("   bad rows(b0):", (d[0].max(dim=1).values > 0.05).nonzero().flatten()[:12].tolist(),
                  " bad cols(b0):", (d[0].max(dim=0).values > 0.05).nonzero().flatten()[:12].tolist(), flush=True)
            print("   got", got[0, 0, :6].float().tolist(), "ref", ref[0, 0, :6].tolist(), flush=True)
    # kv_lens
    B, H, L, hd, hp = 3, 4, 200, 128, 128
    q = torch.randn(B, L, H * hp, device=dev).bfloat16(); k = torch.randn(B, L, H * hp, device=dev).bfloat16(); v = torch.randn(B, L, H * hp, device=dev).bfloat16()
    kvl = torch.tensor([200, 77, 130], device=dev, dtype=torch.int32)
    got = ops.attention(q, k, v, H, hp, hd ** -0.5, kv_lens=kvl); torch.cuda.synchronize()
    r = rel(got, attn_ref(q, k, v, H, hd, hp, hd ** -0.5, kv_lens=kvl))
    print(f"[{'OK ' if r < 1e-2 else 'BAD'}] attn kv_lens: rel={r:.3e}", flush=True); ok &= r < 1e-2
    # ones-column (packed exp2) mode
    for (B, H, Lq, Lk, hd, hp) in [(2, 8, 1024, 1024, 40, 64), (2, 8, 1024, 77, 40, 64), (2, 8, 256, 256, 80, 128), (2, 8, 64, 64, 160, 192)]:
        def mk(L, ones=False):
            t = torch.zeros(B, L, H, hp, device=dev); t[..., :hd] = torch.randn(B, L, H, hd, device=dev)
            if ones: t[..., hd] = 1.0
            return t.view(B, L, H * hp).half()
        q, k, v = mk(Lq), mk(Lk), mk(Lk, True)
        got = ops.attention(q, k, v, H, hp, hd ** -0.5, ones_col=hd); torch.cuda.synchronize()
        r = rel(got, attn_ref(q, k, v, H, hd, hp, hd ** -0.5))
        good = r < 3e-3; ok &= good
        print(f"[{'OK ' if good else 'BAD'}] attn ones_col Lq{Lq} Lk{Lk} hd{hd}/{hp}: rel={r:.3e}", flush=True)
    # perf
    B, H, L, hd, hp = 16, 8, 4096, 40, 64
    q = torch.randn(B, L, H * hp, device=dev).half(); k = torch.randn_like(q); v = torch.randn_like(q)
    out = torch.empty_like(q)
    vv = v.view(B, L, H, hp); vv[..., hd] = 1.0
    for _ in range(3): ops.attention(q, k, v, H, hp, hd ** -0.5, out=out, ones_col=hd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.attention(q, k, v, H, hp, hd ** -0.5, out=out, ones_col=hd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"perf attn (ones_col/packed) B16 H8 L4096 hd40: {ms:.3f} ms, {4 * B * H * L * L * hd / ms / 1e9:.1f} TFLOP/s (algorithmic)", flush=True)
    for _ in range(3): ops.attention(q, k, v, H, hp, hd ** -0.5, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.attention(q, k, v, H, hp, hd ** -0.5, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"perf attn B16 H8 L4096 hd40: {ms:.3f} ms, {4 * B * H * L * L * hd / ms / 1e9:.1f} TFLOP/s (algorithmic)", flush=True)
    return ok

def check_topk():
    ok = True
    for (N, D, Q, K, exact) in [(5000, 256, 1, 3, True), (70000, 256, 130, 16, True), (300000, 768, 1024, 16, True),
                                (300000, 768, 1024, 16, False), (100000, 256, 7, 3, False)]:
        bank = orc.synthetic_bank_chunk(0, N, D, exact).to(dev)
        q = orc.synthetic_queries(Q, D, exact).to(dev)
        excl = torch.tensor([5, 17], dtype=torch.int64, device=dev) if Q <= 7 else None
        v, i = ops.topk_scores(bank, q, K, index_base=1000, exclude_idx=(excl + 1000) if excl is not None else None)
        torch.cuda.synchronize()
        s = q.float() @ bank.float().T
        if excl is not None: s[:, excl] -= 1000
        rv, ri = orc.topk_lowest_index(s.cpu(), K, 1000)
        if exact:
            good = torch.equal(i.cpu(), ri) and torch.equal(v.cpu(), rv)
        else:
            mism = (i.cpu() != ri)
            gap_ok = ((v.cpu() - rv).abs() <= 1e-4 * rv.abs() + 1e-5).all().item()
            good = gap_ok and mism.float().mean().item() < 0.01
        ok &= good
        print(f"[{'OK ' if good else 'BAD'}] topk N{N} D{D} Q{Q} K{K} exact={exact}: idx_mismatch={(i.cpu() != ri).float().mean().item():.4f}", flush=True)
        if not good:
            print("   got", i[0].tolist(), v[0].tolist()); print("   ref", ri[0].tolist(), rv[0].tolist(), flush=True)
    # merge
    R, Q, K = 8, 100, 16
    cv = torch.randint(-50, 50, (R, Q, K), device=dev).float().sort(dim=2, descending=True).values
    ci = torch.randint(0, 1 << 40, (R, Q, K), device=dev)
    mv, mi = ops.topk_merge(cv, ci, K); torch.cuda.synchronize()
    rv, ri = orc.merge_topk(cv.cpu(), ci.cpu(), K)
    good = torch.equal(mv.cpu(), rv) and torch.equal(mi.cpu(), ri); ok &= good
    print(f"[{'OK ' if good else 'BAD'}] topk_merge", flush=True)
    # perf: 3M x 768 bf16 = 4.6 GB
    for (N, D, Q) in [(3_000_000, 768, 1024), (3_000_000, 256, 1024), (3_000_000, 768, 1), (3_000_000, 256, 1)]:
        bank = torch.randn(N, D, device=dev).bfloat16(); q = torch.randn(Q, D, device=dev).bfloat16()
        ws = torch.empty(64 << 20, device=dev, dtype=torch.uint8)
        for _ in range(2): ops.topk_scores(bank, q, 16, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.topk_scores(bank, q, 16, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"perf topk N{N} D{D} Q{Q}: {ms:.3f} ms  {2.0 * N * D * Q / ms / 1e9:.1f} TFLOP/s  {N * D * 2 / ms / 1e6:.0f} GB/s  {Q / ms * 1e3:.0f} QPS", flush=True)
        del bank
    return ok

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ok = True
    if which in ("all", "topk"): ok &= check_topk()
    if which in ("all", "attn"): ok &= check_attn()
    print("ALL OK" if ok else "SOME BAD", flush=True)
