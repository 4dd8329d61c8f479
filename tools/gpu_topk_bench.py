"""Retrieval kernel timing vs shard size (Q=1024, D=768, K=16): fixed per-launch cost = intercept of t(n)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import retrieval
dev, tag = "cuda", (sys.argv[1] if len(sys.argv) > 1 else "")
q = torch.randn(1024, 768, device=dev); q = (q / q.norm(dim=1, keepdim=True)).bfloat16()
ws = torch.empty(1 << 27, device=dev, dtype=torch.uint8)
for n in (375_000, 750_000, 1_500_000, 3_000_000):
    bank = torch.randn(n, 768, device=dev); bank = (bank / bank.norm(dim=1, keepdim=True) * 14.24).bfloat16()
    f = lambda: retrieval.retrieval_topk(bank, q, 16, workspace=ws)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag} rows {n:>9d}: {ms:7.3f} ms  {2.0 * n * 768 * 1024 / ms / 1e9:7.1f} TF/s", flush=True)
    del bank
