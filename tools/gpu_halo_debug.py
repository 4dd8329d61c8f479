"""Small halo-conv cases for compute-sanitizer / correctness triage. Usage: python tools/gpu_halo_debug.py [case ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gill_b200 import ops
dev = "cuda"
CASES = {
    "wide_c64": (2, 64, 64, 64, 320, 320, "res"),
    "wide_c128_n640": (3, 32, 32, 128, 640, 320, "rowbias"),
    "wide_c128_n320": (3, 32, 32, 128, 320, 320, "none"),
    "bn256_n320": (2, 32, 32, 128, 320, 256, "res"),
    "bn256_n256": (2, 32, 32, 128, 256, 256, "none"),
    "bn160_n320": (2, 32, 32, 128, 320, 160, "res"),
    "bn128_n128": (2, 32, 32, 128, 128, 128, "stats"),
}
for name in (sys.argv[1:] or list(CASES)):
    B, H, W, C, Co, bn, mode = CASES[name]
    torch.manual_seed(5)
    x = torch.randn(B, H, W, C, device=dev).half()
    w = (torch.randn(Co, C, 3, 3, device=dev) / (3 * C ** 0.5)).half()
    bias = torch.randn(Co, device=dev)
    wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    kw = {}
    if mode == "res":
        res = torch.randn(B, H, W, Co, device=dev).half(); kw["residual"] = res; ref = ref + res.float()
    elif mode == "rowbias":
        rb = torch.randn(B, Co, device=dev); kw["rowbias"] = rb; ref = ref + rb[:, None, None, :]
    try:
        got = ops.conv3x3(x, wk, bias=bias, block_n=bn, cta_pair=2, stats=(mode == "stats"), **kw)
        torch.cuda.synchronize()
        err = ((got.float() - ref).norm() / ref.norm()).item()
        print(f"{name}: rel err {err:.3e}", flush=True)
    except Exception as e:
        print(f"{name}: EXCEPTION {str(e)[:200]}", flush=True)
        break
