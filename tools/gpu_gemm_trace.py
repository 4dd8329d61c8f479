"""GILLB200_GEMM_DEBUG=3: where CTA 0's MMA warp and first epilogue warp wait (clock64 sums left in the stream-K scratch)."""
import os, sys
os.environ["GILLB200_GEMM_DEBUG"] = "3"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev = "cuda"
for (M, N, K, res) in [(65536, 2560, 320, False), (65536, 1152, 320, False), (65536, 320, 320, True), (65536, 2560, 640, False), (16384, 2304, 640, False),
                       (16384, 640, 640, True), (16384, 640, 768, True), (4096, 1280, 1280, True), (65536, 320, 384, True)]:
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half() * 0.05; bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float16); r = torch.randn(M, N, device=dev).half() if res else None
    kw = dict(block_n=256 if N >= 256 and N % 256 == 0 or N > 640 else 160, cta_pair=1, stream_k=0)
    if (M, N) in ((16384, 640), (4096, 1280)): kw = dict(block_n=160 if N == 640 else 256, cta_pair=1, stream_k=0)
    for _ in range(3): ops.gemm(a, b, out=out, bias=bias, residual=r, **kw)
    torch.cuda.synchronize()
    ws = list(ops._sk_ws.values())[0]
    d = ws[16384:16384 + 128].view(torch.int64).cpu().tolist()
    mt = max(d[3], 1); et = max(d[13], 1)
    print(f"M{M} N{N} K{K}{' +res' if res else ''} bn{kw['block_n']}: MMA warp per tile: wait tmem_empty {d[0]/mt:7.0f}  wait full {d[1]/mt:7.0f}  total {d[2]/mt:7.0f} clk ({d[3]} tiles) | "
          f"epilogue warp per tile: wait tmem_full {d[8]/et:7.0f}  wait store-drain {d[9]/et:7.0f}  tmem ld {d[10]/et:7.0f}  wait residual {d[11]/et:7.0f}  fence+syncwarp {d[14]/et:7.0f}  store issue {d[15]/et:7.0f}  total {d[12]/et:7.0f} clk ({d[13]} tiles)", flush=True)
