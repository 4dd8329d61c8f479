import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
dev = "cuda"
B, HW, C, Co = 16, int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(B, HW, HW, C, device=dev).half(); w = torch.randn(Co, 9 * C, device=dev).half() * 0.02; bias = torch.randn(Co, device=dev)
res = torch.randn(B, HW, HW, Co, device=dev).half(); out = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
for _ in range(4): ops.conv3x3(x, w, out=out, bias=bias, residual=res, stats=True)
torch.cuda.synchronize()
