"""Full SD-1.5-shaped UNet, one CFG pair, vs the CPU fp32 oracle: prints the relative error (test tolerance 5e-3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import sd as psd
from oracle import sd15 as osd
dev = "cuda"
usd = {k: v.half().float() for k, v in osd.init_unet(0).items()}
unet = psd.UNetB200(usd, device=dev)
table = psd.plms_table(50)
unet.prepare_timesteps([t for t, _, _, _ in table])
g = torch.Generator().manual_seed(5)
lat = torch.randn(1, 4, 64, 64, generator=g).half().float()
ctx = torch.randn(2, 77, 768, generator=g).half().float()
kv = unet.precompute_ctx(ctx.to(dev))
pair = torch.cat([lat, lat], 0).permute(0, 2, 3, 1).contiguous().to(dev).half()
ref = osd.unet_forward(usd, torch.cat([lat, lat], 0), table[0][0], ctx)
for i in range(3):
    eps = unet.forward(pair, 0, kv)
    e = eps.permute(0, 3, 1, 2).float().cpu()
    print(f"run {i}: rel = {((e - ref).norm() / ref.norm()).item():.3e}  max|d| = {(e - ref).abs().max().item():.3e}  finite={torch.isfinite(e).all().item()}")
