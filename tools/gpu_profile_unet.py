"""Per-shape time breakdown of one eager SD-1.5 UNet evaluation at batch 16 (CUDA events around every launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops, sd as psd
from harness import synthetic
dev = "cuda"
pipe, *_ = synthetic.build_sd(dev, tiny=False)
table = psd.plms_table(50)
pipe.unet.prepare_timesteps([t for t, _, _, _ in table])
pair = torch.randn(16, 64, 64, 4, device=dev).half()
kv = pipe.unet.precompute_ctx(torch.randn(16, 77, 768, device=dev).half())
for _ in range(2): pipe.unet.forward(pair, 5, kv)
torch.cuda.synchronize()
ops.PROFILE = []
for _ in range(3): pipe.unet.forward(pair, 5, kv)
torch.cuda.synchronize()
rec = ops.PROFILE; ops.PROFILE = None
fam = ops.profile_summary(rec)
tot = sum(d["ms"] for d in fam.values()) / 3
print(f"total {tot:.2f} ms/eval")
for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:12s} {d['ms']/3:7.3f} ms  {d['launches']//3:4d} launches  " + (f"{d['flops']/d['ms']/1e9:7.1f} TF/s" if d["flops"] else f"{d['bytes']/d['ms']/1e6:7.0f} GB/s"))
print("top shapes:")
shp = ops.profile_summary(rec, by_shape=True)
for k, d in sorted(shp.items(), key=lambda kv: -kv[1]["ms"])[:70]:
    print(f"  {d['ms']/3:7.3f} ms  x{d['launches']//3:3d}  {d['ms']/d['launches']*1e3:8.1f} us  " + (f"{d['flops']/d['ms']/1e9:7.1f} TF/s" if d["flops"] else f"{d['bytes']/d['ms']/1e6:7.0f} GB/s") + f"  {k}")
