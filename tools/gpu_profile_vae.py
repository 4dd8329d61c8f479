"""Per-shape time breakdown of one eager SD-1.5 VAE decode at batch 8 (CUDA events around every launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
from harness import synthetic
dev = "cuda"
pipe, *_ = synthetic.build_sd(dev, tiny=False)
lat = torch.randn(8, 64, 64, 4, device=dev)
for _ in range(2): pipe.vae.decode_u8(lat)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): pipe.vae.decode_u8(lat)
e1.record(); torch.cuda.synchronize()
print(f"eager decode: {e0.elapsed_time(e1) / 3:.2f} ms")
ops.PROFILE = []
for _ in range(3): pipe.vae.decode_u8(lat)
torch.cuda.synchronize()
rec = ops.PROFILE; ops.PROFILE = None
fam = ops.profile_summary(rec)
print(f"sum of launches {sum(d['ms'] for d in fam.values()) / 3:.2f} ms/decode")
for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:12s} {d['ms']/3:7.3f} ms  {d['launches']//3:4d} launches  " + (f"{d['flops']/d['ms']/1e9:7.1f} TF/s" if d["flops"] else f"{d['bytes']/d['ms']/1e6:7.0f} GB/s"))
shp = ops.profile_summary(rec, by_shape=True)
for k, d in sorted(shp.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    print(f"  {d['ms']/3:7.3f} ms  x{d['launches']//3:3d}  {d['ms']/d['launches']*1e3:8.1f} us  " + (f"{d['flops']/d['ms']/1e9:7.1f} TF/s" if d["flops"] else f"{d['bytes']/d['ms']/1e6:7.0f} GB/s") + f"  {k}")
