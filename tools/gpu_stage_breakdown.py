"""Where one bench step (8 prompts -> 8 images) spends its time (bench.stage_breakdown): OPT prefill / GILLMapper /
51 graph-replayed UNet evaluations + PLMS / VAE decode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from harness import synthetic

dev = torch.device("cuda", 0)
gill, _ = synthetic.build_gill(dev, "opt-6.7b", tiny_sd=False, with_sd=True)
vis, ids, lat = [t.to(dev) for t in bench.make_inputs(0)]
st = bench.stage_breakdown(gill, vis, ids, lat, reps=3)
tot = st["opt_prefill"] + st["gill_mapper"] + st["unet_51_evals_plms"] + st["vae_decode"]
for k, v in st.items():
    print(f"{k:22s} {v:9.2f} ms")
print(f"total {tot:.2f} ms -> {8 / tot * 1e3:.2f} images/s")
