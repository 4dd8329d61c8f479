"""Per-kernel-family time of one GILLMapper forward at B=256 (BASELINE configs[1]), eager launches with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
from gill_b200.layers import TextFcLayer
from harness import synthetic
dev = "cuda"
m = TextFcLayer(4096, 768, num_input_tokens=8, num_output_tokens=77, mode="gill_mapper")
ck = torch.load(os.path.join(synthetic.CKPT_DIR, "pretrained_ckpt.pth.tar"), map_location="cpu")["state_dict"]
pre = "module.model.gen_text_hidden_fcs.0."
m.load_state_dict({k[len(pre):]: v for k, v in ck.items() if k.startswith(pre)}, strict=True)
m = m.to(dev)
img = ck["module.model.input_embeddings.weight"].float()[None].to(dev)
x = torch.randn(256, 8, 4096, generator=torch.Generator().manual_seed(1234)).bfloat16().float().to(dev)
for _ in range(2): y = m(x, img)
torch.cuda.synchronize()
ops.PROFILE = []
for _ in range(3): y = m(x, img)
torch.cuda.synchronize()
prof = ops.profile_summary(ops.PROFILE, by_shape=True)
ops.PROFILE = None
tot = sum(d["ms"] for d in prof.values()) / 3
print(f"total {tot:.3f} ms per forward (eager, event overhead included)")
for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:24]:
    print(f"  {d['ms'] / 3:7.3f} ms x{d['launches'] // 3:3d}  {d['flops'] / max(d['ms'], 1e-9) / 1e9:7.1f} TF/s  {k}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = m(x, img)
e1.record(); torch.cuda.synchronize()
print(f"back-to-back: {e0.elapsed_time(e1) / 10:.3f} ms per forward")
