"""CLIP ViT-L/14 vision tower throughput: images/s and TFLOP/s at batch 32 (random weights), plus the
device-side re-rank pre-processing (512x512 uint8 -> 224x224 pixel_values)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops
from gill_b200.clip import CLIPVisionB200
from oracle import clip as oclip
dev = "cuda"
cfg = oclip.CLIP_L14
tower = CLIPVisionB200(oclip.init_clip(cfg, seed=0), device=dev)
B = 32
u8 = torch.randint(0, 256, (B, 512, 512, 3), device=dev, dtype=torch.uint8)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
px = ops.clip_preprocess_u8(u8, 224)
t_pre = timeit(lambda: ops.clip_preprocess_u8(u8, 224))
t_tow = timeit(lambda: tower.forward(px))
T, D, F, L = 257, 1024, 4096, 24
flop = B * L * (2 * T * D * 3 * D + 4 * T * T * D + 2 * T * D * D + 4 * T * D * F) + B * 2 * T * 588 * D
print(f"clip_preprocess_u8 B{B} 512->224: {t_pre*1e3:.1f} us")
print(f"CLIP ViT-L/14 tower B{B}: {t_tow:.2f} ms  {B/t_tow*1e3:.0f} images/s  {flop/t_tow/1e9:.0f} TFLOP/s ({flop/B/1e9:.1f} GFLOP/image)")
