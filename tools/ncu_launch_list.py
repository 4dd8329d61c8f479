"""One pass of every stage of the hot path (eager, no graph) for the ncu launch list: OPT-6.7B prefill (B=8, T=81),
GILLMapper (B=8), one UNet evaluation (B=16) + PLMS step, VAE decode (B=8), retrieval top-k (3M x 768, Q=1024)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops, sd as psd, retrieval
from harness import synthetic
dev = "cuda"
gill, kind = synthetic.build_gill(dev, "opt-6.7b", tiny_sd=False, with_sd=True)
m = gill.model
g = torch.Generator().manual_seed(0)
embs = (torch.randn(8, 81, 4096, generator=g) * 0.05).bfloat16().to(dev)
sdp = gill.sd_pipe
table = psd.plms_table(50)
sdp.unet.prepare_timesteps([t for t, _, _, _ in table])
pair = torch.randn(16, 64, 64, 4, device=dev).half()
kv = sdp.unet.precompute_ctx(torch.randn(16, 77, 768, device=dev).half())
lat = torch.randn(8, 64, 64, 4, device=dev)
ets = torch.zeros(4, lat.numel(), device=dev); cur = torch.zeros(lat.numel(), device=dev)
bank = torch.randn(3_000_000, 768, device=dev).bfloat16(); q = torch.randn(1024, 768, device=dev).bfloat16()
def one_pass():
    hs, lg = m.lm.forward(embs, logit_positions=[72])
    raw = hs[:, 73:81].float().contiguous()
    gen = m.gen_text_hidden_fcs[0](raw, torch.zeros(1, 8, 4096, device=dev))
    eps = sdp.unet.forward(pair, 5, kv)
    ops.plms_step(eps, 7.5, ets, 0, 4, 1.01, 0.01, lat, cur, pair)
    sdp.vae.decode_u8(lat)
    retrieval.retrieval_topk(bank, q, 16)
one_pass(); torch.cuda.synchronize()
torch.cuda.profiler.start()
one_pass(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
