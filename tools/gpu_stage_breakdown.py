"""Where one bench step (8 prompts -> 8 images) spends its time: OPT prefill / GILLMapper / 51 graph-replayed UNet
evaluations + PLMS / VAE decode. CUDA events on the current stream, warm caches, 3 repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gill_b200 import synthetic

dev = torch.device("cuda", 0)
gill, _ = synthetic.build_gill(dev, "opt-6.7b", tiny_sd=False, with_sd=True)
m = gill.model
vis, ids, lat = [t.to(dev) for t in bench.make_inputs(0)]


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def step():
    marks = [ev()]
    txt = m.input_embeddings(ids)
    embs = torch.cat([vis, txt], dim=1)
    B, P, D = embs.shape
    img_ids = torch.tensor(m.retrieval_token_idx, dtype=torch.int64, device=dev)
    img_embs = m.input_embeddings(img_ids[None, :])
    full = torch.cat([embs.to(m.lm.dt), img_embs.expand(B, -1, -1).to(m.lm.dt)], dim=1)
    hs, lg = m.lm.forward(full, logit_positions=[P - 1])
    raw = hs[:, P:P + m.num_tokens, :].float().contiguous()
    marks.append(ev())
    gen = m.gen_text_hidden_fcs[0](raw, img_embs.float())
    marks.append(ev())
    latn = gill.sd_pipe.denoise(gen, lat)
    marks.append(ev())
    u8 = gill.sd_pipe.vae.decode_u8(latn)
    marks.append(ev())
    return marks


for _ in range(2):
    step()
torch.cuda.synchronize()
acc = [0.0] * 4
R = 3
for _ in range(R):
    mk = step()
    torch.cuda.synchronize()
    for i in range(4):
        acc[i] += mk[i].elapsed_time(mk[i + 1]) / R
names = ["OPT-6.7B prefill (8x81 tokens)", "GILLMapper (B=8)", "denoise: 51 x (UNet B=16 + PLMS)", "VAE decode (B=8)"]
tot = sum(acc)
for n, t in zip(names, acc):
    print(f"{n:36s} {t:9.2f} ms  {100 * t / tot:5.1f} %")
print(f"{'total':36s} {tot:9.2f} ms   -> {8 / tot * 1e3:.2f} images/s;  UNet eval {acc[2] / 51:.2f} ms")
