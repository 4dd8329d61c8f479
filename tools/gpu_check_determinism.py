"""Bitwise run-to-run repeatability of every op and of the UNet/VAE graphs (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gill_b200 import ops, sd as psd
from harness import synthetic
dev = "cuda"
torch.manual_seed(0)

def rep(name, fn, n=6):
    outs = [fn().clone() for _ in range(n)]
    torch.cuda.synchronize()
    bad = sum(int(not torch.equal(outs[0], o)) for o in outs[1:])
    mx = max((outs[0].float() - o.float()).abs().max().item() for o in outs[1:])
    print(f"[{'OK ' if bad == 0 else 'BAD'}] {name}: {bad}/{n-1} differ, max|d|={mx:.3e}", flush=True)

a = torch.randn(4096, 320, device=dev).half(); b = torch.randn(2560, 320, device=dev).half(); bias = torch.randn(2560, device=dev)
rep("gemm geglu", lambda: ops.gemm(a, b, bias=bias, act="geglu"))
rep("gemm plain bn160", lambda: ops.gemm(a, b[:320], block_n=160))
rep("gemm plain bn32", lambda: ops.gemm(a, b[:64], block_n=32))
x = torch.randn(4, 32, 32, 128, device=dev).half(); w = torch.randn(128, 9 * 128, device=dev).half() * 0.03
rep("conv3x3", lambda: ops.conv3x3(x, w))
x8 = torch.randn(4, 8, 8, 128, device=dev).half()
rep("conv3x3 8x8", lambda: ops.conv3x3(x8, w))
gw, gb = torch.randn(128, device=dev), torch.randn(128, device=dev)
rep("groupnorm", lambda: ops.groupnorm(x, gw, gb, 32, 1e-5, silu=True))
rep("groupnorm concat", lambda: ops.groupnorm(x, torch.cat([gw, gw]), torch.cat([gb, gb]), 32, 1e-5, silu=True, x2=x))
h = torch.randn(4096, 128, device=dev).half()
rep("layernorm", lambda: ops.layernorm(h, gw, gb, 1e-5))
qkv = torch.randn(4, 1024, 3 * 2 * 64, device=dev).half()
rep("attention self", lambda: ops.attention(qkv[:, :, :128], qkv[:, :, 128:256], qkv[:, :, 256:], 2, 64, 0.125))
kv = torch.randn(4, 77, 256, device=dev).half()
rep("attention cross", lambda: ops.attention(qkv[:, :, :128], kv[:, :, :128], kv[:, :, 128:], 2, 64, 0.125))
qkv2 = torch.randn(4, 256, 3 * 2 * 128, device=dev).half()
rep("attention hd128", lambda: ops.attention(qkv2[:, :, :256], qkv2[:, :, 256:512], qkv2[:, :, 512:], 2, 128, 0.1))
qkv3 = torch.randn(4, 64, 3 * 2 * 192, device=dev).half()
rep("attention hd192", lambda: ops.attention(qkv3[:, :, :384], qkv3[:, :, 384:768], qkv3[:, :, 768:], 2, 192, 0.1))
rep("im2col", lambda: ops.im2col3x3(x, 2))
rep("upsample", lambda: ops.upsample2x(x))

pipe, usd, vsd, neg = synthetic.build_sd(dev, tiny=True)
table = psd.plms_table(50)
pipe.unet.prepare_timesteps([t for t, _, _, _ in table])
g = torch.Generator().manual_seed(2)
lat = torch.randn(2, 4, 32, 32, generator=g).half()
ctx = torch.randn(4, 77, 768, generator=g).half().to(dev)
kvs = pipe.unet.precompute_ctx(ctx)
pair = torch.cat([lat, lat], 0).permute(0, 2, 3, 1).contiguous().to(dev)
rep("unet eager eval", lambda: pipe.unet.forward(pair, 3, kvs))
emb = torch.randn(2, 77, 768, generator=g).to(dev)
rep("denoise eager 4 steps", lambda: pipe.denoise(emb, lat.to(dev), 7.5, 4, use_graph=False), n=4)
rep("denoise graph 4 steps", lambda: pipe.denoise(emb, lat.to(dev), 7.5, 4, use_graph=True), n=4)
z = torch.randn(2, 16, 16, 4, device=dev)
rep("vae decode", lambda: pipe.vae.decode_u8(z), n=4)
# per-layer hunt inside the UNet: hook ops to record outputs of two eager runs
import gill_b200.ops as O
names = ["gemm", "conv3x3", "groupnorm", "layernorm", "attention", "im2col3x3", "upsample2x"]
def record():
    rec = []
    orig = {n: getattr(O, n) for n in names}
    def wrap(n):
        def f(*a, **k):
            o = orig[n](*a, **k); rec.append((n, o.clone())); return o
        return f
    for n in names: setattr(O, n, wrap(n))
    try:
        pipe.unet.forward(pair, 3, kvs); torch.cuda.synchronize()
    finally:
        for n in names: setattr(O, n, orig[n])
    return rec
r1, r2 = record(), record()
for i, ((n1, o1), (n2, o2)) in enumerate(zip(r1, r2)):
    if not torch.equal(o1, o2):
        print(f"first divergence at op #{i} ({n1}) shape {tuple(o1.shape)} max|d|={(o1.float()-o2.float()).abs().max().item():.3e}", flush=True)
        break
else:
    print("no divergence inside a hooked eager UNet eval", flush=True)
