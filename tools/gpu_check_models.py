"""GPU check of the model-level graphs against the CPU oracle (run under gpurun; not a pytest)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gill_b200 import ops
from oracle import mapper as omap, opt as oopt, sd15 as osd

dev = "cuda"
torch.manual_seed(0)

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

def rep(name, r, tol):
    ok = r < tol and r == r
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel={r:.3e} (tol {tol})", flush=True)
    return ok

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def check_small_ops():
    ok = True
    for C, dt in [(320, torch.float16), (512, torch.float32), (1280, torch.float16), (4096, torch.float32), (640, torch.bfloat16)]:
        x = torch.randn(300, C, device=dev).to(dt); w = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
        got = ops.layernorm(x, w, b, 1e-5, out_dtype=torch.float32)
        ok &= rep(f"layernorm C{C} {dt}", rel(got, F.layer_norm(x.float(), (C,), w, b, 1e-5)), 1e-5)
    x = torch.randn(77, 512, device=dev); w = torch.randn(512, device=dev); b = torch.randn(512, device=dev)
    hi = torch.empty(77, 512, device=dev, dtype=torch.bfloat16); lo = torch.empty_like(hi)
    ops.layernorm(x, w, b, 1e-5, out=hi, out_lo=lo)
    ok &= rep("layernorm hi+lo", rel(hi.float() + lo.float(), F.layer_norm(x, (512,), w, b, 1e-5)), 2e-5)
    for (B, H, W, C0, C1, silu) in [(2, 64, 64, 320, 0, True), (2, 32, 32, 640, 320, True), (2, 8, 8, 1280, 1280, False),
                                    (1, 128, 128, 256, 0, True), (2, 16, 16, 1280, 640, True)]:
        x0 = torch.randn(B, H, W, C0, device=dev).half() * 2 + 0.5
        x1 = torch.randn(B, H, W, C1, device=dev).half() if C1 else None
        C = C0 + C1
        w = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
        got = ops.groupnorm(x0, w, b, 32, 1e-5, silu=silu, x2=x1, out_dtype=torch.float32)
        xx = torch.cat([x0, x1], -1) if C1 else x0
        ref = F.group_norm(xx.float().permute(0, 3, 1, 2), 32, w, b, 1e-5)
        if silu: ref = F.silu(ref)
        ok &= rep(f"groupnorm B{B} {H}x{W} C{C0}+{C1} silu={silu}", rel(got, ref.permute(0, 2, 3, 1)), 2e-5)
    x = torch.randn(2, 8, 8, 64, device=dev).half()
    ok &= rep("upsample2x", rel(ops.upsample2x(x), F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)), 1e-9)
    x = torch.randn(2, 16, 16, 64, device=dev).half()
    w = torch.randn(96, 64, 3, 3, device=dev).half() * 0.05
    cols = ops.im2col3x3(x, 2)
    got = ops.gemm(cols, w.permute(0, 2, 3, 1).reshape(96, -1).contiguous(), out_dtype=torch.float32).view(2, 8, 8, 96)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    ok &= rep("im2col s2 conv", rel(got, ref), 1e-5)
    s = torch.randn(500, 4096, device=dev)
    ok &= rep("softmax_rows", rel(ops.softmax_rows(s, 0.3, torch.float32), torch.softmax(s * 0.3, -1)), 1e-5)
    tab = torch.randn(100, 256, device=dev).bfloat16(); idx = torch.randint(0, 90, (37,), device=dev)
    xx = torch.randn(37, 256, device=dev).bfloat16()
    ok &= rep("gather_add_rows", rel(ops.gather_add_rows(tab, idx, x=xx, idx_offset=2), (xx.float() + tab[idx + 2].float()).bfloat16()), 1e-9)
    ok &= rep("gather rows", rel(ops.gather_add_rows(tab, idx), tab[idx]), 1e-9)
    x = torch.randn(9, 256, device=dev)
    ok &= rep("l2norm_rows", rel(ops.l2norm_rows(x, torch.float32), x / x.norm(dim=-1, keepdim=True)), 1e-6)
    img = torch.randn(2, 16, 16, 8, device=dev).half()
    ref = ((img[..., :3].float() / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
    got = ops.image_to_u8(img, 3)
    ok &= rep("image_to_u8", (got.int() - ref.int()).abs().max().item() / 255.0, 1e-9)
    # attn_small_f32
    B, Lq, Lk = 3, 77, 8
    q = torch.randn(B, Lq, 512, device=dev); kv = torch.randn(B, Lk, 1024, device=dev)
    out = torch.empty(B, Lq, 512, device=dev)
    ops.attn_small_f32(q, kv[:, :, :512], kv[:, :, 512:], 4, 128 ** -0.5, out=out)
    qh = q.view(B, Lq, 4, 128).transpose(1, 2); kh = kv[:, :, :512].reshape(B, Lk, 4, 128).transpose(1, 2); vh = kv[:, :, 512:].reshape(B, Lk, 4, 128).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 128 ** -0.5, -1) @ vh).transpose(1, 2).reshape(B, Lq, 512)
    ok &= rep("attn_small_f32 77x8", rel(out, ref), 1e-5)
    qkv = torch.randn(B, 77, 1536, device=dev)
    ops.attn_small_f32(qkv[:, :, :512], qkv[:, :, 512:1024], qkv[:, :, 1024:], 4, 128 ** -0.5, out=out)
    qh, kh, vh = (qkv[:, :, i * 512:(i + 1) * 512].reshape(B, 77, 4, 128).transpose(1, 2) for i in range(3))
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 128 ** -0.5, -1) @ vh).transpose(1, 2).reshape(B, 77, 512)
    ok &= rep("attn_small_f32 77x77", rel(out, ref), 1e-5)
    # plms_step vs oracle scheduler
    from gill_b200 import sd as psd
    table = psd.plms_table(50)
    sched = osd.PNDM(); sched.set_timesteps(50)
    n = 4 * 8 * 8 * 4
    lat = torch.randn(n, device=dev); lat_ref = lat.clone().cpu()
    ets = torch.zeros(4, n, device=dev); cur = torch.zeros(n, device=dev); head = 0
    pair = torch.zeros(2, n, device=dev, dtype=torch.float16)
    worst = 0.0
    for i, (t, cs, ce, mode) in enumerate(table[:12]):
        eps = torch.randn(2, n, device=dev)
        ops.plms_step(eps, 7.5, ets, head, mode, cs, ce, lat, cur, pair)
        if mode != 1: head = (head + 1) & 3
        e = eps.cpu(); e = e[0] + 7.5 * (e[1] - e[0])
        lat_ref = sched.step(e, t, lat_ref)
        worst = max(worst, rel(lat, lat_ref))
    ok &= rep("plms_step x12 vs oracle PNDM", worst, 1e-5)
    ok &= rep("plms lat16 pair", rel(pair[1], lat), 1e-3)
    return ok

def check_mapper():
    from gill_b200.layers import TextFcLayer
    ok = True
    sd = omap.synthetic_mapper_state_dict(1234)
    m = TextFcLayer(4096, 768, num_input_tokens=8, num_output_tokens=77, mode="gill_mapper")
    m.load_state_dict(sd, strict=True); m = m.to(dev)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 8, 4096, generator=g).bfloat16().float()
    ie = (torch.randn(1, 8, 4096, generator=g) * 0.024).bfloat16().float()
    ref = omap.mapper_forward({k: v.double() for k, v in sd.items()}, x.double(), ie.double())
    got = m(x.to(dev), ie.to(dev))
    torch.cuda.synchronize()
    ok &= rep("mapper B4 (synthetic weights) vs fp64 oracle", rel(got, ref), 1e-3)
    got16 = m(x.to(dev).bfloat16(), ie.to(dev).bfloat16())
    ok &= rep("mapper bf16-in/bf16-out", rel(got16, ref), 5e-3)
    # linear (retrieval head) mode
    g2 = torch.Generator().manual_seed(6)
    lin = TextFcLayer(4096, 256, num_input_tokens=8, num_output_tokens=1, mode="linear")
    lsd = {"model.weight": (torch.randn(256, 4096, generator=g2) / 64).bfloat16().float(), "model.bias": (torch.randn(256, generator=g2) * 0.02).bfloat16().float()}
    lin.load_state_dict(lsd); lin = lin.to(dev)
    ok &= rep("ret head linear", rel(lin(x.to(dev), None), omap.linear_head_forward({k: v.double() for k, v in lsd.items()}, x.double())), 1e-4)
    xb = torch.randn(256, 8, 4096, device=dev).bfloat16()
    ie_d = ie.to(dev).bfloat16()
    ms = timeit(lambda: m(xb, ie_d), n=5, warm=2)
    print(f"perf mapper B=256: {ms:.3f} ms ({676.8 / ms:.1f} TFLOP/s algorithmic)", flush=True)
    return ok

def check_opt():
    from gill_b200.opt import OPTB200
    ok = True
    cfg = oopt.opt_config("opt-tiny")
    sd = oopt.init_opt(cfg, seed=3)
    sd = {k: v.bfloat16().float() for k, v in sd.items()}
    m = OPTB200(sd, cfg["hidden"], cfg["layers"], cfg["heads"], cfg["ffn"], device=dev)
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(3, 21, cfg["hidden"], generator=g) * 0.05).bfloat16().float()
    hs_ref, lg_ref = oopt.opt_forward(sd, cfg, x)
    hs, lg = m.forward(x.to(dev))
    torch.cuda.synchronize()
    ok &= rep("opt-tiny hidden_states[-1]", rel(hs, hs_ref), 1.5e-2)
    ok &= rep("opt-tiny last logits", rel(lg, lg_ref[:, -1]), 1.5e-2)
    ok &= rep("embed_tokens", rel(m.embed_tokens(torch.tensor([[1, 5, 7]])), sd["model.decoder.embed_tokens.weight"][[1, 5, 7]][None]), 1e-9)
    return ok

def check_sd(full: bool):
    from gill_b200 import sd as psd
    ok = True
    ucfg, vcfg = osd.tiny_unet_cfg(), osd.tiny_vae_cfg()
    usd = {k: v.half().float() for k, v in osd.init_unet(0, ucfg).items()}
    vsd = {k: v.half().float() for k, v in osd.init_vae_decoder(1, vcfg).items()}
    unet = psd.UNetB200(usd, ucfg, device=dev)
    vae = psd.VAEDecoderB200(vsd, vcfg, device=dev)
    g = torch.Generator().manual_seed(2)
    b = 2
    lat = torch.randn(b, 4, 32, 32, generator=g).half().float()
    ctx = torch.randn(b, 77, 768, generator=g).half().float()
    neg = torch.randn(1, 77, 768, generator=g).half().float()
    # one UNet eval
    table = psd.plms_table(50)
    unet.prepare_timesteps([t for t, _, _, _ in table])
    cc = torch.cat([neg.expand(b, -1, -1), ctx], 0)
    kv = unet.precompute_ctx(cc.to(dev))
    pair = torch.cat([lat, lat], 0).permute(0, 2, 3, 1).contiguous().to(dev).half()
    eps = unet.forward(pair, 3, kv)
    torch.cuda.synchronize()
    ref = osd.unet_forward(usd, torch.cat([lat, lat], 0), table[3][0], cc, ucfg)
    ok &= rep("tiny UNet single eval (t index 3)", rel(eps.permute(0, 3, 1, 2), ref), 1e-2)
    # short denoise loop
    pipe = psd.StableDiffusionB200(unet, vae, neg)
    tr = []
    out = pipe.denoise(ctx.to(dev), lat.to(dev), 7.5, 10, trace=tr)
    torch.cuda.synchronize()
    ref_lat, ref_tr = osd.denoise_loop(usd, ctx, neg, lat, 7.5, 10, ucfg, return_all=True)
    ok &= rep("tiny denoise 10 steps (11 evals) final latents", rel(out.permute(0, 3, 1, 2), ref_lat), 1e-2)
    # VAE
    z = torch.randn(b, 4, 16, 16, generator=g)
    u8 = vae.decode_u8(z.permute(0, 2, 3, 1).contiguous().to(dev))
    torch.cuda.synchronize()
    ref_img = osd.to_uint8_nhwc(osd.vae_decode(vsd, z, vcfg))
    d = (u8.cpu().int() - ref_img.int()).abs()
    print(f"tiny VAE uint8: max|diff|={d.max().item()} mean|diff|={d.float().mean().item():.4f}", flush=True)
    ok &= d.max().item() <= 3
    if full:
        t0 = time.time()
        usd = osd.init_unet(0)
        print(f"full UNet init {time.time() - t0:.1f}s params={osd.param_count(usd)}", flush=True)
        usd = {k: v.half().float() for k, v in usd.items()}
        unet = psd.UNetB200(usd, device=dev)
        unet.prepare_timesteps([t for t, _, _, _ in table])
        b = 1
        lat = torch.randn(b, 4, 64, 64, generator=g).half().float()
        ctx = torch.randn(2 * b, 77, 768, generator=g).half().float()
        kv = unet.precompute_ctx(ctx.to(dev))
        pair = torch.cat([lat, lat], 0).permute(0, 2, 3, 1).contiguous().to(dev).half()
        eps = unet.forward(pair, 0, kv); torch.cuda.synchronize()
        t0 = time.time()
        ref = osd.unet_forward(usd, torch.cat([lat, lat], 0), table[0][0], ctx)
        print(f"oracle full UNet eval (batch 2, CPU fp32): {time.time() - t0:.1f}s", flush=True)
        ok &= rep("FULL SD-1.5 UNet single eval", rel(eps.permute(0, 3, 1, 2), ref), 1e-2)
        # perf at batch 16 (8 images x CFG)
        pair16 = torch.randn(16, 64, 64, 4, device=dev).half()
        kv16 = unet.precompute_ctx(torch.randn(16, 77, 768, device=dev).half())
        ms = timeit(lambda: unet.forward(pair16, 5, kv16), n=3, warm=2)
        print(f"perf FULL UNet eval batch 16 (eager launches): {ms:.2f} ms  ({16 * 0.8033 / ms:.1f} PFLOP/s... {16 * 803.3 / ms:.0f} TFLOP/s algorithmic)", flush=True)
        vsd = {k: v.half().float() for k, v in osd.init_vae_decoder(1).items()}
        vae = psd.VAEDecoderB200(vsd, device=dev)
        z8 = torch.randn(8, 64, 64, 4, device=dev)
        ms = timeit(lambda: vae.decode_u8(z8), n=2, warm=1)
        print(f"perf FULL VAE decode batch 8: {ms:.2f} ms ({8 * 2510 / ms:.0f} TFLOP/s algorithmic)", flush=True)
        z1 = torch.randn(1, 4, 64, 64, generator=g)
        u8 = vae.decode_u8(z1.permute(0, 2, 3, 1).contiguous().to(dev)); torch.cuda.synchronize()
        t0 = time.time()
        ref_img = osd.to_uint8_nhwc(osd.vae_decode(vsd, z1))
        print(f"oracle full VAE decode (CPU fp32): {time.time() - t0:.1f}s", flush=True)
        d = (u8.cpu().int() - ref_img.int()).abs()
        print(f"FULL VAE uint8: max|diff|={d.max().item()} mean|diff|={d.float().mean().item():.4f}", flush=True)
        ok &= d.max().item() <= 3
    return ok

if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "mapper", "opt", "sd"]
    ok = True
    for w in which:
        try:
            if w == "ops": ok &= check_small_ops()
            if w == "mapper": ok &= check_mapper()
            if w == "opt": ok &= check_opt()
            if w == "sd": ok &= check_sd(False)
            if w == "sdfull": ok &= check_sd(True)
        except Exception as e:
            import traceback; traceback.print_exc(); ok = False
            print(f"[BAD] {w} raised {e}", flush=True)
    print("ALL OK" if ok else "SOME BAD", flush=True)
