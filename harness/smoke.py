"""__graft_entry__.smoke(): one small invocation of the whole hot path on cuda:0, checked against the CPU oracle.

OPT (2 OPT-6.7B-shaped layers) prefill -> [IMG] hidden states -> GILLMapper -> retrieval top-k -> tiny SD-1.5-shaped
UNet (PLMS, 6 steps => 7 evaluations, CFG) -> tiny VAE -> uint8. Every stage is compared with oracle/ on identical
inputs; the oracle is only the checker here (the product path above never touches it)."""
import torch


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def run():
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device (sm_100a)")
    from gill_b200 import ops, retrieval
    from harness import synthetic
    from oracle import mapper as omap, retrieval as oret, sd15 as osd

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    gill, kind = synthetic.build_gill(dev, "opt-2l", tiny_sd=True, with_sd=True)
    m = gill.model
    B, P = 2, 24
    g = torch.Generator().manual_seed(5)
    vis = (torch.randn(B, 8, 4096, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    ids = torch.randint(3, 50265, (B, P - 8), generator=g).to(dev)
    embs = torch.cat([vis, m.input_embeddings(ids)], dim=1)
    lat = torch.randn(B, 4, 32, 32, generator=g).to(torch.float16)
    # retrieval bank (tier A: exactly representable values => bit-exact expectation)
    bank = oret.synthetic_bank_chunk(0, 20000, 256, exact=True)
    gill.emb_matrix = bank.to(dev)
    out = gill.emit_images_batch(embs, latents=lat.to(dev), num_inference_steps=6, top_k=0)
    torch.cuda.synchronize()
    imgs = out["images"]
    assert imgs.shape == (B, 64, 64, 3) and imgs.dtype == torch.uint8, imgs.shape  # tiny VAE: one 2x upsample

    # --- GILLMapper vs oracle on the same [IMG] hidden states
    img_ids = torch.tensor(m.retrieval_token_idx, device=dev)
    full = torch.cat([embs, m.input_embeddings(img_ids[None]).expand(B, -1, -1)], 1)
    hs, _ = m.lm.forward(full, logit_positions=[P - 1])
    raw = hs[:, P:P + 8].float()
    msd = {k: v.detach().float().cpu() for k, v in m.gen_text_hidden_fcs[0].state_dict().items()}
    ie = m.input_embeddings(img_ids[None]).float().cpu()
    ref_gen = omap.mapper_forward({k: v.double() for k, v in msd.items()}, raw.cpu().double(), ie.double())
    r_map = _rel(out["gen_emb"], ref_gen)
    assert r_map < 1e-3, f"GILLMapper parity {r_map:.3e}"

    # --- retrieval top-k vs oracle (bit-exact on exactly representable data)
    q = oret.synthetic_queries(5, 256, exact=True)
    v, i = retrieval.retrieval_topk(gill.emb_matrix, q.to(dev), 3, exclude_idx=[7, 11])
    rv, ri = oret.retrieval_topk(bank, q, 3, exclude_idx=[7, 11])
    assert torch.equal(i.cpu(), ri) and torch.equal(v.cpu(), rv), "retrieval top-k mismatch"

    # --- UNet loop + VAE vs oracle
    sdp = gill.sd_pipe
    ucfg, vcfg = osd.tiny_unet_cfg(), osd.tiny_vae_cfg()
    usd, vsd = osd.init_unet(0, ucfg), osd.init_vae_decoder(1, vcfg)
    g2 = torch.Generator().manual_seed(77)
    neg = torch.randn(1, 77, 768, generator=g2)
    cond = out["gen_emb"].float().cpu()
    ref_lat = osd.denoise_loop(usd, cond, neg, lat.float(), 7.5, 6, ucfg)
    got_lat = sdp.denoise(out["gen_emb"], lat.to(dev), 7.5, 6).permute(0, 3, 1, 2)
    r_lat = _rel(got_lat, ref_lat)
    assert r_lat < 2e-2, f"denoising-loop latent parity {r_lat:.3e}"
    ref_img = osd.to_uint8_nhwc(osd.vae_decode(vsd, ref_lat, vcfg))
    d = (imgs.cpu().int() - ref_img.int()).abs()
    assert d.float().mean().item() < 2.0, f"decoded image mean |diff| {d.float().mean().item():.3f} uint8 levels"
    print(f"smoke OK: weights={kind} mapper rel={r_map:.2e} latents rel={r_lat:.2e} "
          f"image mean|diff|={d.float().mean().item():.3f}/255 max={d.max().item()} retrieval bit-exact")
