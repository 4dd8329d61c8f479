"""CPU tests: the oracle restatements against the fixtures produced by the reference itself (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import mapper as omap, opt as oopt, retrieval as oret, sd15 as osd


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def mapper_inputs(B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 8, 4096, generator=g).bfloat16().float()


def test_mapper_oracle_matches_reference_class_synthetic_weights(golden):
    g = golden("mapper_synth.npz")
    sd = omap.synthetic_mapper_state_dict(1234)
    gi = torch.Generator().manual_seed(99)
    img = (torch.randn(8, 4096, generator=gi) * 0.024).bfloat16().float()[None]
    out = omap.mapper_forward(sd, mapper_inputs(2, 1234), img)
    assert rel(out, g["out"]) < 5e-6


def test_mapper_oracle_matches_reference_class_real_checkpoint(golden):
    from harness import synthetic

    if not synthetic.real_checkpoint_available():
        pytest.skip("shipped checkpoint not present")
    ck = torch.load(synthetic.CKPT_DIR + "/pretrained_ckpt.pth.tar", map_location="cpu")["state_dict"]
    pre = "module.model.gen_text_hidden_fcs.0."
    sd = {k[len(pre):]: v.float() for k, v in ck.items() if k.startswith(pre)}
    img = ck["module.model.input_embeddings.weight"].float()[None]
    out = omap.mapper_forward(sd, mapper_inputs(2, 1234), img)
    assert rel(out, golden("mapper_real.npz")["out"]) < 5e-6
    # retrieval head + normalisation (gill/models.py:673-675)
    rsd = {"model.weight": ck["module.model.ret_text_hidden_fcs.0.model.weight"].float(),
           "model.bias": ck["module.model.ret_text_hidden_fcs.0.model.bias"].float()}
    r = omap.linear_head_forward(rsd, mapper_inputs(2, 1234))[:, 0, :]
    r = r / r.norm(dim=-1, keepdim=True)
    assert rel(r, golden("rethead_real.npz")["ret_emb"]) < 1e-6


def test_retrieval_oracle_matches_reference_expression(golden):
    g = golden("retrieval_tierA.npz")
    bank = oret.synthetic_bank_chunk(0, 4096, 256, exact=True)
    q = oret.synthetic_queries(6, 256, exact=True)
    v, i = oret.retrieval_topk(bank, q, 3, exclude_idx=g["seen"].tolist())
    assert np.array_equal(v.numpy(), g["values"])                 # bit exact
    assert np.array_equal(i.numpy(), g["indices_lowest_tie"])
    ref_i = g["indices_reference"]
    for r in range(v.shape[0]):                                   # torch.topk tie order is unspecified
        for c in range(3):
            if (g["values"][r] == g["values"][r, c]).sum() == 1:
                assert ref_i[r, c] == i[r, c]


def test_retrieval_edge_cases():
    bank = oret.synthetic_bank_chunk(1, 50, 64, exact=True)
    q = oret.synthetic_queries(3, 64, exact=True)
    v, i = oret.retrieval_topk(bank, q, 50)                       # k == N: a full stable sort
    assert (v[:, :-1] >= v[:, 1:]).all()
    ties = v[:, :-1] == v[:, 1:]
    assert (i[:, :-1][ties] < i[:, 1:][ties]).all()               # ties -> lowest index first
    # excluding everything only shifts scores by 1000; order is preserved
    v2, i2 = oret.retrieval_topk(bank, q, 5, exclude_idx=list(range(50)))
    assert torch.equal(i2, i[:, :5]) and torch.allclose(v2, v[:, :5] - 1000)
    # sharding + merge == single shard
    parts = [oret.retrieval_topk(bank[s:e], q, 5, index_base=s) for s, e in ((0, 20), (20, 37), (37, 50))]
    mv, mi = oret.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), 5)
    assert torch.equal(mi, i[:, :5]) and torch.equal(mv, v[:, :5])


def test_opt_oracle_matches_transformers(golden):
    g = golden("opt_tiny.npz")
    cfg = oopt.opt_config("opt-tiny")
    sd = {k: v.bfloat16().float() for k, v in oopt.init_opt(cfg, seed=3).items()}
    gx = torch.Generator().manual_seed(1)
    x = (torch.randn(3, 21, cfg["hidden"], generator=gx) * 0.05).bfloat16().float()
    hs, lg = oopt.opt_forward(sd, cfg, x)
    assert rel(hs, g["hidden"]) < 2e-6 and rel(lg[:, -1], g["last_logits"]) < 2e-6


@pytest.mark.parametrize("name", ["opt_125m.npz", "opt_wide.npz"])
def test_opt_oracle_matches_transformers_at_the_benchmarked_shapes(golden, name):
    """the OPT oracle pinned by transformers' OPTForCausalLM at hidden 4096 / 32 heads / ffn 16384 (the
    benchmarked width, 2 layers, B=8, T=81) and at the OPT-125M head size (BASELINE configs[0])."""
    g = golden(name)
    cfg, sd, x, T = oopt.shape_case(name)
    hs, lg = oopt.opt_forward(sd, cfg, x)
    assert rel(hs[:, T - 8:], g["hidden_img"].astype(np.float32)) < 1e-3          # fixture stored as fp16
    assert rel(lg[:, T - 9][:, torch.as_tensor(g["sel"])], g["logits_sel"]) < 1e-4


def test_generate_oracle_matches_reference_generate(golden):
    g = golden("generate_tiny.npz")
    cfg = oopt.opt_config("opt-tiny")
    sd = {k: v.bfloat16().float() for k, v in oopt.init_opt(cfg, seed=3).items()}
    ge = torch.Generator().manual_seed(11)
    emb = (torch.randn(1, 9, cfg["hidden"], generator=ge) * 0.05).bfloat16().float()
    img = g["img_ids"].tolist()
    for name, kw in (("forced", dict(max_len=2, gen_scale_factor=1e5)), ("greedy", dict(max_len=4)),
                     ("minwords", dict(max_len=3, min_word_tokens=2, gen_scale_factor=1e5))):
        ids, embs, logits = oopt.generate(sd, cfg, emb, img, img, **kw)
        assert np.array_equal(ids.numpy(), g[name + "_ids"]), name
        assert rel(embs[-1], g[name + "_hidden_last"]) < 2e-6
    with pytest.raises(ValueError):
        oopt.generate(sd, cfg, emb, img, img, max_len=1, top_p=0.5)


def test_sd15_parameter_counts_match_published_models():
    assert osd.param_count(osd.init_unet(0)) == 859_520_964
    assert osd.param_count(osd.init_vae_decoder(1)) == 49_490_179 + 20


def test_pndm_schedule_properties():
    s = osd.PNDM()
    ts = s.set_timesteps(50)
    assert len(ts) == 51 and ts[:4] == [981, 961, 961, 941] and ts[-1] == 1
    tab = osd.plms_schedule(50)
    assert [m for _, _, _, m in tab[:6]] == [0, 1, 2, 3, 4, 4]
    # linearity of step_plms in (eps, sample): step(a e1 + b e2) == a step(e1) + b step(e2) along a trajectory
    g = torch.Generator().manual_seed(0)
    def run(scale_e, scale_x):
        sc = osd.PNDM(); sc.set_timesteps(50)
        gg = torch.Generator().manual_seed(0)
        x = torch.randn(4, generator=gg) * scale_x
        for t in sc.timesteps[:8]:
            x = sc.step(torch.randn(4, generator=gg) * scale_e, t, x)
        return x
    assert torch.allclose(run(2.0, 2.0), 2 * run(1.0, 1.0), rtol=1e-5, atol=1e-6)
    # the product's host-side table is the same arithmetic
    from gill_b200.sd import plms_table
    for a, b in zip(plms_table(50), tab):
        assert a[0] == b[0] and a[3] == b[3] and abs(a[1] - b[1]) < 1e-7 and abs(a[2] - b[2]) < 1e-7


def test_tiny_unet_and_vae_oracle_shapes():
    ucfg, vcfg = osd.tiny_unet_cfg(), osd.tiny_vae_cfg()
    u, v = osd.init_unet(0, ucfg), osd.init_vae_decoder(1, vcfg)
    g = torch.Generator().manual_seed(0)
    eps = osd.unet_forward(u, torch.randn(2, 4, 32, 32, generator=g), 981, torch.randn(2, 77, 768, generator=g), ucfg)
    assert eps.shape == (2, 4, 32, 32) and torch.isfinite(eps).all()
    img = osd.vae_decode(v, torch.randn(1, 4, 16, 16, generator=g), vcfg)
    assert img.shape == (1, 3, 32, 32) and img.min() >= 0 and img.max() <= 1
    assert osd.to_uint8_nhwc(img).dtype == torch.uint8


def test_clip_oracle_matches_transformers_clip_vision_model():
    """oracle/clip.py is pinned by the installed transformers CLIPVisionModel (stand-in for the un-vendored 4.30.2 the
    reference uses at gill/models.py:79,135): same weights, same pixels -> same last_hidden_state / pooler_output."""
    import torch
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from oracle import clip as oclip

    cfg = oclip.tiny_cfg()
    hf = CLIPVisionModel(CLIPVisionConfig(hidden_size=cfg["hidden"], intermediate_size=cfg["mlp"],
                                          num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
                                          image_size=cfg["image"], patch_size=cfg["patch"], hidden_act="quick_gelu")).eval()
    sd = oclip.init_clip(cfg, seed=4)
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(5)
    px = torch.randn(2, 3, cfg["image"], cfg["image"], generator=g)
    with torch.no_grad():
        ref = hf(pixel_values=px)
        hs, pooled = oclip.clip_vision_forward(sd, px, cfg)
    assert (hs - ref.last_hidden_state).abs().max() < 2e-5
    assert (pooled - ref.pooler_output).abs().max() < 2e-5


def test_clip_preprocess_oracle_is_bit_exact_with_pil():
    """The restated 8-bit bicubic resample equals PIL's `img.resize((224, 224))` (gill/models.py:735) bit for bit, for the
    512 -> 224 case of the pipeline, a non-square input and an up-scaling one."""
    import numpy as np
    from PIL import Image
    from oracle import clip as oclip

    rng = np.random.default_rng(0)
    for (h, w, s) in ((512, 512, 224), (300, 400, 224), (96, 128, 224), (64, 64, 28)):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        img[: h // 3] = (np.linspace(0, 255, w)[None, :, None]).astype(np.uint8)       # smooth band + noise
        ref = np.asarray(Image.fromarray(img).resize((s, s)).convert("RGB"))
        got = oclip.pil_bicubic_resize_u8(img, s)
        assert np.array_equal(got, ref), (h, w, s, int(np.abs(got.astype(int) - ref.astype(int)).max()))
    pv = oclip.clip_preprocess(img, 28)
    assert pv.shape == (3, 28, 28) and pv.dtype == torch.float32
