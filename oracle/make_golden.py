"""ORACLE tooling: generate the fixtures under tests/golden/ by running the REFERENCE itself in the build container
(/root/reference imported unmodified through oracle/ref_shim.py; HF transformers for the OPT stand-in).

    python -m oracle.make_golden

Fixtures (all inputs are seeded; every file records how it was made):
  mapper_real.npz       reference `gill.layers.TextFcLayer` (fp32) with the SHIPPED checkpoint weights
  mapper_synth.npz      reference `TextFcLayer` loaded with oracle.mapper.synthetic_mapper_state_dict(1234)
  rethead_real.npz      reference retrieval head (`TextFcLayer` linear mode, shipped weights) + normalisation
  retrieval_tierA.npz   the reference expression `emb_matrix @ ret_emb.T; scores[seen] -= 1000; topk(3)` (fp32, CPU)
  opt_tiny.npz          transformers OPTForCausalLM (config-built, seeded) hidden_states[-1] / logits
  generate_tiny.npz     reference `GILLModel.generate` (patched to a config-built tiny OPT) ids / hidden / logits
  opt_wide.npz          transformers OPTForCausalLM at the BENCHMARKED width (hidden 4096, 32 heads, ffn 16384; 2 layers),
                        B=8, T=81: the 8 [IMG]-position hidden states + strided last-real-position logits
  opt_125m.npz          the same at the OPT-125M shape (hidden 768, 12 heads of 64, ffn 3072; 2 layers), BASELINE configs[0]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mapper as omap, opt as oopt, ref_shim, retrieval as oret  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in kw.items()})
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()})


def mapper_inputs(B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 8, 4096, generator=g).bfloat16().float()


@torch.no_grad()
def golden_mapper():
    layers, _, _ = ref_shim.import_reference()
    full = ref_shim.load_real_state_dict()
    img = full["model.input_embeddings.weight"].float()[None]            # the 8 [IMG] embedding rows
    x = mapper_inputs(2, 1234)
    m = ref_shim.reference_mapper()
    save("mapper_real.npz", x_seed=np.int64(1234), out=m(x, img), note="reference TextFcLayer fp32, shipped ckpt, "
         "x=randn(2,8,4096,seed 1234).bfloat16().float(), input_embs=[IMG] rows of the ckpt")
    ssd = omap.synthetic_mapper_state_dict(1234)
    ms = layers.TextFcLayer(4096, 768, num_input_tokens=8, num_output_tokens=77, mode="gill_mapper")
    ms.load_state_dict(ssd, strict=True)
    ms.eval()
    g = torch.Generator().manual_seed(99)
    img_s = (torch.randn(8, 4096, generator=g) * 0.024).bfloat16().float()[None]
    save("mapper_synth.npz", x_seed=np.int64(1234), out=ms(x, img_s), note="reference TextFcLayer fp32, "
         "oracle.mapper.synthetic_mapper_state_dict(1234), input_embs=(randn(8,4096,seed 99)*0.024).bfloat16()")
    # retrieval head (linear mode) + L2 normalisation + bf16 cast  (gill/models.py:673-675)
    rh = layers.TextFcLayer(4096, 256, num_input_tokens=8, num_output_tokens=1, mode="linear")
    rh.load_state_dict({"model.weight": full["model.ret_text_hidden_fcs.0.model.weight"].float(),
                        "model.bias": full["model.ret_text_hidden_fcs.0.model.bias"].float()})
    r = rh(x, None)[:, 0, :]
    r = r / r.norm(dim=-1, keepdim=True)
    save("rethead_real.npz", x_seed=np.int64(1234), ret_emb=r, note="reference ret head fp32 + normalise")


@torch.no_grad()
def golden_retrieval():
    bank = oret.synthetic_bank_chunk(0, 4096, 256, exact=True)
    q = oret.synthetic_queries(6, 256, exact=True)
    seen = [3, 100, 2047]
    vals, idxs = [], []
    for qi in range(q.shape[0]):
        scores = bank.float() @ q[qi : qi + 1].float().T                  # gill/models.py:676 (fp32 on CPU)
        for s in seen:
            scores[s, :] -= 1000                                          # :679-680
        v, i = scores.squeeze().topk(3)                                   # :683
        vals.append(v)
        idxs.append(i)
    v_ref, i_ref = torch.stack(vals), torch.stack(idxs)
    v_or, i_or = oret.retrieval_topk(bank, q, 3, exclude_idx=seen)
    assert torch.equal(v_ref, v_or), "oracle values differ from the reference expression"
    # torch.topk leaves tie order unspecified: indices must agree wherever the value is unique in its row
    for r in range(v_ref.shape[0]):
        for c in range(3):
            if (v_ref[r] == v_ref[r, c]).sum() == 1:
                assert i_ref[r, c] == i_or[r, c]
    save("retrieval_tierA.npz", values=v_ref, indices_reference=i_ref, indices_lowest_tie=i_or,
         seen=np.array(seen), note="bank=synthetic_bank_chunk(0,4096,256,exact), q=synthetic_queries(6,256,exact)")


@torch.no_grad()
def golden_opt():
    from transformers import OPTConfig, OPTForCausalLM

    cfg = oopt.opt_config("opt-tiny")
    sd = {k: v.bfloat16().float() for k, v in oopt.init_opt(cfg, seed=3).items()}
    hc = OPTConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                   num_attention_heads=cfg["heads"], ffn_dim=cfg["ffn"], max_position_embeddings=cfg["max_pos"],
                   word_embed_proj_dim=cfg["hidden"], do_layer_norm_before=True, activation_function="relu")
    m = OPTForCausalLM(hc).eval()
    m.load_state_dict({**sd, "lm_head.weight": sd["model.decoder.embed_tokens.weight"]}, strict=True)
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(3, 21, cfg["hidden"], generator=g) * 0.05).bfloat16().float()
    o = m(inputs_embeds=x, use_cache=False, output_hidden_states=True)
    save("opt_tiny.npz", hidden=o.hidden_states[-1], last_logits=o.logits[:, -1],
         note="transformers OPTForCausalLM(opt-tiny cfg), weights=init_opt(seed 3).bfloat16(), x=randn(3,21,256,seed 1)*0.05 bf16")
    return hc, sd, cfg


def wide_inputs(cfg, B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, T, cfg["hidden"], generator=g) * 0.05).bfloat16().float()


@torch.no_grad()
def golden_opt_shapes():
    """OPT at the benchmarked width and at the OPT-125M head size: what the hot path consumes -- hidden
    states of the 8 trailing [IMG] positions and the logits of the last prompt position (every 16th vocabulary entry
    plus the last 16, which hold the [IMG] ids), from transformers' own OPTForCausalLM."""
    from transformers import OPTConfig, OPTForCausalLM

    for name, base, seed, B, T in (("opt_wide.npz", "opt-6.7b", 5, 8, 81), ("opt_125m.npz", "opt-125m", 6, 2, 33)):
        cfg = dict(oopt.opt_config(base), layers=2)
        sd = {k: v.bfloat16().float() for k, v in oopt.init_opt(cfg, seed=seed).items()}
        hc = OPTConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                       num_attention_heads=cfg["heads"], ffn_dim=cfg["ffn"], max_position_embeddings=cfg["max_pos"],
                       word_embed_proj_dim=cfg["hidden"], do_layer_norm_before=True, activation_function="relu")
        m = OPTForCausalLM(hc).eval()
        m.load_state_dict({**sd, "lm_head.weight": sd["model.decoder.embed_tokens.weight"]}, strict=True)
        x = wide_inputs(cfg, B, T, seed + 100)
        o = m(inputs_embeds=x, use_cache=False, output_hidden_states=True)
        hs, lg = o.hidden_states[-1], o.logits[:, T - 9]
        o_hs, o_lg = oopt.opt_forward(sd, cfg, x)
        assert torch.allclose(o_hs, hs, atol=2e-4, rtol=1e-4), (name, (o_hs - hs).abs().max())
        assert torch.allclose(o_lg[:, T - 9], lg, atol=2e-4, rtol=1e-4), name
        sel = torch.cat([torch.arange(0, cfg["vocab"] - 16, 16), torch.arange(cfg["vocab"] - 16, cfg["vocab"])])
        save(name, hidden_img=hs[:, T - 8:].half(), logits_sel=lg[:, sel], sel=sel.numpy(),
             hidden_rms=np.float32(hs.pow(2).mean().sqrt().item()),
             note=f"transformers OPTForCausalLM({base} shape, 2 layers), weights=init_opt(seed {seed}).bfloat16(), "
                  f"x=randn({B},{T},{cfg['hidden']},seed {seed + 100})*0.05 bf16; hidden_img = hidden_states[-1][:, -8:] (fp16), "
                  f"logits_sel = logits[:, T-9, sel]")


@torch.no_grad()
def golden_generate(hc, sd, cfg):
    """The reference's own GILLModel.generate with OPT/CLIP patched to config-built models."""
    import transformers
    from transformers import CLIPVisionConfig, CLIPVisionModel, OPTForCausalLM

    _, models, utils = ref_shim.import_reference()
    V = cfg["vocab"]
    img_ids = list(range(V - 8, V))

    class Tok:
        cls_token_id, pad_token_id, bos_token_id = V - 9, 2, 2

        def __len__(self):
            return V

    def fake_opt(*a, **k):
        m = OPTForCausalLM(hc).eval()
        m.load_state_dict({**sd, "lm_head.weight": sd["model.decoder.embed_tokens.weight"]}, strict=True)
        return m

    vc = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                          image_size=32, patch_size=16)
    models.OPTForCausalLM.from_pretrained = staticmethod(fake_opt)
    models.CLIPVisionModel.from_pretrained = staticmethod(lambda *a, **k: CLIPVisionModel(vc))
    utils.get_feature_extractor_for_model = lambda *a, **k: None
    args = models.GILLArgs()
    args.opt_version, args.retrieval_token_idx, args.gen_token_idx = "facebook/opt-tiny", img_ids, img_ids
    args.ret_emb_dim, args.gen_emb_dim = 256, 768
    gm = models.GILLModel(Tok(), args)
    g = torch.Generator().manual_seed(11)
    emb = (torch.randn(1, 9, cfg["hidden"], generator=g) * 0.05).bfloat16().float()
    out = {}
    for name, kw in (("forced", dict(max_len=2, gen_scale_factor=1e5)), ("greedy", dict(max_len=4)),
                     ("minwords", dict(max_len=3, min_word_tokens=2, gen_scale_factor=1e5))):
        ids, embs, logits = gm.generate(emb, **kw)
        out[name + "_ids"] = ids
        out[name + "_hidden_last"] = embs[-1]
        out[name + "_logits0"] = logits[0]
        o_ids, o_embs, o_logits = oopt.generate(sd, cfg, emb, img_ids, img_ids, **kw)
        assert torch.equal(o_ids, ids), (name, o_ids, ids)
        assert torch.allclose(o_embs[-1], embs[-1], atol=1e-5), name
    save("generate_tiny.npz", img_ids=np.array(img_ids), **out,
         note="reference GILLModel.generate on the opt-tiny stand-in; emb=randn(1,9,256,seed 11)*0.05 bf16")


if __name__ == "__main__":
    if not ref_shim.available():
        sys.exit("/root/reference is not available: fixtures can only be regenerated in the build container")
    torch.manual_seed(0)
    golden_mapper()
    golden_retrieval()
    hc, sd, cfg = golden_opt()
    golden_generate(hc, sd, cfg)
    golden_opt_shapes()
