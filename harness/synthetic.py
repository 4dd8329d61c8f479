"""Seeded synthetic stand-ins for the artefacts that cannot be downloaded offline (HF_HUB_OFFLINE, no network):
the OPT tokenizer, OPT-6.7B / SD-1.5 weights, the CLIP-text embedding of "" and the cc3m retrieval bank. The GILL-
trained weights (mapper, retrieval head, [IMG] embeddings, decision MLP) are the REAL shipped checkpoint whenever
`checkpoints/gill_opt/` is present. Used by bench.py, __graft_entry__.smoke() and the tests; never by load_gill()."""
import math
import os
from collections import namedtuple
from types import SimpleNamespace
from typing import Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT_DIR = os.path.join(ROOT, "checkpoints", "gill_opt")
IMG_IDS = list(range(50266, 50274))  # checkpoints/gill_opt/model_args.json:18-37


class SyntheticTokenizer:
    """Minimal stand-in for `AutoTokenizer.from_pretrained('facebook/opt-6.7b', use_fast=False)` after the additions
    of gill/models.py:845-860: 50265 base tokens + <|image|> (50265) + [IMG0..7] (50266..50273); BOS </s> = 2.
    Text is 'tokenised' by hashing whitespace-separated words (the real BPE vocabulary is not available offline)."""

    cls_token_id, pad_token_id, bos_token_id, eos_token_id = 50265, 2, 2, 2

    def __init__(self, vocab: int = 50274, num_img: int = 8):
        self.vocab, self.num_img = vocab, num_img
        self.img_base = vocab - num_img

    def __len__(self):
        return self.vocab

    def _encode(self, text: str):
        ids, i = [], 0
        while i < len(text):
            if text.startswith("[IMG", i) and "]" in text[i:]:
                j = text.index("]", i)
                tok = text[i + 4:j]
                if tok.isdigit() and int(tok) < self.num_img:
                    ids.append(self.img_base + int(tok))
                    i = j + 1
                    continue
            if text[i] == "\n":
                ids.append(50118)  # GPT-2 BPE id of "\n" in the OPT vocabulary
                i += 1
                continue
            j = i
            while j < len(text) and not text[j].isspace() and not text.startswith("[IMG", j):
                j += 1
            word = text[i:j]
            if word:
                h = 0
                for ch in word:
                    h = (h * 131 + ord(ch)) % 50000
                ids.append(3 + h)
            i = max(j, i + 1)
        return ids

    def __call__(self, text, add_special_tokens=True, return_tensors=None):
        ids = ([self.bos_token_id] if add_special_tokens else []) + self._encode(text)
        if return_tensors == "pt":
            return SimpleNamespace(input_ids=torch.tensor([ids], dtype=torch.int64))
        return SimpleNamespace(input_ids=ids)

    def batch_decode(self, ids, skip_special_tokens=True):
        out = []
        for row in ids.tolist():
            toks = []
            for t in row:
                if t >= self.img_base:
                    toks.append(f"[IMG{t - self.img_base}]")
                elif skip_special_tokens and t in (self.bos_token_id, self.cls_token_id):
                    continue
                else:
                    toks.append(f"<{t}>")
            out.append(" ".join(toks))
        return out


def model_args(num_visual_tokens: int = 4):
    """namedtuple like load_gill builds from model_args.json (gill/models.py:841-864)."""
    kw = dict(opt_version="facebook/opt-6.7b", freeze_lm=True, visual_encoder="openai/clip-vit-large-patch14",
              freeze_vm=True, n_visual_tokens=num_visual_tokens, ret_emb_dim=256, gen_emb_dim=768,
              text_emb_layers=[-1], text_fc_mode="gill_mapper", ret_text_fc_mode="linear", num_tokens=8,
              num_clip_tokens=77, share_ret_gen=True, norm_image_embed="none", retrieval_token_idx=list(IMG_IDS),
              gen_token_idx=list(IMG_IDS))
    return namedtuple("args", kw)(**kw)


def real_checkpoint_available() -> bool:
    return os.path.exists(os.path.join(CKPT_DIR, "pretrained_ckpt.pth.tar"))


def load_gill_trained_weights(gill, device):
    """Loads the shipped GILL-trained weights (gill/models.py:880-893) when present, else seeded synthetic ones.
    Returns 'real' or 'synthetic'."""
    lm = gill.model.lm
    if real_checkpoint_available():
        ck = torch.load(os.path.join(CKPT_DIR, "pretrained_ckpt.pth.tar"), map_location="cpu")
        sd = {k.replace("module.", ""): v for k, v in ck["state_dict"].items()}
        img = sd.pop("model.input_embeddings.weight")
        gill.load_state_dict(sd, strict=False)
        lm.embed[-8:].copy_(img.to(lm.embed.dtype))
        return "real"
    from oracle.mapper import synthetic_mapper_state_dict  # seeded stand-in shared with the tests

    gill.model.gen_text_hidden_fcs[0].load_state_dict(synthetic_mapper_state_dict(1234))
    g = torch.Generator().manual_seed(99)
    lm.embed[-8:].copy_((torch.randn(8, lm.D, generator=g) * 0.024).to(lm.embed.dtype))
    return "synthetic"


def build_sd(device="cuda", seed_unet=0, seed_vae=1, tiny=False):
    """SD-1.5-shaped UNet + VAE decoder with seeded PyTorch-default init (oracle/sd15.py holds the init so that the
    CPU oracle and the CUDA path see identical weights) and a seeded stand-in for the CLIP-text embedding of ""."""
    from oracle import sd15 as osd
    from gill_b200 import sd as psd

    ucfg, vcfg = (osd.tiny_unet_cfg(), osd.tiny_vae_cfg()) if tiny else (None, None)
    usd = osd.init_unet(seed_unet, ucfg)
    vsd = osd.init_vae_decoder(seed_vae, vcfg)
    unet = psd.UNetB200(usd, ucfg, device=device)
    vae = psd.VAEDecoderB200(vsd, vcfg, device=device)
    g = torch.Generator().manual_seed(77)
    neg = torch.randn(1, 77, 768, generator=g)
    return psd.StableDiffusionB200(unet, vae, neg), usd, vsd, neg


def build_gill(device="cuda", opt="opt-6.7b", tiny_sd=False, with_sd=True, seed=0):
    """The full pipeline object with synthetic frozen models + (real if present) GILL-trained weights."""
    from gill_b200.models import GILL
    from gill_b200.opt import OPTB200

    cfgs = {"opt-6.7b": dict(hidden=4096, layers=32, heads=32, ffn=16384),
            "opt-2l": dict(hidden=4096, layers=2, heads=32, ffn=16384)}
    c = cfgs[opt]
    tok = SyntheticTokenizer()
    lm = OPTB200.random_init(c["hidden"], c["layers"], c["heads"], c["ffn"], vocab=len(tok), seed=seed, device=device)
    sd_pipe = build_sd(device, tiny=tiny_sd)[0] if with_sd else None
    gill = GILL(tok, model_args(), load_sd=with_sd, num_gen_images=1, lm=lm, sd_pipe=sd_pipe)
    gill = gill.eval().to(device)
    kind = load_gill_trained_weights(gill, device)
    return gill, kind


# ---- synthetic retrieval banks: identical on every machine / GPU count ------------------------
BANK_CHUNKS = 8


def synthetic_bank_chunk(c: int, rows: int, d: int, exact: bool = False, device="cpu") -> torch.Tensor:
    """Chunk c of the synthetic bank: seed 7000+c, randn -> row-normalise -> x14.24 -> bf16 (tier B), or values
    randint(-4,5)/8 (tier A: exactly representable, order-independent fp32 sums). `device` selects the generator too:
    CPU chunks are what the oracle tests use, CUDA chunks (a different but equally deterministic stream) fill the
    3M-row benchmark bank quickly; a given (c, rows, d, device type) always yields the same rows, whatever the GPU count."""
    g = torch.Generator(device=device).manual_seed(7000 + c)
    if exact:
        return (torch.randint(-4, 5, (rows, d), generator=g, device=device).float() / 8).bfloat16()
    m = torch.randn(rows, d, generator=g, device=device)
    m = m / m.norm(dim=1, keepdim=True)
    return (m * 14.24).bfloat16()


def synthetic_queries(q: int, d: int, exact: bool = False, seed: int = 8) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    if exact:
        return (torch.randint(-4, 5, (q, d), generator=g).float() / 8).bfloat16()
    m = torch.randn(q, d, generator=g)
    return (m / m.norm(dim=1, keepdim=True)).bfloat16()


def synthetic_bank_shard(n_total: int, d: int, world: int, rank: int, exact: bool = False, device="cuda") -> torch.Tensor:
    """Rows [rank*n/world, (rank+1)*n/world) of the 8-chunk synthetic bank (world must divide 8, n_total % 8 == 0)."""
    assert BANK_CHUNKS % world == 0 and n_total % BANK_CHUNKS == 0
    per, rows = BANK_CHUNKS // world, n_total // BANK_CHUNKS
    return torch.cat([synthetic_bank_chunk(c, rows, d, exact, device) for c in range(rank * per, (rank + 1) * per)], 0)


def build_clip_tower(device="cuda", layers: int = 24, seed: int = 4):
    """CLIP ViT-L/14-shaped vision tower (seeded random init; the real weights cannot be downloaded offline) on the B200
    kernels, for the re-rank step (gill/models.py:733-751) and raw-pixel prompts."""
    from gill_b200.clip import CLIPVisionB200
    from oracle import clip as oclip

    cfg = dict(oclip.CLIP_L14, layers=layers)
    return CLIPVisionB200(oclip.init_clip(cfg, seed=seed), cfg["hidden"], cfg["layers"], cfg["heads"], cfg["mlp"],
                          cfg["patch"], cfg["image"], device=device)


def config5_prompts(n: int, seed: int):
    """BASELINE configs[4] prompts: 2 CLIP-encoded images (pooled 1024-d features) + ~64 text tokens each, of UNEQUAL
    length (56..72 tokens) so that the batched prefill really is ragged."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        f1, f2 = torch.randn(1024, generator=g), torch.randn(1024, generator=g)
        n1 = 24 + int(torch.randint(0, 9, (1,), generator=g))
        n2 = 32 + int(torch.randint(0, 9, (1,), generator=g))
        w = lambda k, o: " ".join(f"w{seed}x{i}y{o + j}" for j in range(k))
        out.append([f1, w(n1, 0), f2, w(n2, 100)])
    return out
