"""ORACLE (test infrastructure, not product code): CPU restatement of the Stable Diffusion safety checker the reference
reaches through `sd_pipe(...)` (gill/models.py:730 -> gill/custom_sd.py:375-383 `run_safety_checker`; module:
diffusers==0.17.1 `StableDiffusionSafetyChecker.forward`, NOT under /root/reference and not installable offline --
PARITY UNPINNED for the module math, restated from the published implementation; control flow per custom_sd.py).

CLIP ViT-L/14 pooled output -> visual_projection (1024 -> 768, no bias) -> cosine similarity against 3 "special care" and
17 concept embeddings -> per-concept thresholds (special-care hits lower the concept thresholds by 0.01) -> images
with any concept score > 0 are replaced by black images.
"""
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import clip as oclip

SD = Dict[str, torch.Tensor]


def init_safety_checker(cfg, seed: int = 0, n_concepts: int = 17, n_special: int = 3, proj: int = 768) -> SD:
    g = torch.Generator().manual_seed(seed)
    sd = {"vision_model." + k: v for k, v in oclip.init_clip(cfg, seed).items()}      # vision_model.vision_model.*
    sd["visual_projection.weight"] = torch.randn(proj, cfg["hidden"], generator=g) * 0.03
    sd["concept_embeds"] = torch.randn(n_concepts, proj, generator=g)
    sd["special_care_embeds"] = torch.randn(n_special, proj, generator=g)
    sd["concept_embeds_weights"] = torch.full((n_concepts,), 0.02)
    sd["special_care_embeds_weights"] = torch.full((n_special,), 0.03)
    return sd


def flags_from_cosines(special_cos, cos, special_thr, thr) -> List[bool]:
    """The per-image threshold logic (host side in diffusers as well)."""
    out = []
    for i in range(cos.shape[0]):
        adjustment = 0.0
        for c in range(special_cos.shape[1]):
            if round(float(special_cos[i, c]) - float(special_thr[c]) + adjustment, 3) > 0:
                adjustment = 0.01
        bad = [c for c in range(cos.shape[1]) if round(float(cos[i, c]) - float(thr[c]) + adjustment, 3) > 0]
        out.append(len(bad) > 0)
    return out


def safety_check(sd: SD, clip_input: torch.Tensor, cfg) -> Tuple[List[bool], torch.Tensor, torch.Tensor]:
    """clip_input: CLIP pixel_values [B,3,S,S]. Returns (has_nsfw per image, special cosines, concept cosines)."""
    vsd = {k[len("vision_model."):]: v for k, v in sd.items() if k.startswith("vision_model.")}
    _, pooled = oclip.clip_vision_forward(vsd, clip_input, cfg)
    emb = F.linear(pooled, sd["visual_projection.weight"])
    n = lambda x: x / x.norm(dim=-1, keepdim=True)
    special_cos = n(emb) @ n(sd["special_care_embeds"]).T
    cos = n(emb) @ n(sd["concept_embeds"]).T
    flags = flags_from_cosines(special_cos, cos, sd["special_care_embeds_weights"], sd["concept_embeds_weights"])
    return flags, special_cos, cos
