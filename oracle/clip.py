"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the CLIP vision tower the reference calls at
gill/models.py:129-143 (`self.visual_model(pixel_values).pooler_output`, `CLIPVisionModel` of transformers==4.30.2,
`openai/clip-vit-large-patch14`; module math in transformers/models/clip/modeling_clip.py -- not under
/root/reference, the installed transformers 5.5 copy is the stand-in: tests/test_oracle.py pins this file against it).

ViT: patch conv (stride = kernel = patch, no bias) -> [CLS | patches] + learned positions -> pre-LN -> L x
[LN -> MHA(heads, scale hd^-0.5) -> + ; LN -> fc1 -> quick_gelu -> fc2 -> +] -> post-LN of the CLS token = pooler_output.
State-dict names are HF's (`vision_model.*`).
"""
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

CLIP_L14 = dict(hidden=1024, layers=24, heads=16, mlp=4096, patch=14, image=224)


def tiny_cfg():
    return dict(hidden=256, layers=2, heads=4, mlp=1024, patch=14, image=56)


def init_clip(cfg, seed: int = 0) -> SD:
    """Seeded random weights with HF-like scales (no pretrained weights exist offline)."""
    g = torch.Generator().manual_seed(seed)
    h, m, p = cfg["hidden"], cfg["mlp"], cfg["patch"]
    n = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    npos = (cfg["image"] // p) ** 2 + 1
    v = "vision_model."
    sd = {v + "embeddings.class_embedding": n(h), v + "embeddings.patch_embedding.weight": n(h, 3, p, p),
          v + "embeddings.position_embedding.weight": n(npos, h)}
    for nm in ("pre_layrnorm", "post_layernorm"):
        sd[v + nm + ".weight"] = 1 + n(h, std=0.05)
        sd[v + nm + ".bias"] = n(h, std=0.05)
    for i in range(cfg["layers"]):
        l = f"{v}encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[l + f"self_attn.{nm}.weight"], sd[l + f"self_attn.{nm}.bias"] = n(h, h, std=0.03), n(h)
        sd[l + "mlp.fc1.weight"], sd[l + "mlp.fc1.bias"] = n(m, h, std=0.03), n(m)
        sd[l + "mlp.fc2.weight"], sd[l + "mlp.fc2.bias"] = n(h, m, std=0.03), n(h)
        for nm in ("layer_norm1", "layer_norm2"):
            sd[l + nm + ".weight"] = 1 + n(h, std=0.05)
            sd[l + nm + ".bias"] = n(h, std=0.05)
    return sd


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def clip_vision_forward(sd: SD, pixel_values: torch.Tensor, cfg):
    """pixel_values [B,3,S,S] -> (last_hidden_state [B,1+P,h], pooler_output [B,h])."""
    v = "vision_model."
    h, heads, p = cfg["hidden"], cfg["heads"], cfg["patch"]
    B = pixel_values.shape[0]
    x = F.conv2d(pixel_values, sd[v + "embeddings.patch_embedding.weight"], stride=p)      # [B,h,g,g]
    x = x.flatten(2).transpose(1, 2)                                                         # [B,P,h]
    cls = sd[v + "embeddings.class_embedding"].expand(B, 1, h)
    x = torch.cat([cls, x], 1) + sd[v + "embeddings.position_embedding.weight"][None]
    x = F.layer_norm(x, (h,), sd[v + "pre_layrnorm.weight"], sd[v + "pre_layrnorm.bias"], 1e-5)
    hd = h // heads
    for i in range(cfg["layers"]):
        l = f"{v}encoder.layers.{i}."
        n = F.layer_norm(x, (h,), sd[l + "layer_norm1.weight"], sd[l + "layer_norm1.bias"], 1e-5)
        q, k, vv = (F.linear(n, sd[l + f"self_attn.{nm}.weight"], sd[l + f"self_attn.{nm}.bias"])
                    .view(B, -1, heads, hd).transpose(1, 2) for nm in ("q_proj", "k_proj", "v_proj"))
        a = torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2), -1) @ vv
        a = a.transpose(1, 2).reshape(B, -1, h)
        x = x + F.linear(a, sd[l + "self_attn.out_proj.weight"], sd[l + "self_attn.out_proj.bias"])
        n = F.layer_norm(x, (h,), sd[l + "layer_norm2.weight"], sd[l + "layer_norm2.bias"], 1e-5)
        f = quick_gelu(F.linear(n, sd[l + "mlp.fc1.weight"], sd[l + "mlp.fc1.bias"]))
        x = x + F.linear(f, sd[l + "mlp.fc2.weight"], sd[l + "mlp.fc2.bias"])
    pooled = F.layer_norm(x[:, 0], (h,), sd[v + "post_layernorm.weight"], sd[v + "post_layernorm.bias"], 1e-5)
    return x, pooled


# ---------------------------------------------------------------------------------------------------------------
# Pre-processing of generated images for the re-rank step (gill/models.py:733-737; gill/utils.py:117-119):
# PIL `img.resize((224, 224))` (bicubic) + HF CLIP feature extractor (rescale 1/255, normalise). The resize is restated
# from Pillow's libImaging/Resample.c (8-bit path: double coefficients -> 22-bit fixed point, horizontal then vertical
# pass, each rounded and clipped to uint8); tests/test_oracle.py pins it bit-exactly against PIL itself.
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _pil_bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _pil_coeffs(in_size: int, out_size: int):
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    out = []
    for o in range(out_size):
        center = (o + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_pil_bicubic((x + xmin - center + 0.5) / filterscale) for x in range(xmax)]
        ww = sum(w)
        k = [wi / ww if ww != 0.0 else wi for wi in w]
        out.append((xmin, [int(-0.5 + ki * (1 << 22)) if ki < 0 else int(0.5 + ki * (1 << 22)) for ki in k]))
    return out


def pil_bicubic_resize_u8(img, size, size_w=None):
    """img: uint8 numpy [H,W,3] -> uint8 [size, size_w or size, 3], bit-identical to PIL.Image.resize((w, h), BICUBIC)."""
    import numpy as np

    out_h, out_w = size, (size if size_w is None else size_w)
    H, W, _ = img.shape
    a = img.astype(np.int64)
    tmp = np.empty((H, out_w, 3), dtype=np.int64)
    for o, (x0, kk) in enumerate(_pil_coeffs(W, out_w)):
        acc = (1 << 21) + (a[:, x0:x0 + len(kk), :] * np.asarray(kk, dtype=np.int64)[None, :, None]).sum(1)
        tmp[:, o, :] = np.clip(acc >> 22, 0, 255)
    out = np.empty((out_h, out_w, 3), dtype=np.int64)
    for o, (y0, kk) in enumerate(_pil_coeffs(H, out_h)):
        acc = (1 << 21) + (tmp[y0:y0 + len(kk), :, :] * np.asarray(kk, dtype=np.int64)[:, None, None]).sum(0)
        out[o] = np.clip(acc >> 22, 0, 255)
    return out.astype(np.uint8)


def hf_clip_geometry(H: int, W: int, size: int = 224):
    """transformers==4.30.2 CLIPImageProcessor (openai/clip-vit-large-patch14 preprocessor_config.json: resize shortest edge
    224 bicubic, center crop 224): get_resize_output_image_size(default_to_square=False) + image_transforms.center_crop."""
    if H <= W:
        RH, RW = size, int(size * W / H)
    else:
        RH, RW = int(size * H / W), size
    return RH, RW, (RH - size) // 2, (RW - size) // 2


def clip_feature_extractor(img, size: int = 224):
    """What `utils.get_pixel_values_for_model(feature_extractor, img)` (gill/utils.py:117-119) returns for a uint8 RGB image
    [H,W,3]: resize shortest edge (PIL bicubic, 8-bit) -> centre crop -> /255 -> normalise. float32 [3,size,size] plus the
    cropped uint8 image."""
    import numpy as np

    H, W, _ = img.shape
    RH, RW, top, left = hf_clip_geometry(H, W, size)
    r = pil_bicubic_resize_u8(img, RH, RW)[top:top + size, left:left + size]
    x = (r.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.asarray(CLIP_MEAN, dtype=np.float32)) / np.asarray(CLIP_STD, dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1))), r


def clip_preprocess(img, size: int = 224):
    """uint8 [H,W,3] -> float32 [3,size,size] pixel_values (HF CLIPImageProcessor: rescale in float64 -> float32, then
    (x - mean) / std in float32)."""
    import numpy as np

    r = pil_bicubic_resize_u8(img, size)
    x = (r.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.asarray(CLIP_MEAN, dtype=np.float32)) / np.asarray(CLIP_STD, dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1)))
