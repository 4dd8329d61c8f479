"""CLIP ViT vision tower on the libgillb200 kernels: the `self.visual_model(pixel_values).pooler_output` call of
gill/models.py:129-143 (prefix encoding) and :733-751 (re-ranking of generated images) -- SURVEY 8f-1.

Drop-in for transformers' `CLIPVisionModel` at that call site: built from its state dict (HF names, `vision_model.*`),
called with `pixel_values [B,3,S,S]`, returns an object with `.pooler_output [B,hidden]` and `.last_hidden_state`.
Same structure as OPTB200: bf16 GEMM operands (the reference runs the tower in bf16, models.py:876), fp32 residual
stream / LayerNorm statistics / softmax, fused-QKV tcgen05 GEMM, the 4-warp flash-attention kernel (head dim 64, no
padding), bias + quick_gelu + residual in the GEMM epilogues. The patch embedding (stride = kernel = 14, no bias) is a
GEMM over patch rows [B*P, 3*14*14 -> padded to a multiple of 8]; cutting the patches is a view + one copy.
"""
from types import SimpleNamespace
from typing import Dict

import torch

from . import ops

SD = Dict[str, torch.Tensor]


class CLIPVisionB200:
    def __init__(self, sd: SD, hidden: int = 1024, layers: int = 24, heads: int = 16, mlp: int = 4096, patch: int = 14,
                 image: int = 224, device="cuda", dtype=torch.bfloat16):
        if hidden // heads != 64:
            raise ValueError(f"CLIPVisionB200 supports head_dim 64 (ViT-B/L families), got {hidden // heads}")
        self.D, self.L, self.H, self.F, self.P, self.S = hidden, layers, heads, mlp, patch, image
        self.dev, self.dt = torch.device(device), dtype
        f32 = torch.float32
        v = "vision_model."
        g = lambda k, dt=None: sd[k].to(self.dev, dt or dtype).contiguous()
        w = sd[v + "embeddings.patch_embedding.weight"].reshape(hidden, -1)            # [h, 3*p*p], k = (c, py, px)
        self.kp = (w.shape[1] + 7) // 8 * 8                                           # 16-byte row pitch for TMA
        wp = torch.zeros((hidden, self.kp), dtype=w.dtype)
        wp[:, : w.shape[1]] = w
        self.patch_w = wp.to(self.dev, dtype).contiguous()
        # [CLS | patches] + positions: the class embedding is folded into position 0
        pos = sd[v + "embeddings.position_embedding.weight"].float().clone()
        pos[0] += sd[v + "embeddings.class_embedding"].float()
        self.pos = pos.to(self.dev).contiguous()                                      # [1+P, h] fp32
        self.pre_w, self.pre_b = g(v + "pre_layrnorm.weight", f32), g(v + "pre_layrnorm.bias", f32)
        self.post_w, self.post_b = g(v + "post_layernorm.weight", f32), g(v + "post_layernorm.bias", f32)
        self.layers = []
        for i in range(layers):
            l = f"{v}encoder.layers.{i}."
            self.layers.append(dict(
                ln1_w=g(l + "layer_norm1.weight", f32), ln1_b=g(l + "layer_norm1.bias", f32),
                qkv_w=torch.cat([g(l + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")], 0),
                qkv_b=torch.cat([g(l + f"self_attn.{n}.bias", f32) for n in ("q_proj", "k_proj", "v_proj")], 0),
                o_w=g(l + "self_attn.out_proj.weight"), o_b=g(l + "self_attn.out_proj.bias", f32),
                ln2_w=g(l + "layer_norm2.weight", f32), ln2_b=g(l + "layer_norm2.bias", f32),
                fc1_w=g(l + "mlp.fc1.weight"), fc1_b=g(l + "mlp.fc1.bias", f32),
                fc2_w=g(l + "mlp.fc2.weight"), fc2_b=g(l + "mlp.fc2.bias", f32)))

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor):
        """pixel_values [B,3,S,S] -> (last_hidden_state [B,1+P,h] fp32 residual stream, pooler_output [B,h] fp32)."""
        if not pixel_values.is_cuda:
            raise RuntimeError("gill_b200.CLIPVisionB200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, C, S, _ = pixel_values.shape
        p, D, H = self.P, self.D, self.H
        if C != 3 or S != self.S:
            raise ValueError(f"expected pixel_values [B,3,{self.S},{self.S}], got {tuple(pixel_values.shape)}")
        gsz = S // p
        T = gsz * gsz + 1
        # patch rows [B*P, (c,py,px)] (+ zero padding of K); row 0 of every image is the (all-zero) CLS slot
        rows = torch.zeros((B, T, self.kp), device=self.dev, dtype=self.dt)
        patches = pixel_values.to(self.dev, self.dt).view(B, 3, gsz, p, gsz, p).permute(0, 2, 4, 1, 3, 5)
        rows[:, 1:, : 3 * p * p].copy_(patches.reshape(B, gsz * gsz, 3 * p * p))
        x = ops.gemm(rows.view(B * T, self.kp), self.patch_w, out_dtype=torch.float32)          # CLS rows stay 0
        h = ops.cast_add(x, self.pos, torch.float32, y_period=T * D)                             # + positions (+ CLS)
        h = ops.layernorm(h, self.pre_w, self.pre_b, 1e-5, out_dtype=torch.float32)
        for ly in self.layers:
            n = ops.layernorm(h, ly["ln1_w"], ly["ln1_b"], 1e-5, out_dtype=self.dt)
            qkv = ops.gemm(n, ly["qkv_w"], bias=ly["qkv_b"]).view(B, T, 3 * D)
            a = ops.attention(qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:], H, 64, 64 ** -0.5)
            h = ops.gemm(a.view(B * T, D), ly["o_w"], bias=ly["o_b"], residual=h, out_dtype=torch.float32)
            n = ops.layernorm(h, ly["ln2_w"], ly["ln2_b"], 1e-5, out_dtype=self.dt)
            f = ops.gemm(n, ly["fc1_w"], bias=ly["fc1_b"], act="quick_gelu")
            h = ops.gemm(f, ly["fc2_w"], bias=ly["fc2_b"], residual=h, out_dtype=torch.float32)
        hs = h.view(B, T, D)
        pooled = ops.layernorm(hs[:, 0, :].contiguous(), self.post_w, self.post_b, 1e-5, out_dtype=torch.float32)
        return hs, pooled

    def __call__(self, pixel_values=None, **kw):
        hs, pooled = self.forward(pixel_values)
        return SimpleNamespace(last_hidden_state=hs, pooler_output=pooled)


class SafetyCheckerB200:
    """`StableDiffusionSafetyChecker` (diffusers) for `sd_pipe` (gill/custom_sd.py:375-383): CLIP tower -> visual_projection
    -> cosine against the concept embeddings -> thresholds; flagged images are blacked out by the caller. Built from the
    checker's own state dict (`vision_model.vision_model.*`, `visual_projection.weight`, `concept_embeds`,
    `special_care_embeds`, `*_weights`). Takes the generated uint8 images on the device: the feature-extractor step
    (PIL bicubic resize + normalise) is `ops.clip_preprocess_u8`."""

    def __init__(self, sd: SD, hidden: int = 1024, layers: int = 24, heads: int = 16, mlp: int = 4096, patch: int = 14,
                 image: int = 224, device="cuda", dtype=torch.float16):
        vsd = {k[len("vision_model."):]: v for k, v in sd.items() if k.startswith("vision_model.")}
        self.tower = CLIPVisionB200(vsd, hidden, layers, heads, mlp, patch, image, device=device, dtype=dtype)
        self.dev, self.dt, self.image = self.tower.dev, dtype, image
        self.proj = sd["visual_projection.weight"].to(self.dev, dtype).contiguous()
        embeds = torch.cat([sd["special_care_embeds"], sd["concept_embeds"]], 0).float()
        self.n_special = sd["special_care_embeds"].shape[0]
        self.embeds = (embeds / embeds.norm(dim=-1, keepdim=True)).to(self.dev, dtype).contiguous()   # [3+17, 768]
        self.special_thr = sd["special_care_embeds_weights"].float().tolist()
        self.thr = sd["concept_embeds_weights"].float().tolist()

    @torch.no_grad()
    def __call__(self, images_u8: torch.Tensor):
        """images_u8: uint8 NHWC [B,H,W,3] (square) on the device -> list of bool (True = flagged)."""
        B, H, W, _ = images_u8.shape
        if H != W:
            raise ValueError("SafetyCheckerB200 expects square images (the CLIP processor's centre crop is not built)")
        px = ops.clip_preprocess_u8(images_u8.contiguous(), self.image, out_dtype=self.dt)
        _, pooled = self.tower.forward(px)
        emb = ops.gemm(pooled.to(self.dt).contiguous(), self.proj, out_dtype=torch.float32)       # [B, 768]
        emb = ops.l2norm_rows(emb, self.dt)
        cos = ops.gemm(emb, self.embeds, out_dtype=torch.float32).cpu()                            # [B, 3+17]
        flags = []
        for i in range(B):                                                                         # host logic, as upstream
            adjustment = 0.0
            for c in range(self.n_special):
                if round(cos[i, c].item() - self.special_thr[c] + adjustment, 3) > 0:
                    adjustment = 0.01
            flags.append(any(round(cos[i, self.n_special + c].item() - self.thr[c] + adjustment, 3) > 0
                             for c in range(len(self.thr))))
        return flags
