"""ORACLE (test infrastructure only -- never imported by the product path).  PARITY UNPINNED for module math.

CPU restatement (PyTorch fp32, functional, NCHW like the original) of what `self.sd_pipe(prompt_embeds=...)`
(gill/models.py:730) runs: the stock diffusers==0.17.1 StableDiffusionPipeline, whose control flow is restated
in-repo by gill/custom_sd.py:567-666. The module arithmetic (UNet2DConditionModel, AutoencoderKL.decode,
PNDMScheduler) lives in diffusers==0.17.1 (requirements.txt:9), which is NOT vendored under /root/reference and not
installable offline; it is restated here from the published SD-1.5 configs. What pins it:
  - parameter counts match the published models exactly (UNet 859,520,964; VAE decoder + post_quant 49,490,199),
    checked in tests/test_oracle_sd15.py;
  - state-dict key names follow diffusers' so real `runwayml/stable-diffusion-v1-5` weights load unchanged;
  - the loop follows gill/custom_sd.py:626-651 line by line.
Weights here are seeded PyTorch-default-init (no pretrained weights exist offline).
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------------------------------
# configs (runwayml/stable-diffusion-v1-5: unet/config.json, vae/config.json, scheduler/scheduler_config.json)
# --------------------------------------------------------------------------------------------------------------
UNET_CFG = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                cross_attention_dim=768, heads=8, norm_groups=32, has_attn_down=(True, True, True, False),
                has_attn_up=(False, True, True, True))
VAE_CFG = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
               norm_groups=32, scaling_factor=0.18215)


def tiny_unet_cfg():
    """A structurally identical but small UNet used by fast parity tests (all block types, 3 resolutions)."""
    return dict(in_channels=4, out_channels=4, block_out_channels=(64, 128, 128), layers_per_block=1,
                cross_attention_dim=768, heads=2, norm_groups=32, has_attn_down=(True, True, False),
                has_attn_up=(False, True, True))


def tiny_vae_cfg():
    return dict(latent_channels=4, out_channels=3, block_out_channels=(64, 128), layers_per_block=1,
                norm_groups=32, scaling_factor=0.18215)


# --------------------------------------------------------------------------------------------------------------
# seeded default-init weights with diffusers' key names
# --------------------------------------------------------------------------------------------------------------
class _Init:
    def __init__(self, seed: int, dtype=torch.float32):
        self.g = torch.Generator().manual_seed(seed)
        self.sd: SD = {}
        self.dtype = dtype

    def _u(self, shape, bound):
        return ((torch.rand(shape, generator=self.g) * 2 - 1) * bound).to(self.dtype)

    def conv(self, name, cin, cout, k, bias=True):
        fan_in = cin * k * k
        b = 1.0 / math.sqrt(fan_in)  # nn.Conv2d default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
        self.sd[name + ".weight"] = self._u((cout, cin, k, k), b)
        if bias:
            self.sd[name + ".bias"] = self._u((cout,), b)

    def linear(self, name, cin, cout, bias=True):
        b = 1.0 / math.sqrt(cin)
        self.sd[name + ".weight"] = self._u((cout, cin), b)
        if bias:
            self.sd[name + ".bias"] = self._u((cout,), b)

    def norm(self, name, c):
        # default init is weight=1, bias=0; perturb slightly so parity tests exercise the affine terms
        self.sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=self.g)).to(self.dtype)
        self.sd[name + ".bias"] = (0.1 * torch.randn(c, generator=self.g)).to(self.dtype)


def _init_resnet(I: _Init, p, cin, cout, temb_ch: Optional[int]):
    I.norm(p + ".norm1", cin)
    I.conv(p + ".conv1", cin, cout, 3)
    if temb_ch:
        I.linear(p + ".time_emb_proj", temb_ch, cout)
    I.norm(p + ".norm2", cout)
    I.conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        I.conv(p + ".conv_shortcut", cin, cout, 1)


def _init_transformer(I: _Init, p, c, ctx_dim):
    I.norm(p + ".norm", c)
    I.conv(p + ".proj_in", c, c, 1)
    t = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        I.norm(t + "." + n, c)
    for a, kv in (("attn1", c), ("attn2", ctx_dim)):
        I.linear(f"{t}.{a}.to_q", c, c, bias=False)
        I.linear(f"{t}.{a}.to_k", kv, c, bias=False)
        I.linear(f"{t}.{a}.to_v", kv, c, bias=False)
        I.linear(f"{t}.{a}.to_out.0", c, c)
    I.linear(t + ".ff.net.0.proj", c, 8 * c)
    I.linear(t + ".ff.net.2", 4 * c, c)
    I.conv(p + ".proj_out", c, c, 1)


def init_unet(seed: int = 0, cfg=None, dtype=torch.float32) -> SD:
    cfg = cfg or UNET_CFG
    I = _Init(seed, dtype)
    boc = cfg["block_out_channels"]
    temb = boc[0] * 4
    I.linear("time_embedding.linear_1", boc[0], temb)
    I.linear("time_embedding.linear_2", temb, temb)
    I.conv("conv_in", cfg["in_channels"], boc[0], 3)
    ch = boc[0]
    skips = [ch]
    for i, c in enumerate(boc):
        for j in range(cfg["layers_per_block"]):
            _init_resnet(I, f"down_blocks.{i}.resnets.{j}", ch, c, temb)
            ch = c
            if cfg["has_attn_down"][i]:
                _init_transformer(I, f"down_blocks.{i}.attentions.{j}", c, cfg["cross_attention_dim"])
            skips.append(ch)
        if i < len(boc) - 1:
            I.conv(f"down_blocks.{i}.downsamplers.0.conv", ch, ch, 3)
            skips.append(ch)
    _init_resnet(I, "mid_block.resnets.0", ch, ch, temb)
    _init_transformer(I, "mid_block.attentions.0", ch, cfg["cross_attention_dim"])
    _init_resnet(I, "mid_block.resnets.1", ch, ch, temb)
    rev = list(reversed(boc))
    for i, c in enumerate(rev):
        for j in range(cfg["layers_per_block"] + 1):
            s = skips.pop()
            _init_resnet(I, f"up_blocks.{i}.resnets.{j}", ch + s, c, temb)
            ch = c
            if cfg["has_attn_up"][i]:
                _init_transformer(I, f"up_blocks.{i}.attentions.{j}", c, cfg["cross_attention_dim"])
        if i < len(boc) - 1:
            I.conv(f"up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
    I.norm("conv_norm_out", ch)
    I.conv("conv_out", ch, cfg["out_channels"], 3)
    return I.sd


def init_vae_decoder(seed: int = 1, cfg=None, dtype=torch.float32) -> SD:
    cfg = cfg or VAE_CFG
    I = _Init(seed, dtype)
    boc = cfg["block_out_channels"]
    lc = cfg["latent_channels"]
    I.conv("post_quant_conv", lc, lc, 1)
    top = boc[-1]
    I.conv("decoder.conv_in", lc, top, 3)
    _init_resnet(I, "decoder.mid_block.resnets.0", top, top, None)
    a = "decoder.mid_block.attentions.0"
    I.norm(a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        I.linear(f"{a}.{n}", top, top)
    _init_resnet(I, "decoder.mid_block.resnets.1", top, top, None)
    ch = top
    rev = list(reversed(boc))
    for i, c in enumerate(rev):
        for j in range(cfg["layers_per_block"] + 1):
            _init_resnet(I, f"decoder.up_blocks.{i}.resnets.{j}", ch, c, None)
            ch = c
        if i < len(boc) - 1:
            I.conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
    I.norm("decoder.conv_norm_out", ch)
    I.conv("decoder.conv_out", ch, cfg["out_channels"], 3)
    return I.sd


def param_count(sd: SD) -> int:
    return sum(v.numel() for v in sd.values())


# --------------------------------------------------------------------------------------------------------------
# UNet forward
# --------------------------------------------------------------------------------------------------------------
def _conv(x, sd, p, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _gn(x, sd, p, groups, eps):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): cat([cos, sin])."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def resnet_block(x, temb, sd, p, groups, eps):
    h = _conv(F.silu(_gn(x, sd, p + ".norm1", groups, eps)), sd, p + ".conv1")
    if temb is not None:
        h = h + _lin(F.silu(temb), sd, p + ".time_emb_proj")[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, p + ".norm2", groups, eps)), sd, p + ".conv2")
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(x, sd, p + ".conv_shortcut", padding=0)
    return x + h


def attention(x, ctx, sd, p, heads):
    """diffusers Attention (AttnProcessor): q/k/v without bias, softmax(q k^T / sqrt(hd)) v, to_out.0 with bias."""
    q, k, v = _lin(x, sd, p + ".to_q"), _lin(ctx, sd, p + ".to_k"), _lin(ctx, sd, p + ".to_v")
    B, L, C = q.shape
    hd = C // heads
    q = q.view(B, L, heads, hd).transpose(1, 2)
    k = k.view(B, -1, heads, hd).transpose(1, 2)
    v = v.view(B, -1, heads, hd).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, L, C)
    return _lin(o, sd, p + ".to_out.0")


def transformer_2d(x, ctx, sd, p, heads, groups):
    B, C, H, W = x.shape
    res = x
    h = _conv(_gn(x, sd, p + ".norm", groups, 1e-6), sd, p + ".proj_in", padding=0)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    t = p + ".transformer_blocks.0"
    ln = lambda z, n: F.layer_norm(z, (C,), sd[f"{t}.{n}.weight"], sd[f"{t}.{n}.bias"], 1e-5)
    n1 = ln(h, "norm1")
    h = h + attention(n1, n1, sd, t + ".attn1", heads)
    h = h + attention(ln(h, "norm2"), ctx, sd, t + ".attn2", heads)
    g = _lin(ln(h, "norm3"), sd, t + ".ff.net.0.proj")
    a, gate = g.chunk(2, dim=-1)
    h = h + _lin(a * F.gelu(gate), sd, t + ".ff.net.2")
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _conv(h, sd, p + ".proj_out", padding=0) + res


def unet_forward(sd: SD, sample: torch.Tensor, t, ctx: torch.Tensor, cfg=None) -> torch.Tensor:
    """UNet2DConditionModel.forward(sample [B,4,H,W], timestep, encoder_hidden_states [B,77,768]) -> eps."""
    cfg = cfg or UNET_CFG
    boc, G, heads = cfg["block_out_channels"], cfg["norm_groups"], cfg["heads"]
    B = sample.shape[0]
    tt = torch.as_tensor(t, dtype=torch.float32, device=sample.device).reshape(-1).expand(B)
    temb = timestep_embedding(tt, boc[0]).to(sample.dtype)
    temb = _lin(F.silu(_lin(temb, sd, "time_embedding.linear_1")), sd, "time_embedding.linear_2")
    h = _conv(sample, sd, "conv_in")
    skips = [h]
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"]):
            h = resnet_block(h, temb, sd, f"down_blocks.{i}.resnets.{j}", G, 1e-5)
            if cfg["has_attn_down"][i]:
                h = transformer_2d(h, ctx, sd, f"down_blocks.{i}.attentions.{j}", heads, G)
            skips.append(h)
        if i < len(boc) - 1:
            h = _conv(h, sd, f"down_blocks.{i}.downsamplers.0.conv", stride=2)
            skips.append(h)
    h = resnet_block(h, temb, sd, "mid_block.resnets.0", G, 1e-5)
    h = transformer_2d(h, ctx, sd, "mid_block.attentions.0", heads, G)
    h = resnet_block(h, temb, sd, "mid_block.resnets.1", G, 1e-5)
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"] + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(h, temb, sd, f"up_blocks.{i}.resnets.{j}", G, 1e-5)
            if cfg["has_attn_up"][i]:
                h = transformer_2d(h, ctx, sd, f"up_blocks.{i}.attentions.{j}", heads, G)
        if i < len(boc) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, f"up_blocks.{i}.upsamplers.0.conv")
    h = F.silu(_gn(h, sd, "conv_norm_out", G, 1e-5))
    return _conv(h, sd, "conv_out")


# --------------------------------------------------------------------------------------------------------------
# VAE decoder (gill/custom_sd.py:385-392)
# --------------------------------------------------------------------------------------------------------------
def vae_decode(sd: SD, latents: torch.Tensor, cfg=None) -> torch.Tensor:
    """decode_latents: latents / 0.18215 -> post_quant_conv -> decoder -> (x/2+0.5).clamp(0,1); returns NCHW [0,1]."""
    cfg = cfg or VAE_CFG
    G = cfg["norm_groups"]
    boc = cfg["block_out_channels"]
    z = latents / cfg["scaling_factor"]                              # custom_sd.py:386
    z = _conv(z, sd, "post_quant_conv", padding=0)
    h = _conv(z, sd, "decoder.conv_in")
    h = resnet_block(h, None, sd, "decoder.mid_block.resnets.0", G, 1e-6)
    a = "decoder.mid_block.attentions.0"
    B, C, H, W = h.shape
    n = _gn(h, sd, a + ".group_norm", G, 1e-6).reshape(B, C, H * W).transpose(1, 2)
    q, k, v = _lin(n, sd, a + ".to_q"), _lin(n, sd, a + ".to_k"), _lin(n, sd, a + ".to_v")
    att = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1)
    o = _lin(att @ v, sd, a + ".to_out.0").transpose(1, 2).reshape(B, C, H, W)
    h = h + o
    h = resnet_block(h, None, sd, "decoder.mid_block.resnets.1", G, 1e-6)
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"] + 1):
            h = resnet_block(h, None, sd, f"decoder.up_blocks.{i}.resnets.{j}", G, 1e-6)
        if i < len(boc) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, f"decoder.up_blocks.{i}.upsamplers.0.conv")
    h = F.silu(_gn(h, sd, "decoder.conv_norm_out", G, 1e-6))
    img = _conv(h, sd, "decoder.conv_out")
    return (img / 2 + 0.5).clamp(0, 1)                               # custom_sd.py:389


def to_uint8_nhwc(img01: torch.Tensor) -> torch.Tensor:
    """custom_sd.py:391 + numpy_to_pil: NHWC, (x*255).round().astype(uint8)."""
    return (img01.permute(0, 2, 3, 1).float() * 255).round().to(torch.uint8)


# --------------------------------------------------------------------------------------------------------------
# PNDMScheduler (PLMS branch; skip_prk_steps=True, steps_offset=1, set_alpha_to_one=False)
# --------------------------------------------------------------------------------------------------------------
class PNDM:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n: int):
        self.n = n
        ratio = self.num_train_timesteps // n
        ts = (torch.arange(0, n) * ratio).round().long() + 1          # steps_offset = 1
        plms = torch.cat([ts[:-1], ts[-2:-1], ts[-1:]]).flip(0)       # [981, 961, 961, 941, ..., 1]
        self.timesteps = plms.tolist()
        self.ets: List[torch.Tensor] = []
        self.counter = 0
        self.cur_sample = None
        return self.timesteps

    def coeffs(self, t: int, prev_t: int) -> Tuple[float, float]:
        """_get_prev_sample: x' = c_sample * x - c_eps * eps'."""
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t, b_p = 1 - a_t, 1 - a_p
        c_sample = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return float(c_sample), float((a_p - a_t) / denom)

    def step(self, eps: torch.Tensor, t: int, sample: torch.Tensor) -> torch.Tensor:
        """step_plms."""
        ratio = self.num_train_timesteps // self.n
        prev_t = t - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(eps)
        else:
            prev_t = t
            t = t + ratio
        if len(self.ets) == 1 and self.counter == 0:
            e = eps
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            e = (eps + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            e = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            e = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            e = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        cs, ce = self.coeffs(t, prev_t)
        self.counter += 1
        return cs * sample - ce * e


def plms_schedule(n: int = 50):
    """Host-side table: for each of the n+1 UNet evaluations -> (timestep, c_sample, c_eps, mode) where mode says
    which linear-multistep formula `step` applies. Used by both the oracle tests and the product's host code check."""
    s = PNDM()
    ts = s.set_timesteps(n)
    ratio = s.num_train_timesteps // n
    out = []
    n_ets, counter = 0, 0
    for t in ts:
        prev_t, tt = t - ratio, t
        if counter != 1:
            n_ets = min(n_ets + 1, 4)
        else:
            prev_t, tt = t, t + ratio
        if n_ets == 1 and counter == 0:
            mode = 0
        elif n_ets == 1 and counter == 1:
            mode = 1
        else:
            mode = n_ets  # 2, 3, 4
        cs, ce = s.coeffs(tt, prev_t)
        out.append((t, cs, ce, mode))
        counter += 1
    return out


def denoise_loop(unet_sd: SD, prompt_embeds: torch.Tensor, negative_embeds: torch.Tensor, latents: torch.Tensor,
                 guidance_scale: float = 7.5, num_inference_steps: int = 50, cfg=None, return_all=False):
    """gill/custom_sd.py:606-651 (do_classifier_free_guidance == True)."""
    sched = PNDM()
    timesteps = sched.set_timesteps(num_inference_steps)               # custom_sd.py:607
    latents = latents * sched.init_noise_sigma                          # custom_sd.py:472
    neg = negative_embeds.expand(prompt_embeds.shape[0], -1, -1)        # custom_sd.py:351-357 (repeat per prompt)
    ctx = torch.cat([neg, prompt_embeds])                               # custom_sd.py:371
    trace = []
    for t in timesteps:                                                 # custom_sd.py:628
        inp = torch.cat([latents] * 2)                                  # :630 (scale_model_input is identity)
        eps = unet_forward(unet_sd, inp, t, ctx, cfg)                   # :633-638
        eu, et = eps.chunk(2)                                           # :642
        eps = eu + guidance_scale * (et - eu)                           # :643
        latents = sched.step(eps, t, latents)                           # :646
        if return_all:
            trace.append(latents.clone())
    return (latents, trace) if return_all else latents
