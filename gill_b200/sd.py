"""Stable Diffusion v1.5 image emission on the libgillb200 kernels: UNet denoising loop, PLMS scheduler, VAE decoder.

Drop-in for the object the reference calls at gill/models.py:730-731
    self.sd_pipe(prompt_embeds=gen_emb, generator=g, guidance_scale=7.5, num_inference_steps=50).images
whose control flow is restated by gill/custom_sd.py:567-666 (the arithmetic lives in diffusers==0.17.1).

B200-first layout decisions
  * activations are NHWC fp16 (the reference runs SD in fp16): a [B,H,W,C] tensor IS the [B*H*W, C] token matrix, so
    conv <-> transformer transitions are views, 1x1 convs are plain GEMMs and 3x3 convs are implicit GEMMs whose taps
    are TMA box shifts (no im2col buffer);
  * attention heads are stored at a pitch of 48/96/176 columns (head dims 40/80/160; zero weight rows, folded into the
    projection weights at load time); the attention kernels still fetch full 128-byte swizzled rows per TMA box and
    simply run fewer K-steps / a narrower P.V;
  * GEGLU value/gate rows are interleaved so the gate is applied in the GEMM epilogue;
  * everything that does not depend on the latent is hoisted out of the 51-step loop: the time-embedding MLP and all
    22 resnet time projections (a [51, sum_c] table built once), the 16 cross-attention K/V pairs (once per prompt),
    the PLMS coefficients (host scalars);
  * CFG + PLMS is one fused elementwise kernel; the VAE epilogue writes uint8 NHWC on the device.
"""
import math
import os
from typing import Dict, List, Optional

import torch

from . import ops

SD = Dict[str, torch.Tensor]
# conv_out (320 -> 4 in the UNet, 128 -> 3 in the VAE) as a plain GEMM onto per-tap products + shifted sum; "0": the implicit
# 3x3 conv through a 32-column tile (A/B aid)
NARROW_CONV_OUT = os.environ.get("GILLB200_NARROW_CONV_OUT", "1") != "0"
# up blocks: `conv3x3(upsample2x(h))` as four 2 x 2 phase convolutions of the low-res tensor (ops.conv3x3_up2); "0": upsample
# kernel + 3x3 conv on the 4x larger tensor (A/B aid)
FUSED_UPSAMPLE = os.environ.get("GILLB200_FUSED_UPSAMPLE", "1") != "0"


# resnets: GroupNorm + SiLU applied inside the consuming 3x3 conv (ops.conv3x3(..., gn=...)) instead of a separate
# elementwise pass over the tensor (four transform warps rewrite each halo tile in place; the next tile is requested one
# whole channel block ahead). Same box: UNet evaluation 17.76 -> 17.43 ms, VAE decode 19.7 -> 18.3 ms
# (profiles/r02_gn_fused_conv.log). "0": separate GroupNorm kernel (A/B aid)
FUSED_GN = os.environ.get("GILLB200_FUSED_GN", "1") != "0"


def _gn_silu_conv(x, w, pn, pc, G, eps, x2=None, **conv_kw):
    """norm -> SiLU -> conv3x3 of a resnet (x2: second source of a channel concat)."""
    wt = w[pc + ".weight"]
    if FUSED_GN and ops.conv3x3_gn_supported(x, wt.shape[0], x2):
        ss = ops.groupnorm_scale_shift(x, w[pn + ".weight"], w[pn + ".bias"], G, eps, x2=x2)
        return ops.conv3x3(x, wt, bias=w[pc + ".bias"], gn=(ss, True), x2=x2, **conv_kw)
    n = ops.groupnorm(x, w[pn + ".weight"], w[pn + ".bias"], G, eps, silu=True, x2=x2)
    return ops.conv3x3(n, wt, bias=w[pc + ".bias"], **conv_kw)


def _upsample_conv(h, w, p, stats=True):
    """Upsample2D of diffusers (nearest 2x + 3x3 conv): fused phase form when the shape qualifies."""
    if FUSED_UPSAMPLE and (p + ".weight_up2") in w and ops.conv3x3_up2_supported(h):
        return ops.conv3x3_up2(h, w[p + ".weight_up2"], bias=w[p + ".bias"], stats=stats)
    return ops.conv3x3(ops.upsample2x(h), w[p + ".weight"], bias=w[p + ".bias"], stats=stats)

UNET_CFG = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                cross_attention_dim=768, heads=8, norm_groups=32, has_attn_down=(True, True, True, False),
                has_attn_up=(False, True, True, True))
VAE_CFG = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
               norm_groups=32, scaling_factor=0.18215)


def _hd_pad(hd: int) -> int:
    for p in (64, 128, 192):
        if hd <= p:
            return p
    raise ValueError(f"head dim {hd} > 192 not supported by the fused attention kernel")


def _hd_pitch(hd: int) -> int:
    """Column pitch of one head in the q/k/v/out projections. Head dims that fill a 64-column tile exactly keep it;
    the others (SD-1.5: 40 / 80 / 160) get the next multiple of 16 ABOVE hd, i.e. 48 / 96 / 176: at least one spare
    column per head (V's ones column: the softmax denominator comes out of the P.V product) while the projections, the
    attention K-steps and the P.V width shrink by 25 / 25 / 8 % against padding every head to 64 / 128 / 192."""
    if os.environ.get("GILLB200_HEAD_PITCH", "1") == "0":        # A/B aid: round-1 layout (pitch == tile width)
        return _hd_pad(hd)
    return hd if hd % 64 == 0 else (hd // 16 + 1) * 16


def _conv_w(w: torch.Tensor, dt) -> torch.Tensor:
    """[Co, Ci, 3, 3] -> [Co, 9*Ci], k = (ky*3+kx)*Ci + c."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(dt).contiguous()


def _pad_heads_rows(w: torch.Tensor, heads: int, hd: int, hp: int) -> torch.Tensor:
    """[heads*hd, K] -> [heads*hp, K] with zero rows in each head's padding."""
    out = torch.zeros((heads, hp, w.shape[1]), dtype=w.dtype, device=w.device)
    out[:, :hd] = w.view(heads, hd, -1)
    return out.view(heads * hp, -1)


def _pad_heads_cols(w: torch.Tensor, heads: int, hd: int, hp: int) -> torch.Tensor:
    """[N, heads*hd] -> [N, heads*hp] with zero columns in each head's padding."""
    out = torch.zeros((w.shape[0], heads, hp), dtype=w.dtype, device=w.device)
    out[:, :, :hd] = w.view(w.shape[0], heads, hd)
    return out.view(w.shape[0], heads * hp)


def plms_table(num_inference_steps: int = 50, num_train_timesteps: int = 1000, beta_start=0.00085, beta_end=0.012):
    """Host-side PNDM/PLMS schedule (diffusers PNDMScheduler with SD-1.5's scheduler_config: scaled_linear betas,
    skip_prk_steps, steps_offset 1, set_alpha_to_one False). Returns [(timestep, c_sample, c_eps, mode)] for the
    n+1 UNet evaluations: x' = c_sample * x - c_eps * e'. mode selects the multistep formula (see plms_step)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    final = ac[0]
    ratio = num_train_timesteps // num_inference_steps
    ts = (torch.arange(0, num_inference_steps) * ratio).round().long() + 1
    plms = torch.cat([ts[:-1], ts[-2:-1], ts[-1:]]).flip(0).tolist()
    out, n_ets, counter = [], 0, 0
    for t in plms:
        prev_t, tt = t - ratio, t
        if counter != 1:
            n_ets = min(n_ets + 1, 4)
        else:
            prev_t, tt = t, t + ratio
        mode = 0 if (n_ets == 1 and counter == 0) else 1 if (n_ets == 1 and counter == 1) else n_ets
        a_t = ac[tt]
        a_p = ac[prev_t] if prev_t >= 0 else final
        b_t, b_p = 1 - a_t, 1 - a_p
        cs = (a_p / a_t) ** 0.5
        ce = (a_p - a_t) / (a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5)
        out.append((int(t), float(cs), float(ce), mode))
        counter += 1
    return out


class _Out:
    def __init__(self, images):
        self.images = images
        self.nsfw_content_detected = None


# ================================================================================================================
class UNetB200:
    """SD-1.5 UNet2DConditionModel forward on B200 kernels. Weights: diffusers-named state dict (any float dtype)."""

    def __init__(self, sd: SD, cfg=None, device="cuda", dtype=torch.float16):
        self.cfg = cfg or UNET_CFG
        self.dev, self.dt = torch.device(device), dtype
        self.G = self.cfg["norm_groups"]
        self.heads = self.cfg["heads"]
        self.w: Dict[str, torch.Tensor] = {}
        self._temb_names: List[str] = []
        self._temb_off: Dict[str, tuple] = {}
        self._pack(sd)
        self._temb_table = None
        self._temb_steps = None
        # Optional (GILLB200_FOLD_LN=1): transformer-block LayerNorms folded into their consuming GEMMs. Off by default:
        # measured on one B200 box, graph-replayed B=16 evaluation, 21.49 ms folded vs 21.32 ms with the separate
        # LayerNorm kernels -- the K=320 consumers (qkv, GEGLU) are epilogue-bound, so the extra epilogue FMAs and the
        # producers' row sums cost more than the 0.88 ms of LayerNorm launches they remove.
        self.fold_ln = os.environ.get("GILLB200_FOLD_LN", "0") == "1"

    # ------------------------------------------------------------------------------------------------ packing
    def _put(self, k, v, dt=None):
        self.w[k] = v.to(self.dev, dt or self.dt).contiguous()

    def _pack_resnet(self, sd, p):
        f32 = torch.float32
        for n in ("norm1", "norm2"):
            self._put(f"{p}.{n}.weight", sd[f"{p}.{n}.weight"], f32)
            self._put(f"{p}.{n}.bias", sd[f"{p}.{n}.bias"], f32)
        for n in ("conv1", "conv2"):
            self._put(f"{p}.{n}.weight", _conv_w(sd[f"{p}.{n}.weight"], self.dt))
            self._put(f"{p}.{n}.bias", sd[f"{p}.{n}.bias"], f32)
        if f"{p}.conv_shortcut.weight" in sd:
            w = sd[f"{p}.conv_shortcut.weight"]
            self._put(f"{p}.conv_shortcut.weight", w.reshape(w.shape[0], w.shape[1]))
            self._put(f"{p}.conv_shortcut.bias", sd[f"{p}.conv_shortcut.bias"], f32)
        if f"{p}.time_emb_proj.weight" in sd:
            self._temb_names.append(p)

    def _pack_transformer(self, sd, p, c):
        f32 = torch.float32
        H = self.heads
        hd = c // H
        hp = _hd_pitch(hd)          # column pitch of a head in every packed projection below
        for n in ("norm",):
            self._put(f"{p}.{n}.weight", sd[f"{p}.{n}.weight"], f32)
            self._put(f"{p}.{n}.bias", sd[f"{p}.{n}.bias"], f32)
        for n in ("proj_in", "proj_out"):
            w = sd[f"{p}.{n}.weight"]
            self._put(f"{p}.{n}.weight", w.reshape(w.shape[0], w.shape[1]))
            self._put(f"{p}.{n}.bias", sd[f"{p}.{n}.bias"], f32)
        t = p + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            self._put(f"{t}.{n}.weight", sd[f"{t}.{n}.weight"], f32)
            self._put(f"{t}.{n}.bias", sd[f"{t}.{n}.bias"], f32)
        q, k, v = (sd[f"{t}.attn1.to_{x}.weight"].float() for x in "qkv")
        self._put(f"{t}.attn1.qkv", torch.cat([_pad_heads_rows(x, H, hd, hp) for x in (q, k, v)], 0))
        # V carries 1.0 in the first padding column of every head (a bias on zero weight rows): the attention kernel
        # then reads the softmax denominator out of the P.V accumulator instead of summing P on the CUDA cores
        ones = torch.zeros(H, hp)
        if hp > hd:
            ones[:, hd] = 1.0
        self._put(f"{t}.attn1.qkv_bias", torch.cat([torch.zeros(2 * H * hp), ones.reshape(-1)]), f32)
        self._put(f"{t}.attn2.kv_bias", torch.cat([torch.zeros(H * hp), ones.reshape(-1)]), f32)
        self._put(f"{t}.attn1.out.weight", _pad_heads_cols(sd[f"{t}.attn1.to_out.0.weight"].float(), H, hd, hp))
        self._put(f"{t}.attn1.out.bias", sd[f"{t}.attn1.to_out.0.bias"], f32)
        self._put(f"{t}.attn2.q", _pad_heads_rows(sd[f"{t}.attn2.to_q.weight"].float(), H, hd, hp))
        self._put(f"{t}.attn2.kv", torch.cat([_pad_heads_rows(sd[f"{t}.attn2.to_{x}.weight"].float(), H, hd, hp)
                                              for x in "kv"], 0))
        self._put(f"{t}.attn2.out.weight", _pad_heads_cols(sd[f"{t}.attn2.to_out.0.weight"].float(), H, hd, hp))
        self._put(f"{t}.attn2.out.bias", sd[f"{t}.attn2.to_out.0.bias"], f32)
        # GEGLU: interleave (value_j, gate_j) rows so that the gate is applied in the GEMM epilogue
        w, b = sd[f"{t}.ff.net.0.proj.weight"], sd[f"{t}.ff.net.0.proj.bias"]
        half = w.shape[0] // 2
        self._put(f"{t}.ff.geglu.weight", torch.stack([w[:half], w[half:]], 1).reshape(w.shape[0], w.shape[1]))
        self._put(f"{t}.ff.geglu.bias", torch.stack([b[:half], b[half:]], 1).reshape(-1), f32)
        self._put(f"{t}.ff.out.weight", sd[f"{t}.ff.net.2.weight"])
        self._put(f"{t}.ff.out.bias", sd[f"{t}.ff.net.2.bias"], f32)
        # LayerNorm folded into the consuming GEMM (norm1 -> qkv, norm2 -> cross-attention q, norm3 -> GEGLU):
        #   LN(x) W^T + b = rstd * (x W'^T - mean * colsum(W')) + (b + W beta),  W' = W diag(gamma)
        def fold(wkey, bias, nkey):
            W = self.w[wkey].float().cpu()                      # the packed (padded / interleaved) fp16 weight
            gam, bet = sd[f"{t}.{nkey}.weight"].float(), sd[f"{t}.{nkey}.bias"].float()
            Wf = (W * gam[None, :]).to(self.dt)
            self._put(wkey + "_ln", Wf)
            self._put(wkey + "_cs", Wf.float().sum(1), f32)     # column sums of what the tensor core multiplies
            b0 = torch.zeros(W.shape[0]) if bias is None else self.w[bias].float().cpu()
            self._put(wkey + "_lnb", b0 + W @ bet, f32)
        fold(f"{t}.attn1.qkv", f"{t}.attn1.qkv_bias", "norm1")
        fold(f"{t}.attn2.q", None, "norm2")
        fold(f"{t}.ff.geglu.weight", f"{t}.ff.geglu.bias", "norm3")
        self._tf_meta = getattr(self, "_tf_meta", {})
        self._tf_meta[p] = (c, hd, _hd_pad(hp), hp)

    def _pack(self, sd):
        cfg, f32 = self.cfg, torch.float32
        boc = cfg["block_out_channels"]
        self.temb_dim = boc[0] * 4
        for n in ("time_embedding.linear_1", "time_embedding.linear_2"):
            self._put(n + ".weight", sd[n + ".weight"], f32)
            self._put(n + ".bias", sd[n + ".bias"], f32)
        w = _conv_w(sd["conv_in.weight"], self.dt)  # [320, 36] -> zero-pad K to 64 for the im2col GEMM
        wp = torch.zeros((w.shape[0], 64), dtype=w.dtype)
        wp[:, : w.shape[1]] = w
        self._put("conv_in.weight", wp)
        self._put("conv_in.bias", sd["conv_in.bias"], f32)
        self.tf_layers: List[str] = []
        ch = boc[0]
        for i, c in enumerate(boc):
            for j in range(cfg["layers_per_block"]):
                self._pack_resnet(sd, f"down_blocks.{i}.resnets.{j}")
                ch = c
                if cfg["has_attn_down"][i]:
                    self._pack_transformer(sd, f"down_blocks.{i}.attentions.{j}", c)
                    self.tf_layers.append(f"down_blocks.{i}.attentions.{j}")
            if i < len(boc) - 1:
                p = f"down_blocks.{i}.downsamplers.0.conv"
                self._put(p + ".weight", _conv_w(sd[p + ".weight"], self.dt))
                self._put(p + ".bias", sd[p + ".bias"], f32)
        self._pack_resnet(sd, "mid_block.resnets.0")
        self._pack_transformer(sd, "mid_block.attentions.0", ch)
        self.tf_layers.append("mid_block.attentions.0")
        self._pack_resnet(sd, "mid_block.resnets.1")
        for i, c in enumerate(reversed(boc)):
            for j in range(cfg["layers_per_block"] + 1):
                self._pack_resnet(sd, f"up_blocks.{i}.resnets.{j}")
                if cfg["has_attn_up"][i]:
                    self._pack_transformer(sd, f"up_blocks.{i}.attentions.{j}", c)
                    self.tf_layers.append(f"up_blocks.{i}.attentions.{j}")
            if i < len(boc) - 1:
                p = f"up_blocks.{i}.upsamplers.0.conv"
                self._put(p + ".weight", _conv_w(sd[p + ".weight"], self.dt))
                self._put(p + ".weight_up2", ops.conv3x3_up2_weights(_conv_w(sd[p + ".weight"], torch.float32)))
                self._put(p + ".bias", sd[p + ".bias"], f32)
        self._put("conv_norm_out.weight", sd["conv_norm_out.weight"], f32)
        self._put("conv_norm_out.bias", sd["conv_norm_out.bias"], f32)
        self._put("conv_out.weight", _conv_w(sd["conv_out.weight"], self.dt))
        self._put("conv_out.bias", sd["conv_out.bias"], f32)
        # conv_out (C -> 4) as one plain GEMM onto the 36 per-tap products + a shifted sum (ops.conv3x3_narrow)
        self._put("conv_out.taps", ops.conv3x3_taps_weight(_conv_w(sd["conv_out.weight"], self.dt), sd["conv_out.weight"].shape[0]))
        # all resnet time projections as one [sum_c, 1280] matrix
        ws, bs, off = [], [], 0
        for p in self._temb_names:
            w = sd[p + ".time_emb_proj.weight"]
            ws.append(w)
            bs.append(sd[p + ".time_emb_proj.bias"])
            self._temb_off[p] = (off, off + w.shape[0])
            off += w.shape[0]
        self._put("temb_proj_all.weight", torch.cat(ws, 0), torch.float16)
        self._put("temb_proj_all.bias", torch.cat(bs, 0), f32)
        self.temb_total = off
        # fixed-address buffer holding the current step's time projections (captured graphs read it)
        self._temb_cur = torch.zeros((1, off), device=self.dev, dtype=torch.float32)

    # ------------------------------------------------------------------------------------------------ hoisted work
    def prepare_timesteps(self, timesteps: List[int]):
        """Time-embedding MLP + every resnet's time projection for all timesteps of the schedule: [n_steps, sum_c].
        Depends only on the weights and the schedule, so it is built once (cached per schedule)."""
        key = tuple(timesteps)
        if self._temb_steps == key:
            return
        boc0 = self.cfg["block_out_channels"][0]
        half = boc0 // 2
        t = torch.tensor(timesteps, dtype=torch.float32, device=self.dev)
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=self.dev) / half)
        args = t[:, None] * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)  # flip_sin_to_cos
        n = len(timesteps)
        # tiny [n,320]x[320,1280] MLP on the tensor cores with fp16 hi+lo operands (~22 mantissa bits)
        def split(x):
            hi = x.to(torch.float16)
            return hi, (x - hi.float()).to(torch.float16)
        w1, w2 = self.w["time_embedding.linear_1.weight"], self.w["time_embedding.linear_2.weight"]
        w1h, w1l = split(w1)
        w2h, w2l = split(w2)
        eh, el = split(emb)
        h = ops.gemm(eh, w1h, a2=el, a2_mode=2, bias=self.w["time_embedding.linear_1.bias"], out_dtype=torch.float32)
        h = h + ops.gemm(eh, w1l, out_dtype=torch.float32)
        h = torch.nn.functional.silu(h)
        hh, hl = split(h)
        te = ops.gemm(hh, w2h, a2=hl, a2_mode=2, bias=self.w["time_embedding.linear_2.bias"], out_dtype=torch.float32)
        te = te + ops.gemm(hh, w2l, out_dtype=torch.float32)
        act = torch.nn.functional.silu(te)
        ah, al = split(act)
        self._temb_table = ops.gemm(ah, self.w["temb_proj_all.weight"], a2=al, a2_mode=2,
                                    bias=self.w["temb_proj_all.bias"], out_dtype=torch.float32)  # [n, sum_c]
        self._temb_steps = key

    def set_step(self, step: int):
        """Select the schedule entry the next forward() uses: copies that row of the time-projection table into the
        fixed buffer every resnet reads its per-channel bias from (fixed address => the forward is graph-replayable)."""
        ops.cast_add(self._temb_table[step], None, torch.float32, out=self._temb_cur)

    def precompute_ctx(self, ctx: torch.Tensor, out: Optional[Dict[str, torch.Tensor]] = None):
        """Cross-attention K/V of all transformer layers for a batch of conditionings [B2,77,768]: once per prompt.
        `out` (from a previous call with the same batch size) is overwritten in place, keeping addresses stable."""
        B2, L, D = ctx.shape
        c2 = ctx.to(self.dt).reshape(B2 * L, D).contiguous()
        res = {} if out is None else out
        for p in self.tf_layers:
            w = self.w[f"{p}.transformer_blocks.0.attn2.kv"]
            bias = self.w[f"{p}.transformer_blocks.0.attn2.kv_bias"]
            if out is None:
                res[p] = ops.gemm(c2, w, bias=bias).view(B2, L, -1)
            else:
                ops.gemm(c2, w, bias=bias, out=out[p].view(B2 * L, -1))
        return res

    # ------------------------------------------------------------------------------------------------ blocks
    def _resnet(self, x, p, x2=None):
        w, G = self.w, self.G
        B, H, W, _ = x.shape
        lo, hi = self._temb_off[p]
        rb = self._temb_cur[:, lo:hi]
        h = _gn_silu_conv(x, w, p + ".norm1", p + ".conv1", G, 1e-5, x2=x2, rowbias=rb, stats=True)
        if (p + ".conv_shortcut.weight") in w:
            M = B * H * W
            if x2 is not None:
                sc = ops.gemm(x.view(M, -1), w[p + ".conv_shortcut.weight"], a2=x2.view(M, -1), a2_mode=1,
                              bias=w[p + ".conv_shortcut.bias"])
            else:
                sc = ops.gemm(x.view(M, -1), w[p + ".conv_shortcut.weight"], bias=w[p + ".conv_shortcut.bias"])
            sc = sc.view(B, H, W, -1)
        else:
            assert x2 is None
            sc = x
        return _gn_silu_conv(h, w, p + ".norm2", p + ".conv2", G, 1e-5, residual=sc, stats=True)

    def _transformer(self, x, p, ctx_kv):
        w, G, Hh = self.w, self.G, self.heads
        B, H, W, C = x.shape
        _, hd, tile, hp = self._tf_meta[p]      # hp = head pitch (columns per head), tile = kernel tile width
        M, L = B * H * W, H * W
        t = p + ".transformer_blocks.0"
        n = ops.groupnorm(x, w[p + ".norm.weight"], w[p + ".norm.bias"], G, 1e-6)
        if self.fold_ln:
            return self._transformer_folded(x, n, p, ctx_kv)
        h = ops.gemm(n.view(M, C), w[p + ".proj_in.weight"], bias=w[p + ".proj_in.bias"])
        # self attention
        n1 = ops.layernorm(h, w[t + ".norm1.weight"], w[t + ".norm1.bias"], 1e-5)
        # V carries 1.0 in the first padding column of each head, so the P.V accumulator's column `hd` IS the softmax
        # denominator: the kernel skips the 64 row-sum FADDs per KV tile (fp32 exp2 kept; the packed f16x2 ex2 variant
        # measured slower on B200)
        oc = hd if hp > hd else 0
        qkv = ops.gemm(n1, w[t + ".attn1.qkv"], bias=w[t + ".attn1.qkv_bias"]).view(B, L, 3 * Hh * hp)
        a = ops.attention(qkv[:, :, : Hh * hp], qkv[:, :, Hh * hp : 2 * Hh * hp], qkv[:, :, 2 * Hh * hp :], Hh, tile,
                          hd ** -0.5, ones_col=oc, head_dim=hd, head_stride=hp)
        h = ops.gemm(a.view(M, Hh * hp), w[t + ".attn1.out.weight"], bias=w[t + ".attn1.out.bias"], residual=h)
        # cross attention against the precomputed K/V of the conditioning
        n2 = ops.layernorm(h, w[t + ".norm2.weight"], w[t + ".norm2.bias"], 1e-5)
        q = ops.gemm(n2, w[t + ".attn2.q"]).view(B, L, Hh * hp)
        kv = ctx_kv[p]
        a = ops.attention(q, kv[:, :, : Hh * hp], kv[:, :, Hh * hp :], Hh, tile, hd ** -0.5, ones_col=oc, head_dim=hd,
                          head_stride=hp)
        h = ops.gemm(a.view(M, Hh * hp), w[t + ".attn2.out.weight"], bias=w[t + ".attn2.out.bias"], residual=h)
        # GEGLU feed-forward
        n3 = ops.layernorm(h, w[t + ".norm3.weight"], w[t + ".norm3.bias"], 1e-5)
        g = ops.gemm(n3, w[t + ".ff.geglu.weight"], bias=w[t + ".ff.geglu.bias"], act="geglu")
        h = ops.gemm(g, w[t + ".ff.out.weight"], bias=w[t + ".ff.out.bias"], residual=h)
        out = ops.gemm(h, w[p + ".proj_out.weight"], bias=w[p + ".proj_out.bias"], residual=x.view(M, C), stats=True)
        o4 = out.view(B, H, W, C)
        o4.gn_stats = out.gn_stats      # the next resnet's GroupNorm takes its statistics from this epilogue
        return o4

    def _transformer_folded(self, x, n, p, ctx_kv):
        """Same block with the three LayerNorms folded into their consuming GEMMs: the producers of the residual stream
        leave per-row partial sums (rowstats), the consumers correct `x W'^T` with (mean, rstd) in their epilogues -- no
        LayerNorm kernel, no normalised copy of the stream."""
        w, Hh = self.w, self.heads
        B, H, W, C = x.shape
        _, hd, tile, hp = self._tf_meta[p]
        M, L = B * H * W, H * W
        t = p + ".transformer_blocks.0"
        oc = hd if hp > hd else 0
        h = ops.gemm(n.view(M, C), w[p + ".proj_in.weight"], bias=w[p + ".proj_in.bias"], rowstats=True)
        qkv = ops.gemm(h, w[t + ".attn1.qkv_ln"], bias=w[t + ".attn1.qkv_lnb"],
                       ln=(h.ln_stats, w[t + ".attn1.qkv_cs"], 1e-5)).view(B, L, 3 * Hh * hp)
        a = ops.attention(qkv[:, :, : Hh * hp], qkv[:, :, Hh * hp : 2 * Hh * hp], qkv[:, :, 2 * Hh * hp :], Hh, tile,
                          hd ** -0.5, ones_col=oc, head_dim=hd, head_stride=hp)
        h = ops.gemm(a.view(M, Hh * hp), w[t + ".attn1.out.weight"], bias=w[t + ".attn1.out.bias"], residual=h, rowstats=True)
        q = ops.gemm(h, w[t + ".attn2.q_ln"], bias=w[t + ".attn2.q_lnb"],
                     ln=(h.ln_stats, w[t + ".attn2.q_cs"], 1e-5)).view(B, L, Hh * hp)
        kv = ctx_kv[p]
        a = ops.attention(q, kv[:, :, : Hh * hp], kv[:, :, Hh * hp :], Hh, tile, hd ** -0.5, ones_col=oc, head_dim=hd,
                          head_stride=hp)
        h = ops.gemm(a.view(M, Hh * hp), w[t + ".attn2.out.weight"], bias=w[t + ".attn2.out.bias"], residual=h, rowstats=True)
        g = ops.gemm(h, w[t + ".ff.geglu.weight_ln"], bias=w[t + ".ff.geglu.weight_lnb"], act="geglu",
                     ln=(h.ln_stats, w[t + ".ff.geglu.weight_cs"], 1e-5))
        h = ops.gemm(g, w[t + ".ff.out.weight"], bias=w[t + ".ff.out.bias"], residual=h)
        out = ops.gemm(h, w[p + ".proj_out.weight"], bias=w[p + ".proj_out.bias"], residual=x.view(M, C), stats=True)
        o4 = out.view(B, H, W, C)
        o4.gn_stats = out.gn_stats
        return o4

    def _conv_s2(self, x, p):
        # stride-2 downsampler as an implicit GEMM: the TMA tensor map walks the input with element strides {1,2,2,1}
        # (no [M, 9C] im2col buffer: 94 MB at the 64x64 level)
        return ops.conv3x3(x, self.w[p + ".weight"], bias=self.w[p + ".bias"], stride=2, stats=True)

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, step: Optional[int], ctx_kv: Dict[str, torch.Tensor]) -> torch.Tensor:
        """x: NHWC [B2,H,W,4] fp16 latent pair; step: index into the prepared timestep table (None: keep the row
        selected by the last set_step, as graph replays do); returns eps NHWC."""
        cfg, w = self.cfg, self.w
        if step is not None:
            self.set_step(step)
        boc = cfg["block_out_channels"]
        B, H, W, _ = x.shape
        cols = ops.im2col3x3(x, 1, ld_out=64)
        h2 = ops.gemm(cols, w["conv_in.weight"], bias=w["conv_in.bias"], stats=True)
        h = h2.view(B, H, W, boc[0])
        h.gn_stats = h2.gn_stats
        skips = [h]
        for i in range(len(boc)):
            for j in range(cfg["layers_per_block"]):
                h = self._resnet(h, f"down_blocks.{i}.resnets.{j}")
                if cfg["has_attn_down"][i]:
                    h = self._transformer(h, f"down_blocks.{i}.attentions.{j}", ctx_kv)
                skips.append(h)
            if i < len(boc) - 1:
                h = self._conv_s2(h, f"down_blocks.{i}.downsamplers.0.conv")
                skips.append(h)
        h = self._resnet(h, "mid_block.resnets.0")
        h = self._transformer(h, "mid_block.attentions.0", ctx_kv)
        h = self._resnet(h, "mid_block.resnets.1")
        for i in range(len(boc)):
            for j in range(cfg["layers_per_block"] + 1):
                h = self._resnet(h, f"up_blocks.{i}.resnets.{j}", x2=skips.pop())
                if cfg["has_attn_up"][i]:
                    h = self._transformer(h, f"up_blocks.{i}.attentions.{j}", ctx_kv)
            if i < len(boc) - 1:
                h = _upsample_conv(h, w, f"up_blocks.{i}.upsamplers.0.conv")
        n = ops.groupnorm(h, w["conv_norm_out.weight"], w["conv_norm_out.bias"], self.G, 1e-5, silu=True)
        if NARROW_CONV_OUT:
            return ops.conv3x3_narrow(n, w["conv_out.taps"], w["conv_out.bias"].numel(), bias=w["conv_out.bias"])
        return ops.conv3x3(n, w["conv_out.weight"], bias=w["conv_out.bias"], block_n=32)


# ================================================================================================================
class VAEDecoderB200:
    """AutoencoderKL.decode + the uint8 epilogue of gill/custom_sd.py:385-392 on B200 kernels (NHWC fp16)."""

    def __init__(self, sd: SD, cfg=None, device="cuda", dtype=torch.float16):
        self.cfg = cfg or VAE_CFG
        self.dev, self.dt = torch.device(device), dtype
        self.G = self.cfg["norm_groups"]
        self.w: Dict[str, torch.Tensor] = {}
        f32 = torch.float32
        for k, v in sd.items():
            if k.endswith("conv_shortcut.weight") or k == "post_quant_conv.weight":
                self.w[k] = v.reshape(v.shape[0], v.shape[1]).to(self.dev, dtype).contiguous()
            elif v.dim() == 4:
                self.w[k] = _conv_w(v, dtype).to(self.dev)
                if "upsamplers" in k and k.endswith("conv.weight"):   # pre-summed phase kernels of the fused upsample conv
                    self.w[k[: -len("weight")] + "weight_up2"] = \
                        ops.conv3x3_up2_weights(_conv_w(v, f32).to(self.dev)).to(dtype).contiguous()
            elif v.dim() == 2:
                self.w[k] = v.to(self.dev, dtype).contiguous()
            else:
                self.w[k] = v.to(self.dev, f32).contiguous()
        # fold latents / scaling_factor and the 1x1 post_quant_conv into one tiny 4x4 map applied before conv_in
        lc = self.cfg["latent_channels"]
        pq = sd["post_quant_conv.weight"].reshape(lc, lc).float() / self.cfg["scaling_factor"]
        self.pq_w = pq.contiguous().to(self.dev)
        self.pq_b = sd["post_quant_conv.bias"].float().to(self.dev)
        w = self.w["decoder.conv_in.weight"]  # [512, 36] -> K padded to 64
        wp = torch.zeros((w.shape[0], 64), dtype=w.dtype, device=self.dev)
        wp[:, : w.shape[1]] = w
        self.w["decoder.conv_in.weight"] = wp
        w = self.w["decoder.conv_out.weight"]  # [3, 9*128] -> 8 rows so the epilogue can use vector stores
        wp = torch.zeros((8, w.shape[1]), dtype=w.dtype, device=self.dev)
        wp[: w.shape[0]] = w
        self.w["decoder.conv_out.weight"] = wp
        self.w["decoder.conv_out.taps"] = ops.conv3x3_taps_weight(w, self.cfg["out_channels"])
        self.w["decoder.conv_out.bias3"] = self.w["decoder.conv_out.bias"].clone()
        b = torch.zeros(8, dtype=f32, device=self.dev)
        b[: self.cfg["out_channels"]] = self.w["decoder.conv_out.bias"]
        self.w["decoder.conv_out.bias"] = b

    def _resnet(self, x, p):
        w, G = self.w, self.G
        B, H, W, _ = x.shape
        h = _gn_silu_conv(x, w, p + ".norm1", p + ".conv1", G, 1e-6, stats=True)
        if (p + ".conv_shortcut.weight") in w:
            sc = ops.gemm(x.view(B * H * W, -1), w[p + ".conv_shortcut.weight"],
                          bias=w[p + ".conv_shortcut.bias"]).view(B, H, W, -1)
        else:
            sc = x
        return _gn_silu_conv(h, w, p + ".norm2", p + ".conv2", G, 1e-6, residual=sc, stats=True)

    def _mid_attention(self, x):
        """Single-head attention over H*W tokens with head dim 512: too wide for the fused kernel's smem tiles, and
        it runs once per image, so it is three GEMMs around a row softmax (scores kept fp32)."""
        w, a = self.w, "decoder.mid_block.attentions.0"
        B, H, W, C = x.shape
        L = H * W
        n = ops.groupnorm(x, w[a + ".group_norm.weight"], w[a + ".group_norm.bias"], self.G, 1e-6).view(B * L, C)
        q = ops.gemm(n, w[a + ".to_q.weight"], bias=w[a + ".to_q.bias"])
        k = ops.gemm(n, w[a + ".to_k.weight"], bias=w[a + ".to_k.bias"])
        o = torch.empty((B * L, C), device=x.device, dtype=x.dtype)
        for b in range(B):
            nb = n[b * L : (b + 1) * L]
            vT = ops.gemm(w[a + ".to_v.weight"], nb, bias=w[a + ".to_v.bias"], bias_along_m=True)   # [C, L] = V^T
            s = ops.gemm(q[b * L : (b + 1) * L], k[b * L : (b + 1) * L], out_dtype=torch.float32)   # [L, L]
            pr = ops.softmax_rows(s, C ** -0.5, x.dtype)
            ops.gemm(pr, vT, out=o[b * L : (b + 1) * L])
        out = ops.gemm(o, w[a + ".to_out.0.weight"], bias=w[a + ".to_out.0.bias"], residual=x.view(B * L, C), stats=True)
        o4 = out.view(B, H, W, C)
        o4.gn_stats = out.gn_stats
        return o4

    def decode_u8(self, latents: torch.Tensor) -> torch.Tensor:
        """latents NHWC [B,h,w,4] (fp32 or fp16) -> uint8 NHWC [B,8h,8w,3]."""
        cfg, w = self.cfg, self.w
        boc = cfg["block_out_channels"]
        B, H, W, lc = latents.shape
        # latents/0.18215 -> post_quant_conv: one folded 4x4 channel map
        z = ops.channel_mix(latents.float().contiguous(), self.pq_w, self.pq_b, self.dt)
        cols = ops.im2col3x3(z, 1, ld_out=64)
        h2 = ops.gemm(cols, w["decoder.conv_in.weight"], bias=w["decoder.conv_in.bias"], stats=True)
        h = h2.view(B, H, W, boc[-1])
        h.gn_stats = h2.gn_stats
        h = self._resnet(h, "decoder.mid_block.resnets.0")
        h = self._mid_attention(h)
        h = self._resnet(h, "decoder.mid_block.resnets.1")
        for i in range(len(boc)):
            for j in range(cfg["layers_per_block"] + 1):
                h = self._resnet(h, f"decoder.up_blocks.{i}.resnets.{j}")
            if i < len(boc) - 1:
                h = _upsample_conv(h, w, f"decoder.up_blocks.{i}.upsamplers.0.conv")
        n = ops.groupnorm(h, w["decoder.conv_norm_out.weight"], w["decoder.conv_norm_out.bias"], self.G, 1e-6,
                          silu=True)
        if NARROW_CONV_OUT:
            img = ops.conv3x3_narrow(n, w["decoder.conv_out.taps"], cfg["out_channels"], bias=w["decoder.conv_out.bias3"])
        else:
            img = ops.conv3x3(n, w["decoder.conv_out.weight"], bias=w["decoder.conv_out.bias"], block_n=32)
        return ops.image_to_u8(img, cfg["out_channels"])


# ================================================================================================================
class StableDiffusionB200:
    """Callable with the `sd_pipe` signature GILL uses (gill/models.py:730; gill/custom_sd.py:477-496)."""

    def __init__(self, unet: UNetB200, vae: VAEDecoderB200, negative_prompt_embeds: torch.Tensor, safety_checker=None):
        """negative_prompt_embeds: the (77,768) CLIP-text embedding of "" that gill/custom_sd.py:319-357 recomputes on
        every call; it is a constant of the pipeline, so it is supplied once here."""
        self.unet, self.vae = unet, vae
        self.safety_checker = safety_checker   # optional gill_b200.clip.SafetyCheckerB200 (custom_sd.py:375-383)
        self.device = unet.dev
        self.neg = negative_prompt_embeds.to(unet.dev, unet.dt).reshape(1, 77, -1)
        self.latent_hw = 64
        self._graphs = {}
        self.use_graph = True      # replay one captured UNet evaluation per step instead of ~650 launches
        self.graph_replays = 0     # bookkeeping for bench.py's gpu_launches

    def _state(self, b: int, h: int, w: int):
        """Static per-(batch, size) buffers: every pointer the denoising loop touches stays fixed across calls, so one
        captured UNet evaluation can be replayed for all 51 steps of every call."""
        key = (b, h, w)
        st = self._graphs.get(key)
        if st is None:
            dev, n = self.device, b * h * w * 4
            st = dict(lat=torch.empty((b, h, w, 4), device=dev, dtype=torch.float32),
                      pair=torch.empty((2 * b, h, w, 4), device=dev, dtype=self.unet.dt),
                      ets=torch.empty((4, n), device=dev, dtype=torch.float32),
                      cur=torch.empty(n, device=dev, dtype=torch.float32),
                      ctx=torch.empty((2 * b, 77, self.neg.shape[-1]), device=dev, dtype=self.unet.dt),
                      ctx_kv=None, graph=None, eps=None, launches_per_eval=0)
            self._graphs[key] = st
        return st

    @torch.no_grad()
    def denoise(self, prompt_embeds: torch.Tensor, latents_nchw: torch.Tensor, guidance_scale: float = 7.5,
                num_inference_steps: int = 50, trace: Optional[list] = None, use_graph: Optional[bool] = None):
        """gill/custom_sd.py:606-651. Returns the final latents, NHWC fp32 [b,h,w,4] (a static buffer)."""
        b, _, h, w = latents_nchw.shape
        use_graph = self.use_graph if use_graph is None else use_graph
        table = plms_table(num_inference_steps)
        self.unet.prepare_timesteps([t for t, _, _, _ in table])
        st = self._state(b, h, w)
        st["ctx"][:b].copy_(self.neg.expand(b, -1, -1))                                         # custom_sd.py:371
        st["ctx"][b:].copy_(prompt_embeds.to(self.unet.dt))
        st["ctx_kv"] = self.unet.precompute_ctx(st["ctx"], st["ctx_kv"])
        lat, pair, ets, cur = st["lat"], st["pair"], st["ets"], st["cur"]
        lat.copy_(latents_nchw.float().permute(0, 2, 3, 1))                                     # NHWC fp32
        pair[:b].copy_(lat)                                                                     # custom_sd.py:630
        pair[b:].copy_(lat)
        if use_graph and st["graph"] is None:
            from ._lib import lib
            self.unet.set_step(0)
            st["eps"] = self.unet.forward(pair, None, st["ctx_kv"])                             # eager warm-up
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = lib().gillb200_launch_count()
            with ops.graph_capture(g, self.device):
                st["eps"] = self.unet.forward(pair, None, st["ctx_kv"])
            st["launches_per_eval"] = lib().gillb200_launch_count() - n0
            st["graph"] = g
        head = 0
        for i, (t, cs, ce, mode) in enumerate(table):                                           # custom_sd.py:628
            if use_graph:
                self.unet.set_step(i)
                st["graph"].replay()                                                            # :633-638
                self.graph_replays += 1
                eps = st["eps"]
            else:
                eps = self.unet.forward(pair, i, st["ctx_kv"])
            ops.plms_step(eps, guidance_scale, ets, head, mode, cs, ce, lat, cur, pair)         # :641-646
            if mode != 1:
                head = (head + 1) & 3
            if trace is not None:
                trace.append(lat.clone())
        return lat

    @torch.no_grad()
    def __call__(self, prompt=None, prompt_embeds: Optional[torch.Tensor] = None, generator=None,
                 guidance_scale: float = 7.5, num_inference_steps: int = 50, latents: Optional[torch.Tensor] = None,
                 output_type: str = "pil", height: int = 512, width: int = 512, **unused):
        if prompt_embeds is None:
            raise ValueError("gill_b200 StableDiffusion takes `prompt_embeds` (GILL never passes a text prompt; "
                             "gill/models.py:730)")
        if prompt_embeds.dim() != 3 or prompt_embeds.shape[1] != 77:
            raise ValueError(f"prompt_embeds must be (b, 77, 768), got {tuple(prompt_embeds.shape)}")
        if height % 8 != 0 or width % 8 != 0:                                                    # custom_sd.py:424
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        b = prompt_embeds.shape[0]
        shape = (b, 4, height // 8, width // 8)
        if latents is None:                                                                      # custom_sd.py:466-470
            latents = torch.randn(shape, generator=generator, device=self.device, dtype=torch.float16)
        elif tuple(latents.shape[:2]) != shape[:2] or latents.dim() != 4:
            raise ValueError(f"Unexpected latents shape, got {tuple(latents.shape)}, expected {shape}")
        lat = self.denoise(prompt_embeds.to(self.device), latents.to(self.device), guidance_scale, num_inference_steps)
        u8 = self.vae.decode_u8(lat)                                                            # custom_sd.py:654
        nsfw = None
        if self.safety_checker is not None:                                                      # custom_sd.py:657
            nsfw = self.safety_checker(u8)
            for i, bad in enumerate(nsfw):
                if bad:
                    u8[i].zero_()                                                                # black image
        if output_type == "uint8":
            o = _Out(u8)
            o.nsfw_content_detected = nsfw
            return o
        arr = u8.cpu().numpy()                                                                  # custom_sd.py:391
        if output_type == "np":
            o = _Out(arr.astype("float32") / 255.0)
            o.nsfw_content_detected = nsfw
            return o
        from PIL import Image

        o = _Out([Image.fromarray(a) for a in arr])                                              # custom_sd.py:661
        o.nsfw_content_detected = nsfw
        return o
