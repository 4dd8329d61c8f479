"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of GILL's `TextFcLayer.forward` in `gill_mapper` and `linear` modes
(reference: gill/layers.py:28-53; constructor gill/layers.py:8-26). The transformer is torch.nn.Transformer
(batch_first, norm_first, d_model 512, 4 enc + 4 dec layers, 4 heads, ff 2048, ReLU, dropout 0, final LayerNorm on both
stacks, no masks) restated functionally from a plain state_dict so that tests do not depend on nn.Transformer internals.

Pinned by: tests/golden/mapper_*.npz, produced by running the reference's own class (oracle/make_golden.py).
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

D_MODEL, N_HEAD, N_LAYERS = 512, 4, 4


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _mha(q_in, kv_in, sd, name):
    """nn.MultiheadAttention forward (packed in_proj q|k|v, scale head_dim^-0.5, no mask)."""
    w, b = sd[name + ".in_proj_weight"], sd[name + ".in_proj_bias"]
    d = w.shape[1]
    q = F.linear(q_in, w[:d], b[:d])
    k = F.linear(kv_in, w[d : 2 * d], b[d : 2 * d])
    v = F.linear(kv_in, w[2 * d :], b[2 * d :])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    hd = d // N_HEAD
    q = q.view(B, Lq, N_HEAD, hd).transpose(1, 2)
    k = k.view(B, Lk, N_HEAD, hd).transpose(1, 2)
    v = v.view(B, Lk, N_HEAD, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, sd[name + ".out_proj.weight"], sd[name + ".out_proj.bias"])


def _ffn(x, sd, pre):
    return F.linear(F.relu(F.linear(x, sd[pre + ".linear1.weight"], sd[pre + ".linear1.bias"])),
                    sd[pre + ".linear2.weight"], sd[pre + ".linear2.bias"])


def mapper_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, input_embs: torch.Tensor) -> torch.Tensor:
    """gill/layers.py:31-44 for mode == 'gill_mapper'. sd: TextFcLayer.state_dict() (any float dtype, same as x)."""
    x = x + input_embs                                             # layers.py:32
    x = F.linear(x, sd["fc.weight"], sd["fc.bias"])               # layers.py:42
    # encoder (norm_first): x += sa(LN1 x); x += ff(LN2 x); final LN
    for i in range(N_LAYERS):
        p = f"tfm.encoder.layers.{i}"
        h = _ln(x, sd, p + ".norm1")
        x = x + _mha(h, h, sd, p + ".self_attn")
        x = x + _ffn(_ln(x, sd, p + ".norm2"), sd, p)
    mem = _ln(x, sd, "tfm.encoder.norm")
    # decoder: y += sa(LN1 y); y += ca(LN2 y, mem); y += ff(LN3 y); final LN
    y = sd["query_embs"].expand(x.shape[0], -1, -1)               # layers.py:43
    for i in range(N_LAYERS):
        p = f"tfm.decoder.layers.{i}"
        h = _ln(y, sd, p + ".norm1")
        y = y + _mha(h, h, sd, p + ".self_attn")
        y = y + _mha(_ln(y, sd, p + ".norm2"), mem, sd, p + ".multihead_attn")
        y = y + _ffn(_ln(y, sd, p + ".norm3"), sd, p)
    y = _ln(y, sd, "tfm.decoder.norm")
    out = F.linear(y, sd["model.weight"], sd["model.bias"])       # layers.py:44
    assert out.shape[1] * out.shape[2] == 77 * 768                 # layers.py:52
    return out


def linear_head_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_output_tokens: int = 1) -> torch.Tensor:
    """gill/layers.py:44-48 for mode == 'linear' (the retrieval head): Linear on every token, keep the first n."""
    out = F.linear(x, sd["model.weight"], sd["model.bias"])
    if out.shape[1] != num_output_tokens:
        out = out[:, :num_output_tokens, :]
    return out


def synthetic_mapper_state_dict(seed: int = 1234, in_dim: int = 4096, out_dim: int = 768, n_query: int = 77,
                                dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded stand-in for the shipped checkpoint (same names/shapes as `gen_text_hidden_fcs.0.*`,
    checkpoints/gill_opt/pretrained_ckpt.pth.tar), values rounded to bf16 like the real weights. Used on machines
    that do not have the checkpoint (the GPU box); reproducible across machines (CPU generator)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, scale):
        return (torch.randn(*shape, generator=g) * scale).bfloat16().to(dtype)

    sd = {"fc.weight": rnd(D_MODEL, in_dim, scale=in_dim ** -0.5), "fc.bias": rnd(D_MODEL, scale=0.02),
          "model.weight": rnd(out_dim, D_MODEL, scale=D_MODEL ** -0.5), "model.bias": rnd(out_dim, scale=0.02),
          "query_embs": rnd(1, n_query, D_MODEL, scale=1.0)}

    def attn(p):
        sd[p + ".in_proj_weight"] = rnd(3 * D_MODEL, D_MODEL, scale=D_MODEL ** -0.5)
        sd[p + ".in_proj_bias"] = rnd(3 * D_MODEL, scale=0.02)
        sd[p + ".out_proj.weight"] = rnd(D_MODEL, D_MODEL, scale=D_MODEL ** -0.5)
        sd[p + ".out_proj.bias"] = rnd(D_MODEL, scale=0.02)

    def norm(p):
        sd[p + ".weight"] = (1.0 + 0.05 * torch.randn(D_MODEL, generator=g)).bfloat16().to(dtype)
        sd[p + ".bias"] = rnd(D_MODEL, scale=0.02)

    def ffn(p):
        sd[p + ".linear1.weight"] = rnd(4 * D_MODEL, D_MODEL, scale=D_MODEL ** -0.5)
        sd[p + ".linear1.bias"] = rnd(4 * D_MODEL, scale=0.02)
        sd[p + ".linear2.weight"] = rnd(D_MODEL, 4 * D_MODEL, scale=(4 * D_MODEL) ** -0.5)
        sd[p + ".linear2.bias"] = rnd(D_MODEL, scale=0.02)

    for i in range(N_LAYERS):
        p = f"tfm.encoder.layers.{i}"
        attn(p + ".self_attn"); ffn(p); norm(p + ".norm1"); norm(p + ".norm2")
        p = f"tfm.decoder.layers.{i}"
        attn(p + ".self_attn"); attn(p + ".multihead_attn"); ffn(p)
        norm(p + ".norm1"); norm(p + ".norm2"); norm(p + ".norm3")
    norm("tfm.encoder.norm"); norm("tfm.decoder.norm")
    return sd
