"""Frozen OPT decoder forward on the libgillb200 kernels (the `self.lm(...)` call of gill/models.py:465).

The reference calls HF `OPTForCausalLM(inputs_embeds=..., use_cache=False, output_hidden_states=True)` on the WHOLE
sequence at every decode step and consumes only `hidden_states[-1]` (all positions) and `logits[:, -1]`. This class
computes exactly those two things: one fused-QKV tcgen05 GEMM, one causal flash-attention launch, out_proj / fc1 / fc2
GEMMs with bias + ReLU + residual in their epilogues, fp32 residual stream, and the tied lm_head only on the last
position of each sequence.
"""
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops

SD = Dict[str, torch.Tensor]


class KVCache:
    """Key/value buffers of every decoder layer, [B, max_len, D] in the model dtype, plus the number of tokens held."""

    def __init__(self, layers: int, batch: int, max_len: int, hidden: int, device, dtype):
        self.batch, self.max_len, self.len = batch, max_len, 0
        self.k = [torch.empty((batch, max_len, hidden), device=device, dtype=dtype) for _ in range(layers)]
        self.v = [torch.empty((batch, max_len, hidden), device=device, dtype=dtype) for _ in range(layers)]


class OPTB200:
    def __init__(self, sd: SD, hidden: int, layers: int, heads: int, ffn: int, device="cuda", dtype=torch.bfloat16):
        """sd: OPTForCausalLM-named state dict (`model.decoder.*`); weights are stored bf16, biases / norms fp32."""
        self.D, self.L, self.H, self.F = hidden, layers, heads, ffn
        self.hd = hidden // heads
        if self.hd not in (64, 128):
            raise ValueError(f"OPTB200 supports head_dim 64 (OPT-125M, BASELINE configs[0]) and 128 (OPT-6.7B family), "
                             f"got {self.hd}")
        self.dev, self.dt = torch.device(device), dtype
        f32 = torch.float32
        p = "model.decoder."
        g = lambda k, dt=None: sd[k].to(self.dev, dt or dtype).contiguous()
        self.embed = g(p + "embed_tokens.weight")
        self.pos = g(p + "embed_positions.weight")
        self.layers = []
        for i in range(layers):
            l = f"{p}layers.{i}."
            self.layers.append(dict(
                ln1_w=g(l + "self_attn_layer_norm.weight", f32), ln1_b=g(l + "self_attn_layer_norm.bias", f32),
                qkv_w=torch.cat([g(l + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")], 0),
                qkv_b=torch.cat([g(l + f"self_attn.{n}.bias", f32) for n in ("q_proj", "k_proj", "v_proj")], 0),
                o_w=g(l + "self_attn.out_proj.weight"), o_b=g(l + "self_attn.out_proj.bias", f32),
                ln2_w=g(l + "final_layer_norm.weight", f32), ln2_b=g(l + "final_layer_norm.bias", f32),
                fc1_w=g(l + "fc1.weight"), fc1_b=g(l + "fc1.bias", f32),
                fc2_w=g(l + "fc2.weight"), fc2_b=g(l + "fc2.bias", f32)))
        self.lnf_w, self.lnf_b = g(p + "final_layer_norm.weight", f32), g(p + "final_layer_norm.bias", f32)

    @classmethod
    def random_init(cls, hidden, layers, heads, ffn, vocab, max_pos=2048, seed=0, device="cuda", std=0.02):
        """Seeded HF-style init generated directly on the device (no pretrained weights exist offline)."""
        g = torch.Generator(device=device).manual_seed(seed)
        n = lambda *s: (torch.randn(*s, generator=g, device=device, dtype=torch.float32) * std).to(torch.bfloat16)
        p = "model.decoder."
        sd = {p + "embed_tokens.weight": n(vocab, hidden), p + "embed_positions.weight": n(max_pos + 2, hidden)}
        for i in range(layers):
            l = f"{p}layers.{i}."
            for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
                sd[l + f"self_attn.{nm}.weight"], sd[l + f"self_attn.{nm}.bias"] = n(hidden, hidden), n(hidden)
            sd[l + "fc1.weight"], sd[l + "fc1.bias"] = n(ffn, hidden), n(ffn)
            sd[l + "fc2.weight"], sd[l + "fc2.bias"] = n(hidden, ffn), n(hidden)
            for nm in ("self_attn_layer_norm", "final_layer_norm"):
                sd[l + nm + ".weight"] = torch.ones(hidden, device=device)
                sd[l + nm + ".bias"] = torch.zeros(hidden, device=device)
        sd[p + "final_layer_norm.weight"] = torch.ones(hidden, device=device)
        sd[p + "final_layer_norm.bias"] = torch.zeros(hidden, device=device)
        return cls(sd, hidden, layers, heads, ffn, device=device)

    # ------------------------------------------------------------------------------------------------------------
    def get_input_embeddings(self):
        return self.embed

    def embed_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        """`self.input_embeddings(ids)` (gill/models.py:620, :529, :709): [..] int64 -> [.., D]."""
        flat = ids.reshape(-1).to(self.dev, torch.int64).contiguous()
        return ops.gather_add_rows(self.embed, flat).view(*ids.shape, self.D)

    def new_cache(self, batch: int, max_len: int) -> "KVCache":
        """Per-layer K/V buffers for incremental decoding (the reference itself decodes without a cache,
        gill/models.py:465 `use_cache=False`, recomputing the whole sequence at every step)."""
        return KVCache(self.L, batch, max_len, self.D, self.dev, self.dt)

    @torch.no_grad()
    def forward(self, inputs_embeds: torch.Tensor, need_logits: bool = True, logit_positions=None,
                cache: Optional["KVCache"] = None):
        """inputs_embeds [B,T,D] -> (hidden_states[-1] [B,T,D] in the model dtype, logits fp32).
        logits: [B,V] at the last position, or [B,len(logit_positions),V] at the requested positions.
        With `cache`, inputs_embeds are the T NEW tokens following the cache.len tokens already processed: their K/V are
        appended and attention runs over the cached keys (causal with offset). Every position's arithmetic is the same
        as in the uncached forward (row-independent GEMMs, same KV tiling), so hidden states are bit-identical."""
        B, T, D = inputs_embeds.shape
        H, hd = self.H, self.hd
        past = 0 if cache is None else cache.len
        if cache is not None and (B != cache.batch or past + T > cache.max_len):
            raise ValueError(f"KV cache holds batch {cache.batch} x {cache.max_len} tokens; got batch {B}, {past}+{T} tokens")
        x = inputs_embeds.to(self.dev, self.dt).contiguous().view(B * T, D)
        pos_idx = torch.arange(past, past + T, device=self.dev, dtype=torch.int64).repeat(B)
        h16 = ops.gather_add_rows(self.pos, pos_idx, x=x, idx_offset=2)           # + embed_positions(pos + 2)
        h = ops.cast_add(h16, None, torch.float32)                                # fp32 residual stream
        for li, ly in enumerate(self.layers):
            n = ops.layernorm(h, ly["ln1_w"], ly["ln1_b"], 1e-5, out_dtype=self.dt)
            qkv = ops.gemm(n, ly["qkv_w"], bias=ly["qkv_b"]).view(B, T, 3 * D)
            if cache is None:
                a = ops.attention(qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:], H, hd, hd ** -0.5, causal=True)
            else:
                cache.k[li][:, past:past + T].copy_(qkv[:, :, D:2 * D])
                cache.v[li][:, past:past + T].copy_(qkv[:, :, 2 * D:])
                a = ops.attention(qkv[:, :, :D], cache.k[li][:, :past + T], cache.v[li][:, :past + T], H, hd, hd ** -0.5,
                                  causal=True, causal_offset=past)
            h = ops.gemm(a.view(B * T, D), ly["o_w"], bias=ly["o_b"], residual=h, out_dtype=torch.float32)
            n = ops.layernorm(h, ly["ln2_w"], ly["ln2_b"], 1e-5, out_dtype=self.dt)
            f = ops.gemm(n, ly["fc1_w"], bias=ly["fc1_b"], act="relu")
            h = ops.gemm(f, ly["fc2_w"], bias=ly["fc2_b"], residual=h, out_dtype=torch.float32)
        if cache is not None:
            cache.len = past + T
        hs = ops.layernorm(h, self.lnf_w, self.lnf_b, 1e-5, out_dtype=self.dt).view(B, T, D)
        logits = None
        if logit_positions is not None:
            # (slices + stack: an index list would be an implicit host-to-device copy, illegal inside a graph capture)
            sel = torch.stack([hs[:, int(pp), :] for pp in logit_positions], dim=1).reshape(B * len(logit_positions), D)
            logits = ops.gemm(sel, self.embed, out_dtype=torch.float32).view(B, len(logit_positions), -1)
        elif need_logits:
            last = hs[:, -1, :].contiguous()
            logits = ops.gemm(last, self.embed, out_dtype=torch.float32)          # tied lm_head, last position only
        return hs, logits

    @torch.no_grad()
    def forward_graphed(self, inputs_embeds: torch.Tensor, logit_positions=None):
        """`forward` (no KV cache) replayed from a CUDA graph captured per (B, T, logit positions): a prefill is ~12
        launches per layer whose eager submission through ctypes costs more CPU time than the GPU needs at batch 8 x 81
        tokens. Same kernels, same arithmetic; the returned tensors are the graph's static outputs (valid until the next
        call with the same shape). For fixed-shape callers (the batched emission path); `generate` keeps eager launches."""
        gc = getattr(self, "_graphed", None)
        if gc is None:
            gc = self._graphed = {}
        key = tuple(logit_positions) if logit_positions is not None else None
        g = gc.get(key)
        if g is None:
            if len(gc) >= 4:
                gc.pop(next(iter(gc)))
            g = gc[key] = ops.GraphedCall(lambda x: self.forward(x, logit_positions=logit_positions), self.dev)
        return g(inputs_embeds.to(self.dev, self.dt).contiguous())

    @torch.no_grad()
    def logits_of(self, hidden_rows: torch.Tensor) -> torch.Tensor:
        """Tied lm_head on selected post-final-LayerNorm hidden states [n, D] -> fp32 logits [n, V]."""
        return ops.gemm(hidden_rows.to(self.dev, self.dt).contiguous(), self.embed, out_dtype=torch.float32)

    def __call__(self, inputs_embeds=None, use_cache=False, output_hidden_states=True, **kw):
        """HF-shaped result for the reference's call site: `.logits[:, -1, :]` and `.hidden_states[-1]` are valid."""
        if inputs_embeds is None:
            raise ValueError("OPTB200 is driven with inputs_embeds (gill/models.py:465)")
        hs, logits = self.forward(inputs_embeds)
        return SimpleNamespace(logits=logits[:, None, :], hidden_states=(hs,))
