"""Drop-in for the inference surface of the reference's `gill/models.py`: `GILLArgs`, `GILLModel` (`generate`),
`GILL` (`generate_for_images_and_texts`) and `load_gill`, executing on the libgillb200 kernels.

Same names, argument order, defaults, return structure and error behaviour as the reference (file:line cited inline).
Differences that are deliberate and documented in DESIGN.md:
  * the three frozen third-party models cannot be downloaded here, so they are injected: `lm` (an `OPTB200`),
    `sd_pipe` (a `StableDiffusionB200`) and optionally `visual_model` (CLIP vision tower: `gill_b200.clip.CLIPVisionB200`;
    prompts may instead carry already CLIP-encoded images as tensors);
  * `generate` speculatively appends the 8 [IMG] embeddings to every forward: OPT is causal, so the logits at the last
    real position are unchanged, and when [IMG0] is emitted the same forward already contains the next step's hidden
    states -- one prefill replaces the reference's two no-cache passes on the image-emission path;
  * `GILL.emit_images_batch` runs B independent equal-length prompts through the forced-emission path at once
    (one OPT prefill, one GILLMapper call, SD in chunks of 8) -- the per-sample equivalent of the reference's batch-1 loop.
"""
import glob
import json
import os
import pickle as pkl
from collections import namedtuple
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import layers, ops, retrieval
from .opt import OPTB200

# batched emission path: OPT prefill and GILLMapper replayed from CUDA graphs (one per input shape); "0": eager launches
GRAPH_STAGES = os.environ.get("GILLB200_GRAPH_STAGES", "1") != "0"

try:  # PIL is only needed for image prompts / PIL outputs
    from PIL import Image, UnidentifiedImageError
except Exception:  # pragma: no cover
    Image, UnidentifiedImageError = None, OSError


def truncate_caption(caption: str) -> str:
    """Truncate captions at periods and newlines (behaviour of gill/utils.py:32-40)."""
    caption = caption.strip("\n")
    cut = caption.find("\n") + 1
    if cut <= 0:
        cut = caption.find(".") + 1
    return caption[:cut] if cut > 0 else caption


def get_image_from_url(url: str):
    """gill/utils.py:24-29."""
    import requests
    from io import BytesIO

    response = requests.get(url)
    img = Image.open(BytesIO(response.content))
    return img.resize((224, 224)).convert("RGB")


class GILLArgs:  # gill/models.py:21-36
    freeze_lm: bool = True
    freeze_vm: bool = True
    opt_version: str = "facebook/opt-6.7b"
    visual_encoder: str = "openai/clip-vit-large-patch14"
    n_visual_tokens: int = 1
    task: str = "captioning"
    ret_emb_dim: Optional[int] = 256
    gen_emb_dim: Optional[int] = 256
    text_emb_layers: List[int] = [-1]
    gen_token_idx: List[int] = [0]
    retrieval_token_idx: List[int] = [0]
    text_fc_mode: str = "gill_mapper"
    ret_text_fc_mode: str = "linear"
    num_tokens: int = 8
    num_clip_tokens: int = 77


class _InputEmbeddings:
    """`self.input_embeddings` (gill/models.py:75): callable ids -> embeddings over the OPT table, `.weight` exposed."""

    def __init__(self, lm: OPTB200):
        self.lm = lm

    @property
    def weight(self):
        return self.lm.embed

    @property
    def embedding_dim(self):
        return self.lm.D

    def __call__(self, ids: torch.Tensor) -> torch.Tensor:
        return self.lm.embed_tokens(ids)


class _LinearB200(nn.Linear):
    """nn.Linear parameters, forward on the tcgen05 GEMM (split-precision activations, fp32 accumulate)."""

    def _packed(self):
        """bf16 weight / fp32 bias copies, rebuilt only when the parameters change (load_state_dict, .to(), .bfloat16())."""
        key = (self.weight.data_ptr(), self.weight._version, self.weight.dtype,
               None if self.bias is None else (self.bias.data_ptr(), self.bias._version))
        if getattr(self, "_pk_key", None) != key:
            self._pk = (self.weight.detach().to(torch.bfloat16).contiguous(),
                        self.bias.detach().float().contiguous() if self.bias is not None else None)
            self._pk_key = key
        return self._pk

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("gill_b200 linears run on CUDA (sm_100a) only; there is no CPU fallback")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        hi = torch.empty(x2.shape, device=x.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        ops.cast_add(x2, None, torch.bfloat16, out=hi, out_lo=lo)
        w, b = self._packed()
        o = ops.gemm(hi, w, a2=lo, a2_mode=2, bias=b, out_dtype=torch.float32)
        return o.view(*shp[:-1], -1).to(x.dtype)


class GILLModel(nn.Module):
    def __init__(self, tokenizer, args: GILLArgs = GILLArgs(), lm: Optional[OPTB200] = None, visual_model=None,
                 feature_extractor=None, visual_hidden_size: int = 1024):
        super().__init__()
        self.tokenizer = tokenizer
        self.feature_extractor = feature_extractor
        self.image_token = getattr(tokenizer, "cls_token_id", None)
        self.args = args
        self.num_tokens = args.num_tokens
        self.num_clip_tokens = args.num_clip_tokens
        if lm is None:
            raise ValueError("gill_b200.GILLModel needs an OPTB200 `lm` (pretrained OPT weights cannot be fetched "
                             "offline; build one with OPTB200(state_dict, ...) or OPTB200.random_init(...))")
        self.lm = lm
        self.opt_version = args.opt_version
        self.retrieval_token_idx = args.retrieval_token_idx
        self.gen_token_idx = args.gen_token_idx
        if lm.embed.shape[0] != len(tokenizer):                                    # models.py:73
            raise ValueError(f"OPT embedding table has {lm.embed.shape[0]} rows but the tokenizer has {len(tokenizer)}"
                             " tokens; resize the table (resize_token_embeddings) before constructing OPTB200")
        self.input_embeddings = _InputEmbeddings(lm)                               # models.py:75
        self.visual_model = visual_model
        self.visual_model_name = args.visual_encoder
        hidden_size = visual_hidden_size
        in_dim = lm.D
        embedding_dim = in_dim * args.n_visual_tokens                              # models.py:98
        self.ret_text_hidden_fcs = nn.ModuleList([])
        self.gen_text_hidden_fcs = nn.ModuleList([])
        for layer_idx in args.text_emb_layers:                                     # models.py:102-120
            if layer_idx == -1 or layer_idx == lm.L:
                self.ret_text_hidden_fcs.append(layers.TextFcLayer(
                    in_dim, args.ret_emb_dim, num_input_tokens=args.num_tokens, num_output_tokens=1,
                    mode=args.ret_text_fc_mode))
                self.gen_text_hidden_fcs.append(layers.TextFcLayer(
                    in_dim, args.gen_emb_dim, num_input_tokens=args.num_tokens,
                    num_output_tokens=args.num_clip_tokens, mode=args.text_fc_mode))
            else:
                raise ValueError(f"Only the last hidden layer is available from the B200 OPT forward "
                                 f"(text_emb_layers={args.text_emb_layers}); the shipped model uses [-1].")
        self.visual_embeddings = _LinearB200(hidden_size, embedding_dim)           # models.py:122
        self.visual_fc = _LinearB200(hidden_size, args.ret_emb_dim)                # models.py:125
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))         # models.py:126

    # -------------------------------------------------------------------------------------------- visual prefix
    def get_visual_embs(self, pixel_values: torch.Tensor, mode: str = "captioning"):
        """gill/models.py:129-152. `pixel_values` may be raw pixels (needs a CLIP `visual_model`) or already pooled
        CLIP features [n, hidden] (BASELINE config 5 feeds CLIP-encoded images)."""
        if mode not in ["captioning", "retrieval", "generation"]:
            raise ValueError(f"mode should be one of ['captioning', 'retrieval', 'generation'], got {mode} instead.")
        if pixel_values.dim() == 2:
            encoder_outputs = pixel_values
        else:
            if self.visual_model is None:
                raise NotImplementedError("raw pixels need a vision tower: pass visual_model=gill_b200.clip."
                                          "CLIPVisionB200(...) (SURVEY.md §8f-1) or pooled CLIP features [n, hidden]")
            encoder_outputs = self.visual_model(pixel_values).pooler_output
        if mode == "captioning":
            v = self.visual_embeddings(encoder_outputs)
            return torch.reshape(v, (v.shape[0], self.args.n_visual_tokens, -1))
        if mode == "retrieval":
            v = self.visual_fc(encoder_outputs)
            return torch.reshape(v, (v.shape[0], 1, -1))
        return torch.zeros((pixel_values.shape[0], 1, 768), device=pixel_values.device)

    def train(self, mode=True):  # gill/models.py:155-161 (returns None, like the reference)
        super(GILLModel, self).train(mode=mode)

    # -------------------------------------------------------------------------------------------- decode
    def _postprocess_logits(self, logits, i, min_word_tokens, ret_scale_factor, gen_scale_factor, filter_value):
        """gill/models.py:475-489 (in place, same order of operations)."""
        logits[:, self.retrieval_token_idx[1:]] = filter_value
        logits[:, self.gen_token_idx[1:]] = filter_value
        if (self.retrieval_token_idx or self.gen_token_idx) and self.retrieval_token_idx[0] != -1 \
                and self.gen_token_idx[0] != -1:
            if i < min_word_tokens:
                logits[:, self.retrieval_token_idx] = filter_value
                logits[:, self.gen_token_idx] = filter_value
            else:
                if ret_scale_factor > 1:
                    logits[:, self.retrieval_token_idx[0]] = logits[:, self.retrieval_token_idx[0]].abs() * ret_scale_factor
                if gen_scale_factor > 1:
                    logits[:, self.gen_token_idx[0]] = logits[:, self.gen_token_idx[0]].abs() * gen_scale_factor
        return logits

    def generate(self, embeddings=torch.FloatTensor, max_len: int = 32, temperature: float = 0.0, top_p: float = 1.0,
                 min_word_tokens: int = 0, ret_scale_factor: float = 1.0, gen_scale_factor: float = 1.0,
                 filter_value: float = -float("Inf"), speculative: bool = True, use_cache: bool = False):
        """gill/models.py:443-532. Returns (out ids [N,T'], [hidden_states[-1] per step], [last logits per step]).

        use_cache=True (not in the reference): incremental decoding with a per-layer K/V cache -- each step
        runs the decoder over the newly appended tokens only (1, or the 8 forced [IMG] tokens) instead of the whole
        sequence; hidden states and logits are bit-identical to the uncached path (tested), the cost per step drops
        from O(T) tokens to O(1).

        speculative=True (default): each forward runs over [embeddings | 8 speculative [IMG] embeddings]. Causality
        makes the first T positions identical to the reference's forward; if step i emits [IMG0] (batch 1), step i+1's
        forward IS the speculative one and is not recomputed."""
        import torch.nn.functional as F

        with torch.no_grad():
            out = None
            output_embeddings, output_logits = [], []
            dev = self.lm.dev
            embeddings = embeddings.to(dev)
            img_ids = torch.tensor(self.retrieval_token_idx, dtype=torch.int64, device=dev)
            can_spec = speculative and embeddings.shape[0] == 1 and self.retrieval_token_idx[0] != -1 \
                and self.retrieval_token_idx == self.gen_token_idx
            img_embs = self.input_embeddings(img_ids[None, :]) if can_spec else None
            n_img = len(self.retrieval_token_idx)  # tokens appended on an [IMG0] hit (models.py:518-520); != num_tokens
                                                   # under the default GILLArgs (retrieval_token_idx=[0])
            cached = None  # (hidden_states over T+8, logits at position T+7) carried over from a speculative hit
            kv, hs_all, n_new = None, None, embeddings.shape[1]
            if use_cache:
                can_spec = False
                kv = self.lm.new_cache(embeddings.shape[0], embeddings.shape[1] + max_len * max(1, n_img))
            for i in range(max_len):
                T = embeddings.shape[1]
                if kv is not None:
                    hs_new, lg = self.lm.forward(embeddings[:, T - n_new:], logit_positions=[n_new - 1], cache=kv)
                    hs_all = hs_new if hs_all is None else torch.cat([hs_all, hs_new], dim=1)
                    hs, logits, spec_hs = hs_all, lg[:, 0], None
                elif cached is not None:
                    hs, logits = cached
                    cached = None
                    spec_hs = None
                elif can_spec:
                    full = torch.cat([embeddings, img_embs.to(embeddings.dtype)], dim=1)
                    hs_full, lg2 = self.lm.forward(full, logit_positions=[T - 1, T + n_img - 1])
                    hs, logits = hs_full[:, :T], lg2[:, 0]
                    spec_hs = (hs_full, lg2[:, 1])
                else:
                    hs, lg = self.lm.forward(embeddings, logit_positions=[T - 1])
                    logits, spec_hs = lg[:, 0], None
                for idx in self.args.text_emb_layers:
                    output_embeddings.append(hs)                                          # models.py:467-468
                logits = logits.float().clone()                                            # models.py:470
                if top_p == 1.0:
                    logits = logits.cpu()                                                  # models.py:471-472
                output_logits.append(logits)
                self._postprocess_logits(logits, i, min_word_tokens, ret_scale_factor, gen_scale_factor, filter_value)
                if temperature == 0.0:
                    if top_p != 1.0:
                        raise ValueError("top_p cannot be set if temperature is 0 (greedy decoding).")  # :493
                    next_token = torch.argmax(logits, keepdim=True, dim=-1)                # models.py:494
                else:
                    logits = logits / temperature                                          # models.py:496
                    if top_p < 1.0:                                                        # models.py:499-512
                        assert top_p > 0, f"top_p should be above 0, got {top_p} instead."
                        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
                        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
                        sorted_indices_to_remove = cumulative_probs > top_p
                        sorted_indices_to_remove[..., 1:] = sorted_indices_to_remove[..., :-1].clone()
                        sorted_indices_to_remove[..., 0] = 0
                        for j in range(sorted_indices.shape[0]):
                            indices_to_remove = sorted_indices[j, sorted_indices_to_remove[j, :]]
                            logits[j, indices_to_remove] = filter_value
                    token_weights = logits.exp()                                           # models.py:514
                    next_token = torch.multinomial(token_weights, 1)                       # models.py:515
                hit = next_token.shape[0] == 1 and next_token.item() == self.retrieval_token_idx[0]
                if hit:                                                                    # models.py:518-520
                    assert self.retrieval_token_idx == self.gen_token_idx, (self.retrieval_token_idx, self.gen_token_idx)
                    next_token = torch.tensor(self.retrieval_token_idx)[None, :].long().to(dev)
                    if spec_hs is not None:
                        cached = spec_hs
                else:
                    next_token = next_token.long().to(dev)
                out = next_token if out is None else torch.cat([out, next_token], dim=-1)  # models.py:524-527
                next_embedding = self.input_embeddings(next_token)                         # models.py:529
                n_new = next_embedding.shape[1]
                embeddings = torch.cat([embeddings, next_embedding.to(embeddings.dtype)], dim=1)
        return out, output_embeddings, output_logits


class GILL(nn.Module):
    def __init__(self, tokenizer, model_args: Optional[GILLArgs] = None, path_array: Optional[List[str]] = None,
                 emb_matrix: Optional[torch.Tensor] = None, load_sd: bool = False, num_gen_images: int = 1,
                 decision_model_path: Optional[str] = None, *, lm: Optional[OPTB200] = None, sd_pipe=None,
                 visual_model=None, feature_extractor=None):
        super().__init__()
        self.model = GILLModel(tokenizer, model_args, lm=lm, visual_model=visual_model,
                               feature_extractor=feature_extractor)
        self.path_array = path_array
        self.emb_matrix = emb_matrix
        self.load_sd = load_sd
        self.num_gen_images = num_gen_images
        self.idx2dec = {0: "gen", 1: "ret", 2: "same"}
        self.decision_model = None
        if load_sd:                                                                        # models.py:549-551
            if sd_pipe is None:
                raise ValueError("load_sd=True needs an `sd_pipe` (StableDiffusionB200); SD-1.5 weights cannot be "
                                 "downloaded offline")
            self.sd_pipe = sd_pipe
        if decision_model_path is not None:                                                # models.py:553-561
            print("Loading decision model...")
            self.decision_model = nn.Sequential(*[nn.Dropout(0.5), _LinearB200(4096, 2)])
            mlp_checkpoint = torch.load(decision_model_path, map_location="cpu")
            self.decision_model.load_state_dict(mlp_checkpoint["state_dict"], strict=True)
            self.decision_model.eval()

    def __call__(self, images, tgt_tokens=None, caption_len=None, generate: bool = False, num_words: int = 32,
                 temperature: float = 1.0, top_p: float = 1.0, ret_scale_factor: float = 1.0,
                 gen_scale_factor: float = 1.0, min_word_tokens: int = 0, mode: str = "captioning",
                 concat_captions: bool = False, input_prefix: Optional[str] = None):
        if generate:                                                                       # models.py:568-571
            return self.model.generate(images, num_words, temperature=temperature, top_p=top_p,
                                       min_word_tokens=min_word_tokens, ret_scale_factor=ret_scale_factor,
                                       gen_scale_factor=gen_scale_factor)
        raise NotImplementedError("the training forward (gill/models.py:164-441) is outside the B200 hot path")

    # -------------------------------------------------------------------------------------------- prompt encoding
    def _encode_prompts(self, prompts: List, always_add_bos: bool):
        """gill/models.py:600-626. Extra prompt types: a tensor [hidden] / [n, hidden] of pooled CLIP features, or a
        tensor [n_visual_tokens, D] of ready visual-prefix embeddings."""
        m = self.model
        dev, dt = m.lm.dev, m.lm.dt
        input_embs, input_ids = [], []
        add_bos = True
        for p in prompts:
            if Image is not None and isinstance(p, Image.Image):
                if m.visual_model is None:
                    raise NotImplementedError("PIL prompts need a CLIP visual_model (gill_b200.clip.CLIPVisionB200, "
                                              "SURVEY.md §8f-1); pass CLIP-encoded tensors instead")
                if m.feature_extractor is not None:                               # models.py:608 (host, PIL/numpy)
                    pixel_values = m.feature_extractor(p.convert("RGB"), return_tensors="pt").pixel_values[0, ...]
                    pixel_values = pixel_values.to(device=dev, dtype=dt)[None, ...]
                else:
                    # the same pre-processing on the device: PIL-exact 8-bit bicubic resize of the shortest edge to 224,
                    # centre crop, rescale, normalise (gill/utils.py:117-119)
                    u8 = torch.from_numpy(np.asarray(p.convert("RGB")).copy()).to(dev)[None].contiguous()
                    pixel_values = ops.clip_preprocess_u8(u8, 224, out_dtype=dt, mode="feature_extractor")
                input_embs.append(m.get_visual_embs(pixel_values, mode="captioning").to(dt))
            elif isinstance(p, torch.Tensor):
                t = p.to(dev)
                if t.shape[-1] == m.lm.D:
                    input_embs.append(t.reshape(1, -1, m.lm.D).to(dt))
                else:
                    input_embs.append(m.get_visual_embs(t.reshape(-1, t.shape[-1]).to(dt), mode="captioning")
                                      .reshape(1, -1, m.lm.D).to(dt))
            elif type(p) == str:
                text_ids = m.tokenizer(p, add_special_tokens=add_bos, return_tensors="pt").input_ids.to(dev)
                if not always_add_bos:
                    add_bos = False
                input_embs.append(m.input_embeddings(text_ids))
                input_ids.append(text_ids)
            else:
                raise ValueError(f"Input prompts should be either PIL.Image.Image or str types, got {type(p)} instead.")
        return torch.cat(input_embs, dim=1), (torch.cat(input_ids, dim=1) if input_ids else None)

    # -------------------------------------------------------------------------------------------- the hot path
    def generate_for_images_and_texts(
            self, prompts: List, num_words: int = 0, min_word_tokens: int = 0, ret_scale_factor: float = 1.0,
            gen_scale_factor: float = 1.0, top_p: float = 1.0, temperature: float = 0.0, max_num_rets: int = 1,
            generator=None, always_add_bos: bool = False, guidance_scale: float = 7.5, num_inference_steps: int = 50):
        """gill/models.py:582-762: encode prompts, decode, and for each emitted [IMG0] run retrieval, the decision
        head, the GILLMapper and Stable Diffusion. Same return structure as the reference."""
        m = self.model
        with torch.no_grad():
            input_embs, input_ids = self._encode_prompts(prompts, always_add_bos)
            if num_words == 0:
                raise NotImplementedError("Generation not implemented for num_words=0.")     # models.py:629
            elif num_words > 0:
                generated_ids, generated_embeddings, _ = m.generate(
                    input_embs, num_words, min_word_tokens=min_word_tokens, temperature=temperature, top_p=top_p,
                    ret_scale_factor=ret_scale_factor, gen_scale_factor=gen_scale_factor)
                embeddings = generated_embeddings[-1][:, input_embs.shape[1]:]                # models.py:633
                newline_token_id = m.tokenizer("\n", add_special_tokens=False).input_ids[0]   # models.py:636-644
                trunc_idx = 0
                for j in range(generated_ids.shape[1]):
                    if generated_ids[0, j] == newline_token_id:
                        trunc_idx = j
                        break
                if trunc_idx > 0:
                    generated_ids = generated_ids[:, :trunc_idx]
                    embeddings = embeddings[:, :trunc_idx]
            else:
                raise ValueError

            return_outputs = []
            ids_host = generated_ids[0].tolist()
            all_ret_idx = [i for i, x in enumerate(ids_host) if x == m.retrieval_token_idx[0]][:max_num_rets]  # :651
            seen_image_idx = []
            last_ret_idx = 0
            if len(all_ret_idx) == 0:
                caption = m.tokenizer.batch_decode(generated_ids, skip_special_tokens=True)[0]
                return_outputs.append(truncate_caption(caption))
            else:
                for ret_idx in all_ret_idx:
                    assert ids_host[ret_idx:ret_idx + m.num_tokens] == m.retrieval_token_idx, \
                        (ids_host[ret_idx:ret_idx + m.num_tokens], m.retrieval_token_idx)     # models.py:661
                    raw_emb = embeddings[:, ret_idx:ret_idx + m.num_tokens, :]                # (1, 8, 4096)
                    assert len(m.args.text_emb_layers) == 1
                    image_outputs = {"gen": [], "ret": [], "decision": None}
                    ret_emb = None
                    if self.emb_matrix is not None:
                        ret_emb = m.ret_text_hidden_fcs[0](raw_emb.float(), None)[:, 0, :]    # models.py:673
                        ret_emb = ops.l2norm_rows(ret_emb.float().contiguous(), self.emb_matrix.dtype)  # :674-675
                        _, top_image_idx = retrieval.retrieval_topk(self.emb_matrix, ret_emb, 3,
                                                                    exclude_idx=seen_image_idx)       # :676-683
                        top_vals = _.cpu()[0].tolist()
                        for rank_i, img_idx in enumerate(top_image_idx[0].tolist()):
                            try:                                                               # models.py:686-693
                                seen_image_idx.append(img_idx)
                                img = get_image_from_url(self.path_array[img_idx])
                                image_outputs["ret"].append((img, "ret", top_vals[rank_i]))
                                if len(image_outputs) == max_num_rets:   # (sic) dict length, as in the reference
                                    break
                            except (UnidentifiedImageError, ConnectionError, OSError):
                                pass
                            except Exception as e:  # requests' own ConnectionError is an OSError subclass
                                if e.__class__.__name__ != "ConnectionError":
                                    raise
                        if self.decision_model is not None:                                    # models.py:696-701
                            decision_emb = raw_emb[:, 0, :]
                            assert decision_emb.shape[1] == 4096, decision_emb.shape
                            decision_logits = self.decision_model(decision_emb.float())
                            probs = decision_logits.softmax(dim=-1).cpu().float().numpy().tolist()
                            image_outputs["decision"] = [self.idx2dec[decision_logits.argmax().item()]] + probs
                    else:
                        image_outputs["decision"] = ["gen", [0, 1]]                            # models.py:704

                    gen_prefix = "".join([f"[IMG{i}]" for i in range(m.args.num_tokens)])     # models.py:707-710
                    gen_prefx_ids = m.tokenizer(gen_prefix, add_special_tokens=False,
                                                return_tensors="pt").input_ids.to(m.lm.dev)
                    gen_prefix_embs = m.input_embeddings(gen_prefx_ids)
                    gen_emb = m.gen_text_hidden_fcs[0](raw_emb.float(), gen_prefix_embs.float())  # (1, 77, 768)
                    if gen_emb.shape[1] != 77:                                                 # models.py:712-719
                        print(f"Padding {gen_emb.shape} with zeros")
                        bs, clip_emb = gen_emb.shape[0], 768
                        gen_emb = gen_emb.reshape(bs, -1, clip_emb)
                        seq_len = gen_emb.shape[1]
                        gen_emb = torch.cat([gen_emb, torch.zeros((bs, 77 - seq_len, clip_emb), device=gen_emb.device,
                                                                  dtype=gen_emb.dtype)], dim=1)
                        print("Padded to", gen_emb.shape)
                    gen_emb = gen_emb.repeat(self.num_gen_images, 1, 1)                        # models.py:721

                    if self.load_sd:                                                           # models.py:724-731
                        gen_max_bs = 8
                        gen_images = []
                        rerank = self.emb_matrix is not None and m.visual_model is not None
                        # our pipeline can hand the uint8 images over on the device: the re-rank pre-processing
                        # (PIL-exact bicubic 512 -> 224 + CLIP normalisation) then runs there too, models.py:733-737
                        on_device = rerank and m.feature_extractor is None and hasattr(self.sd_pipe, "vae")
                        gen_u8 = []
                        for i in range(0, self.num_gen_images, gen_max_bs):
                            res = self.sd_pipe(
                                prompt_embeds=gen_emb[i:i + gen_max_bs], generator=generator,
                                guidance_scale=guidance_scale, num_inference_steps=num_inference_steps,
                                **({"output_type": "uint8"} if on_device else {})).images
                            if on_device:
                                from PIL import Image

                                gen_u8.append(res)
                                gen_images.extend(Image.fromarray(a) for a in res.cpu().numpy())
                            else:
                                gen_images.extend(res)
                        if rerank:                                                             # models.py:733-751
                            if on_device:
                                all_gen_pixels = ops.clip_preprocess_u8(torch.cat(gen_u8, 0).contiguous(), 224,
                                                                        out_dtype=m.lm.dt)
                            else:
                                all_gen_pixels = []
                                for img in gen_images:
                                    pv = m.feature_extractor(img.resize((224, 224)).convert("RGB"),
                                                             return_tensors="pt").pixel_values[0, ...]
                                    all_gen_pixels.append(pv.to(device=m.lm.dev, dtype=m.lm.dt))
                                all_gen_pixels = torch.stack(all_gen_pixels, dim=0)
                            gen_visual_embs = m.get_visual_embs(all_gen_pixels, mode="retrieval")
                            gen_visual_embs = gen_visual_embs / gen_visual_embs.norm(dim=-1, keepdim=True)
                            gen_visual_embs = gen_visual_embs.type(self.emb_matrix.dtype)
                            gen_rank_scores = (gen_visual_embs @ ret_emb.T).squeeze()
                            sorted_score_idx = torch.argsort(-gen_rank_scores)
                            if self.num_gen_images > 1:
                                image_outputs["gen"] = [(gen_images[idx], gen_rank_scores[idx].item())
                                                        for idx in sorted_score_idx]
                            else:
                                image_outputs["gen"] = [(gen_images[0], gen_rank_scores.item())]
                        else:
                            # no bank (reference behaviour, models.py:753) or no CLIP tower to re-rank with
                            image_outputs["gen"] = [(gen_images[0], 0)]
                    else:
                        image_outputs["gen"] = [gen_emb]                                       # models.py:755

                    caption = m.tokenizer.batch_decode(generated_ids[:, last_ret_idx:ret_idx],
                                                       skip_special_tokens=True)[0]            # models.py:757-760
                    last_ret_idx = ret_idx + 1
                    return_outputs.append(truncate_caption(caption) + f" {gen_prefix}")
                    return_outputs.append(image_outputs)
        return return_outputs

    # -------------------------------------------------------------------------------------------- batched emission
    @torch.no_grad()
    def emit_images_batch(self, input_embs: torch.Tensor, latents: Optional[torch.Tensor] = None, generator=None,
                          guidance_scale: float = 7.5, num_inference_steps: int = 50, top_k: int = 0,
                          output_type: str = "uint8"):
        """B independent prompts of equal length through the forced-emission path the evals use
        (`num_words=2, gen_scale_factor=1e5`, evals/generate_vist_images.py:72-73): one OPT prefill over
        [prompt | 8 [IMG] embeddings], GILLMapper on the 8 [IMG] hidden states, optional retrieval top-k, SD in chunks
        of 8 (models.py:726). Per-sample equivalent of `generate_for_images_and_texts`; returns a dict with
        'images' (uint8 NHWC), 'gen_emb' [B,77,768], 'forced_ok' (bool per prompt: argmax after scaling is [IMG0]),
        and optionally 'ret' = (values, indices)."""
        m = self.model
        B, P, D = input_embs.shape
        if len(m.retrieval_token_idx) != m.num_tokens:                                         # models.py:661
            raise ValueError(f"emit_images_batch needs all {m.num_tokens} [IMG] token ids in retrieval_token_idx, "
                             f"got {m.retrieval_token_idx}")
        img_ids = torch.tensor(m.retrieval_token_idx, dtype=torch.int64, device=m.lm.dev)
        img_embs = m.input_embeddings(img_ids[None, :])                                        # (1, 8, D)
        full = torch.cat([input_embs.to(m.lm.dev, m.lm.dt), img_embs.expand(B, -1, -1).to(m.lm.dt)], dim=1)
        if GRAPH_STAGES and hasattr(m.lm, "forward_graphed"):
            hs, lg = m.lm.forward_graphed(full, logit_positions=[P - 1])   # ~400 launches replayed from one CUDA graph
        else:
            hs, lg = m.lm.forward(full, logit_positions=[P - 1])
        logits = lg[:, 0].float()
        m._postprocess_logits(logits, 0, 0, 1.0, 1e5, -float("Inf"))
        forced_ok = logits.argmax(dim=-1) == m.retrieval_token_idx[0]
        raw_emb = hs[:, P:P + m.num_tokens, :].float().contiguous()                             # (B, 8, D)
        out = {"forced_ok": forced_ok}
        if top_k > 0 and self.emb_matrix is not None:
            ret_emb = m.ret_text_hidden_fcs[0](raw_emb, None)[:, 0, :]
            q = ops.l2norm_rows(ret_emb.float().contiguous(), self.emb_matrix.dtype)
            out["ret"] = retrieval.retrieval_topk(self.emb_matrix, q, top_k)
        mapper = m.gen_text_hidden_fcs[0]
        if GRAPH_STAGES:
            gc = getattr(mapper, "_graphed", None)
            if gc is None:
                gc = mapper._graphed = ops.GraphedCall(lambda x, e: mapper(x, e), raw_emb.device)
            gen_emb = gc(raw_emb, img_embs.float().contiguous()).clone()                       # (B, 77, 768)
        else:
            gen_emb = mapper(raw_emb, img_embs.float())
        out["gen_emb"] = gen_emb
        if self.load_sd:
            imgs = []
            for i in range(0, B, 8):                                                           # models.py:726-731
                lat = None if latents is None else latents[i:i + 8]
                imgs.append(self.sd_pipe(prompt_embeds=gen_emb[i:i + 8], generator=generator, latents=lat,
                                         guidance_scale=guidance_scale, num_inference_steps=num_inference_steps,
                                         output_type=output_type).images)
            out["images"] = torch.cat(imgs, 0) if output_type == "uint8" else sum(imgs, [])
        return out

    # -------------------------------------------------------------------------------------------- batched surface
    @torch.no_grad()
    def generate_for_images_and_texts_batch(
            self, prompt_lists: List[List], num_gen_images: Optional[int] = None, max_num_rets: int = 3,
            generator=None, latents: Optional[torch.Tensor] = None, guidance_scale: float = 7.5,
            num_inference_steps: int = 50, always_add_bos: bool = False, bank=None, seen: Optional[List[List[int]]] = None,
            fetch_images: bool = False, output_type: str = "uint8"):
        """B independent conversations through the forced-emission path of `generate_for_images_and_texts`
        (`num_words=2, gen_scale_factor=1e5`, evals/generate_vist_images.py:72-73), each the per-sample equivalent of the
        reference's strictly batch-1 call (models.py:582-762) -- BASELINE configs[4]:

          * prompts of UNEQUAL length are right-padded to one [B, P_max + 8, D] prefill: OPT is causal, so the 8 [IMG]
            embeddings placed directly after each prompt's last token see exactly that prompt and nothing of the padding
            (per-sample [IMG] offsets; positions are left-aligned, so the learned position embeddings match too);
          * retrieval for all prompts at once, each with its own seen list (`seen`), through `bank` (a ShardedBank:
            row-sharded across the ranks, one candidate exchange) or `self.emb_matrix`;
          * decision head, GILLMapper, `num_gen_images` Stable-Diffusion samples per prompt in chunks of 8
            (models.py:726), then the CLIP re-rank of models.py:733-751 on the device when a vision tower is present.

        Returns one `[caption + ' [IMG0]...[IMG7]', {'gen': [(image, score)...], 'ret': [...], 'decision': ...}]` pair per
        prompt list; images are uint8 HWC tensors (output_type='uint8') or PIL images ('pil'); 'ret' holds
        (row index, 'ret', score) triples, or fetched images when fetch_images=True."""
        m = self.model
        dev, dt = m.lm.dev, m.lm.dt
        G = self.num_gen_images if num_gen_images is None else num_gen_images
        if len(m.retrieval_token_idx) != m.num_tokens:
            raise ValueError(f"needs all {m.num_tokens} [IMG] token ids in retrieval_token_idx, got {m.retrieval_token_idx}")
        B = len(prompt_lists)
        embs = [self._encode_prompts(pl, always_add_bos)[0][0] for pl in prompt_lists]            # [P_i, D] each
        lens = [int(e.shape[0]) for e in embs]
        n_img = m.num_tokens
        T = max(lens) + n_img
        img_ids = torch.tensor(m.retrieval_token_idx, dtype=torch.int64, device=dev)
        img_embs = m.input_embeddings(img_ids[None, :])                                          # (1, 8, D)
        full = torch.zeros((B, T, m.lm.D), device=dev, dtype=dt)
        for i, e in enumerate(embs):
            full[i, :lens[i]] = e.to(dt)
            full[i, lens[i]:lens[i] + n_img] = img_embs[0].to(dt)
        hs, _ = m.lm.forward(full, need_logits=False)
        rows = torch.arange(B, device=dev)
        lens_t = torch.tensor(lens, device=dev)
        logits = m.lm.logits_of(hs[rows, lens_t - 1]).float()                                     # last prompt position
        m._postprocess_logits(logits, 0, 0, 1.0, 1e5, -float("Inf"))                              # models.py:475-489
        forced_ok = (logits.argmax(dim=-1) == m.retrieval_token_idx[0]).tolist()
        gather = (lens_t[:, None] + torch.arange(n_img, device=dev)[None, :])                     # per-sample [IMG] offsets
        raw_emb = hs[rows[:, None], gather].float().contiguous()                                   # (B, 8, D) models.py:662
        gen_prefix = "".join([f"[IMG{i}]" for i in range(n_img)])
        outs = [{"gen": [], "ret": [], "decision": None} for _ in range(B)]
        # ---- retrieval (models.py:671-693) for all prompts in one kernel launch / one exchange
        ret_q = None
        have_bank = bank is not None or self.emb_matrix is not None
        if have_bank:
            ret_emb = m.ret_text_hidden_fcs[0](raw_emb, None)[:, 0, :]
            ret_q = ops.l2norm_rows(ret_emb.float().contiguous(), torch.bfloat16)
            seen_l = seen if seen is not None else [[] for _ in range(B)]
            ex = seen_l if any(len(x) for x in seen_l) else None
            if bank is not None:
                vals, idx = bank.search(ret_q, 3, exclude_idx=ex)
            else:
                vals, idx = retrieval.retrieval_topk(self.emb_matrix, ret_q, 3, exclude_idx=ex)
            vals, idx = vals.cpu().tolist(), idx.cpu().tolist()
            for i in range(B):
                for rank_i, img_idx in enumerate(idx[i][:max(1, max_num_rets)]):
                    if fetch_images:
                        try:
                            outs[i]["ret"].append((get_image_from_url(self.path_array[img_idx]), "ret", vals[i][rank_i]))
                        except Exception:
                            pass
                    else:
                        outs[i]["ret"].append((img_idx, "ret", vals[i][rank_i]))
            if self.decision_model is not None:                                                   # models.py:696-701
                dl = self.decision_model(raw_emb[:, 0, :].float())
                probs = dl.softmax(dim=-1).cpu().float().tolist()
                arg = dl.argmax(dim=-1).tolist()
                for i in range(B):
                    outs[i]["decision"] = [self.idx2dec[arg[i]], probs[i]]
        else:
            for i in range(B):
                outs[i]["decision"] = ["gen", [0, 1]]                                             # models.py:704
        # ---- GILLMapper + Stable Diffusion (models.py:707-731)
        gen_emb = m.gen_text_hidden_fcs[0](raw_emb, img_embs.float())                             # (B, 77, 768)
        result_imgs = None
        if self.load_sd:
            rep = gen_emb.repeat_interleave(G, dim=0)                                             # models.py:721, per prompt
            chunks = []
            for c0 in range(0, B * G, 8):                                                         # gen_max_bs = 8
                lat = None if latents is None else latents[c0:c0 + 8]
                chunks.append(self.sd_pipe(prompt_embeds=rep[c0:c0 + 8], generator=generator, latents=lat,
                                           guidance_scale=guidance_scale, num_inference_steps=num_inference_steps,
                                           output_type="uint8").images)
            result_imgs = torch.cat(chunks, 0)                                                    # (B*G, H, W, 3) uint8
            scores = None
            if have_bank and m.visual_model is not None:                                          # models.py:733-751
                px = ops.clip_preprocess_u8(result_imgs.contiguous(), 224, out_dtype=dt)
                ve = m.get_visual_embs(px, mode="retrieval")[:, 0, :].float()
                ve = ve / ve.norm(dim=-1, keepdim=True)
                scores = (ve.view(B, G, -1).to(torch.bfloat16).float() * ret_q.float()[:, None, :]).sum(-1)   # (B, G)
            for i in range(B):
                imgs_i = result_imgs[i * G:(i + 1) * G]
                if output_type == "pil":
                    imgs_i = [Image.fromarray(a) for a in imgs_i.cpu().numpy()]
                if scores is not None:
                    order = torch.argsort(-scores[i]).tolist() if G > 1 else [0]
                    sc = scores[i].tolist()
                    outs[i]["gen"] = [(imgs_i[j], sc[j]) for j in order]
                else:
                    outs[i]["gen"] = [(imgs_i[0], 0)]                                             # models.py:753
        else:
            for i in range(B):
                outs[i]["gen"] = [gen_emb[i:i + 1]]                                               # models.py:755
        ret = []
        for i in range(B):
            ret.append([f" {gen_prefix}", outs[i]])
        self.last_batch_info = {"forced_ok": forced_ok, "prompt_lens": lens, "gen_emb": gen_emb, "images": result_imgs,
                                "ret_q": ret_q, "raw_emb": raw_emb}
        return ret

    @torch.no_grad()
    def get_log_likelihood_scores(self, prompts: List):
        """gill/models.py:764-807: log likelihood of an interleaved prompt = minus the mean next-token cross entropy over
        the text positions (image positions carry label -100; <bos> only on the first string). The reference gets it
        from `lm(..., labels=input_ids).loss`; here: OPT forward -> tied lm_head on every position -> row softmax kernel."""
        m = self.model
        dev, dt = m.lm.dev, m.lm.dt
        input_embs, input_ids = [], []
        add_bos = True
        for p in prompts:
            if (Image is not None and isinstance(p, Image.Image)) or isinstance(p, torch.Tensor):
                e, _ = self._encode_prompts([p], always_add_bos=False)
                input_embs.append(e)
                input_ids.append(torch.full(e.shape[:2], -100, dtype=torch.int64, device=dev))
            elif type(p) == str:
                text_ids = m.tokenizer(p, add_special_tokens=True, return_tensors="pt").input_ids.to(dev)
                if not add_bos:
                    text_ids = text_ids[:, 1:]                                           # models.py:792-797
                else:
                    add_bos = False
                input_embs.append(m.input_embeddings(text_ids))
                input_ids.append(text_ids)
            else:
                raise ValueError(f"Input prompts should be either PIL.Image.Image or str types, got {type(p)} instead.")
        embs = torch.cat([e.to(dt) for e in input_embs], dim=1)
        ids = torch.cat(input_ids, dim=1)
        hs, _ = m.lm.forward(embs, need_logits=False)
        labels = ids[0, 1:]
        keep = labels != -100
        if not bool(keep.any()):
            return float("nan")                                                          # mean over zero tokens, as torch does
        # position t predicts token t+1; the vocabulary is padded to a multiple of 8 columns of -inf for the vector kernels
        V = m.lm.embed.shape[0]
        Vp = (V + 7) // 8 * 8
        logits = torch.full((embs.shape[1] - 1, Vp), -float("inf"), device=dev, dtype=torch.float32)
        ops.gemm(hs[0, :-1].contiguous(), m.lm.embed, out=logits[:, :V])
        probs = ops.softmax_rows(logits, 1.0, torch.float32)
        tok = probs[torch.arange(labels.numel(), device=dev), labels.clamp_min(0)]
        return float(tok[keep].log().mean().item())


def load_gill(model_dir: str, load_ret_embs: bool = True, decision_model_fn: str = "decision_model.pth.tar", *,
              tokenizer=None, lm: Optional[OPTB200] = None, sd_pipe=None, visual_model=None, feature_extractor=None,
              device: str = "cuda") -> GILL:
    """gill/models.py:810-902. Reads model_args.json, pretrained_ckpt.pth.tar (keys `module.model.*`) and the
    cc3m*.npy pickles exactly like the reference. The frozen third-party models are injected (see module docstring)."""
    if lm is None:
        raise ValueError("load_gill needs `lm=` (an OPTB200 holding the OPT weights): pretrained OPT cannot be fetched "
                         "offline, see the module docstring")
    model_args_path = os.path.join(model_dir, "model_args.json")
    model_ckpt_path = os.path.join(model_dir, "pretrained_ckpt.pth.tar")
    embs_paths = [s for s in glob.glob(os.path.join(model_dir, "cc3m*.npy"))]
    if not os.path.exists(model_args_path):
        raise ValueError(f"model_args.json does not exist in {model_dir}.")
    if not os.path.exists(model_ckpt_path):
        raise ValueError(f"pretrained_ckpt.pth.tar does not exist in {model_dir}.")
    if not load_ret_embs or len(embs_paths) == 0:
        if len(embs_paths) == 0:
            print(f"cc3m.npy files do not exist in {model_dir}.")
        print("Running the model without retrieval.")
        path_array, emb_matrix = None, None
    else:
        path_array, emb_matrix = [], []
        for p in embs_paths:                                                                   # models.py:831-835
            with open(p, "rb") as wf:
                train_embs_data = pkl.load(wf)
                path_array.extend(train_embs_data["paths"])
                emb_matrix.extend(train_embs_data["embeddings"])
        emb_matrix = np.stack(emb_matrix, axis=0)
        assert len(path_array) == emb_matrix.shape[0], (len(path_array), emb_matrix.shape)

    with open(model_args_path, "r") as f:
        model_kwargs = json.load(f)
    if tokenizer is None:
        from transformers import AutoTokenizer

        tokenizer = AutoTokenizer.from_pretrained(model_kwargs["opt_version"], use_fast=False)   # models.py:845
        if tokenizer.pad_token is None:
            tokenizer.pad_token_id = tokenizer.eos_token_id
        tokenizer.add_special_tokens({"cls_token": "<|image|>"})
        for i in range(model_kwargs["num_tokens"]):
            tokenizer.add_tokens(f"[IMG{i}]")
    model_kwargs["retrieval_token_idx"] = []
    for i in range(model_kwargs["num_tokens"]):                                                # models.py:853-860
        ret_token_idx = tokenizer(f"[IMG{i}]", add_special_tokens=False).input_ids
        assert len(ret_token_idx) == 1, ret_token_idx
        model_kwargs["retrieval_token_idx"].append(ret_token_idx[0])
    model_kwargs["gen_token_idx"] = model_kwargs["retrieval_token_idx"]                        # models.py:862
    args = namedtuple("args", model_kwargs)(**model_kwargs)
    # like the reference (models.py:866-873): a named decision model that is missing is an error, not a silent skip
    decision_model_path = os.path.join(model_dir, decision_model_fn) if decision_model_fn is not None else None

    model = GILL(tokenizer, args, path_array=path_array, emb_matrix=emb_matrix, load_sd=sd_pipe is not None,
                 num_gen_images=1, decision_model_path=decision_model_path, lm=lm, sd_pipe=sd_pipe,
                 visual_model=visual_model, feature_extractor=feature_extractor)
    model = model.eval()
    model = model.bfloat16()
    model = model.to(device)

    checkpoint = torch.load(model_ckpt_path, map_location="cpu")                               # models.py:880-884
    state_dict = {k.replace("module.", ""): v for k, v in checkpoint["state_dict"].items()}
    img_token_embeddings = state_dict["model.input_embeddings.weight"].cpu().detach()
    del state_dict["model.input_embeddings.weight"]
    # the reference loads with strict=False (models.py:887) because the checkpoint holds the trained parameters only;
    # here every checkpoint key must land somewhere (no silent key-name drift)
    res = model.load_state_dict(state_dict, strict=False)
    if res.unexpected_keys:
        raise RuntimeError(f"pretrained_ckpt.pth.tar has keys this model does not know: {res.unexpected_keys[:8]}")
    missing = [k for k in res.missing_keys if not k.startswith("decision_model.")]
    if missing:
        raise RuntimeError(f"pretrained_ckpt.pth.tar lacks trained parameters: {missing[:8]}")
    with torch.no_grad():                                                                      # models.py:890-893
        if "share_ret_gen" in model_kwargs:
            assert model_kwargs["share_ret_gen"], "Model loading only supports share_ret_gen=True for now."
        lm.embed[-model_kwargs["num_tokens"]:, :].copy_(img_token_embeddings.to(lm.embed.dtype))
    if load_ret_embs and len(embs_paths) > 0:                                                  # models.py:895-900
        model.emb_matrix = retrieval.prepare_bank(emb_matrix, model.model.logit_scale.detach())
    return model
