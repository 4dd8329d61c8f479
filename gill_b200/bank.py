"""On-disk format of the retrieval bank.

The reference ships the CC3M bank as `cc3m*.npy` files that are really PICKLES of {'paths': [str...], 'embeddings':
[tensor...]} (gill/models.py:813-839): loading means unpickling ~3 M Python objects, stacking them and only then
preparing the matrix on the GPU (:895-900). This module stores the PREPARED bank (bf16, row-normalised, multiplied by
exp(logit_scale) -- exactly what `retrieval_topk` streams) as flat little-endian row-major shards that are memory-mapped
and copied straight into HBM, one contiguous row range per GPU:

    <dir>/meta.json            {"n": N, "d": D, "dtype": "bfloat16", "shards": S, "rows": [[lo, hi], ...], "logit_scale": x}
    <dir>/bank.<s>.bf16        rows [lo_s, hi_s) as raw uint16 (bf16 bit patterns), row-major
    <dir>/paths.txt            one path/URL per bank row (line i <-> global row i)

Row ranges are `retrieval.shard_rows(N, S, s)`, so a world of S ranks loads one file each; any other world size reads
the overlapping files.
"""
import json
import os
import pickle
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .retrieval import prepare_bank, shard_rows

_META = "meta.json"


def save_prepared_bank(bank: torch.Tensor, paths: Sequence[str], out_dir: str, shards: int = 1,
                       logit_scale: Optional[float] = None) -> None:
    """bank: the PREPARED [N,D] bf16 matrix (retrieval.prepare_bank). Writes the layout described above."""
    if bank.dtype != torch.bfloat16 or bank.dim() != 2:
        raise ValueError(f"expected a prepared [N,D] bfloat16 bank, got {tuple(bank.shape)} {bank.dtype}")
    n, d = bank.shape
    if len(paths) != n:
        raise ValueError(f"{len(paths)} paths for {n} bank rows")
    os.makedirs(out_dir, exist_ok=True)
    raw = bank.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
    rows = []
    for s in range(shards):
        lo, hi = shard_rows(n, shards, s)
        raw[lo:hi].tofile(os.path.join(out_dir, f"bank.{s}.bf16"))
        rows.append([lo, hi])
    with open(os.path.join(out_dir, "paths.txt"), "w") as f:
        for p in paths:
            if "\n" in p:
                raise ValueError("bank paths must not contain newlines")
            f.write(p + "\n")
    with open(os.path.join(out_dir, _META), "w") as f:
        json.dump({"n": n, "d": d, "dtype": "bfloat16", "shards": shards, "rows": rows, "logit_scale": logit_scale}, f)


def read_reference_bank(npy_paths: Sequence[str]) -> Tuple[List[str], np.ndarray]:
    """The reference's loader (gill/models.py:829-839): unpickle every file, concatenate paths and embeddings."""
    path_array, embs = [], []
    for p in npy_paths:
        with open(p, "rb") as f:
            d = pickle.load(f)
        path_array.extend(d["paths"])
        embs.extend(np.asarray(e, dtype=np.float32) if not torch.is_tensor(e) else e.float().numpy()
                    for e in d["embeddings"])
    emb = np.stack(embs, axis=0)
    assert len(path_array) == emb.shape[0], (len(path_array), emb.shape)          # models.py:838
    return path_array, emb


def convert_reference_bank(npy_paths: Sequence[str], logit_scale: torch.Tensor, out_dir: str, shards: int = 1) -> None:
    """cc3m*.npy pickles -> prepared flat shards. `logit_scale` is the checkpoint's (bf16) parameter: the preparation is
    the reference's own expression evaluated in its dtype on its device (retrieval.prepare_bank, models.py:896-899)."""
    paths, emb = read_reference_bank(npy_paths)
    bank = prepare_bank(emb, logit_scale)
    save_prepared_bank(bank.to(torch.bfloat16), paths, out_dir, shards, float(logit_scale.float().item()))


def bank_meta(bank_dir: str) -> dict:
    with open(os.path.join(bank_dir, _META)) as f:
        return json.load(f)


def load_bank_rows(bank_dir: str, lo: int, hi: int, device="cuda") -> torch.Tensor:
    """Rows [lo, hi) of the global bank as a bf16 tensor on `device` (memory-mapped files -> pinned staging -> HBM)."""
    meta = bank_meta(bank_dir)
    n, d = meta["n"], meta["d"]
    if not (0 <= lo <= hi <= n):
        raise ValueError(f"row range [{lo},{hi}) outside the bank's {n} rows")
    out = torch.empty((hi - lo, d), dtype=torch.bfloat16, device=device)
    for s, (slo, shi) in enumerate(meta["rows"]):
        a, b = max(lo, slo), min(hi, shi)
        if a >= b:
            continue
        mm = np.memmap(os.path.join(bank_dir, f"bank.{s}.bf16"), dtype=np.uint16, mode="r", shape=(shi - slo, d))
        chunk = torch.from_numpy(np.array(mm[a - slo:b - slo]).view(np.int16)).view(torch.bfloat16)
        if torch.device(device).type == "cuda":
            chunk = chunk.pin_memory()
        out[a - lo:b - lo].copy_(chunk, non_blocking=True)
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    return out


def load_bank_shard(bank_dir: str, rank: int, world: int, device="cuda") -> Tuple[torch.Tensor, int, int]:
    """This rank's row block of the bank (retrieval.shard_rows partition). Returns (tensor, lo, n_total)."""
    n = bank_meta(bank_dir)["n"]
    lo, hi = shard_rows(n, world, rank)
    return load_bank_rows(bank_dir, lo, hi, device), lo, n


def load_paths(bank_dir: str) -> List[str]:
    with open(os.path.join(bank_dir, "paths.txt")) as f:
        return [ln.rstrip("\n") for ln in f]


def build_bank(model, images: Iterable, paths: Sequence[str], out_dir: str, shards: int = 1, batch: int = 64) -> None:
    """The bank BUILDER (scripts/extract_img_embs.py:16-43): image -> HF feature extractor -> CLIP ViT-L/14 ->
    `visual_fc` (`get_visual_embs(mode="retrieval")`) -> one 256-d embedding per image, then the load-time preparation of
    gill/models.py:895-900 (cast, row-normalise, x exp(logit_scale)) and the flat sharded layout of this module.

    `model` is a gill_b200.models.GILL with a CLIP tower; `images` yields PIL images or uint8 HWC arrays / tensors in the
    order of `paths`. Everything after the JPEG decode runs on the device, `batch` images of equal size at a time
    (the reference pushes one image at a time through the host-side feature extractor)."""
    import numpy as np

    from . import ops, retrieval

    m = model.model
    if m.visual_model is None:
        raise ValueError("build_bank needs a CLIP vision tower (GILL(..., visual_model=CLIPVisionB200(...)))")
    dev, dt = m.lm.dev, m.lm.dt
    embs: List[torch.Tensor] = []
    pend: List[torch.Tensor] = []

    def flush():
        if not pend:
            return
        u8 = torch.stack(pend).to(dev).contiguous()
        px = ops.clip_preprocess_u8(u8, 224, out_dtype=dt, mode="feature_extractor")      # scripts/extract_img_embs.py:37
        embs.append(m.get_visual_embs(px, mode="retrieval")[:, 0, :].float().cpu())       # :39-40
        pend.clear()

    n = 0
    for img in images:
        a = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.asarray(
            img.convert("RGB") if hasattr(img, "convert") else img).copy())
        if a.dtype != torch.uint8 or a.dim() != 3 or a.shape[-1] != 3:
            raise ValueError(f"image {n}: expected uint8 [H,W,3], got {a.dtype} {tuple(a.shape)}")
        if pend and (a.shape != pend[0].shape or len(pend) == batch):
            flush()
        pend.append(a)
        n += 1
    flush()
    if n != len(paths):
        raise ValueError(f"{n} images for {len(paths)} paths")
    emb = torch.cat(embs, 0)
    # the reference prepares the bank in the model dtype, bf16 after load_gill's model.bfloat16() (models.py:876, :896-899)
    bank = retrieval.prepare_bank(emb.numpy(), m.logit_scale.detach().to(torch.bfloat16))
    save_prepared_bank(bank.cpu(), list(paths), out_dir, shards=shards)
