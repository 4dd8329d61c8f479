"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of GILL's retrieval branch:
  bank preparation  gill/models.py:895-900   (cast to model dtype, row-normalise, multiply by exp(logit_scale))
  query             gill/models.py:673-675   (retrieval head on the 8 [IMG] hiddens, keep token 0, L2-normalise, cast)
  scoring + top-k   gill/models.py:676-683   (bank @ q.T, -1000 on already-returned rows, topk)

Determinism: the reference leaves tie order to torch.topk and rounds scores to bf16 (many ties). The oracle defines
scores as the fp32 accumulation of the bf16 operands and breaks ties by LOWEST global row index; tier-A fixtures use
values k/8 so every partial sum is exact in fp32 and the result is order-independent, i.e. truly bit-exact.
"""
from typing import Optional, Sequence, Tuple

import numpy as np
import torch


def prepare_bank(emb: np.ndarray, logit_scale: torch.Tensor) -> torch.Tensor:
    """gill/models.py:896-899 in the model dtype (bf16 when loaded by load_gill)."""
    ls = logit_scale.exp()
    m = torch.tensor(emb, dtype=ls.dtype)
    m = m / m.norm(dim=1, keepdim=True)
    return ls * m


def normalize_query(ret_emb: torch.Tensor, dtype) -> torch.Tensor:
    """gill/models.py:674-675."""
    return (ret_emb / ret_emb.norm(dim=-1, keepdim=True)).type(dtype)


def scores_fp32(bank: torch.Tensor, q: torch.Tensor, chunk: int = 262144) -> torch.Tensor:
    """[N,D] x [Q,D] -> [Q,N] fp32 (inputs are taken at their stored 16-bit values)."""
    q32 = q.float()
    out = torch.empty(q.shape[0], bank.shape[0], dtype=torch.float32)
    for s in range(0, bank.shape[0], chunk):
        out[:, s : s + chunk] = q32 @ bank[s : s + chunk].float().T
    return out


def topk_lowest_index(scores: torch.Tensor, k: int, index_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row-wise top-k, descending value, ties -> lowest index. scores: [Q,N] fp32."""
    # stable descending sort == ties keep ascending index order
    vals, idx = torch.sort(scores, dim=1, descending=True, stable=True)
    return vals[:, :k].contiguous(), (idx[:, :k] + index_base).contiguous()


def retrieval_topk(bank: torch.Tensor, q: torch.Tensor, k: int, exclude_idx: Optional[Sequence[int]] = None,
                   index_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """gill/models.py:676-683 for a batch of queries. Returns (values [Q,k] fp32, global indices [Q,k] int64)."""
    s = scores_fp32(bank, q)
    if exclude_idx is not None:
        for e in exclude_idx:
            if index_base <= e < index_base + bank.shape[0]:
                s[:, e - index_base] -= 1000.0            # models.py:679-680
    return topk_lowest_index(s, k, index_base)


def merge_topk(vals: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge per-shard candidates [R,Q,K] -> [Q,k]; order by (value desc, index asc)."""
    R, Q, K = vals.shape
    v = vals.permute(1, 0, 2).reshape(Q, R * K)
    i = idx.permute(1, 0, 2).reshape(Q, R * K)
    # sort by index asc first, then stable by value desc => (value desc, index asc)
    o1 = torch.argsort(i, dim=1, stable=True)
    v, i = torch.gather(v, 1, o1), torch.gather(i, 1, o1)
    o2 = torch.argsort(v, dim=1, descending=True, stable=True)
    return torch.gather(v, 1, o2)[:, :k].contiguous(), torch.gather(i, 1, o2)[:, :k].contiguous()


# ---- synthetic banks: one definition, shared with bench.py's product arm through harness/ ----
from harness.synthetic import BANK_CHUNKS, synthetic_bank_chunk, synthetic_queries  # noqa: E402,F401
