"""Retrieval branch of GILL on the libgillb200 kernels (gill/models.py:671-683 and the bank preparation :895-900).

Single GPU:  values, indices = retrieval_topk(bank, q, k, exclude_idx)
Multi GPU :  ShardedBank row-partitions the bank across the ranks of one box. The only cross-GPU
             step is the candidate exchange: all_gather(queries) -> local fused GEMM+top-k over the shard ->
             all_gather(candidates [Q,K] x (fp32, int64)) -> merge (value desc, index asc). ~200 KB per rank, i.e.
             latency-bound over NVLink/NVSwitch; no other collective exists on the hot path.
"""
from typing import Callable, Optional, Sequence, Tuple

import torch

from . import ops


def prepare_bank(emb_matrix, logit_scale: torch.Tensor) -> torch.Tensor:
    """Load-time bank preparation, exactly the reference's torch expressions (gill/models.py:896-899):
    cast to the model dtype on the model device, row-normalise, multiply by exp(logit_scale)."""
    ls = logit_scale.exp()
    m = torch.as_tensor(emb_matrix).to(device=ls.device, dtype=ls.dtype)
    m = m / m.norm(dim=1, keepdim=True)
    return (ls * m).contiguous()


def retrieval_topk(bank: torch.Tensor, q: torch.Tensor, k: int,
                   exclude_idx: Optional[Sequence[int]] = None, index_base: int = 0,
                   workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """`scores = bank @ q.T; scores[seen] -= 1000; scores.topk(k)` (gill/models.py:676-683) without materialising the
    scores. bank [N,D] bf16, q [Q,D] bf16 (already L2-normalised, models.py:674-675). Returns fp32 values [Q,k] and
    int64 global row indices [Q,k]; ties -> lowest index."""
    if not bank.is_cuda:
        raise RuntimeError("retrieval_topk runs on CUDA (sm_100a) only; there is no CPU fallback")
    ex = None
    if exclude_idx is not None and len(exclude_idx) > 0:
        ex = torch.as_tensor([int(e) for e in exclude_idx], dtype=torch.int64, device=bank.device)
    return ops.topk_scores(bank, q.to(bank.dtype).contiguous(), k, index_base=index_base, exclude_idx=ex,
                           workspace=workspace)


def shard_rows(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Row partition [lo, hi) of rank `rank`: contiguous blocks of ceil(n/world) rows (last shard may be shorter)."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


class ShardedBank:
    """Bank rows [lo, hi) of the global bank live on this rank; `search` returns the GLOBAL top-k for this rank's
    queries. `local_topk` / `merge` default to the CUDA kernels; tests on CPU (gloo) inject stand-ins to exercise the
    host-side exchange logic."""

    def __init__(self, local_bank: torch.Tensor, n_total: int, group=None,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.lo, self.hi = shard_rows(n_total, self.world, self.rank)
        if local_bank.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank}: local bank has {local_bank.shape[0]} rows, expected {self.hi - self.lo}")
        self.bank = local_bank
        self._local_topk = local_topk or (lambda bank, q, k, ex, base: retrieval_topk(bank, q, k, ex, base))
        self._merge = merge or ops.topk_merge

    def search(self, q_local: torch.Tensor, k: int, exclude_idx: Optional[Sequence[int]] = None):
        """q_local [Q_local, D] (same Q_local on every rank). Returns (values [Q_local,k], indices [Q_local,k])."""
        dist, W = self.dist, self.world
        if W == 1:
            return self._local_topk(self.bank, q_local, k, exclude_idx, self.lo)
        Ql = q_local.shape[0]
        q_all = torch.empty((W * Ql, q_local.shape[1]), dtype=q_local.dtype, device=q_local.device)
        dist.all_gather_into_tensor(q_all, q_local.contiguous(), group=self.group)           # (1) queries
        v, i = self._local_topk(self.bank, q_all, k, exclude_idx, self.lo)                   # (2) shard top-k
        Qa, Kc = v.shape
        cv = torch.empty((W * Qa, Kc), dtype=v.dtype, device=v.device)                       # rank-major concat
        ci = torch.empty((W * Qa, Kc), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(cv, v.contiguous(), group=self.group)                    # (3) candidates
        dist.all_gather_into_tensor(ci, i.contiguous(), group=self.group)
        cv, ci = cv.view(W, Qa, Kc), ci.view(W, Qa, Kc)
        mine = slice(self.rank * Ql, (self.rank + 1) * Ql)
        return self._merge(cv[:, mine].contiguous(), ci[:, mine].contiguous(), k)            # (4) owner merges
