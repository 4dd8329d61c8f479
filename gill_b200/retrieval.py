"""Retrieval branch of GILL on the libgillb200 kernels (gill/models.py:671-683 and the bank preparation :895-900).

Single GPU:  values, indices = retrieval_topk(bank, q, k, exclude_idx)
Multi GPU :  ShardedBank row-partitions the bank across the ranks of one box. The only cross-GPU
             step is the candidate exchange, TWO collectives per search, each over one packed byte buffer:
               all_gather([queries | per-query seen lists]) -> local fused GEMM+top-k over the shard (writes straight
               into the send buffer) -> all_gather([fp32 values | int64 indices]) -> strided merge of this rank's own
               queries out of the receive buffer (value desc, index asc).
             ~200 KB per rank, i.e. latency-bound over NVLink/NVSwitch; every buffer is allocated once per shape and the
             whole search can be replayed as one CUDA graph. No other collective exists on the hot path.
"""
from typing import Callable, Optional, Sequence, Tuple, Union

import torch

from . import ops

Exclude = Union[None, Sequence[int], Sequence[Sequence[int]], torch.Tensor]


def prepare_bank(emb_matrix, logit_scale: torch.Tensor) -> torch.Tensor:
    """Load-time bank preparation, exactly the reference's torch expressions (gill/models.py:896-899):
    cast to the model dtype on the model device, row-normalise, multiply by exp(logit_scale)."""
    ls = logit_scale.exp()
    m = torch.as_tensor(emb_matrix).to(device=ls.device, dtype=ls.dtype)
    m = m / m.norm(dim=1, keepdim=True)
    return (ls * m).contiguous()


def exclude_tensor(exclude_idx: Exclude, n_queries: int, device, width: Optional[int] = None) -> Optional[torch.Tensor]:
    """Seen-row lists as the kernel takes them: int64 [1, n] (one list shared by all queries, the reference's
    `seen_image_idx`, gill/models.py:679) or [Q, n] (one list per query: batched prompts keep their own lists).
    Ragged per-query lists are padded with -1; `width` pads (or checks) the list length."""
    if exclude_idx is None:
        rows = []
    elif isinstance(exclude_idx, torch.Tensor):
        t = exclude_idx.to(torch.int64)
        rows = [t.tolist()] if t.dim() == 1 else t.tolist()
    else:
        seq = list(exclude_idx)
        rows = [list(map(int, r)) for r in seq] if seq and isinstance(seq[0], (list, tuple, torch.Tensor)) \
            else [list(map(int, seq))]
    if len(rows) not in (0, 1, n_queries):
        raise ValueError(f"exclude_idx holds {len(rows)} lists for {n_queries} queries (expected 1 or {n_queries})")
    n = max([len(r) for r in rows], default=0)
    if n == 0:
        return None                      # nothing seen yet: no list at all (the exchange then sends all -1 slots)
    if width is not None:
        if n > width:
            raise ValueError(f"{n} excluded rows exceed the exchange's max_exclude={width}")
        n = width
    return torch.tensor([r + [-1] * (n - len(r)) for r in rows], dtype=torch.int64, device=device)


def retrieval_topk(bank: torch.Tensor, q: torch.Tensor, k: int, exclude_idx: Exclude = None, index_base: int = 0,
                   workspace: Optional[torch.Tensor] = None, out=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """`scores = bank @ q.T; scores[seen] -= 1000; scores.topk(k)` (gill/models.py:676-683) without materialising the
    scores. bank [N,D] bf16, q [Q,D] bf16 (already L2-normalised, models.py:674-675). Returns fp32 values [Q,k] and
    int64 global row indices [Q,k]; ties -> lowest index. exclude_idx: global row ids, one list or one list per query."""
    if not bank.is_cuda:
        raise RuntimeError("retrieval_topk runs on CUDA (sm_100a) only; there is no CPU fallback")
    if q.dtype != bank.dtype:
        q = q.to(bank.dtype)
    if q.stride(-1) != 1:
        q = q.contiguous()
    ex = exclude_idx if isinstance(exclude_idx, torch.Tensor) and exclude_idx.is_cuda and exclude_idx.dim() == 2 \
        else exclude_tensor(exclude_idx, q.shape[0], bank.device)
    return ops.topk_scores(bank, q, k, index_base=index_base, exclude_idx=ex, workspace=workspace, out=out)


def shard_rows(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Row partition [lo, hi) of rank `rank`: contiguous blocks of ceil(n/world) rows (last shard may be shorter)."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def _align(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class ShardedBank:
    """Bank rows [lo, hi) of the global bank live on this rank; `search` returns the GLOBAL top-k for this rank's
    queries, each query filtered by ITS OWN seen list wherever its candidates are produced. `local_topk` / `merge`
    default to the CUDA kernels; tests on CPU (gloo) inject stand-ins to exercise the host-side exchange logic."""

    def __init__(self, local_bank: torch.Tensor, n_total: int, group=None, local_topk: Optional[Callable] = None,
                 merge: Optional[Callable] = None, max_exclude: int = 8):
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
        self.max_exclude = _align(max_exclude, 2)        # message rows stay 16-byte multiples
        self._local_topk = local_topk                    # (bank, q, k, ex [Q,E] int64, base) -> (values, indices)
        self._merge = merge                              # (cand_val [R,Q,K], cand_idx [R,Q,K], k) -> (values, indices)
        self._state = {}
        self._graphs = {}
        self.last_phase_events = None

    # ------------------------------------------------------------------------------------------------ buffers
    def _buffers(self, Ql: int, D: int, k: int, dev):
        key = (Ql, D, k)
        st = self._state.get(key)
        if st is None:
            W, E = self.world, self.max_exclude
            row = 2 * D + 8 * E                                              # [bf16 query | int64 seen list]
            if row % 16:
                raise ValueError(f"query rows of {row} bytes are not 16-byte multiples (D={D} must be a multiple of 8)")
            Qa = W * Ql
            vb = _align(Qa * k * 4, 16)
            rb = vb + Qa * k * 8
            u8 = dict(device=dev, dtype=torch.uint8)
            st = dict(row=row, vb=vb, rb=rb, msg=torch.zeros((Ql, row), **u8), gath=torch.zeros((Qa, row), **u8),
                      send=torch.zeros(rb, **u8), recv=torch.zeros(W * rb, **u8),
                      out_v=torch.empty((Ql, k), device=dev, dtype=torch.float32),
                      out_i=torch.empty((Ql, k), device=dev, dtype=torch.int64), ws=None)
            if dev.type == "cuda":
                from ._lib import lib

                st["ws"] = torch.empty(lib().gillb200_topk_workspace_bytes(Qa, self.hi - self.lo), **u8)
            self._state[key] = st
        return st

    # ------------------------------------------------------------------------------------------------ search
    def search(self, q_local: torch.Tensor, k: int, exclude_idx: Exclude = None, record_phases: bool = False):
        """q_local [Q_local, D] (same Q_local on every rank). exclude_idx: this rank's seen rows -- one list for all of
        its queries or one list per query (at most max_exclude ids). Returns (values [Q_local,k], indices [Q_local,k]);
        the returned tensors are reused by the next search of the same shape."""
        dist, W = self.dist, self.world
        dev = q_local.device
        Ql, D = q_local.shape
        if W == 1:
            ex = exclude_tensor(exclude_idx, Ql, dev)
            if self._local_topk is not None:
                return self._local_topk(self.bank, q_local, k, ex, self.lo)
            return retrieval_topk(self.bank, q_local, k, ex, self.lo)
        st = self._buffers(Ql, D, k, dev)
        E, row, vb = self.max_exclude, st["row"], st["vb"]
        Qa = W * Ql
        ev = []

        def mark():
            if record_phases and dev.type == "cuda":
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append(e)

        mark()
        # (1) one message per query: [bf16 query | its seen list]
        msg, gath = st["msg"], st["gath"]
        msg[:, :2 * D].view(q_local.dtype).copy_(q_local)
        ex = exclude_tensor(exclude_idx, Ql, dev, width=E)
        exv = msg[:, 2 * D:].view(torch.int64)
        if ex is None:
            exv.fill_(-1)
        else:
            exv.copy_(ex.expand(Ql, E))
        dist.all_gather_into_tensor(gath, msg, group=self.group)
        mark()
        # (2) shard top-k for every rank's queries, written straight into the send buffer
        q_all = gath[:, :2 * D].view(q_local.dtype)
        ex_all = gath[:, 2 * D:].view(torch.int64)
        send, recv = st["send"], st["recv"]
        sv = send[:Qa * k * 4].view(torch.float32).view(Qa, k)
        si = send[vb:].view(torch.int64).view(Qa, k)
        if self._local_topk is not None:
            v, i = self._local_topk(self.bank, q_all, k, ex_all, self.lo)
            sv.copy_(v)
            si.copy_(i)
        else:
            ops.topk_scores(self.bank, q_all, k, index_base=self.lo, exclude_idx=ex_all, workspace=st["ws"], out=(sv, si))
        mark()
        # (3) one candidate exchange
        dist.all_gather_into_tensor(recv, send, group=self.group)
        mark()
        # (4) the owner merges its own queries out of the receive buffer (list r = rank r's shard)
        mine = self.rank * Ql
        if self._merge is not None:
            r2 = recv.view(W, st["rb"])
            cv = r2[:, :Qa * k * 4].contiguous().view(torch.float32).view(W, Qa, k)[:, mine:mine + Ql]
            ci = r2[:, vb:].contiguous().view(torch.int64).view(W, Qa, k)[:, mine:mine + Ql]
            v, i = self._merge(cv.contiguous(), ci.contiguous(), k)
            st["out_v"].copy_(v)
            st["out_i"].copy_(i)
        else:
            ops.topk_merge_packed(recv, W, st["rb"], vb, mine, Ql, k, k, out=(st["out_v"], st["out_i"]))
        mark()
        if record_phases:
            self.last_phase_events = ev
        return st["out_v"], st["out_i"]

    def phase_ms(self):
        """After search(record_phases=True) and a synchronize: milliseconds of (pack + query gather, shard top-k, candidate
        gather, merge)."""
        ev = self.last_phase_events
        if not ev or len(ev) != 5:
            return None
        names = ("pack_and_query_allgather", "shard_topk_kernel", "candidate_allgather", "merge")
        return {n: round(ev[i].elapsed_time(ev[i + 1]), 4) for i, n in enumerate(names)}

    # ------------------------------------------------------------------------------------------------ CUDA graph
    def capture(self, q_static: torch.Tensor, k: int):
        """Capture one search over `q_static` (its storage is read on every replay; no seen lists) into a CUDA graph.
        Returns (graph, values, indices): refill q_static, graph.replay(), read the two output tensors."""
        key = (q_static.data_ptr(), tuple(q_static.shape), k)
        got = self._graphs.get(key)
        if got is None:
            for _ in range(2):                       # warm-up: buffers, NCCL channels, kernel attributes
                self.search(q_static, k)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with ops.graph_capture(g, q_static.device):
                v, i = self.search(q_static, k)
            got = (g, v, i)
            self._graphs[key] = got
        return got
