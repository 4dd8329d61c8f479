"""CPU tests of the host-side logic: tokenizer stand-in, caption truncation, shard arithmetic and the 2-rank candidate
exchange of ShardedBank over gloo (the CUDA kernels are replaced by oracle stand-ins: this tests the plumbing only)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def test_truncate_caption_behaviour():
    from gill_b200.models import truncate_caption

    assert truncate_caption("a dog. a cat.") == "a dog."
    assert truncate_caption("\nfirst line\nsecond") == "first line\n"
    assert truncate_caption("no terminator") == "no terminator"
    assert truncate_caption("") == ""


def test_synthetic_tokenizer_contract():
    from harness.synthetic import IMG_IDS, SyntheticTokenizer

    t = SyntheticTokenizer()
    assert len(t) == 50274 and t.cls_token_id == 50265
    pre = "".join(f"[IMG{i}]" for i in range(8))
    assert t(pre, add_special_tokens=False).input_ids == IMG_IDS          # model_args.json:18-37
    ids = t("a dog", add_special_tokens=True, return_tensors="pt").input_ids
    assert ids.shape == (1, 3) and ids[0, 0] == 2
    assert t("\n", add_special_tokens=False).input_ids == [50118]
    assert "[IMG0]" in t.batch_decode(torch.tensor([[2, 50266]]))[0]


def test_shard_rows_cover_the_bank():
    from gill_b200.retrieval import shard_rows

    for n, w in ((3_000_000, 8), (10, 3), (7, 8), (1, 1)):
        spans = [shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_plms_table_is_the_51_step_schedule():
    from gill_b200.sd import plms_table

    tab = plms_table(50)
    assert len(tab) == 51 and [t for t, *_ in tab[:4]] == [981, 961, 961, 941] and tab[-1][0] == 1
    assert [m for *_, m in tab[:6]] == [0, 1, 2, 3, 4, 4]
    assert all(cs > 1.0 for _, cs, _, _ in tab)        # sqrt(a_prev / a_t) > 1 while denoising


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from gill_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GillB200Error):
        _lib.lib()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_out):
    import torch.distributed as dist

    from gill_b200.retrieval import ShardedBank, shard_rows
    from oracle import retrieval as orc

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, D, K, Ql = 1001, 64, 5, 3
    bank = orc.synthetic_bank_chunk(0, N, D, exact=True)
    lo, hi = shard_rows(N, world, rank)
    q_all = orc.synthetic_queries(world * Ql, D, exact=True)

    def local_topk(b, q, k, ex, base):
        # per-query seen lists [Q, E] (-1 = unused slot), as the CUDA kernel takes them
        vs, is_ = [], []
        for j in range(q.shape[0]):
            seen = [int(e) for e in ex[j].tolist() if e >= 0] if ex is not None else None
            v, i = orc.retrieval_topk(b, q[j:j + 1].contiguous(), k, seen, base)
            vs.append(v)
            is_.append(i)
        return torch.cat(vs), torch.cat(is_)

    sb = ShardedBank(bank[lo:hi], N, local_topk=local_topk, merge=orc.merge_topk)
    # every rank has its OWN seen lists (one per query): remote shards must filter with the owner's list
    mine = q_all[rank * Ql:(rank + 1) * Ql]
    first = orc.retrieval_topk(bank, mine, K)[1]
    seen = [[int(first[j, 0]), int(first[j, 2])] if (j + rank) % 2 == 0 else [int(first[j, 1])] for j in range(Ql)]
    v, i = sb.search(mine, K, exclude_idx=seen)
    rv = torch.empty(Ql, K)
    ri = torch.empty(Ql, K, dtype=torch.int64)
    for j in range(Ql):
        rv[j], ri[j] = (t[0] for t in orc.retrieval_topk(bank, mine[j:j + 1], K, exclude_idx=seen[j]))
    ok = bool(torch.equal(v, rv) and torch.equal(i, ri))
    # a shared list and no list at all go through the same exchange
    v2, i2 = sb.search(mine, K, exclude_idx=[0, 500, 1000])
    rv2, ri2 = orc.retrieval_topk(bank, mine, K, exclude_idx=[0, 500, 1000])
    ok = ok and bool(torch.equal(v2, rv2) and torch.equal(i2, ri2))
    v3, i3 = sb.search(mine, K)
    rv3, ri3 = orc.retrieval_topk(bank, mine, K)
    ok = ok and bool(torch.equal(v3, rv3) and torch.equal(i3, ri3))
    v, i, rv, ri = v3, i3, rv3, ri3
    q_out.put((rank, bool(torch.equal(v, rv) and torch.equal(i, ri))))
    dist.destroy_process_group()


def test_sharded_bank_exchange_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_bank_disk_format_roundtrip_and_reference_conversion(tmp_path):
    """flat prepared-bank shards. Conversion from the reference's pickle format reproduces
    retrieval.prepare_bank bit for bit; any (rank, world) row block reads back exactly; paths keep their order."""
    import pickle
    import numpy as np
    import torch
    from gill_b200 import bank as gbank, retrieval

    g = torch.Generator().manual_seed(3)
    n, d = 1003, 256
    emb = torch.randn(n, d, generator=g)
    paths = [f"http://example.invalid/{i}.jpg" for i in range(n)]
    files = []
    for j, (lo, hi) in enumerate(((0, 400), (400, n))):                    # two cc3m*.npy-style pickles
        fn = tmp_path / f"cc3m_{j}.npy"
        with open(fn, "wb") as f:
            pickle.dump({"paths": paths[lo:hi], "embeddings": [emb[i] for i in range(lo, hi)]}, f)
        files.append(str(fn))
    logit_scale = torch.tensor(2.6562, dtype=torch.bfloat16)
    out = tmp_path / "bank"
    gbank.convert_reference_bank(files, logit_scale, str(out), shards=3)
    expect = retrieval.prepare_bank(emb.numpy(), logit_scale)
    meta = gbank.bank_meta(str(out))
    assert meta["n"] == n and meta["d"] == d and meta["shards"] == 3
    assert meta["rows"] == [list(retrieval.shard_rows(n, 3, s)) for s in range(3)]
    full = gbank.load_bank_rows(str(out), 0, n, device="cpu")
    assert full.dtype == torch.bfloat16 and torch.equal(full.view(torch.int16), expect.view(torch.int16))
    for world in (1, 2, 4, 8):                                              # any world size reads its own row block
        got = []
        for r in range(world):
            t, lo, ntot = gbank.load_bank_shard(str(out), r, world, device="cpu")
            assert ntot == n and lo == retrieval.shard_rows(n, world, r)[0]
            got.append(t)
        assert torch.equal(torch.cat(got).view(torch.int16), expect.view(torch.int16))
    assert gbank.load_paths(str(out)) == paths
    import pytest
    with pytest.raises(ValueError):
        gbank.load_bank_rows(str(out), 10, n + 1, device="cpu")
    with pytest.raises(ValueError):
        gbank.save_prepared_bank(expect.float(), paths, str(out))


def test_up2_phase_weights_reproduce_upsample_then_conv_on_cpu():
    """The algebra behind ops.conv3x3_up2 (gillb200_gemm_args::conv_phase), checked on the CPU with plain torch: nearest 2x
    upsample + 3x3 / pad 1 conv == four 2 x 2 convolutions of the LOW-RES tensor with the pre-summed phase weights, output
    phase (a, b) reading low-res offsets (a - 1 + u, b - 1 + v)."""
    import torch.nn.functional as F

    from gill_b200 import ops

    torch.manual_seed(0)
    B, C, Co, H, W = 2, 5, 7, 6, 4
    x = torch.randn(B, C, H, W, dtype=torch.float64)
    w4 = torch.randn(Co, C, 3, 3, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w4, padding=1)
    wk = w4.permute(0, 2, 3, 1).reshape(Co, 9 * C)                         # k = (ky*3+kx)*C + c, as sd._conv_w lays it out
    wph = ops.conv3x3_up2_weights(wk)                                      # [4, Co, 4*C], k = (u*2+v)*C + c
    assert wph.shape == (4, Co, 4 * C)
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for a in (0, 1):
        for b in (0, 1):
            w2 = wph[a * 2 + b].view(Co, 2, 2, C).permute(0, 3, 1, 2)      # [Co, C, u, v]
            out[:, :, a::2, b::2] = F.conv2d(xp[:, :, a:a + H + 1, b:b + W + 1], w2)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)


def test_narrow_conv_tap_weights_reproduce_conv_on_cpu():
    """The algebra behind ops.conv3x3_narrow: a 3x3 / pad 1 conv == one GEMM onto the nine per-tap products followed by the
    nine-tap shifted sum of gillb200_tap_sum3x3 (restated here in torch)."""
    import torch.nn.functional as F

    from gill_b200 import ops

    torch.manual_seed(1)
    B, C, Co, H, W = 2, 6, 4, 5, 7
    x = torch.randn(B, H, W, C, dtype=torch.float64)
    w4 = torch.randn(Co, C, 3, 3, dtype=torch.float64)
    wk = w4.permute(0, 2, 3, 1).reshape(Co, 9 * C)
    wt = ops.conv3x3_taps_weight(wk, Co, pad_to=16)                        # [48, C], row tap*Co + o
    assert wt.shape == (48, C) and (wt[9 * Co:] == 0).all()
    y = (x.reshape(-1, C) @ wt.T).view(B, H, W, -1)                        # per-tap products at every pixel
    out = torch.zeros(B, H, W, Co, dtype=torch.float64)
    for dy in range(3):
        for dx in range(3):
            t = dy * 3 + dx
            ys = F.pad(y[..., t * Co:(t + 1) * Co], (0, 0, 1, 1, 1, 1))    # zero outside the image
            out += ys[:, dy:dy + H, dx:dx + W]
    ref = F.conv2d(x.permute(0, 3, 1, 2), w4, padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
