#!/usr/bin/env python
"""Benchmark of the GILL image-emission hot path on B200 (contract: see the task's bench.py section).

    python bench.py --gpus N --steps K --warmup W            # our arm (libgillb200 kernels)
    python bench.py --impl reference --gpus N ...            # reference arm: the path's CPU implementation (oracle port)

One step = one pass of the hot path over one batch of synthetic prompts: BASELINE.json configs[3]
  8 prompts x (2x4 visual-prefix embeddings + BOS + 64 token ids) -> OPT-6.7B prefill -> GILLMapper -> SD-1.5 UNet
  51 evaluations (50-step PLMS, CFG 7.5) at 512x512 -> VAE decode -> uint8 images.
Metric: images/sec (whole job, all ranks). The second half of BASELINE.json's metric (retrieval top-k QPS over a 3M-row
bank, configs[2]) is measured in the same run and reported under "retrieval".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch

BATCH = 8
N_VIS, N_TXT = 8, 64          # 2 images x 4 visual tokens; 64 text tokens (+ BOS)
UNET_GFLOP_PER_EVAL = 803.3   # BASELINE.md §2 (per sample-evaluation)
VAE_GFLOP = 2510.0
OPT_GFLOP_T81 = 1080.0        # one prefill over 73 + 8 tokens
MAPPER_GFLOP = 2.644
BANK_N, BANK_D, BANK_Q, BANK_K = 3_000_000, 768, 1024, 16


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the upper half of the samples == "under load" (the sampler also sees idle gaps)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n):
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


# ------------------------------------------------------------------------------------------------------------------
def make_inputs(rank):
    """SURVEY.md §8d C4 recipe (seeds offset by rank so that every rank has its own prompts)."""
    g = torch.Generator().manual_seed(21 + 1000 * rank)
    vis = (torch.randn(BATCH, N_VIS, 4096, generator=g) * 0.05).to(torch.bfloat16)
    g = torch.Generator().manual_seed(22 + 1000 * rank)
    ids = torch.cat([torch.full((BATCH, 1), 2, dtype=torch.int64),
                     torch.randint(3, 50265, (BATCH, N_TXT), generator=g)], 1)
    g = torch.Generator().manual_seed(42 + 1000 * rank)
    lat = torch.randn(BATCH, 4, 64, 64, generator=g).to(torch.float16)
    return vis, ids, lat


def stage_breakdown(gill, vis, ids, lat, reps=2):
    """CUDA-event time of each stage of one step (device-resident inputs): OPT prefill / GILLMapper / 51 graph-replayed
    UNet evaluations + PLMS / VAE decode. Mirrors GILL.emit_images_batch (gill_b200/models.py)."""
    m = gill.model
    dev = vis.device

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step():
        marks = [ev()]
        txt = m.input_embeddings(ids)
        embs = torch.cat([vis, txt], dim=1)
        B, P, D = embs.shape
        img_ids = torch.tensor(m.retrieval_token_idx, dtype=torch.int64, device=dev)
        img_embs = m.input_embeddings(img_ids[None, :])
        full = torch.cat([embs.to(m.lm.dt), img_embs.expand(B, -1, -1).to(m.lm.dt)], dim=1)
        from gill_b200 import models as gm
        graphed = gm.GRAPH_STAGES and hasattr(m.lm, "forward_graphed")       # as emit_images_batch runs them
        hs, lg = (m.lm.forward_graphed if graphed else m.lm.forward)(full, logit_positions=[P - 1])
        raw = hs[:, P:P + m.num_tokens, :].float().contiguous()
        marks.append(ev())
        mp = m.gen_text_hidden_fcs[0]
        gen = mp._graphed(raw, img_embs.float().contiguous()).clone() if graphed and hasattr(mp, "_graphed") \
            else mp(raw, img_embs.float())
        marks.append(ev())
        latn = gill.sd_pipe.denoise(gen, lat)
        marks.append(ev())
        gill.sd_pipe.vae.decode_u8(latn)
        marks.append(ev())
        return marks

    step()
    torch.cuda.synchronize()
    acc = [0.0] * 4
    for _ in range(reps):
        mk = step()
        torch.cuda.synchronize()
        for i in range(4):
            acc[i] += mk[i].elapsed_time(mk[i + 1]) / reps
    return {"opt_prefill": round(acc[0], 2), "gill_mapper": round(acc[1], 2), "unet_51_evals_plms": round(acc[2], 2),
            "vae_decode": round(acc[3], 2), "unet_ms_per_eval": round(acc[2] / 51, 3)}


WORKLOAD = ("BASELINE configs[3]: 8 prompts (2x4 CLIP visual-prefix tokens + BOS + 64 text tokens) -> OPT-6.7B prefill over "
            "81 tokens -> GILLMapper -> SD-1.5 UNet 51 evals (50-step PLMS, CFG 7.5) 512x512 -> VAE -> uint8")


def retrieval_section(dev, rank, world, pk, timed):
    """BASELINE configs[2]: cosine-sim + top-16 over the 3M x 768 bank (SURVEY C3's eight fixed 375 k-row chunks, so the
    bank is the same at every GPU count), 1024 queries per batch split over the ranks, bank row-sharded. At N > 1 the
    NCCL + CUDA search is first checked bit-exact against the unsharded kernel on a tier-A (exactly representable) bank.
    Rank 0 at N = 1 also times the reference's own call shape (Q = 1, K = 3; D = 256 and 768), an HBM-bound stream."""
    import torch.distributed as dist

    from gill_b200 import retrieval
    from harness import synthetic

    ql = BANK_Q // world
    checks = {}
    if world > 1:
        shard_a = synthetic.synthetic_bank_shard(BANK_N, BANK_D, world, rank, exact=True, device=dev)
        sba = retrieval.ShardedBank(shard_a, BANK_N)
        qa = synthetic.synthetic_queries(64, BANK_D, exact=True, seed=8 + rank).to(dev)
        seen = [[int(5 + j), int(BANK_N - 1 - j)] for j in range(64)]
        va, ia = (t.clone() for t in sba.search(qa, BANK_K, exclude_idx=seen))
        if rank == 0:
            full = torch.cat([shard_a] + [synthetic.synthetic_bank_shard(BANK_N, BANK_D, world, r, exact=True, device=dev)
                                          for r in range(1, world)], 0)
            vs, is_ = retrieval.retrieval_topk(full, qa, BANK_K, exclude_idx=seen)
            checks["sharded_equals_unsharded_tierA"] = bool(torch.equal(va, vs) and torch.equal(ia, is_))
            del full
        del shard_a, sba
        torch.cuda.empty_cache()
    bank = synthetic.synthetic_bank_shard(BANK_N, BANK_D, world, rank, device=dev)
    q = synthetic.synthetic_queries(BANK_Q, BANK_D)[rank * ql:(rank + 1) * ql].to(dev)
    phases, graphed = None, False
    if world > 1:
        sb = retrieval.ShardedBank(bank, BANK_N)
        search = lambda: sb.search(q, BANK_K)
        ms_eager, _ = timed(search, 10, 3)
        sb.search(q, BANK_K, record_phases=True)
        torch.cuda.synchronize()
        phases = sb.phase_ms()
        ms_ret = ms_eager
        if os.environ.get("GILLB200_RETRIEVAL_GRAPH", "1") != "0":
            try:
                g, _, _ = sb.capture(q, BANK_K)
                ms_graph, _ = timed(g.replay, 10, 3)
                graphed, ms_ret = True, min(ms_graph, ms_eager)
                checks["ms_eager"], checks["ms_graph"] = round(ms_eager, 3), round(ms_graph, 3)
            except Exception as e:  # capture of NCCL collectives not available: the eager number stands
                checks["graph_error"] = str(e)[:200]
    else:
        ws = torch.empty(1 << 26, device=dev, dtype=torch.uint8)
        search = lambda: retrieval.retrieval_topk(bank, q, BANK_K, workspace=ws)
        ms_ret, _ = timed(search, 10, 3)
    qps = BANK_Q / (ms_ret / 1e3)
    ret_flops = 2.0 * BANK_N * BANK_D * BANK_Q / world
    obj = {"metric": "retrieval top-k QPS over 3M", "value": round(qps, 1), "unit": "queries/s",
           "ms_per_batch": round(ms_ret, 3), "phases_ms": phases, "cuda_graph": graphed, "checks": checks,
           "config": {"bank": f"{BANK_N}x{BANK_D} bf16, 8 fixed seeded chunks (identical at every GPU count)",
                      "queries": BANK_Q, "k": BANK_K, "bank_shards": world,
                      "exchange": "one packed all_gather of queries+seen lists, one of candidates" if world > 1 else "none",
                      "cache": "bank shard (>= 576 MB) larger than L2"},
           "roofline": {"kernel": "topk_scores_kernel", "bound": "tensor",
                        "achieved": round(ret_flops / ms_ret / 1e9, 1), "peak": pk["tf_burst"],
                        "unit": "TFLOP/s", "frac": round(ret_flops / ms_ret / 1e9 / pk["tf_burst"], 3),
                        "traffic": traffic_db().get("topk_scores_kernel"),
                        "peak_source": pk["src"] + " burst bf16 (kernel timed alone); per-GPU share of the batch at N > 1"}}
    if rank == 0 and world == 1:
        # the reference's own call shape (gill/models.py:676-683): one query, top-3, two seen rows; HBM-bound
        small = []
        for d_ in (256, 768):
            bk = bank if d_ == BANK_D else synthetic.synthetic_bank_shard(BANK_N, d_, 1, 0, device=dev)
            q1 = synthetic.synthetic_queries(1, d_).to(dev)
            f1 = lambda: retrieval.retrieval_topk(bk, q1, 3, exclude_idx=[17, 4242], workspace=ws)
            ms1, _ = timed(f1, 20, 3)
            gbs = BANK_N * d_ * 2 / ms1 / 1e6
            small.append({"shape": f"Q=1 K=3 D={d_} N={BANK_N}", "ms": round(ms1, 4), "qps": round(1e3 / ms1, 1),
                          "roofline": {"kernel": "topk_stream_kernel", "bound": "hbm", "achieved": round(gbs, 1),
                                       "peak": pk["hbm"], "unit": "GB/s", "frac": round(gbs / pk["hbm"], 3),
                                       "traffic": traffic_db().get(f"topk_stream_kernel_d{d_}"),
                                       "algorithmic_bytes": BANK_N * d_ * 2}})
            if bk is not bank:
                del bk
        obj["reference_call_shape"] = small
    del bank
    torch.cuda.empty_cache()
    return obj


def config5_section(gill, dev, rank, world, timed):
    """BASELINE configs[4]: the full generate_for_images_and_texts surface, batched -- per rank 4 conversations (2 CLIP-
    encoded images + ~64 text tokens each, unequal lengths) -> ragged OPT prefill -> retrieval top-3 over the 3M x 256 bank
    (row-sharded across the ranks, one candidate exchange) -> decision -> GILLMapper -> 4 images per prompt (SD in chunks
    of 8, 50-step PLMS) -> CLIP ViT-L/14 re-rank on the device. images/s = all ranks' images / max-over-ranks time."""
    from gill_b200 import retrieval
    from harness import synthetic

    m = gill.model
    n_prompts, n_gen = 4, 4
    t0 = time.time()
    old = (m.visual_model, gill.emb_matrix, gill.num_gen_images)
    m.visual_model = synthetic.build_clip_tower(dev)
    shard = synthetic.synthetic_bank_shard(BANK_N, 256, world, rank, device=dev)
    sb = retrieval.ShardedBank(shard, BANK_N)
    prompts = synthetic.config5_prompts(n_prompts, 100 + 1000 * rank)
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    setup_s = time.time() - t0

    def step():
        out = gill.generate_for_images_and_texts_batch(prompts, num_gen_images=n_gen, generator=gen, bank=sb)
        return torch.stack([o[1]["gen"][0][0] for o in out]).cpu()          # best image of each prompt reaches the host

    try:
        ms, launches = timed(step, 2, 1)
        info = gill.last_batch_info
        ok = bool(all(info["forced_ok"]))
        lens = info["prompt_lens"]
    finally:
        m.visual_model, gill.emb_matrix, gill.num_gen_images = old
        del shard, sb
        torch.cuda.empty_cache()
    imgs = world * n_prompts * n_gen
    return {"metric": "images/sec, full generate_for_images_and_texts surface (BASELINE configs[4])",
            "value": round(imgs / (ms / 1e3), 4), "unit": "images/s", "ms_per_step": round(ms, 1), "gpu_launches": int(launches),
            "config": {"prompts_per_gpu": n_prompts, "images_per_prompt": n_gen, "global_images_per_step": imgs,
                       "prompt_lens_rank0": lens, "bank": f"{BANK_N}x256 bf16 sharded {world}-way, top-3 + per-prompt seen lists",
                       "rerank": "CLIP ViT-L/14-shaped tower (24 layers, seeded init) on 224x224 PIL-exact device resize",
                       "forced_emission_ok": ok, "setup_s": round(setup_s, 1)}}


_TRAFFIC = None


def traffic_db():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures."""
    global _TRAFFIC
    if _TRAFFIC is None:
        _TRAFFIC = {}
        for fn in ("r01_traffic.json", "r02_traffic.json"):
            tp = os.path.join(ROOT, "profiles", fn)
            if os.path.exists(tp):
                _TRAFFIC.update(json.load(open(tp)))
    return _TRAFFIC


def hf_eager_gpu_baseline(dev, vis, ids, lat):
    """BASELINE leg, not the product: SURVEY 8(d) "reference HF pipeline on the same B200". transformers' own
    OPTForCausalLM (eager, bf16) driven the way gill/models.py:464-530 drives it -- two no-cache passes (73 then 81
    tokens), logits of the last position copied to the CPU each step -- then the GILLMapper and the diffusers-0.17.1
    UNet / PLMS / VAE module math in plain eager PyTorch fp16 (oracle/ restatement: diffusers itself is not installable
    offline), batch 8, 51 evaluations, CFG, `.cpu()` of the fp32 images. None of this repo's kernels run here."""
    import torch.nn.functional as F
    from transformers import OPTConfig, OPTForCausalLM

    from oracle import mapper as omap, sd15 as osd

    t_build = time.time()
    hc = OPTConfig(vocab_size=50274, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32, ffn_dim=16384,
                   max_position_embeddings=2048, word_embed_proj_dim=4096, do_layer_norm_before=True,
                   activation_function="relu")
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            lm = OPTForCausalLM(hc).eval()
    finally:
        torch.set_default_dtype(prev)
    usd = {k: v.to(dev, torch.float16) for k, v in osd.init_unet(0).items()}
    vsd = {k: v.to(dev, torch.float16) for k, v in osd.init_vae_decoder(1).items()}
    msd = {k: v.to(dev, torch.bfloat16) for k, v in omap.synthetic_mapper_state_dict(1234).items()}
    neg = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(77)).to(dev, torch.float16)
    emb_table = lm.get_input_embeddings().weight
    img_ids = torch.arange(50266, 50274, device=dev)
    build_s = time.time() - t_build

    @torch.no_grad()
    def step(n_steps=50):
        embs = torch.cat([vis.to(torch.bfloat16), F.embedding(ids, emb_table)], 1)                 # (8, 73, 4096)
        hs = None
        for i in range(2):                                                                         # models.py:464
            o = lm(inputs_embeds=embs, use_cache=False, output_hidden_states=True)                 # :465
            hs = o.hidden_states[-1]
            _ = o.logits[:, -1, :].float().cpu()                                                   # :470-473
            if i == 0:
                embs = torch.cat([embs, F.embedding(img_ids, emb_table)[None].expand(embs.shape[0], -1, -1)], 1)
        raw = hs[:, 73:81]
        gen = omap.mapper_forward(msd, raw, F.embedding(img_ids, emb_table)[None])                 # layers.py:28-53 (bf16)
        latn = osd.denoise_loop(usd, gen.to(torch.float16), neg, lat.to(torch.float16), 7.5, n_steps)
        img = osd.vae_decode(vsd, latn)                                                            # custom_sd.py:385-392
        return img.permute(0, 2, 3, 1).float().cpu()

    step(2)
    torch.cuda.synchronize()
    t0 = time.time()
    out = step(50)
    torch.cuda.synchronize()
    dt = time.time() - t0
    del lm, usd, vsd
    torch.cuda.empty_cache()
    return {"value": round(BATCH / dt, 4), "unit": "images/s", "s_per_step": round(dt, 3), "steps_timed": 1,
            "build_s": round(build_s, 1), "finite": bool(torch.isfinite(out).all()),
            "what": "transformers OPTForCausalLM eager bf16 (2 no-cache passes, per-step logits .cpu()) + eager PyTorch fp16 "
                    "UNet/PLMS/VAE module math (oracle restatement of diffusers 0.17.1, bmm+softmax attention as under "
                    "torch 1.13), batch 8, fp32 images copied to the host; wall clock around one full step after a warm-up"}


def run_ours(args):
    from gill_b200 import ops
    from harness import synthetic
    from gill_b200._lib import lib
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    pk = peaks()
    t0 = time.time()
    gill, wkind = synthetic.build_gill(dev, "opt-6.7b", tiny_sd=False, with_sd=True)
    m = gill.model
    vis_h, ids_h, lat_h = make_inputs(rank)
    vis_p, ids_p, lat_p = vis_h.pin_memory(), ids_h.pin_memory(), lat_h.pin_memory()
    vis_d, ids_d, lat_d = vis_h.to(dev), ids_h.to(dev), lat_h.to(dev)
    out_host = torch.empty((BATCH, 512, 512, 3), dtype=torch.uint8).pin_memory()
    setup_s = time.time() - t0

    def hot_path(vis, ids, lat):
        """device tensors in -> uint8 images on the device"""
        txt = m.input_embeddings(ids)                                            # (B, 65, 4096) gather kernel
        embs = torch.cat([vis, txt], dim=1)                                      # (B, 73, 4096)
        return gill.emit_images_batch(embs, latents=lat)

    def step_device():
        return hot_path(vis_d, ids_d, lat_d)["images"]

    def step_e2e():
        v = vis_p.to(dev, non_blocking=True)
        i = ids_p.to(dev, non_blocking=True)
        l = lat_p.to(dev, non_blocking=True)
        img = hot_path(v, i, l)["images"]
        out_host.copy_(img, non_blocking=True)
        torch.cuda.current_stream().synchronize()                                 # the caller reads the images now
        return out_host

    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        n0, r0 = lib().gillb200_launch_count(), gill.sd_pipe.graph_replays
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        n1, r1 = lib().gillb200_launch_count(), gill.sd_pipe.graph_replays
        lpe = max(st["launches_per_eval"] for st in gill.sd_pipe._graphs.values()) if gill.sd_pipe._graphs else 0
        launches = (n1 - n0) + (r1 - r0) * lpe                                   # direct launches + graph-replayed ones
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps, launches // steps

    res = step_device()
    torch.cuda.synchronize()
    forced_ok = True
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 3))
    img_per_s = world * BATCH / (ms_dev / 1e3)
    e2e_per_s = world * BATCH / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel family, measured live with CUDA events (one eager, un-graphed UNet eval)
    roof, breakdown, stages, rooflines = None, None, None, None
    if rank == 0:
        sdp = gill.sd_pipe
        st = next(iter(sdp._graphs.values()))
        ops.PROFILE = []
        for _ in range(2):
            sdp.unet.forward(st["pair"], 5, st["ctx_kv"])
        torch.cuda.synchronize()
        prof = ops.profile_summary(ops.PROFILE)
        ops.PROFILE = None
        tot = sum(d["ms"] for d in prof.values())
        breakdown = {k: {"ms_per_eval": round(d["ms"] / 2, 3), "share": round(d["ms"] / tot, 3),
                         "launches_per_eval": d["launches"] // 2,
                         "tflops": round(d["flops"] / d["ms"] / 1e9, 1) if d["flops"] else None}
                     for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        stages = stage_breakdown(gill, vis_d, ids_d, lat_d)
        rooflines = []
        for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            if d["flops"]:
                a_ = d["flops"] / d["ms"] / 1e9
                rooflines.append({"family": k, "bound": "tensor", "achieved": round(a_, 1), "peak": pk["tf_sustained"],
                                  "unit": "TFLOP/s", "frac": round(a_ / pk["tf_sustained"], 3)})
            elif d["bytes"]:
                a_ = d["bytes"] / d["ms"] / 1e6
                rooflines.append({"family": k, "bound": "hbm", "achieved": round(a_, 1), "peak": pk["hbm"],
                                  "unit": "GB/s", "frac": round(a_ / pk["hbm"], 3)})
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        ach = top[1]["flops"] / top[1]["ms"] / 1e9
        roof = {"kernel": {"conv3x3": "gemm_kernel<BN> (implicit 3x3 conv mode)", "gemm": "gemm_kernel<BN>",
                           "attention": "attn_kernel<HD_PAD,BLOCK_KV>"}.get(top[0], top[0]),
                "bound": "tensor", "achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tf_sustained"], 3), "traffic": traffic_db().get(top[0]),
                "peak_source": pk["src"] + " sustained bf16 (kernel timed inside a long step)",
                "launches_timed": top[1]["launches"], "avg_launch_ms": round(top[1]["ms"] / top[1]["launches"], 4),
                "share_of_unet_eval": round(top[1]["ms"] / tot, 3)}

    # ---- retrieval: 3M x 768 bank row-sharded over the ranks, Q=1024, K=16 (BASELINE configs[2])
    retrieval_obj = retrieval_section(dev, rank, world, pk, timed)

    # ---- BASELINE configs[4]: the batched full surface (every rank takes part: the bank exchange is a collective)
    config5 = None
    if os.environ.get("GILLB200_BENCH_C5", "1") != "0":
        try:
            config5 = config5_section(gill, dev, rank, world, timed)
        except Exception as e:
            config5 = {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}
            if world > 1:
                raise

    # ---- GILLMapper-only forward, B=256 (BASELINE configs[1]); one captured CUDA graph replayed per step
    mapper_obj = None
    if rank == 0:
        mp = m.gen_text_hidden_fcs[0]
        gx = torch.Generator().manual_seed(1234)
        # bf16-rounded inputs held in fp32: the mapper's fp32-output mode (bf16 hi+lo operands) is the one
        # that meets the <= 1e-3 parity bar; a bf16 input tensor would select the bf16-in/bf16-out drop-in mode whose
        # output rounding alone is 1.7e-3
        x256 = torch.randn(256, 8, 4096, generator=gx).to(torch.bfloat16).float().to(dev)
        img_embs = m.input_embeddings(torch.tensor([m.retrieval_token_idx], device=dev)).float()
        for _ in range(2):
            y = mp(x256, img_embs)
        torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        n0 = lib().gillb200_launch_count()
        with ops.graph_capture(gph, dev):
            y = mp(x256, img_embs)
        mapper_launches = lib().gillb200_launch_count() - n0
        for _ in range(3):
            gph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            gph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms_map = e0.elapsed_time(e1) / 20
        from oracle import mapper as omap

        sdm = {k: v.detach().double().cpu() for k, v in mp.state_dict().items()}
        ref = omap.mapper_forward(sdm, x256[:4].double().cpu(), img_embs.double().cpu())
        relerr = ((y[:4].double().cpu() - ref).norm() / ref.norm()).item()
        t0 = time.time()
        omap.mapper_forward({k: v.float() for k, v in sdm.items()}, x256[:64].float().cpu(), img_embs.float().cpu())
        cpu_s = (time.time() - t0) * 4
        mapper_obj = {"metric": "GILLMapper forward, B=256 (8x4096 [IMG] hiddens -> 77x768)", "ms_per_batch": round(ms_map, 3),
                      "samples_per_s": round(256 / ms_map * 1e3, 1), "gpu_launches": int(mapper_launches),
                      "rel_err_vs_fp64_oracle_B4": float(f"{relerr:.3e}"), "output_dtype": "fp32 (bf16 hi+lo operands, fp32 accumulate)",
                      "roofline": {"kernel": "gemm_kernel<BN> (split-precision A)", "bound": "tensor", "achieved": round(676.8 / ms_map, 1),
                                   "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": round(676.8 / ms_map / pk["tf_burst"], 3),
                                   "note": "algorithmic 676.8 GFLOP per batch; the split-precision path issues 2x that on the tensor cores"},
                      "cpu_baseline": {"value_ms": round(cpu_s * 1e3, 1), "cores": os.cpu_count(), "kind": "port",
                                       "sample": "oracle fp32 on B=64, x4"}}

    if rank != 0:
        return
    hf_gpu = None
    if world == 1 and os.environ.get("GILLB200_BENCH_HF", "1") != "0":
        try:
            hf_gpu = hf_eager_gpu_baseline(dev, vis_d, ids_d, lat_d)
            hf_gpu["ours_over_hf_eager_e2e"] = round(e2e_per_s / hf_gpu["value"], 2)
        except Exception as e:
            hf_gpu = {"unavailable": str(e)[:300]}
    cpu = cpu_baseline_sample(bounded_s=20.0)
    flop_per_batch = BATCH * (2 * 51 * UNET_GFLOP_PER_EVAL + VAE_GFLOP + OPT_GFLOP_T81 + MAPPER_GFLOP) / 1e3  # TFLOP
    line = {
        "metric": "images/sec OPT-6.7B->GILLMapper->SD1.5(50-step)", "value": round(img_per_s, 4), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_dev, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16 (SD) / bf16 (OPT) operands, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * world,
                   "not_in_step": "safety checker (custom_sd.py:657; 162 GFLOP/image = 0.2 % of the step; built and tested, "
                                  "gill_b200.clip.SafetyCheckerB200, but the synthetic pipeline is constructed without one) and the "
                                  "PIL conversion of the uint8 images",
                   "weights": f"OPT-6.7B/UNet/VAE seeded random init (no pretrained weights offline); GILL-trained weights: {wkind}",
                   "cache": "working set per step (13.3 GB OPT + 1.7 GB UNet weights) larger than L2",
                   "parallelism": f"dp{world} (independent prompt batches per GPU; no collective on the image path)",
                   "setup_s": round(setup_s, 1), "algorithmic_tflop_per_step": round(flop_per_batch, 1)},
        "e2e": {"value": round(e2e_per_s, 4), "unit": "images/s", "ms_per_step": round(ms_e2e, 2),
                "h2d_bytes_per_step": int(vis_h.numel() * 2 + ids_h.numel() * 8 + lat_h.numel() * 2),
                "d2h_bytes_per_step": int(out_host.numel())},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "achieved_tflops_whole_step": round(flop_per_batch / (ms_dev / 1e3), 1),
        "roofline": roof, "rooflines_by_family": rooflines, "stages_ms": stages, "unet_eval_breakdown": breakdown,
        "cpu_baseline": cpu, "hf_eager_gpu": hf_gpu, "retrieval": retrieval_obj, "config5_full_surface": config5,
        "mapper": mapper_obj,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
_CPU_STATE = {}


def cpu_sample_step(n_evals=2):
    """ONE bounded sample of configs[3] on the host cores (oracle port = the path's CPU implementation, PyTorch fp32, all
    threads): `n_evals` UNet evaluations at batch 2 (the CFG pair of one image), one VAE decode, the GILLMapper at B=8 and
    one OPT-6.7B decoder layer over the reference's two no-cache passes (73 and 81 tokens, one prompt). Returns the
    measured pieces; images/s is extrapolated from them (51 evaluations and 32 layers + lm_head per image)."""
    from oracle import mapper as omap, opt as oopt, sd15 as osd

    st = _CPU_STATE
    if not st:
        torch.set_num_threads(os.cpu_count() or 1)
        st["usd"], st["vsd"] = osd.init_unet(0), osd.init_vae_decoder(1)
        st["msd"] = omap.synthetic_mapper_state_dict(1234)
        cfg = dict(oopt.opt_config("opt-6.7b"), layers=1, vocab=64, max_pos=128)   # one full-width layer; tiny tables
        st["ocfg"], st["osd"] = cfg, oopt.init_opt(cfg, seed=0)
        st["lm_head"] = torch.randn(50274, 4096, generator=torch.Generator().manual_seed(1)) * 0.02
        g = torch.Generator().manual_seed(0)
        st["x"], st["ctx"] = torch.randn(2, 4, 64, 64, generator=g), torch.randn(2, 77, 768, generator=g)
        st["z"], st["xm"] = torch.randn(1, 4, 64, 64, generator=g), torch.randn(8, 8, 4096, generator=g)
        st["e73"], st["e81"] = torch.randn(1, 73, 4096, generator=g) * 0.05, torch.randn(1, 81, 4096, generator=g) * 0.05
        with torch.no_grad():
            osd.unet_forward(st["usd"], st["x"], 981, st["ctx"])                    # warm-up (allocator, threads)
    with torch.no_grad():
        t0 = time.time()
        for _ in range(n_evals):
            osd.unet_forward(st["usd"], st["x"], 961, st["ctx"])
        t_eval = (time.time() - t0) / n_evals
        t0 = time.time()
        osd.vae_decode(st["vsd"], st["z"])
        t_vae = time.time() - t0
        t0 = time.time()
        omap.mapper_forward(st["msd"], st["xm"], torch.zeros(1, 8, 4096))
        t_map = (time.time() - t0) / 8
        t0 = time.time()
        for e in (st["e73"], st["e81"]):                                            # models.py:464-465, no KV cache
            hs, _ = oopt.opt_forward(st["osd"], st["ocfg"], e)
        t_layer2 = time.time() - t0                                                 # one layer, both passes
        t0 = time.time()
        for e in (st["e73"], st["e81"]):
            _ = e[:, -1] @ st["lm_head"].T                                           # logits of the last position
        t_head = time.time() - t0
    t_opt = 32 * t_layer2 + t_head
    per_image = 51 * t_eval + t_vae + t_map + t_opt
    return dict(per_image_s=per_image, t_eval=t_eval, t_vae=t_vae, t_map=t_map, t_opt=t_opt, n_evals=n_evals)


def cpu_baseline_sample(bounded_s=20.0):
    """cpu_baseline of our line: a bounded sample (~10-30 s of CPU work) of configs[3], extrapolated."""
    t0 = time.time()
    best = None
    while best is None or time.time() - t0 < bounded_s * 0.5:
        r = cpu_sample_step(2)
        if best is None or r["per_image_s"] < best["per_image_s"]:
            best = r
    return {"value": round(1.0 / best["per_image_s"], 5), "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": _sample_text(best)}


def _sample_text(r):
    return (f"{r['n_evals']} UNet evals at batch 2 (CFG pair of one image, {r['t_eval']:.2f} s each) + 1 VAE decode "
            f"({r['t_vae']:.2f} s) + GILLMapper B=8 ({r['t_map'] * 1e3:.0f} ms/prompt) + one OPT-6.7B decoder layer over the "
            f"reference's two no-cache passes x32 + lm_head ({r['t_opt']:.2f} s/prompt), PyTorch fp32; images/s extrapolated "
            "to 51 evals + VAE + mapper + OPT per image")


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the host cores (oracle port: the third-party
    module math is not installable), same metric / config. Each of the K timed steps is ONE bounded sample
    (cpu_sample_step); `ms_per_step` is the measured time of a sample step, `value` the images/s extrapolated from the
    best sample (stated in cpu_baseline.sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(min(args.warmup, 2)):
        cpu_sample_step(1)
    t0 = time.time()
    rs = [cpu_sample_step(2) for _ in range(args.steps)]
    ms_step = (time.time() - t0) / max(1, args.steps) * 1e3
    best = min(rs, key=lambda r: r["per_image_s"])
    v = round(1.0 / best["per_image_s"], 5)
    cpu = {"value": v, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": _sample_text(best)}
    line = {"impl": "reference", "metric": "images/sec OPT-6.7B->GILLMapper->SD1.5(50-step)", "value": v,
            "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH,
                       "note": "a timed step is one bounded CPU sample of this workload, not 8 whole images (one image is "
                               f"~{best['per_image_s']:.0f} s on these cores); value is extrapolated, see cpu_baseline.sample"},
            "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_hf_eager_gpu(args):
    """`--impl hf_eager_gpu`: the HF-eager-on-B200 baseline alone (same leg our line reports under "hf_eager_gpu")."""
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    vis, ids, lat = (t.to(dev) for t in make_inputs(0))
    print(json.dumps({"impl": "hf_eager_gpu", "metric": "images/sec OPT-6.7B->GILLMapper->SD1.5(50-step)",
                      "config": {"workload": WORKLOAD}, **hf_eager_gpu_baseline(dev, vis, ids, lat)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "hf_eager_gpu"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "hf_eager_gpu":
        run_hf_eager_gpu(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
