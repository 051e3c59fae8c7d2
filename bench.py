#!/usr/bin/env python
"""bench.py — the OpenP5 hot path on B200: T5-base train step (BASELINE.json configs[1]) + constrained beam-search eval.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # this repo's engine (libp5b200.so)
  python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's HF+PyTorch CPU path (oracle/hf_pin)

A "step" is one pass of the hot path over one synthetic batch: forward, runner loss, backward, (gradient all-reduce
when N > 1), clip_grad_norm_, AdamW, zero_grad  — ref src/src_t5/runner/DistributedRunner.py:63-87.
Prints ONE JSON line (rank 0).  `value` = whole-job training samples/s with inputs resident in HBM;
`e2e` = the same metric through the public Python API with pinned HOST inputs and a D2H read of the loss every step.
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

# BASELINE.json configs: [1] is the headline (the default `python bench.py`); [2] and [3] are 8-GPU jobs whose per-GPU share
# is what one rank runs (--workload ...; the driver's scaling run multiplies by N); [4] is the eval leg of every run
WORKLOADS = {
    "ml1m_base": dict(backbone="t5-base", vocab=32100, B=64, Le=256, Ld=8, n_items=3416,
                      name="T5-base train step, ML-1M-shaped synthetic sequences (BASELINE configs[1])"),
    "beauty_collab": dict(backbone="t5-base", vocab=32600, B=32, Le=256, Ld=8, n_items=12101, collab=True,
                          name="T5-base train step, Beauty-shaped, collaborative indexing (+500 <CI> tokens), global batch 256 on 8 GPUs "
                               "= 32 per GPU (BASELINE configs[2])"),
    "yelp_large": dict(backbone="t5-large", vocab=32100, B=64, Le=512, Ld=8, n_items=112394,
                       name="T5-large train step, Yelp-shaped random indexing, global batch 512 on 8 GPUs = 64 per GPU, seq_len 512 "
                            "(BASELINE configs[3])"),
}
WORKLOAD = WORKLOADS["ml1m_base"]
EVAL = dict(B=20, K=20, max_length=50, Le=256)                                       # BASELINE.json configs[4]
DIMS = {"t5-base": dict(d=768, ff=3072, H=12, N=12), "t5-large": dict(d=1024, ff=4096, H=16, N=24)}


def fwd_flops(backbone, Le, Ld, V):
    """SURVEY.md §8d algorithmic forward FLOPs of ONE sample with Le encoder tokens (dense, norms / softmax / bias ignored)"""
    q = DIMS[backbone]
    d, ff, A, N = q["d"], q["ff"], q["H"] * 64, q["N"]
    enc = N * (2 * Le * 4 * d * A + 2 * Le * 2 * d * ff + 4 * Le * Le * A)
    dec = N * (2 * Ld * 4 * d * A + 2 * Ld * 2 * d * A + 2 * Le * 2 * d * A + 2 * Ld * 2 * d * ff + 4 * Ld * Ld * A + 4 * Ld * Le * A)
    return enc + dec + 2 * Ld * d * V
# AdamW of step t on a side stream under the forward of step t+1 (what B200Runner.train_batch does; P5_BENCH_SYNC_OPT=1
# runs the optimiser in stream order instead).  Every step's update completes inside the timed region: the closing
# torch.cuda.synchronize() waits for the side stream too.
OVERLAP_OPT = os.environ.get("P5_BENCH_SYNC_OPT") is None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tflops=j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1400.0)), hbm=j.get("hbm_gbs", 6650.0),
                    source="MEASURED_PEAKS.json (sustained bf16 cuBLAS, STREAM copy)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        """keep the samples taken inside [t0, t1] (the timed region); when the region is shorter than nvidia-smi's sampling
        period, the samples of the warm-up right before it (same load) stand in"""
        inside = [x for x in self.lines if t0 <= x[0] <= t1]
        self.lines = inside if len(inside) >= 2 else [x for x in self.lines if x[0] <= t1][-8:]

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_batches(n, B, Le, Ld, vocab, items, seed0):
    from openp5_b200.synth import synth_batch
    return [synth_batch(B, Le, Ld, vocab, items, seed=seed0 + i) for i in range(n)]


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own stack (HuggingFace T5 + PyTorch) on the host CPU cores
# ----------------------------------------------------------------------------------------------------------------
def effective_cores():
    """host cores this process may really use: affinity mask capped by the cgroup CPU quota (a container can show
    128 logical CPUs and be allowed 16 — running 128 threads there is 10x slower than running 16)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = min(n, max(1, int(float(q[0]) / float(q[1]) + 0.5)))
    except Exception:
        pass
    return n


_BEST_THREADS = None


def best_threads():
    """pick the intra-op thread count that runs a T5-base-sized GEMM fastest (the reference would use torch's default,
    which oversubscribes badly on many-core hosts; we give the CPU arm its best setting)"""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    import torch
    cores = effective_cores()
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores} or {cores})
    a, b = torch.randn(2048, 768), torch.randn(768, 3072)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        (a @ b)
        t0 = time.perf_counter()
        for _ in range(5):
            (a @ b)
        t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = c, t
    _BEST_THREADS = best
    return best


def cpu_reference_train(sample_B, steps, warmup, threads=None, budget_s=90.0, w=None):
    """HF T5-base (installed transformers) driven as P5_T5 drives it, fp32, runner loss, clip, HF-4.26 AdamW.
    Stops early (after >= 1 timed step) once `budget_s` seconds have been spent."""
    import torch
    from oracle import p5_oracle as po, hf_pin   # CPU baseline leg: the one place bench.py may execute oracle/
    from openp5_b200.synth import synth_items
    torch.set_num_threads(threads or best_threads())
    t_start = time.perf_counter()
    w = w or WORKLOAD
    cfg = po.t5_cfg(w["backbone"], vocab_size=w["vocab"])
    weights = po.init_weights(cfg, seed=2023)
    m, wwe = hf_pin.build_hf(cfg, weights)
    params = {k: p for k, p in m.named_parameters()}
    params["encoder.whole_word_embeddings.weight"] = wwe.weight
    mom = {k: torch.zeros_like(p) for k, p in params.items()}
    var = {k: torch.zeros_like(p) for k, p in params.items()}
    items = synth_items(min(w["n_items"], 3416), seed=2023)
    batches = make_batches(2, sample_B, w["Le"], w["Ld"], w["vocab"], items, 100)
    times = []
    for s in range(warmup + steps):
        ids, attn, ww, labels, oattn = batches[s % len(batches)]
        t0 = time.perf_counter()
        _, _, _, grads = hf_pin.hf_loss_and_grads(m, wwe, ids, ww, attn, labels, oattn)
        grads = {k: grads[k] for k in params}
        po.clip_grad_norm(grads, 1.0)
        with torch.no_grad():
            for k, p in params.items():
                po.adamw_hf426(p.data, grads[k], mom[k], var[k], s + 1, 1e-3, eps=1e-6, weight_decay=po.adamw_weight_decay_for(k, 0.01))
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        elif time.perf_counter() - t_start > budget_s:
            times.append(dt)          # budget exhausted during warm-up: keep the one measurement we have
        if times and time.perf_counter() - t_start > budget_s:
            break
    mean = sum(times) / len(times)
    return dict(samples_per_s=sample_B / mean, s_per_step=mean, best_s=min(times), B=sample_B, steps=len(times),
                cores=torch.get_num_threads())


def cpu_reference_eval(users, K, batches, threads=None):
    """the reference's eval path on the host: HF generate(num_beams=K, prefix_allowed_tokens_fn=Trie) with the encoder run
    once with whole-word ids (ref DistributedRunner.py:344-371 through oracle/hf_pin.hf_generate), T5-base fp32, the
    3416-item ML-1M-shaped trie.  A bounded sample: `users` users per batch."""
    import torch
    from oracle import p5_oracle as po, hf_pin   # CPU baseline leg
    from openp5_b200.synth import synth_items
    torch.set_num_threads(threads or best_threads())
    w = WORKLOAD
    cfg = po.t5_cfg(w["backbone"], vocab_size=w["vocab"])
    m, wwe = hf_pin.build_hf(cfg, po.init_weights(cfg, seed=2023))
    items = synth_items(w["n_items"], seed=2023)
    trie = po.Trie(items)
    bs = make_batches(batches, users, EVAL["Le"], w["Ld"], w["vocab"], items, 5000)
    t0 = time.perf_counter()
    for b in bs:
        hf_pin.hf_generate(m, wwe, b[0], b[2], b[1], trie, K, K, EVAL["max_length"])
    dt = time.perf_counter() - t0
    return dict(items_per_s=users * K * batches / dt, users_per_s=users * batches / dt, s_per_batch=dt / batches, users=users,
                batches=batches, cores=torch.get_num_threads())


def gpu_reference_train(w, steps=3, warmup=2):
    """SURVEY §8d "the number to beat": the reference's own stack — HuggingFace T5 + PyTorch eager (library kernels: cuBLAS /
    ATen) — on THIS B200, same shapes, fp32 with TF32 matmuls and bf16 autocast.  Random-init HF model (no oracle involved),
    whole-word embedding added through inputs_embeds as P5_T5 does, un-reduced CE + runner loss, clip, fused AdamW."""
    import torch
    from transformers import T5Config, T5ForConditionalGeneration
    q = DIMS[w["backbone"]]
    hc = T5Config(vocab_size=w["vocab"], d_model=q["d"], d_kv=64, d_ff=q["ff"], num_layers=q["N"], num_decoder_layers=q["N"],
                  num_heads=q["H"], dropout_rate=0.1, feed_forward_proj="relu", tie_word_embeddings=True, pad_token_id=0,
                  eos_token_id=1, decoder_start_token_id=0)
    out = {}
    from openp5_b200.synth import synth_items
    items = synth_items(min(w["n_items"], 3416), seed=2023)
    ids, attn, ww, labels, oattn = [t.cuda() for t in make_batches(1, w["B"], w["Le"], w["Ld"], w["vocab"], items, 1000)[0]]
    for name in ("fp32_tf32", "bf16_autocast"):
        torch.manual_seed(0)
        m = T5ForConditionalGeneration(hc).cuda().train()
        wwe = torch.nn.Embedding(512, q["d"]).cuda()
        opt = torch.optim.AdamW(list(m.parameters()) + list(wwe.parameters()), lr=1e-3, eps=1e-6, weight_decay=0.01, fused=True)
        torch.backends.cuda.matmul.allow_tf32 = True

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(name == "bf16_autocast")):
                emb = m.shared(ids) + wwe(ww)
                logits = m(inputs_embeds=emb, attention_mask=attn, labels=labels).logits
            lt = torch.nn.functional.cross_entropy(logits.float().view(-1, logits.size(-1)), labels.view(-1), reduction="none")
            lm = (oattn != 0).float()
            loss = ((lt.view(labels.shape) * lm).sum(1) / lm.sum(1).clamp(min=1)).mean()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"samples_per_s": w["B"] / (ms / 1e3), "ms_per_step": ms}
        del m, wwe, opt
        torch.cuda.empty_cache()
    out["what"] = ("HuggingFace transformers %s T5ForConditionalGeneration + torch %s eager on this GPU (library kernels), padded "
                   "batch B=%d Le=%d Ld=%d, dropout 0.1, fused AdamW; %d timed steps" %
                   (__import__("transformers").__version__, torch.__version__, w["B"], w["Le"], w["Ld"], steps))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a FIXED bounded sample (4 rows of the B=64 batch per step) on every box, so that the CPU number does not move with the
    # host; the step count is what the time budget bounds
    B = 4
    w = WORKLOADS[args.workload]
    r = cpu_reference_train(B, args.steps, args.warmup, budget_s=170.0, w=w)
    out = {
        "impl": "reference", "metric": "train_samples_per_sec", "value": r["samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": w["name"], "sample": "B=%d rows of the B=%d batch per step" % (B, w["B"])},
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                         "sample": "HF transformers %s fp32 + restated P5 glue (oracle/hf_pin.py), B=%d, %d timed steps"
                                   % (w["backbone"], B, r["steps"])},
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from openp5_b200 import _lib
    from openp5_b200.model import P5B200
    from openp5_b200.synth import synth_items, random_init_
    from openp5_b200.runner import linear_schedule

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    headline = args.workload == "ml1m_base"
    B, Le, Ld = w["B"], w["Le"], w["Ld"]
    model = P5B200(w["backbone"], vocab_size=w["vocab"], device=local, precision="bf16", dropout=0.1, max_batch=B,
                   max_enc_len=Le, max_dec_len=Ld, max_beams=EVAL["K"])
    random_init_(model, seed=2023)
    if world > 1:
        model.init_data_parallel()
    if w.get("collab"):
        # collaborative indexing: item ids are paths of <CIk> tokens appended to the vocabulary (ids >= 32100), depth 2-4
        import random as _r
        rng = _r.Random(2023)
        seen, items = set(), []
        while len(items) < w["n_items"]:
            pth = tuple(rng.randrange(32100, w["vocab"]) for _ in range(rng.randrange(2, 5)))
            if pth not in seen:
                seen.add(pth)
                items.append([0, 300, 301] + list(pth) + [1])
    else:
        items = synth_items(w["n_items"], seed=2023)
    nb = 4
    host = make_batches(nb, B, Le, Ld, w["vocab"], items, 1000 + 97 * rank)     # a different shard per rank
    pinned = [tuple(t.pin_memory() for t in b) for b in host]
    resident = [tuple(t.to(dev) for t in b) for b in host]
    # encoder lengths = attention_mask.sum(1), known on the host where the collator builds the batch (4 B / sequence)
    lengths = [b[1].sum(dim=1).to(torch.int32).tolist() for b in host]
    total = args.warmup + args.steps
    sched_total, sched_warm = max(total * 4, 20), max(1, int(0.05 * max(total * 4, 20)))
    lr_at = lambda s: 1e-3 * max(linear_schedule(s, sched_warm, sched_total), 0.05)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(s):
        ids, attn, ww, labels, oattn = resident[s % nb]
        return model.train_step(ids, ww, attn, labels, oattn, lr=lr_at(s), clip=1.0, enc_lengths=lengths[s % nb],
                                overlap_optimizer=OVERLAP_OPT)

    def step_e2e(s):
        ids, attn, ww, labels, oattn = pinned[s % nb]
        loss = model.train_step(ids, ww, attn, labels, oattn, lr=lr_at(s), clip=1.0, enc_lengths=lengths[s % nb],
                                overlap_optimizer=OVERLAP_OPT)
        return loss.item()      # D2H read of the step's result (4 bytes), synchronises

    # ---------------- device-resident timing (the clock sampler runs from before the warm-up: nvidia-smi needs ~1 s to start
    # and the timed region of a short run is shorter than its sampling period)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for s in range(args.warmup):
        step_resident(s)
    barrier()
    l0 = _lib.load().p5_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.time()
    e0.record()
    for s in range(args.steps):
        loss = step_resident(args.warmup + s)
    e1.record()
    barrier()
    t_w1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = _lib.load().p5_launch_count() - l0
    clk = None
    if rank == 0:
        clocks.window(t_w0, t_w1)
        clk = clocks.stop()
    final_loss = float(loss.item())
    free_b, total_b = torch.cuda.mem_get_info()      # device memory in use after the timed steps (engine buffers + torch)
    hbm_used_gb = (total_b - free_b) / 1e9
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * args.steps / (ms / 1e3)

    # ---------------- end-to-end timing (host inputs, loss read back every step)
    for s in range(min(args.warmup, 3)):
        step_e2e(s)
    barrier()
    e0.record()
    for s in range(args.steps):
        step_e2e(s)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    h2d = sum(x.numel() * x.element_size() for x in pinned[0])

    # data-parallel sanity: after identical all-reduced updates every replica must hold the same weights
    dp_diff = None
    if world > 1:
        probe = torch.stack([p.detach().double().sum() for _, p in list(model.named_parameters())[:40]])
        gathered = [torch.zeros_like(probe) for _ in range(world)]
        dist.all_gather(gathered, probe)
        dp_diff = max(float((g - gathered[0]).abs().max() / gathered[0].abs().max().clamp_min(1e-30)) for g in gathered)

    # ---------------- eval leg (BASELINE configs[4]): constrained beam search, items-ranked/s.  Users shard across
    # ranks with no data-path collective (ref DistributedSampler in DistributedRunner.py:186): every rank ranks its own
    # users; the whole-job figure is N * users * K / max-over-ranks time.
    ev = EVAL
    eval_result = None
    if headline:
        import ctypes as C
        trie = model.build_trie(items)
        eb = make_batches(2, ev["B"], ev["Le"], Ld, w["vocab"], items, 5000 + 13 * rank)
        eres = [tuple(t.to(dev) for t in b) for b in eb]
        epin = [tuple(t.pin_memory() for t in b) for b in eb]
        model.eval()
        gen = lambda b: model.generate(input_ids=b[0], attention_mask=b[1], whole_word_ids=b[2], max_length=ev["max_length"],
                                       trie=trie, num_beams=ev["K"], num_return_sequences=ev["K"])
        for i in range(3):
            gen(eres[i % 2])
        barrier()
        n_ev = 8
        e0.record()
        for i in range(n_ev):
            gen(eres[i % 2])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / n_ev], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ev_ms = t.item()
        # roofline of the dominant kernel of the leg: the persistent decode kernel (HBM-bound weight / KV streaming)
        dms, dby, dst = C.c_float(), C.c_double(), C.c_int()
        eval_roof = None
        if _lib.load().p5_decode_last_launch(C.byref(dms), C.byref(dby), C.byref(dst)) == 0 and dms.value > 0:
            pk_hbm = peaks()["hbm"]
            gbs = dby.value / (dms.value * 1e-3) / 1e9
            etraffic = None
            try:   # DRAM bytes of ONE captured launch (ncu --set full, tools/collect_profiles_r02.sh)
                tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_decode_traffic.json")))
                etraffic = {"dram_bytes_per_launch": tj["dram_bytes"], "source": tj.get("source")}
            except Exception:  # noqa
                pass
            eval_roof = {"bound": "hbm", "kernel": "p5::decode_persistent_kernel (one cooperative launch per generate(): every decode "
                                                   "position, grid barriers between phases)",
                         "achieved": gbs, "peak": pk_hbm, "unit": "GB/s", "frac": gbs / pk_hbm, "traffic": etraffic,
                         "launch_ms": dms.value, "positions": dst.value, "algorithmic_bytes_per_launch": dby.value,
                         "share_of_batch": dms.value / ev_ms,
                         "peak_source": peaks()["source"] + " — of measured",
                         "measured": "CUDA events on the launch stream around the persistent launch of the last timed batch; algorithmic "
                                     "bytes = decoder passes x (decoder-block weights + tied LM head, bf16, + every user's cross K|V "
                                     "once); passes = positions - forced-prefix length (the prefix is decoded in ONE prefill pass) "
                                     "(DESIGN.md §3)"}
        # end to end: pinned host inputs -> generate -> sequences and scores back on the host, every batch
        barrier()
        e0.record()
        for i in range(n_ev):
            o = gen(epin[i % 2])
            o["sequences"].cpu(); o["sequences_scores"].cpu()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / n_ev], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ev_e2e_ms = t.item()
        model.train()
        eval_result = {"metric": "eval_items_ranked_per_sec", "value": world * ev["B"] * ev["K"] / (ev_ms / 1e3), "unit": "items/s",
                       "users_per_sec": world * ev["B"] / (ev_ms / 1e3), "ms_per_batch": ev_ms,
                       "e2e": {"value": world * ev["B"] * ev["K"] / (ev_e2e_ms / 1e3), "unit": "items/s", "ms_per_batch": ev_e2e_ms,
                               "h2d_bytes_per_step": sum(x.numel() * x.element_size() for x in epin[0][:3]),
                               "d2h_bytes_per_step": ev["B"] * ev["K"] * (9 * 8 + 4)},
                       "roofline": eval_roof,
                       "config": {"workload": "T5-base constrained beam search, ML-1M-shaped 3416-item trie (BASELINE configs[4])",
                                  "users_per_gpu_batch": ev["B"], "num_beams": ev["K"], "max_length": ev["max_length"], "Le": ev["Le"],
                                  "n_gpus": world,
                                  "timed": "host-visible generate() calls incl. the output-length D2H sync per batch, max over ranks"}}

    out = None
    if rank == 0:
        pk = peaks()
        # ---------------- roofline leg: per-launch CUDA events around every tcgen05 GEMM over `steps` more steps
        import ctypes as C
        lib = _lib.load()
        _lib.check(lib.p5_prof_enable(1))
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    nprof = min(args.steps, 5)
    for s in range(nprof):
        step_resident(s)
    pe1.record()
    torch.cuda.synchronize()
    if rank == 0:
        buf = C.create_string_buffer(1024)
        _lib.check(lib.p5_prof_summary(buf, 1024))
        _lib.check(lib.p5_prof_enable(0))
        prof = json.loads(buf.value.decode())
        prof_step_ms = pe0.elapsed_time(pe1) / nprof
        tot_ms = sum(v["ms"] for v in prof.values())
        tot_fl = sum(v["flops"] for v in prof.values())
        tot_n = sum(v["launches"] for v in prof.values())
        all_tf = tot_fl / (tot_ms / 1e3) / 1e12 if tot_ms > 0 else 0.0
        # dominant kernel = the 128x256 tile instantiation (97 % of the step's GEMM FLOPs); the 128x64 instantiation
        # carries the latency-bound decoder GEMMs (M = B*Ld = 512 rows) and is reported separately in per_class
        dom = prof["bn256"]
        achieved = dom["flops"] / (dom["ms"] / 1e3) / 1e12 if dom["ms"] > 0 else 0.0
        traffic = None
        try:   # DRAM bytes of ONE captured launch of that kernel (ncu --set full, tools/collect_profiles.sh)
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_gemm_traffic.json")))
            traffic = {"dram_bytes_per_launch": tj["dram_bytes"], "algorithmic_bytes_per_launch": tj.get("algorithmic_bytes"),
                       "captured_launch": tj.get("launch"), "source": "profiles/r02_gemm_ncu_full_summary.txt"}
        except Exception:  # noqa
            pass
        roofline = {
            "bound": "tensor", "kernel": "p5::gemm_tc_kernel<256, EPI> (tcgen05.mma cta_group::1 128x256x16, TMA, TMEM x2)",
            "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
            "traffic": traffic, "peak_source": pk["source"] + " — of measured",
            "launches_per_step": dom["launches"] / nprof, "avg_launch_us": 1e3 * dom["ms"] / max(dom["launches"], 1),
            "share_of_step": (dom["ms"] / nprof) / prof_step_ms,
            "all_gemm_classes": {"achieved": all_tf, "launches_per_step": tot_n / nprof, "share_of_step": (tot_ms / nprof) / prof_step_ms},
            "per_class": prof,
            "measured": "CUDA events on the launch stream around every GEMM launch over %d extra steps (algorithmic FLOPs "
                        "2*M*N*K of the rows actually computed; the event pairs serialise the launches, so PDL overlap is "
                        "off in this leg and the in-step rate is slightly higher)" % nprof,
        }
        # whole-step model FLOPs: PADDED = every sequence charged Le tokens (what a padded implementation
        # computes); VALID = the tokens that exist (the engine skips padding, so this is its algorithmic work and the
        # honest utilisation figure)
        step_s = ms / args.steps / 1e3
        f_pad = 3.0 * fwd_flops(w["backbone"], Le, Ld, w["vocab"]) * B
        f_valid = sum(3.0 * fwd_flops(w["backbone"], int(n), Ld, w["vocab"]) for lens in lengths for n in lens) / nb
        roofline["step_model_flops"] = {
            "padded_gflop_per_sample": f_pad / B / 1e9, "valid_gflop_per_sample": f_valid / B / 1e9,
            "padded_frac_of_peak": f_pad / step_s / (pk["tflops"] * 1e12),
            "valid_frac_of_peak": f_valid / step_s / (pk["tflops"] * 1e12),
            "note": "valid_frac_of_peak is the utilisation figure; padded is reported for comparison with padded baselines"}
        eval_out = eval_result
        # ---------------- the reference's own stack on THIS GPU (side number)
        gpu_ref = None
        if not args.no_gpu_reference and world == 1:
            try:
                gpu_ref = gpu_reference_train(w)
            except Exception as ex:  # noqa
                gpu_ref = {"failed": repr(ex)[:300]}
        # ---------------- CPU baseline on this box's host cores (bounded sample)
        cpu = None
        if not args.no_cpu_baseline:
            if eval_out is not None:
                try:
                    r = cpu_reference_eval(4, ev["K"], 2)
                    eval_out["cpu_baseline"] = {
                        "value": r["items_per_s"], "unit": "items/s", "cores": r["cores"], "kind": "port",
                        "sample": "HF transformers generate(num_beams=%d, prefix_allowed_tokens_fn=Trie) T5-base fp32 through "
                                  "oracle/hf_pin.hf_generate, %d users per batch (of the 20-user batch), %d batches, %.1f s per batch"
                                  % (ev["K"], r["users"], r["batches"], r["s_per_batch"])}
                except Exception as ex:  # noqa
                    eval_out["cpu_baseline"] = {"value": None, "unit": "items/s", "kind": "port", "sample": "failed: %r" % (ex,)}
            try:
                r = cpu_reference_train(4, 2, 1, budget_s=60.0, w=w)
                cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                       "sample": "HF transformers T5-base fp32 (oracle/hf_pin.py: the reference's own HF+PyTorch stack with the "
                                 "P5 glue restated), 4 rows of the B=%d batch, <=1 warm-up + %d timed train steps, %d of %d "
                                 "usable cores (fastest setting)" % (B, r["steps"], r["cores"], effective_cores())}
            except Exception as ex:  # noqa
                cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (ex,)}
        out = {
            "metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": w["name"] + " — fwd+loss+bwd+clip+AdamW%s" % ("+NCCL grad all-reduce" if world > 1 else ""),
                       "backbone": w["backbone"], "global_batch": world * B, "per_gpu_batch": B, "seq_len": Le, "dec_len": Ld,
                       "vocab": w["vocab"], "dropout": 0.1, "parallelism": "dp%d" % world,
                       "l2": "per-step working set (4 GB parameters+moments, >5 GB activations) >> 126 MB L2; no flush needed",
                       "residual_stream": "fp32", "gemm_operands": "bf16, fp32 accumulate (tcgen05/TMEM)",
                       "padding": "removed: encoder kernels run on sum(valid tokens) = %.0f%% of B*Le rows (lengths from the "
                                  "host-side collator batch)" % (100.0 * sum(map(sum, lengths)) / (nb * B * Le)),
                       "optimizer": ("AdamW of step t runs on a side stream under the forward of step t+1 (per-layer waits); "
                                     "all K updates complete inside the timed region") if OVERLAP_OPT else "stream order"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "gpu_reference": gpu_ref,
            "eval": eval_out,
            "final_loss": final_loss,
            "hbm_used_gb": round(hbm_used_gb, 2),
            "dp_replica_max_rel_diff": dp_diff,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--workload", default="ml1m_base", choices=sorted(WORKLOADS),
                    help="ml1m_base = BASELINE configs[1] (default, with the configs[4] eval leg); beauty_collab = configs[2]; "
                         "yelp_large = configs[3] (per-GPU share of the 8-GPU jobs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
