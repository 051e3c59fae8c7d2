"""Generate the committed golden vectors in tests/golden/ (run in the build container, NOT on the GPU box).

Sources of truth:
  (1) installed HuggingFace T5 (transformers, driven exactly as the reference's P5_T5 drives it — oracle/hf_pin.py):
      logits, per-token loss, runner loss, gradients, 3 AdamW steps, constrained beam search.
  (2) the reference's own pure-Python helpers imported from /root/reference/src/src_t5 (they import cleanly):
      utils/generation_trie.Trie.get, processor/Collator.calculate_whole_word_ids, utils/evaluate metrics.
Usage:  python tests/golden/make_golden.py
"""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import p5_oracle as po, hf_pin  # noqa: E402

REF = "/root/reference/src/src_t5"


def hf_vectors():
    torch.manual_seed(0)
    cfg = po.t5_cfg("t5-tiny", vocab_size=1200)
    w = po.init_weights(cfg, seed=11)
    items = po.synth_items(60, seed=13)
    ids, attn, ww, labels, oattn = po.synth_batch(3, 21, 8, cfg.vocab_size, items, seed=17)
    m, wwe = hf_pin.build_hf(cfg, w)
    loss, loss_tok, logits, grads = hf_pin.hf_loss_and_grads(m, wwe, ids, ww, attn, labels, oattn)
    out = dict(ids=ids.numpy(), attn=attn.numpy(), ww=ww.numpy(), labels=labels.numpy(), oattn=oattn.numpy(),
               logits=logits.numpy().astype(np.float32), loss_tok=loss_tok.numpy(), loss=np.array(loss.item(), dtype=np.float32))
    names = sorted(k for k in grads if k in w)
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([grads[k].norm().item() for k in names], dtype=np.float64)
    for k in ["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
              "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
              "encoder.final_layer_norm.weight", "decoder.block.1.layer.1.EncDecAttention.q.weight",
              "encoder.block.1.layer.1.DenseReluDense.wi.weight"]:
        out["grad::" + k] = grads[k].numpy()
    # three optimiser steps with HF arithmetic + the transformers-4.26 AdamW formula (restated: 4.26's AdamW class no
    # longer exists in transformers 5.x; its step is documented in SURVEY.md §8a-11)
    wo = {k: v.clone() for k, v in w.items()}
    mo = {k: torch.zeros_like(v) for k, v in w.items()}
    vo = {k: torch.zeros_like(v) for k, v in w.items()}
    losses = []
    for step in range(1, 4):
        mm, wwe2 = hf_pin.build_hf(cfg, wo)
        l, _, _, g = hf_pin.hf_loss_and_grads(mm, wwe2, ids, ww, attn, labels, oattn)
        g = {k: g[k] for k in wo}
        po.clip_grad_norm(g, 1.0)
        for k in wo:
            # the reference's two parameter groups (SingleRunner.py:186-205): no decay for names containing "bias"
            po.adamw_hf426(wo[k], g[k], mo[k], vo[k], step, 1e-3, eps=1e-6, weight_decay=po.adamw_weight_decay_for(k, 0.01))
        losses.append(l.item())
    out["adamw_losses"] = np.array(losses, dtype=np.float64)
    out["adamw::shared.weight[:8]"] = wo["shared.weight"][:8].numpy()
    out["adamw::encoder.block.0.layer.0.SelfAttention.q.weight"] = wo["encoder.block.0.layer.0.SelfAttention.q.weight"].numpy()
    out["adamw::decoder.final_layer_norm.weight"] = wo["decoder.final_layer_norm.weight"].numpy()
    for k in ("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
              "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"):
        out["adamw::" + k] = wo[k].numpy()        # weight_decay 0 group
    # constrained beam search
    trie = po.Trie(items)
    seqs, scores = hf_pin.hf_generate(m, wwe, ids, ww, attn, trie, 5, 5, 20)
    out["beam_sequences"] = seqs.numpy()
    out["beam_scores"] = scores.numpy()
    out["items"] = np.array(items)
    np.savez_compressed(os.path.join(HERE, "hf_t5_tiny.npz"), **out)
    meta = dict(cfg=cfg.to_dict(), weights_seed=11, items_seed=13, batch_seed=17, B=3, Le=21, Ld=8, n_items=60, K=5,
                max_length=20, transformers=__import__("transformers").__version__, torch=torch.__version__)
    json.dump(meta, open(os.path.join(HERE, "hf_t5_tiny.json"), "w"), indent=1)
    print("hf vectors written; loss", loss.item(), "beam top", seqs[0].tolist(), scores[0].item())


def reference_helpers():
    if not os.path.isdir(REF):
        print("reference tree not present; skipping helper goldens")
        return
    sys.path.insert(0, REF)
    import utils.generation_trie as gt          # ref utils/generation_trie.py
    import utils.evaluate as ev                 # ref utils/evaluate.py
    from processor.Collator import calculate_whole_word_ids  # ref processor/Collator.py:72-83
    items = po.synth_items(40, seed=21, min_digits=1, max_digits=3)
    trie = gt.Trie(items)
    probes = [[], [0], [0, 300], [0, 300, 301, 302, 303], items[0][:6], items[1][:7], items[2], [0, 5], [7]]
    trie_get = [sorted(trie.get(p)) for p in probes]
    toks = ["▁ML", "100", "K", "▁user", "_", "1", "▁item", "_", "100", "1", "</s>", "<pad>", "<pad>"]
    wwids = calculate_whole_word_ids(toks, list(range(len(toks))))
    toks2 = ["▁a", "b", "▁c", "<pad>", "▁d", "e", "</s>"]
    wwids2 = calculate_whole_word_ids(toks2, list(range(len(toks2))))
    preds = ["a", "b", "c", "d", "e", "f", "x", "y", "z", "w", "q", "r"]
    scores = [0.1, 0.9, 0.5, 0.3, 0.8, 0.2, 0.7, 0.6, 0.4, 0.05, 0.95, 0.0]
    targets = ["b", "x", "k"]
    rel = ev.rel_results(preds, targets, scores, 4)
    metrics = ev.get_metrics_results(rel, ["hit@1", "hit@3", "ndcg@3", "ndcg@4"]).tolist()
    # filtered variant (evaluate.py:6-35): 3 users x 4 returned rows, top-2 after removing each user's positives
    positive = {"u1": {"q", "a"}, "u2": {"e"}, "u3": set()}
    id2user = {0: "u1", 1: "u2", 2: "u3"}
    rel_f = ev.rel_results_filtered(positive, id2user, [0, 1, 2], 4, preds, targets, scores, 2)
    metrics_f = ev.get_metrics_results(rel_f, ["hit@1", "hit@2", "ndcg@2"]).tolist()
    json.dump(dict(items=items, probes=probes, trie_get=trie_get, tokens=toks, whole_word_ids=wwids, tokens2=toks2,
                   whole_word_ids2=wwids2, preds=preds, scores=scores, targets=targets, rel=rel, metrics=metrics,
                   positive={k: sorted(v) for k, v in positive.items()}, user_order=["u1", "u2", "u3"], rel_filtered=rel_f,
                   metrics_filtered=metrics_f),
              open(os.path.join(HERE, "reference_helpers.json"), "w"), indent=1, ensure_ascii=False)
    print("reference helper goldens written", metrics)


if __name__ == "__main__":
    hf_vectors()
    reference_helpers()
