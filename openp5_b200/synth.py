"""ML-1M-shaped synthetic inputs and random-init weights.  Pure torch-CPU generators, shared by
bench.py, the tests and the oracle so that every arm sees identical token ids.  No tokenizer / dataset / checkpoint
exists offline, hence synthetic ids and HF-distribution random weights (as BASELINE.json's north_star asks)."""
from __future__ import annotations

from typing import List, Optional

import torch

DIGIT_BASE = 1000                   # 100-token "digit" sub-vocabulary [DIGIT_BASE, DIGIT_BASE + 100)
ITEM_PREFIX = [300, 301, 302, 303]  # fixed 4-token "{dataset} item_" prefix


def synth_items(n_items: int, seed: int = 2023, min_digits: int = 3, max_digits: int = 3) -> List[List[int]]:
    """n unique trie paths [0, p1..p4, d1..dn, 1] with digit tokens from the digit sub-vocabulary.
    Default depth is uniform (3 digit tokens): with ragged depths a running beam can end in EOS while others
    continue, and transformers 5.5 raises on the resulting empty allowed-token list (4.26 produced an all -inf
    row) — ragged tries are exercised with on_empty="neg_inf" only."""
    g = torch.Generator().manual_seed(seed)
    seen, out = set(), []
    while len(out) < n_items:
        nd = int(torch.randint(min_digits, max_digits + 1, (1,), generator=g))
        digs = tuple(int(x) + DIGIT_BASE for x in torch.randint(0, 100, (nd,), generator=g))
        if digs in seen:
            continue
        seen.add(digs)
        out.append([0] + ITEM_PREFIX + list(digs) + [1])
    return out


def synth_batch(B: int, Le: int, Ld: int, vocab: int, items: Optional[List[List[int]]] = None, seed: int = 2023):
    """(input_ids, attention_mask, whole_word_ids, labels, output_attention), all int64 [B, L] — the five tensors
    the reference Collator returns (ref processor/Collator.py:8-34)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(2, vocab, (B, Le), generator=g)
    lens = torch.randint(max(1, Le // 2), Le + 1, (B,), generator=g)
    lens[0] = Le  # pad-to-longest: at least one full row
    pos = torch.arange(Le)[None, :]
    ids = torch.where(pos < lens[:, None], ids, torch.zeros_like(ids))
    ids[torch.arange(B), lens - 1] = 1
    attn = (ids != 0).long()
    new_word = (torch.rand(B, Le, generator=g) < 0.4).long()
    new_word[:, 0] = 1
    ww = torch.cumsum(new_word, dim=1).clamp(max=511) * attn
    ww[torch.arange(B), lens - 1] = 0
    labels = torch.zeros(B, Ld, dtype=torch.long)
    if items is None:
        items = synth_items(max(B, 64), seed)
    pick = torch.randint(0, len(items), (B,), generator=g)
    for b in range(B):
        path = items[int(pick[b])][1:][:Ld]
        labels[b, : len(path)] = torch.tensor(path)
    out_attn = (labels != 0).long()
    return ids, attn, ww, labels, out_attn


def hf_init_std(name: str, d_model: int, d_kv: int, inner: int, d_ff: int) -> Optional[float]:
    """std of HF T5's _init_weights (HF:models/t5/modeling_t5.py:541-593, factor 1.0); None = ones (RMSNorm gains)"""
    if name.endswith("layer_norm.weight"):
        return None
    if name == "shared.weight" or "whole_word" in name:
        return 1.0
    if name.endswith(".q.weight"):
        return (d_model * d_kv) ** -0.5
    if name.endswith(".k.weight") or name.endswith(".v.weight"):
        return d_model ** -0.5
    if name.endswith(".o.weight"):
        return inner ** -0.5
    if "relative_attention_bias" in name:
        return d_model ** -0.5
    if ".wi" in name:
        return d_model ** -0.5
    if name.endswith(".wo.weight"):
        return d_ff ** -0.5
    raise KeyError(name)


def random_init_(model, seed: int = 2023):
    """fill a P5B200's parameters in place on the device with the HF T5 init distributions"""
    g = torch.Generator(device=model.device).manual_seed(seed)
    c = model.cfg
    with torch.no_grad():
        for name, p in model.named_parameters():
            std = hf_init_std(name, c.d_model, c.d_kv, c.num_heads * c.d_kv, c.d_ff)
            if std is None:
                p.fill_(1.0)
            else:
                p.normal_(0.0, std, generator=g)
    model.mark_params_changed()
