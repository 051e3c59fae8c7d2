"""Pin for the oracle: the installed HuggingFace T5 (transformers 5.5.0 here; the reference pins 4.26.0 —
same T5 arithmetic) driven exactly the way the reference's P5_T5 drives it.  TEST INFRASTRUCTURE ONLY.

The reference's own model class (src/src_t5/model/P5_T5.py) cannot be imported under transformers 5.x
(SURVEY.md §8c), so the executable pin is: T5ForConditionalGeneration + the whole-word embedding added through
`inputs_embeds` (restating P5_T5.py:94-100), the reference loss (P5_T5.py:364-369; DistributedRunner.py:72-77),
and `generate(num_beams=K, prefix_allowed_tokens_fn=trie)` (DistributedRunner.py:361-371).
"""
from __future__ import annotations

import torch

from . import p5_oracle as po


def build_hf(cfg: po.T5Cfg, weights):
    from transformers import T5Config, T5ForConditionalGeneration
    hc = T5Config(
        vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff, num_layers=cfg.num_layers,
        num_decoder_layers=cfg.num_decoder_layers, num_heads=cfg.num_heads,
        relative_attention_num_buckets=cfg.rel_buckets, relative_attention_max_distance=cfg.rel_max_distance,
        dropout_rate=0.0, layer_norm_epsilon=cfg.ln_eps,
        feed_forward_proj="gated-gelu" if cfg.ffn_gated_gelu else "relu", tie_word_embeddings=True,
        pad_token_id=cfg.pad_id, eos_token_id=cfg.eos_id, decoder_start_token_id=cfg.decoder_start_id,
    )
    m = T5ForConditionalGeneration(hc)
    sd = {k: v.clone().float() for k, v in weights.items() if k != "encoder.whole_word_embeddings.weight"}
    sd["encoder.embed_tokens.weight"] = sd["shared.weight"]
    sd["decoder.embed_tokens.weight"] = sd["shared.weight"]
    sd["lm_head.weight"] = sd["shared.weight"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("embed_tokens" in k or "lm_head" in k for k in missing), missing
    m.eval()
    wwe = torch.nn.Embedding(cfg.whole_word_rows, cfg.d_model)
    wwe.weight.data.copy_(weights["encoder.whole_word_embeddings.weight"].float())
    return m, wwe


def hf_forward(m, wwe, input_ids, whole_word_ids, attention_mask, labels):
    """logits and un-reduced per-token CE the way P5_T5.forward produces them."""
    emb = m.shared(input_ids) + wwe(whole_word_ids)
    out = m(inputs_embeds=emb, attention_mask=attention_mask, labels=labels)
    logits = out.logits
    loss_tok = torch.nn.functional.cross_entropy(
        logits.view(-1, logits.size(-1)), labels.view(-1), ignore_index=-100, reduction="none")
    return loss_tok, logits


def hf_loss_and_grads(m, wwe, input_ids, whole_word_ids, attention_mask, labels, output_attention):
    m.zero_grad()
    wwe.zero_grad()
    loss_tok, logits = hf_forward(m, wwe, input_ids, whole_word_ids, attention_mask, labels)
    loss = po.runner_loss(loss_tok, output_attention, labels.shape[0], labels.shape[1])
    loss.backward()
    grads = {}
    for k, p in m.named_parameters():
        grads[k] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)
    grads["encoder.whole_word_embeddings.weight"] = wwe.weight.grad.detach().clone()
    return loss.detach(), loss_tok.detach(), logits.detach(), grads


def hf_generate(m, wwe, input_ids, whole_word_ids, attention_mask, trie, num_beams, num_return, max_length):
    """HF generate with the prefix trie, encoder run once with whole-word ids (P5_T5.py:519-578)."""
    with torch.no_grad():
        emb = m.shared(input_ids) + wwe(whole_word_ids)
        enc = m.encoder(inputs_embeds=emb, attention_mask=attention_mask)

        def allowed(batch_id, sent):
            return trie.get(sent.tolist())
        out = m.generate(encoder_outputs=enc, attention_mask=attention_mask, max_length=max_length,
                         prefix_allowed_tokens_fn=allowed, num_beams=num_beams, num_return_sequences=num_return,
                         output_scores=True, return_dict_in_generate=True, do_sample=False,
                         early_stopping=False, length_penalty=1.0)
    return out["sequences"], out["sequences_scores"]
