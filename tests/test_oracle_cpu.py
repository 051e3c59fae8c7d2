"""CPU tests that PIN the oracle (oracle/p5_oracle.py):
  * against the committed golden vectors generated from the installed HF T5 (tests/golden/make_golden.py),
  * against HF executed live (when transformers is importable),
  * against goldens produced by the reference's own pure-Python helpers (Trie.get, whole-word ids, metrics).
"""
import math

import numpy as np
import pytest
import torch

from oracle import p5_oracle as po


def _setup(golden):
    meta = golden["meta"]
    cfg = po.T5Cfg(**meta["cfg"])
    w = po.init_weights(cfg, seed=meta["weights_seed"])
    t = lambda k: torch.from_numpy(golden[k])
    return cfg, w, t("ids"), t("attn"), t("ww"), t("labels"), t("oattn")


def test_synthetic_inputs_are_reproducible(golden):
    meta = golden["meta"]
    items = po.synth_items(meta["n_items"], seed=meta["items_seed"])
    assert np.array_equal(np.array(items), golden["items"])
    b = po.synth_batch(meta["B"], meta["Le"], meta["Ld"], meta["cfg"]["vocab_size"], items, seed=meta["batch_seed"])
    for got, key in zip(b, ["ids", "attn", "ww", "labels", "oattn"]):
        assert np.array_equal(got.numpy(), golden[key]), key


def test_forward_matches_hf_golden(golden):
    cfg, w, ids, attn, ww, labels, oattn = _setup(golden)
    loss_tok, logits = po.forward(w, cfg, ids, ww, attn, labels)
    ref = torch.from_numpy(golden["logits"])
    assert (logits - ref).abs().max() <= 1e-4 * ref.abs().max()           # fp32: 1e-4 rel of max logit
    assert torch.allclose(loss_tok, torch.from_numpy(golden["loss_tok"]), rtol=1e-4, atol=1e-5)
    loss = po.runner_loss(loss_tok, oattn, *labels.shape)
    assert abs(loss.item() - float(golden["loss"])) < 1e-5


def test_gradients_match_hf_golden(golden):
    cfg, w, ids, attn, ww, labels, oattn = _setup(golden)
    _, _, _, grads = po.loss_and_grads(w, cfg, ids, ww, attn, labels, oattn)
    names = [str(n) for n in golden["grad_names"]]
    for n, ref_norm in zip(names, golden["grad_norms"]):
        assert abs(grads[n].norm().item() - ref_norm) <= 1e-4 * max(ref_norm, 1e-6), n
    for k in golden:
        if k.startswith("grad::"):
            ref = torch.from_numpy(golden[k])
            assert (grads[k[6:]] - ref).abs().max() <= 1e-4 * ref.abs().max() + 1e-8, k


def test_adamw_three_steps_match_golden(golden):
    cfg, w, ids, attn, ww, labels, oattn = _setup(golden)
    wo = {k: v.clone() for k, v in w.items()}
    mo = {k: torch.zeros_like(v) for k, v in w.items()}
    vo = {k: torch.zeros_like(v) for k, v in w.items()}
    for step in range(1, 4):
        l, _, _, g = po.loss_and_grads(wo, cfg, ids, ww, attn, labels, oattn)
        po.clip_grad_norm(g, 1.0)
        for k in wo:
            po.adamw_hf426(wo[k], g[k], mo[k], vo[k], step, 1e-3, eps=1e-6, weight_decay=po.adamw_weight_decay_for(k, 0.01))
        assert abs(l.item() - golden["adamw_losses"][step - 1]) < 1e-4
    for k in golden:
        if k.startswith("adamw::"):
            name = k[7:]
            ref = torch.from_numpy(golden[k])
            got = wo["shared.weight"][:8] if name == "shared.weight[:8]" else wo[name]
            assert torch.allclose(got, ref, rtol=1e-4, atol=1e-6), k


def test_adamw_known_answer():
    # one scalar, hand-computed with the transformers-4.26 formula (eps outside bias correction, decay after update)
    p, g, m, v = (torch.tensor([x]) for x in (1.0, 0.5, 0.0, 0.0))
    po.adamw_hf426(p, g, m, v, step=1, lr=0.1, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.01)
    m1, v1 = 0.05, 0.00025
    step_size = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = 1.0 - step_size * m1 / (math.sqrt(v1) + 1e-6)
    want = want - 0.1 * 0.01 * want
    assert abs(p.item() - want) < 1e-6 and abs(m.item() - m1) < 1e-8 and abs(v.item() - v1) < 1e-9


def test_linear_schedule():
    assert po.linear_schedule(0, 5, 100) == 0.0          # first optimizer step runs with lr 0
    assert po.linear_schedule(5, 5, 100) == 1.0
    assert abs(po.linear_schedule(52, 5, 100) - 48 / 95) < 1e-12
    assert po.linear_schedule(100, 5, 100) == 0.0


def test_beam_search_matches_hf_golden(golden):
    cfg, w, ids, attn, ww, labels, oattn = _setup(golden)
    meta = golden["meta"]
    trie = po.Trie(golden["items"].tolist())
    seqs, scores = po.beam_search(w, cfg, ids, ww, attn, trie, meta["K"], meta["K"], meta["max_length"])
    assert np.array_equal(seqs.numpy(), golden["beam_sequences"])
    assert np.allclose(scores.numpy(), golden["beam_scores"], rtol=0, atol=2e-5)
    # every returned sequence is a trie path
    paths = {tuple(p) for p in golden["items"].tolist()}
    for row in seqs.tolist():
        assert tuple(row[: row.index(1) + 1]) in paths


def test_oracle_matches_hf_live_other_seed():
    hf_pin = pytest.importorskip("oracle.hf_pin")
    pytest.importorskip("transformers")
    cfg = po.t5_cfg("t5-tiny", vocab_size=1150, num_heads=4, d_model=96)
    w = po.init_weights(cfg, seed=5)
    items = po.synth_items(50, seed=6)
    ids, attn, ww, labels, oattn = po.synth_batch(2, 13, 8, cfg.vocab_size, items, seed=7)
    m, wwe = hf_pin.build_hf(cfg, w)
    l_hf, lt_hf, lg_hf, g_hf = hf_pin.hf_loss_and_grads(m, wwe, ids, ww, attn, labels, oattn)
    l_o, lt_o, lg_o, g_o = po.loss_and_grads(w, cfg, ids, ww, attn, labels, oattn)
    assert (lg_hf - lg_o).abs().max() <= 1e-4 * lg_hf.abs().max()
    for k in g_o:
        assert (g_hf[k] - g_o[k]).abs().max() <= 1e-4 * g_hf[k].abs().max() + 1e-8, k
    trie = po.Trie(items)
    s_hf, sc_hf = hf_pin.hf_generate(m, wwe, ids, ww, attn, trie, 4, 4, 16)
    s_o, sc_o = po.beam_search(w, cfg, ids, ww, attn, trie, 4, 4, 16)
    assert torch.equal(s_hf, s_o) and torch.allclose(sc_hf, sc_o, atol=2e-5)


def test_gated_gelu_oracle_matches_hf_live():
    hf_pin = pytest.importorskip("oracle.hf_pin")
    pytest.importorskip("transformers")
    cfg = po.t5_cfg("t5-tiny", vocab_size=1150, ffn_gated_gelu=True)
    w = po.init_weights(cfg, seed=9)
    ids, attn, ww, labels, oattn = po.synth_batch(2, 10, 8, cfg.vocab_size, None, seed=3)
    m, wwe = hf_pin.build_hf(cfg, w)
    _, lt_hf, lg_hf, _ = hf_pin.hf_loss_and_grads(m, wwe, ids, ww, attn, labels, oattn)
    lt_o, lg_o = po.forward(w, cfg, ids, ww, attn, labels)
    assert (lg_hf - lg_o).abs().max() <= 1e-4 * lg_hf.abs().max()


def test_relative_position_bucket_matches_hf():
    tr = pytest.importorskip("transformers.models.t5.modeling_t5")
    rp = torch.arange(-600, 601)
    for bidir in (True, False):
        want = tr.T5Attention._relative_position_bucket(rp, bidirectional=bidir, num_buckets=32, max_distance=128)
        assert torch.equal(po.relative_position_bucket(rp, bidir, 32, 128), want)


def test_padding_invariance():
    # pad-to-longest must not change the valid positions (the engine pads Le to a multiple of 8 internally)
    cfg = po.t5_cfg("t5-tiny", vocab_size=1150)
    w = po.init_weights(cfg, seed=2)
    ids, attn, ww, labels, _ = po.synth_batch(2, 11, 8, cfg.vocab_size, None, seed=4)
    pad = lambda t: torch.cat([t, torch.zeros(t.shape[0], 5, dtype=t.dtype)], dim=1)
    lt1, lg1 = po.forward(w, cfg, ids, ww, attn, labels)
    lt2, lg2 = po.forward(w, cfg, pad(ids), pad(ww), pad(attn), labels)
    assert torch.allclose(lg1, lg2, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------- reference helper goldens
def test_trie_matches_reference_trie(ref_helpers):
    trie = po.Trie(ref_helpers["items"])
    for probe, want in zip(ref_helpers["probes"], ref_helpers["trie_get"]):
        assert sorted(trie.get(probe)) == want, probe
    assert len(trie) == len(ref_helpers["items"])


def test_whole_word_ids_match_reference(ref_helpers):
    assert po.calculate_whole_word_ids(ref_helpers["tokens"]) == ref_helpers["whole_word_ids"]
    assert po.calculate_whole_word_ids(ref_helpers["tokens2"]) == ref_helpers["whole_word_ids2"]
    assert ref_helpers["whole_word_ids"] == [1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 3, 0, 0]   # SURVEY §8c known answer


def test_metrics_match_reference(ref_helpers):
    rel = po.rel_results(ref_helpers["preds"], ref_helpers["targets"], ref_helpers["scores"], 4)
    assert rel == ref_helpers["rel"]
    got = [po.hit_at_k(rel, 1), po.hit_at_k(rel, 3), po.ndcg_at_k(rel, 3), po.ndcg_at_k(rel, 4)]
    assert np.allclose(got, ref_helpers["metrics"])


def test_adamw_parameter_groups_follow_the_reference_substring_rule():
    """ref SingleRunner.py:186-205: no_decay = ["bias", "LayerNorm.weight"] matched by substring against the parameter
    name: only the two relative_attention_bias tables fall into the weight_decay = 0 group"""
    cfg = po.t5_cfg("t5-tiny", vocab_size=1200)
    names = list(po.param_shapes(cfg))
    nd = [n for n in names if po.adamw_weight_decay_for(n, 0.01) == 0.0]
    assert nd == ["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
                  "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    no_decay = ["bias", "LayerNorm.weight"]                     # the reference's own expression, verbatim
    assert nd == [n for n in names if any(x in n for x in no_decay)]


def test_filtered_metrics_match_reference(ref_helpers):
    """openp5_b200.runner.rel_results_filtered (token-id paths, what the device kernel is tested against) ==
    ref utils/evaluate.py:6-35 on strings (golden produced by the reference function)"""
    from openp5_b200 import runner as R
    vocab = {s: i + 10 for i, s in enumerate(sorted(set(ref_helpers["preds"] + ref_helpers["targets"] + sum(ref_helpers["positive"].values(), []))))}
    seqs = [[0, vocab[p], 1, 0] for p in ref_helpers["preds"]]
    gold = [[vocab[t], 1, 0] for t in ref_helpers["targets"]]
    users = ref_helpers["user_order"]
    pmax = max(len(ref_helpers["positive"][u]) for u in users)
    pos = [[[vocab[x], 1] for x in ref_helpers["positive"][u]] + [[0, 0]] * (pmax - len(ref_helpers["positive"][u])) for u in users]
    npos = [len(ref_helpers["positive"][u]) for u in users]
    rel = R.rel_results_filtered(seqs, ref_helpers["scores"], gold, 4, pos, npos, 2)
    assert rel == ref_helpers["rel_filtered"]
    assert np.allclose(R.metric_sums(rel, ["hit@1", "hit@2", "ndcg@2"]), ref_helpers["metrics_filtered"])


def test_cached_beam_search_equals_uncached(golden):
    """beam_search(cached=True) (self K/V cache re-ordered by beam index, cross K/V once per user: what makes the BASELINE
    eval shape affordable on CPU) is the same function as the uncached restatement that is pinned on the HF golden"""
    cfg, w, ids, attn, ww, labels, oattn = _setup(golden)
    meta = golden["meta"]
    trie = po.Trie(golden["items"].tolist())
    seqs, scores = po.beam_search(w, cfg, ids, ww, attn, trie, meta["K"], meta["K"], meta["max_length"], cached=True)
    assert np.array_equal(seqs.numpy(), golden["beam_sequences"])
    assert np.allclose(scores.numpy(), golden["beam_scores"], rtol=0, atol=2e-5)
    items = po.synth_items(80, seed=8, min_digits=1, max_digits=3)          # ragged depths
    b = po.synth_batch(4, 16, 8, cfg.vocab_size, items, seed=9)
    a1 = po.beam_search(w, cfg, b[0], b[2], b[1], po.Trie(items), 6, 6, 20, on_empty="neg_inf")
    a2 = po.beam_search(w, cfg, b[0], b[2], b[1], po.Trie(items), 6, 6, 20, on_empty="neg_inf", cached=True)
    assert torch.equal(a1[0], a2[0]) and torch.allclose(a1[1], a2[1], atol=2e-6)


def test_bf16_emulation_is_the_same_function_up_to_operand_rounding():
    """`with po.bf16_emulation()` rounds the engine's bf16 storage points and nothing else: outputs stay within bf16
    rounding of the fp32 restatement; outside the context the restatement is bit-identical to before.  The gradient of
    the decoder's first FFN `wi` is the most rounding-sensitive tensor (ReLU units whose pre-activation changes sign
    under operand rounding flip whole rows of it): 10-25 % max-norm distance between two CORRECT implementations, which is
    why the bf16 engine is gated against the emulation and only reported against fp32 (tests/gpu_cases_model.py)."""
    cfg = po.t5_cfg("t5-small", vocab_size=2100, num_layers=2, num_decoder_layers=2)
    w = po.init_weights(cfg, seed=1)
    items = po.synth_items(300, seed=3)
    b = po.synth_batch(4, 64, 8, cfg.vocab_size, items, seed=5)
    l0, lt0, lg0, g0 = po.loss_and_grads(w, cfg, b[0], b[2], b[1], b[3], b[4])
    with po.bf16_emulation():
        l1, lt1, lg1, g1 = po.loss_and_grads(w, cfg, b[0], b[2], b[1], b[3], b[4])
    l2, lt2, lg2, g2 = po.loss_and_grads(w, cfg, b[0], b[2], b[1], b[3], b[4])
    assert torch.equal(lg0, lg2) and all(torch.allclose(g0[k], g2[k], rtol=1e-5, atol=1e-8) for k in g0)
    assert 0 < ((lg0 - lg1).abs().max() / lg0.abs().max()).item() < 1e-2
    assert abs(l0.item() - l1.item()) < 2e-3 * abs(l0.item())
    errs = {k: ((g0[k] - g1[k]).abs().max() / g0[k].abs().max().clamp_min(1e-9)).item() for k in g0}
    worst = max(errs, key=errs.get)
    assert "DenseReluDense.wi" in worst and 0.02 < errs[worst] < 0.5, (worst, errs[worst])
    assert sorted(errs.values())[len(errs) // 2] < 0.03                # the typical tensor: percent level
