"""CPU stand-in for openp5_b200.model.P5B200 backed by the ORACLE (TEST INFRASTRUCTURE): same protocol (train_step,
__call__, generate, build_trie, eval_metric_sums[_filtered], state dicts), fp32 torch on the host.  It lets the -m "not gpu"
suite drive the reference's unmodified data plumbing through openp5_b200.runner.B200Runner end to end; the GPU suite runs
the same runner on the real engine and compares it with this arithmetic."""
from collections import OrderedDict

import torch

from oracle import p5_oracle as po
from openp5_b200 import runner as R


class _PathTrie:
    def __init__(self, paths):
        self.paths = [list(map(int, p)) for p in paths]
        self.trie = po.Trie(self.paths)


class OracleModel:
    def __init__(self, cfg, weights):
        self.cfg, self.w = cfg, {k: v.clone() for k, v in weights.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.w.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.w.items()}
        self.world_size, self.rank, self.training, self._opt_step = 1, 0, True, 0
        self.module = self
        self.calls = dict(train_step=0, forward=0, generate=0)

    # ---- module protocol
    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def zero_grad(self, set_to_none=False):
        pass

    def state_dict(self):
        sd = OrderedDict((k, v.clone()) for k, v in self.w.items())
        for alias in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight"):
            sd[alias] = sd["shared.weight"]
        return sd

    def load_state_dict(self, sd, strict=True):
        for k in self.w:
            self.w[k] = sd[k].clone().float()

    def optimizer_state_dict(self):
        return {"step": self._opt_step, "state": {k: {"step": self._opt_step, "exp_avg": self.m[k].clone(),
                                                      "exp_avg_sq": self.v[k].clone()} for k in self.w}}

    def load_optimizer_state_dict(self, sd):
        for k in self.w:
            self.m[k] = sd["state"][k]["exp_avg"].clone()
            self.v[k] = sd["state"][k]["exp_avg_sq"].clone()
        self._opt_step = int(sd["step"])

    # ---- hot path (oracle arithmetic; dropout 0)
    def train_step(self, input_ids, whole_word_ids, attention_mask, labels, labels_attention, *, lr, clip=1.0, betas=(0.9, 0.999),
                   eps=1e-6, weight_decay=0.01, **_):
        self.calls["train_step"] += 1
        loss, _, _, g = po.loss_and_grads(self.w, self.cfg, input_ids, whole_word_ids, attention_mask, labels, labels_attention)
        if clip > 0:
            po.clip_grad_norm(g, clip)
        self._opt_step += 1
        for k in self.w:
            po.adamw_hf426(self.w[k], g[k], self.m[k], self.v[k], self._opt_step, lr, betas[0], betas[1], eps,
                           po.adamw_weight_decay_for(k, weight_decay))
        return loss.reshape(1)

    def __call__(self, input_ids=None, whole_word_ids=None, attention_mask=None, labels=None, return_dict=True, **_):
        self.calls["forward"] += 1
        with torch.no_grad():
            lt, lg = po.forward(self.w, self.cfg, input_ids, whole_word_ids, attention_mask, labels)
        return {"loss": lt, "logits": lg}

    def build_trie(self, paths):
        return _PathTrie(paths)

    def generate(self, input_ids=None, attention_mask=None, whole_word_ids=None, max_length=50, prefix_allowed_tokens_fn=None,
                 trie=None, num_beams=1, num_return_sequences=1, **_):
        self.calls["generate"] += 1
        if trie is None:    # the reference's opaque callback: recover its Trie exactly like P5B200._trie_from_callback
            src = [c.cell_contents for c in prefix_allowed_tokens_fn.__closure__ if hasattr(c.cell_contents, "trie_dict")][0]
            trie = _PathTrie(list(iter(src)))
        with torch.no_grad():
            s, sc = po.beam_search(self.w, self.cfg, input_ids, whole_word_ids, attention_mask, trie.trie, num_beams,
                                   num_return_sequences, max_length, cached=True, on_empty="neg_inf")
        return {"sequences": s, "sequences_scores": sc}

    def eval_metric_sums(self, sequences, scores, gold, num_beams, ks, out=None):
        rel = R.rel_results(sequences.tolist(), scores.tolist(), gold.tolist(), num_beams)
        names = ["hit@%d" % k for k in ks] + ["ndcg@%d" % k for k in ks]
        t = torch.tensor(R.metric_sums(rel, names), dtype=torch.float32)
        return t if out is None else out + t

    def eval_metric_sums_filtered(self, sequences, scores, gold, rows_per_user, ks, positives, n_positives, k_cut, out=None):
        rel = R.rel_results_filtered(sequences.tolist(), scores.tolist(), gold.tolist(), rows_per_user, positives.tolist(),
                                     n_positives.tolist(), k_cut)
        names = ["hit@%d" % k for k in ks] + ["ndcg@%d" % k for k in ks]
        t = torch.tensor(R.metric_sums(rel, names), dtype=torch.float32)
        return t if out is None else out + t
