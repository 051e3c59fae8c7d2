"""CPU ORACLE for the OpenP5 T5 hot path — TEST INFRASTRUCTURE, NOT THE PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product path (openp5_b200/) never does and fails loudly when the CUDA library is missing.

What this restates (plain torch on CPU, fp32 or fp64), with the reference file:line each function follows.
`ref:` paths are relative to agiresearch/OpenP5; `HF:` is the `transformers` package the reference imports
(pinned transformers==4.26.0 in src/src_t5/environment_t5.txt:2 — a third-party dependency that is NOT
vendored under the reference tree; the installed copy here is 5.5.0, same T5 arithmetic).

PINNING: tests/test_oracle_cpu.py checks this restatement against the installed HF implementation
(`oracle/hf_pin.py`: T5ForConditionalGeneration forward/backward, generate(num_beams=K) with a prefix trie)
and against the committed golden vectors in tests/golden/ that were generated from HF by
tests/golden/make_golden.py.  The reference itself ships no tests or golden vectors for this path
(SURVEY.md §4), so HF-as-executed-here is the pin.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

NEG_INF_BEAM = -1.0e9  # HF:generation/utils.py:3200-3201, :3011, :3052-3058


# ----------------------------------------------------------------------------------------------------
# bf16 OPERAND EMULATION (test infrastructure for the benchmarked precision="bf16" engine mode)
#
# The fp32 restatement below IS the reference arithmetic (the reference runs fp32, no autocast).  The B200 engine's
# throughput mode stores GEMM operands as bf16 (weights, normed activations, q/k/v, probabilities, FFN hidden, every
# activation gradient) and keeps accumulators, softmax statistics, the residual stream, logits and weight gradients
# in fp32.  Comparing that mode with the fp32 oracle measures bf16 rounding, not bugs; under
# `with bf16_emulation():` the SAME restatement rounds to bf16 at exactly those storage points (values in the forward,
# gradients in the backward), so the engine can be gated tightly (<= 1e-2) on every tensor at full benchmark shapes.
# ----------------------------------------------------------------------------------------------------
_EMU = False


class _RoundAct(torch.autograd.Function):
    """a bf16-stored activation: value rounded in the forward, its gradient rounded in the backward"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundFwd(torch.autograd.Function):
    """bf16 shadow of an fp32 tensor (weights; enc_out, whose gradient is accumulated in fp32)"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """an fp32 GEMM output whose incoming gradient is cast to bf16 before the dgrad / wgrad GEMMs consume it"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _ra(x):
    return _RoundAct.apply(x) if _EMU else x


def _rw(x):
    return _RoundFwd.apply(x) if _EMU else x


def _rg(x):
    return _RoundGrad.apply(x) if _EMU else x


class bf16_emulation:
    """context manager: run the restatement with the engine's bf16 storage points rounded (see above)"""

    def __enter__(self):
        global _EMU
        self.prev = _EMU
        _EMU = True
        return self

    def __exit__(self, *a):
        global _EMU
        _EMU = self.prev


# ----------------------------------------------------------------------------------------------------
# configuration (HF T5Config subset used by the reference: ref:src/src_t5/main.py:176-184)
# ----------------------------------------------------------------------------------------------------
@dataclass
class T5Cfg:
    vocab_size: int = 32100
    d_model: int = 512
    d_kv: int = 64
    d_ff: int = 2048
    num_layers: int = 6
    num_decoder_layers: int = 6
    num_heads: int = 8
    rel_buckets: int = 32
    rel_max_distance: int = 128
    ffn_gated_gelu: bool = False
    whole_word_rows: int = 512  # ref:src/src_t5/model/P5_T5.py:64-66
    dropout: float = 0.1
    ln_eps: float = 1e-6
    pad_id: int = 0
    eos_id: int = 1
    decoder_start_id: int = 0

    @property
    def inner(self) -> int:
        return self.num_heads * self.d_kv

    def to_dict(self):
        return asdict(self)


def t5_cfg(name: str, vocab_size: int = 32100, **kw) -> T5Cfg:
    """Model dimensions of the backbones the reference accepts (--backbone t5-*, SingleRunner.py:30)."""
    dims = {
        "t5-tiny": dict(d_model=64, d_ff=128, num_layers=2, num_decoder_layers=2, num_heads=2),  # test-only
        "t5-small": dict(d_model=512, d_ff=2048, num_layers=6, num_decoder_layers=6, num_heads=8),
        "t5-base": dict(d_model=768, d_ff=3072, num_layers=12, num_decoder_layers=12, num_heads=12),
        "t5-large": dict(d_model=1024, d_ff=4096, num_layers=24, num_decoder_layers=24, num_heads=16),
    }[name]
    dims.update(kw)
    return T5Cfg(vocab_size=vocab_size, **dims)


# ----------------------------------------------------------------------------------------------------
# parameters: HF state_dict key names so engine <-> oracle <-> HF exchange plain dicts
# ----------------------------------------------------------------------------------------------------
def param_shapes(cfg: T5Cfg) -> Dict[str, Tuple[int, ...]]:
    d, A, ff, H = cfg.d_model, cfg.inner, cfg.d_ff, cfg.num_heads
    s: Dict[str, Tuple[int, ...]] = {"shared.weight": (cfg.vocab_size, d)}
    s["encoder.whole_word_embeddings.weight"] = (cfg.whole_word_rows, d)

    def attn(prefix, rel):
        s[prefix + ".q.weight"] = (A, d)
        s[prefix + ".k.weight"] = (A, d)
        s[prefix + ".v.weight"] = (A, d)
        s[prefix + ".o.weight"] = (d, A)
        if rel:
            s[prefix + ".relative_attention_bias.weight"] = (cfg.rel_buckets, H)

    def ffn(prefix):
        if cfg.ffn_gated_gelu:
            s[prefix + ".DenseReluDense.wi_0.weight"] = (ff, d)
            s[prefix + ".DenseReluDense.wi_1.weight"] = (ff, d)
        else:
            s[prefix + ".DenseReluDense.wi.weight"] = (ff, d)
        s[prefix + ".DenseReluDense.wo.weight"] = (d, ff)

    for i in range(cfg.num_layers):
        b = f"encoder.block.{i}.layer"
        attn(b + ".0.SelfAttention", i == 0)
        s[b + ".0.layer_norm.weight"] = (d,)
        ffn(b + ".1")
        s[b + ".1.layer_norm.weight"] = (d,)
    s["encoder.final_layer_norm.weight"] = (d,)
    for i in range(cfg.num_decoder_layers):
        b = f"decoder.block.{i}.layer"
        attn(b + ".0.SelfAttention", i == 0)
        s[b + ".0.layer_norm.weight"] = (d,)
        attn(b + ".1.EncDecAttention", False)
        s[b + ".1.layer_norm.weight"] = (d,)
        ffn(b + ".2")
        s[b + ".2.layer_norm.weight"] = (d,)
    s["decoder.final_layer_norm.weight"] = (d,)
    return s


def init_weights(cfg: T5Cfg, seed: int = 2023, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init weights with the HF T5 distributions (HF:models/t5/modeling_t5.py:541-593, factor 1.0);
    whole-word embeddings keep nn.Embedding's N(0,1) (ref P5_T5.py:64-67)."""
    g = torch.Generator().manual_seed(seed)
    d, A, ff = cfg.d_model, cfg.inner, cfg.d_ff
    w: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("layer_norm.weight"):
            t = torch.ones(shape)
        else:
            if name == "shared.weight" or "whole_word" in name:
                std = 1.0
            elif name.endswith(".q.weight"):
                std = (d * cfg.d_kv) ** -0.5
            elif name.endswith(".k.weight") or name.endswith(".v.weight"):
                std = d ** -0.5
            elif name.endswith(".o.weight"):
                std = A ** -0.5
            elif "relative_attention_bias" in name:
                std = d ** -0.5
            elif ".wi" in name:
                std = d ** -0.5
            elif name.endswith(".wo.weight"):
                std = ff ** -0.5
            else:
                raise KeyError(name)
            t = torch.randn(shape, generator=g) * std
        w[name] = t.to(dtype)
    return w


# ----------------------------------------------------------------------------------------------------
# T5 arithmetic
# ----------------------------------------------------------------------------------------------------
def relative_position_bucket(rel_pos: torch.Tensor, bidirectional: bool, num_buckets: int, max_distance: int):
    """HF:models/t5/modeling_t5.py:189-235 (_relative_position_bucket); rel_pos = key_pos - query_pos."""
    buckets = torch.zeros_like(rel_pos)
    if bidirectional:
        num_buckets //= 2
        buckets = buckets + (rel_pos > 0).to(torch.long) * num_buckets
        rel_pos = rel_pos.abs()
    else:
        rel_pos = -torch.min(rel_pos, torch.zeros_like(rel_pos))
    max_exact = num_buckets // 2
    is_small = rel_pos < max_exact
    large = max_exact + (
        torch.log(rel_pos.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rel_pos, large)


def compute_bias(table: torch.Tensor, q_len: int, k_len: int, bidirectional: bool, cfg: T5Cfg,
                 q_offset: int = 0) -> torch.Tensor:
    """HF:models/t5/modeling_t5.py:237-251 (compute_bias) -> [1, H, q_len, k_len]."""
    ctx = torch.arange(q_len)[:, None] + q_offset
    mem = torch.arange(k_len)[None, :]
    b = relative_position_bucket(mem - ctx, bidirectional, cfg.rel_buckets, cfg.rel_max_distance)
    return table[b].permute(2, 0, 1).unsqueeze(0)


def rms_norm(x, w, eps, fp32_grad=False):
    """HF:models/t5/modeling_t5.py:55-70 (T5LayerNorm): no mean subtraction, no bias.
    (emulation: the normed activation is a bf16 GEMM operand; `fp32_grad` = its gradient is kept in fp32: enc_out)"""
    var = x.pow(2).mean(-1, keepdim=True)
    n = w * (x * torch.rsqrt(var + eps))
    return _rw(n) if fp32_grad else _ra(n)


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class _EmuFusedAttn(torch.autograd.Function):
    """bf16 emulation of the FUSED encoder self-attention (csrc/fattn.cu forward, csrc/fattn_bwd.cu backward, Le <= 256):
    the backward recomputes the normalised P from the saved row statistic, takes delta_i = sum_c dO_ic O_ic from the
    bf16-stored forward output (not from sum_j P_ij dP_ij), and feeds P and dS to the tensor cores as bf16."""

    @staticmethod
    def forward(ctx, q, k, v, bias):
        bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
        s = (q @ k.transpose(-1, -2) + bias).float()
        m = s.max(dim=-1, keepdim=True)[0]
        pu = torch.exp(s - m)
        l = pu.sum(dim=-1, keepdim=True)
        o = (bf(pu) @ v) / l
        ctx.save_for_backward(q, k, v, s, m + torch.log(l), bf(o))
        return o

    @staticmethod
    def backward(ctx, do):
        bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
        q, k, v, s, lse, o_r = ctx.saved_tensors
        p = torch.exp(s - lse)
        dp = do @ v.transpose(-1, -2)
        delta = (do * o_r).sum(dim=-1, keepdim=True)
        ds = p * (dp - delta)
        ds16 = bf(ds)
        dv = bf(p).transpose(-1, -2) @ do
        dq = ds16 @ k
        dk = ds16.transpose(-1, -2) @ q
        return dq, dk, dv, ds


def _attn_core(q, k, v, bias, fused_bwd=False):
    """softmax(q k^T + bias) v on [.., H, L, d_kv] tensors: unscaled scores, fp32 softmax (HF:modeling_t5.py:308-334)"""
    if _EMU and fused_bwd:
        return _EmuFusedAttn.apply(q, k, v, bias.expand(q.shape[0], q.shape[1], q.shape[2], k.shape[2]))
    scores = _rg(q @ k.transpose(-1, -2) + bias)      # emulation: the score gradient dS is a bf16 MMA operand in the backward
    if _EMU:
        # engine: fp32 scores and statistics; UN-normalised probabilities rounded to bf16 for the P.V product, the
        # fp32 row sum divides the fp32 accumulator afterwards (fattn.cu / dattn.cu)
        sf = scores.float()
        pu = torch.exp(sf - sf.max(dim=-1, keepdim=True)[0])
        return (_ra(pu) @ v) / pu.sum(dim=-1, keepdim=True)
    p = torch.softmax(scores.float(), dim=-1).to(scores.dtype)
    return p @ v


def attention(w, prefix, x_q, x_kv, bias, cfg: T5Cfg, trace=None, fused_bwd=False):
    """HF:models/t5/modeling_t5.py:253-344 (T5Attention.forward): NO 1/sqrt(d) scaling; softmax in fp32;
    `bias` already contains position bias + additive mask."""
    B, Lq, _ = x_q.shape
    Lk = x_kv.shape[1]
    H, dk = cfg.num_heads, cfg.d_kv
    q = _ra(x_q @ _rw(w[prefix + ".q.weight"]).T).view(B, Lq, H, dk).transpose(1, 2)
    k = _ra(x_kv @ _rw(w[prefix + ".k.weight"]).T).view(B, Lk, H, dk).transpose(1, 2)
    v = _ra(x_kv @ _rw(w[prefix + ".v.weight"]).T).view(B, Lk, H, dk).transpose(1, 2)
    ctx = _attn_core(q, k, v, bias, fused_bwd)
    ctx = _ra(ctx.transpose(1, 2).reshape(B, Lq, H * dk))
    if trace is not None:
        trace[prefix + ".ctx"] = ctx
    return _rg(ctx @ _rw(w[prefix + ".o.weight"]).T)


def ffn(w, prefix, x, cfg: T5Cfg):
    """HF:models/t5/modeling_t5.py:84-132 (T5DenseActDense / T5DenseGatedActDense), dropout omitted (p=0)."""
    if cfg.ffn_gated_gelu:
        h = _ra(gelu_new(_ra(x @ _rw(w[prefix + ".wi_0.weight"]).T)) * _ra(x @ _rw(w[prefix + ".wi_1.weight"]).T))
    else:
        h = _ra(torch.relu(x @ _rw(w[prefix + ".wi.weight"]).T))
    return _rg(h @ _rw(w[prefix + ".wo.weight"]).T)


def extended_mask(attention_mask: torch.Tensor, dtype) -> torch.Tensor:
    """(1 - mask) * finfo.min, [B,1,1,Lk]  (ref P5_T5.py:111-113 -> HF get_extended_attention_mask).
    The additive constant is always fp32's finfo.min (what the reference, which runs fp32, adds)."""
    m = attention_mask.to(dtype)
    return ((1.0 - m) * torch.finfo(torch.float32).min)[:, None, None, :]


def encode(w, cfg: T5Cfg, input_ids, whole_word_ids, attention_mask, trace=None):
    """ref:src/src_t5/model/P5_T5.py:74-204 (JointEncoder.forward), eval mode / dropout 0."""
    dt = w["shared.weight"].dtype
    x = w["shared.weight"][input_ids]
    if whole_word_ids is not None:
        x = x + w["encoder.whole_word_embeddings.weight"][whole_word_ids]  # P5_T5.py:94-100
    L = input_ids.shape[1]
    bias = compute_bias(w["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], L, L, True, cfg)
    bias = bias + extended_mask(attention_mask, dt)  # P5_T5.py:136-143
    for i in range(cfg.num_layers):
        b = f"encoder.block.{i}.layer"
        n = rms_norm(x, w[b + ".0.layer_norm.weight"], cfg.ln_eps)
        # (emulation) the engine's fused tcgen05 attention backward covers encoder lengths up to 256 (padded to 8)
        x = x + attention(w, b + ".0.SelfAttention", n, n, bias, cfg, trace, fused_bwd=((L + 7) // 8) * 8 <= 256)
        n = rms_norm(x, w[b + ".1.layer_norm.weight"], cfg.ln_eps)
        x = x + ffn(w, b + ".1.DenseReluDense", n, cfg)
        if trace is not None:
            trace[f"enc.{i}"] = x
    return rms_norm(x, w["encoder.final_layer_norm.weight"], cfg.ln_eps, fp32_grad=True)


def shift_right(labels: torch.Tensor, cfg: T5Cfg) -> torch.Tensor:
    """HF:models/t5/modeling_t5.py (_shift_right); labels hold pad=0, never -100 (Collator.py:25,32)."""
    out = torch.zeros_like(labels)
    out[:, 1:] = labels[:, :-1]
    out[:, 0] = cfg.decoder_start_id
    return out.masked_fill(out == -100, cfg.pad_id)


def decode(w, cfg: T5Cfg, dec_ids, enc_out, attention_mask, trace=None):
    """HF:models/t5/modeling_t5.py:637-793 (T5Stack as decoder) as called from ref P5_T5.py:338-350:
    causal self-attention with the decoder's unidirectional relative bias, NO decoder padding mask
    (decoder_attention_mask=None, P5_T5.py:340), cross-attention with zero position bias + encoder pad mask."""
    dt = w["shared.weight"].dtype
    y = w["shared.weight"][dec_ids]
    Ld = dec_ids.shape[1]
    self_bias = compute_bias(w["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], Ld, Ld, False, cfg)
    causal = torch.triu(torch.ones(Ld, Ld, dtype=torch.bool), diagonal=1)
    self_bias = self_bias + torch.zeros(Ld, Ld, dtype=dt).masked_fill(causal, torch.finfo(torch.float32).min)
    cross_bias = extended_mask(attention_mask, dt)
    for i in range(cfg.num_decoder_layers):
        b = f"decoder.block.{i}.layer"
        n = rms_norm(y, w[b + ".0.layer_norm.weight"], cfg.ln_eps)
        y = y + attention(w, b + ".0.SelfAttention", n, n, self_bias, cfg, trace)
        n = rms_norm(y, w[b + ".1.layer_norm.weight"], cfg.ln_eps)
        y = y + attention(w, b + ".1.EncDecAttention", n, enc_out, cross_bias, cfg, trace)
        n = rms_norm(y, w[b + ".2.layer_norm.weight"], cfg.ln_eps)
        y = y + ffn(w, b + ".2.DenseReluDense", n, cfg)
        if trace is not None:
            trace[f"dec.{i}"] = y
    return rms_norm(y, w["decoder.final_layer_norm.weight"], cfg.ln_eps)


def lm_logits(w, cfg: T5Cfg, y):
    """ref P5_T5.py:352-361: tied embeddings -> hidden * d_model^-0.5, then lm_head = shared^T."""
    if _EMU:   # engine: bf16 operands, the d_model^-0.5 factor is applied to the fp32 accumulator in the epilogue
        return _rg((y @ _rw(w["shared.weight"]).T) * (cfg.d_model ** -0.5))
    return (y * (cfg.d_model ** -0.5)) @ w["shared.weight"].T


def forward(w, cfg: T5Cfg, input_ids, whole_word_ids, attention_mask, labels, trace=None):
    """ref:src/src_t5/model/P5_T5.py:275-386 (P5_T5.forward), dropout 0.
    Returns (loss_tok [B*Ld] un-reduced CE as in P5_T5.py:364-369, logits [B, Ld, V])."""
    enc = encode(w, cfg, input_ids, whole_word_ids, attention_mask, trace)
    if trace is not None:
        trace["enc_out"] = enc
    dec_ids = shift_right(labels, cfg)
    y = decode(w, cfg, dec_ids, enc, attention_mask, trace)
    if trace is not None:
        trace["dec_out"] = y
    logits = lm_logits(w, cfg, y)
    loss_tok = torch.nn.functional.cross_entropy(
        logits.view(-1, logits.shape[-1]).float(), labels.reshape(-1), ignore_index=-100, reduction="none")
    return loss_tok.to(logits.dtype), logits


def runner_loss(loss_tok, output_attention, B, Ld):
    """ref:src/src_t5/runner/DistributedRunner.py:72-77."""
    m = (output_attention != 0).to(loss_tok.dtype)
    l = loss_tok.view(B, Ld) * m
    return (l.sum(dim=1) / m.sum(dim=1).clamp(min=1)).mean()


def loss_and_grads(w, cfg, input_ids, whole_word_ids, attention_mask, labels, output_attention):
    """forward + autograd backward of the runner loss; returns (loss, loss_tok, logits, grads dict)."""
    wr = {k: v.detach().clone().requires_grad_(True) for k, v in w.items()}
    loss_tok, logits = forward(wr, cfg, input_ids, whole_word_ids, attention_mask, labels)
    loss = runner_loss(loss_tok, output_attention, labels.shape[0], labels.shape[1])
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in wr.items()}
    return loss.detach(), loss_tok.detach(), logits.detach(), grads


# ----------------------------------------------------------------------------------------------------
# optimiser / schedule / clipping
# ----------------------------------------------------------------------------------------------------
def clip_grad_norm(grads: Dict[str, torch.Tensor], max_norm: float) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ as called at ref DistributedRunner.py:81: total L2 norm,
    coef = max_norm / (norm + 1e-6) clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(torch.float32)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads.values():
        g.mul_(coef.to(g.dtype))
    return total


def adamw_hf426(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.01, correct_bias=True):
    """transformers==4.26 optimization.AdamW.step (constructed at ref SingleRunner.py:191-214):
    eps is added to sqrt(v) WITHOUT bias correction, bias correction folded into the step size, decoupled decay
    applied AFTER the update with the un-corrected lr.  `step` is 1-based.  In-place on p, m, v."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


def adamw_weight_decay_for(name: str, weight_decay: float) -> float:
    """the reference's optimizer parameter groups (ref SingleRunner.py:186-205): `no_decay = ["bias", "LayerNorm.weight"]`
    matched BY SUBSTRING against the parameter name.  T5 calls its norms `layer_norm` (no match), but "bias" matches the
    two `...SelfAttention.relative_attention_bias.weight` tables, which therefore get weight_decay = 0."""
    return 0.0 if any(nd in name for nd in ("bias", "LayerNorm.weight")) else weight_decay


def linear_schedule(step: int, warmup: int, total: int) -> float:
    """get_linear_schedule_with_warmup multiplier (HF:optimization.py:101-104; ref SingleRunner.py:181-183,217).
    `step` = number of scheduler.step() calls so far (the first optimizer step runs with multiplier 0)."""
    if step < warmup:
        return float(step) / float(max(1, warmup))
    return max(0.0, float(total - step) / float(max(1, total - warmup)))


# ----------------------------------------------------------------------------------------------------
# constrained decoding: trie + beam search
# ----------------------------------------------------------------------------------------------------
class Trie:
    """ref:src/src_t5/utils/generation_trie.py:7-97 restated: nested-dict trie over token paths;
    get(prefix) = allowed next tokens (children of the node reached by prefix; [] if the prefix leaves the trie)."""

    def __init__(self, sequences: Sequence[Sequence[int]] = ()):
        self.root: dict = {}
        self.len = 0
        for s in sequences:
            self.add(s)

    def add(self, seq):
        node = self.root
        for t in seq:
            node = node.setdefault(int(t), {})
        self.len += 1

    def get(self, prefix) -> List[int]:
        node = self.root
        for t in prefix:
            if int(t) not in node:
                return []
            node = node[int(t)]
        return list(node.keys())

    def __len__(self):
        return self.len


def decoder_step_logits(w, cfg, dec_ids, enc_out, attention_mask):
    """logits of the LAST position for a batch of prefixes (recomputes the prefix: the oracle keeps no KV cache;
    result is identical to HF's cached decode)."""
    y = decode(w, cfg, dec_ids, enc_out, attention_mask)
    return lm_logits(w, cfg, y[:, -1:, :])[:, 0, :]


class _DecodeCache:
    """Incremental decoding for beam_search(cached=True): self-attention K/V per beam row (re-ordered by the beam
    indices each step, as HF's cache reorder does) and cross-attention K/V computed ONCE per user.  Arithmetic is the
    same as decode() position by position (the test suite checks cached == uncached); it only avoids recomputing the
    prefix and the K-fold repeated cross projections, so that the BASELINE eval shape (20 users x 20 beams, T5-base)
    finishes in seconds on CPU."""

    def __init__(self, w, cfg: T5Cfg, enc_out, attention_mask, K: int):
        self.w, self.cfg, self.K = w, cfg, K
        B, Le, _ = enc_out.shape
        H, dk = cfg.num_heads, cfg.d_kv
        self.B = B
        self.cross = []
        for i in range(cfg.num_decoder_layers):
            pre = f"decoder.block.{i}.layer.1.EncDecAttention"
            k = _ra(enc_out @ _rw(w[pre + ".k.weight"]).T).view(B, Le, H, dk).transpose(1, 2)
            v = _ra(enc_out @ _rw(w[pre + ".v.weight"]).T).view(B, Le, H, dk).transpose(1, 2)
            self.cross.append((k, v))
        self.cross_bias = extended_mask(attention_mask, enc_out.dtype)          # [B,1,1,Le]
        self.self_kv = [None] * cfg.num_decoder_layers
        self.t = 0

    def reorder(self, beam_idx):
        """beam_idx [B, K]: parent beam of every new running beam"""
        flat = (beam_idx + torch.arange(self.B)[:, None] * self.K).reshape(-1)
        self.self_kv = [None if kv is None else (kv[0][flat], kv[1][flat]) for kv in self.self_kv]

    def step(self, tok):
        """tok [B*K]: decoder input token at position self.t -> logits [B*K, V] of that position"""
        w, cfg, K, B = self.w, self.cfg, self.K, self.B
        H, dk, t = cfg.num_heads, cfg.d_kv, self.t
        R = tok.shape[0]
        y = w["shared.weight"][tok][:, None, :]
        table = w["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
        self_bias = compute_bias(table, 1, t + 1, False, cfg, q_offset=t)      # [1,H,1,t+1]; every cached key is <= t
        for i in range(cfg.num_decoder_layers):
            b = f"decoder.block.{i}.layer"
            n = rms_norm(y, w[b + ".0.layer_norm.weight"], cfg.ln_eps)
            pre = b + ".0.SelfAttention"
            q = _ra(n @ _rw(w[pre + ".q.weight"]).T).view(R, 1, H, dk).transpose(1, 2)
            k = _ra(n @ _rw(w[pre + ".k.weight"]).T).view(R, 1, H, dk).transpose(1, 2)
            v = _ra(n @ _rw(w[pre + ".v.weight"]).T).view(R, 1, H, dk).transpose(1, 2)
            if self.self_kv[i] is not None:
                k = torch.cat([self.self_kv[i][0], k], dim=2)
                v = torch.cat([self.self_kv[i][1], v], dim=2)
            self.self_kv[i] = (k, v)
            ctx = _ra(_attn_core(q, k, v, self_bias).transpose(1, 2).reshape(R, 1, H * dk))
            y = y + ctx @ _rw(w[pre + ".o.weight"]).T
            n = rms_norm(y, w[b + ".1.layer_norm.weight"], cfg.ln_eps)
            pre = b + ".1.EncDecAttention"
            q = _ra(n @ _rw(w[pre + ".q.weight"]).T).view(B, K, H, dk).transpose(1, 2)      # beams of a user = query rows
            ck, cv = self.cross[i]
            ctx = _attn_core(q, ck, cv, self.cross_bias)                                    # [B,H,K,dk]
            ctx = _ra(ctx.transpose(1, 2).reshape(R, 1, H * dk))
            y = y + ctx @ _rw(w[pre + ".o.weight"]).T
            n = rms_norm(y, w[b + ".2.layer_norm.weight"], cfg.ln_eps)
            y = y + ffn(w, b + ".2.DenseReluDense", n, cfg)
        y = rms_norm(y, w["decoder.final_layer_norm.weight"], cfg.ln_eps)
        self.t += 1
        return lm_logits(w, cfg, y)[:, 0, :]


def beam_search(w, cfg: T5Cfg, input_ids, whole_word_ids, attention_mask, trie: Trie, num_beams: int,
                num_return: int, max_length: int, length_penalty: float = 1.0, on_empty: str = "raise", cached: bool = False):
    """Restatement of HF:generation/utils.py:3076-3400 (_beam_search, transformers 5.5, early_stopping=False,
    do_sample=False) with PrefixConstrainedLogitsProcessor (HF:generation/logits_process.py:1532-1549) driven by
    the reference trie (ref DistributedRunner.py:344-371; encoder run once with whole-word ids, P5_T5.py:519-578).
    Returns (sequences [B*R, T] int64 pad-filled, scores [B*R] fp32)."""
    B = input_ids.shape[0]
    K, V = num_beams, cfg.vocab_size
    enc = encode(w, cfg, input_ids, whole_word_ids, attention_mask)
    enc_k = None if cached else enc.repeat_interleave(K, dim=0)   # _expand_inputs_for_generation
    mask_k = None if cached else attention_mask.repeat_interleave(K, dim=0)
    cur_len = 1
    prompt_len = 1
    running_seq = torch.full((B, K, max_length), cfg.pad_id, dtype=torch.long)
    running_seq[:, :, 0] = cfg.decoder_start_id
    sequences = running_seq.clone()
    running_scores = torch.zeros(B, K)
    running_scores[:, 1:] = NEG_INF_BEAM
    beam_scores = torch.full((B, K), NEG_INF_BEAM)
    is_finished = torch.zeros(B, K, dtype=torch.bool)
    unsat = torch.ones(B, 1, dtype=torch.bool)
    gen_len = torch.zeros(B, K, dtype=torch.long)          # generated length of each finished hypothesis
    run_gen_len = torch.zeros(B, K, dtype=torch.long)
    top_mask = torch.cat([torch.ones(K, dtype=torch.bool), torch.zeros(K, dtype=torch.bool)])
    cache = _DecodeCache(w, cfg, enc, attention_mask, K) if cached else None
    while True:
        flat = running_seq[:, :, :cur_len].reshape(B * K, cur_len)
        if cached:
            logits = cache.step(flat[:, -1]).float()
        else:
            logits = decoder_step_logits(w, cfg, flat, enc_k, mask_k).float()
        logp = torch.log_softmax(logits, dim=-1)
        # prefix-constrained mask (HF:logits_process.py:1532-1549): -inf everywhere except allowed tokens
        mask = torch.full_like(logp, -math.inf)
        for r in range(B * K):
            allowed = trie.get(flat[r].tolist())
            if len(allowed) == 0:
                if on_empty == "raise":  # transformers 5.5 (HF:logits_process.py:1540-1545)
                    raise ValueError("prefix_allowed_tokens_fn returned an empty list")
                continue                 # transformers 4.26: the row stays all -inf
            mask[r, allowed] = 0
        logp = logp + mask
        logp = logp.view(B, K, V) + running_scores[:, :, None]
        logp = logp.view(B, K * V)
        topk_lp, topk_idx = torch.topk(logp, k=2 * K)                      # :2981
        topk_beam = topk_idx // V
        topk_tok = topk_idx % V
        topk_seq = torch.gather(running_seq, 1, topk_beam[:, :, None].expand(-1, -1, max_length)).clone()
        topk_seq[:, :, cur_len] = topk_tok
        hits = (topk_tok == cfg.eos_id) | (cur_len + 1 >= max_length)     # EosTokenCriteria | MaxLengthCriteria
        # e. running beams for next iteration (:2999-3019)
        run_lp = topk_lp + hits.float() * NEG_INF_BEAM
        nxt = torch.topk(run_lp, k=K)[1]
        running_seq = torch.gather(topk_seq, 1, nxt[:, :, None].expand(-1, -1, max_length))
        running_scores = torch.gather(run_lp, 1, nxt)
        if cached:
            cache.reorder(torch.gather(topk_beam, 1, nxt))
        # f. finished beams (:3021-3073)
        just_finished = hits & top_mask[None, :]
        fin_lp = topk_lp / ((cur_len + 1 - prompt_len) ** length_penalty)
        fin_lp = fin_lp + (~unsat).float() * NEG_INF_BEAM
        fin_lp = fin_lp + (~just_finished).float() * NEG_INF_BEAM
        merged_seq = torch.cat([sequences, topk_seq], dim=1)
        merged_sc = torch.cat([beam_scores, fin_lp], dim=1)
        merged_fin = torch.cat([is_finished, just_finished], dim=1)
        merged_len = torch.cat([gen_len, torch.full((B, 2 * K), cur_len + 1 - prompt_len, dtype=torch.long)], dim=1)
        sel = torch.topk(merged_sc, k=K)[1]
        sequences = torch.gather(merged_seq, 1, sel[:, :, None].expand(-1, -1, max_length))
        beam_scores = torch.gather(merged_sc, 1, sel)
        is_finished = torch.gather(merged_fin, 1, sel)
        gen_len = torch.gather(merged_len, 1, sel)
        cur_len += 1
        # early-stop heuristic (:2876-2921), early_stopping=False
        best_running = running_scores[:, :1] / ((cur_len - prompt_len) ** length_penalty)
        worst_fin = torch.where(is_finished, beam_scores.min(dim=1, keepdim=True)[0], torch.tensor(NEG_INF_BEAM))
        unsat = unsat & (best_running > worst_fin).any(dim=-1, keepdim=True)
        # loop condition (:2923-2943)
        if not (bool(unsat.any()) and not bool(hits.all())):
            break
    seqs = sequences[:, :num_return, :].reshape(B * num_return, max_length)
    scores = beam_scores[:, :num_return].reshape(B * num_return)
    # crop to prompt + longest generated among returned rows that were actually written (beam_indices != -1)
    lens = gen_len[:, :num_return].reshape(-1) * is_finished[:, :num_return].reshape(-1).long()
    out_len = prompt_len + int(lens.max().item()) if lens.numel() else prompt_len
    return seqs[:, :out_len], scores


# ----------------------------------------------------------------------------------------------------
# metrics (ref:src/src_t5/utils/evaluate.py:37-92)
# ----------------------------------------------------------------------------------------------------
def rel_results(predictions: Sequence, targets: Sequence, scores: Sequence[float], k: int) -> List[List[int]]:
    """evaluate.py:37-58: per user, sort the k predictions by score (desc, stable) and mark 0/1 vs the gold."""
    res = []
    for b, gt in enumerate(targets):
        pairs = list(zip(predictions[b * k:(b + 1) * k], scores[b * k:(b + 1) * k]))
        pairs = sorted(pairs, key=lambda x: x[1], reverse=True)
        res.append([1 if p == gt else 0 for p, _ in pairs])
    return res


def hit_at_k(rel, k):   # evaluate.py:86-92 (sum over users, not mean)
    return float(sum(1 for row in rel if sum(row[:k]) > 0))


def ndcg_at_k(rel, k):  # evaluate.py:72-83
    return float(sum(sum(r / math.log(i + 2, 2) for i, r in enumerate(row[:k])) for row in rel))


# ----------------------------------------------------------------------------------------------------
# whole-word ids (ref:src/src_t5/processor/Collator.py:72-83)
# ----------------------------------------------------------------------------------------------------
def calculate_whole_word_ids(tokens: Sequence[str]) -> List[int]:
    out, curr = [], 0
    for t in tokens:
        if t == "<pad>":
            curr = 0
        if t.startswith("▁"):
            curr += 1
        out.append(curr)
    return out[: len(tokens) - 1] + [0]


# ----------------------------------------------------------------------------------------------------
# synthetic ML-1M-shaped data: the generators live in openp5_b200/synth.py (pure torch-CPU, no
# engine dependency) so that bench.py, the tests and this oracle all see identical token ids
# ----------------------------------------------------------------------------------------------------
from openp5_b200.synth import DIGIT_BASE, ITEM_PREFIX, synth_items, synth_batch  # noqa: E402,F401


# ----------------------------------------------------------------------------------------------------
# collaborative indexing (ref:src/src_t5/utils/indexing.py:149-282) — restated with the reference's own numpy / Python
# loops; pinned against the reference function itself in tests/test_dropin_cpu.py (same process, SpectralClustering
# recorded), and used on the GPU box as the checker of csrc/indexing.cu + openp5_b200/indexing.py
# ----------------------------------------------------------------------------------------------------
class collab_helpers:
    """utils/indexing.py:258-282 (token bookkeeping), verbatim semantics"""

    @staticmethod
    def add_token_to_indexing(item_map, grouping, index_now, token_size):
        for group in grouping:
            index_now = index_now % token_size
            for (item, idx) in grouping[group]:
                if item not in item_map:
                    item_map[item] = ''
                item_map[item] += f'<CI{index_now}>'
            index_now += 1
        return item_map, index_now

    @staticmethod
    def add_last_token_to_indexing_sequential(item_map, item_list, token_size):
        for i in range(len(item_list)):
            item = item_list[i]
            if item not in item_map:
                item_map[item] = ''
            item_map[item] += f'<CI{i}>'
        return item_map

    @staticmethod
    def add_last_token_to_indexing_random(item_map, item_list, token_size):
        import random
        last_tokens = random.sample([i for i in range(token_size)], len(item_list))
        for i in range(len(item_list)):
            item = item_list[i]
            if item not in item_map:
                item_map[item] = ''
            item_map[item] += f'<CI{last_tokens[i]}>'
        return item_map


def collab_item_ids(user_sequence_dict):
    """utils/indexing.py:153-165: (all_items, train_items, item2id, id2item) — ids follow the iteration order of the set"""
    all_items, train_items = set(), set()
    for user in user_sequence_dict:
        all_items.update(set(user_sequence_dict[user]))
        train_items.update(set(user_sequence_dict[user][:-2]))
    item2id, id2item = dict(), dict()
    for item in train_items:
        item2id[item] = len(item2id)
        id2item[len(id2item)] = item
    return all_items, train_items, item2id, id2item


def cooccurrence_matrix_ref(user_sequence_dict, item2id, float32=0):
    """utils/indexing.py:168-180"""
    import numpy as np
    from itertools import combinations
    adj = np.zeros((len(item2id), len(item2id)), dtype=np.float32 if float32 > 0 else np.float64)
    for user in user_sequence_dict:
        for a, b in combinations(user_sequence_dict[user][:-2], 2):
            adj[item2id[a]][item2id[b]] += 1
            adj[item2id[b]][item2id[a]] += 1
    return adj


def submatrix_ref(adj, idx):
    """utils/indexing.py:220-231"""
    import numpy as np
    m = len(idx)
    sub = np.zeros((m, m), dtype=adj.dtype)
    for i in range(m):
        for j in range(i + 1, m):
            sub[i][j] = adj[idx[i]][idx[j]]
            sub[j][i] = adj[idx[j]][idx[i]]
    return sub


def generate_collaborative_id_ref(user_sequence_dict, token_size, cluster_num, last_token, float32):
    """utils/indexing.py:149-256 restated"""
    from collections import defaultdict
    from sklearn.cluster import SpectralClustering
    H = collab_helpers
    all_items, train_items, item2id, id2item = collab_item_ids(user_sequence_dict)
    adj = cooccurrence_matrix_ref(user_sequence_dict, item2id, float32)
    sc = lambda m: SpectralClustering(n_clusters=cluster_num, assign_labels="cluster_qr", random_state=0,
                                      affinity="precomputed").fit(m).labels_.tolist()
    labels = sc(adj)
    grouping = defaultdict(list)
    for i in range(len(labels)):
        grouping[labels[i]].append((id2item[i], i))
    item_map, index_now = H.add_token_to_indexing(dict(), grouping, 0, token_size)
    queue = [grouping[g] for g in grouping]
    while queue:
        group_items = queue.pop(0)
        if len(group_items) <= token_size:
            item_list = [it[0] for it in group_items]
            item_map = (H.add_last_token_to_indexing_sequential if last_token == 'sequential' else
                        H.add_last_token_to_indexing_random)(item_map, item_list, token_size)
        else:
            labels = sc(submatrix_ref(adj, [it[1] for it in group_items]))
            grouping = defaultdict(list)
            for i in range(len(labels)):
                grouping[labels[i]].append(group_items[i])
            item_map, index_now = H.add_token_to_indexing(item_map, grouping, index_now, token_size)
            for g in grouping:
                queue.append(grouping[g])
    remaining = list(all_items - train_items)
    if remaining:
        item_map = (H.add_last_token_to_indexing_sequential if last_token == 'sequential' else
                    H.add_last_token_to_indexing_random)(item_map, remaining, token_size)
    return item_map
