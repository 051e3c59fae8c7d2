"""Host-side mirror of the reference model object for the B200 engine.

`P5B200` exposes the call surface the reference runner uses on `model`:
  model(input_ids=, whole_word_ids=, attention_mask=, labels=, return_dict=True) -> {"loss": [B*Ld], "logits"}
      (ref src/src_t5/runner/DistributedRunner.py:63-70 -> model/P5_T5.py:275-386)
  model.generate(input_ids=, attention_mask=, whole_word_ids=, max_length=, prefix_allowed_tokens_fn= | trie=,
                 num_beams=, num_return_sequences=, output_scores=True, return_dict_in_generate=True)
      (ref DistributedRunner.py:361-371)
  named_parameters / parameters / state_dict / load_state_dict / zero_grad / train / eval / shared.weight
      (ref main.py:184-198, utils/initialization.py:27, utils/utils.py:119-129)
All arithmetic runs in libp5b200.so (hand-written sm_100a CUDA) through the C-ABI in include/p5_b200.h; torch is
used for device memory views, streams and autograd plumbing only.  There is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, Iterable, List, Optional, Sequence

import torch

from . import _lib

BACKBONES = {
    "t5-tiny": dict(d_model=64, d_ff=128, num_layers=2, num_decoder_layers=2, num_heads=2),
    "t5-small": dict(d_model=512, d_ff=2048, num_layers=6, num_decoder_layers=6, num_heads=8),
    "t5-base": dict(d_model=768, d_ff=3072, num_layers=12, num_decoder_layers=12, num_heads=12),
    "t5-large": dict(d_model=1024, d_ff=4096, num_layers=24, num_decoder_layers=24, num_heads=16),
}


class _DevArray:
    """zero-copy view of engine-owned device memory for torch.as_tensor"""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
            "strides": None,
        }


def _i32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()


class _P5Function(torch.autograd.Function):
    """loss.backward() on the per-token loss calls p5_backward, which accumulates into the engine's grad buffers"""

    @staticmethod
    def forward(ctx, anchor, model, ids, mask, ww, labels, training, seed, want_logits):
        loss_tok, logits = model._forward_raw(ids, mask, ww, labels, training, seed, want_logits)
        ctx.model = model
        ctx.mark_non_differentiable(logits)
        return loss_tok, logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        ctx.model._backward_raw(dloss)
        return (None,) * 9


class Trie:
    """Device trie handle (replaces ref utils/generation_trie.py Trie + prefix_allowed_tokens_fn for generate)."""

    def __init__(self, model: "P5B200", sequences: Sequence[Sequence[int]]):
        lib = _lib.load()
        flat: List[int] = []
        offs = [0]
        for s in sequences:
            flat.extend(int(t) for t in s)
            offs.append(len(flat))
        self._paths = (C.c_int32 * len(flat))(*flat)
        self._offs = (C.c_int64 * len(offs))(*offs)
        self.handle = C.c_void_p()
        _lib.check(lib.p5_trie_build(model.handle, self._paths, self._offs, len(sequences), C.byref(self.handle)))
        self.n_paths = len(sequences)

    def stats(self):
        n, e, d = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.load().p5_trie_stats(self.handle, C.byref(n), C.byref(e), C.byref(d)))
        return dict(nodes=n.value, edges=e.value, max_depth=d.value)

    def get(self, prefix: Sequence[int], cap: int = 4096) -> List[int]:
        """allowed next tokens after `prefix`, looked up in the DEVICE CSR copy"""
        pre = (C.c_int32 * max(1, len(prefix)))(*[int(t) for t in prefix])
        out = (C.c_int32 * cap)()
        n = C.c_int()
        _lib.check(_lib.load().p5_trie_get(self.handle, pre, len(prefix), out, cap, C.byref(n)))
        return [out[i] for i in range(min(n.value, cap))]

    def __del__(self):
        try:
            if self.handle:
                _lib.load().p5_trie_free(self.handle)
                self.handle = None
        except Exception:
            pass


class P5B200:
    """P5 T5 encoder-decoder on one B200.  `precision`: "bf16" (tcgen05 tensor-core path, the benchmarked mode), "fp32"
    (SIMT fp32 parity path) or "bf16x3" (fp32 storage, every linear layer through the tcgen05 kernel as a hi/lo-split
    three-product GEMM: the tensor-core path at fp32-class accuracy, gated at the north star's 1e-3)."""

    def __init__(self, backbone: str = "t5-small", vocab_size: int = 32100, device: Optional[int] = None,
                 precision: str = "bf16", dropout: float = 0.1, max_batch: int = 64, max_enc_len: int = 512,
                 max_dec_len: int = 16, max_beams: int = 20, ffn_gated_gelu: bool = False, use_mn_major: bool = True,
                 **dims):
        if not torch.cuda.is_available():
            raise _lib.P5LibraryError("P5B200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        d = dict(BACKBONES[backbone]) if backbone in BACKBONES else {}
        d.update(dims)
        cfg = _lib.P5Config()
        cfg.vocab_size = vocab_size
        cfg.d_model, cfg.d_kv, cfg.d_ff = d["d_model"], 64, d["d_ff"]
        cfg.num_layers, cfg.num_decoder_layers, cfg.num_heads = d["num_layers"], d["num_decoder_layers"], d["num_heads"]
        cfg.rel_buckets, cfg.rel_max_distance = 32, 128
        cfg.ffn_gated_gelu = int(ffn_gated_gelu)
        cfg.whole_word_rows = 512
        cfg.dropout, cfg.ln_eps = float(dropout), 1e-6
        cfg.precision = {"fp32": 0, "bf16": 1, "bf16x3": 2}[precision]
        cfg.max_batch, cfg.max_enc_len, cfg.max_dec_len, cfg.max_beams = max_batch, max_enc_len, max_dec_len, max_beams
        cfg.use_mn_major = int(use_mn_major)
        self.cfg = cfg
        self.backbone, self.precision = backbone, precision
        self.config = SimpleNamespace(vocab_size=vocab_size, d_model=cfg.d_model, pad_token_id=0, eos_token_id=1,
                                      decoder_start_token_id=0, dropout_rate=dropout)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device_index):
            self.stream = torch.cuda.current_stream()
            _lib.check(self.lib.p5_create(C.byref(cfg), self.device_index, C.c_void_p(self.stream.cuda_stream),
                                          C.byref(self.handle)))
        self._build_param_views()
        self.module = self                                                      # DDP `.module` alias
        self.training = True
        self._versions = self._version_sum()
        self._shadow_trusted = False
        self._anchor = torch.zeros((), device=self.device, requires_grad=True)
        self._step_seed = 0
        self._opt_step = 0
        self.world_size, self.rank = 1, 0

    def _build_param_views(self):
        """torch views (zero-copy, `.grad` attached) over the engine's flat parameter / gradient buffers"""
        self._params: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        n = C.c_int()
        _lib.check(self.lib.p5_param_count(self.handle, C.byref(n)))
        for i in range(n.value):
            name, ndim = C.c_char_p(), C.c_int()
            shape = (C.c_int64 * 2)()
            data, grad = C.c_void_p(), C.c_void_p()
            _lib.check(self.lib.p5_param_info(self.handle, i, C.byref(name), C.byref(ndim), C.byref(shape), C.byref(data),
                                              C.byref(grad)))
            shp = (shape[0],) if ndim.value == 1 else (shape[0], shape[1])
            p = torch.as_tensor(_DevArray(data.value, shp), device=self.device).requires_grad_(True)
            p.grad = torch.as_tensor(_DevArray(grad.value, shp), device=self.device)
            self._params[name.value.decode()] = p
        self.shared = SimpleNamespace(weight=self._params["shared.weight"])   # ref utils/initialization.py:27

    def resize_token_embeddings(self, new_num_tokens: int):
        """ref main.py:193 `model.resize_token_embeddings(len(tokenizer))`: the engine is rebuilt in place for the new
        vocabulary (tied embedding / LM head rows kept, new rows ~ N(0, 1) as HF's T5 initialiser draws them); every
        parameter tensor obtained before this call is invalid afterwards."""
        new_num_tokens = int(new_num_tokens)
        if new_num_tokens == self.cfg.vocab_size:
            return self.shared
        self._on_stream()
        _lib.check(self.lib.p5_resize_vocab(self.handle, new_num_tokens))
        self.cfg.vocab_size = new_num_tokens
        self.config.vocab_size = new_num_tokens
        self._build_param_views()
        self._versions = self._version_sum()
        self._shadow_trusted = False
        return self.shared

    # ---------------------------------------------------------------- module protocol
    def _join_optimizer(self):
        """an asynchronous AdamW (train_step(..., overlap_optimizer=True)) may still be updating the buffers behind the
        torch views: order it before anything else that is enqueued on this stream"""
        _lib.check(self.lib.p5_optimizer_join(self.handle))

    def named_parameters(self):
        self._join_optimizer()
        return iter(self._params.items())

    def parameters(self):
        self._join_optimizer()
        return iter(self._params.values())

    def num_parameters(self) -> int:
        return sum(p.numel() for p in self._params.values())

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def to(self, *_a, **_k):
        return self

    def zero_grad(self, set_to_none: bool = False):
        _lib.check(self.lib.p5_zero_grad(self.handle))

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        """HF T5 key names incl. the tied aliases (ref DistributedRunner.py:155,169)."""
        self._join_optimizer()
        sd = OrderedDict((k, v.detach().clone()) for k, v in self._params.items())
        sd["encoder.embed_tokens.weight"] = sd["shared.weight"]
        sd["decoder.embed_tokens.weight"] = sd["shared.weight"]
        sd["lm_head.weight"] = sd["shared.weight"]
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        self._join_optimizer()
        aliases = {"encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight"}
        missing = [k for k in self._params if k not in sd]
        unexpected = [k for k in sd if k not in self._params and k not in aliases]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing={missing} unexpected={unexpected}")
        with torch.no_grad():
            for k, p in self._params.items():
                if k in sd:
                    if tuple(sd[k].shape) != tuple(p.shape):
                        raise RuntimeError(f"shape mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(p.shape)}")
                    p.copy_(sd[k].to(device=self.device, dtype=torch.float32))
        self.mark_params_changed()
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def mark_params_changed(self):
        """call after writing parameters through `.data` / raw pointers BETWEEN fused train_step() calls: such writes do
        not bump tensor versions, and train_step trusts the engine's own bf16 shadows otherwise"""
        _lib.check(self.lib.p5_params_changed(self.handle))
        self._versions = self._version_sum()

    # ---------------------------------------------------------------- optimiser state (checkpoint / resume)
    def _opt_views(self):
        views = OrderedDict()
        for i, (name, p) in enumerate(self._params.items()):
            m, v = C.c_void_p(), C.c_void_p()
            _lib.check(self.lib.p5_opt_state_info(self.handle, i, C.byref(m), C.byref(v)))
            views[name] = (torch.as_tensor(_DevArray(m.value, p.shape), device=self.device),
                           torch.as_tensor(_DevArray(v.value, p.shape), device=self.device))
        return views

    def optimizer_state_dict(self):
        """AdamW state in torch.optim layout keyed by parameter NAME: {"state": {name: {"step", "exp_avg", "exp_avg_sq"}},
        "step": n}.  The reference saves weights only (DistributedRunner.py:155,169), so its runs cannot resume."""
        self._join_optimizer()
        return {"step": int(self._opt_step),
                "state": OrderedDict((k, {"step": int(self._opt_step), "exp_avg": m.detach().clone().cpu(),
                                          "exp_avg_sq": v.detach().clone().cpu()}) for k, (m, v) in self._opt_views().items())}

    def load_optimizer_state_dict(self, sd):
        self._join_optimizer()
        views = self._opt_views()
        with torch.no_grad():
            for k, (m, v) in views.items():
                st = sd["state"][k]
                m.copy_(st["exp_avg"].to(self.device, torch.float32))
                v.copy_(st["exp_avg_sq"].to(self.device, torch.float32))
        self._opt_step = int(sd["step"])

    def _version_sum(self) -> int:
        return sum(p._version for p in self._params.values())

    def _sync_params(self, fast_path: bool = False):
        """Host code may write parameters in place through the torch views (`model.shared.weight.data[idx] = ...`,
        ref utils/initialization.py:27; torch optimisers).  Writes through `.data` do not bump tensor versions, so
        the bf16 GEMM shadows are refreshed (one 0.2 ms cast pass) before every call except between consecutive
        fused train_step() calls, where the engine's own AdamW keeps them current."""
        if not (fast_path and self._shadow_trusted and self._version_sum() == self._versions):
            self.mark_params_changed()
        self._shadow_trusted = fast_path

    def _on_stream(self):
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.stream.cuda_stream:
            raise _lib.P5LibraryError("P5B200 must be driven on the CUDA stream it was created on")

    # ---------------------------------------------------------------- encoder lengths (padding removal)
    def _set_enc_lengths(self, enc_lengths, attention_mask):
        """Tell the engine how many (right-padded) tokens each encoder row holds, so that it processes sum(lens)
        rows instead of B*Le.  Lengths come from the caller, or for free from a HOST attention mask (the collator's
        tensors start on the CPU, ref Collator.py:8-34); a device-resident mask without lengths keeps the padded path."""
        if enc_lengths is None and attention_mask is not None and not attention_mask.is_cuda:
            # the packed layout needs CONTIGUOUS right padding with >= 1 token per row (what the collator produces,
            # Collator.py:12-21); any other mask keeps the padded layout, where the mask itself is applied
            m = attention_mask != 0
            lens = m.sum(dim=1)
            prefix = torch.arange(m.shape[1])[None, :] < lens[:, None]
            if bool((m == prefix).all()) and bool((lens >= 1).all()):
                enc_lengths = lens
        if enc_lengths is None:
            return
        if torch.is_tensor(enc_lengths):
            enc_lengths = enc_lengths.to("cpu", torch.int32).tolist()
        arr = (C.c_int32 * len(enc_lengths))(*[int(x) for x in enc_lengths])
        _lib.check(self.lib.p5_set_enc_lengths(self.handle, arr, len(enc_lengths)))

    # ---------------------------------------------------------------- forward / backward
    def _forward_raw(self, ids, mask, ww, labels, training, seed, want_logits):
        B, Le = ids.shape
        Ld = labels.shape[1]
        loss_tok = torch.empty(B * Ld, device=self.device, dtype=torch.float32)
        logits = torch.empty((B, Ld, self.cfg.vocab_size), device=self.device, dtype=torch.float32) if want_logits \
            else torch.empty(0, device=self.device)
        _lib.check(self.lib.p5_forward(
            self.handle, ids.data_ptr(), mask.data_ptr(), ww.data_ptr() if ww is not None else None, labels.data_ptr(),
            B, Le, Ld, loss_tok.data_ptr(), logits.data_ptr() if want_logits else None, int(training), C.c_uint64(seed)))
        return loss_tok, logits

    def _backward_raw(self, dloss):
        dloss = dloss.to(torch.float32).contiguous()
        _lib.check(self.lib.p5_backward(self.handle, dloss.data_ptr()))

    def __call__(self, input_ids=None, whole_word_ids=None, attention_mask=None, labels=None, return_dict=True,
                 return_logits=True, enc_lengths=None, **_unused):
        """ref P5_T5.forward: returns {"loss": flat un-reduced per-token CE [B*Ld], "logits": [B, Ld, V]}"""
        self._on_stream()
        self._sync_params()
        self._set_enc_lengths(enc_lengths, attention_mask)
        ids = _i32(input_ids, self.device)
        mask = _i32(attention_mask if attention_mask is not None else (input_ids != 0), self.device)
        ww = _i32(whole_word_ids, self.device) if whole_word_ids is not None else None
        lab = _i32(labels, self.device)
        self._step_seed += 1
        grad = torch.is_grad_enabled()
        if grad:
            loss_tok, logits = _P5Function.apply(self._anchor, self, ids, mask, ww, lab, self.training, self._step_seed,
                                                 return_logits)
        else:
            loss_tok, logits = self._forward_raw(ids, mask, ww, lab, self.training, self._step_seed, return_logits)
        out = {"loss": loss_tok, "logits": logits if return_logits else None}
        return out if return_dict else (out["loss"], out["logits"])

    forward = __call__

    # ---------------------------------------------------------------- fused training step (fast path)
    def train_step(self, input_ids, whole_word_ids, attention_mask, labels, labels_attention, *, lr: float,
                   clip: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.01,
                   seed: Optional[int] = None, step: Optional[int] = None, enc_lengths=None,
                   overlap_optimizer: bool = False) -> torch.Tensor:
        """One optimisation step = ref DistributedRunner.py:63-87 (forward, runner loss, backward, clip, AdamW,
        zero_grad) + the gradient all-reduce DDP was meant to do.  Returns the scalar loss as a device tensor
        (no host sync).

        overlap_optimizer=True issues AdamW range by range on an engine-owned side stream so that it overlaps the
        forward of the NEXT train_step (which waits per layer).  Same arithmetic, same results; the only contract is
        that parameter tensors obtained earlier (`p = model.shared.weight`) are not read by the caller until the next
        call into this object (state_dict / named_parameters / generate / ... all join first) — which is how the
        reference's training loop behaves."""
        self._on_stream()
        self._sync_params(fast_path=True)
        self._set_enc_lengths(enc_lengths, attention_mask)
        ids = _i32(input_ids, self.device)
        mask = _i32(attention_mask, self.device)
        ww = _i32(whole_word_ids, self.device) if whole_word_ids is not None else None
        lab = _i32(labels, self.device)
        lmask = _i32(labels_attention, self.device)
        B, Le = ids.shape
        Ld = lab.shape[1]
        self._step_seed += 1
        loss = torch.empty(1, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.p5_train_fwd_bwd(
            self.handle, ids.data_ptr(), mask.data_ptr(), ww.data_ptr() if ww is not None else None, lab.data_ptr(),
            lmask.data_ptr(), B, Le, Ld, loss.data_ptr(), C.c_uint64(seed if seed is not None else self._step_seed)))
        if self.world_size > 1:
            _lib.check(self.lib.p5_allreduce_grads(self.handle))
        self._opt_step = step if step is not None else self._opt_step + 1
        # optimizer.step() + model.zero_grad() (DistributedRunner.py:85-87) in one pass over the flat buffers
        step_fn = self.lib.p5_adamw_step_zero_grad_async if overlap_optimizer else self.lib.p5_adamw_step_zero_grad
        _lib.check(step_fn(self.handle, lr, betas[0], betas[1], eps, weight_decay, self._opt_step, clip))
        self._versions = self._version_sum()
        return loss

    def grad_norm(self) -> torch.Tensor:
        out = torch.empty(1, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.p5_grad_norm(self.handle, out.data_ptr()))
        return out

    def adamw_step(self, lr, step, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, clip=0.0):
        _lib.check(self.lib.p5_adamw_step(self.handle, lr, betas[0], betas[1], eps, weight_decay, int(step), clip))
        self._versions = self._version_sum()

    # ---------------------------------------------------------------- data parallel
    def init_data_parallel(self, process_group=None):
        """create the NCCL communicator through the existing torch.distributed group (rank 0 broadcasts the id)"""
        import torch.distributed as dist
        world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            _lib.check(self.lib.p5_comm_unique_id(buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(self.device)
        dist.broadcast(uid, src=0, group=process_group)
        host = uid.cpu().numpy().tobytes()
        _lib.check(self.lib.p5_comm_init(self.handle, host, rank, world))
        self.world_size, self.rank = world, rank
        # replicas start from rank 0's weights (DDP ctor broadcast, ref DistributedRunner.py:26)
        for p in self._params.values():
            dist.broadcast(p, src=0, group=process_group)
        self.mark_params_changed()

    def allreduce_grads(self):
        _lib.check(self.lib.p5_allreduce_grads(self.handle))

    # ---------------------------------------------------------------- generation
    def build_trie(self, sequences: Sequence[Sequence[int]]) -> Trie:
        return Trie(self, sequences)

    def generate(self, input_ids=None, attention_mask=None, whole_word_ids=None, max_length: int = 50,
                 prefix_allowed_tokens_fn=None, trie: Optional[Trie] = None, num_beams: int = 1,
                 num_return_sequences: int = 1, length_penalty: float = 1.0, output_scores: bool = True,
                 return_dict_in_generate: bool = True, **_unused):
        """ref model.generate(...) call site DistributedRunner.py:361-371.  The reference passes an opaque Python
        callback built from a Trie (gt.prefix_allowed_tokens_fn(trie)); the closure's trie is recovered and flattened
        to the device CSR form once and cached on the callable."""
        self._on_stream()
        # an eval loop calls generate() batch after batch on unchanged weights (ref DistributedRunner.py:359-371): the bf16
        # shadows are re-cast only when a parameter tensor's version moved or after load_state_dict / mark_params_changed
        self._sync_params(fast_path=True)
        if trie is None:
            trie = self._trie_from_callback(prefix_allowed_tokens_fn)
        ids = _i32(input_ids, self.device)
        mask = _i32(attention_mask if attention_mask is not None else (input_ids != 0), self.device)
        ww = _i32(whole_word_ids, self.device) if whole_word_ids is not None else None
        B, Le = ids.shape
        R = num_return_sequences
        seqs = torch.zeros((B * R, max_length), device=self.device, dtype=torch.int32)
        scores = torch.empty(B * R, device=self.device, dtype=torch.float32)
        out_len = C.c_int()
        _lib.check(self.lib.p5_generate(self.handle, ids.data_ptr(), mask.data_ptr(),
                                        ww.data_ptr() if ww is not None else None, B, Le, trie.handle, num_beams, R,
                                        max_length, length_penalty, seqs.data_ptr(), scores.data_ptr(), C.byref(out_len)))
        sequences = seqs[:, : out_len.value].to(torch.int64)
        if return_dict_in_generate:
            return {"sequences": sequences, "sequences_scores": scores if output_scores else None}
        return sequences

    def eval_metric_sums(self, sequences: torch.Tensor, sequences_scores: torch.Tensor, gold: torch.Tensor, num_beams: int,
                         ks: Sequence[int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """hit@k / ndcg@k SUMMED over the users of one eval batch, on the device (ref utils/evaluate.py:37-92 and
        DistributedRunner.py:376-393 on token-id paths).  Returns / accumulates into a float32 tensor
        [hit@ks..., ndcg@ks...]; no host synchronisation."""
        self._on_stream()
        seqs = _i32(sequences, self.device)
        sc = sequences_scores.to(self.device, torch.float32).contiguous()
        g = _i32(gold, self.device)
        B = g.shape[0]
        assert seqs.shape[0] == B * num_beams and sc.numel() == B * num_beams
        kd = torch.tensor([int(k) for k in ks], dtype=torch.int32, device=self.device)
        if out is None:
            out = torch.zeros(2 * len(ks), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.p5_eval_metrics(self.handle, seqs.data_ptr(), sc.data_ptr(), B, num_beams, seqs.shape[1], g.data_ptr(),
                                            g.shape[1], kd.data_ptr(), len(ks), out.data_ptr()))
        return out

    def eval_metric_sums_filtered(self, sequences, sequences_scores, gold, rows_per_user: int, ks: Sequence[int], positives,
                                  n_positives, k_cut: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """filtered hit@k / ndcg@k sums (ref utils/evaluate.py:6-35 via DistributedRunner.py:204-265): `positives`
        [B, Pmax, Tp] token paths of each user's already-interacted items, `n_positives` [B]; rows equal to a positive
        are skipped, the first `k_cut` remaining rows (by score) form the relevance list."""
        self._on_stream()
        seqs = _i32(sequences, self.device)
        sc = sequences_scores.to(self.device, torch.float32).contiguous()
        g = _i32(gold, self.device)
        pos = _i32(positives, self.device)
        npos = _i32(n_positives, self.device)
        B = g.shape[0]
        assert seqs.shape[0] == B * rows_per_user and pos.dim() == 3 and pos.shape[0] == B
        kd = torch.tensor([int(k) for k in ks], dtype=torch.int32, device=self.device)
        if out is None:
            out = torch.zeros(2 * len(ks), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.p5_eval_metrics_filtered(self.handle, seqs.data_ptr(), sc.data_ptr(), B, rows_per_user, seqs.shape[1],
                                                     g.data_ptr(), g.shape[1], pos.data_ptr(), npos.data_ptr(), pos.shape[1],
                                                     pos.shape[2], kd.data_ptr(), len(ks), int(k_cut), out.data_ptr()))
        return out

    def _trie_from_callback(self, fn) -> Trie:
        if fn is None:
            raise ValueError("generate() needs trie= or a prefix_allowed_tokens_fn built from a Trie")
        cached = getattr(fn, "_p5_device_trie", None)
        if cached is not None:
            return cached
        src = None
        for cell in (fn.__closure__ or ()):
            obj = cell.cell_contents
            if hasattr(obj, "trie_dict") or hasattr(obj, "root"):
                src = obj
                break
        if src is None:
            raise ValueError("prefix_allowed_tokens_fn does not close over a Trie; pass trie= explicitly")
        root = getattr(src, "trie_dict", None)
        if root is None:
            root = src.root
        paths: List[List[int]] = []

        def walk(node, prefix):
            if not node:
                paths.append(prefix)
                return
            for tok, child in node.items():
                walk(child, prefix + [int(tok)])
        walk(root, [])
        t = Trie(self, paths)
        try:
            fn._p5_device_trie = t
        except Exception:
            pass
        return t

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.p5_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
