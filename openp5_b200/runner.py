"""B200Runner — drop-in for the reference's DistributedRunner (ref src/src_t5/runner/DistributedRunner.py) over the
P5B200 engine.  The data plumbing (MultiTaskDataset, samplers, Collator, TestDataset, tokenizer) stays the
reference's own host code and is consumed through the same loader objects; only the five hot-path call sites
 are replaced:

  train fwd        DistributedRunner.py:63-70   ->  P5B200.train_step (fused fwd + runner loss + bwd)
  bwd/clip/step    :80-87                        ->  same call: clip_grad_norm_ + AdamW(4.26 semantics) on device
  grad all-reduce  (DDP wrapper :26, bypassed at :63 in the reference)  ->  NCCL mean all-reduce, overlapped
  validation fwd   :121-128                      ->  P5B200.__call__ under no_grad
  eval generate    :361-371                      ->  P5B200.generate with the device trie

Differences kept on purpose: gradients ARE averaged over ranks; the two per-step
dist.barrier() calls and the per-step loss all-reduce (:83,:90-93) are dropped (loss is reduced once per epoch).
"""
from __future__ import annotations

import logging
import math
from typing import List, Sequence

import torch


# ---------------------------------------------------------------------------------------------------------------
# host-side pieces (no GPU needed; covered by the CPU tests)
# ---------------------------------------------------------------------------------------------------------------
def linear_schedule(step: int, warmup: int, total: int) -> float:
    """get_linear_schedule_with_warmup multiplier (HF:optimization.py:101-104 via ref SingleRunner.py:181-183,217);
    `step` = number of scheduler.step() calls so far, so the first optimizer step runs with lr = 0."""
    if step < warmup:
        return float(step) / float(max(1, warmup))
    return max(0.0, float(total - step) / float(max(1, total - warmup)))


def schedule_plan(batches_per_epoch: int, epochs: int, warmup_prop: float, grad_accum: int = 1):
    """(total_steps, warmup_steps) exactly as ref SingleRunner.py:181-183"""
    total = batches_per_epoch // max(1, grad_accum) * epochs
    return total, int(total * warmup_prop)


def strip_special(row: Sequence[int], pad: int = 0, eos: int = 1) -> tuple:
    """token path of a generated / gold row without pad and eos (== batch_decode(skip_special_tokens=True) as a key)"""
    return tuple(int(t) for t in row if int(t) != pad and int(t) != eos)


def rel_results(seqs, scores, gold, k: int) -> List[List[int]]:
    """ref utils/evaluate.py:37-58 on token-id paths: per user, the k predictions sorted by score (stable, desc)
    marked 1 where they equal the gold item.  seqs [B*k, T], scores [B*k], gold [B, Ld] (any int sequences)."""
    out = []
    B = len(gold)
    for b in range(B):
        g = strip_special(gold[b])
        pairs = [(strip_special(seqs[b * k + i]), float(scores[b * k + i])) for i in range(k)]
        pairs = sorted(pairs, key=lambda x: x[1], reverse=True)
        out.append([1 if p == g else 0 for p, _ in pairs])
    return out


def metric_sums(rel: List[List[int]], names: Sequence[str]) -> List[float]:
    """ref utils/evaluate.py:60-92: hit@k / ndcg@k SUMMED over users (divided by the all-reduced count later)"""
    res = []
    for m in names:
        kind, k = m.lower().split("@")
        k = int(k)
        if kind.startswith("hit"):
            res.append(float(sum(1 for r in rel if sum(r[:k]) > 0)))
        elif kind.startswith("ndcg"):
            res.append(float(sum(sum(x / math.log(i + 2, 2) for i, x in enumerate(r[:k])) for r in rel)))
        else:
            raise ValueError(m)
    return res


def allreduce_metrics(sums: Sequence[float], count: int, device="cpu"):
    """ref DistributedRunner.py:389-395: SUM all-reduce of the metric sums and the user count, then divide"""
    import torch.distributed as dist
    t = torch.tensor(list(sums) + [float(count)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total = t[-1].item()
    return [x / max(total, 1.0) for x in t[:-1].tolist()], int(total)


def shard_for_rank(indices: Sequence[int], rank: int, world: int) -> List[int]:
    """ref processor/DistMultiDataTaskSampler.py:30-34,55-65: rank r takes list[r::world]; short shards wrap around so
    every rank emits the same number of samples"""
    mine = list(indices[rank::world])
    target = math.ceil(len(indices) / world)
    i = 0
    while len(mine) < target and mine:
        mine.append(mine[i % len(mine)])
        i += 1
    return mine


# ---------------------------------------------------------------------------------------------------------------
class B200Runner:
    def __init__(self, model, tokenizer, train_loader, valid_loader, device, args, rank: int = 0):
        self.model, self.tokenizer = model, tokenizer
        self.train_loader, self.valid_loader = train_loader, valid_loader
        self.device, self.args, self.rank = device, args, rank
        self.metrics = getattr(args, "metrics", "hit@5,hit@10,ndcg@5,ndcg@10").split(",")
        self.generate_num = max(int(m.split("@")[1]) for m in self.metrics)
        self.testloaders = []
        self.global_step = 0
        self.total_steps, self.warmup_steps = self.create_optimizer_and_scheduler()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and model.world_size == 1:
            model.init_data_parallel()

    def create_optimizer_and_scheduler(self):
        a = self.args
        n = len(self.train_loader) if self.train_loader is not None else 0
        return schedule_plan(n, getattr(a, "epochs", 1), getattr(a, "warmup_prop", 0.05),
                             getattr(a, "gradient_accumulation_steps", 1))

    def _lr(self):
        return self.args.lr * linear_schedule(self.global_step, self.warmup_steps, self.total_steps)

    def train_batch(self, batch):
        """one optimisation step on a collator batch (input_ids, attention, whole_word_ids, output_ids, output_attention)"""
        a = self.args
        loss = self.model.train_step(batch[0], batch[2], batch[1], batch[3], batch[4], lr=self._lr(),
                                     clip=getattr(a, "clip", 1.0), eps=getattr(a, "adam_eps", 1e-6),
                                     weight_decay=getattr(a, "weight_decay", 0.01),
                                     overlap_optimizer=True)   # the loop never reads parameters between steps
        self.global_step += 1
        return loss

    def train(self):
        import torch.distributed as dist
        a = self.args
        self.model.zero_grad()
        for epoch in range(a.epochs):
            if hasattr(self.train_loader, "sampler") and hasattr(self.train_loader.sampler, "set_epoch"):
                self.train_loader.sampler.set_epoch(epoch)
            self.model.train()
            losses = []
            for batch in self.train_loader:
                losses.append(self.train_batch(batch))
            ep = torch.stack([l.reshape(()) for l in losses]).mean() if losses else torch.zeros((), device=self.device)
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(ep, op=dist.ReduceOp.SUM)
                ep /= dist.get_world_size()
            if self.rank == 0:
                logging.info(f"The average training loss for epoch {epoch + 1} is {ep.item()}")
            if getattr(a, "valid_select", 0) > 0 and self.valid_loader is not None:
                v = self.validate()
                if self.rank == 0:
                    logging.info(f"The average valid loss for epoch {epoch + 1} is {v}")
            if getattr(a, "test_epoch", 0) > 0 and (epoch + 1) % a.test_epoch == 0 and self.testloaders:
                self.test()
        if self.rank == 0 and getattr(a, "model_path", None):
            torch.save(self.model.state_dict(), a.model_path)   # ref DistributedRunner.py:169
        return

    def validate(self) -> float:
        """ref DistributedRunner.py:105-156: masked mean loss over the validation loader"""
        import torch.distributed as dist
        self.model.eval()
        tot = []
        with torch.no_grad():
            for batch in self.valid_loader:
                out = self.model(input_ids=batch[0], whole_word_ids=batch[2], attention_mask=batch[1], labels=batch[3])
                B, L = batch[3].shape
                m = (batch[4].to(out["loss"].device) != 0).float()
                l = out["loss"].view(B, L) * m
                tot.append((l.sum(1) / m.sum(1).clamp(min=1)).mean())
        v = torch.stack(tot).mean() if tot else torch.zeros((), device=self.device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            v /= dist.get_world_size()
        return v.item()

    def test(self, path=None):
        self.model.eval()
        if path:
            self.model.load_state_dict(torch.load(path, map_location="cpu"))
        res = []
        for loader in self.testloaders:
            res.append(self.test_dataset_task(loader))
        return res

    def candidate_paths(self, dataset_name: str, candidates) -> List[List[int]]:
        """ref DistributedRunner.py:344-351: [0] + tokenizer.encode(f"{dataset} item_{candidate}")"""
        return [[0] + list(self.tokenizer.encode(f"{dataset_name} item_{c}")) for c in candidates]

    def test_dataset_task(self, testloader, paths=None):
        """ref DistributedRunner.py:339-399: constrained beam search over all items, HR@k / NDCG@k"""
        ds = testloader.dataset
        if paths is None:
            paths = self.candidate_paths(ds.dataset, ds.all_items)
        trie = self.model.build_trie(paths)
        sums = [0.0] * len(self.metrics)
        total = 0
        K = self.generate_num
        # metrics on the device: the generated sequences never leave the GPU, one small D2H read at the very end
        kinds = [m.lower().split("@") for m in self.metrics]
        ks = sorted({int(k) for _, k in kinds})
        dev_sums = None
        for batch in testloader:
            pred = self.model.generate(input_ids=batch[0], attention_mask=batch[1], whole_word_ids=batch[2], max_length=50,
                                       trie=trie, num_beams=K, num_return_sequences=K)
            dev_sums = self.model.eval_metric_sums(pred["sequences"], pred["sequences_scores"], batch[3], K, ks, out=dev_sums)
            total += int(batch[3].shape[0])
        if dev_sums is not None:
            host = dev_sums.tolist()
            sums = [host[(0 if kind.startswith("hit") else len(ks)) + ks.index(int(k))] for kind, k in kinds]
        res, n = allreduce_metrics(sums, total, device=self.device)
        if self.rank == 0:
            for name, v in zip(self.metrics, res):
                logging.info(f"{name}: {v}")
        return dict(zip(self.metrics, res)), n
