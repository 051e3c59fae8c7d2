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


def rel_results_filtered(seqs, scores, gold, rows_per_user: int, positives, n_positives, k: int) -> List[List[int]]:
    """ref utils/evaluate.py:6-35 on token-id paths: per user the `rows_per_user` predictions sorted by score (stable,
    desc); predictions that are one of the user's positive items are skipped; the first k survivors are marked 0/1
    against the gold item.  positives [B, Pmax, Tp] (first n_positives[b] rows valid)."""
    out = []
    for b in range(len(gold)):
        g = strip_special(gold[b])
        pos = {strip_special(positives[b][j]) for j in range(int(n_positives[b]))}
        pairs = [(strip_special(seqs[b * rows_per_user + i]), float(scores[b * rows_per_user + i])) for i in range(rows_per_user)]
        pairs = sorted(pairs, key=lambda x: x[1], reverse=True)
        one = []
        for p, _ in pairs:
            if p in pos:
                continue
            one.append(1 if p == g else 0)
            if len(one) >= k:
                break
        out.append(one)
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
def _ref_import(module: str, name: str):
    """the reference's own host classes (TestDataset, Collator, TestCollator) are imported under their in-tree names
    (`from data.TestDataset import TestDataset`, ref runner/DistributedRunner.py:11-14) — the drop-in runs with
    src/src_t5 on sys.path exactly like the reference's main.py"""
    import importlib
    return getattr(importlib.import_module(module), name)


class B200Runner:
    """same constructor and public methods as ref DistributedRunner (runner/DistributedRunner.py:21-27 on top of
    runner/SingleRunner.py:41-64): train(), test(path), get_testloader(), test_dataset_task*(loader)."""

    def __init__(self, model, tokenizer, train_loader, valid_loader, device, args, rank: int = 0, test_dataset_cls=None,
                 collator_cls=None, test_collator_cls=None):
        self.model, self.tokenizer = model, tokenizer
        self.train_loader, self.valid_loader = train_loader, valid_loader
        self.device, self.args, self.rank = device, args, rank
        g = lambda k, dflt: getattr(args, k, dflt)
        # ref SingleRunner.py:48-55
        ds0 = None
        try:
            ds0 = self.train_loader.dataset.datasets[0]
        except Exception:
            pass
        self.regenerate_candidate = bool(ds0 is not None and "candidate_items" in getattr(ds0, "info", ()))
        self.reconstruct_data = g("sample_prompt", 0)
        self.test_epoch = g("test_epoch", 0)
        self.valid_select = g("valid_select", 0)
        self.test_before_train = g("test_before_train", 0)
        self.test_filtered = g("test_filtered", 0)
        self.test_filtered_batch = g("test_filtered_batch", 1)
        self._cls = dict(test_dataset=test_dataset_cls, collator=collator_cls, test_collator=test_collator_cls)
        self.metrics = g("metrics", "hit@5,hit@10,ndcg@5,ndcg@10").split(",")
        self.generate_num = max(int(m.split("@")[1]) for m in self.metrics)
        self.testloaders = []
        self.get_testloader()
        self.global_step = 0
        self.start_epoch = 0
        self.total_steps, self.warmup_steps = (0, 0)
        if g("train", 1):
            self.total_steps, self.warmup_steps = self.create_optimizer_and_scheduler()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and model.world_size == 1:
            model.init_data_parallel()       # replaces the DDP wrapper of ref DistributedRunner.py:26
        if g("resume", None):
            self.load_checkpoint(args.resume)

    # ------------------------------------------------------------------ optimiser / schedule (host side)
    def create_optimizer_and_scheduler(self):
        """ref SingleRunner.py:178-219.  AdamW state lives in the engine (flat m / v buffers); what remains on the host
        is the step count of get_linear_schedule_with_warmup.  The reference's two parameter groups (no weight decay
        for names containing "bias" = the relative_attention_bias tables) are built into p5_adamw_step."""
        a = self.args
        n = len(self.train_loader) if self.train_loader is not None else 0
        return schedule_plan(n, getattr(a, "epochs", 1), getattr(a, "warmup_prop", 0.05),
                             getattr(a, "gradient_accumulation_steps", 1))

    def _lr(self):
        return self.args.lr * linear_schedule(self.global_step, self.warmup_steps, self.total_steps)

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self, epoch: int):
        """everything a run needs to continue bit-for-bit: weights (HF key names, what the reference saves at
        DistributedRunner.py:155,169), Adam moments + step, scheduler position, epoch"""
        return {"model": self.model.state_dict(), "optimizer": self.model.optimizer_state_dict(),
                "global_step": self.global_step, "epoch": epoch, "total_steps": self.total_steps,
                "warmup_steps": self.warmup_steps}

    def save_checkpoint(self, path: str, epoch: int):
        if self.rank == 0:
            torch.save(self.state_dict(epoch), path)

    def load_checkpoint(self, path: str):
        ck = torch.load(path, map_location="cpu")
        if "model" not in ck:                       # a plain reference state_dict (weights only)
            self.model.load_state_dict(ck, strict=False)
            return
        self.model.load_state_dict(ck["model"])
        self.model.load_optimizer_state_dict(ck["optimizer"])
        self.global_step = int(ck["global_step"])
        self.start_epoch = int(ck["epoch"])

    # ------------------------------------------------------------------ training
    def _staged(self, loader):
        """tokenisation + collation (the reference's Collator, unchanged) run on a background thread, batches
        arrive through pinned double buffers and one H2D copy each (openp5_b200/pipeline.py); args.stage_batches=0 keeps the
        reference's inline loop"""
        if getattr(self.args, "stage_batches", 1) and torch.cuda.is_available() and str(self.device).startswith("cuda"):
            from .pipeline import BatchStager
            return BatchStager(loader, self.device)
        return loader

    def train_batch(self, batch):
        """one optimisation step on a collator batch (input_ids, attention, whole_word_ids, output_ids, output_attention):
        ref DistributedRunner.py:56-87"""
        a = self.args
        loss = self.model.train_step(batch[0], batch[2], batch[1], batch[3], batch[4], lr=self._lr(),
                                     clip=getattr(a, "clip", 1.0), eps=getattr(a, "adam_eps", 1e-6),
                                     weight_decay=getattr(a, "weight_decay", 0.01), enc_lengths=getattr(batch, "enc_lengths", None),
                                     overlap_optimizer=True)   # the loop never reads parameters between steps
        self.global_step += 1
        return loss

    def train(self):
        """ref DistributedRunner.py:28-177"""
        import torch.distributed as dist
        a = self.args
        ddp = dist.is_available() and dist.is_initialized()
        self.model.zero_grad()
        valid_losses = []
        best_epoch = -1
        if self.test_before_train > 0:
            self.test()
        # resume: the reference's samplers shuffle the task index lists IN PLACE, cumulatively epoch after epoch
        # (MultiTaskDataset.py:189-195 from SingleMultiDataTaskSampler.py:31-33), so epoch e's order depends on the shuffles
        # of epochs 0..e-1: replay them (index work only) before continuing
        sampler = getattr(self.train_loader, "sampler", None)
        for e in range(self.start_epoch):
            if hasattr(sampler, "set_epoch"):
                sampler.set_epoch(e)
                for _ in iter(sampler):
                    break
        for epoch in range(self.start_epoch, a.epochs):
            if self.rank == 0:
                logging.info(f"Start training for epoch {epoch + 1}")
            # per-epoch regeneration of candidates / sampled prompts (ref :42-49)
            datasets = getattr(getattr(self.train_loader, "dataset", None), "datasets", [])
            if self.regenerate_candidate:
                for ds in datasets:
                    ds.generate_candidates()
                    ds.construct_sentence()
            elif self.reconstruct_data:
                for ds in datasets:
                    ds.construct_sentence()
            if hasattr(self.train_loader, "sampler") and hasattr(self.train_loader.sampler, "set_epoch"):
                self.train_loader.sampler.set_epoch(epoch)
            self.model.train()
            losses = []
            for batch in self._staged(self.train_loader):
                losses.append(self.train_batch(batch))
            ep = torch.stack([l.reshape(()) for l in losses]).mean() if losses else torch.zeros((), device=self.device)
            if ddp:
                dist.all_reduce(ep, op=dist.ReduceOp.SUM)
                ep /= dist.get_world_size()
            self.last_train_loss = ep.item()
            if self.rank == 0:
                logging.info(f"The average training loss for epoch {epoch + 1} is {self.last_train_loss}")
            if self.valid_select > 0 and self.valid_loader is not None:
                if self.rank == 0:
                    logging.info(f"Start validation for epoch {epoch + 1}")
                v = self.validate()
                valid_losses.append(v)
                if self.rank == 0:
                    logging.info(f"The average valid loss for epoch {epoch + 1} is {v}")
                    if v == min(valid_losses):                                   # ref :152-156
                        logging.info("The minimal validation loss so far.")
                        best_epoch = epoch + 1
                        if getattr(a, "model_path", None):
                            torch.save(self.model.state_dict(), a.model_path)
                            logging.info(f"Save the current model to {a.model_path}")
            if getattr(a, "checkpoint_path", None):
                self.save_checkpoint(a.checkpoint_path, epoch + 1)
            if self.test_epoch > 0 and (epoch + 1) % self.test_epoch == 0:
                self.model.eval()
                self.test()
            if ddp:
                dist.barrier()
        if self.valid_select > 0:
            if self.rank == 0:
                logging.info(f"The best validation at Epoch {best_epoch}")
        elif self.rank == 0 and getattr(a, "model_path", None):
            torch.save(self.model.state_dict(), a.model_path)                    # ref :167-170
            logging.info(f"Save the current model to {a.model_path}")
        self.best_epoch = best_epoch
        return

    def validate(self) -> float:
        """ref DistributedRunner.py:105-150: masked mean loss over the validation loader"""
        import torch.distributed as dist
        self.model.eval()
        if getattr(self.args, "valid_prompt_sample", 0) > 0:
            for ds in getattr(getattr(self.valid_loader, "dataset", None), "datasets", []):
                ds.construct_sentence()
        tot = []
        with torch.no_grad():
            for batch in self.valid_loader:
                out = self.model(input_ids=batch[0], whole_word_ids=batch[2], attention_mask=batch[1], labels=batch[3],
                                 return_logits=False)
                B, L = batch[3].shape
                m = (batch[4].to(out["loss"].device) != 0).float()
                l = out["loss"].view(B, L) * m
                tot.append((l.sum(1) / m.sum(1).clamp(min=1)).mean())
        v = torch.stack(tot).mean() if tot else torch.zeros((), device=self.device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            v /= dist.get_world_size()
        return v.item()

    # ------------------------------------------------------------------ evaluation
    def get_testloader(self):
        """ref DistributedRunner.py:179-192: one loader per (dataset, task): the reference's TestDataset, sharded over ranks
        by DistributedSampler (default shuffle=True, as in the reference), collated by its Collator / TestCollator."""
        self.testloaders = []
        a = self.args
        if not (hasattr(a, "datasets") and hasattr(a, "tasks")):
            return
        from torch.utils.data import DataLoader
        from torch.utils.data.distributed import DistributedSampler
        import torch.distributed as dist
        try:
            TestDataset = self._cls["test_dataset"] or _ref_import("data.TestDataset", "TestDataset")
            Coll = self._cls["collator"] or _ref_import("processor.Collator", "Collator")
            TestColl = self._cls["test_collator"] or _ref_import("processor.Collator", "TestCollator")
        except ImportError as ex:
            logging.warning("B200Runner: the reference's data/TestDataset.py and processor/Collator.py are not importable "
                            "(%s): no test loaders were built; runner.test() will raise", ex)
            return
        collator = TestColl(self.tokenizer) if self.test_filtered > 0 else Coll(self.tokenizer)
        ddp = dist.is_available() and dist.is_initialized()
        for dataset in a.datasets.split(","):
            for task in a.tasks.split(","):
                testdata = TestDataset(a, dataset, task)
                sampler = DistributedSampler(testdata) if ddp else DistributedSampler(testdata, num_replicas=1, rank=0)
                self.testloaders.append(DataLoader(dataset=testdata, sampler=sampler, batch_size=a.eval_batch_size,
                                                   collate_fn=collator, shuffle=False))

    def test(self, path=None):
        """ref DistributedRunner.py:194-207"""
        self.model.eval()
        if path:
            sd = torch.load(path, map_location="cpu")
            self.model.load_state_dict(sd["model"] if "model" in sd else sd)
        if not self.testloaders:
            raise RuntimeError("B200Runner.test(): no test loaders (args.datasets / args.tasks missing, or the reference's "
                               "TestDataset is not importable) — nothing would be evaluated")
        res = []
        for loader in self.testloaders:
            if self.test_filtered > 0:
                if self.test_filtered_batch > 0:
                    res.append(self.test_dataset_task_filtered_batch(loader))
                else:
                    assert self.args.eval_batch_size == 1
                    res.append(self.test_dataset_task_filtered(loader))
            else:
                res.append(self.test_dataset_task(loader))
        return res

    def candidate_paths(self, dataset_name: str, candidates) -> List[List[int]]:
        """ref DistributedRunner.py:344-351: [0] + tokenizer.encode(f"{dataset} item_{candidate}")"""
        return [[0] + list(self.tokenizer.encode(f"{dataset_name} item_{c}")) for c in candidates]

    def _metric_plan(self):
        kinds = [m.lower().split("@") for m in self.metrics]
        ks = sorted({int(k) for _, k in kinds})
        return kinds, ks

    def _finish(self, dev_sums, total, kinds, ks):
        sums = [0.0] * len(self.metrics)
        if dev_sums is not None:
            host = dev_sums.tolist()       # the one D2H read of the loader
            sums = [host[(0 if kind.startswith("hit") else len(ks)) + ks.index(int(k))] for kind, k in kinds]
        res, n = allreduce_metrics(sums, total, device=self.device)
        if self.rank == 0:
            for name, v in zip(self.metrics, res):
                logging.info(f"{name}: {v}")
        return dict(zip(self.metrics, res)), n

    def test_dataset_task(self, testloader, paths=None, prefix_allowed_tokens_fn=None):
        """ref DistributedRunner.py:339-399: constrained beam search over all items, HR@k / NDCG@k.  The candidate trie is
        flattened to the device CSR form once per loader (or recovered from a reference `prefix_allowed_tokens_fn`); the
        generated sequences never leave the GPU: metrics are reduced there, one small D2H read at the end."""
        ds = testloader.dataset
        if self.rank == 0:
            logging.info(f"testing {ds.dataset} dataset on {ds.task} task")
        gen_kw = {}
        if prefix_allowed_tokens_fn is not None:
            gen_kw["prefix_allowed_tokens_fn"] = prefix_allowed_tokens_fn
        else:
            gen_kw["trie"] = self.model.build_trie(paths if paths is not None else self.candidate_paths(ds.dataset, ds.all_items))
        K = self.generate_num
        kinds, ks = self._metric_plan()
        dev_sums, total = None, 0
        for batch in testloader:
            pred = self.model.generate(input_ids=batch[0], attention_mask=batch[1], whole_word_ids=batch[2], max_length=50,
                                       num_beams=K, num_return_sequences=K, **gen_kw)
            dev_sums = self.model.eval_metric_sums(pred["sequences"], pred["sequences_scores"], batch[3], K, ks, out=dev_sums)
            total += int(batch[3].shape[0])
        return self._finish(dev_sums, total, kinds, ks)

    def _positive_paths(self, ds, user_idx, cap):
        """token paths of the strings in ds.positive[user] (what ref evaluate.rel_results_filtered compares the decoded
        predictions with, utils/evaluate.py:6-35), padded to [B, cap, Tp]"""
        rows, counts, width = [], [], 1
        for u in user_idx:
            pos = [list(self.tokenizer.encode(str(p))) for p in sorted(ds.positive[ds.id2user[int(u)]])][:cap]
            rows.append(pos)
            counts.append(len(pos))
            width = max([width] + [len(p) for p in pos])
        out = torch.zeros((len(rows), max(cap, 1), width), dtype=torch.int32)
        for b, pos in enumerate(rows):
            for j, p in enumerate(pos):
                out[b, j, : len(p)] = torch.tensor(p, dtype=torch.int32)
        return out, torch.tensor(counts, dtype=torch.int32)

    def test_dataset_task_filtered_batch(self, testloader):
        """ref DistributedRunner.py:209-270: beams = generate_num + max_positive, max_length 30, predictions that are
        positives of the user are skipped before the top-generate_num relevance list is formed"""
        ds = testloader.dataset
        if self.rank == 0:
            logging.info(f"testing filtered {ds.dataset} dataset on {ds.task} task")
        trie = self.model.build_trie(self.candidate_paths(ds.dataset, sorted(set(ds.all_items))))
        R = self.generate_num + int(getattr(ds, "max_positive", 0))
        kinds, ks = self._metric_plan()
        dev_sums, total = None, 0
        for batch in testloader:
            pred = self.model.generate(input_ids=batch[0], attention_mask=batch[1], whole_word_ids=batch[2], max_length=30,
                                       trie=trie, num_beams=R, num_return_sequences=R)
            pos, npos = self._positive_paths(ds, batch[5].tolist(), int(getattr(ds, "max_positive", 0)))
            dev_sums = self.model.eval_metric_sums_filtered(pred["sequences"], pred["sequences_scores"], batch[3], R, ks,
                                                            pos, npos, self.generate_num, out=dev_sums)
            total += int(batch[3].shape[0])
        return self._finish(dev_sums, total, kinds, ks)

    def test_dataset_task_filtered(self, testloader):
        """ref DistributedRunner.py:272-337 (eval_batch_size 1): the user's own history is REMOVED from the candidate
        trie, then plain beam search + unfiltered metrics"""
        ds = testloader.dataset
        if self.rank == 0:
            logging.info(f"testing filtered {ds.dataset} dataset on {ds.task} task")
        candidates = set(ds.all_items)
        K = self.generate_num
        kinds, ks = self._metric_plan()
        dev_sums, total = None, 0
        for batch in testloader:
            user_idx = int(batch[5][0])
            user_candidate = candidates - ds.positive[ds.id2user[user_idx]]
            trie = self.model.build_trie(self.candidate_paths(ds.dataset, sorted(user_candidate)))
            pred = self.model.generate(input_ids=batch[0], attention_mask=batch[1], whole_word_ids=batch[2], max_length=30,
                                       trie=trie, num_beams=K, num_return_sequences=K)
            dev_sums = self.model.eval_metric_sums(pred["sequences"], pred["sequences_scores"], batch[3], K, ks, out=dev_sums)
            total += int(batch[3].shape[0])
        return self._finish(dev_sums, total, kinds, ks)
