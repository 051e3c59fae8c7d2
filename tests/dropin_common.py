"""Shared pieces of the drop-in tests (TEST INFRASTRUCTURE): a synthetic `user_sequence.txt`, the argparse namespace the
reference's host code reads, and the reference's own loader construction (main.py:24-66 restated — main.py itself
cannot be imported under transformers 5.x because it imports the HF-4.26-only model class).
Everything under /root/reference is used UNMODIFIED and only here in the build container (never on the GPU box)."""
import argparse
import os
import random
import sys

REF = "/root/reference/src/src_t5"
PROMPT_FILE = "/root/reference/prompt.txt"


def have_reference():
    return os.path.isdir(REF) and os.path.exists(PROMPT_FILE)


def import_reference():
    """the reference's modules under their in-tree names (what `cd src/src_t5; python main.py` sees)"""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    mods = {}
    for name in ("data.MultiTaskDataset", "data.TestDataset", "processor.Collator", "processor.DistMultiDataTaskSampler",
                 "processor.SingleMultiDataTaskSampler", "utils.generation_trie", "utils.evaluate", "utils.indexing"):
        mods[name.split(".")[-1]] = importlib.import_module(name)
    return argparse.Namespace(**mods)


def write_user_sequences(data_dir, dataset="ML100K", n_users=52, n_items=40, seed=2023):
    """`user item item ...` lines like the reference's preprocessed data (README.md:46): >= 5 interactions per user"""
    rng = random.Random(seed)
    os.makedirs(os.path.join(data_dir, dataset), exist_ok=True)
    lines = []
    for u in range(1, n_users + 1):
        n = rng.randrange(5, 7)     # (the reference logs validation sample 101: it needs > 50 users, MultiTaskDataset.py:305-306)
        items = rng.sample(range(1, n_items + 1), n)
        lines.append(" ".join([str(u)] + [str(1000 + i) for i in items]))
    with open(os.path.join(data_dir, dataset, "user_sequence.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return lines


def make_args(data_dir, **over):
    """every attribute the reference's MultiTaskDataset / TestDataset / samplers / runner read (their argparse defaults:
    MultiTaskDataset.py:19-52, SingleRunner.py:14-39, SingleMultiDataTaskSampler.py:8-16, utils/utils.py:12-23), with
    BASELINE configs[0] values: ML-100K, sequential indexing, batch 4"""
    a = dict(data_path=data_dir, item_indexing="sequential", tasks="sequential,straightforward", datasets="ML100K",
             prompt_file=PROMPT_FILE, sequential_order="original", collaborative_token_size=200, collaborative_cluster=20,
             collaborative_last_token="sequential", collaborative_float32=0, max_his=10, his_prefix=1, his_sep=" , ",
             skip_empty_his=1, valid_prompt="seen:0", valid_prompt_sample=0, valid_sample_num="3,3", test_prompt="seen:0",
             sample_prompt=0, sample_num="2,2", batch_size=4, eval_batch_size=4, dist_sampler=0, seed=2023, distributed=0,
             rank=0, world_size=1, gpu=0, optim="AdamW", epochs=1, lr=1e-3, clip=1.0, logging_step=100, warmup_prop=0.05,
             gradient_accumulation_steps=1, weight_decay=0.01, adam_eps=1e-6, dropout=0.1, alpha=2, train=1, backbone="t5-small",
             metrics="hit@5,hit@10,ndcg@5,ndcg@10", load=0, random_initialize=1, test_epoch=0, valid_select=1,
             test_before_train=0, test_filtered=0, test_filtered_batch=1, model_path=None)
    a.update(over)
    return argparse.Namespace(**a)


def build_loaders(ref, args, tokenizer):
    """ref main.py:24-66 (get_dataset + get_loader) with the reference's own classes"""
    from torch.utils.data import ConcatDataset, DataLoader
    MTD = ref.MultiTaskDataset.MultiTaskDataset
    train_sets = [MTD(args, d, "train") for d in args.datasets.split(",")]
    TrainSet = ConcatDataset(train_sets)
    ValidSet = ConcatDataset([MTD(args, d, "validation") for d in args.datasets.split(",")]) if args.valid_select > 0 else None
    sampler = ref.DistMultiDataTaskSampler.DistMultiDataTaskSampler(TrainSet, args.batch_size, args.world_size, args.rank,
                                                                    args.seed, shuffle=True)
    collator = ref.Collator.Collator(tokenizer)
    train_loader = DataLoader(dataset=TrainSet, sampler=sampler, batch_size=args.batch_size, collate_fn=collator, shuffle=False)
    valid_loader = (DataLoader(dataset=ValidSet, sampler=None, batch_size=args.batch_size, collate_fn=collator, shuffle=False)
                    if ValidSet is not None else None)
    return train_loader, valid_loader
