"""Generate tests/golden/dropin_ml100k.npz (run in the build container, where /root/reference exists): the collated
batches that the reference's UNMODIFIED host pipeline produces for BASELINE configs[0] (ML-100K-shaped synthetic
user sequences, sequential indexing, batch 4) — MultiTaskDataset -> DistMultiDataTaskSampler -> Collator for training /
validation, TestDataset -> DistributedSampler -> Collator for evaluation, and the candidate item paths that
DistributedRunner.py:344-351 puts into the Trie.  tests/test_dropin_gpu.py feeds them through B200Runner on the engine.
Usage:  python tests/golden/make_dropin_fixture.py"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import dropin_common as D  # noqa: E402
from standin_tokenizer import StandInTokenizer  # noqa: E402


def pad_stack(rows, width):
    out = np.zeros((len(rows), rows[0].shape[0], width), dtype=np.int32)
    for i, r in enumerate(rows):
        out[i, :, : r.shape[1]] = r
    return out


def main():
    from openp5_b200.runner import B200Runner
    tmp = tempfile.mkdtemp(prefix="p5_fixture_")
    D.write_user_sequences(tmp)
    ref = D.import_reference()
    args = D.make_args(tmp, train=0)
    tok = StandInTokenizer(vocab_size=32100)
    train_loader, valid_loader = D.build_loaders(ref, args, tok)
    train_loader.sampler.set_epoch(0)
    out = {}
    for name, loader, n in (("train", train_loader, 12), ("valid", valid_loader, 4)):
        batches = []
        for i, b in enumerate(loader):
            if i >= n:
                break
            batches.append([t.numpy().astype(np.int32) for t in b])
        Le = max(b[0].shape[1] for b in batches)
        Ld = max(b[3].shape[1] for b in batches)
        for j, key, w in ((0, "ids", Le), (1, "attn", Le), (2, "ww", Le), (3, "labels", Ld), (4, "oattn", Ld)):
            out[f"{name}_{key}"] = pad_stack([b[j] for b in batches], w)
        out[f"{name}_Le"] = np.array([b[0].shape[1] for b in batches], dtype=np.int32)     # the collator pads per batch
        out[f"{name}_Ld"] = np.array([b[3].shape[1] for b in batches], dtype=np.int32)
    # evaluation batches through the runner's own get_testloader (reference TestDataset + DistributedSampler + Collator)
    r = B200Runner(None, tok, train_loader, valid_loader, "cpu", args, rank=0)
    loader = r.testloaders[0]
    loader.sampler.set_epoch(0)
    tb = [[t.numpy().astype(np.int32) for t in b] for b in loader]
    tb = [b for b in tb if b[0].shape[0] == args.eval_batch_size]
    Le = max(b[0].shape[1] for b in tb)
    Ld = max(b[3].shape[1] for b in tb)
    for j, key, w in ((0, "ids", Le), (1, "attn", Le), (2, "ww", Le), (3, "labels", Ld), (4, "oattn", Ld)):
        out[f"test_{key}"] = pad_stack([b[j] for b in tb], w)
    out["test_Le"] = np.array([b[0].shape[1] for b in tb], dtype=np.int32)
    ds = loader.dataset
    paths = r.candidate_paths(ds.dataset, ds.all_items)
    out["item_paths"] = pad_stack([np.array([p], dtype=np.int32) for p in paths], max(len(p) for p in paths))[:, 0, :]
    out["item_path_len"] = np.array([len(p) for p in paths], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "dropin_ml100k.npz"), **out)
    print({k: v.shape for k, v in out.items()})
    print("first train input :", tok.decode(out["train_ids"][0, 0].tolist()))
    print("first train output:", tok.decode(out["train_labels"][0, 0].tolist()))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
