"""BASELINE configs[0] ("ML-100K sequential-indexing T5-small-class model, batch=4, reference plumbing, no GPU"):
the reference's UNMODIFIED host code — utils/indexing.sequential_indexing, data/MultiTaskDataset, data/TestDataset,
processor/DistMultiDataTaskSampler, processor/Collator, utils/generation_trie (Trie + prefix_allowed_tokens_fn),
utils/evaluate — drives openp5_b200.runner.B200Runner exactly as its own main.py drives DistributedRunner
(main.py:150-212, runner/DistributedRunner.py:23-28,339-399).  The model behind the runner is the CPU oracle
(tests/oracle_model.py), so this runs in the -m "not gpu" suite; tests/test_dropin_gpu.py runs the same runner on the
real engine over batches this pipeline produced (tests/golden/dropin_ml100k.npz).

Needs /root/reference (build container only): skipped elsewhere."""
import os
import shutil
import sys
import tempfile

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dropin_common as D  # noqa: E402
from standin_tokenizer import StandInTokenizer  # noqa: E402

pytestmark = pytest.mark.skipif(not D.have_reference(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def pipeline():
    from oracle import p5_oracle as po
    from oracle_model import OracleModel
    tmp = tempfile.mkdtemp(prefix="p5_dropin_")
    D.write_user_sequences(tmp)
    ref = D.import_reference()
    args = D.make_args(tmp, model_path=os.path.join(tmp, "model.pt"), checkpoint_path=os.path.join(tmp, "ckpt.pt"))
    tok = StandInTokenizer(vocab_size=1200)
    train_loader, valid_loader = D.build_loaders(ref, args, tok)
    cfg = po.t5_cfg("t5-tiny", vocab_size=len(tok))
    model = OracleModel(cfg, po.init_weights(cfg, seed=2023))
    yield dict(tmp=tmp, ref=ref, args=args, tok=tok, train_loader=train_loader, valid_loader=valid_loader, model=model, cfg=cfg)
    shutil.rmtree(tmp, ignore_errors=True)


def test_reference_plumbing_runs_through_b200runner(pipeline):
    from openp5_b200.runner import B200Runner
    p = pipeline
    r = B200Runner(p["model"], p["tok"], p["train_loader"], p["valid_loader"], "cpu", p["args"], rank=0)
    # get_testloader built one loader per (dataset, task) from the reference's TestDataset + Collator
    assert len(r.testloaders) == 2
    assert type(r.testloaders[0].dataset).__name__ == "TestDataset"
    n_batches = len(p["train_loader"])
    assert (r.total_steps, r.warmup_steps) == (n_batches * p["args"].epochs, int(n_batches * p["args"].epochs * 0.05))
    r.train()
    assert p["model"].calls["train_step"] == n_batches and r.global_step == n_batches
    assert np.isfinite(r.last_train_loss)
    assert r.best_epoch == 1 and os.path.exists(p["args"].model_path)          # valid_select > 0: min-valid-loss epoch saved
    sd = torch.load(p["args"].model_path)
    assert "shared.weight" in sd and "encoder.whole_word_embeddings.weight" in sd and "lm_head.weight" in sd
    ck = torch.load(p["args"].checkpoint_path)
    assert ck["global_step"] == n_batches and ck["epoch"] == 1 and "exp_avg" in next(iter(ck["optimizer"]["state"].values()))
    # every batch was task-homogeneous and came from the reference sampler's rank shard
    res = r.test()
    assert len(res) == 2 and all(set(m) == set(r.metrics) for m, _ in res)
    n_users = len(r.testloaders[0].dataset)
    assert all(n == n_users for _, n in res)


def test_runner_metrics_equal_reference_string_metrics(pipeline):
    """B200Runner.test_dataset_task (token-id paths) == the reference's own flow on decoded strings
    (DistributedRunner.py:361-393: generate -> batch_decode -> evaluate.rel_results -> get_metrics_results), for the
    verbatim call with the reference's opaque `prefix_allowed_tokens_fn` closure and for the explicit trie."""
    from openp5_b200.runner import B200Runner
    p = pipeline
    ref, tok, model = p["ref"], p["tok"], p["model"]
    r = B200Runner(model, tok, p["train_loader"], p["valid_loader"], "cpu", p["args"], rank=0)
    loader = r.testloaders[0]
    ds = loader.dataset
    gt, ev = ref.generation_trie, ref.evaluate
    candidate_trie = gt.Trie([[0] + tok.encode(f"{ds.dataset} item_{c}") for c in ds.all_items])
    fn = gt.prefix_allowed_tokens_fn(candidate_trie)
    K = r.generate_num
    want = np.zeros(len(r.metrics))
    total = 0
    loader.sampler.set_epoch(0)
    for batch in loader:
        pred = model.generate(input_ids=batch[0], attention_mask=batch[1], whole_word_ids=batch[2], max_length=50,
                              prefix_allowed_tokens_fn=fn, num_beams=K, num_return_sequences=K)
        gold = tok.batch_decode(batch[3], skip_special_tokens=True)
        gen = tok.batch_decode(pred["sequences"], skip_special_tokens=True)
        rel = ev.rel_results(gen, gold, pred["sequences_scores"], K)
        want += ev.get_metrics_results(rel, r.metrics)
        total += len(rel)
    want /= total
    loader.sampler.set_epoch(0)
    got_cb, n1 = r.test_dataset_task(loader, prefix_allowed_tokens_fn=fn)
    loader.sampler.set_epoch(0)
    got_trie, n2 = r.test_dataset_task(loader)
    assert n1 == n2 == total
    assert np.allclose([got_cb[m] for m in r.metrics], want, atol=1e-6)
    assert np.allclose([got_trie[m] for m in r.metrics], want, atol=1e-6)


def test_filtered_variants_equal_reference(pipeline):
    """test_filtered=1: the reference's TestCollator batches (user_idx column) through both filtered variants;
    the batch variant is compared with evaluate.rel_results_filtered on decoded strings (DistributedRunner.py:209-270)"""
    from openp5_b200.runner import B200Runner
    p = pipeline
    args = D.make_args(p["tmp"], test_filtered=1, test_filtered_batch=1, eval_batch_size=4, train=0)
    r = B200Runner(p["model"], p["tok"], p["train_loader"], p["valid_loader"], "cpu", args, rank=0)
    loader = r.testloaders[0]
    ds = loader.dataset
    ev, tok, model = p["ref"].evaluate, p["tok"], p["model"]
    R = r.generate_num + ds.max_positive
    trie = model.build_trie(r.candidate_paths(ds.dataset, sorted(set(ds.all_items))))
    want, total = np.zeros(len(r.metrics)), 0
    loader.sampler.set_epoch(0)
    for batch in loader:
        pred = model.generate(input_ids=batch[0], attention_mask=batch[1], whole_word_ids=batch[2], max_length=30, trie=trie,
                              num_beams=R, num_return_sequences=R)
        gold = tok.batch_decode(batch[3], skip_special_tokens=True)
        gen = tok.batch_decode(pred["sequences"], skip_special_tokens=True)
        rel = ev.rel_results_filtered(ds.positive, ds.id2user, batch[5].numpy(), R, gen, gold, pred["sequences_scores"], r.generate_num)
        want += ev.get_metrics_results(rel, r.metrics)
        total += len(rel)
    want /= total
    loader.sampler.set_epoch(0)
    got, n = r.test_dataset_task_filtered_batch(loader)
    assert n == total and np.allclose([got[m] for m in r.metrics], want, atol=1e-6)
    # per-user candidate tries (eval_batch_size 1)
    args1 = D.make_args(p["tmp"], test_filtered=1, test_filtered_batch=0, eval_batch_size=1, train=0, datasets="ML100K",
                        tasks="sequential")
    r1 = B200Runner(p["model"], p["tok"], p["train_loader"], p["valid_loader"], "cpu", args1, rank=0)
    res = r1.test()
    assert len(res) == 1 and res[0][1] == len(r1.testloaders[0].dataset)


def test_resume_continues_where_the_checkpoint_stopped(pipeline):
    from oracle import p5_oracle as po
    from oracle_model import OracleModel
    from openp5_b200.runner import B200Runner
    p = pipeline
    tmp = p["tmp"]
    a2 = D.make_args(tmp, epochs=2, valid_select=0, checkpoint_path=os.path.join(tmp, "ck2.pt"))
    m_full = OracleModel(p["cfg"], po.init_weights(p["cfg"], seed=7))
    tl, vl = D.build_loaders(p["ref"], a2, p["tok"])
    B200Runner(m_full, p["tok"], tl, None, "cpu", a2).train()
    # interrupted after epoch 1, resumed for epoch 2
    a1 = D.make_args(tmp, epochs=2, valid_select=0, checkpoint_path=os.path.join(tmp, "ck1.pt"))
    m_a = OracleModel(p["cfg"], po.init_weights(p["cfg"], seed=7))
    tl, _ = D.build_loaders(p["ref"], a1, p["tok"])
    ra = B200Runner(m_a, p["tok"], tl, None, "cpu", a1)
    a1.epochs = 1
    ra.train()            # total_steps was planned for 2 epochs; stop after the first
    a1.epochs = 2
    a1.resume = a1.checkpoint_path
    m_b = OracleModel(p["cfg"], po.init_weights(p["cfg"], seed=99))      # different init: everything must come from the checkpoint
    tl, _ = D.build_loaders(p["ref"], a1, p["tok"])
    rb = B200Runner(m_b, p["tok"], tl, None, "cpu", a1)
    assert rb.start_epoch == 1 and rb.global_step == len(tl)
    rb.train()
    for k in m_full.w:
        assert torch.allclose(m_b.w[k], m_full.w[k], rtol=1e-5, atol=1e-7), k


def test_collaborative_indexing_restatement_equals_reference(pipeline):
    """oracle.generate_collaborative_id_ref == the reference's utils/indexing.generate_collaborative_id (same process, so
    the set iteration order that defines item ids is the same), and the co-occurrence matrix the reference hands to
    SpectralClustering == oracle.cooccurrence_matrix_ref — the matrices csrc/indexing.cu builds on the GPU are checked
    against that restatement in tests/test_model_gpu.py."""
    import random
    from oracle import p5_oracle as po
    idx = pipeline["ref"].indexing
    rng = random.Random(5)
    seqs = {str(u): [str(1000 + i) for i in rng.sample(range(120), rng.randrange(6, 14))] for u in range(1, 80)}
    want = idx.generate_collaborative_id(seqs, 20, 4, "sequential", 0)
    got = po.generate_collaborative_id_ref(seqs, 20, 4, "sequential", 0)
    assert got == want
    recorded = []

    class Recorder:
        def __init__(self, **kw):
            self.kw = kw

        def fit(self, m):
            recorded.append(np.array(m))
            self.labels_ = np.arange(m.shape[0]) % self.kw["n_clusters"]
            return self
    orig = idx.SpectralClustering
    idx.SpectralClustering = Recorder
    try:
        idx.generate_collaborative_id(seqs, 20, 4, "sequential", 1)
    finally:
        idx.SpectralClustering = orig
    _, _, item2id, _ = po.collab_item_ids(seqs)
    adj = po.cooccurrence_matrix_ref(seqs, item2id, 1)
    assert recorded[0].dtype == np.float32 and np.array_equal(recorded[0], adj)
    # the first BFS sub-matrix the reference built == the restated extraction for the same group
    labels = (np.arange(adj.shape[0]) % 4).tolist()
    group0 = [i for i in range(len(labels)) if labels[i] == 0]
    assert np.array_equal(recorded[1], po.submatrix_ref(adj, group0))
