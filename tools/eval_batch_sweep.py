"""Constrained beam-search throughput against the number of users per generate() call (T5-base, 20 beams, 3416-item trie,
Le = 256): the bench line reports the reference scripts' --eval_batch_size 20; the reference's own default is 32 and the
persistent decode kernel takes up to 64 users.  Device-resident inputs, CUDA events, max_length 50."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from openp5_b200.model import P5B200  # noqa: E402
from openp5_b200.synth import synth_items, synth_batch, random_init_  # noqa: E402

Le, Ld, K = 256, 8, 20
m = P5B200("t5-base", vocab_size=32100, precision="bf16", dropout=0.1, max_batch=64, max_enc_len=Le, max_dec_len=Ld)
random_init_(m, seed=2023)
items = synth_items(3416, seed=2023)
m.eval()
trie = m.build_trie(items)
for B in (20, 32, 48, 64):
    try:
        bs = [[t.cuda() for t in synth_batch(B, Le, Ld, 32100, items, seed=2 + i)] for i in range(2)]
        g = lambda b: m.generate(input_ids=b[0], attention_mask=b[1], whole_word_ids=b[2], max_length=50, trie=trie, num_beams=K,
                                 num_return_sequences=K)
        for i in range(3):
            g(bs[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 8
        e0.record()
        for i in range(n):
            o = g(bs[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("users/batch %3d: %7.3f ms/batch  %9.0f items ranked/s  (%d sequences returned)" % (B, ms, B * K / (ms / 1e3), o["sequences"].shape[0]), flush=True)
    except Exception as e:  # noqa
        print("users/batch %3d: failed: %s" % (B, repr(e)[:300]), flush=True)
        break
