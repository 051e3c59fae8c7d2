"""One T5-base train step (BASELINE configs[1]) between cudaProfilerStart/Stop, for ncu --profile-from-start off:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
Optional argv[1] = "eval" profiles one constrained beam-search batch instead."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from openp5_b200.model import P5B200  # noqa: E402
from openp5_b200.synth import synth_items, synth_batch, random_init_  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B, Le, Ld = 64, 256, 8
m = P5B200("t5-base", vocab_size=32100, precision="bf16", dropout=0.1, max_batch=B, max_enc_len=Le, max_dec_len=Ld)
random_init_(m, seed=2023)
items = synth_items(3416, seed=2023)
if mode == "train":
    hb = synth_batch(B, Le, Ld, 32100, items, seed=1)
    lens = hb[1].sum(1).tolist() if os.environ.get("P5_PADDED") != "1" else None   # packed (padding-free) by default
    b = [t.cuda() for t in hb]
    for s in range(3):
        m.train_step(b[0], b[2], b[1], b[3], b[4], lr=1e-3, clip=1.0, enc_lengths=lens)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m.train_step(b[0], b[2], b[1], b[3], b[4], lr=1e-3, clip=1.0, enc_lengths=lens)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    b = [t.cuda() for t in synth_batch(20, Le, Ld, 32100, items, seed=2)]
    m.eval()
    trie = m.build_trie(items)
    g = lambda: m.generate(input_ids=b[0], attention_mask=b[1], whole_word_ids=b[2], max_length=50, trie=trie, num_beams=20,
                           num_return_sequences=20)
    g()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    g()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled", mode)
