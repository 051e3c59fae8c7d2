"""Repeat the second half of tests/test_dropin_gpu.py::test_batch_stager_delivers_the_collator_batches: 12 fp32 train steps on
the reference-produced batches, inline loop vs staged loop, several times; prints the final losses (run-to-run spread of the
fp32-atomics order vs the difference between the two loops)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle import p5_oracle as po  # noqa: E402
from openp5_b200.model import P5B200  # noqa: E402
from openp5_b200.runner import B200Runner  # noqa: E402
import test_dropin_gpu as T  # noqa: E402

fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "dropin_ml100k.npz")))
train = T._batches(fx, "train") * 3
cfg = po.t5_cfg("t5-small", vocab_size=32100)
w = po.init_weights(cfg, seed=2023)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
out = {0: [], 1: []}
for rep in range(reps):
    for stage in (0, 1):
        m = P5B200("t5-small", vocab_size=32100, precision="fp32", dropout=0.0, max_batch=4, max_enc_len=64, max_dec_len=16)
        m.load_state_dict(w)
        r = B200Runner(m, None, T._Loader(train[:12]), None, m.device, T._args(valid_select=0, stage_batches=stage))
        r.train()
        out[stage].append(r.last_train_loss)
        del m, r
print("env P5_NO_EARLY_NORM=%s" % os.environ.get("P5_NO_EARLY_NORM"))
for s in (0, 1):
    print("stage %d:" % s, " ".join("%.7f" % x for x in out[s]))
allv = out[0] + out[1]
print("spread (max-min)/mean = %.2e" % ((max(allv) - min(allv)) / (sum(allv) / len(allv))))
