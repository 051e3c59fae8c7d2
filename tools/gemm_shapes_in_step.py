"""Per-shape table of every tcgen05 GEMM launch of ONE T5-base train step (CUDA events around each launch).
usage (under gpurun): python tools/gemm_shapes_in_step.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from openp5_b200 import _lib  # noqa: E402
from openp5_b200.model import P5B200  # noqa: E402
from openp5_b200.synth import synth_items, synth_batch, random_init_  # noqa: E402

B, Le, Ld = 64, 256, 8
m = P5B200("t5-base", vocab_size=32100, precision="bf16", dropout=0.1, max_batch=B, max_enc_len=Le, max_dec_len=Ld)
random_init_(m, seed=2023)
items = synth_items(3416, seed=2023)
hb = synth_batch(B, Le, Ld, 32100, items, seed=1)
lens = hb[1].sum(1).tolist()
b = [t.cuda() for t in hb]
for s in range(3):
    m.train_step(b[0], b[2], b[1], b[3], b[4], lr=1e-3, clip=1.0, enc_lengths=lens)
torch.cuda.synchronize()
lib = _lib.load()
_lib.check(lib.p5_prof_enable(1))
for s in range(2):
    m.train_step(b[0], b[2], b[1], b[3], b[4], lr=1e-3, clip=1.0, enc_lengths=lens)
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
_lib.check(lib.p5_prof_shapes(buf, 1 << 16))
print("two train steps, every tcgen05 GEMM launch grouped by shape:")
print(buf.value.decode())
