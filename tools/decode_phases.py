"""Per-phase time of the persistent decode kernel on the BASELINE eval shape (20 users x 20 beams, T5-base, 3416-item trie).
usage (under gpurun): python tools/decode_phases.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from openp5_b200 import _lib  # noqa: E402
from openp5_b200.model import P5B200  # noqa: E402
from openp5_b200.synth import synth_items, synth_batch, random_init_  # noqa: E402

B, K = int(os.environ.get("P5_EVAL_B", "20")), int(os.environ.get("P5_EVAL_K", "20"))
m = P5B200("t5-base", vocab_size=32100, precision="bf16", dropout=0.0, max_batch=B, max_enc_len=256, max_dec_len=8)
random_init_(m, seed=2023)
items = synth_items(3416, seed=2023)
b = [t.cuda() for t in synth_batch(B, 256, 8, 32100, items, seed=2)]
m.eval()
trie = m.build_trie(items)
g = lambda: m.generate(input_ids=b[0], attention_mask=b[1], whole_word_ids=b[2], max_length=50, trie=trie, num_beams=K, num_return_sequences=K)
for _ in range(3):
    g()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g()
e1.record()
torch.cuda.synchronize()
print("generate() ms per batch: %.3f" % (e0.elapsed_time(e1) / 5))
ms, by, st = C.c_float(), C.c_double(), C.c_int()
_lib.check(_lib.load().p5_decode_last_launch(C.byref(ms), C.byref(by), C.byref(st)))
print("persistent launch: %.3f ms, %d positions, %.1f MB algorithmic -> %.0f GB/s" % (ms.value, st.value, by.value / 1e6, by.value / ms.value / 1e6))
arr = (C.c_uint64 * 32)()
_lib.check(_lib.load().p5_decode_phase_ns(arr))
names = ["qkv", "self_attn", "o", "cq", "cross_attn", "co", "wi", "wo", "lm_head", "user", "list"]
for off, tag in ((0, "light positions"), (16, "heavy positions")):
    tot = sum(arr[off + i] for i in range(11))
    print("%s: %.1f us total" % (tag, tot / 1e3))
    for i, n in enumerate(names):
        print("   %-11s %9.1f us" % (n, arr[off + i] / 1e3))
if os.environ.get("P5_DECODE_PROF_FINE"):
    print("fine (heavy, CTA 0): cross_attn work %.1f us, self_attn work %.1f us, bare barrier %.1f us total over %d positions" % (
        arr[27] / 1e3, arr[29] / 1e3, arr[31] / 1e3, st.value))
    print("cross-attention pair (CTA 0, group 0): issue+K arrive %.1f, S %.1f, softmax %.1f, V wait+PV %.1f, reduce+write %.1f us" % tuple(arr[11 + i] / 1e3 for i in range(5)))
    print("self-attention (CTA 0, warp 0): loads+scores %.1f, merge+store %.1f us" % (arr[28] / 1e3, arr[30] / 1e3))
