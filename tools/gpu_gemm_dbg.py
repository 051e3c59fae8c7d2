"""perf-debug: time the tcgen05 GEMM with parts of the epilogue disabled (P5_GEMM_DBG bitmask, set by the caller)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openp5_b200 import _lib
out = []
for (M, N, K) in [(16384, 2304, 768), (16384, 768, 3072), (16384, 768, 768)]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for fbn in (256, 128):
        f = lambda: _lib.op_gemm(A, B, C, M=M, N=N, K=K, backend=1, force_block_n=fbn, lda=K, ldb=K, ldc=N)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out.append((M, N, K, fbn, round(ms * 1000, 1), round(2.0 * M * N * K / ms / 1e9)))
print("DBG=%s" % os.environ.get("P5_GEMM_DBG", "0"), out)
