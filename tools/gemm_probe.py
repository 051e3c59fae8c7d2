"""Where does the 128x256 tcgen05 GEMM lose its time?  Runs the step's big shapes under the kernel's debug switches
(each in its own process: the switches are read once):
  base            : as shipped
  P5_GEMM_DBG=1   : epilogue without global stores
  P5_GEMM_DBG=3   : ... and without the shared-memory transpose
  P5_GEMM_DBG=7   : ... and without the TMEM loads (main loop only)
  P5_GEMM_PAIR=1  : cta_group::2 (CTA pair shares B)
  P5_GEMM_EPI=direct : TMEM -> registers -> global, no staging
usage (under gpurun): python tools/gemm_probe.py > gpurun_out/gemm_probe.log"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(12800, 2304, 768, 0, 0), (12800, 3072, 768, 0, 0), (12800, 768, 3072, 0, 0), (12800, 768, 768, 0, 0), (8192, 8192, 8192, 0, 0),
          (768, 3072, 12800, 1, 1)]


def child():
    import torch
    from openp5_b200 import _lib
    out = []
    for (M, N, K, am, bm) in SHAPES:
        A = torch.randn((K, M) if am else (M, K), device="cuda").to(torch.bfloat16)
        B = torch.randn((K, N) if bm else (N, K), device="cuda").to(torch.bfloat16)
        ldc = ((N + 63) // 64) * 64
        Cc = torch.empty(M, ldc, device="cuda", dtype=torch.bfloat16)
        f = lambda: _lib.op_gemm(A, B, Cc, a_major=am, b_major=bm, M=M, N=N, K=K, backend=1, force_block_n=256, lda=A.shape[1],
                                 ldb=B.shape[1], ldc=ldc)
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
        out.append(round(2.0 * M * N * K / ms / 1e9, 1))
    print("PROBE " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child()
        sys.exit(0)
    print("shapes (M,N,K,a_major,b_major):", SHAPES)
    for name, env in [("base", {}), ("dbg1_nostore", {"P5_GEMM_DBG": "1"}), ("dbg3_nostage", {"P5_GEMM_DBG": "3"}),
                      ("dbg7_mainloop", {"P5_GEMM_DBG": "7"}), ("pair", {"P5_GEMM_PAIR": "1"}), ("epi_direct", {"P5_GEMM_EPI": "direct"}),
                      ("pair_dbg7", {"P5_GEMM_PAIR": "1", "P5_GEMM_DBG": "7"})]:
        p = subprocess.run([sys.executable, __file__, "--child"], capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
        line = [l for l in p.stdout.splitlines() if l.startswith("PROBE ")]
        print("%-16s TFLOP/s %s" % (name, line[-1][6:] if line else ("FAILED " + p.stderr[-300:])), flush=True)
