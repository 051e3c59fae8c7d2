"""GPU bring-up check for the tcgen05 GEMM: every case is run in a child process (a trap in one case must not
take the others down) and compared with torch fp32 matmul on the bf16-rounded operands.
Run under gpurun:  python tools/gpu_check_gemm.py > gpurun_out/gemm_check.log 2>&1
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # name, M, N, K, a_major, b_major, nb1, nb2, force_bn, flags
    ("kk_small_bn256", 256, 512, 128, 0, 0, 1, 1, 256, 0),
    ("kk_small_bn128", 256, 512, 128, 0, 0, 1, 1, 128, 0),
    ("kk_small_bn64", 256, 512, 128, 0, 0, 1, 1, 64, 0),
    ("kk_k64_single", 128, 256, 64, 0, 0, 1, 1, 256, 0),
    ("kk_ragged", 200, 300, 136, 0, 0, 1, 1, 0, 0),
    ("kk_deepk", 384, 768, 3072, 0, 0, 1, 1, 0, 0),
    ("kk_manytiles", 2048, 2304, 768, 0, 0, 1, 1, 256, 0),
    ("mnA", 256, 512, 128, 1, 0, 1, 1, 256, 0),
    ("mnB", 256, 512, 128, 0, 1, 1, 1, 256, 0),
    ("mnB_bn64", 256, 64, 256, 0, 1, 1, 1, 64, 0),
    ("mnAB", 256, 512, 192, 1, 1, 1, 1, 128, 0),
    ("mnAB_ragged", 200, 328, 136, 1, 1, 1, 1, 0, 0),
    ("batched_kk", 256, 256, 64, 0, 0, 3, 2, 0, 0),
    ("batched_mnB", 256, 64, 256, 0, 1, 3, 2, 0, 0),
    ("epi_relu", 256, 512, 128, 0, 0, 1, 1, 0, 1),
    ("epi_mulpos", 256, 512, 128, 0, 0, 1, 1, 0, 2),
    ("epi_dropout", 256, 512, 128, 0, 0, 1, 1, 0, 4),
    ("epi_resid_f32", 256, 512, 128, 0, 0, 1, 1, 0, 8),
    ("epi_accum_f32", 256, 512, 128, 0, 0, 1, 1, 0, 16),
    ("vocab_head", 512, 32100, 768, 0, 0, 1, 1, 0, 0),
    # split-K weight gradients: EPI_ATOMIC (red.global.add.v2.f32 into ONE fp32 C from every split), both operands
    # MN-major as linear_wgrad issues them — gemm_tc_kernel<256, EPI_ATOMIC|F32> is the most-launched kernel of a train step
    ("epi_atomic_splitk_bn256", 2304, 768, 512, 1, 1, 8, 1, 256, 32),
    ("epi_atomic_splitk_bn128", 768, 768, 256, 1, 1, 4, 1, 128, 32),
    ("epi_atomic_splitk_bn64", 768, 64, 256, 1, 1, 4, 1, 64, 32),
    ("epi_atomic_ragged", 200, 328, 136, 1, 1, 3, 2, 0, 32),
    ("kk_bn192", 512, 768, 256, 0, 0, 1, 1, 192, 0),
    ("kk_bn192_ragged", 200, 700, 136, 0, 0, 1, 1, 192, 0),
    ("mnAB_bn192", 256, 576, 192, 1, 1, 1, 1, 192, 0),
    ("epi_resid_bn192", 384, 768, 128, 0, 0, 1, 1, 192, 8),
    ("epi_dropout_bn192", 384, 768, 128, 0, 0, 1, 1, 192, 4),
    ("auto_bn192_n768", 12800, 768, 768, 0, 0, 1, 1, 0, 0),
    # big enough for the CTA-pair (cta_group::2) path when P5_GEMM_PAIR=1 (>= 74 units of 256 x 256)
    ("pair_kk", 4096, 2304, 768, 0, 0, 1, 1, 256, 0),
    ("pair_mnAB", 2304, 3072, 4096, 1, 1, 1, 1, 256, 0),
    ("pair_mnB", 4096, 2304, 768, 0, 1, 1, 1, 256, 0),
    ("pair_ragged", 4000, 2500, 520, 0, 0, 1, 1, 256, 0),
    ("pair_odd_blocks", 3968, 2304, 256, 0, 0, 1, 1, 256, 0),
    ("pair_batched", 2048, 1536, 256, 0, 0, 2, 1, 256, 0),
    ("pair_epi_resid", 4096, 2304, 256, 0, 0, 1, 1, 256, 8),
    ("pair_epi_mulpos", 4096, 2304, 256, 0, 0, 1, 1, 256, 2),
    ("pair_epi_dropout", 4096, 2304, 256, 0, 0, 1, 1, 256, 4),
    ("pair_epi_accum", 4096, 2304, 256, 0, 0, 1, 1, 256, 16),
]


def run_case(name):
    import torch
    from openp5_b200 import _lib
    spec = [c for c in CASES if c[0] == name][0]
    _, M, N, K, am, bm, nb1, nb2, fbn, flags = spec
    torch.manual_seed(0)
    dev = "cuda"
    nb = nb1 * nb2
    A = torch.randn(nb, M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(nb, N, K, device=dev).to(torch.bfloat16)
    ref = torch.matmul(A.float(), B.float().transpose(1, 2))  # [nb, M, N]
    A_st = A.transpose(1, 2).contiguous() if am else A.contiguous()
    B_st = B.transpose(1, 2).contiguous() if bm else B.contiguous()
    c_f32 = bool(flags & (8 | 16 | 32)) or name == "vocab_head"
    ldc = ((N + 63) // 64) * 64
    Cout = torch.zeros(nb, M, ldc, device=dev, dtype=torch.float32 if c_f32 else torch.bfloat16)
    aux = resid = None
    kw = {}
    if flags & 1:
        ref = ref.clamp_min(0)
    if flags & 2:
        aux_full = torch.randn(nb, M, ldc, device=dev).to(torch.bfloat16)
        aux = aux_full
        ref = torch.where(aux_full[:, :, :N].float() > 0, ref, torch.zeros_like(ref))
    if flags & 8:
        resid = torch.randn(nb, M, ldc, device=dev, dtype=torch.float32)
        ref = ref + resid[:, :, :N]
    if flags & 16:
        Cout.normal_()
        ref = ref + Cout[:, :, :N]
    if flags & 4:
        kw = dict(seed=1234, site=7, drop_p=0.25)
    if flags & 32:
        # every (b1, b2) batch accumulates into the SAME C (batch strides 0) on top of what is already there
        Cout = torch.randn(1, M, ldc, device=dev, dtype=torch.float32)
        ref = (Cout[:, :, :N] + ref.sum(dim=0, keepdim=True)).clone()
    a_rows = A_st.shape[1]
    b_rows = B_st.shape[1]
    common = dict(a_major=am, b_major=bm, M=M, N=N, K=K, nb1=nb1, nb2=nb2,
                  a_bs=(a_rows * A_st.shape[2], a_rows * A_st.shape[2] * nb1),
                  b_bs=(b_rows * B_st.shape[2], b_rows * B_st.shape[2] * nb1),
                  c_bs=(0, 0) if flags & 32 else (M * ldc, M * ldc * nb1), lda=A_st.shape[2], ldb=B_st.shape[2], ldc=ldc,
                  flags=flags, aux=aux, resid=resid, **kw)
    if flags & 4:
        # dropout: the SIMT backend uses the same counter RNG -> masks must agree exactly
        Cs = Cout.clone()
        _lib.op_gemm(A_st, B_st, Cs, backend=0, **common)
        _lib.op_gemm(A_st, B_st, Cout, backend=1, force_block_n=fbn, **common)
        torch.cuda.synchronize()
        got, want = Cout[:, :, :N].float(), Cs[:, :, :N].float()
        keep = (got != 0).float().mean().item()
        mask_equal = bool(((got != 0) == (want != 0)).all())
        err = (got - want).abs().max().item()
        ok = mask_equal and err <= 0.05 * want.abs().max().item() + 1e-3 and abs(keep - 0.75) < 0.02
        return dict(name=name, ok=bool(ok), max_err=err, keep=keep, mask_equal=mask_equal)
    _lib.op_gemm(A_st, B_st, Cout, backend=1, force_block_n=fbn, **common)
    torch.cuda.synchronize()
    got = Cout[:, :, :N].float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-2 if not c_f32 else 2e-3) * scale + 1e-3
    pad_clean = bool((Cout[:, :, N:] == 0).all()) if not (flags & (16 | 32)) else True
    res = dict(name=name, ok=bool(err <= tol and pad_clean), max_err=err, ref_scale=scale, pad_clean=pad_clean)
    if not res["ok"]:
        bad = ((got - ref).abs() > tol).nonzero()
        res["n_bad"] = int(bad.shape[0])
        res["first_bad"] = bad[:6].tolist()
        res["bad_rows_mod128"] = sorted(set((bad[:, 1] % 128).tolist()))[:16]
        res["bad_cols_mod64"] = sorted(set((bad[:, 2] % 64).tolist()))[:16]
    return res


def bench():
    import torch
    from openp5_b200 import _lib
    out = []
    for (M, N, K, am, bm) in [(16384, 2304, 768, 0, 0), (16384, 3072, 768, 0, 0), (16384, 768, 3072, 0, 0),
                              (16384, 768, 768, 0, 0), (8192, 8192, 8192, 0, 0), (768, 3072, 16384, 1, 1),
                              (16384, 768, 3072, 0, 1), (512, 768, 768, 0, 0), (512, 32100, 768, 0, 0)]:
        A = torch.randn((K, M) if am else (M, K), device="cuda").to(torch.bfloat16)
        B = torch.randn((K, N) if bm else (N, K), device="cuda").to(torch.bfloat16)
        ldc = ((N + 63) // 64) * 64
        Cc = torch.empty(M, ldc, device="cuda", dtype=torch.bfloat16)
        for fbn in (0, 256, 128):
            f = lambda: _lib.op_gemm(A, B, Cc, a_major=am, b_major=bm, M=M, N=N, K=K, backend=1, force_block_n=fbn,
                                     lda=A.shape[1], ldb=B.shape[1], ldc=ldc)
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                f()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            out.append(dict(M=M, N=N, K=K, am=am, bm=bm, bn=fbn, ms=ms, tflops=2.0 * M * N * K / ms / 1e9))
        # torch (cuBLAS) for scale
        At = A.t().contiguous() if am else A
        Bt = B.t().contiguous() if bm else B
        for _ in range(3):
            torch.matmul(At, Bt.t())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.matmul(At, Bt.t())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out.append(dict(M=M, N=N, K=K, am=am, bm=bm, bn="cublas", ms=ms, tflops=2.0 * M * N * K / ms / 1e9))
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        try:
            print("RESULT " + json.dumps(run_case(sys.argv[2])))
        except Exception as e:  # noqa
            print("RESULT " + json.dumps(dict(name=sys.argv[2], ok=False, error=repr(e)[:400])))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--bench":
        for r in bench():
            print("BENCH " + json.dumps(r))
        sys.exit(0)
    results = []
    for c in CASES:
        try:
            p = subprocess.run([sys.executable, __file__, "--case", c[0]], capture_output=True, text=True, timeout=120)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                results.append(json.loads(line[-1][7:]))
            else:
                results.append(dict(name=c[0], ok=False, rc=p.returncode, stdout=p.stdout[-600:], stderr=p.stderr[-600:]))
        except subprocess.TimeoutExpired:
            results.append(dict(name=c[0], ok=False, error="timeout"))
        print(json.dumps(results[-1]), flush=True)
    print("SUMMARY passed %d / %d" % (sum(1 for r in results if r.get("ok")), len(results)))
    if all(r.get("ok") for r in results if r["name"].startswith("kk_")):
        p = subprocess.run([sys.executable, __file__, "--bench"], capture_output=True, text=True, timeout=300)
        print(p.stdout[-6000:])
        print(p.stderr[-2000:])
