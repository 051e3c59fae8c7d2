"""ctypes binding of libp5b200.so (the C-ABI declared in include/p5_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every entry point
needs a CUDA device.  Build with `python -m openp5_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libp5b200.so")


class P5LibraryError(RuntimeError):
    pass


class P5Config(C.Structure):
    _fields_ = [
        ("vocab_size", C.c_int32), ("d_model", C.c_int32), ("d_kv", C.c_int32), ("d_ff", C.c_int32),
        ("num_layers", C.c_int32), ("num_decoder_layers", C.c_int32), ("num_heads", C.c_int32),
        ("rel_buckets", C.c_int32), ("rel_max_distance", C.c_int32), ("ffn_gated_gelu", C.c_int32),
        ("whole_word_rows", C.c_int32), ("dropout", C.c_float), ("ln_eps", C.c_float),
        ("precision", C.c_int32), ("max_batch", C.c_int32), ("max_enc_len", C.c_int32),
        ("max_dec_len", C.c_int32), ("max_beams", C.c_int32), ("use_mn_major", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class P5GemmDesc(C.Structure):
    _fields_ = [
        ("backend", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("nb1", C.c_int32),
        ("nb2", C.c_int32),
        ("A", C.c_void_p), ("a_dtype", C.c_int32), ("a_major", C.c_int32), ("lda", C.c_int64),
        ("a_bs1", C.c_int64), ("a_bs2", C.c_int64),
        ("B", C.c_void_p), ("b_dtype", C.c_int32), ("b_major", C.c_int32), ("ldb", C.c_int64),
        ("b_bs1", C.c_int64), ("b_bs2", C.c_int64),
        ("C", C.c_void_p), ("c_dtype", C.c_int32), ("pad0", C.c_int32), ("ldc", C.c_int64),
        ("c_bs1", C.c_int64), ("c_bs2", C.c_int64),
        ("alpha", C.c_float), ("flags", C.c_int32),
        ("aux", C.c_void_p), ("aux_dtype", C.c_int32), ("pad1", C.c_int32),
        ("resid", C.c_void_p),
        ("seed", C.c_uint64), ("site", C.c_uint32), ("drop_p", C.c_float),
        ("force_block_n", C.c_int32), ("pad2", C.c_int32),
    ]


# every symbol include/p5_b200.h declares (tests check the library exports all of them)
DECLARED_SYMBOLS = [
    "p5_last_error", "p5_version", "p5_create", "p5_destroy", "p5_param_count", "p5_param_info",
    "p5_params_changed", "p5_resize_vocab", "p5_forward", "p5_set_enc_lengths", "p5_backward", "p5_train_fwd_bwd", "p5_grad_norm", "p5_grad_scale",
    "p5_zero_grad", "p5_adamw_step", "p5_adamw_step_zero_grad", "p5_adamw_step_zero_grad_async", "p5_optimizer_join", "p5_eval_metrics", "p5_comm_unique_id", "p5_comm_init", "p5_allreduce_grads",
    "p5_trie_build", "p5_trie_free", "p5_trie_stats", "p5_trie_get", "p5_generate", "p5_op_gemm",
    "p5_launch_count", "p5_prof_enable", "p5_prof_summary", "p5_prof_shapes", "p5_eval_metrics_filtered", "p5_opt_state_info",
    "p5_decode_last_launch", "p5_decode_phase_ns", "p5_cooccurrence", "p5_submatrix", "p5_gemm_tile_width",
]

_lib = None


def load():
    """Load the shared library once; raises P5LibraryError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise P5LibraryError(
            "libp5b200.so is missing (%s): build it with `python -m openp5_b200.build`; "
            "there is no CPU fallback for the B200 engine" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.p5_last_error.restype = C.c_char_p
    i32p, f32p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_void_p

    def sig(name, *argtypes):
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.argtypes = list(argtypes)
            fn.restype = C.c_int
    sig("p5_create", C.POINTER(P5Config), C.c_int, vp, C.POINTER(vp))
    sig("p5_destroy", vp)
    sig("p5_param_count", vp, C.POINTER(C.c_int))
    sig("p5_param_info", vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int64 * 2),
        C.POINTER(vp), C.POINTER(vp))
    sig("p5_params_changed", vp)
    sig("p5_forward", vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_uint64)
    sig("p5_set_enc_lengths", vp, vp, C.c_int)
    sig("p5_backward", vp, vp)
    sig("p5_train_fwd_bwd", vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_uint64)
    sig("p5_grad_norm", vp, vp)
    sig("p5_grad_scale", vp, C.c_float)
    sig("p5_resize_vocab", vp, C.c_int)
    sig("p5_eval_metrics", vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp)
    sig("p5_eval_metrics_filtered", vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, C.c_int,
        C.c_int, vp)
    sig("p5_opt_state_info", vp, C.c_int, C.POINTER(vp), C.POINTER(vp))
    sig("p5_decode_last_launch", C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int))
    sig("p5_decode_phase_ns", vp)
    sig("p5_cooccurrence", vp, vp, C.c_int, C.c_int, C.c_int, vp, vp)
    sig("p5_submatrix", vp, C.c_int, C.c_int, vp, C.c_int, vp, vp)
    sig("p5_zero_grad", vp)
    sig("p5_adamw_step", vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float)
    sig("p5_adamw_step_zero_grad", vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float)
    sig("p5_adamw_step_zero_grad_async", vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float)
    sig("p5_optimizer_join", vp)
    sig("p5_comm_unique_id", vp)
    sig("p5_comm_init", vp, vp, C.c_int, C.c_int)
    sig("p5_allreduce_grads", vp)
    sig("p5_trie_build", vp, vp, vp, C.c_int, C.POINTER(vp))
    sig("p5_trie_free", vp)
    sig("p5_trie_stats", vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int))
    sig("p5_trie_get", vp, vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int))
    sig("p5_generate", vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp,
        C.POINTER(C.c_int))
    sig("p5_op_gemm", C.POINTER(P5GemmDesc), vp)
    sig("p5_prof_enable", C.c_int)
    sig("p5_prof_summary", C.c_char_p, C.c_int)
    sig("p5_prof_shapes", C.c_char_p, C.c_int)
    lib.p5_version.restype = C.c_int
    lib.p5_launch_count.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().p5_last_error()
        raise P5LibraryError("libp5b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_DT = {"torch.float32": 0, "torch.bfloat16": 1}


def op_gemm(A, B, C_out, *, a_major=0, b_major=0, M=None, N=None, K=None, backend=1, alpha=1.0, flags=0, aux=None,
            resid=None, seed=0, site=0, drop_p=0.0, nb1=1, nb2=1, a_bs=(0, 0), b_bs=(0, 0), c_bs=(0, 0), lda=None,
            ldb=None, ldc=None, force_block_n=0):
    """Run one (batched) GEMM through the C-ABI op hook.  Tensors are CUDA torch tensors; the logical problem
    is C[M,N] = A[M,K] * B[N,K]^T with A/B stored K-major ([rows,K]) or MN-major ([K,rows])."""
    lib = load()
    d = P5GemmDesc()
    d.backend = backend
    d.M, d.N, d.K, d.nb1, d.nb2 = M, N, K, nb1, nb2
    d.A, d.a_dtype, d.a_major = A.data_ptr(), _DT[str(A.dtype)], a_major
    d.lda = lda if lda is not None else A.stride(-2)
    d.a_bs1, d.a_bs2 = a_bs
    d.B, d.b_dtype, d.b_major = B.data_ptr(), _DT[str(B.dtype)], b_major
    d.ldb = ldb if ldb is not None else B.stride(-2)
    d.b_bs1, d.b_bs2 = b_bs
    d.C, d.c_dtype = C_out.data_ptr(), _DT[str(C_out.dtype)]
    d.ldc = ldc if ldc is not None else C_out.stride(-2)
    d.c_bs1, d.c_bs2 = c_bs
    d.alpha, d.flags = alpha, flags
    d.aux = aux.data_ptr() if aux is not None else None
    d.aux_dtype = _DT[str(aux.dtype)] if aux is not None else 1
    d.resid = resid.data_ptr() if resid is not None else None
    d.seed, d.site, d.drop_p = seed, site, drop_p
    d.force_block_n = force_block_n
    check(lib.p5_op_gemm(C.byref(d), current_stream_ptr()))
