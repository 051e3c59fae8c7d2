// engine.cu — parameter layout, workspace, and the forward / backward pass of the P5 T5 model on one B200.
//
// Reference path restated here (what the kernels are wired to compute):
//   encoder  : ref src/src_t5/model/P5_T5.py:74-204  (JointEncoder: token + whole-word embedding, shared position
//              bias + padding mask built once, T5 blocks, final RMSNorm + dropout)
//   decoder  : HF:models/t5/modeling_t5.py:637-793 via P5_T5.py:338-350 (causal self-attention, NO decoder pad mask,
//              cross-attention with zero position bias + encoder pad mask)
//   head/loss: P5_T5.py:352-369 (hidden * d_model^-0.5, tied lm_head, un-reduced CE)
// Data layout in HBM: token-major row-major activations [B*L, features]; the residual stream is fp32, GEMM
// operands are `dt` (bf16 on the tensor-core path, fp32 on the parity path); heads are column slices of width 64.
#include "engine.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace p5 {

void comm_allreduce_range(Engine* e, int64_t off, int64_t n);   // comm.cu (no-op without a communicator)
cudaStream_t comm_stream(Engine* e);                              // comm.cu: the all-reduce side stream (null without a communicator)

static inline void* poff(const void* p, int64_t elems, int dt) {
    return (void*)((const char*)p + elems * (int64_t)dtype_size(dt));
}

// ------------------------------------------------------------------------------------------------------------
// construction
// ------------------------------------------------------------------------------------------------------------
void* Engine::dalloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 256;
    P5_CUDA(cudaMalloc(&p, bytes));
    P5_CUDA(cudaMemsetAsync(p, 0, bytes, st));
    allocs.push_back(p);
    return p;
}

Engine::Engine(const P5Config& c, int dev, cudaStream_t stream) : cfg(c), device(dev), st(stream) {
    P5_CHECK(c.d_kv == 64, "d_kv must be 64");
    P5_CHECK(c.d_model % 32 == 0 && c.d_model <= 1024, "d_model must be a multiple of 32 and <= 1024");
    P5_CHECK(c.d_ff % 8 == 0, "d_ff must be a multiple of 8");
    P5_CHECK(c.max_enc_len >= 1 && c.max_enc_len <= 512, "max_enc_len must be in [1, 512] (Collator.py:13)");
    P5_CHECK(c.max_batch >= 1 && c.max_dec_len >= 1, "max_batch / max_dec_len");
    P5_CUDA(cudaSetDevice(dev));
    P5_CHECK(c.precision >= 0 && c.precision <= 2, "precision must be 0 (fp32 SIMT), 1 (bf16 tcgen05) or 2 (bf16x3 tcgen05 parity)");
    dt = c.precision == 1 ? DT_BF16 : DT_F32;
    x3 = c.precision == 2;
    mn = c.use_mn_major != 0;
    d = c.d_model; H = c.num_heads; A = H * 64; ff = c.d_ff; V = c.vocab_size; Vpad = (int)round_up(V, 64);
    NE = c.num_layers; ND = c.num_decoder_layers; gated = c.ffn_gated_gelu != 0; p_drop = c.dropout;

    // ---- parameter table (HF state_dict names); q,k,v / cross k,v / wi_0,wi_1 are adjacent so
    //      that each group is ONE GEMM weight
    int64_t off = 0;
    auto add = [&](const std::string& name, int64_t r, int64_t cdim) {
        ParamInfo pi;
        pi.name = name; pi.ndim = cdim > 0 ? 2 : 1; pi.shape[0] = r; pi.shape[1] = cdim > 0 ? cdim : 0;
        pi.numel = cdim > 0 ? r * cdim : r;
        pi.off = off;
        off += round_up(pi.numel, 64);
        params.push_back(pi);
        return pi.off;
    };
    auto add_attn = [&](const std::string& pre, AttnOff& a) {
        a.q = add(pre + ".q.weight", A, d);
        a.k = add(pre + ".k.weight", A, d);
        a.v = add(pre + ".v.weight", A, d);
        a.o = add(pre + ".o.weight", d, A);
    };
    auto add_ffn = [&](const std::string& pre, FfnOff& f) {
        if (gated) {
            f.wi = add(pre + ".DenseReluDense.wi_0.weight", ff, d);
            f.wi1 = add(pre + ".DenseReluDense.wi_1.weight", ff, d);
        } else {
            f.wi = add(pre + ".DenseReluDense.wi.weight", ff, d);
            f.wi1 = -1;
        }
        f.wo = add(pre + ".DenseReluDense.wo.weight", d, ff);
    };
    off_shared = add("shared.weight", V, d);
    off_ww = add("encoder.whole_word_embeddings.weight", c.whole_word_rows, d);
    enc.resize(NE);
    for (int i = 0; i < NE; ++i) {
        const std::string b = "encoder.block." + std::to_string(i) + ".layer";
        add_attn(b + ".0.SelfAttention", enc[i].sa);
        if (i == 0) off_enc_rel = add(b + ".0.SelfAttention.relative_attention_bias.weight", c.rel_buckets, H);
        enc[i].ln0 = add(b + ".0.layer_norm.weight", d, 0);
        add_ffn(b + ".1", enc[i].ff);
        enc[i].ln1 = add(b + ".1.layer_norm.weight", d, 0);
    }
    off_enc_final = add("encoder.final_layer_norm.weight", d, 0);
    dec.resize(ND);
    for (int i = 0; i < ND; ++i) {
        const std::string b = "decoder.block." + std::to_string(i) + ".layer";
        add_attn(b + ".0.SelfAttention", dec[i].sa);
        if (i == 0) off_dec_rel = add(b + ".0.SelfAttention.relative_attention_bias.weight", c.rel_buckets, H);
        dec[i].ln0 = add(b + ".0.layer_norm.weight", d, 0);
        add_attn(b + ".1.EncDecAttention", dec[i].ca);
        dec[i].ln1 = add(b + ".1.layer_norm.weight", d, 0);
        add_ffn(b + ".2", dec[i].ff);
        dec[i].ln2 = add(b + ".2.layer_norm.weight", d, 0);
    }
    off_dec_final = add("decoder.final_layer_norm.weight", d, 0);
    n_flat = off;
    P = dalloc_t<float>(n_flat);
    G = dalloc_t<float>(n_flat);
    M1 = dalloc_t<float>(n_flat);
    V2 = dalloc_t<float>(n_flat);
    if (dt == DT_BF16) P16 = dalloc_t<bf16>(n_flat);

    // ---- workspace
    Bm = c.max_batch; Lem = (int)round_up(c.max_enc_len, 8); Ldm = c.max_dec_len;
    // +512 rows: the packed layout rounds its row count up to a multiple of 512 with inert filler rows
    const int64_t Mem = (int64_t)Bm * Lem + 512, Mdm = (int64_t)Bm * Ldm, Mx = Mem > Mdm ? Mem : Mdm;
    const size_t e = esz();
    ids_e = dalloc_t<int>(Mem); mask_e = dalloc_t<int>(Mem); ww_e = dalloc_t<int>(Mem);
    labels = dalloc_t<int>(Mdm); dec_ids = dalloc_t<int>(Mdm); lmask = dalloc_t<int>(Mdm);
    const bool tc_attn = dt == DT_BF16;
    xe.resize(2 * NE + 1); rstd_e.resize(2 * NE + 1); ne.resize(2 * NE);
    for (auto& p : xe) p = dalloc_t<float>(Mem * d);
    for (auto& p : rstd_e) p = dalloc_t<float>(Mem);
    for (auto& p : ne) p = dalloc(Mem * d * e);
    enc_out = dalloc(Mem * d * e);
    qkv_e.resize(NE); ctx_e.resize(NE); h_e.resize(NE); z_e.assign(NE, nullptr); P_e.assign(NE, nullptr);
    lse_e.assign(NE, nullptr); p_unnorm.assign(NE, false); p_fbwd.assign(NE, false);
    for (int i = 0; i < NE; ++i) {
        qkv_e[i] = dalloc(Mem * 3 * A * e);
        ctx_e[i] = dalloc(Mem * A * e);
        h_e[i] = dalloc(Mem * ff * e);
        if (gated) z_e[i] = dalloc(Mem * 2 * ff * e);
        lse_e[i] = dalloc_t<float>((int64_t)Bm * H * Lem);   // fp32 mode: LSE; bf16 mode: 1 / row sum of the fused kernel
    }
    // the L^2-sized buffers of the materialised attention paths (P per layer, fp32 scores, Pd, dS) are allocated on first
    // use (ensure_attn_scratch): with the fused forward + backward (Le <= 512) they are never touched, and at T5-large /
    // Le = 512 / B = 64 they would take 15 GB
    yd.resize(3 * ND + 1); rstd_d.resize(3 * ND + 1); nd.resize(3 * ND);
    for (auto& p : yd) p = dalloc_t<float>(Mdm * d);
    for (auto& p : rstd_d) p = dalloc_t<float>(Mdm);
    // the saved inputs of the decoder's linear layers are slabs with a constant per-layer stride (batched weight gradients)
    nd_all = dalloc((size_t)3 * (ND > 0 ? ND : 1) * Mdm * d * e);
    for (int i = 0; i < 3 * ND; ++i) nd[i] = poff(nd_all, (int64_t)i * Mdm * d, dt);
    dec_out = dalloc(Mdm * d * e);
    sqkv.resize(ND); sctx.resize(ND); cq.resize(ND); ckv.resize(ND); cctx.resize(ND); h_d.resize(ND);
    z_d.assign(ND, nullptr); slse.resize(ND); clse.resize(ND);
    ckv_ld = (int64_t)ND * 2 * A;
    if (ND > 0) ckv_all = dalloc(Mem * ckv_ld * e);
    dec_layer_stride = ND > 1 ? dec[1].ca.k - dec[0].ca.k : 0;
    {
        static const bool off = getenv("P5_NO_BATCHED_CKV") != nullptr;
        batched_ckv = dt == DT_BF16 && ND > 1 && !off;
        for (int i = 1; i < ND; ++i) batched_ckv = batched_ckv && (dec[i].ca.k - dec[0].ca.k == i * dec_layer_stride);
        if (batched_ckv) g_ckv_all = dalloc(Mem * ckv_ld * e);
    }
    h_d_all = dalloc((size_t)(ND > 0 ? ND : 1) * Mdm * ff * e);
    sctx_all = dalloc((size_t)(ND > 0 ? ND : 1) * Mdm * A * e);
    cctx_all = dalloc((size_t)(ND > 0 ? ND : 1) * Mdm * A * e);
    {
        static const bool off = getenv("P5_NO_BATCHED_DWG") != nullptr;
        batched_dwg = dt == DT_BF16 && ND > 2 && !gated && !off;
        for (int i = 2; i < ND; ++i)      // layer 0 carries the relative-bias table between sa.o and ln0: strides are checked from layer 1 on
            batched_dwg = batched_dwg && (dec[i].sa.q - dec[1].sa.q == (i - 1) * dec_layer_stride) && (dec[i].ff.wo - dec[0].ff.wo == i * dec_layer_stride);
        if (batched_dwg) {
            gdw_all = dalloc((size_t)ND * Mdm * d * e); gdc_all = dalloc((size_t)ND * Mdm * d * e); gds_all = dalloc((size_t)ND * Mdm * d * e);
            gff_all = dalloc((size_t)ND * Mdm * ff * e); gcq_all = dalloc((size_t)ND * Mdm * A * e); gsq_all = dalloc((size_t)ND * Mdm * 3 * A * e);
        }
    }
    for (int i = 0; i < ND; ++i) {
        sqkv[i] = dalloc(Mdm * 3 * A * e);
        sctx[i] = poff(sctx_all, (int64_t)i * Mdm * A, dt);
        cq[i] = dalloc(Mdm * A * e);
        ckv[i] = poff(ckv_all, (int64_t)i * 2 * A, dt);
        cctx[i] = poff(cctx_all, (int64_t)i * Mdm * A, dt);
        h_d[i] = poff(h_d_all, (int64_t)i * Mdm * ff, dt);
        if (gated) z_d[i] = dalloc(Mdm * 2 * ff * e);
        slse[i] = dalloc_t<float>((int64_t)Bm * H * Ldm);
        clse[i] = dalloc_t<float>((int64_t)Bm * H * Ldm);
    }
    logits = dalloc_t<float>(Mdm * Vpad);
    lse_ce = dalloc_t<float>(Mdm); loss_tok = dalloc_t<float>(Mdm); dloss = dalloc_t<float>(Mdm);
    loss_scalar = dalloc_t<float>(4);
    dlogits = dalloc(Mdm * Vpad * e);
    dx_a = dalloc_t<float>(Mx * d);
    d_encout = dalloc_t<float>(Mem * d);
    g_ff = dalloc(Mx * (gated ? 3 : 1) * ff * e);
    g_d = dalloc(Mx * d * e); g_d2 = dalloc(Mx * d * e);
    g_qkv = dalloc(Mx * 3 * A * e); g_ctx = dalloc(Mx * A * e); g_ckv = dalloc(Mem * 2 * A * e);
    f_qkv = dalloc_t<float>((tc_attn ? Mdm : Mx) * 3 * A);
    f_ckv = dalloc_t<float>(Mem * 2 * A);
    bias_enc = dalloc_t<float>((int64_t)H * (2 * Lem)); dbias_enc = dalloc_t<float>((int64_t)H * (2 * Lem));
    const int Ldb = Ldm > 256 ? Ldm : 256;   // generate() builds the decoder table for max_length <= 256 positions
    bias_dec = dalloc_t<float>((int64_t)H * (2 * Ldb)); dbias_dec = dalloc_t<float>((int64_t)H * (2 * Ldb));
    lut_enc = dalloc_t<int>(2 * Lem); lut_dec = dalloc_t<int>(2 * Ldb);
    norm_partial = dalloc_t<float>(1024); norm_out = dalloc_t<float>(4); norm_early = dalloc_t<float>(EARLY_SLOTS);
    lens_d = dalloc_t<int>(Bm + 1); offs_d = dalloc_t<int>(Bm + 1);
    ids_p = dalloc_t<int>(Mem); ww_p = dalloc_t<int>(Mem);
    qkv_pad = dalloc(Mem * 3 * A * e); ctx_pad = dalloc(Mem * A * e); dqkv_pad = dalloc(Mem * 3 * A * e);
    if (!tc_attn) f_qkv_pad = dalloc_t<float>(Mem * 3 * A);
    P5_CUDA(cudaStreamSynchronize(st));
}

void free_gen_ws(struct GenWs* g);
void free_persist_ws();

Engine::~Engine() {
    cudaSetDevice(device);
    if (st_opt) cudaStreamSynchronize(st_opt);
    cudaStreamSynchronize(st);
    for (auto e : ev_opt) cudaEventDestroy(e);
    if (ev_opt_start) cudaEventDestroy(ev_opt_start);
    if (st_opt) cudaStreamDestroy(st_opt);
    if (st_norm) { cudaStreamSynchronize(st_norm); cudaStreamDestroy(st_norm); }
    if (ev_norm_fork) cudaEventDestroy(ev_norm_fork);
    if (ev_norm_join) cudaEventDestroy(ev_norm_join);
    if (gen) free_gen_ws(gen);
    free_persist_ws();
    gemm_x3_release();
    for (void* p : allocs) cudaFree(p);
    gemm_tc_clear_cache();
}

__global__ void init_normal_rows_kernel(float* __restrict__ w, int64_t n, uint64_t seed) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // Box-Muller on two counter hashes: N(0, 1) (HF T5 `_init_weights`: shared.weight ~ N(0, factor * 1.0))
        const uint32_t a = drop_hash(seed, 1u, (uint64_t)(2 * i)), b = drop_hash(seed, 2u, (uint64_t)(2 * i + 1));
        const float u1 = ((float)a + 1.f) * (1.f / 4294967296.f), u2 = (float)b * (1.f / 4294967296.f);
        w[i] = sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2);
    }
}

// model.resize_token_embeddings(n) (ref main.py:193): a NEW engine for the new vocabulary that keeps every tensor other
// than shared.weight, the first min(V, n) embedding rows (with their Adam moments), the data-parallel communicator and
// the mode flags; new rows ~ N(0, 1) as HF's T5 initialiser draws them.  The C-ABI handle is a box around the Engine
// pointer (capi.cu), so the caller swaps the pointer and deletes the old engine — no object is relocated bytewise.
Engine* Engine::resized(int new_vocab) {
    P5_CHECK(new_vocab >= 2 && new_vocab < (1 << 24), "resize_vocab: vocabulary size out of range");
    P5_CUDA(cudaSetDevice(device));
    join_optimizer();
    P5_CUDA(cudaStreamSynchronize(st));
    P5Config c2 = cfg;
    c2.vocab_size = new_vocab;
    Engine* ne = new Engine(c2, device, st);
    try {
        P5_CHECK(ne->params.size() == params.size(), "resize_vocab: parameter lists differ");
        for (size_t i = 0; i < params.size(); ++i) {
            const ParamInfo &o = params[i], &n = ne->params[i];
            P5_CHECK(o.name == n.name, "resize_vocab: parameter order differs");
            const int64_t cnt = o.numel < n.numel ? o.numel : n.numel;    // row-major: common leading rows
            P5_CUDA(cudaMemcpyAsync(ne->P + n.off, P + o.off, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
            P5_CUDA(cudaMemcpyAsync(ne->M1 + n.off, M1 + o.off, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
            P5_CUDA(cudaMemcpyAsync(ne->V2 + n.off, V2 + o.off, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
            if (n.numel > cnt) {
                launch_k(init_normal_rows_kernel, 256, 256, 0, st, ne->P + n.off + cnt, n.numel - cnt, (uint64_t)0x5eed0000 + i);
                P5_CUDA(cudaGetLastError());
            }
        }
        ne->training = training;
        ne->shadow_stale = true;
        P5_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        delete ne;
        throw;
    }
    // the communicator (own stream / events, addresses the gradient buffer through the engine it is called with) moves over
    ne->nccl_comm = nccl_comm; ne->world = world; ne->rank = rank;
    nccl_comm = nullptr; world = 1; rank = 0;
    return ne;
}

DropCfg Engine::drop(uint32_t kind, int layer) const {
    DropCfg c;
    if (training && p_drop > 0.f) {
        c.seed = seed; c.site = kind * 64u + (uint32_t)layer; c.thr = drop_threshold(p_drop);
        c.inv_keep = 1.f / (1.f - p_drop);
    }
    return c;
}

void Engine::refresh_shadow() {
    join_optimizer();
    if (dt == DT_BF16) cast_f32_to(P, P16, DT_BF16, n_flat, st);
    shadow_stale = false;
}

// ------------------------------------------------------------------------------------------------------------
// GEMM helpers
// ------------------------------------------------------------------------------------------------------------
void Engine::gemm(GemmProblem& p) {
    if (next_gemm_indep) { p.indep_of_prev = true; next_gemm_indep = false; }   // one-shot flag set by the caller
    if (dt == DT_BF16 && gemm_tc_supported(p, mn)) gemm_tc(p, st);
    else if (x3 && gemm_x3_supported(p)) gemm_bf16x3(p, st);   // fp32 storage, tcgen05 arithmetic at fp32-class accuracy
    else { gemm_simt(p, st); ++g_launches; }
}

// Y[M,N] = epi( X[M,K] * W[N,K]^T )
void Engine::linear_fwd(const void* X, int64_t ldx, int64_t w_off, int N, int K, int M, void* Y, int y_dtype,
                        int64_t ldy, int flags, float alpha, const void* aux, const float* resid, DropCfg dc) {
    GemmProblem p;
    p.M = M; p.N = N; p.K = K;
    p.A.ptr = X; p.A.dtype = dt; p.A.major = MAJOR_K; p.A.ld = ldx;
    p.B.ptr = W(w_off); p.B.dtype = dt; p.B.major = MAJOR_K; p.B.ld = K;
    p.epi.C = Y; p.epi.c_dtype = y_dtype; p.epi.ldc = ldy; p.epi.alpha = alpha; p.epi.flags = flags;
    p.epi.aux = aux; p.epi.aux_dtype = dt; p.epi.resid = resid;
    if (dc.thr == 0) p.epi.flags &= ~EPI_DROPOUT;
    p.epi.seed = dc.seed; p.epi.site = dc.site; p.epi.drop_thr = dc.thr; p.epi.inv_keep = dc.inv_keep;
    gemm(p);
}

// dX[M,K] = epi( dY[M,N] * W[N,K] )      (B operand = W read in place as MN-major)
void Engine::linear_dgrad(const void* dY, int64_t lddy, int64_t w_off, int N, int K, int M, void* dX, int dx_dtype,
                          int64_t lddx, int flags, float alpha, const void* aux, bool accum_f32) {
    GemmProblem p;
    p.tail_filled = true;   // every dgrad is followed by its (independent, late-wait) wgrad
    p.M = M; p.N = K; p.K = N;
    p.A.ptr = dY; p.A.dtype = dt; p.A.major = MAJOR_K; p.A.ld = lddy;
    p.B.ptr = W(w_off); p.B.dtype = dt; p.B.major = MAJOR_MN; p.B.ld = K;
    p.epi.C = dX; p.epi.c_dtype = dx_dtype; p.epi.ldc = lddx; p.epi.alpha = alpha; p.epi.flags = flags;
    p.epi.aux = aux; p.epi.aux_dtype = dt;
    if (accum_f32) p.epi.flags |= EPI_ACCUM;
    gemm(p);
}

// G[w_off : N x K] += alpha * dY[M,N]^T * X[M,K]     (both operands read in place as MN-major, split-K over M)
// after_own_dgrad: the kernel launched just before this one is the dgrad of the same linear layer (reads dY and W,
// writes dX) — independent of the weight gradient, so the wgrad may start in the SM slots that idle during the dgrad's
// partial last wave (GemmProblem::indep_of_prev).
void Engine::linear_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t w_off, int N, int K, int M,
                          float alpha, bool after_own_dgrad) {
    GemmProblem p;
    p.indep_of_prev = after_own_dgrad;
    p.M = N; p.N = K; p.K = M;
    // split-K over the token dimension with 128 x 256 tiles: pick the split count (a divisor of M, >= 256 rows per
    // split) that minimises  rounds(tiles * splits / #SM) * (k-blocks per split + fixed tile overhead)
    int splits = 1;
    {
        const int64_t t256 = cdiv(N, 128) * cdiv(K, 256);
        double best = 1e30;
        for (int s = 1; s <= 32; s *= 2) {
            if (M % s != 0 || M / s < 256) break;
            const double rounds = (double)cdiv(t256 * s, 148);
            const double cost = rounds * ((double)(M / s) / 64.0 + 16.0);
            if (cost < best) { best = cost; splits = s; }
        }
    }
    if (x3) splits = 1;   // the bf16x3 parity GEMM takes non-batched operands
    p.prefer_bn = K > 128 ? 256 : (K > 64 ? 128 : 64);
    const int Ks = M / splits;
    p.K = Ks; p.nb1 = splits;
    p.A.ptr = dY; p.A.dtype = dt; p.A.major = MAJOR_MN; p.A.ld = lddy; p.A.bs1 = (int64_t)Ks * lddy;
    p.B.ptr = X; p.B.dtype = dt; p.B.major = MAJOR_MN; p.B.ld = ldx; p.B.bs1 = (int64_t)Ks * ldx;
    p.epi.C = G + w_off; p.epi.c_dtype = DT_F32; p.epi.ldc = K; p.epi.cs1 = 0; p.epi.alpha = alpha;
    p.epi.flags = EPI_ATOMIC;
    gemm(p);
}

// ------------------------------------------------------------------------------------------------------------
// inputs
// ------------------------------------------------------------------------------------------------------------
__global__ void shift_right_kernel(const int* __restrict__ labels, int* __restrict__ dec_ids, int B, int Ld) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Ld) return;
    const int t = i % Ld;
    int v = t == 0 ? 0 : labels[i - 1];
    if (v == -100) v = 0;
    dec_ids[i] = v;
}

void Engine::set_geometry(int B_, int Le_user_, int Ld_) {
    P5_CHECK(B_ >= 1 && B_ <= Bm, "batch exceeds max_batch");
    P5_CHECK(Le_user_ >= 1 && round_up(Le_user_, 8) <= Lem, "encoder length exceeds max_enc_len");
    P5_CHECK(Ld_ >= 1 && Ld_ <= Ldm, "decoder length exceeds max_dec_len");
    B = B_; Le_user = Le_user_; Le = (int)round_up(Le_user_, 8); Ld = Ld_;
    Me = (int64_t)B * Le; Md = (int64_t)B * Ld;
    packed = false; Mt = Me; Mt_true = Me;
}

void Engine::apply_lengths() {
    if ((int)pending_lens.size() != B) { pending_lens.clear(); return; }   // none (or stale) -> padded layout
    lens_h = pending_lens;
    pending_lens.clear();
    offs_h.assign(B + 1, 0);
    for (int b = 0; b < B; ++b) {
        P5_CHECK(lens_h[b] >= 1 && lens_h[b] <= Le_user, "encoder length out of range");
        offs_h[b + 1] = offs_h[b] + lens_h[b];
    }
    // Round the packed row count up to a multiple of 512 with FILLER rows (pad token 0, attended by nobody, gradient
    // exactly zero): token-dimension split-K of the weight-gradient GEMMs needs a row count with power-of-two divisors.
    Mt_true = offs_h[B];
    Mt = round_up(Mt_true, 512);
    packed = true;
    P5_CUDA(cudaMemcpyAsync(lens_d, lens_h.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
    P5_CUDA(cudaMemcpyAsync(offs_d, offs_h.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
}

void Engine::load_inputs(const int32_t* ids, const int32_t* mask, const int32_t* ww, const int32_t* lab) {
    const size_t wbytes = (size_t)Le_user * 4, pitch = (size_t)Le * 4;
    if (Le != Le_user) {
        P5_CUDA(cudaMemsetAsync(ids_e, 0, Me * 4, st));
        P5_CUDA(cudaMemsetAsync(mask_e, 0, Me * 4, st));
        P5_CUDA(cudaMemsetAsync(ww_e, 0, Me * 4, st));
    }
    P5_CUDA(cudaMemcpy2DAsync(ids_e, pitch, ids, wbytes, wbytes, B, cudaMemcpyDeviceToDevice, st));
    P5_CUDA(cudaMemcpy2DAsync(mask_e, pitch, mask, wbytes, wbytes, B, cudaMemcpyDeviceToDevice, st));
    if (ww) P5_CUDA(cudaMemcpy2DAsync(ww_e, pitch, ww, wbytes, wbytes, B, cudaMemcpyDeviceToDevice, st));
    else P5_CUDA(cudaMemsetAsync(ww_e, 0xff, Me * 4, st));  // -1 = no whole-word embedding
    if (packed) {
        if (Mt > Mt_true) {   // filler rows: pad token, no whole-word embedding
            P5_CUDA(cudaMemsetAsync(ids_p + Mt_true, 0, (Mt - Mt_true) * sizeof(int), st));
            P5_CUDA(cudaMemsetAsync(ww_p + Mt_true, 0xff, (Mt - Mt_true) * sizeof(int), st));
        }
        pack_ints(ids_e, ids_p, offs_d, lens_d, B, Le, st);
        pack_ints(ww_e, ww_p, offs_d, lens_d, B, Le, st);
    }
    if (lab) {
        P5_CUDA(cudaMemcpyAsync(labels, lab, Md * 4, cudaMemcpyDeviceToDevice, st));
        launch_k(shift_right_kernel, (unsigned)cdiv(Md, 256), 256, 0, st, labels, dec_ids, B, Ld);
        P5_CUDA(cudaGetLastError());
        ++g_launches;
    }
}

// relative_position_bucket (HF:modeling_t5.py:189-235) for every delta = key_pos - query_pos in [-(L-1), L-1]
static int rel_bucket(int rel, bool bidirectional, int num_buckets, int max_distance) {
    int b = 0, n = num_buckets;
    if (bidirectional) {
        n /= 2;
        if (rel > 0) b += n;
        rel = rel < 0 ? -rel : rel;
    } else {
        rel = rel < 0 ? -rel : 0;
    }
    const int max_exact = n / 2;
    if (rel < max_exact) return b + rel;
    // fp32 arithmetic as torch does it; +1e-5 guards the exact-integer cases (rel = max_exact * 2^k)
    const float v = logf((float)rel / (float)max_exact) / (float)log((double)max_distance / max_exact) * (float)(n - max_exact);
    int l = max_exact + (int)(v + 1e-5f);
    if (l > n - 1) l = n - 1;
    return b + l;
}

void Engine::build_bias(bool encoder, int L) {
    int& cur = encoder ? lut_enc_L : lut_dec_L;
    int* lut = encoder ? lut_enc : lut_dec;
    const int n_delta = 2 * L - 1;
    if (cur != L) {
        std::vector<int> h(n_delta);
        for (int i = 0; i < n_delta; ++i) h[i] = rel_bucket(i - (L - 1), encoder, cfg.rel_buckets, cfg.rel_max_distance);
        // pageable source: the call returns once the data is staged, so `h` may go out of scope
        P5_CUDA(cudaMemcpyAsync(lut, h.data(), n_delta * sizeof(int), cudaMemcpyHostToDevice, st));
        cur = L;
    }
    relbias_build(P + (encoder ? off_enc_rel : off_dec_rel), lut, encoder ? bias_enc : bias_dec, H, n_delta, st);
}

// ------------------------------------------------------------------------------------------------------------
// feed-forward sub-layer (HF:modeling_t5.py:84-150):  x_out = x + drop( Wo * drop(act(Wi * n)) )
// ------------------------------------------------------------------------------------------------------------
void Engine::ffn_fwd(const void* n, int64_t M, const FfnOff& w, void* z, void* h, const float* x_resid, float* x_out,
                     uint32_t kind_act, uint32_t kind_wo, int layer) {
    DropCfg none;
    if (!gated) {
        linear_fwd(n, d, w.wi, ff, d, (int)M, h, dt, ff, EPI_RELU | EPI_DROPOUT, 1.f, nullptr, nullptr, drop(kind_act, layer));
    } else {
        linear_fwd(n, d, w.wi, 2 * ff, d, (int)M, z, dt, 2 * ff, 0, 1.f, nullptr, nullptr, none);
        gated_gelu_fwd(z, h, dt, (int)M, ff, drop(kind_act, layer), st);
    }
    linear_fwd(h, ff, w.wo, d, ff, (int)M, x_out, DT_F32, d, EPI_ADD_RESID | EPI_DROPOUT, 1.f, nullptr, x_resid,
               drop(kind_wo, layer));
}

// in: dx_out = dL/dx_out (fp32).  out: dn_out (dt) = dL/dn; weight grads accumulated.
void Engine::ffn_bwd(const float* dx_out, int64_t M, const FfnOff& w, const void* n, const void* z, const void* h,
                     void* dn_out, uint32_t kind_act, uint32_t kind_wo, int layer, const void* gd, void* gff, bool skip_wgrad) {
    // gd (default g_d) = dropout-cast(dx_out) was already produced by the rmsnorm_bwd that computed dx_out
    // every linear: dgrad first, then its wgrad (independent of the dgrad -> fills the dgrad's partial last wave);
    // skip_wgrad: the caller batches the weight gradients of all layers afterwards (gd / gff are then per-layer buffers)
    if (!gd) gd = g_d;
    if (!gff) gff = g_ff;
    if (!gated) {
        const DropCfg da = drop(kind_act, layer);
        linear_dgrad(gd, d, w.wo, d, ff, (int)M, gff, dt, ff, EPI_MULPOS, da.inv_keep, h, false);
        if (!skip_wgrad) linear_wgrad(gd, d, h, ff, w.wo, d, ff, (int)M, 1.f, true);
        linear_dgrad(gff, ff, w.wi, ff, d, (int)M, dn_out, dt, d, 0, 1.f, nullptr, false);
        if (!skip_wgrad) linear_wgrad(gff, ff, n, d, w.wi, ff, d, (int)M, 1.f, true);
    } else {
        void* dh = poff(g_ff, (int64_t)M * 2 * ff, dt);   // g_ff = [dz (M x 2ff) | dh (M x ff)]
        linear_dgrad(g_d, d, w.wo, d, ff, (int)M, dh, dt, ff, 0, 1.f, nullptr, false);
        linear_wgrad(g_d, d, h, ff, w.wo, d, ff, (int)M, 1.f, true);
        gated_gelu_bwd(z, dh, g_ff, dt, (int)M, ff, drop(kind_act, layer), st);
        linear_dgrad(g_ff, 2 * ff, w.wi, 2 * ff, d, (int)M, dn_out, dt, d, 0, 1.f, nullptr, false);
        linear_wgrad(g_ff, 2 * ff, n, d, w.wi, 2 * ff, d, (int)M, 1.f, true);
    }
}

// ------------------------------------------------------------------------------------------------------------
// encoder self-attention
// ------------------------------------------------------------------------------------------------------------
// The attention kernels that work on the padded [B, Le] geometry (batched GEMMs of the backward, the SIMT kernels of
// the parity mode) see packed activations through padded scratch copies: unpack -> kernel -> pack.
void Engine::ensure_attn_scratch() {
    if (S_scr || dt != DT_BF16) return;
    const int64_t SS = (int64_t)Bm * H * Lem * Lem;
    const size_t e = esz();
    for (int i = 0; i < NE; ++i) P_e[i] = dalloc(SS * e);
    S_scr = dalloc_t<float>(SS);
    Pd_scr = dalloc(SS * e);
    dS_scr = dalloc(SS * e);
}

void Engine::enc_attention_fwd(int l) {
    const int64_t SS1 = (int64_t)Le * Le;
    static int fused = -1;
    if (fused < 0) { const char* e = getenv("P5_ATTN"); fused = (e && strcmp(e, "unfused") == 0) ? 0 : 1; }
    const DropCfg dc = drop(S_ENC_P, l);
    if (dt == DT_BF16 && (fused || packed)) {
        // the fused backward (Le <= 512) recomputes P from the row statistic, nothing of size L^2 is saved
        const bool fbwd = fattn_bwd_supported(Le);
        if (!fbwd) ensure_attn_scratch();
        const bool ok = fattn_fwd(qkv_e[l], 3 * A, A, B, H, Le, bias_enc, mask_e, fbwd ? nullptr : P_e[l],
                                  fbwd ? nullptr : lse_e[l], fbwd ? lse_e[l] : nullptr, ctx_e[l], A, dc, st,
                                  packed ? offs_d : nullptr, packed ? lens_d : nullptr, Mt);
        P5_CHECK(ok || !packed, "packed encoder attention needs the fused kernel (Le <= 512)");
        p_unnorm[l] = ok && !fbwd;     // P_e[l] holds un-normalised probabilities, lse_e[l] the 1 / row-sum factors
        p_fbwd[l] = ok && fbwd;        // lse_e[l] holds lse2 for fattn_bwd
        if (ok) return;
    }
    const void* qkv = qkv_e[l];
    void* ctx = ctx_e[l];
    if (packed) {
        unpack_rows(qkv_e[l], qkv_pad, offs_d, lens_d, B, Le, (int64_t)3 * A * esz(), st);
        qkv = qkv_pad; ctx = ctx_pad;
    }
    if (dt == DT_BF16) {
        ensure_attn_scratch();
        GemmProblem p;   // S = Q K^T   (unscaled, HF:modeling_t5.py:308)
        p.M = Le; p.N = Le; p.K = 64; p.nb1 = H; p.nb2 = B;
        p.A.ptr = qkv; p.A.dtype = dt; p.A.major = MAJOR_K; p.A.ld = 3 * A; p.A.bs1 = 64; p.A.bs2 = (int64_t)Le * 3 * A;
        p.B = p.A; p.B.ptr = poff(qkv, A, dt);
        p.epi.C = S_scr; p.epi.c_dtype = DT_F32; p.epi.ldc = Le; p.epi.cs1 = SS1; p.epi.cs2 = SS1 * H;
        gemm(p);
        softmax_fwd(S_scr, bias_enc, mask_e, P_e[l], Pd_scr, dt, B, H, Le, Le, 0, dc, st);
        GemmProblem q;   // ctx = Pd V
        q.M = Le; q.N = 64; q.K = Le; q.nb1 = H; q.nb2 = B;
        q.A.ptr = dc.thr ? Pd_scr : P_e[l]; q.A.dtype = dt; q.A.major = MAJOR_K; q.A.ld = Le; q.A.bs1 = SS1; q.A.bs2 = SS1 * H;
        q.B.ptr = poff(qkv, 2 * A, dt); q.B.dtype = dt; q.B.major = MAJOR_MN; q.B.ld = 3 * A; q.B.bs1 = 64;
        q.B.bs2 = (int64_t)Le * 3 * A;
        q.epi.C = ctx; q.epi.c_dtype = dt; q.epi.ldc = A; q.epi.cs1 = 64; q.epi.cs2 = (int64_t)Le * A;
        gemm(q);
    } else {
        AttnArgs a;
        a.B = B; a.H = H; a.Lq = Le; a.Lk = Le;
        a.q = {qkv, dt, 3 * A, (int64_t)Le * 3 * A};
        a.k = {poff(qkv, A, dt), dt, 3 * A, (int64_t)Le * 3 * A};
        a.v = {poff(qkv, 2 * A, dt), dt, 3 * A, (int64_t)Le * 3 * A};
        a.bias_rel = bias_enc; a.bias_off = Le - 1; a.n_delta = 2 * Le - 1;
        a.key_mask = mask_e; a.causal = 0; a.q_pos_offset = 0; a.row_map = nullptr; a.drop = dc;
        attn_simt_fwd(a, ctx, dt, A, (int64_t)Le * A, lse_e[l], st);
    }
    if (packed) pack_rows(ctx_pad, ctx_e[l], offs_d, lens_d, B, Le, (int64_t)A * esz(), st);
}

void Engine::enc_attention_bwd(int l, const void* dctx_in, void* dqkv_out) {
    const int64_t SS1 = (int64_t)Le * Le;
    const DropCfg dc = drop(S_ENC_P, l);
    const void* qkv = qkv_e[l];
    const void* dctx = dctx_in;
    void* dqkv = dqkv_out;
    const void* ctx_fwd = ctx_e[l];
    if (dt == DT_BF16 && p_fbwd[l]) {
        const bool ok = fattn_bwd(qkv_e[l], 3 * A, A, B, H, Le, bias_enc, mask_e, lse_e[l], ctx_e[l], A, dctx_in, A, dqkv_out, 3 * A,
                                  dbias_enc, dc, st, packed ? offs_d : nullptr, packed ? lens_d : nullptr, Mt);
        P5_CHECK(ok, "fused attention backward refused a shape its forward accepted");
        if (packed && Mt > Mt_true)   // filler rows carry zero gradient
            P5_CUDA(cudaMemsetAsync((char*)dqkv_out + Mt_true * 3 * A * esz(), 0, (Mt - Mt_true) * 3 * A * esz(), st));
        return;
    }
    if (packed) {
        unpack_rows(qkv_e[l], qkv_pad, offs_d, lens_d, B, Le, (int64_t)3 * A * esz(), st);
        unpack_rows(dctx_in, ctx_pad, offs_d, lens_d, B, Le, (int64_t)A * esz(), st);
        qkv = qkv_pad; dctx = ctx_pad;
        dqkv = dt == DT_F32 ? (void*)f_qkv_pad : dqkv_pad;
    }
    if (dt == DT_BF16) {
        ensure_attn_scratch();
        GemmProblem p;   // dPd = dctx V^T
        p.M = Le; p.N = Le; p.K = 64; p.nb1 = H; p.nb2 = B;
        p.A.ptr = dctx; p.A.dtype = dt; p.A.major = MAJOR_K; p.A.ld = A; p.A.bs1 = 64; p.A.bs2 = (int64_t)Le * A;
        p.B.ptr = poff(qkv, 2 * A, dt); p.B.dtype = dt; p.B.major = MAJOR_K; p.B.ld = 3 * A; p.B.bs1 = 64;
        p.B.bs2 = (int64_t)Le * 3 * A;
        p.epi.C = S_scr; p.epi.c_dtype = DT_F32; p.epi.ldc = Le; p.epi.cs1 = SS1; p.epi.cs2 = SS1 * H;
        gemm(p);
        // regenerate Pd when dropout is on, when packed (P_save holds stale rows/columns outside the sequences) or when
        // the fused forward saved un-normalised probabilities
        const bool regen = dc.thr || packed || p_unnorm[l];
        softmax_bwd(S_scr, P_e[l], dS_scr, regen ? Pd_scr : nullptr, dt, dbias_enc, B, H, Le, Le, dc, st, packed ? lens_d : nullptr,
                    p_unnorm[l] ? lse_e[l] : nullptr);
        const void* Pd = regen ? Pd_scr : P_e[l];
        GemmProblem v;   // dV[j,c] = sum_i Pd[i,j] dctx[i,c]
        v.M = Le; v.N = 64; v.K = Le; v.nb1 = H; v.nb2 = B;
        v.A.ptr = Pd; v.A.dtype = dt; v.A.major = MAJOR_MN; v.A.ld = Le; v.A.bs1 = SS1; v.A.bs2 = SS1 * H;
        v.B.ptr = dctx; v.B.dtype = dt; v.B.major = MAJOR_MN; v.B.ld = A; v.B.bs1 = 64; v.B.bs2 = (int64_t)Le * A;
        v.epi.C = poff(dqkv, 2 * A, dt); v.epi.c_dtype = dt; v.epi.ldc = 3 * A; v.epi.cs1 = 64; v.epi.cs2 = (int64_t)Le * 3 * A;
        gemm(v);
        GemmProblem q;   // dQ[i,c] = sum_j dS[i,j] K[j,c]
        q.M = Le; q.N = 64; q.K = Le; q.nb1 = H; q.nb2 = B;
        q.A.ptr = dS_scr; q.A.dtype = dt; q.A.major = MAJOR_K; q.A.ld = Le; q.A.bs1 = SS1; q.A.bs2 = SS1 * H;
        q.B.ptr = poff(qkv, A, dt); q.B.dtype = dt; q.B.major = MAJOR_MN; q.B.ld = 3 * A; q.B.bs1 = 64;
        q.B.bs2 = (int64_t)Le * 3 * A;
        q.epi.C = dqkv; q.epi.c_dtype = dt; q.epi.ldc = 3 * A; q.epi.cs1 = 64; q.epi.cs2 = (int64_t)Le * 3 * A;
        gemm(q);
        GemmProblem k;   // dK[j,c] = sum_i dS[i,j] Q[i,c]
        k.M = Le; k.N = 64; k.K = Le; k.nb1 = H; k.nb2 = B;
        k.A.ptr = dS_scr; k.A.dtype = dt; k.A.major = MAJOR_MN; k.A.ld = Le; k.A.bs1 = SS1; k.A.bs2 = SS1 * H;
        k.B.ptr = qkv; k.B.dtype = dt; k.B.major = MAJOR_MN; k.B.ld = 3 * A; k.B.bs1 = 64; k.B.bs2 = (int64_t)Le * 3 * A;
        k.epi.C = poff(dqkv, A, dt); k.epi.c_dtype = dt; k.epi.ldc = 3 * A; k.epi.cs1 = 64; k.epi.cs2 = (int64_t)Le * 3 * A;
        gemm(k);
    } else {
        // fp32 parity path: dqkv IS an fp32 buffer
        if (packed) {   // the forward output in padded geometry (for delta = sum dO * O)
            unpack_rows(ctx_e[l], dqkv_pad, offs_d, lens_d, B, Le, (int64_t)A * esz(), st);
            ctx_fwd = dqkv_pad;
        }
        if (attn_bwd_needs_zero(Le)) P5_CUDA(cudaMemsetAsync(dqkv, 0, Me * 3 * A * sizeof(float), st));
        AttnArgs a;
        a.B = B; a.H = H; a.Lq = Le; a.Lk = Le;
        a.q = {qkv, dt, 3 * A, (int64_t)Le * 3 * A};
        a.k = {poff(qkv, A, dt), dt, 3 * A, (int64_t)Le * 3 * A};
        a.v = {poff(qkv, 2 * A, dt), dt, 3 * A, (int64_t)Le * 3 * A};
        a.bias_rel = bias_enc; a.bias_off = Le - 1; a.n_delta = 2 * Le - 1;
        a.key_mask = mask_e; a.causal = 0; a.q_pos_offset = 0; a.row_map = nullptr; a.drop = dc;
        float* f = (float*)dqkv;
        attn_simt_bwd(a, ctx_fwd, dctx, dt, A, (int64_t)Le * A, lse_e[l], f, 3 * A, (int64_t)Le * 3 * A, f + A,
                      f + 2 * A, 3 * A, (int64_t)Le * 3 * A, dbias_enc, st);
    }
    if (packed) {
        pack_rows(dqkv, dqkv_out, offs_d, lens_d, B, Le, (int64_t)3 * A * esz(), st);
        if (Mt > Mt_true)   // filler rows carry zero gradient (pack_rows writes the real rows only)
            P5_CUDA(cudaMemsetAsync((char*)dqkv_out + Mt_true * 3 * A * esz(), 0, (Mt - Mt_true) * 3 * A * esz(), st));
    }
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
void Engine::encoder_forward() {
    wait_opt(0);                 // embeddings
    if (NE > 0) wait_opt(1);     // encoder layer 0 holds the shared relative-bias table
    build_bias(true, Le);
    const int* ids_in = packed ? ids_p : ids_e;
    const int* ww_in = packed ? ww_p : ww_e;
    embed_fwd(P + off_shared, P + off_ww, ids_in, ww_in, xe[0], (int)Mt, d, V, cfg.whole_word_rows, drop(S_EMB_E, 0), st);
    DropCfg none;
    for (int l = 0; l < NE; ++l) {
        const EncLayerOff& w = enc[l];
        wait_opt(1 + l);          // asynchronous AdamW: this layer's range must be updated before its first GEMM
        float *x_in = xe[2 * l], *x_mid = xe[2 * l + 1], *x_out = xe[2 * l + 2];
        rmsnorm_fwd(x_in, P + w.ln0, ne[2 * l], dt, rstd_e[2 * l], (int)Mt, d, cfg.ln_eps, none, st);
        linear_fwd(ne[2 * l], d, w.sa.q, 3 * A, d, (int)Mt, qkv_e[l], dt, 3 * A, 0, 1.f, nullptr, nullptr, none);
        enc_attention_fwd(l);
        linear_fwd(ctx_e[l], A, w.sa.o, d, A, (int)Mt, x_mid, DT_F32, d, EPI_ADD_RESID | EPI_DROPOUT, 1.f, nullptr, x_in,
                   drop(S_ENC_O, l));
        rmsnorm_fwd(x_mid, P + w.ln1, ne[2 * l + 1], dt, rstd_e[2 * l + 1], (int)Mt, d, cfg.ln_eps, none, st);
        ffn_fwd(ne[2 * l + 1], Mt, w.ff, z_e[l], h_e[l], x_mid, x_out, S_ENC_ACT, S_ENC_WO, l);
    }
    rmsnorm_fwd(xe[2 * NE], P + off_enc_final, enc_out, dt, rstd_e[2 * NE], (int)Mt, d, cfg.ln_eps, drop(S_ENC_FINAL, 0), st);
}

static AttnArgs dec_self_args(Engine& e, int l, DropCfg dc) {
    AttnArgs a;
    const int A = e.A, Ld = e.Ld;
    a.B = e.B; a.H = e.H; a.Lq = Ld; a.Lk = Ld;
    a.q = {e.sqkv[l], e.dt, 3 * A, (int64_t)Ld * 3 * A};
    a.k = {poff(e.sqkv[l], A, e.dt), e.dt, 3 * A, (int64_t)Ld * 3 * A};
    a.v = {poff(e.sqkv[l], 2 * A, e.dt), e.dt, 3 * A, (int64_t)Ld * 3 * A};
    a.bias_rel = e.bias_dec; a.bias_off = Ld - 1; a.n_delta = 2 * Ld - 1;
    a.key_mask = nullptr; a.causal = 1; a.q_pos_offset = 0; a.row_map = nullptr; a.drop = dc;
    return a;
}
static AttnArgs dec_cross_args(Engine& e, int l, DropCfg dc) {
    AttnArgs a;
    const int A = e.A, Ld = e.Ld, Le = e.Le;
    a.B = e.B; a.H = e.H; a.Lq = Ld; a.Lk = Le;
    a.q = {e.cq[l], e.dt, A, (int64_t)Ld * A};
    a.k = {e.ckv[l], e.dt, e.ckv_ld, (int64_t)Le * e.ckv_ld};
    a.v = {poff(e.ckv[l], A, e.dt), e.dt, e.ckv_ld, (int64_t)Le * e.ckv_ld};
    a.bias_rel = nullptr; a.bias_off = 0; a.n_delta = 0;   // cross-attention position bias is zero (HF:...:317-322)
    a.key_mask = e.mask_e; a.causal = 0; a.q_pos_offset = 0; a.row_map = nullptr; a.drop = dc;
    if (e.packed) {   // encoder rows are packed: per-user row offset + key count replace (batch stride, key mask)
        a.key_mask = nullptr; a.kv_off = e.offs_d; a.kv_len = e.lens_d;
    }
    return a;
}

// cross K|V of every decoder layer in ONE batched GEMM: C[:, l*2A : (l+1)*2A] = enc_out . W_l^T  (batch dim 2 = layer:
// A broadcast, B strided by the per-layer parameter stride, C strided by 2A columns).  7200 tiles instead of 12 x 600:
// no per-layer wave quantisation (600 tiles = 4.05 waves on 148 SMs).
void Engine::project_cross_kv_all(int64_t rows) {
    GemmProblem p;
    p.M = (int)rows; p.N = 2 * A; p.K = d; p.nb1 = 1; p.nb2 = ND;
    p.A.ptr = enc_out; p.A.dtype = dt; p.A.major = MAJOR_K; p.A.ld = d; p.A.bcast2 = true;
    p.B.ptr = W(dec[0].ca.k); p.B.dtype = dt; p.B.major = MAJOR_K; p.B.ld = d; p.B.bs2 = dec_layer_stride;
    p.epi.C = ckv_all; p.epi.c_dtype = dt; p.epi.ldc = ckv_ld; p.epi.cs2 = 2 * A; p.epi.alpha = 1.f; p.epi.flags = 0;
    gemm(p);
}

void Engine::decoder_forward() {
    wait_opt(0);
    if (ND > 0) wait_opt(1 + NE);
    if (batched_ckv) {
        for (int l = 1; l < ND; ++l) wait_opt(1 + NE + l);
        project_cross_kv_all(Mt);
    }
    build_bias(false, Ld);
    embed_fwd(P + off_shared, nullptr, dec_ids, nullptr, yd[0], (int)Md, d, V, cfg.whole_word_rows, drop(S_EMB_D, 0), st);
    DropCfg none;
    for (int l = 0; l < ND; ++l) {
        const DecLayerOff& w = dec[l];
        wait_opt(1 + NE + l);
        float *y0 = yd[3 * l], *y1 = yd[3 * l + 1], *y2 = yd[3 * l + 2], *y3 = yd[3 * l + 3];
        rmsnorm_fwd(y0, P + w.ln0, nd[3 * l], dt, rstd_d[3 * l], (int)Md, d, cfg.ln_eps, none, st);
        linear_fwd(nd[3 * l], d, w.sa.q, 3 * A, d, (int)Md, sqkv[l], dt, 3 * A, 0, 1.f, nullptr, nullptr, none);
        {
            const AttnArgs sa = dec_self_args(*this, l, drop(S_DEC_SP, l));
            if (dattn_supported(sa)) dattn_fwd(sa, sctx[l], A, (int64_t)Ld * A, slse[l], st);
            else attn_simt_fwd(sa, sctx[l], dt, A, (int64_t)Ld * A, slse[l], st);
        }
        linear_fwd(sctx[l], A, w.sa.o, d, A, (int)Md, y1, DT_F32, d, EPI_ADD_RESID | EPI_DROPOUT, 1.f, nullptr, y0,
                   drop(S_DEC_SO, l));
        rmsnorm_fwd(y1, P + w.ln1, nd[3 * l + 1], dt, rstd_d[3 * l + 1], (int)Md, d, cfg.ln_eps, none, st);
        linear_fwd(nd[3 * l + 1], d, w.ca.q, A, d, (int)Md, cq[l], dt, A, 0, 1.f, nullptr, nullptr, none);
        if (!batched_ckv) {
            next_gemm_indep = true;   // cross K|V projection: needs enc_out only, not the (small) cross-Q GEMM just before it
            linear_fwd(enc_out, d, w.ca.k, 2 * A, d, (int)Mt, ckv[l], dt, ckv_ld, 0, 1.f, nullptr, nullptr, none);
        }
        {
            const AttnArgs ca = dec_cross_args(*this, l, drop(S_DEC_CP, l));
            if (dattn_supported(ca)) dattn_fwd(ca, cctx[l], A, (int64_t)Ld * A, clse[l], st);
            else attn_simt_fwd(ca, cctx[l], dt, A, (int64_t)Ld * A, clse[l], st);
        }
        linear_fwd(cctx[l], A, w.ca.o, d, A, (int)Md, y2, DT_F32, d, EPI_ADD_RESID | EPI_DROPOUT, 1.f, nullptr, y1,
                   drop(S_DEC_CO, l));
        rmsnorm_fwd(y2, P + w.ln2, nd[3 * l + 2], dt, rstd_d[3 * l + 2], (int)Md, d, cfg.ln_eps, none, st);
        ffn_fwd(nd[3 * l + 2], Md, w.ff, z_d[l], h_d[l], y2, y3, S_DEC_ACT, S_DEC_WO, l);
    }
    rmsnorm_fwd(yd[3 * ND], P + off_dec_final, dec_out, dt, rstd_d[3 * ND], (int)Md, d, cfg.ln_eps, drop(S_DEC_FINAL, 0), st);
}

void Engine::head_forward() {
    DropCfg none;
    // logits = (h * d_model^-0.5) * shared^T   (P5_T5.py:357-361)
    linear_fwd(dec_out, d, off_shared, V, d, (int)Md, logits, DT_F32, Vpad, 0, 1.f / sqrtf((float)d), nullptr, nullptr, none);
    ce_fwd(logits, Vpad, labels, loss_tok, lse_ce, (int)Md, V, st);
}

void Engine::forward(const int32_t* ids, const int32_t* mask, const int32_t* ww, const int32_t* lab, int B_, int Le_,
                     int Ld_, bool train, uint64_t seed_) {
    P5_CUDA(cudaSetDevice(device));
    set_geometry(B_, Le_, Ld_);
    apply_lengths();
    training = train; seed = seed_;
    if (shadow_stale) refresh_shadow();
    load_inputs(ids, mask, ww, lab);
    encoder_forward();
    decoder_forward();
    head_forward();
    opt_pending = false;      // every optimiser range has been waited for by now
    have_fwd = true;
}

// ------------------------------------------------------------------------------------------------------------
// decoder weight gradients, batched over the layers: dW_l = dY_l^T . X_l for the six linear layers of a decoder block
// (FFN wo / wi, cross-attention o / q, self-attention o / q|k|v).  One GEMM per weight type, batch dimension 2 = layer:
// dY and X are read from the per-layer slabs (MN-major, in place), dW lands in the flat gradient buffer at the layer
// stride.  The self-attention weights of layer 0 sit before the relative-bias table, so their stride to layer 1 differs:
// layer 0 is its own launch.
// ------------------------------------------------------------------------------------------------------------
void Engine::decoder_wgrads_batched() {
    const int64_t rs = (int64_t)Ldm * Bm;            // rows per layer slab
    auto run = [&](const void* dY, int64_t wy, int64_t sy, const void* X, int64_t wx, int64_t sx, int64_t w_off, int N, int K, int l0, int nl) {
        if (nl <= 0) return;
        GemmProblem p;
        p.indep_of_prev = true;
        p.M = N; p.N = K; p.K = (int)Md; p.nb1 = 1; p.nb2 = nl;
        p.prefer_bn = K > 128 ? 256 : (K > 64 ? 128 : 64);
        p.A.ptr = poff(dY, (int64_t)l0 * sy, dt); p.A.dtype = dt; p.A.major = MAJOR_MN; p.A.ld = wy; p.A.bs2 = sy;
        p.B.ptr = poff(X, (int64_t)l0 * sx, dt); p.B.dtype = dt; p.B.major = MAJOR_MN; p.B.ld = wx; p.B.bs2 = sx;
        p.epi.C = G + w_off; p.epi.c_dtype = DT_F32; p.epi.ldc = K; p.epi.cs2 = dec_layer_stride; p.epi.alpha = 1.f;
        p.epi.flags = EPI_ATOMIC;
        gemm(p);
    };
    // FFN and cross-attention weights come after the relative-bias table of layer 0: one stride for all layers
    run(gdw_all, d, rs * d, h_d_all, ff, rs * ff, dec[0].ff.wo, d, ff, 0, ND);
    run(gff_all, ff, rs * ff, poff(nd_all, 2 * rs * d, dt), d, 3 * rs * d, dec[0].ff.wi, ff, d, 0, ND);
    run(gdc_all, d, rs * d, cctx_all, A, rs * A, dec[0].ca.o, d, A, 0, ND);
    run(gcq_all, A, rs * A, poff(nd_all, rs * d, dt), d, 3 * rs * d, dec[0].ca.q, A, d, 0, ND);
    // self-attention weights: layer 0 alone, layers 1.. batched
    run(gds_all, d, rs * d, sctx_all, A, rs * A, dec[0].sa.o, d, A, 0, 1);
    run(gds_all, d, rs * d, sctx_all, A, rs * A, dec[1].sa.o, d, A, 1, ND - 1);
    run(gsq_all, 3 * A, rs * 3 * A, nd_all, d, 3 * rs * d, dec[0].sa.q, 3 * A, d, 0, 1);
    run(gsq_all, 3 * A, rs * 3 * A, nd_all, d, 3 * rs * d, dec[1].sa.q, 3 * A, d, 1, ND - 1);
}

// ------------------------------------------------------------------------------------------------------------
// backward (consumes this->dloss = dL/dloss_tok)
// ------------------------------------------------------------------------------------------------------------
void Engine::backward() {
    P5_CHECK(have_fwd, "p5_backward called without a preceding p5_forward");
    P5_CUDA(cudaSetDevice(device));
    join_optimizer();
    invalidate_norm();
    early_np = 0; early_cov = 0; early_lo = n_flat;
    const float hs = 1.f / sqrtf((float)d);
    auto as_T = [&](float* src, void* dst, int64_t n) -> void* {
        if (dt == DT_F32) return (void*)src;
        cast_f32_to(src, dst, dt, n, st);
        return dst;
    };
    // ---- head
    ce_bwd(logits, Vpad, lse_ce, labels, dloss, dlogits, dt, (int)Md, V, Vpad, st);
    linear_dgrad(dlogits, Vpad, off_shared, V, d, (int)Md, g_d2, dt, d, 0, hs, nullptr, false);
    linear_wgrad(dlogits, Vpad, dec_out, d, off_shared, V, d, (int)Md, hs, true);
    float* dy = dx_a;
    // per-layer dY buffers when the decoder weight gradients are batched over the layers afterwards (decoder_wgrads_batched)
    const bool bw = batched_dwg && dattn_supported(dec_self_args(*this, 0, DropCfg())) && dattn_supported(dec_cross_args(*this, 0, DropCfg()));
    auto L_ = [&](void* slab, int l, int64_t width) -> void* { return poff(slab, (int64_t)l * Ldm * Bm * width, dt); };
    auto gdw = [&](int l) -> void* { return bw ? L_(gdw_all, l, d) : g_d; };
    auto gdc = [&](int l) -> void* { return bw ? L_(gdc_all, l, d) : g_d; };
    auto gds = [&](int l) -> void* { return bw ? L_(gds_all, l, d) : g_d; };
    rmsnorm_bwd(g_d2, dt, yd[3 * ND], rstd_d[3 * ND], P + off_dec_final, nullptr, dy, G + off_dec_final, (int)Md, d,
                drop(S_DEC_FINAL, 0), st, ND > 0 ? gdw(ND - 1) : g_d, dt, drop(S_DEC_WO, ND - 1));
    P5_CUDA(cudaMemsetAsync(d_encout, 0, Mt * d * sizeof(float), st));
    P5_CUDA(cudaMemsetAsync(dbias_dec, 0, (size_t)H * (2 * Ld) * sizeof(float), st));
    DropCfg none;
    // ---- decoder blocks
    for (int l = ND - 1; l >= 0; --l) {
        const DecLayerOff& w = dec[l];
        float *y0 = yd[3 * l], *y1 = yd[3 * l + 1], *y2 = yd[3 * l + 2];
        ffn_bwd(dy, Md, w.ff, nd[3 * l + 2], z_d[l], h_d[l], g_d2, S_DEC_ACT, S_DEC_WO, l, gdw(l), bw ? L_(gff_all, l, ff) : nullptr, bw);
        rmsnorm_bwd(g_d2, dt, y2, rstd_d[3 * l + 2], P + w.ln2, dy, dy, G + w.ln2, (int)Md, d, none, st, gdc(l), dt,
                    drop(S_DEC_CO, l));
        // cross attention (gdc = dropout-cast(dy))
        linear_dgrad(gdc(l), d, w.ca.o, d, A, (int)Md, g_ctx, dt, A, 0, 1.f, nullptr, false);
        if (!bw) linear_wgrad(g_d, d, cctx[l], A, w.ca.o, d, A, (int)Md, 1.f, true);
        void *gq, *gkv = nullptr;
        const AttnArgs ca = dec_cross_args(*this, l, drop(S_DEC_CP, l));
        const bool batch_kv = batched_ckv && dattn_supported(ca);     // dK|dV of all layers -> one wgrad / dgrad after the loop
        if (dattn_supported(ca)) {   // bf16: tensor-core kernel writes dQ and dK|dV as bf16 in place
            void* dk = batch_kv ? poff(g_ckv_all, (int64_t)l * 2 * A, dt) : g_ckv;
            const int64_t ldkv = batch_kv ? ckv_ld : 2 * A;
            void* dq = bw ? L_(gcq_all, l, A) : g_qkv;
            dattn_bwd(ca, g_ctx, A, (int64_t)Ld * A, clse[l], dq, A, (int64_t)Ld * A, dk, poff(dk, A, dt), ldkv,
                      (int64_t)Le * ldkv, nullptr, st);
            if (!batch_kv && packed && Mt > Mt_true)   // filler rows of dK|dV: zero gradient
                P5_CUDA(cudaMemsetAsync(poff(g_ckv, Mt_true * 2 * A, dt), 0, (Mt - Mt_true) * 2 * A * dtype_size(dt), st));
            gq = dq;
            gkv = g_ckv;
        } else {
            if (attn_bwd_needs_zero(Ld)) P5_CUDA(cudaMemsetAsync(f_ckv, 0, Mt * 2 * A * sizeof(float), st));
            attn_simt_bwd(ca, cctx[l], g_ctx, dt, A, (int64_t)Ld * A, clse[l], f_qkv, A, (int64_t)Ld * A, f_ckv, f_ckv + A, 2 * A,
                          (int64_t)Le * 2 * A, nullptr, st);
            if (packed && Mt > Mt_true)   // filler rows of dK|dV: zero gradient
                P5_CUDA(cudaMemsetAsync(f_ckv + Mt_true * 2 * A, 0, (Mt - Mt_true) * 2 * A * sizeof(float), st));
            gq = as_T(f_qkv, g_qkv, Md * A);
            gkv = as_T(f_ckv, g_ckv, Mt * 2 * A);
        }
        if (!batch_kv) {
            linear_dgrad(gkv, 2 * A, w.ca.k, 2 * A, d, (int)Mt, d_encout, DT_F32, d, 0, 1.f, nullptr, true);
            linear_wgrad(gkv, 2 * A, enc_out, d, w.ca.k, 2 * A, d, (int)Mt, 1.f, true);
            next_gemm_indep = true;   // reads gq and W only: independent of the cross-K|V dgrad / wgrad before it
        }
        linear_dgrad(gq, A, w.ca.q, A, d, (int)Md, g_d2, dt, d, 0, 1.f, nullptr, false);
        if (!bw) linear_wgrad(gq, A, nd[3 * l + 1], d, w.ca.q, A, d, (int)Md, 1.f, true);
        rmsnorm_bwd(g_d2, dt, y1, rstd_d[3 * l + 1], P + w.ln1, dy, dy, G + w.ln1, (int)Md, d, none, st, gds(l), dt,
                    drop(S_DEC_SO, l));
        // self attention (gds = dropout-cast(dy))
        linear_dgrad(gds(l), d, w.sa.o, d, A, (int)Md, g_ctx, dt, A, 0, 1.f, nullptr, false);
        if (!bw) linear_wgrad(g_d, d, sctx[l], A, w.sa.o, d, A, (int)Md, 1.f, true);
        void* gqkv;
        const AttnArgs sa = dec_self_args(*this, l, drop(S_DEC_SP, l));
        if (dattn_supported(sa)) {
            void* dq = bw ? L_(gsq_all, l, 3 * A) : g_qkv;
            dattn_bwd(sa, g_ctx, A, (int64_t)Ld * A, slse[l], dq, 3 * A, (int64_t)Ld * 3 * A, poff(dq, A, dt),
                      poff(dq, 2 * A, dt), 3 * A, (int64_t)Ld * 3 * A, dbias_dec, st);
            gqkv = dq;
        } else {
            if (attn_bwd_needs_zero(Ld)) P5_CUDA(cudaMemsetAsync(f_qkv, 0, Md * 3 * A * sizeof(float), st));
            attn_simt_bwd(sa, sctx[l], g_ctx, dt, A, (int64_t)Ld * A, slse[l], f_qkv, 3 * A, (int64_t)Ld * 3 * A, f_qkv + A,
                          f_qkv + 2 * A, 3 * A, (int64_t)Ld * 3 * A, dbias_dec, st);
            gqkv = as_T(f_qkv, g_qkv, Md * 3 * A);
        }
        linear_dgrad(gqkv, 3 * A, w.sa.q, 3 * A, d, (int)Md, g_d2, dt, d, 0, 1.f, nullptr, false);
        if (!bw) linear_wgrad(gqkv, 3 * A, nd[3 * l], d, w.sa.q, 3 * A, d, (int)Md, 1.f, true);
        rmsnorm_bwd(g_d2, dt, y0, rstd_d[3 * l], P + w.ln0, dy, dy, G + w.ln0, (int)Md, d, none, st, l > 0 ? gdw(l - 1) : nullptr, dt,
                    drop(S_DEC_WO, l > 0 ? l - 1 : 0));
    }
    if (bw) decoder_wgrads_batched();
    if (batched_ckv && ND > 0 && dattn_supported(dec_cross_args(*this, 0, none))) {
        if (packed && Mt > Mt_true)   // filler rows of every layer's dK|dV: zero gradient
            P5_CUDA(cudaMemsetAsync(poff(g_ckv_all, Mt_true * ckv_ld, dt), 0, (Mt - Mt_true) * ckv_ld * dtype_size(dt), st));
        {   // d_encout += sum_l gkv_l . W_l : the layers are batch dimension 2, all accumulating into the same (zeroed) C
            GemmProblem p;
            p.tail_filled = true;
            p.M = (int)Mt; p.N = d; p.K = 2 * A; p.nb1 = 1; p.nb2 = ND;
            p.A.ptr = g_ckv_all; p.A.dtype = dt; p.A.major = MAJOR_K; p.A.ld = ckv_ld; p.A.bs2 = 2 * A;
            p.B.ptr = W(dec[0].ca.k); p.B.dtype = dt; p.B.major = MAJOR_MN; p.B.ld = d; p.B.bs2 = dec_layer_stride;
            p.epi.C = d_encout; p.epi.c_dtype = DT_F32; p.epi.ldc = d; p.epi.cs2 = 0; p.epi.alpha = 1.f; p.epi.flags = EPI_ATOMIC;
            gemm(p);
        }
        {   // dW_l = gkv_l^T . enc_out for every layer: batch dim 2 = layer (B broadcast), batch dim 1 = split-K
            GemmProblem p;
            p.indep_of_prev = true;
            p.M = 2 * A; p.N = d;
            int splits = 1;
            {
                const int64_t t256 = cdiv(2 * A, 128) * cdiv(d, 256) * ND;
                double best = 1e30;
                for (int sp = 1; sp <= 32; sp *= 2) {
                    if (Mt % sp != 0 || Mt / sp < 256) break;
                    const double rounds = (double)cdiv(t256 * sp, 148);
                    const double cost = rounds * ((double)(Mt / sp) / 64.0 + 16.0);
                    if (cost < best) { best = cost; splits = sp; }
                }
            }
            const int64_t Ks = Mt / splits;
            p.K = (int)Ks; p.nb1 = splits; p.nb2 = ND; p.prefer_bn = 256;
            p.A.ptr = g_ckv_all; p.A.dtype = dt; p.A.major = MAJOR_MN; p.A.ld = ckv_ld; p.A.bs1 = Ks * ckv_ld; p.A.bs2 = 2 * A;
            p.B.ptr = enc_out; p.B.dtype = dt; p.B.major = MAJOR_MN; p.B.ld = d; p.B.bs1 = Ks * d; p.B.bcast2 = true;
            p.epi.C = G + dec[0].ca.k; p.epi.c_dtype = DT_F32; p.epi.ldc = d; p.epi.cs1 = 0; p.epi.cs2 = dec_layer_stride;
            p.epi.alpha = 1.f; p.epi.flags = EPI_ATOMIC;
            gemm(p);
        }
    }
    embed_bwd(dy, dec_ids, nullptr, G + off_shared, nullptr, (int)Md, d, V, cfg.whole_word_rows, drop(S_EMB_D, 0), st);
    relbias_scatter_grad(dbias_dec, lut_dec, G + off_dec_rel, H, 2 * Ld - 1, st);
    // every decoder gradient is final: hand the range to NCCL while the encoder backward runs
    if (ND > 0) range_final(dec[0].sa.q, n_flat - dec[0].sa.q);

    // ---- encoder
    float* dx = dx_a;
    rmsnorm_bwd(d_encout, DT_F32, xe[2 * NE], rstd_e[2 * NE], P + off_enc_final, nullptr, dx, G + off_enc_final, (int)Mt, d,
                drop(S_ENC_FINAL, 0), st, g_d, dt, drop(S_ENC_WO, NE - 1));
    P5_CUDA(cudaMemsetAsync(dbias_enc, 0, (size_t)H * (2 * Le) * sizeof(float), st));
    for (int l = NE - 1; l >= 0; --l) {
        const EncLayerOff& w = enc[l];
        float *x_in = xe[2 * l], *x_mid = xe[2 * l + 1];
        ffn_bwd(dx, Mt, w.ff, ne[2 * l + 1], z_e[l], h_e[l], g_d2, S_ENC_ACT, S_ENC_WO, l);
        rmsnorm_bwd(g_d2, dt, x_mid, rstd_e[2 * l + 1], P + w.ln1, dx, dx, G + w.ln1, (int)Mt, d, none, st, g_d, dt,
                    drop(S_ENC_O, l));
        linear_dgrad(g_d, d, w.sa.o, d, A, (int)Mt, g_ctx, dt, A, 0, 1.f, nullptr, false);
        linear_wgrad(g_d, d, ctx_e[l], A, w.sa.o, d, A, (int)Mt, 1.f, true);
        void* dqkv = dt == DT_F32 ? (void*)f_qkv : g_qkv;
        enc_attention_bwd(l, g_ctx, dqkv);
        linear_dgrad(dqkv, 3 * A, w.sa.q, 3 * A, d, (int)Mt, g_d2, dt, d, 0, 1.f, nullptr, false);
        linear_wgrad(dqkv, 3 * A, ne[2 * l], d, w.sa.q, 3 * A, d, (int)Mt, 1.f, true);
        rmsnorm_bwd(g_d2, dt, x_in, rstd_e[2 * l], P + w.ln0, dx, dx, G + w.ln0, (int)Mt, d, none, st, l > 0 ? g_d : nullptr, dt,
                    drop(S_ENC_WO, l > 0 ? l - 1 : 0));
        // block l >= 1 is final (block 0 also holds the shared relative bias, reduced with the embeddings at the end)
        if (l >= 1) {
            const int64_t hi = (l + 1 < NE) ? enc[l + 1].sa.q : (ND > 0 ? dec[0].sa.q : n_flat);
            range_final(w.sa.q, hi - w.sa.q);
        }
    }
    embed_bwd(dx, packed ? ids_p : ids_e, packed ? ww_p : ww_e, G + off_shared, G + off_ww, (int)Mt, d, V, cfg.whole_word_rows,
              drop(S_EMB_E, 0), st);
    relbias_scatter_grad(dbias_enc, lut_enc, G + off_enc_rel, H, 2 * Le - 1, st);
    if (early_norm && early_cov > 0 && early_lo + early_cov == n_flat) {     // the early ranges tile [early_lo, n_flat)
        P5_CUDA(cudaEventRecord(ev_norm_join, overlap_comm ? comm_stream(this) : st_norm));
        early_norm_ready = true;
    }
}

// A gradient range [off, off + n) of the flat buffer has received its last contribution of this backward pass.
void Engine::range_final(int64_t off, int64_t n) {
    if (overlap_comm) comm_allreduce_range(this, off, n);
    if (!early_norm || n <= 0) return;
    if (!ev_norm_join) {
        P5_CUDA(cudaStreamCreateWithFlags(&st_norm, cudaStreamNonBlocking));
        P5_CUDA(cudaEventCreateWithFlags(&ev_norm_fork, cudaEventDisableTiming));
        P5_CUDA(cudaEventCreateWithFlags(&ev_norm_join, cudaEventDisableTiming));
    }
    cudaStream_t s = overlap_comm ? comm_stream(this) : st_norm;     // data parallel: ordered after the range's all-reduce
    P5_CHECK(s != nullptr, "range_final: no side stream");
    if (!overlap_comm) {
        P5_CUDA(cudaEventRecord(ev_norm_fork, st));
        P5_CUDA(cudaStreamWaitEvent(s, ev_norm_fork, 0));
    }
    int64_t nb = n / 65536;
    nb = nb < 16 ? 16 : (nb > 512 ? 512 : nb);
    if (early_np + nb > EARLY_SLOTS - 256) return;     // out of slots: coverage stays incomplete and grad_norm() takes the full pass
    sumsq_partial(G + off, n, norm_early + early_np, (int)nb, s);
    early_np += (int)nb;
    early_cov += n;
    if (off < early_lo) early_lo = off;
}

// ------------------------------------------------------------------------------------------------------------
// optimiser
// ------------------------------------------------------------------------------------------------------------
// the reference's optimizer groups (SingleRunner.py:186-205): `no_decay = ["bias", "LayerNorm.weight"]` is matched by
// substring against the parameter names; T5 names its norms "layer_norm" (no match) but "bias" matches the two
// `relative_attention_bias.weight` tables, which therefore take weight_decay = 0.  Ranges relative to `base`.
NoDecay Engine::no_decay(int64_t base) const {
    NoDecay nd;
    const int64_t n = (int64_t)cfg.rel_buckets * H;
    if (NE > 0) { nd.lo0 = off_enc_rel - base; nd.hi0 = nd.lo0 + n; }
    if (ND > 0) { nd.lo1 = off_dec_rel - base; nd.hi1 = nd.lo1 + n; }
    return nd;
}

void Engine::grad_norm() {
    join_optimizer();
    if (early_norm_ready) {
        // [early_lo, n_flat) was summed under the backward; the head of the buffer (embeddings + encoder block 0) is left
        P5_CUDA(cudaStreamWaitEvent(st, ev_norm_join, 0));
        int np = early_np;
        if (early_lo > 0) { sumsq_partial(G, early_lo, norm_early + np, 256, st); np += 256; }
        sumsq_final(norm_early, np, norm_out, st);
        early_norm_ready = false;
    } else {
        sumsq_norm(G, n_flat, norm_partial, norm_out, st);
    }
    norm_valid = true;
}
void Engine::zero_grad() {
    join_optimizer();
    P5_CUDA(cudaMemsetAsync(G, 0, n_flat * sizeof(float), st));
    invalidate_norm();
}
void Engine::wait_opt(int r) {
    if (opt_pending) P5_CUDA(cudaStreamWaitEvent(st, ev_opt[r], 0));
}
void Engine::join_optimizer() {
    if (!opt_pending) return;
    for (auto e : ev_opt) P5_CUDA(cudaStreamWaitEvent(st, e, 0));
    opt_pending = false;
}
void Engine::adamw_async(float lr, float b1, float b2, float eps, float wd, int step, float clip) {
    join_optimizer();
    if (clip > 0.f && !norm_valid) grad_norm();
    if (!st_opt) {
        P5_CUDA(cudaStreamCreateWithFlags(&st_opt, cudaStreamNonBlocking));
        P5_CUDA(cudaEventCreateWithFlags(&ev_opt_start, cudaEventDisableTiming));
        // ranges in the order the forward first touches them: [embeddings] [encoder layer l (+ final norm)] [decoder layer l]
        std::vector<int64_t> cuts;
        cuts.push_back(0);
        for (int l = 0; l < NE; ++l) cuts.push_back(enc[l].sa.q);
        for (int l = 0; l < ND; ++l) cuts.push_back(dec[l].sa.q);
        cuts.push_back(n_flat);
        for (size_t i = 0; i + 1 < cuts.size(); ++i) {
            P5_CHECK(cuts[i + 1] > cuts[i], "optimizer ranges must be increasing");
            opt_ranges.push_back({cuts[i], cuts[i + 1] - cuts[i]});
            cudaEvent_t e;
            P5_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev_opt.push_back(e);
        }
    }
    P5_CUDA(cudaEventRecord(ev_opt_start, st));
    P5_CUDA(cudaStreamWaitEvent(st_opt, ev_opt_start, 0));
    for (size_t r = 0; r < opt_ranges.size(); ++r) {
        const int64_t o = opt_ranges[r].first, n = opt_ranges[r].second;
        adamw_flat(P + o, G + o, M1 + o, V2 + o, P16 ? P16 + o : nullptr, n, lr, b1, b2, eps, wd, step, clip,
                   clip > 0.f ? norm_out : nullptr, 1.f, st_opt, true, no_decay(o));
        P5_CUDA(cudaEventRecord(ev_opt[r], st_opt));
    }
    opt_pending = true;
    shadow_stale = false;
    invalidate_norm();
}
void Engine::adamw(float lr, float b1, float b2, float eps, float wd, int step, float clip, bool zero_grad_after) {
    join_optimizer();
    if (clip > 0.f && !norm_valid) grad_norm();
    adamw_flat(P, G, M1, V2, P16, n_flat, lr, b1, b2, eps, wd, step, clip, clip > 0.f ? norm_out : nullptr, 1.f, st,
               zero_grad_after, no_decay(0));
    shadow_stale = false;
    if (zero_grad_after) invalidate_norm();
}

}  // namespace p5
