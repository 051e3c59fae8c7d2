// kernels.cuh — launchers of the HBM-bound (non-GEMM) kernels of the T5 hot path.  Implementations: kernels.cu,
// attention.cu, optim.cu, beam.cu.  All launchers enqueue on `st` and throw P5Error on launch failure.
#pragma once
#include "common.cuh"

namespace p5 {

struct DropCfg {
    uint64_t seed = 0;
    uint32_t site = 0;
    uint32_t thr = 0;       // 0 = dropout off
    float inv_keep = 1.f;
};

// ---- embeddings (ref P5_T5.py:94-100,125; HF T5Stack embed + dropout) --------------------------------------
void embed_fwd(const float* E, const float* Wword, const int* ids, const int* ww, float* x, int M, int d, int vocab,
               int ww_rows, DropCfg drop, cudaStream_t st);
void embed_bwd(const float* dx, const int* ids, const int* ww, float* dE, float* dWword, int M, int d, int vocab,
               int ww_rows, DropCfg drop, cudaStream_t st);

// ---- RMSNorm (HF:modeling_t5.py:55-70) -----------------------------------------------------------------------
// n = drop( w * x * rsqrt(mean(x^2)+eps) ), rstd saved for backward (rstd may be null in inference)
void rmsnorm_fwd(const float* x, const float* w, void* n, int n_dtype, float* rstd, int M, int d, float eps,
                 DropCfg drop, cudaStream_t st);
// dx = (dres ? dres : 0) + d/dx[ rmsnorm ](mask(dn));  dw += sum_rows(mask(dn) * xhat)   (atomic)
// optional dx_cast: cast_dtype copy of dx with `cast_drop` applied = the dropout-cast the NEXT backward op needs
void rmsnorm_bwd(const void* dn, int dn_dtype, const float* x, const float* rstd, const float* w, const float* dres,
                 float* dx, float* dw, int M, int d, DropCfg drop, cudaStream_t st, void* dx_cast = nullptr,
                 int cast_dtype = DT_BF16, DropCfg cast_drop = DropCfg());

// out = drop(in) cast to out_dtype (backward of the `x + drop(y)` sites: the mask is regenerated, never stored)
void drop_cast(const float* in, void* out, int out_dtype, int64_t n, DropCfg drop, cudaStream_t st);
void cast_f32_to(const float* in, void* out, int out_dtype, int64_t n, cudaStream_t st);
void cast_to_f32(const void* in, int in_dtype, float* out, int64_t n, cudaStream_t st);
// out[r, c] (dtype) = in_f32[r, c] for a [rows, cols] block with leading dims
void cast_block_f32_to(const float* in, int64_t ld_in, void* out, int out_dtype, int64_t ld_out, int rows, int cols,
                       cudaStream_t st);
void add_f32(float* dst, const float* src, int64_t n, cudaStream_t st);

// packed (variable-length, padding removed) <-> padded [B, L] row layouts; offs[b] = first packed row, lens[b] = tokens
void pack_rows(const void* src_padded, void* dst_packed, const int* offs, const int* lens, int B, int L, int64_t row_bytes,
               cudaStream_t st);
void unpack_rows(const void* src_packed, void* dst_padded, const int* offs, const int* lens, int B, int L, int64_t row_bytes,
                 cudaStream_t st);
void pack_ints(const int* src_padded, int* dst_packed, const int* offs, const int* lens, int B, int L, cudaStream_t st);

// gated-GELU (HF:modeling_t5.py:106-132): z = [z0 | z1] per row (ld = 2*ff); h = drop(gelu_new(z0) * z1)
void gated_gelu_fwd(const void* z, void* h, int dtype, int M, int ff, DropCfg drop, cudaStream_t st);
void gated_gelu_bwd(const void* z, const void* dh, void* dz, int dtype, int M, int ff, DropCfg drop, cudaStream_t st);

// ---- LM head loss (ref P5_T5.py:364-369) ------------------------------------------------------------------------
void ce_fwd(const float* logits, int64_t ld, const int* labels, float* loss_tok, float* lse, int M, int V,
            cudaStream_t st);
// dlogits[m, :] = dloss[m] * (softmax(logits[m]) - onehot(label)); columns [V, Vpad) are written as zero
void ce_bwd(const float* logits, int64_t ld, const float* lse, const int* labels, const float* dloss, void* dlogits,
            int d_dtype, int M, int V, int Vpad, cudaStream_t st);
// runner loss (ref DistributedRunner.py:72-77): loss = mean_b( sum_t l*m / max(sum_t m, 1) ); also d loss / d l
void runner_loss_fwd_bwd(const float* loss_tok, const int* labels_mask, int B, int Ld, float* loss_out,
                         float* dloss_tok, cudaStream_t st);

// ---- relative position bias (HF:modeling_t5.py:189-251) ---------------------------------------------------------
// bias_rel[h, delta + (Lq-1)] = table[bucket_lut[delta + (Lq-1)], h] for delta = j - i in [-(Lq-1), Lk-1]
void relbias_build(const float* table, const int* bucket_lut, float* bias_rel, int H, int n_delta, cudaStream_t st);
void relbias_scatter_grad(const float* dbias_rel, const int* bucket_lut, float* dtable, int H, int n_delta,
                          cudaStream_t st);

// ---- attention --------------------------------------------------------------------------------------------------
// Q, K, V are strided views: element (b, pos, h, c) at ptr[b*bs + pos*ld + h*64 + c].  d_kv = 64.
struct AttnView {
    const void* ptr; int dtype; int64_t ld; int64_t bs;
};
struct AttnArgs {
    int B, H, Lq, Lk;
    AttnView q, k, v;
    const float* bias_rel;   // [H, n_delta] or null (cross-attention: zero bias); entry (j - i_pos) + bias_off
    int bias_off, n_delta;
    const int* key_mask;     // [B, Lk] (1 = keep) or null
    int causal;              // 1: key j > query i is masked (decoder self-attention); q_pos_offset shifts i
    int q_pos_offset;        // position of query row 0 (decode step with KV cache)
    const int* row_map;      // optional [B] indirection for K/V/mask batch index (beam -> user), null = identity
    const int* kv_off = nullptr;   // packed K/V: first row of batch kb (replaces kb * bs); null = padded layout
    const int* kv_len = nullptr;   // packed K/V: number of keys of batch kb (replaces Lk and the key mask)
    DropCfg drop;
};
// fused SIMT attention (fp32 math): O[b, i, h*64 + c] (ld_o) and LSE[b, h, i]
void attn_simt_fwd(const AttnArgs& a, void* O, int o_dtype, int64_t ld_o, int64_t bs_o, float* lse, cudaStream_t st);
// backward: dQ/dK/dV written as fp32 with the (ld, bs) geometry given.  When Lq > 16 (several q-blocks per (b,h))
// dK/dV are accumulated with atomics and the caller zeroes them first (attn_bwd_needs_zero); otherwise they are
// plain stores.  dbias_rel [H, n_delta] accumulated atomically if non-null.
static inline bool attn_bwd_needs_zero(int Lq) { return Lq > 8; }   // one q-block: Lq <= 8 (1 row/warp) — else zero first
void attn_simt_bwd(const AttnArgs& a, const void* O, const void* dO, int o_dtype, int64_t ld_o, int64_t bs_o,
                   const float* lse, float* dQ, int64_t ld_dq, int64_t bs_dq, float* dK, float* dV, int64_t ld_dkv,
                   int64_t bs_dkv, float* dbias_rel, cudaStream_t st);

// decoder attention of the bf16 train step on mma.sync tiles (dattn.cu): Lq <= 16, Lk <= 512, all views bf16 with
// 16-byte aligned rows.  Same semantics / dropout stream as attn_simt_*; gradients are written as bf16 in place
// (every (key row, head) slice is owned by exactly one CTA: no zero-fill, no fp32 staging).
bool dattn_supported(const AttnArgs& a);
bool dattn_infer_supported(const AttnArgs& a);   // forward without dropout / LSE: up to 32 query rows (decode-step beams)
void dattn_fwd(const AttnArgs& a, void* O, int64_t ld_o, int64_t bs_o, float* lse, cudaStream_t st);
void dattn_bwd(const AttnArgs& a, const void* dO, int64_t ld_do, int64_t bs_do, const float* lse, void* dQ, int64_t ld_dq,
               int64_t bs_dq, void* dK, void* dV, int64_t ld_dkv, int64_t bs_dkv, float* dbias_rel, cudaStream_t st);

// materialised softmax for the tensor-core attention path: S fp32 [B,H,Lq,Lk]
//   P  = softmax(S + bias + mask)             -> P_save (dtype)           (needed by backward)
//   Pd = drop(P)                              -> Pd (dtype, may alias P_save when dropout is off)
void softmax_fwd(const float* S, const float* bias_rel, const int* key_mask, void* P_save, void* Pd, int dtype, int B,
                 int H, int Lq, int Lk, int causal, DropCfg drop, cudaStream_t st);
//   dP_in = gradient wrt Pd (fp32); dS = P * (mask(dP) - rowsum(mask(dP) * P)) -> dS (dtype);  Pd regenerated
// lens (optional, packed training): rows / columns >= lens[b] hold no probabilities -> dS and Pd are written as zeros
// row_scale (optional, [B,H,Lq]): P holds un-normalised probabilities, the true P is P * row_scale[row] (fattn_fwd)
void softmax_bwd(const float* dPd, const void* P, void* dS, void* Pd_out, int dtype, float* dbias_rel, int B, int H,
                 int Lq, int Lk, DropCfg drop, cudaStream_t st, const int* lens = nullptr, const float* row_scale = nullptr);

// fused tcgen05 encoder self-attention forward (fattn.cu): qkv [B*L, 3A] bf16 -> ctx [B*L, A] bf16.  For the backward
// it saves P_save bf16 [B,H,L,L] = UN-normalised un-dropped probabilities 2^(s2 - m2) and row_scale fp32 [B,H,L] =
// 1 / row sum (P = P_save * row_scale; softmax_bwd takes the pair), and/or row_lse2 fp32 [B,H,L] for the fused backward
// (each output is skipped when its pointer is null).  Returns false when the shape is unsupported.
// packed mode (offs/lens non-null): qkv / ctx are [packed_rows, .] with sequence b at rows offs[b] .. offs[b]+lens[b];
// P_save keeps the padded [B,H,L,L] geometry (rows/cols < lens[b] written).
bool fattn_fwd(const void* qkv, int64_t ld_qkv, int A, int B, int H, int L, const float* bias_rel, const int* key_mask,
               void* P_save, float* row_scale, float* row_lse2, void* ctx, int64_t ld_ctx, DropCfg drop, cudaStream_t st,
               const int* offs = nullptr, const int* lens = nullptr, int64_t packed_rows = 0);
// fused tcgen05 backward of the same attention (fattn_bwd.cu), L <= 512: recomputes P from row_lse2 (= m2 + log2 l,
// written by fattn_fwd; P_save / row_scale may then be null), writes dQ | dK | dV as bf16 into dqkv (rows as qkv) and
// accumulates d(bias_rel).  Returns false when the shape is unsupported (caller falls back to the GEMM chain).
// fattn_bwd_supported(L): the encoder length is covered, i.e. the forward should save row_lse2 instead of P_save.
bool fattn_bwd_supported(int L);
bool fattn_bwd(const void* qkv, int64_t ld_qkv, int A, int B, int H, int L, const float* bias_rel, const int* key_mask,
               const float* row_lse2, const void* ctx, int64_t ld_ctx, const void* dctx, int64_t ld_dctx, void* dqkv,
               int64_t ld_dqkv, float* dbias_rel, DropCfg drop, cudaStream_t st, const int* offs = nullptr,
               const int* lens = nullptr, int64_t packed_rows = 0);


// ---- optimiser (optim.cu) ---------------------------------------------------------------------------------------
void sumsq_norm(const float* g, int64_t n, float* partial /*>=1024 floats*/, float* out_norm, cudaStream_t st);
// the two halves of sumsq_norm: per-block sums of squares of a range (nblocks floats), sqrt of the sum of np partials
void sumsq_partial(const float* g, int64_t n, float* partial, int nblocks, cudaStream_t st);
void sumsq_final(const float* partial, int np, float* out_norm, cudaStream_t st);
void scale_f32(float* g, int64_t n, float s, cudaStream_t st);
// element ranges [lo, hi) (relative to the pointers passed to adamw_flat) that take weight_decay = 0: the reference's
// no_decay group (SingleRunner.py:186-205, names containing "bias" -> the two relative_attention_bias tables)
struct NoDecay { int64_t lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0; };
// transformers-4.26 AdamW on the flat parameter buffer; clip_norm_ptr (device) optional
void adamw_flat(float* p, float* g, float* m, float* v, bf16* p16, int64_t n, float lr, float b1, float b2,
                float eps, float wd, int step, float clip, const float* norm_ptr, float grad_div, cudaStream_t st,
                bool zero_grad_after = false, NoDecay nd = NoDecay());

// ranking metrics of one eval batch on the device (beam.cu)
void eval_metrics(const int32_t* seqs, const float* scores, int B, int K, int T, const int32_t* gold, int Tg, const int32_t* ks_dev,
                  int n_k, float* out_sums, cudaStream_t st);

void eval_metrics_filtered(const int32_t* seqs, const float* scores, int B, int R, int T, const int32_t* gold, int Tg,
                           const int32_t* pos, const int32_t* npos, int Pmax, int Tp, const int32_t* ks_dev, int n_k, int k_cut,
                           float* out_sums, cudaStream_t st);

}  // namespace p5
