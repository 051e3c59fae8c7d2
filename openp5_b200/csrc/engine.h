// engine.h — the P5 T5 engine: flat parameter storage, activation workspace and the forward / backward /
// optimiser / generate orchestration over the kernels in gemm_tc.cu, gemm_simt.cu, kernels.cu, attention.cu,
// optim.cu and beam.cu.   Path replaced: ref src/src_t5/model/P5_T5.py:74-204,275-386 + HF modeling_t5.py blocks.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "../../include/p5_b200.h"
#include <string>
#include <vector>
#include <utility>

namespace p5 {
extern int g_launches;
void gemm_tc_force_block_n(int bn);
// tcgen05 when the problem qualifies (bf16 operands, aligned), SIMT otherwise
void gemm_auto(const GemmProblem& p, cudaStream_t stream, bool allow_mn_major);
// gemm_x3.cu: fp32 operands split into bf16 hi/lo pairs, three products in ONE tcgen05 GEMM (K-concatenated operands)
bool gemm_x3_supported(const GemmProblem& p);
void gemm_bf16x3(const GemmProblem& p, cudaStream_t st);
void gemm_x3_release();

struct ParamInfo {
    std::string name;
    int ndim;
    int64_t shape[2];
    int64_t off;    // offset in the flat fp32 buffers
    int64_t numel;
};

struct AttnOff { int64_t q, k, v, o; };
struct FfnOff { int64_t wi, wi1, wo; };
struct EncLayerOff { AttnOff sa; int64_t ln0; FfnOff ff; int64_t ln1; };
struct DecLayerOff { AttnOff sa; int64_t ln0; AttnOff ca; int64_t ln1; FfnOff ff; int64_t ln2; };

// dropout sites: site id = kind * 64 + layer
enum SiteKind : uint32_t {
    S_EMB_E = 0, S_EMB_D, S_ENC_P, S_ENC_O, S_ENC_ACT, S_ENC_WO, S_ENC_FINAL,
    S_DEC_SP, S_DEC_SO, S_DEC_CP, S_DEC_CO, S_DEC_ACT, S_DEC_WO, S_DEC_FINAL
};

struct Trie;  // beam.cu

struct Engine {
    P5Config cfg;
    int device = 0;
    cudaStream_t st = nullptr;
    int dt = DT_BF16;          // activation / GEMM operand dtype
    bool x3 = false;           // precision 2: fp32 storage, every linear layer through the tcgen05 kernel as a bf16x3 product
    bool mn = true;            // MN-major operands allowed on the tcgen05 path
    int d = 0, A = 0, H = 0, ff = 0, V = 0, Vpad = 0, NE = 0, ND = 0;
    bool gated = false;
    float p_drop = 0.f;

    // ---- parameters (one flat allocation each) ----
    std::vector<ParamInfo> params;
    int64_t n_flat = 0;
    float *P = nullptr, *G = nullptr, *M1 = nullptr, *V2 = nullptr;
    bf16* P16 = nullptr;
    int64_t off_shared = 0, off_ww = 0, off_enc_rel = 0, off_dec_rel = 0, off_enc_final = 0, off_dec_final = 0;
    std::vector<EncLayerOff> enc;
    std::vector<DecLayerOff> dec;
    bool shadow_stale = true;

    // ---- workspace ----
    std::vector<void*> allocs;
    int Bm = 0, Lem = 0, Ldm = 0;      // capacity
    int B = 0, Le = 0, Ld = 0;          // current step geometry (Le padded to a multiple of 8)
    int Le_user = 0;
    int64_t Me = 0, Md = 0;
    // packed (padding-free) encoder token layout: active when the caller supplied the encoder lengths for this step
    // (p5_set_enc_lengths).  All token-wise encoder work then runs on Mt = sum(lens) rows instead of B*Le.
    bool packed = false;
    int64_t Mt = 0;                      // encoder token rows actually processed (Me when not packed)
    int64_t Mt_true = 0;                 // packed: real tokens (Mt = Mt_true rounded up to 512 with inert filler rows)
    std::vector<int> lens_h, offs_h;     // host copies (offs has B+1 entries)
    std::vector<int> pending_lens;       // set by p5_set_enc_lengths, consumed by the next forward
    int *lens_d = nullptr, *offs_d = nullptr, *ids_p = nullptr, *ww_p = nullptr;
    void *qkv_pad = nullptr, *ctx_pad = nullptr, *dqkv_pad = nullptr;   // padded scratch around the attention backward
    float* f_qkv_pad = nullptr;
    bool training = false;
    uint64_t seed = 0;
    bool have_fwd = false;

    int *ids_e = nullptr, *mask_e = nullptr, *ww_e = nullptr, *labels = nullptr, *dec_ids = nullptr, *lmask = nullptr;
    // encoder saved activations
    std::vector<float*> xe;             // [2*NE+1] residual stream (x_in[l] = xe[2l], x_mid[l] = xe[2l+1], out = xe[2NE])
    std::vector<float*> rstd_e;         // [2*NE+1]
    std::vector<void*> ne;              // [2*NE] normed inputs (dt)
    void* enc_out = nullptr;            // dt [Me, d]
    std::vector<void*> qkv_e, ctx_e, h_e, z_e, P_e;
    bool next_gemm_indep = false; // one-shot: the next gemm() neither consumes nor clobbers the kernel launched before it
    std::vector<bool> p_fbwd;     // per encoder layer: lse_e holds lse2, the fused attention backward applies
    std::vector<bool> p_unnorm;   // per encoder layer: P_e holds un-normalised probabilities (fused attention forward)
    std::vector<float*> lse_e;
    float* S_scr = nullptr;             // fp32 [B,H,Le,Le] scores / dP
    void* Pd_scr = nullptr;             // dt dropped probabilities (forward + regenerated in backward)
    void* dS_scr = nullptr;             // dt
    // decoder saved activations
    std::vector<float*> yd;             // [3*ND+1]
    std::vector<float*> rstd_d;         // [3*ND+1]
    std::vector<void*> nd;              // [3*ND]
    void* dec_out = nullptr;
    std::vector<void*> sqkv, sctx, cq, ckv, cctx, h_d, z_d;
    // cross K|V of ALL decoder layers live in one [rows, ND * 2A] matrix (ckv[l] = column block l, row stride ckv_ld):
    // the ND projections of enc_out are ONE batched GEMM (no per-layer wave quantisation), and so are their weight /
    // input gradients (g_ckv_all holds dK|dV of every layer until the decoder backward is done)
    void *ckv_all = nullptr, *g_ckv_all = nullptr;
    int64_t ckv_ld = 0, dec_layer_stride = 0;
    bool batched_ckv = false;
    // Decoder weight gradients of ALL layers as one batched GEMM per weight type (batch dimension = layer): the decoder has
    // only B * Ld (= 512) tokens, so its 6 wgrads per layer are launch- and pipeline-fill-bound (~13 us each for < 1 us of
    // tensor work, 72 launches per step); none of them is on the critical path of the backward.  Their operands (the
    // incoming gradient dY of every linear and its saved input X) live in per-layer slabs with a constant stride.
    bool batched_dwg = false;
    void *gdw_all = nullptr, *gdc_all = nullptr, *gds_all = nullptr, *gff_all = nullptr, *gcq_all = nullptr, *gsq_all = nullptr;
    void *nd_all = nullptr, *h_d_all = nullptr, *sctx_all = nullptr, *cctx_all = nullptr;
    void decoder_wgrads_batched();
    void project_cross_kv_all(int64_t rows);
    // model.resize_token_embeddings(n) (main.py:193): rebuilds the flat buffers for the new vocabulary, keeps every other
    // tensor and the first min(V, n) embedding rows (and their Adam moments); new rows ~ N(0, 1) like HF's T5 init
    Engine* resized(int new_vocab);
    std::vector<float*> slse, clse;
    // head
    float *logits = nullptr, *lse_ce = nullptr, *loss_tok = nullptr, *dloss = nullptr, *loss_scalar = nullptr;
    void* dlogits = nullptr;
    // backward scratch
    float *dx_a = nullptr, *dx_b = nullptr, *d_encout = nullptr;
    void *g_ff = nullptr, *g_d = nullptr, *g_d2 = nullptr, *g_qkv = nullptr, *g_ctx = nullptr, *g_ckv = nullptr;
    float *f_qkv = nullptr, *f_ckv = nullptr;   // fp32 scratch for SIMT attention backward
    // relative position bias tables
    float *bias_enc = nullptr, *bias_dec = nullptr, *dbias_enc = nullptr, *dbias_dec = nullptr;
    int *lut_enc = nullptr, *lut_dec = nullptr;
    int lut_enc_L = -1, lut_dec_L = -1;
    // optimiser scratch
    float *norm_partial = nullptr, *norm_out = nullptr;
    bool norm_valid = false;
    // Early gradient norm (fused train step, p5_train_fwd_bwd): the sum of squares of a gradient range is taken on a side
    // stream as soon as the backward has finished the range (after its all-reduce when data parallel), under the rest of
    // the backward; grad_norm() then only has the last range (embeddings + encoder block 0) left instead of a pass over
    // the whole 0.9 GB buffer between the backward and AdamW.
    bool early_norm = false;           // set around backward() by the fused train step
    bool early_norm_ready = false;     // the partials cover [early_lo, n_flat) of the CURRENT gradient buffer
    int64_t early_lo = 0, early_cov = 0;
    int early_np = 0;                  // block partials written so far
    static constexpr int EARLY_SLOTS = 4096;
    float* norm_early = nullptr;       // [EARLY_SLOTS]
    cudaStream_t st_norm = nullptr;
    cudaEvent_t ev_norm_fork = nullptr, ev_norm_join = nullptr;
    void range_final(int64_t off, int64_t n);   // the gradient range is complete: all-reduce (data parallel) + early sum of squares
    void invalidate_norm() { norm_valid = false; early_norm_ready = false; }
    // comm
    void* nccl_comm = nullptr;
    int world = 1, rank = 0;
    // ---- asynchronous optimiser (p5_adamw_step_zero_grad_async): AdamW runs range by range (embeddings, encoder layer
    // 0..NE-1, decoder layer 0..ND-1) on a side stream; the NEXT forward waits for a range just before it first reads
    // it, every other entry point joins all ranges first.  HBM-bound AdamW then overlaps the tensor-bound forward.
    cudaStream_t st_opt = nullptr;
    cudaEvent_t ev_opt_start = nullptr;
    std::vector<cudaEvent_t> ev_opt;
    std::vector<std::pair<int64_t, int64_t>> opt_ranges;   // (offset, count) in forward-use order
    bool opt_pending = false;
    void adamw_async(float lr, float b1, float b2, float eps, float wd, int step, float clip);
    void wait_opt(int range);      // st waits for one range (no-op when nothing is pending)
    void join_optimizer();         // st waits for every range
    bool overlap_comm = false;   // set for the fused train step: ranges are all-reduced as backward completes them

    // ---- generate workspace (beam.cu) ----
    struct GenWs* gen = nullptr;

    Engine(const P5Config& c, int device, cudaStream_t st);
    ~Engine();

    void* dalloc(size_t bytes);
    template <typename T> T* dalloc_t(size_t n) { return (T*)dalloc(n * sizeof(T)); }
    const void* W(int64_t off) const { return dt == DT_BF16 ? (const void*)(P16 + off) : (const void*)(P + off); }
    size_t esz() const { return dtype_size(dt); }
    DropCfg drop(uint32_t kind, int layer) const;
    void refresh_shadow();
    void set_geometry(int B, int Le_user, int Ld);
    void load_inputs(const int32_t* ids, const int32_t* mask, const int32_t* ww, const int32_t* labels);
    void apply_lengths();   // consumes pending_lens -> packed / Mt / offs / lens
    void build_bias(bool encoder, int L);

    // GEMM helpers (C = A * B^T forms; see engine.cu)
    void gemm(GemmProblem& p);
    void linear_fwd(const void* X, int64_t ldx, int64_t w_off, int N, int K, int M, void* Y, int y_dtype, int64_t ldy,
                    int flags, float alpha, const void* aux, const float* resid, DropCfg dc);
    void linear_dgrad(const void* dY, int64_t lddy, int64_t w_off, int N, int K, int M, void* dX, int dx_dtype,
                      int64_t lddx, int flags, float alpha, const void* aux, bool accum_f32);
    void linear_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t w_off, int N, int K, int M,
                      float alpha, bool after_own_dgrad = false);

    void encoder_forward();
    void decoder_forward();
    void head_forward();
    void forward(const int32_t* ids, const int32_t* mask, const int32_t* ww, const int32_t* labels, int B, int Le,
                 int Ld, bool training, uint64_t seed);
    void backward();   // consumes this->dloss
    void ensure_attn_scratch();   // L^2-sized buffers of the materialised attention paths, allocated on first use
    void enc_attention_fwd(int l);
    void enc_attention_bwd(int l, const void* dctx, void* dqkv);
    void ffn_fwd(const void* n, int64_t M, const FfnOff& w, void* z, void* h, const float* x_resid, float* x_out,
                 uint32_t kind_act, uint32_t kind_wo, int layer);
    void ffn_bwd(const float* dx_out, int64_t M, const FfnOff& w, const void* n, const void* z, const void* h,
                 void* dn_out, uint32_t kind_act, uint32_t kind_wo, int layer, const void* gd = nullptr, void* gff = nullptr,
                 bool skip_wgrad = false);

    NoDecay no_decay(int64_t base) const;
    void grad_norm();
    void adamw(float lr, float b1, float b2, float eps, float wd, int step, float clip, bool zero_grad_after = false);
    void zero_grad();
};

}  // namespace p5
