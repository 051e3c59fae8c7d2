/* p5_b200.h — C-ABI of libp5b200.so, the B200 (sm_100a) engine for the OpenP5 T5 hot path.
 *
 * The reference (agiresearch/OpenP5) has no FFI layer: its hot path is the Python object protocol the runner
 * uses on `model` / `optimizer`.  Each entry point below replaces one of those call sites;
 * the citation after "replaces:" is the reference file:line (relative to the reference tree, `HF:` = the
 * transformers package the reference imports).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; p5_last_error() returns the message of the
 *     last failing call on this thread.
 *   - all tensor pointers are DEVICE pointers unless the name ends in `_host`; they are borrowed for the
 *     duration of the call.  Work is enqueued on the cudaStream_t given at p5_create (stream-ordered, no
 *     internal threads).  One handle per process/GPU; a handle is not thread-safe.
 *   - token ids / masks are int32 on this boundary (the Python wrapper converts the collator's int64).
 *   - dtype codes: 0 = fp32, 1 = bf16.   major codes: 0 = K-major, 1 = MN-major.
 */
#ifndef P5_B200_H
#define P5_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct p5_engine* p5_handle;
typedef struct p5_trie_s* p5_trie;

/* Model / run configuration.  Field meaning follows HF T5Config as used by the reference
 * (src/src_t5/main.py:176-184, HF:models/t5/configuration_t5.py). */
typedef struct {
    int32_t vocab_size;        /* after resize_token_embeddings(len(tokenizer)), main.py:193 */
    int32_t d_model;
    int32_t d_kv;              /* must be 64 */
    int32_t d_ff;
    int32_t num_layers;        /* encoder blocks */
    int32_t num_decoder_layers;
    int32_t num_heads;
    int32_t rel_buckets;       /* relative_attention_num_buckets (32) */
    int32_t rel_max_distance;  /* relative_attention_max_distance (128) */
    int32_t ffn_gated_gelu;    /* 0 = DenseReluDense (t5-small/base/large), 1 = gated-GELU (v1.1) */
    int32_t whole_word_rows;   /* 512, P5_T5.py:64-66 */
    float   dropout;           /* dropout_rate */
    float   ln_eps;            /* layer_norm_epsilon */
    int32_t precision;         /* 0 = fp32 parity path (SIMT fp32), 1 = bf16 tensor-core path, 2 = bf16x3 tensor-core parity
                                  path (fp32 storage; every linear layer is the tcgen05 kernel over hi/lo-split operands) */
    int32_t max_batch;         /* workspace sizing: largest B the handle will see */
    int32_t max_enc_len;       /* largest Le (<= 512, Collator.py:13) */
    int32_t max_dec_len;       /* largest Ld for training / max_length for generate */
    int32_t max_beams;         /* largest num_beams for p5_generate (0 = no generate workspace) */
    int32_t use_mn_major;      /* 1: dgrad/wgrad read forward tensors in place via MN-major UMMA descriptors */
    int32_t reserved[7];
} P5Config;

const char* p5_last_error(void);
int p5_version(void);

/* ---- lifecycle -------------------------------------------------------------------------------------- */
/* replaces: P5_T5.from_pretrained(...).to(device)  (main.py:184)  — parameters are created zeroed; the host
 * fills them through p5_param_info pointers (state_dict protocol). */
int p5_create(const P5Config* cfg, int device, void* cuda_stream, p5_handle* out);
int p5_destroy(p5_handle h);

/* replaces: model.named_parameters() / state_dict() (SingleRunner.py:196; DistributedRunner.py:155,169,193).
 * Names are the HF T5 keys ("shared.weight", "encoder.block.0.layer.0.SelfAttention.q.weight", ...). */
int p5_param_count(p5_handle h, int* n);
int p5_param_info(p5_handle h, int i, const char** name, int* ndim, int64_t shape[2], float** data, float** grad);
/* replaces: model.resize_token_embeddings(len(tokenizer)) (main.py:193; HF:modeling_utils.py resize_token_embeddings):
 * the engine is rebuilt for the new vocabulary size in place (same handle); every tensor other than shared.weight is
 * kept, shared.weight keeps its first min(old, new) rows (and Adam moments), new rows are drawn from N(0, 1) as HF's
 * T5 initialiser does (the reference then overwrites them in utils/initialization.py:27).  All pointers returned by
 * p5_param_info are invalidated: query them again. */
int p5_resize_vocab(p5_handle h, int new_vocab);
/* call after writing parameter data from the host side (refreshes the bf16 GEMM shadows) */
int p5_params_changed(p5_handle h);

/* ---- training step ---------------------------------------------------------------------------------- */
/* replaces: model(input_ids=, whole_word_ids=, attention_mask=, labels=) (DistributedRunner.py:63-70; P5_T5.py:275-386).
 * loss_tok [B*Ld] = un-reduced per-token CE (P5_T5.py:368-369); logits_or_null [B*Ld, vocab] fp32 if non-null.
 * training != 0 enables dropout with the given seed. */
int p5_forward(p5_handle h, const int32_t* input_ids, const int32_t* attention_mask, const int32_t* whole_word_ids,
               const int32_t* labels, int B, int Le, int Ld, float* loss_tok, float* logits_or_null, int training,
               uint64_t seed);
/* Optional, applies to the NEXT p5_forward / p5_train_fwd_bwd only: number of valid (right-padded) encoder tokens of
 * each sequence = attention_mask.sum(1) of the collator batch (Collator.py:8-34 pads to the longest; HF computes the
 * padded positions and masks them).  With the lengths known on the host the engine drops the padding: every
 * token-wise encoder kernel runs on sum(lens) rows instead of B*Le.  Results at valid positions are unchanged. */
int p5_set_enc_lengths(p5_handle h, const int32_t* lens_host, int B);
/* replaces: loss.backward() (DistributedRunner.py:80) given dL/dloss_tok [B*Ld]; accumulates into the grad buffers */
int p5_backward(p5_handle h, const float* dloss_tok);
/* fused form of DistributedRunner.py:72-80: loss = mean_b( sum_t loss_tok*m / max(sum_t m,1) ), m = (labels_mask != 0);
 * runs forward + backward with that reduction; loss_out (device, 1 float) receives the scalar loss. */
int p5_train_fwd_bwd(p5_handle h, const int32_t* input_ids, const int32_t* attention_mask,
                     const int32_t* whole_word_ids, const int32_t* labels, const int32_t* labels_mask, int B, int Le,
                     int Ld, float* loss_out, uint64_t seed);
/* replaces: torch.nn.utils.clip_grad_norm_ (DistributedRunner.py:81): out (device, 1 float) = global L2 norm */
int p5_grad_norm(p5_handle h, float* out);
int p5_grad_scale(p5_handle h, float s);
/* replaces: model.zero_grad() (DistributedRunner.py:87) */
int p5_zero_grad(p5_handle h);
/* replaces: transformers(4.26).AdamW.step (SingleRunner.py:214): m,v update, eps outside bias correction,
 * decoupled decay after the update.  clip > 0 folds clip_grad_norm_(clip) into the same pass using the norm
 * computed on device (no host sync).  step is 1-based.  Parameter groups as the reference builds them
 * (SingleRunner.py:186-205): names containing "bias" — the two relative_attention_bias tables — take weight_decay 0. */
int p5_adamw_step(p5_handle h, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  float clip);
/* replaces: optimizer.step() immediately followed by model.zero_grad() (DistributedRunner.py:85-87) in ONE pass over
 * the flat buffers: identical update to p5_adamw_step, the gradient is cleared while it is read (saves the separate
 * 4 B/parameter memset of p5_zero_grad). */
int p5_adamw_step_zero_grad(p5_handle h, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                            float clip);
/* Same update and gradient clearing as p5_adamw_step_zero_grad, but issued range by range (embeddings, encoder layer
 * 0.., decoder layer 0..) on an engine-owned side stream: the HBM-bound optimiser pass overlaps the tensor-bound
 * forward of the NEXT step, which waits for each range right before it first reads it (the training loop of
 * DistributedRunner.py:59-87 never touches the parameters between optimizer.step() and the next forward).  Every
 * other entry point of this library joins the pending update first; code that reads the parameter buffers directly
 * (torch views) must call p5_optimizer_join before doing so. */
int p5_adamw_step_zero_grad_async(p5_handle h, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  int step, float clip);
int p5_optimizer_join(p5_handle h);

/* replaces: utils/evaluate.py:37-92 (rel_results + hit@k / ndcg@k) and the per-batch accumulation of
 * DistributedRunner.py:376-393, on token-id paths and on the device: seqs [B*K, T] / scores [B*K] as returned by
 * p5_generate, gold [B, Tg] the tokenised target items, ks_dev [n_k] the cut-offs (device).  Pad (0) and eos (1) are
 * ignored in the comparison (== batch_decode(skip_special_tokens=True) as a key).  out_sums (device, 2*n_k floats,
 * NOT cleared): [hit@ks[0..n_k) | ndcg@ks[0..n_k)] summed over the B users; the caller divides by the (all-reduced)
 * user count as the reference does. */
int p5_eval_metrics(p5_handle h, const int32_t* seqs, const float* scores, int B, int K, int T, const int32_t* gold, int Tg,
                    const int32_t* ks_dev, int n_k, float* out_sums);

/* Filtered variant — replaces: utils/evaluate.py:6-35 (rel_results_filtered) as driven by DistributedRunner.py:204-265:
 * the R (= generate_num + max_positive) returned rows of a user are ordered by score, rows that equal one of the user's
 * positive (already interacted) items pos [B, Pmax, Tp] (npos [B] valid rows per user; device pointers) are skipped, the
 * first k_cut remaining rows form the relevance list the hit@k / ndcg@k sums are taken over. */
int p5_eval_metrics_filtered(p5_handle h, const int32_t* seqs, const float* scores, int B, int R, int T, const int32_t* gold,
                             int Tg, const int32_t* pos, const int32_t* npos, int Pmax, int Tp, const int32_t* ks_dev, int n_k,
                             int k_cut, float* out_sums);

/* Optimiser state of parameter i (same index / shape as p5_param_info): device pointers to the fp32 Adam moments
 * exp_avg / exp_avg_sq.  With the step count (kept by the host) this is what a resumable checkpoint adds to the
 * reference's plain state_dict (DistributedRunner.py:155,169 saves weights only). */
int p5_opt_state_info(p5_handle h, int i, float** exp_avg, float** exp_avg_sq);

/* ---- data-parallel gradient exchange (the DDP all-reduce the reference constructs, DistributedRunner.py:26) ---- */
int p5_comm_unique_id(void* id128_host);                      /* 128-byte ncclUniqueId */
int p5_comm_init(p5_handle h, const void* id128_host, int rank, int world);
int p5_allreduce_grads(p5_handle h);                          /* mean over ranks */

/* ---- constrained beam search ------------------------------------------------------------------------ */
/* replaces: gt.Trie([...]) + prefix_allowed_tokens_fn (utils/generation_trie.py:7-97; DistributedRunner.py:344-351).
 * paths_host: concatenated token paths (each starts with decoder_start 0), offsets_host[n_paths+1]. */
int p5_trie_build(p5_handle h, const int32_t* paths_host, const int64_t* offsets_host, int n_paths, p5_trie* out);
int p5_trie_free(p5_trie t);
int p5_trie_stats(p5_trie t, int* n_nodes, int* n_edges, int* max_depth);
/* device lookup used by the tests: for each prefix (host arrays) return allowed next tokens like Trie.get */
int p5_trie_get(p5_trie t, const int32_t* prefix_host, int prefix_len, int32_t* out_tokens_host, int cap, int* n_out);

/* replaces: model.generate(num_beams=K, num_return_sequences=R, max_length=T, prefix_allowed_tokens_fn=...)
 * (DistributedRunner.py:361-371; HF:generation/utils.py:3076-3400).  seqs [B*R, max_len] int32 (pad 0),
 * scores [B*R] fp32 = sum log-prob / generated_len^length_penalty, rows per user sorted by score desc.
 * out_len_host receives the cropped output length (prompt + longest generated). */
int p5_generate(p5_handle h, const int32_t* input_ids, const int32_t* attention_mask, const int32_t* whole_word_ids,
                int B, int Le, p5_trie trie, int num_beams, int num_return, int max_len, float length_penalty,
                int32_t* seqs, float* scores, int* out_len_host);

/* Device time (CUDA events on the launch stream), algorithmic bytes (per position: the decoder-block weights and the tied
 * LM head once, every user's cross K|V once, all bf16) and number of positions of the LAST persistent decode launch of
 * p5_generate (bench.py eval roofline).  Synchronises on that launch. */
int p5_decode_last_launch(float* ms_host, double* bytes_host, int* steps_host);
/* per-phase nanoseconds of that launch as CTA 0 saw them (each phase incl. its grid barrier): out32_host[0..15] = positions
 * run by the weight-streaming GEMM variant (<= 32 live rows), [16..31] = the others; index: 0 qkv, 1 self-attention, 2 o,
 * 3 cross-q, 4 cross-attention, 5 cross-o, 6 wi, 7 wo, 8 LM head, 9 per-user beam phase, 10 live-list build */
int p5_decode_phase_ns(uint64_t* out32_host);

/* ---- collaborative item indexing: the quadratic part ------------------------------------ */
/* replaces: the co-occurrence matrix loop of utils/indexing.py:163-180 (generate_collaborative_id): items [sum len] are
 * the training-prefix item ids (0 .. n_items-1) of every user back to back, offsets [n_users+1]; adj [n_items, n_items]
 * (fp32 when f64 == 0, else fp64; cleared by the call) receives adj[a][b] += 1, adj[b][a] += 1 for every pair of
 * positions i < j of a user.  All pointers are device pointers. */
int p5_cooccurrence(const int32_t* items, const int64_t* offsets, int n_users, int n_items, int f64, void* adj, void* cuda_stream);
/* replaces: the sub-matrix loop of utils/indexing.py:220-231: out[i][j] = adj[idx[i]][idx[j]] (i != j), 0 on the diagonal */
int p5_submatrix(const void* adj, int n_items, int f64, const int32_t* idx, int m, void* out, void* cuda_stream);

/* ---- op-level hooks (unit tests / micro-benchmarks of individual kernels) ------------------------------ */
typedef struct {
    int32_t backend;            /* 0 = SIMT fp32-accumulate kernel, 1 = tcgen05 kernel, 2 = auto */
    int32_t M, N, K, nb1, nb2;
    const void* A; int32_t a_dtype, a_major; int64_t lda, a_bs1, a_bs2;
    const void* B; int32_t b_dtype, b_major; int64_t ldb, b_bs1, b_bs2;
    void* C; int32_t c_dtype; int32_t pad0; int64_t ldc, c_bs1, c_bs2;
    float alpha; int32_t flags;
    const void* aux; int32_t aux_dtype; int32_t pad1;
    const float* resid;
    uint64_t seed; uint32_t site; float drop_p;
    int32_t force_block_n; int32_t pad2;
} P5GemmDesc;
int p5_op_gemm(const P5GemmDesc* d, void* cuda_stream);
int p5_launch_count(void);   /* kernels launched by this library since load */
/* tile width (64 / 128 / 192 / 256) the tcgen05 GEMM picks for an M x N output (x batches) on `sms` SMs: host arithmetic
   only, no device needed (tests/test_cabi_cpu.py); 0 on invalid arguments */
int p5_gemm_tile_width(int M, int N, int batches, int sms);
/* per-launch CUDA-event timing of the tcgen05 GEMM (bench.py roofline leg): enable, run steps, read a JSON summary
 * {"bn256": {"launches", "ms", "flops"}, "bn128": ..., "bn64": ...} (algorithmic FLOPs = 2*M*N*K per launch) */
int p5_prof_enable(int on);
int p5_prof_summary(char* json_out, int cap);
/* the same profiled launches as a per-shape text table (M N K batches epilogue-flags majors tile : launches, us, TFLOP/s) */
int p5_prof_shapes(char* text_out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* P5_B200_H */
