// extern "C" boundary of libp5b200.so — see include/p5_b200.h for the contract of every entry point.
#include "engine.h"
#include <string>
#include <dlfcn.h>

using namespace p5;

static thread_local std::string g_last_error;
namespace p5 { int g_launches = 0; }

#define P5_API_BEGIN try {
#define P5_API_END                                            \
    }                                                         \
    catch (const P5Error& e) {                                \
        g_last_error = e.what();                              \
        return e.code ? e.code : 1;                           \
    }                                                         \
    catch (const std::exception& e) {                         \
        g_last_error = std::string("exception: ") + e.what(); \
        return 99;                                            \
    }                                                         \
    return 0;

// the handle is a box around the engine pointer: p5_resize_vocab replaces the engine behind an unchanged handle
struct p5_engine { Engine* e; };
static Engine* E(p5_handle h) {
    P5_CHECK(h != nullptr && h->e != nullptr, "null engine handle");
    return h->e;
}

namespace p5 {
void gemm_tc_prof_enable(bool on);
std::string gemm_tc_prof_summary();
std::string gemm_tc_prof_shapes();
// beam.cu
int trie_build(Engine* e, const int32_t* paths, const int64_t* offsets, int n_paths, Trie** out);
void trie_free(Trie* t);
void trie_stats(Trie* t, int* n_nodes, int* n_edges, int* max_depth);
int trie_get(Trie* t, const int32_t* prefix, int prefix_len, int32_t* out, int cap);
void generate(Engine* e, const int32_t* ids, const int32_t* mask, const int32_t* ww, int B, int Le, Trie* trie,
              int K, int R, int max_len, float length_penalty, int32_t* seqs, float* scores, int* out_len);
int decode_last_launch(float* ms, double* bytes, int* steps);   // decode_persist.cu
int decode_phase_ns(unsigned long long* out32);
void cooccurrence(const int* items, const long long* offs, int n_users, int n_items, int f64, void* adj, cudaStream_t st);   // indexing.cu
void submatrix(const void* adj, int n_items, int f64, const int* idx, int m, void* out, cudaStream_t st);
// comm.cu
void comm_unique_id(void* id128);
void comm_init(Engine* e, const void* id128, int rank, int world);
void comm_allreduce_grads(Engine* e);
void comm_destroy(Engine* e);
}

extern "C" {

const char* p5_last_error(void) { return g_last_error.c_str(); }
int p5_version(void) { return 100; }
int p5_launch_count(void) { return p5::g_launches + gemm_tc_launch_count(); }
int p5_gemm_tile_width(int M, int N, int batches, int sms) { return (M > 0 && N > 0 && batches > 0 && sms > 0) ? gemm_tc_tile_width(M, N, batches, sms) : 0; }

int p5_create(const P5Config* cfg, int device, void* cuda_stream, p5_handle* out) {
    P5_API_BEGIN
    P5_CHECK(cfg && out, "null argument");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0)
        throw P5Error(4, "p5_create: no CUDA device available — the B200 engine has no CPU fallback");
    P5_CHECK(device >= 0 && device < ndev, "invalid device index");
    Engine* e = new Engine(*cfg, device, (cudaStream_t)cuda_stream);
    *out = new p5_engine{e};
    P5_API_END
}
int p5_destroy(p5_handle h) {
    P5_API_BEGIN
    if (h) {
        if (h->e) { comm_destroy(h->e); delete h->e; }
        delete h;
    }
    P5_API_END
}
int p5_param_count(p5_handle h, int* n) {
    P5_API_BEGIN
    *n = (int)E(h)->params.size();
    P5_API_END
}
int p5_param_info(p5_handle h, int i, const char** name, int* ndim, int64_t shape[2], float** data, float** grad) {
    P5_API_BEGIN
    Engine* e = E(h);
    P5_CHECK(i >= 0 && i < (int)e->params.size(), "parameter index out of range");
    const ParamInfo& p = e->params[i];
    if (name) *name = p.name.c_str();
    if (ndim) *ndim = p.ndim;
    if (shape) { shape[0] = p.shape[0]; shape[1] = p.shape[1]; }
    if (data) *data = e->P + p.off;
    if (grad) *grad = e->G + p.off;
    P5_API_END
}
int p5_resize_vocab(p5_handle h, int new_vocab) {
    P5_API_BEGIN
    Engine* old = E(h);
    if (new_vocab != old->V) {
        h->e = old->resized(new_vocab);
        delete old;
    }
    P5_API_END
}
int p5_params_changed(p5_handle h) {
    P5_API_BEGIN
    E(h)->shadow_stale = true;
    P5_API_END
}
int p5_forward(p5_handle h, const int32_t* input_ids, const int32_t* attention_mask, const int32_t* whole_word_ids,
               const int32_t* labels, int B, int Le, int Ld, float* loss_tok, float* logits_or_null, int training,
               uint64_t seed) {
    P5_API_BEGIN
    Engine* e = E(h);
    P5_CHECK(input_ids && attention_mask && labels, "null input");
    e->forward(input_ids, attention_mask, whole_word_ids, labels, B, Le, Ld, training != 0, seed);
    if (loss_tok)
        P5_CUDA(cudaMemcpyAsync(loss_tok, e->loss_tok, (size_t)B * Ld * 4, cudaMemcpyDeviceToDevice, e->st));
    if (logits_or_null)
        P5_CUDA(cudaMemcpy2DAsync(logits_or_null, (size_t)e->V * 4, e->logits, (size_t)e->Vpad * 4, (size_t)e->V * 4,
                                  (size_t)B * Ld, cudaMemcpyDeviceToDevice, e->st));
    P5_API_END
}
int p5_set_enc_lengths(p5_handle h, const int32_t* lens_host, int B) {
    P5_API_BEGIN
    Engine* e = E(h);
    e->pending_lens.clear();
    if (lens_host && B > 0) e->pending_lens.assign(lens_host, lens_host + B);
    P5_API_END
}
int p5_backward(p5_handle h, const float* dloss_tok) {
    P5_API_BEGIN
    Engine* e = E(h);
    P5_CHECK(dloss_tok != nullptr, "null dloss");
    P5_CUDA(cudaMemcpyAsync(e->dloss, dloss_tok, (size_t)e->Md * 4, cudaMemcpyDeviceToDevice, e->st));
    e->backward();
    P5_API_END
}
// P5_NO_EARLY_NORM=1: the gradient norm is one pass over the whole buffer after the backward (A/B switch)
static bool early_norm_enabled() {
    static const bool on = getenv("P5_NO_EARLY_NORM") == nullptr;
    return on;
}
int p5_train_fwd_bwd(p5_handle h, const int32_t* input_ids, const int32_t* attention_mask,
                     const int32_t* whole_word_ids, const int32_t* labels, const int32_t* labels_mask, int B, int Le,
                     int Ld, float* loss_out, uint64_t seed) {
    P5_API_BEGIN
    Engine* e = E(h);
    P5_CHECK(input_ids && attention_mask && labels && labels_mask, "null input");
    e->forward(input_ids, attention_mask, whole_word_ids, labels, B, Le, Ld, true, seed);
    P5_CUDA(cudaMemcpyAsync(e->lmask, labels_mask, (size_t)B * Ld * 4, cudaMemcpyDeviceToDevice, e->st));
    runner_loss_fwd_bwd(e->loss_tok, e->lmask, B, Ld, e->loss_scalar, e->dloss, e->st);
    if (loss_out) P5_CUDA(cudaMemcpyAsync(loss_out, e->loss_scalar, 4, cudaMemcpyDeviceToDevice, e->st));
    e->overlap_comm = e->world > 1 && e->nccl_comm != nullptr;
    e->early_norm = early_norm_enabled();
    e->backward();
    e->overlap_comm = false;
    e->early_norm = false;
    P5_API_END
}
int p5_grad_norm(p5_handle h, float* out) {
    P5_API_BEGIN
    Engine* e = E(h);
    e->grad_norm();
    if (out) P5_CUDA(cudaMemcpyAsync(out, e->norm_out, 4, cudaMemcpyDeviceToDevice, e->st));
    P5_API_END
}
int p5_grad_scale(p5_handle h, float s) {
    P5_API_BEGIN
    Engine* e = E(h);
    e->join_optimizer();
    scale_f32(e->G, e->n_flat, s, e->st);
    e->invalidate_norm();
    P5_API_END
}
int p5_zero_grad(p5_handle h) {
    P5_API_BEGIN
    E(h)->zero_grad();
    P5_API_END
}
int p5_adamw_step(p5_handle h, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  float clip) {
    P5_API_BEGIN
    P5_CHECK(step >= 1, "AdamW step is 1-based");
    E(h)->adamw(lr, beta1, beta2, eps, weight_decay, step, clip);
    P5_API_END
}

int p5_adamw_step_zero_grad(p5_handle h, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                            float clip) {
    P5_API_BEGIN
    P5_CHECK(step >= 1, "AdamW step is 1-based");
    E(h)->adamw(lr, beta1, beta2, eps, weight_decay, step, clip, true);
    P5_API_END
}

int p5_adamw_step_zero_grad_async(p5_handle h, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                  float clip) {
    P5_API_BEGIN
    P5_CHECK(step >= 1, "AdamW step is 1-based");
    E(h)->adamw_async(lr, beta1, beta2, eps, weight_decay, step, clip);
    P5_API_END
}
int p5_optimizer_join(p5_handle h) {
    P5_API_BEGIN
    E(h)->join_optimizer();
    P5_API_END
}

int p5_eval_metrics(p5_handle h, const int32_t* seqs, const float* scores, int B, int K, int T, const int32_t* gold, int Tg,
                    const int32_t* ks_dev, int n_k, float* out_sums) {
    P5_API_BEGIN
    Engine* e = E(h);
    P5_CUDA(cudaSetDevice(e->device));
    eval_metrics(seqs, scores, B, K, T, gold, Tg, ks_dev, n_k, out_sums, e->st);
    P5_API_END
}

int p5_eval_metrics_filtered(p5_handle h, const int32_t* seqs, const float* scores, int B, int R, int T, const int32_t* gold,
                             int Tg, const int32_t* pos, const int32_t* npos, int Pmax, int Tp, const int32_t* ks_dev, int n_k,
                             int k_cut, float* out_sums) {
    P5_API_BEGIN
    Engine* e = E(h);
    P5_CUDA(cudaSetDevice(e->device));
    P5_CHECK(seqs && scores && gold && npos && ks_dev && out_sums && (pos || Pmax == 0), "null argument");
    eval_metrics_filtered(seqs, scores, B, R, T, gold, Tg, pos, npos, Pmax, Tp, ks_dev, n_k, k_cut, out_sums, e->st);
    P5_API_END
}

int p5_opt_state_info(p5_handle h, int i, float** exp_avg, float** exp_avg_sq) {
    P5_API_BEGIN
    Engine* e = E(h);
    P5_CHECK(i >= 0 && i < (int)e->params.size(), "parameter index out of range");
    e->join_optimizer();
    if (exp_avg) *exp_avg = e->M1 + e->params[i].off;
    if (exp_avg_sq) *exp_avg_sq = e->V2 + e->params[i].off;
    P5_API_END
}

int p5_comm_unique_id(void* id128_host) {
    P5_API_BEGIN
    comm_unique_id(id128_host);
    P5_API_END
}
int p5_comm_init(p5_handle h, const void* id128_host, int rank, int world) {
    P5_API_BEGIN
    comm_init(E(h), id128_host, rank, world);
    P5_API_END
}
int p5_allreduce_grads(p5_handle h) {
    P5_API_BEGIN
    E(h)->join_optimizer();
    comm_allreduce_grads(E(h));
    P5_API_END
}

int p5_trie_build(p5_handle h, const int32_t* paths_host, const int64_t* offsets_host, int n_paths, p5_trie* out) {
    P5_API_BEGIN
    Trie* t = nullptr;
    trie_build(E(h), paths_host, offsets_host, n_paths, &t);
    *out = reinterpret_cast<p5_trie>(t);
    P5_API_END
}
int p5_trie_free(p5_trie t) {
    P5_API_BEGIN
    if (t) trie_free(reinterpret_cast<Trie*>(t));
    P5_API_END
}
int p5_trie_stats(p5_trie t, int* n_nodes, int* n_edges, int* max_depth) {
    P5_API_BEGIN
    P5_CHECK(t != nullptr, "null trie");
    trie_stats(reinterpret_cast<Trie*>(t), n_nodes, n_edges, max_depth);
    P5_API_END
}
int p5_trie_get(p5_trie t, const int32_t* prefix_host, int prefix_len, int32_t* out_tokens_host, int cap, int* n_out) {
    P5_API_BEGIN
    P5_CHECK(t != nullptr, "null trie");
    *n_out = trie_get(reinterpret_cast<Trie*>(t), prefix_host, prefix_len, out_tokens_host, cap);
    P5_API_END
}
int p5_generate(p5_handle h, const int32_t* input_ids, const int32_t* attention_mask, const int32_t* whole_word_ids,
                int B, int Le, p5_trie trie, int num_beams, int num_return, int max_len, float length_penalty,
                int32_t* seqs, float* scores, int* out_len_host) {
    P5_API_BEGIN
    P5_CHECK(trie != nullptr, "null trie");
    generate(E(h), input_ids, attention_mask, whole_word_ids, B, Le, reinterpret_cast<Trie*>(trie), num_beams,
             num_return, max_len, length_penalty, seqs, scores, out_len_host);
    P5_API_END
}

int p5_prof_enable(int on) {
    P5_API_BEGIN
    gemm_tc_prof_enable(on != 0);
    P5_API_END
}
int p5_prof_summary(char* json_out, int cap) {
    P5_API_BEGIN
    P5_CHECK(json_out && cap > 0, "null buffer");
    const std::string s = gemm_tc_prof_summary();
    snprintf(json_out, (size_t)cap, "%s", s.c_str());
    P5_API_END
}

int p5_decode_last_launch(float* ms_host, double* bytes_host, int* steps_host) {
    P5_API_BEGIN
    P5_CHECK(decode_last_launch(ms_host, bytes_host, steps_host) == 0, "no persistent decode launch has been timed yet");
    P5_API_END
}

int p5_decode_phase_ns(uint64_t* out32_host) {
    P5_API_BEGIN
    P5_CHECK(out32_host && decode_phase_ns((unsigned long long*)out32_host) == 0, "no persistent decode launch to report");
    P5_API_END
}
int p5_cooccurrence(const int32_t* items, const int64_t* offsets, int n_users, int n_items, int f64, void* adj, void* cuda_stream) {
    P5_API_BEGIN
    P5_CHECK(items && offsets && adj, "null argument");
    cooccurrence(items, (const long long*)offsets, n_users, n_items, f64, adj, (cudaStream_t)cuda_stream);
    P5_API_END
}
int p5_submatrix(const void* adj, int n_items, int f64, const int32_t* idx, int m, void* out, void* cuda_stream) {
    P5_API_BEGIN
    P5_CHECK(adj && idx && out, "null argument");
    submatrix(adj, n_items, f64, idx, m, out, (cudaStream_t)cuda_stream);
    P5_API_END
}

int p5_prof_shapes(char* text_out, int cap) {
    P5_API_BEGIN
    P5_CHECK(text_out && cap > 0, "null buffer");
    const std::string s = gemm_tc_prof_shapes();
    snprintf(text_out, (size_t)cap, "%s", s.c_str());
    P5_API_END
}

int p5_op_gemm(const P5GemmDesc* d, void* cuda_stream) {
    P5_API_BEGIN
    GemmProblem p;
    p.M = d->M; p.N = d->N; p.K = d->K; p.nb1 = d->nb1 > 0 ? d->nb1 : 1; p.nb2 = d->nb2 > 0 ? d->nb2 : 1;
    p.A.ptr = d->A; p.A.dtype = d->a_dtype; p.A.major = d->a_major; p.A.ld = d->lda; p.A.bs1 = d->a_bs1; p.A.bs2 = d->a_bs2;
    p.B.ptr = d->B; p.B.dtype = d->b_dtype; p.B.major = d->b_major; p.B.ld = d->ldb; p.B.bs1 = d->b_bs1; p.B.bs2 = d->b_bs2;
    p.epi.C = d->C; p.epi.c_dtype = d->c_dtype; p.epi.ldc = d->ldc; p.epi.cs1 = d->c_bs1; p.epi.cs2 = d->c_bs2;
    p.epi.alpha = d->alpha; p.epi.flags = d->flags; p.epi.aux = d->aux; p.epi.aux_dtype = d->aux_dtype;
    p.epi.resid = d->resid; p.epi.seed = d->seed; p.epi.site = d->site;
    p.epi.drop_thr = drop_threshold(d->drop_p);
    p.epi.inv_keep = d->drop_p < 1.f ? 1.f / (1.f - d->drop_p) : 0.f;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (d->backend == 0) {
        gemm_simt(p, st);
        ++p5::g_launches;
    } else if (d->backend == 1) {
        gemm_tc_force_block_n(d->force_block_n);
        gemm_tc(p, st);
        gemm_tc_force_block_n(0);
    } else {
        gemm_auto(p, st, true);
    }
    P5_API_END
}

}  // extern "C"
