// Trie-constrained beam search on device.
//
// Replaces, for `model.generate(num_beams=K, num_return_sequences=R, prefix_allowed_tokens_fn=trie)`
// (ref src/src_t5/runner/DistributedRunner.py:344-371):
//   * utils/generation_trie.py:7-97      nested-dict Trie + per-beam Python callback  ->  CSR trie in HBM, each
//                                         running beam carries its trie NODE (no prefix re-walk, no D2H sync)
//   * HF:generation/logits_process.py:1532-1549 (-inf mask over V) -> only the children of the node are scored
//   * HF:generation/utils.py:3231-3378   _beam_search step: log_softmax, + running score, top-2K over K*V,
//                                         running/finished split, length-normalised finished scores, early-stop
//                                         heuristic, KV-cache reorder -> index indirection (no KV copies)
//   * P5_T5.py:543-578                    encoder states repeated xK  -> cross-K/V stored once per USER
// Semantics mirror transformers 5.5 `_beam_search` (early_stopping=False, do_sample=False); ties are broken by
// lowest flat index (beam * V + token).  The whole search is enqueued without host synchronisation: the step
// count is the trie depth, and finished users are frozen by the same `-1e9` gating HF uses.
#include "engine.h"
#include <map>
#include <vector>
#include <float.h>

namespace p5 {
#define LAUNCHED() do { P5_CUDA(cudaGetLastError()); ++g_launches; } while (0)
static constexpr float NEG_BIG = -1.0e9f;   // HF:generation/utils.py:3200-3201

// ------------------------------------------------------------------------------------------------------------
// trie
// ------------------------------------------------------------------------------------------------------------
struct Trie {
    int device = 0;
    int n_nodes = 0, n_edges = 0, max_depth = 0, max_fanout = 0;
    std::vector<int> h_off, h_tok, h_node;
    int *d_off = nullptr, *d_tok = nullptr, *d_node = nullptr;
};

int trie_build(Engine* e, const int32_t* paths, const int64_t* offsets, int n_paths, Trie** out) {
    P5_CHECK(paths && offsets && n_paths > 0, "trie_build: empty path set");
    std::vector<std::map<int, int>> nodes(1);
    int max_depth = 0;
    for (int i = 0; i < n_paths; ++i) {
        int cur = 0;
        const int64_t n = offsets[i + 1] - offsets[i];
        if ((int)n > max_depth) max_depth = (int)n;
        for (int64_t j = offsets[i]; j < offsets[i + 1]; ++j) {
            const int t = paths[j];
            auto it = nodes[cur].find(t);
            if (it == nodes[cur].end()) {
                nodes.push_back({});
                const int id = (int)nodes.size() - 1;
                nodes[cur][t] = id;
                cur = id;
            } else {
                cur = it->second;
            }
        }
    }
    Trie* t = new Trie();
    t->device = e->device;
    t->n_nodes = (int)nodes.size();
    t->max_depth = max_depth;
    t->h_off.resize(t->n_nodes + 1);
    int edges = 0;
    for (int i = 0; i < t->n_nodes; ++i) {
        t->h_off[i] = edges;
        edges += (int)nodes[i].size();
        if ((int)nodes[i].size() > t->max_fanout) t->max_fanout = (int)nodes[i].size();
    }
    t->h_off[t->n_nodes] = edges;
    t->n_edges = edges;
    t->h_tok.resize(edges > 0 ? edges : 1);
    t->h_node.resize(edges > 0 ? edges : 1);
    for (int i = 0; i < t->n_nodes; ++i) {
        int k = t->h_off[i];
        for (auto& kv : nodes[i]) { t->h_tok[k] = kv.first; t->h_node[k] = kv.second; ++k; }
    }
    P5_CUDA(cudaSetDevice(e->device));
    P5_CUDA(cudaMalloc(&t->d_off, (t->n_nodes + 1) * sizeof(int)));
    P5_CUDA(cudaMalloc(&t->d_tok, t->h_tok.size() * sizeof(int)));
    P5_CUDA(cudaMalloc(&t->d_node, t->h_node.size() * sizeof(int)));
    P5_CUDA(cudaMemcpy(t->d_off, t->h_off.data(), (t->n_nodes + 1) * sizeof(int), cudaMemcpyHostToDevice));
    P5_CUDA(cudaMemcpy(t->d_tok, t->h_tok.data(), t->h_tok.size() * sizeof(int), cudaMemcpyHostToDevice));
    P5_CUDA(cudaMemcpy(t->d_node, t->h_node.data(), t->h_node.size() * sizeof(int), cudaMemcpyHostToDevice));
    *out = t;
    return 0;
}
void trie_free(Trie* t) {
    cudaSetDevice(t->device);
    cudaFree(t->d_off); cudaFree(t->d_tok); cudaFree(t->d_node);
    delete t;
}
void trie_stats(Trie* t, int* n_nodes, int* n_edges, int* max_depth) {
    if (n_nodes) *n_nodes = t->n_nodes;
    if (n_edges) *n_edges = t->n_edges;
    if (max_depth) *max_depth = t->max_depth;
}

__device__ __forceinline__ int trie_child(const int* off, const int* tok, const int* node, int n, int t) {
    if (n < 0) return -1;
    int lo = off[n], hi = off[n + 1] - 1;
    while (lo <= hi) {   // children are sorted by token
        const int mid = (lo + hi) >> 1;
        const int v = tok[mid];
        if (v == t) return node[mid];
        if (v < t) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}
// walks `prefix` from the root on the DEVICE copy and lists the children (Trie.get of the reference)
__global__ void trie_get_kernel(const int* off, const int* tok, const int* node, const int* prefix, int n, int* out,
                                int cap, int* n_out) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    int cur = 0;
    for (int i = 0; i < n && cur >= 0; ++i) cur = trie_child(off, tok, node, cur, prefix[i]);
    int cnt = 0;
    if (cur >= 0)
        for (int k = off[cur]; k < off[cur + 1]; ++k) { if (cnt < cap) out[cnt] = tok[k]; ++cnt; }
    *n_out = cnt;
}
int trie_get(Trie* t, const int32_t* prefix, int prefix_len, int32_t* out, int cap) {
    P5_CUDA(cudaSetDevice(t->device));
    int *d_pre = nullptr, *d_out = nullptr, *d_n = nullptr;
    P5_CUDA(cudaMalloc(&d_pre, (prefix_len + 1) * sizeof(int)));
    P5_CUDA(cudaMalloc(&d_out, (cap + 1) * sizeof(int)));
    P5_CUDA(cudaMalloc(&d_n, sizeof(int)));
    if (prefix_len) P5_CUDA(cudaMemcpy(d_pre, prefix, prefix_len * sizeof(int), cudaMemcpyHostToDevice));
    launch_k(trie_get_kernel, 1, 1, 0, nullptr, t->d_off, t->d_tok, t->d_node, d_pre, prefix_len, d_out, cap, d_n);
    P5_CUDA(cudaGetLastError());
    int n = 0;
    P5_CUDA(cudaMemcpy(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost));
    if (n > 0) P5_CUDA(cudaMemcpy(out, d_out, (n < cap ? n : cap) * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_pre); cudaFree(d_out); cudaFree(d_n);
    return n;
}

// ------------------------------------------------------------------------------------------------------------
// generate workspace
// ------------------------------------------------------------------------------------------------------------
struct GenWs {
    int Rm = 0, Tm = 0, Km = 0, Bm = 0;
    float* y = nullptr;
    void *n = nullptr, *qkv = nullptr, *ctx = nullptr, *cq = nullptr, *h = nullptr, *z = nullptr;
    std::vector<void*> Kc, Vc;
    float *logits = nullptr, *rowmax = nullptr, *logsum = nullptr;
    float* xS = nullptr; void* xP = nullptr; int xLe = 0;   // cross-attention scores / probabilities [B, H, K, Le]
    int* seq[2] = {nullptr, nullptr};      // running sequences [R, T]
    int* fin_seq[2] = {nullptr, nullptr};  // finished sequences [R, T]
    int* src[2] = {nullptr, nullptr};      // KV-cache row indirection [R, T]
    int* node[2] = {nullptr, nullptr};     // trie node of each running beam
    float* run_score[2] = {nullptr, nullptr};
    float* fin_score[2] = {nullptr, nullptr};
    int* is_fin[2] = {nullptr, nullptr};
    int* gen_len[2] = {nullptr, nullptr};
    int *cur_tok = nullptr, *unsat = nullptr, *out_len = nullptr;
    float* cand_lp = nullptr; int *cand_beam = nullptr, *cand_tok = nullptr;
    float* scr_score = nullptr; int* scr_flat = nullptr; int scr_cap = 0;
    std::vector<void*> allocs;
};
// decode_persist.cu: the whole search as one persistent cooperative kernel (bf16 mode)
bool decode_persistent_supported(const Engine* e, int B, int K, int max_len, int Le);
const int* generate_persistent(Engine* e, const int* t_off, const int* t_tok, const int* t_node, int root_child, int max_depth,
                               int max_fanout, int B, int K, int Rret, int max_len, float length_penalty, int32_t* seqs, float* scores,
                               const int* forced_host, int n_forced, int node_forced);
void free_persist_ws();

void free_gen_ws(GenWs* g) {
    for (void* p : g->allocs) cudaFree(p);
    delete g;
}
static GenWs* get_gen_ws(Engine* e, int R, int T, int K, int B, int cand_cap, int Le) {
    GenWs* g = e->gen;
    if (g && g->Rm >= R && g->Tm >= T && g->Km >= K && g->Bm >= B && g->scr_cap >= cand_cap && g->xLe >= Le) return g;
    if (g) { P5_CUDA(cudaStreamSynchronize(e->st)); free_gen_ws(g); e->gen = nullptr; }
    g = new GenWs();
    auto al = [&](size_t bytes) { void* p = nullptr; P5_CUDA(cudaMalloc(&p, bytes ? bytes : 256)); g->allocs.push_back(p); return p; };
    const size_t es = e->esz();
    const int A = e->A, d = e->d, ff = e->ff;
    g->Rm = R; g->Tm = T; g->Km = K; g->Bm = B; g->scr_cap = cand_cap;
    g->y = (float*)al((size_t)R * d * 4);
    g->n = al((size_t)R * d * es); g->qkv = al((size_t)R * 3 * A * es); g->ctx = al((size_t)R * A * es);
    g->cq = al((size_t)R * A * es); g->h = al((size_t)R * ff * es);
    if (e->gated) g->z = al((size_t)R * 2 * ff * es);
    g->Kc.resize(e->ND); g->Vc.resize(e->ND);
    for (int l = 0; l < e->ND; ++l) { g->Kc[l] = al((size_t)R * T * A * es); g->Vc[l] = al((size_t)R * T * A * es); }
    g->logits = (float*)al((size_t)R * e->Vpad * 4);
    g->xLe = Le;
    g->xS = (float*)al((size_t)R * e->H * Le * 4);
    g->xP = al((size_t)R * e->H * Le * es);
    g->rowmax = (float*)al((size_t)R * 4); g->logsum = (float*)al((size_t)R * 4);
    for (int i = 0; i < 2; ++i) {
        g->seq[i] = (int*)al((size_t)R * T * 4); g->fin_seq[i] = (int*)al((size_t)R * T * 4);
        g->src[i] = (int*)al((size_t)R * T * 4); g->node[i] = (int*)al((size_t)R * 4);
        g->run_score[i] = (float*)al((size_t)R * 4); g->fin_score[i] = (float*)al((size_t)R * 4);
        g->is_fin[i] = (int*)al((size_t)R * 4); g->gen_len[i] = (int*)al((size_t)R * 4);
    }
    g->cur_tok = (int*)al((size_t)R * 4); g->unsat = (int*)al((size_t)B * 4); g->out_len = (int*)al(16);
    g->cand_lp = (float*)al((size_t)B * 2 * K * 4); g->cand_beam = (int*)al((size_t)B * 2 * K * 4);
    g->cand_tok = (int*)al((size_t)B * 2 * K * 4);
    g->scr_score = (float*)al((size_t)B * cand_cap * 4); g->scr_flat = (int*)al((size_t)B * cand_cap * 4);
    e->gen = g;
    return g;
}

// ------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------
__global__ void gen_init_kernel(int* seq, int* fin_seq, int* src, int* node, float* run_score, float* fin_score,
                                int* is_fin, int* gen_len, int* cur_tok, int* unsat, int B, int K, int T, int root_child) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B * K) return;
    for (int t = 0; t < T; ++t) { seq[r * T + t] = 0; fin_seq[r * T + t] = 0; src[r * T + t] = r; }
    node[r] = root_child;                        // node reached by the decoder-start token 0
    run_score[r] = (r % K == 0) ? 0.f : NEG_BIG; // only beam 0 is live at the first step
    fin_score[r] = NEG_BIG;
    is_fin[r] = 0; gen_len[r] = 0; cur_tok[r] = 0;
    if (r % K == 0) unsat[r / K] = 1;
}

// decode self-attention for ONE new position per row, with KV append and row indirection.
// grid (H, R), 64 threads.  K/V cache layout [R, T, A].
template <typename T> __device__ __forceinline__ float2 ld2(const T* p);
template <> __device__ __forceinline__ float2 ld2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 ld2<bf16>(const bf16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
template <typename T> __device__ __forceinline__ void st2(T* p, float2 v);
template <> __device__ __forceinline__ void st2<float>(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
template <> __device__ __forceinline__ void st2<bf16>(bf16* p, float2 v) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y); }

// one warp per (beam row, head): lane owns two of the 64 head columns; the <= max_length cached positions are walked
// once with an online softmax (running max / sum / weighted V), scores reduced with warp shuffles
template <typename T>
__global__ void __launch_bounds__(128)
decode_self_attn_kernel(const T* __restrict__ qkv, T* __restrict__ Kc, T* __restrict__ Vc, const int* __restrict__ src,
                        const float* __restrict__ bias_rel, int n_delta, int bias_off, T* __restrict__ ctx, int A, int Tm,
                        int pos, int R, int H) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (gw >= R * H) return;
    const int r = gw / H, h = gw % H, c = 2 * lane;
    const T* row = qkv + (int64_t)r * 3 * A + h * 64 + c;
    const float2 q = ld2<T>(row), kc = ld2<T>(row + A), vc = ld2<T>(row + 2 * A);
    st2<T>(Kc + ((int64_t)r * Tm + pos) * A + h * 64 + c, kc);
    st2<T>(Vc + ((int64_t)r * Tm + pos) * A + h * 64 + c, vc);
    float m = -INFINITY, l = 0.f;
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j <= pos; ++j) {
        float2 k = kc, v = vc;
        if (j < pos) {
            const int64_t o = ((int64_t)src[r * Tm + j] * Tm + j) * A + h * 64 + c;
            k = ld2<T>(Kc + o);
            v = ld2<T>(Vc + o);
        }
        float sc = warp_sum(fmaf(q.x, k.x, q.y * k.y));
        int di = j - pos + bias_off;
        di = di < 0 ? 0 : (di >= n_delta ? n_delta - 1 : di);
        sc += bias_rel[h * n_delta + di];
        const float mn = fmaxf(m, sc);
        const float scale = __expf(m - mn), p = __expf(sc - mn);     // first step: exp(-inf) = 0
        l = l * scale + p;
        acc.x = fmaf(acc.x, scale, p * v.x);
        acc.y = fmaf(acc.y, scale, p * v.y);
        m = mn;
    }
    const float inv = 1.f / l;
    st2<T>(ctx + (int64_t)r * A + h * 64 + c, make_float2(acc.x * inv, acc.y * inv));
}

// row-wise max and log(sum(exp(x - max))) over the vocabulary (torch log_softmax = (x - max) - logsum)
__global__ void __launch_bounds__(512)
rows_logsumexp_kernel(const float* __restrict__ logits, int64_t ld, int V, float* __restrict__ rowmax,
                      float* __restrict__ logsum) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    __shared__ float sh[32];
    const int r = blockIdx.x;
    const float* l = logits + (int64_t)r * ld;
    // single pass: per-thread online (max, sum) over float4 chunks, then a block combine
    float mx = -INFINITY, s = 0.f;
    const int V4 = V >> 2;
    const float4* l4 = reinterpret_cast<const float4*>(l);
    for (int c = threadIdx.x; c < V4; c += blockDim.x) {
        const float4 v = l4[c];
        const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (m4 > mx) { s *= __expf(mx - m4); mx = m4; }
        s += __expf(v.x - mx) + __expf(v.y - mx) + __expf(v.z - mx) + __expf(v.w - mx);
    }
    for (int c = (V4 << 2) + threadIdx.x; c < V; c += blockDim.x) {
        const float v = l[c];
        if (v > mx) { s *= __expf(mx - v); mx = v; }
        s += __expf(v - mx);
    }
    const float bm = block_max(mx, sh);
    s = (mx > -INFINITY) ? s * __expf(mx - bm) : 0.f;
    s = block_sum(s, sh);
    if (threadIdx.x == 0) { rowmax[r] = bm; logsum[r] = logf(s); }
}

// per user: score the trie children of every running beam and keep the best 2K (score desc, flat index asc)
__global__ void __launch_bounds__(256)
topk_candidates_kernel(const float* __restrict__ logits, int64_t ld, int V, const float* __restrict__ rowmax,
                       const float* __restrict__ logsum, const float* __restrict__ run_score,
                       const int* __restrict__ node, const int* __restrict__ t_off, const int* __restrict__ t_tok,
                       int K, float* __restrict__ scr_score, int* __restrict__ scr_flat, int scr_cap,
                       float* __restrict__ cand_lp, int* __restrict__ cand_beam, int* __restrict__ cand_tok) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    constexpr int MAXK = 64, NS = 1024;
    __shared__ int s_off[MAXK + 1], s_nd[MAXK];
    __shared__ float s_sc[NS];
    __shared__ int s_fl[NS];
    __shared__ float s_best[8];
    __shared__ int s_besti[8], s_bestf[8];
    const int b = blockIdx.x;
    float* sc = scr_score + (int64_t)b * scr_cap;
    int* fl = scr_flat + (int64_t)b * scr_cap;
    // children of every running beam: counts -> exclusive offsets (K <= 64: one thread adds them up)
    if (threadIdx.x < K) s_nd[threadIdx.x] = node[b * K + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < K; ++k) {
            s_off[k] = acc;
            const int nd = s_nd[k];
            if (nd >= 0) acc += t_off[nd + 1] - t_off[nd];
        }
        s_off[K] = acc;
    }
    __syncthreads();
    const int n = min(s_off[K], scr_cap);
    const bool in_smem = n <= NS;         // the usual case (a few children per beam): no global scratch, no selection loop
    float* vsc = in_smem ? s_sc : sc;
    int* vfl = in_smem ? s_fl : fl;
    for (int k = 0; k < K; ++k) {
        const int nd = s_nd[k];
        if (nd < 0) continue;
        const int r = b * K + k, e0 = t_off[nd], cnt = s_off[k + 1] - s_off[k], base = s_off[k];
        const float rs = run_score[r], nrm = rowmax[r], ls = logsum[r];
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const int tok = t_tok[e0 + e];
            const int slot = base + e;
            if (slot < scr_cap && tok >= 0 && tok < V) {
                const float lp = (logits[(int64_t)r * ld + tok] - nrm) - ls;
                vsc[slot] = lp + rs;
                vfl[slot] = k * V + tok;
            } else if (slot < scr_cap) {
                vsc[slot] = -INFINITY; vfl[slot] = 0x7fffffff;
            }
        }
    }
    // fewer than 2K continuations: HF's remaining top-k slots hold -inf entries
    for (int sel = threadIdx.x; sel < 2 * K; sel += blockDim.x) {
        const int o = b * 2 * K + sel;
        cand_lp[o] = -INFINITY; cand_beam[o] = 0; cand_tok[o] = 0;
    }
    __syncthreads();
    if (in_smem) {
        // rank of a candidate in the order (score desc, flat index asc) == the slot torch.topk would give it
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float v = s_sc[i];
            const int f = s_fl[i];
            if (!(v > -INFINITY)) continue;
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const float vj = s_sc[j];
                rank += (vj > v || (vj == v && s_fl[j] < f)) ? 1 : 0;
            }
            if (rank < 2 * K) {
                const int o = b * 2 * K + rank;
                cand_lp[o] = v; cand_beam[o] = f / V; cand_tok[o] = f % V;
            }
        }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int sel = 0; sel < 2 * K; ++sel) {
        float best = -INFINITY; int bi = -1, bf = 0x7fffffff;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float v = sc[i]; const int f = fl[i];
            if (v > best || (v == best && v > -INFINITY && f < bf)) { best = v; bi = i; bf = f; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int of = __shfl_xor_sync(0xffffffffu, bf, o);
            if (ov > best || (ov == best && ov > -INFINITY && of < bf)) { best = ov; bi = oi; bf = of; }
        }
        if (lane == 0) { s_best[warp] = best; s_besti[warp] = bi; s_bestf[warp] = bf; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float bb = -INFINITY; int ii = -1, ff = 0x7fffffff;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
                if (s_best[w] > bb || (s_best[w] == bb && bb > -INFINITY && s_bestf[w] < ff)) { bb = s_best[w]; ii = s_besti[w]; ff = s_bestf[w]; }
            const int o = b * 2 * K + sel;
            if (ii >= 0 && bb > -INFINITY) {
                cand_lp[o] = bb; cand_beam[o] = ff / V; cand_tok[o] = ff % V;
                sc[ii] = -INFINITY;
            }
        }
        __syncthreads();
    }
}

// per user: HF:generation/utils.py:2999-3073 (running beams, finished beams) + :2876-2921 (early stop heuristic)
__global__ void __launch_bounds__(128)
beam_update_kernel(const float* __restrict__ cand_lp, const int* __restrict__ cand_beam, const int* __restrict__ cand_tok,
                   const int* __restrict__ seq_in, int* __restrict__ seq_out, const int* __restrict__ fin_in,
                   int* __restrict__ fin_out, const int* __restrict__ src_in, int* __restrict__ src_out,
                   const int* __restrict__ node_in, int* __restrict__ node_out, float* __restrict__ run_out,
                   const float* __restrict__ fscore_in, float* __restrict__ fscore_out, const int* __restrict__ isfin_in,
                   int* __restrict__ isfin_out, const int* __restrict__ glen_in, int* __restrict__ glen_out,
                   int* __restrict__ cur_tok, int* __restrict__ unsat, const int* __restrict__ t_off,
                   const int* __restrict__ t_tok, const int* __restrict__ t_node, int K, int T, int cur_len, int max_len,
                   int eos, float denom_fin, float denom_next) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    extern __shared__ int smi[];
    int* run_sel = smi;              // [K]   candidate index chosen for running slot k
    int* fin_sel = smi + K;          // [K]   merged index chosen for finished slot k
    float* s_lp = reinterpret_cast<float*>(smi + 2 * K);   // [2K] candidates
    int* s_cb = smi + 4 * K;
    int* s_ct = smi + 6 * K;
    float* s_rv = reinterpret_cast<float*>(smi + 8 * K);   // [2K] running-beam keys
    float* s_fv = reinterpret_cast<float*>(smi + 10 * K);  // [3K] finished-beam keys: old finished (K) | candidates (2K)
    int* s_fin = smi + 13 * K;       // [K]   is_finished of the selected finished slots
    float* s_misc = reinterpret_cast<float*>(smi + 14 * K);   // [0] best running score, [1] worst kept finished score
    const int b = blockIdx.x;
    const bool at_max = (cur_len + 1 >= max_len);
    const bool us = unsat[b] != 0;
    for (int c = threadIdx.x; c < 2 * K; c += blockDim.x) {
        const float l = cand_lp[b * 2 * K + c];
        const int tk = cand_tok[b * 2 * K + c];
        s_lp[c] = l; s_cb[c] = cand_beam[b * 2 * K + c]; s_ct[c] = tk;
        const bool hit = (tk == eos) || at_max;
        s_rv[c] = l + (hit ? NEG_BIG : -0.0f);                 // running beams: top-K of lp + hits * -1e9
        float v = l / denom_fin;                                // finished beams: candidates that hit EOS inside the top K
        v += us ? -0.0f : NEG_BIG;
        v += (hit && c < K) ? -0.0f : NEG_BIG;
        s_fv[K + c] = v;
    }
    for (int m = threadIdx.x; m < K; m += blockDim.x) s_fv[m] = fscore_in[b * K + m];
    __syncthreads();
    const float* lp = s_lp;
    const int* cb = s_cb;
    const int* ct = s_ct;
    // Selection by RANK instead of K serial arg-max rounds: an element's rank in the order (value desc, index asc) is
    // the round in which the serial selection (ties -> lowest index, the order torch.topk yields; SURVEY 7 tie-break)
    // would have picked it.
    for (int c = threadIdx.x; c < 2 * K; c += blockDim.x) {
        const float v = s_rv[c];
        int rank = 0;
        for (int j = 0; j < 2 * K; ++j) rank += (s_rv[j] > v || (s_rv[j] == v && j < c)) ? 1 : 0;
        if (rank < K) {
            run_sel[rank] = c;
            run_out[b * K + rank] = v;
            if (rank == 0) s_misc[0] = v;
        }
    }
    for (int m = threadIdx.x; m < 3 * K; m += blockDim.x) {
        const float v = s_fv[m];
        int rank = 0;
        for (int j = 0; j < 3 * K; ++j) rank += (s_fv[j] > v || (s_fv[j] == v && j < m)) ? 1 : 0;
        if (rank < K) {
            fin_sel[rank] = m;
            fscore_out[b * K + rank] = v;
            int fin, gl;
            if (m < K) { fin = isfin_in[b * K + m]; gl = glen_in[b * K + m]; }
            else {
                const int c = m - K;
                const bool hit = (ct[c] == eos) || at_max;
                fin = (hit && c < K) ? 1 : 0;
                gl = cur_len;   // cur_len + 1 - prompt_len(=1)
            }
            isfin_out[b * K + rank] = fin; glen_out[b * K + rank] = gl;
            s_fin[rank] = fin;
            if (rank == K - 1) s_misc[1] = v;      // min over the kept finished scores
        }
    }
    __syncthreads();
    // ---- early-stop heuristic with cur_len already incremented
    if (threadIdx.x == 0) {
        const float best_running = s_misc[0] / denom_next, min_fs = s_misc[1];
        bool any = false;
        for (int k = 0; k < K; ++k) {
            const float worst = s_fin[k] ? min_fs : NEG_BIG;
            if (best_running > worst) any = true;
        }
        unsat[b] = (us && any) ? 1 : 0;
    }
    // ---- materialise the selected rows (all threads)
    for (int idx = threadIdx.x; idx < K * T; idx += blockDim.x) {
        const int k = idx / T, t = idx - k * T;
        const int c = run_sel[k];
        const int parent = b * K + cb[c];
        const int r = b * K + k;
        int v = seq_in[parent * T + t];
        if (t == cur_len) v = ct[c];
        seq_out[r * T + t] = v;
        int sr = src_in[parent * T + t];
        if (t >= cur_len) sr = r;
        src_out[r * T + t] = sr;
        const int m = fin_sel[k];
        int fv;
        if (m < K) fv = fin_in[(b * K + m) * T + t];
        else {
            const int c2 = m - K;
            fv = seq_in[(b * K + cb[c2]) * T + t];
            if (t == cur_len) fv = ct[c2];
        }
        fin_out[r * T + t] = fv;
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int c = run_sel[k];
        const int parent = b * K + cb[c];
        const int r = b * K + k;
        cur_tok[r] = ct[c];
        node_out[r] = (lp[c] > -INFINITY) ? trie_child(t_off, t_tok, t_node, node_in[parent], ct[c]) : -1;
    }
}

__global__ void gen_finalize_kernel(const int* __restrict__ fin_seq, const float* __restrict__ fin_score,
                                    const int* __restrict__ is_fin, const int* __restrict__ gen_len, int B, int K, int R,
                                    int T, int max_len, int32_t* __restrict__ seqs, float* __restrict__ scores,
                                    int* __restrict__ out_len) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    __shared__ int s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < B * R; i += blockDim.x) {
        const int b = i / R, k = i % R;
        const int r = b * K + k;
        scores[i] = fin_score[r];
        for (int t = 0; t < max_len; ++t) seqs[(int64_t)i * max_len + t] = fin_seq[r * T + t];
        if (is_fin[r]) atomicMax(&s_max, gen_len[r]);
    }
    __syncthreads();
    if (threadIdx.x == 0) out_len[0] = 1 + s_max;
}

// ------------------------------------------------------------------------------------------------------------
// ranking metrics on the device (ref utils/evaluate.py:37-92, DistributedRunner.py:376-393): per user, the K generated
// rows are ordered by score (desc, stable), marked 1 where their token path (pad 0 / eos 1 stripped ==
// batch_decode(skip_special_tokens=True) as a key) equals the gold item's; hit@k = any(rel[:k]),
// ndcg@k = sum_r rel[r] / log2(r + 2).  Sums over users are accumulated into out[0 .. n_k) (hit) and out[n_k .. 2 n_k).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
eval_metrics_kernel(const int32_t* __restrict__ seqs, const float* __restrict__ scores, int K, int T,
                    const int32_t* __restrict__ gold, int Tg, const int32_t* __restrict__ ks, int n_k, float* __restrict__ out) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    __shared__ int rel[64];
    const int b = blockIdx.x, i = threadIdx.x;
    if (i < 64) rel[i] = 0;
    __syncthreads();
    if (i < K) {
        const int32_t* p = seqs + (int64_t)(b * K + i) * T;
        const int32_t* g = gold + (int64_t)b * Tg;
        int a = 0, c = 0;
        bool same = true;
        while (true) {     // two cursors over the rows, skipping pad (0) and eos (1)
            while (a < T && (p[a] == 0 || p[a] == 1)) ++a;
            while (c < Tg && (g[c] == 0 || g[c] == 1)) ++c;
            if (a >= T || c >= Tg) { same = (a >= T) && (c >= Tg); break; }
            if (p[a] != g[c]) { same = false; break; }
            ++a; ++c;
        }
        const float sc = scores[b * K + i];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const float sj = scores[b * K + j];
            rank += (sj > sc || (sj == sc && j < i)) ? 1 : 0;
        }
        rel[rank] = same ? 1 : 0;
    }
    __syncthreads();
    if (i < n_k) {
        const int k = min(ks[i], K);
        bool hit = false;
        float ndcg = 0.f;
        for (int r = 0; r < k; ++r)
            if (rel[r]) { hit = true; ndcg += 1.f / log2f((float)r + 2.f); }
        if (hit) atomicAdd(out + i, 1.f);
        if (ndcg != 0.f) atomicAdd(out + n_k + i, ndcg);
    }
}

// two token rows are the same item iff they agree after dropping pad (0) and eos (1)
__device__ __forceinline__ bool same_path(const int32_t* p, int T, const int32_t* g, int Tg) {
    int a = 0, c = 0;
    while (true) {
        while (a < T && (p[a] == 0 || p[a] == 1)) ++a;
        while (c < Tg && (g[c] == 0 || g[c] == 1)) ++c;
        if (a >= T || c >= Tg) return (a >= T) && (c >= Tg);
        if (p[a] != g[c]) return false;
        ++a; ++c;
    }
}

// filtered variant (ref utils/evaluate.py:6-35 rel_results_filtered, DistributedRunner.py:204-265): the R returned rows of a
// user are ordered by score (desc, stable); rows equal to one of the user's POSITIVE (already interacted) items are
// skipped; the first k_cut remaining rows form the relevance list.  pos [B, Pmax, Tp] token paths, npos [B] valid counts.
__global__ void __launch_bounds__(64)
eval_metrics_filtered_kernel(const int32_t* __restrict__ seqs, const float* __restrict__ scores, int R, int T,
                             const int32_t* __restrict__ gold, int Tg, const int32_t* __restrict__ pos,
                             const int32_t* __restrict__ npos, int Pmax, int Tp, const int32_t* __restrict__ ks, int n_k,
                             int k_cut, float* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ int s_gold[64], s_pos[64], rel[64];
    __shared__ int n_rel;
    const int b = blockIdx.x, i = threadIdx.x;
    if (i < R) {
        const int32_t* p = seqs + (int64_t)(b * R + i) * T;
        const bool is_gold = same_path(p, T, gold + (int64_t)b * Tg, Tg);
        bool is_pos = false;
        const int np = min(npos[b], Pmax);
        for (int j = 0; j < np && !is_pos; ++j) is_pos = same_path(p, T, pos + ((int64_t)b * Pmax + j) * Tp, Tp);
        const float sc = scores[b * R + i];
        int rank = 0;
        for (int j = 0; j < R; ++j) {
            const float sj = scores[b * R + j];
            rank += (sj > sc || (sj == sc && j < i)) ? 1 : 0;
        }
        s_gold[rank] = is_gold ? 1 : 0;
        s_pos[rank] = is_pos ? 1 : 0;
    }
    __syncthreads();
    if (i == 0) {
        int n = 0;
        for (int r = 0; r < R && n < k_cut; ++r)
            if (!s_pos[r]) rel[n++] = s_gold[r];
        n_rel = n;
    }
    __syncthreads();
    if (i < n_k) {
        const int k = min(ks[i], n_rel);
        bool hit = false;
        float ndcg = 0.f;
        for (int r = 0; r < k; ++r)
            if (rel[r]) { hit = true; ndcg += 1.f / log2f((float)r + 2.f); }
        if (hit) atomicAdd(out + i, 1.f);
        if (ndcg != 0.f) atomicAdd(out + n_k + i, ndcg);
    }
}

void eval_metrics_filtered(const int32_t* seqs, const float* scores, int B, int R, int T, const int32_t* gold, int Tg,
                           const int32_t* pos, const int32_t* npos, int Pmax, int Tp, const int32_t* ks_dev, int n_k, int k_cut,
                           float* out_sums, cudaStream_t st) {
    if (B <= 0) return;
    P5_CHECK(R >= 1 && R <= 64 && n_k >= 1 && n_k <= 64 && T >= 1 && Tg >= 1 && k_cut >= 1 && k_cut <= 64 && Pmax >= 0 && Tp >= 1,
             "eval_metrics_filtered: rows per user, cut-offs and k must be in [1, 64]");
    launch_k(eval_metrics_filtered_kernel, (unsigned)B, 64, 0, st, seqs, scores, R, T, gold, Tg, pos, npos, Pmax, Tp, ks_dev, n_k, k_cut,
             out_sums);
    LAUNCHED();
}

void eval_metrics(const int32_t* seqs, const float* scores, int B, int K, int T, const int32_t* gold, int Tg, const int32_t* ks_dev,
                  int n_k, float* out_sums, cudaStream_t st) {
    if (B <= 0) return;
    P5_CHECK(K >= 1 && K <= 64 && n_k >= 1 && n_k <= 64 && T >= 1 && Tg >= 1, "eval_metrics: K and the number of cut-offs must be in [1, 64]");
    launch_k(eval_metrics_kernel, (unsigned)B, 64, 0, st, seqs, scores, K, T, gold, Tg, ks_dev, n_k, out_sums);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------------------
void generate(Engine* e, const int32_t* ids, const int32_t* mask, const int32_t* ww, int B, int Le_user, Trie* trie,
              int K, int Rret, int max_len, float length_penalty, int32_t* seqs, float* scores, int* out_len_host) {
    P5_CHECK(K >= 1 && K <= 64, "num_beams must be in [1, 64]");
    P5_CHECK(Rret >= 1 && Rret <= K, "num_return_sequences must be in [1, num_beams]");
    P5_CHECK(max_len >= 2 && max_len <= 256, "max_length must be in [2, 256]");
    P5_CHECK(ids && mask, "null input");
    P5_CUDA(cudaSetDevice(e->device));
    e->join_optimizer();          // an asynchronous AdamW of the last train step must have updated every weight
    cudaStream_t st = e->st;
    const int dt = e->dt, d = e->d, A = e->A, H = e->H, ff = e->ff, V = e->V, Vpad = e->Vpad;
    const int R = B * K, T = max_len;
    // root child reached by the decoder start token (all reference paths start with 0)
    int root_child = -1;
    for (int k = trie->h_off[0]; k < trie->h_off[1]; ++k)
        if (trie->h_tok[k] == 0) root_child = trie->h_node[k];
    P5_CHECK(root_child >= 0, "trie paths must start with the decoder start token 0");
    const int cand_cap = K * (trie->max_fanout > 0 ? trie->max_fanout : 1);
    const bool persist = decode_persistent_supported(e, B, K, max_len, (int)round_up(Le_user, 8));
    GenWs* g = persist ? nullptr : get_gen_ws(e, R, T, K, B, cand_cap, (int)round_up(Le_user, 8));

    // ---- encoder once per user (eval mode) + cross K/V once per user
    e->set_geometry(B, Le_user, 1);
    e->pending_lens.clear();   // generate() always uses the padded encoder layout
    e->training = false; e->seed = 0;
    if (e->shadow_stale) e->refresh_shadow();
    e->load_inputs(ids, mask, ww, nullptr);
    e->encoder_forward();
    e->have_fwd = false;   // training activations are not valid for backward any more
    const int Le = e->Le;
    DropCfg none;
    if (e->batched_ckv) {
        e->project_cross_kv_all(e->Me);
    } else {
        for (int l = 0; l < e->ND; ++l)
            e->linear_fwd(e->enc_out, d, e->dec[l].ca.k, 2 * A, d, (int)e->Me, e->ckv[l], dt, e->ckv_ld, 0, 1.f, nullptr, nullptr, none);
    }
    e->build_bias(false, T);   // decoder relative bias for positions 0..T-1: [H, 2T-1], offset T-1
    const int n_delta = 2 * T - 1, bias_off = T - 1;
    if (persist) {
        // forced prefix: while the node reached so far has exactly one child (and it is not EOS) every beam of every user
        // must take that token ("ML1M item_" ... of the reference's item strings): known before decoding starts
        std::vector<int> forced;
        int node_f = root_child;
        while ((int)forced.size() < 31 && trie->h_off[node_f + 1] - trie->h_off[node_f] == 1 && trie->h_tok[trie->h_off[node_f]] != 1) {
            forced.push_back(trie->h_tok[trie->h_off[node_f]]);
            node_f = trie->h_node[trie->h_off[node_f]];
        }
        // ONE cooperative launch runs every decode position (decode_persist.cu)
        const int* out_len_dev = generate_persistent(e, trie->d_off, trie->d_tok, trie->d_node, root_child, trie->max_depth,
                                                     trie->max_fanout, B, K, Rret, max_len, length_penalty, seqs, scores,
                                                     forced.data(), (int)forced.size(), node_f);
        if (out_len_host) {
            P5_CUDA(cudaMemcpyAsync(out_len_host, out_len_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
            P5_CUDA(cudaStreamSynchronize(st));
        }
        return;
    }

    launch_k(gen_init_kernel, (unsigned)cdiv(R, 128), 128, 0, st, g->seq[0], g->fin_seq[0], g->src[0], g->node[0], g->run_score[0],
                                                           g->fin_score[0], g->is_fin[0], g->gen_len[0], g->cur_tok, g->unsat,
                                                           B, K, T, root_child);
    LAUNCHED();
    // y += X * W^T for the decode rows.  No dropout at inference, so the residual add can be a split-K atomic
    // accumulation straight into the fp32 residual stream: R ~ 400 rows give only 48 output tiles, splitting the
    // reduction spreads each projection over >= 140 CTAs instead of streaming K serially in 48.
    auto resid_gemm = [&](const void* X, int64_t ldx, int64_t w_off, int N, int K) {
        GemmProblem p;
        p.M = R; p.N = N;
        const int kb = (int)cdiv(K, 64);
        const int64_t tiles = cdiv(R, 128) * cdiv(N, 64);
        int splits = 1;
        for (int s2 = 1; s2 <= kb; ++s2)
            if (kb % s2 == 0 && (K % 64 == 0) && !e->x3) { splits = s2; if (tiles * s2 >= 140) break; }
        const int Ks = K / splits;
        p.K = Ks; p.nb1 = splits; p.prefer_bn = 64;
        p.A.ptr = X; p.A.dtype = dt; p.A.major = MAJOR_K; p.A.ld = ldx; p.A.bs1 = Ks;
        p.B.ptr = e->W(w_off); p.B.dtype = dt; p.B.major = MAJOR_K; p.B.ld = K; p.B.bs1 = Ks;
        p.epi.C = g->y; p.epi.c_dtype = DT_F32; p.epi.ldc = d; p.epi.cs1 = 0; p.epi.alpha = 1.f; p.epi.flags = EPI_ATOMIC;
        e->gemm(p);
    };
    const int n_steps = std::min(max_len - 1, trie->max_depth - 1);
    // A trie shallower than max_length: every path has ended by step n_steps, a beam that sits on a leaf without EOS has no
    // allowed continuation (transformers 4.26: an all -inf row; 5.x raises).  The last executed step is then the
    // MaxLengthCriteria step, so such hypotheses are finalised instead of being dropped with a -1e9 score.
    const int max_len_eff = std::min(max_len, n_steps + 1);
    int cur = 0;
    const float hs = 1.f / sqrtf((float)d);
    for (int step = 0; step < n_steps; ++step) {
        const int cur_len = step + 1, pos = step;
        // ---- one decoder step over R rows
        embed_fwd(e->P + e->off_shared, nullptr, g->cur_tok, nullptr, g->y, R, d, V, e->cfg.whole_word_rows, none, st);
        for (int l = 0; l < e->ND; ++l) {
            const DecLayerOff& w = e->dec[l];
            rmsnorm_fwd(g->y, e->P + w.ln0, g->n, dt, nullptr, R, d, e->cfg.ln_eps, none, st);
            e->linear_fwd(g->n, d, w.sa.q, 3 * A, d, R, g->qkv, dt, 3 * A, 0, 1.f, nullptr, nullptr, none);
            const unsigned sa_grid = (unsigned)cdiv((int64_t)R * H, 4);
            if (dt == DT_F32)
                launch_k(decode_self_attn_kernel<float>, sa_grid, 128, 0, st, (const float*)g->qkv, (float*)g->Kc[l], (float*)g->Vc[l],
                         g->src[cur], e->bias_dec, n_delta, bias_off, (float*)g->ctx, A, T, pos, R, H);
            else
                launch_k(decode_self_attn_kernel<bf16>, sa_grid, 128, 0, st, (const bf16*)g->qkv, (bf16*)g->Kc[l], (bf16*)g->Vc[l],
                         g->src[cur], e->bias_dec, n_delta, bias_off, (bf16*)g->ctx, A, T, pos, R, H);
            LAUNCHED();
            resid_gemm(g->ctx, A, w.sa.o, d, A);
            rmsnorm_fwd(g->y, e->P + w.ln1, g->n, dt, nullptr, R, d, e->cfg.ln_eps, none, st);
            e->linear_fwd(g->n, d, w.ca.q, A, d, R, g->cq, dt, A, 0, 1.f, nullptr, nullptr, none);
            AttnArgs xa;   // the K beams of a user are the query rows against that user's cross K/V
            xa.B = B; xa.H = H; xa.Lq = K; xa.Lk = Le;
            xa.q = {g->cq, dt, A, (int64_t)K * A};
            xa.k = {e->ckv[l], dt, e->ckv_ld, (int64_t)Le * e->ckv_ld};
            xa.v = {(const char*)e->ckv[l] + (size_t)A * e->esz(), dt, e->ckv_ld, (int64_t)Le * e->ckv_ld};
            xa.bias_rel = nullptr; xa.bias_off = 0; xa.n_delta = 0; xa.key_mask = e->mask_e; xa.causal = 0; xa.q_pos_offset = 0;
            xa.row_map = nullptr;
            if (dt == DT_BF16 && dattn_infer_supported(xa)) {
                // one mma.sync kernel per layer: S, softmax and P V of a (user, head) pair stay in registers
                dattn_fwd(xa, g->ctx, A, (int64_t)K * A, nullptr, st);
            } else if (dt == DT_BF16) {
                // cross-attention on the tensor cores: the K beams of a user are the M rows of one batched GEMM per
                // (user, head) against that user's cross K/V (stored once per user): S = Q K^T -> softmax -> P V
                const int64_t KL = (int64_t)K * Le;
                GemmProblem sp;
                sp.M = K; sp.N = Le; sp.K = 64; sp.nb1 = H; sp.nb2 = B;
                sp.A.ptr = g->cq; sp.A.dtype = dt; sp.A.major = MAJOR_K; sp.A.ld = A; sp.A.bs1 = 64; sp.A.bs2 = (int64_t)K * A;
                sp.B.ptr = e->ckv[l]; sp.B.dtype = dt; sp.B.major = MAJOR_K; sp.B.ld = e->ckv_ld; sp.B.bs1 = 64; sp.B.bs2 = (int64_t)Le * e->ckv_ld;
                sp.epi.C = g->xS; sp.epi.c_dtype = DT_F32; sp.epi.ldc = Le; sp.epi.cs1 = KL; sp.epi.cs2 = KL * H;
                e->gemm(sp);
                softmax_fwd(g->xS, nullptr, e->mask_e, g->xP, nullptr, dt, B, H, K, Le, 0, none, st);
                GemmProblem pv;
                pv.M = K; pv.N = 64; pv.K = Le; pv.nb1 = H; pv.nb2 = B;
                pv.A.ptr = g->xP; pv.A.dtype = dt; pv.A.major = MAJOR_K; pv.A.ld = Le; pv.A.bs1 = KL; pv.A.bs2 = KL * H;
                pv.B.ptr = (const char*)e->ckv[l] + (size_t)A * e->esz(); pv.B.dtype = dt; pv.B.major = MAJOR_MN; pv.B.ld = e->ckv_ld;
                pv.B.bs1 = 64; pv.B.bs2 = (int64_t)Le * e->ckv_ld;
                pv.epi.C = g->ctx; pv.epi.c_dtype = dt; pv.epi.ldc = A; pv.epi.cs1 = 64; pv.epi.cs2 = (int64_t)K * A;
                e->gemm(pv);
            } else {
                attn_simt_fwd(xa, g->ctx, dt, A, (int64_t)K * A, nullptr, st);
            }
            resid_gemm(g->ctx, A, w.ca.o, d, A);
            rmsnorm_fwd(g->y, e->P + w.ln2, g->n, dt, nullptr, R, d, e->cfg.ln_eps, none, st);
            if (!e->gated) {
                e->linear_fwd(g->n, d, w.ff.wi, ff, d, R, g->h, dt, ff, EPI_RELU, 1.f, nullptr, nullptr, none);
            } else {
                e->linear_fwd(g->n, d, w.ff.wi, 2 * ff, d, R, g->z, dt, 2 * ff, 0, 1.f, nullptr, nullptr, none);
                gated_gelu_fwd(g->z, g->h, dt, R, ff, none, st);
            }
            resid_gemm(g->h, ff, w.ff.wo, d, ff);
        }
        rmsnorm_fwd(g->y, e->P + e->off_dec_final, g->n, dt, nullptr, R, d, e->cfg.ln_eps, none, st);
        e->linear_fwd(g->n, d, e->off_shared, V, d, R, g->logits, DT_F32, Vpad, 0, hs, nullptr, nullptr, none);
        // ---- log-softmax normaliser, constrained top-2K, beam bookkeeping
        launch_k(rows_logsumexp_kernel, R, 512, 0, st, g->logits, Vpad, V, g->rowmax, g->logsum);
        LAUNCHED();
        launch_k(topk_candidates_kernel, B, 256, 0, st, g->logits, Vpad, V, g->rowmax, g->logsum, g->run_score[cur], g->node[cur],
                                                 trie->d_off, trie->d_tok, K, g->scr_score, g->scr_flat, g->scr_cap,
                                                 g->cand_lp, g->cand_beam, g->cand_tok);
        LAUNCHED();
        const int nxt = cur ^ 1;
        const float denom_fin = (float)pow((double)(cur_len + 1 - 1), (double)length_penalty);
        const float denom_next = (float)pow((double)(cur_len + 1 - 1), (double)length_penalty);
        launch_k(beam_update_kernel, B, 128, (14 * K + 4) * sizeof(int), st, g->cand_lp, g->cand_beam, g->cand_tok, g->seq[cur], g->seq[nxt], g->fin_seq[cur], g->fin_seq[nxt], g->src[cur],
            g->src[nxt], g->node[cur], g->node[nxt], g->run_score[nxt], g->fin_score[cur], g->fin_score[nxt], g->is_fin[cur],
            g->is_fin[nxt], g->gen_len[cur], g->gen_len[nxt], g->cur_tok, g->unsat, trie->d_off, trie->d_tok, trie->d_node, K,
            T, cur_len, max_len_eff, 1 /*eos*/, denom_fin, denom_next);
        LAUNCHED();
        cur = nxt;
    }
    launch_k(gen_finalize_kernel, 1, 256, 0, st, g->fin_seq[cur], g->fin_score[cur], g->is_fin[cur], g->gen_len[cur], B, K, Rret, T,
                                          max_len, seqs, scores, g->out_len);
    LAUNCHED();
    if (out_len_host) {
        P5_CUDA(cudaMemcpyAsync(out_len_host, g->out_len, sizeof(int), cudaMemcpyDeviceToHost, st));
        P5_CUDA(cudaStreamSynchronize(st));
    }
}

}  // namespace p5
