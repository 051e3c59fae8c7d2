// decode_persist.cu — trie-constrained beam search as ONE persistent kernel.
//
// Replaces the per-position kernel chain of beam.cu (~150 dependent launches per position, 1200 per eval batch) for
//   model.generate(num_beams=K, num_return_sequences=R, prefix_allowed_tokens_fn=trie)        (ref DistributedRunner.py:361-371)
//   HF:generation/utils.py:3231-3378 (_beam_search step), HF:generation/logits_process.py:1532-1549, utils/generation_trie.py
// One cooperative launch per generate() call (after the encoder and the per-user cross K|V projections): a grid of one
// CTA per SM runs EVERY decode position — token embedding, 12 decoder blocks (RMSNorm, self-attention with KV append and
// row indirection, cross-attention over the user's K|V, ReLU FFN), tied LM head, log-softmax normaliser, trie scoring,
// top-2K, HF beam bookkeeping, KV re-ordering by index — as phases separated by grid barriers.  No host round trip, no
// launch latency, weights stream from HBM/L2 straight into the mma.sync tiles of the phase that needs them.
//
// Design points
//   * RMSNorm never runs as its own pass: the phase that FINISHES a residual row writes y (fp32), y16 = bf16(y * ln_next)
//     (the next GEMM's A operand) and adds the row's sum of squares into rowss[site]; the consuming GEMM multiplies its
//     accumulator rows by rstd = rsqrt(rowss / d + eps) in the epilogue.   (HF:modeling_t5.py:55-70)
//   * the LM-head epilogue emits per-tile (max, sum exp) pairs, so log_softmax costs no extra pass over the 51 MB logits.
//   * DUPLICATE beams are computed once: running beams of a user with the same (parent representative, token) have the
//     same decoder state; only the first (`rep`) row is live, the others alias its logits and KV rows.  HF's K-1 initial
//     "-1e9" beams are exactly such duplicates through the forced item prefix, so the first positions run on one row per
//     user instead of K.  Bitwise neutral: a row's values do not depend on which other rows are computed.
//   * the KV cache is never moved: every row keeps, per position, the index of the row that owns that position's K/V
//     (`src`), re-pointed by the beam update (HF reorders the cache tensors, :3346-3352).
// bf16 operands / fp32 accumulation and statistics, like the precision=1 engine path; the fp32 and bf16x3 parity modes
// keep the per-position kernels of beam.cu.
#include "engine.h"
// #define P5_CA_STAMPS     // sub-phase timers of the two attention phases (tools/decode_phases.py with P5_DECODE_PROF_FINE=1)
#include "dattn_dev.cuh"
#include <cuda_runtime.h>
#include <algorithm>
#include <string.h>
#include <vector>

namespace p5 {
namespace {

constexpr int PD_THREADS = 256;
constexpr int PD_MAXL = 24;        // decoder layers
constexpr int PD_MAXK = 32;        // beams per user (two m16 query tiles of the cross-attention)
constexpr float PD_NEG_BIG = -1.0e9f;

// ---- GEMM tile (heavy): 128 rows x 64 columns, k-chunks of 64 through a cp.async ring
constexpr int G_TM = 128, G_TN = 64, G_KC = 64, G_STAGES = 4, G_LDS = 72;
constexpr int G_A_BYTES = G_TM * G_LDS * 2, G_B_BYTES = G_TN * G_LDS * 2, G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_RING_BYTES = G_STAGES * G_STAGE_BYTES;

// ---- GEMM tile (tcgen05, positions with > 32 live rows): 128 rows x 64 columns per tile, TMA -> 128B-swizzled smem ring ->
//      tcgen05.mma 128x64x16 into two TMEM accumulators -> 4 epilogue warps (thread = row)
constexpr int T_TM = 128, T_TN = 64, T_KB = 64, T_STAGES = 6;
constexpr int T_A_BYTES = T_TM * T_KB * 2, T_B_BYTES = T_TN * T_KB * 2, T_STAGE_BYTES = T_A_BYTES + T_B_BYTES;
constexpr int T_RING_BYTES = T_STAGES * T_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
// tensor-map table (device array): per layer 6 weight maps, then the LM head, then the three activation operands
enum { TM_QKV = 0, TM_O, TM_CQ, TM_CO, TM_WI, TM_WO, TM_PER_LAYER };

struct TcState {        // pipeline positions that persist across phases (each role keeps its own copy in registers)
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
};

struct PdLayer {
    const bf16 *wqkv, *wo, *wcq, *wco, *wi, *wwo;      // bf16 shadows, [N, K] row-major
    const float *ln0, *ln1, *ln2;                       // RMSNorm weights (fp32)
    const bf16 *ck, *cv;                                // cross K / V of this layer: row (b * Le + j), stride ckv_ld
    bf16 *Kc, *Vc;                                      // self K / V cache [R, T, A]
};

struct PdParams {
    int B, K, R, T, Le, d, A, H, ff, V, Vpad, ND;
    int n_steps, max_len, max_len_eff, n_ret, root_child, no_light, no_prefetch, prof_fine;
    int n_forced, node_forced, forced[32];     // forced item prefix (the trie has ONE child per node there): prefilled in one pass
    float eps, hs, length_penalty;
    PdLayer layer[PD_MAXL];
    const float* E;            // shared.weight fp32 [V, d] (embedding lookups)
    const bf16* E16;           // bf16 shadow (LM head)
    const float* ln_final;
    int64_t ckv_ld;
    const int* mask_e;         // [B, Le]
    const float* bias_dec; int n_delta, bias_off;
    // activations (indexed by ORIGINAL row r = b * K + k)
    float* y; bf16* y16; float* rowss;                  // rowss [3 * ND + 1][R]
    bf16 *qkv, *ctx, *cq, *h;
    float* logits; float2* lse_part; int n_ct_head;
    // beam state (double buffered by step parity)
    int *seq[2], *fin_seq[2], *src[2], *node[2], *rep[2], *is_fin[2], *gen_len[2];
    float *run_score[2], *fin_score[2];
    int *cur_tok, *unsat, *live_u, *n_u;
    float* scr_score; int* scr_flat; int scr_cap;
    const int *t_off, *t_tok, *t_node;
    int* tile_cnt;             // arrival counters of the split-K residual tiles (zero between phases)
    const CUtensorMap* tmaps;  // [6 * ND + 1 + 3]: weights per layer, LM head, then A operands y16 / ctx / h
    int no_tc;
    unsigned* bar;
    unsigned long long* prof;      // [2][16] ns per phase kind (CTA 0's view, phase + its barrier), light / heavy positions
    int32_t* out_seqs; float* out_scores; int* out_len;
};
enum PdPhase { PH_QKV = 0, PH_SA, PH_O, PH_CQ, PH_CA, PH_CO, PH_WI, PH_WO, PH_LM, PH_USER, PH_LIST, PH_N };

// ------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(uint32_t saddr, const void* g, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr));
}
__device__ __forceinline__ uint64_t pd_timer() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float2 ldcg_bf2(const bf16* p) {
    const unsigned u = __ldcg(reinterpret_cast<const unsigned*>(p));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ void st_bf2(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// Grid-wide barrier: one arrival per CTA on a monotonically increasing counter (zeroed by the host before the launch).
// A mis-scheduled grid traps after ~4 s instead of hanging the GPU (cooperative launch guarantees co-residency).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        // release: everything this CTA wrote (ordered before by the bar.sync above) is visible to whoever observes the count
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        uint64_t t0 = 0;
        unsigned spins = 0;
        while (true) {
            // acquire: also invalidates this SM's L1, so the plain loads of the next phase see the other CTAs' writes
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if ((int)(v - target) >= 0) break;
            if ((++spins & 0xfff) == 0) {
                const uint64_t now = pd_timer();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) {
                    printf("p5: decode grid barrier timeout (block %d, counter %u, target %u)\n", blockIdx.x, v, target);
                    __trap();
                }
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// GEMM phase:  C[live rows, N] = A[live rows, K] . W[N, K]^T   with a fused epilogue
// ------------------------------------------------------------------------------------------------------------
struct GemmDesc {
    const bf16* A; int64_t lda; int K;
    const bf16* W; int N;
    int mode;                      // 0: bf16 out (row scale, ReLU)   1: residual stream (y +=, y16, rowss)   2: logits (+ lse partials)
    float alpha; const float* rowss_in; float inv_d, eps; int relu;
    bf16* out16; int64_t ldo;
    float* y; bf16* y16; const float* ln_next; float* rowss_out; int d;
    float* logits; int64_t ldl; float2* lse_part; int n_ct_total; int V;
};

__device__ __forceinline__ float row_scale(const GemmDesc& g, int r) {
    float s = g.alpha;
    if (g.rowss_in) s *= rsqrtf(__ldcg(g.rowss_in + r) * g.inv_d + g.eps);
    return s;
}

__device__ void gemm_phase(const GemmDesc& g, const int* __restrict__ live, int n_live, uint8_t* smem) {
    const int n_rt = (n_live + G_TM - 1) / G_TM, n_ct = (g.N + G_TN - 1) / G_TN;
    const int KC = g.K / G_KC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    for (int tile = blockIdx.x; tile < n_rt * n_ct; tile += gridDim.x) {
        const int ct = tile / n_rt, rt = tile - ct * n_rt;
        const int row0 = rt * G_TM, col0 = ct * G_TN;
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

        auto load_chunk = [&](int kc, int stage) {
            uint8_t* sa = smem + stage * G_STAGE_BYTES;
            uint8_t* sb = sa + G_A_BYTES;
#pragma unroll
            for (int p = threadIdx.x; p < G_TM * 8; p += PD_THREADS) {
                const int r = p >> 3, c = p & 7, li = row0 + r;
                const bool ok = li < n_live;
                const bf16* src = g.A + (int64_t)(ok ? live[li] : 0) * g.lda + kc * G_KC + c * 8;
                cp_async16_zfill(smem_u32(sa + (r * G_LDS + c * 8) * 2), src, ok ? 16 : 0);
            }
#pragma unroll
            for (int p = threadIdx.x; p < G_TN * 8; p += PD_THREADS) {
                const int r = p >> 3, c = p & 7, n = col0 + r;
                const bool ok = n < g.N;
                const bf16* src = g.W + (int64_t)(ok ? n : 0) * g.K + kc * G_KC + c * 8;
                cp_async16_zfill(smem_u32(sb + (r * G_LDS + c * 8) * 2), src, ok ? 16 : 0);
            }
        };
#pragma unroll
        for (int s = 0; s < G_STAGES - 1; ++s) {
            if (s < KC) load_chunk(s, s);
            cp_async_commit();
        }
        for (int kc = 0; kc < KC; ++kc) {
            cp_async_wait<G_STAGES - 2>();
            __syncthreads();
            const int nk = kc + G_STAGES - 1;
            if (nk < KC) load_chunk(nk, nk % G_STAGES);
            cp_async_commit();
            const uint32_t sa_u = smem_u32(smem + (kc % G_STAGES) * G_STAGE_BYTES), sb_u = sa_u + G_A_BYTES;
#pragma unroll
            for (int k16 = 0; k16 < G_KC / 16; ++k16) {
                uint32_t a0, a1, a2, a3;
                ldsm_x4(a0, a1, a2, a3, sa_u + ((16 * warp + (lane & 15)) * G_LDS + k16 * 16 + (lane >> 4) * 8) * 2);
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(b0, b1, b2, b3, sb_u + (((2 * jp + (lane >> 4)) * 8 + (lane & 7)) * G_LDS + k16 * 16 + ((lane >> 3) & 1) * 8) * 2);
                    mma16816(acc[2 * jp], a0, a1, a2, a3, b0, b1);
                    mma16816(acc[2 * jp + 1], a0, a1, a2, a3, b2, b3);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();     // every warp is done with the ring before the next tile's prologue refills it

        // ---------------- epilogue: thread holds rows (gq, gq + 8) of its warp's 16, columns 8 j + 2 t (+1)
        const int li_lo = row0 + 16 * warp + gq, li_hi = li_lo + 8;
        const bool ok_lo = li_lo < n_live, ok_hi = li_hi < n_live;
        const int r_lo = ok_lo ? live[li_lo] : 0, r_hi = ok_hi ? live[li_hi] : 0;
        if (g.mode == 0) {
            const float s_lo = ok_lo ? row_scale(g, r_lo) : 0.f, s_hi = ok_hi ? row_scale(g, r_hi) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = col0 + 8 * j + 2 * t;
                if (col >= g.N) continue;
                float v0 = acc[j][0] * s_lo, v1 = acc[j][1] * s_lo, v2 = acc[j][2] * s_hi, v3 = acc[j][3] * s_hi;
                if (g.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
                if (ok_lo) st_bf2(g.out16 + (int64_t)r_lo * g.ldo + col, v0, v1);
                if (ok_hi) st_bf2(g.out16 + (int64_t)r_hi * g.ldo + col, v2, v3);
            }
        } else if (g.mode == 1) {
            float ss_lo = 0.f, ss_hi = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = col0 + 8 * j + 2 * t;
                if (col >= g.N) continue;
                const float2 ln = *reinterpret_cast<const float2*>(g.ln_next + col);
                if (ok_lo) {
                    float2* yp = reinterpret_cast<float2*>(g.y + (int64_t)r_lo * g.d + col);
                    float2 v = __ldcg(yp);
                    v.x += acc[j][0]; v.y += acc[j][1];
                    *yp = v;
                    st_bf2(g.y16 + (int64_t)r_lo * g.d + col, v.x * ln.x, v.y * ln.y);
                    ss_lo += v.x * v.x + v.y * v.y;
                }
                if (ok_hi) {
                    float2* yp = reinterpret_cast<float2*>(g.y + (int64_t)r_hi * g.d + col);
                    float2 v = __ldcg(yp);
                    v.x += acc[j][2]; v.y += acc[j][3];
                    *yp = v;
                    st_bf2(g.y16 + (int64_t)r_hi * g.d + col, v.x * ln.x, v.y * ln.y);
                    ss_hi += v.x * v.x + v.y * v.y;
                }
            }
            ss_lo += __shfl_xor_sync(0xffffffffu, ss_lo, 1); ss_lo += __shfl_xor_sync(0xffffffffu, ss_lo, 2);
            ss_hi += __shfl_xor_sync(0xffffffffu, ss_hi, 1); ss_hi += __shfl_xor_sync(0xffffffffu, ss_hi, 2);
            if (t == 0) {
                if (ok_lo) atomicAdd(g.rowss_out + r_lo, ss_lo);
                if (ok_hi) atomicAdd(g.rowss_out + r_hi, ss_hi);
            }
        } else {
            const float s_lo = ok_lo ? row_scale(g, r_lo) : 0.f, s_hi = ok_hi ? row_scale(g, r_hi) : 0.f;
            float m_lo = -INFINITY, m_hi = -INFINITY;
            float v[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = col0 + 8 * j + 2 * t;
                v[j][0] = acc[j][0] * s_lo; v[j][1] = acc[j][1] * s_lo; v[j][2] = acc[j][2] * s_hi; v[j][3] = acc[j][3] * s_hi;
                const bool c0 = col < g.V, c1 = col + 1 < g.V;
                if (!c0) { v[j][0] = -INFINITY; v[j][2] = -INFINITY; }
                if (!c1) { v[j][1] = -INFINITY; v[j][3] = -INFINITY; }
                m_lo = fmaxf(m_lo, fmaxf(v[j][0], v[j][1]));
                m_hi = fmaxf(m_hi, fmaxf(v[j][2], v[j][3]));
                if (ok_lo) { if (c0) g.logits[(int64_t)r_lo * g.ldl + col] = v[j][0]; if (c1) g.logits[(int64_t)r_lo * g.ldl + col + 1] = v[j][1]; }
                if (ok_hi) { if (c0) g.logits[(int64_t)r_hi * g.ldl + col] = v[j][2]; if (c1) g.logits[(int64_t)r_hi * g.ldl + col + 1] = v[j][3]; }
            }
            m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1)); m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
            m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1)); m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
            float e_lo = 0.f, e_hi = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (m_lo > -INFINITY) e_lo += __expf(v[j][0] - m_lo) + __expf(v[j][1] - m_lo);
                if (m_hi > -INFINITY) e_hi += __expf(v[j][2] - m_hi) + __expf(v[j][3] - m_hi);
            }
            e_lo += __shfl_xor_sync(0xffffffffu, e_lo, 1); e_lo += __shfl_xor_sync(0xffffffffu, e_lo, 2);
            e_hi += __shfl_xor_sync(0xffffffffu, e_hi, 1); e_hi += __shfl_xor_sync(0xffffffffu, e_hi, 2);
            if (t == 0) {
                if (ok_lo) g.lse_part[(int64_t)r_lo * g.n_ct_total + ct] = make_float2(m_lo, e_lo);
                if (ok_hi) g.lse_part[(int64_t)r_hi * g.n_ct_total + ct] = make_float2(m_hi, e_hi);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// GEMM phase on the 5th-generation tensor cores (positions with more than 32 live rows: every beam row is computed).
//   warp 0 lane 0 : TMA producer   cp.async.bulk.tensor -> 128B-swizzled smem ring (6 stages x (128x64 A + 64x64 W))
//   warp 1 lane 0 : MMA issuer     tcgen05.mma.cta_group::1.kind::f16 128x64x16, two TMEM accumulators (2 x 64 columns)
//   warps 4-7     : epilogue       tcgen05.ld 32x32b: thread = output row, 64 fp32 columns in registers -> fused epilogue
// The pipeline barriers and their phase bits live for the whole kernel; `TcState` carries each role's position from one
// phase to the next.  Activations written by other CTAs in the previous phase were published by the grid barrier
// (release / acquire at gpu scope); the producer adds the generic -> async proxy fence before its first TMA read.
// ------------------------------------------------------------------------------------------------------------
__device__ void gemm_tc_phase(const GemmDesc& g, const CUtensorMap* tmA, const CUtensorMap* tmB, int rows, uint32_t ring, uint32_t bars,
                              uint32_t tmem_base, TcState& st, int* __restrict__ tile_cnt, int* s_flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (rows + T_TM - 1) / T_TM, n_tiles = (g.N + T_TN - 1) / T_TN, k_blocks = g.K / T_KB;
    const int total = m_tiles * n_tiles;
    // Residual phases (mode 1) have N = d_model: only a few dozen tiles for 148 SMs, each streaming the whole K extent
    // through ONE SM's L2 port.  They are split along K into `splits` work units per tile: every unit adds its partial
    // product into y with red.global; the unit that arrives last on the tile's counter finalises the rows of the tile
    // (y16 = bf16(y * ln_next), row sums of squares).  Measured r02 on the K = d_model phases (12 k-blocks): 11.8 us split
    // by 3 against 13.2 us unsplit; the unsplit non-residual phase of the same shape (cross-q) takes 8.4 us.
    int splits = 1;
    if (g.mode == 1) { splits = (int)gridDim.x / total; splits = splits < 1 ? 1 : (splits > k_blocks ? k_blocks : splits); if (splits > 4) splits = 4; }
    const int kbs = (k_blocks + splits - 1) / splits;
    const int units = total * splits;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (T_STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * T_STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * T_STAGES + 2 + s); };
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.global;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the ring was attention staging (generic proxy) a phase ago
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int t = u / splits, sp = u - t * splits;
                const int m_blk = t % m_tiles, n_blk = t / m_tiles;
                const int kb0 = sp * kbs, kb1 = min(k_blocks, kb0 + kbs);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(st.stage), st.phase ^ 1);
                    const uint32_t sa = ring + st.stage * T_STAGE_BYTES, sb = sa + T_A_BYTES;
                    mbar_expect_tx(full_bar(st.stage), T_STAGE_BYTES);
                    tma_load_4d(sa, tmA, full_bar(st.stage), kb * T_KB, m_blk * T_TM, 0, 0);
                    tma_load_4d(sb, tmB, full_bar(st.stage), kb * T_KB, n_blk * T_TN, 0, 0);
                    if (++st.stage == T_STAGES) { st.stage = 0; st.phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T_TN >> 3) << 17) | ((uint32_t)(T_TM >> 4) << 24);
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int sp = u % splits;
                const int kb0 = sp * kbs, kb1 = min(k_blocks, kb0 + kbs);
                mbar_wait(tempty_bar(st.acc), st.acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(st.acc * T_TN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(st.stage), st.phase);
                    tc_fence_after();
                    const uint32_t sa = ring + st.stage * T_STAGE_BYTES, sb = sa + T_A_BYTES;
#pragma unroll
                    for (int k = 0; k < T_KB / 16; ++k)
                        umma_bf16(tmem_d, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                                  (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_commit(empty_bar(st.stage));
                    if (++st.stage == T_STAGES) { st.stage = 0; st.phase ^= 1; }
                }
                umma_commit(tfull_bar(st.acc));
                if (++st.acc == 2) { st.acc = 0; st.acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int ew = warp & 3;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int t = u / splits;
            const int m_blk = t % m_tiles, n_blk = t / m_tiles;
            mbar_wait(tfull_bar(st.acc), st.acc_phase);
            tc_fence_after();
            uint32_t rr[64];
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(st.acc * T_TN);
            tmem_ld32(taddr, rr);
            tmem_ld32(taddr + 32, rr + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(st.acc));          // the accumulator may be overwritten
            if (++st.acc == 2) { st.acc = 0; st.acc_phase ^= 1; }
            const int r = m_blk * T_TM + ew * 32 + lane, col0 = n_blk * T_TN;
            const bool row_ok = r < rows;
            if (g.mode == 0) {
                if (!row_ok) continue;
                const float sc = row_scale(g, r);
                bf16* out = g.out16 + (int64_t)r * g.ldo + col0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { v[e] = __uint_as_float(rr[8 * c + e]) * sc; if (g.relu) v[e] = fmaxf(v[e], 0.f); }
                    if (col0 + 8 * c < g.N)
                        *reinterpret_cast<uint4*>(out + 8 * c) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
            } else if (g.mode == 1) {
                float* yp = g.y + (int64_t)(row_ok ? r : 0) * g.d + col0;
                bool finalize = true;
                if (splits > 1) {
                    if (row_ok) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (col0 + 4 * c < g.N)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yp + 4 * c), "f"(__uint_as_float(rr[4 * c])),
                                             "f"(__uint_as_float(rr[4 * c + 1])), "f"(__uint_as_float(rr[4 * c + 2])), "f"(__uint_as_float(rr[4 * c + 3])) : "memory");
                    }
                    __threadfence();
                    asm volatile("bar.sync 1, 128;" ::: "memory");                 // the 4 epilogue warps: every partial of this unit is out
                    if (warp == 4 && lane == 0) {
                        const int old = atomicAdd(tile_cnt + t, 1);
                        const int last = (old == splits - 1) ? 1 : 0;
                        if (last) { tile_cnt[t] = 0; __threadfence(); }
                        *s_flag = last;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    finalize = *s_flag != 0;
                    asm volatile("bar.sync 1, 128;" ::: "memory");                 // s_flag is rewritten by the next unit
#pragma unroll
                    for (int c = 0; c < 64; ++c) rr[c] = 0u;                      // the finaliser adds nothing more
                }
                if (!finalize || !row_ok) continue;
                bf16* y16 = g.y16 + (int64_t)r * g.d + col0;
                float ss = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (col0 + 8 * c >= g.N) continue;
                    float4 a = __ldcg(reinterpret_cast<const float4*>(yp + 8 * c)), b = __ldcg(reinterpret_cast<const float4*>(yp + 8 * c + 4));
                    const float4 l0 = *reinterpret_cast<const float4*>(g.ln_next + col0 + 8 * c), l1 = *reinterpret_cast<const float4*>(g.ln_next + col0 + 8 * c + 4);
                    a.x += __uint_as_float(rr[8 * c]); a.y += __uint_as_float(rr[8 * c + 1]); a.z += __uint_as_float(rr[8 * c + 2]); a.w += __uint_as_float(rr[8 * c + 3]);
                    b.x += __uint_as_float(rr[8 * c + 4]); b.y += __uint_as_float(rr[8 * c + 5]); b.z += __uint_as_float(rr[8 * c + 6]); b.w += __uint_as_float(rr[8 * c + 7]);
                    if (splits == 1) {
                        *reinterpret_cast<float4*>(yp + 8 * c) = a;
                        *reinterpret_cast<float4*>(yp + 8 * c + 4) = b;
                    }
                    *reinterpret_cast<uint4*>(y16 + 8 * c) = make_uint4(pack_bf16(a.x * l0.x, a.y * l0.y), pack_bf16(a.z * l0.z, a.w * l0.w),
                                                                        pack_bf16(b.x * l1.x, b.y * l1.y), pack_bf16(b.z * l1.z, b.w * l1.w));
                    ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
                }
                atomicAdd(g.rowss_out + r, ss);
            } else {
                if (!row_ok) continue;
                const float sc = row_scale(g, r);
                float* lp = g.logits + (int64_t)r * g.ldl + col0;
                float m = -INFINITY;
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    float v = __uint_as_float(rr[c]) * sc;
                    if (col0 + c >= g.V) v = -INFINITY;
                    rr[c] = __float_as_uint(v);
                    m = fmaxf(m, v);
                }
                float e = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float4 v = make_float4(__uint_as_float(rr[4 * c]), __uint_as_float(rr[4 * c + 1]), __uint_as_float(rr[4 * c + 2]), __uint_as_float(rr[4 * c + 3]));
                    if (col0 + 4 * c + 3 < g.V) *reinterpret_cast<float4*>(lp + 4 * c) = v;
                    else {
                        if (col0 + 4 * c < g.V) lp[4 * c] = v.x;
                        if (col0 + 4 * c + 1 < g.V) lp[4 * c + 1] = v.y;
                        if (col0 + 4 * c + 2 < g.V) lp[4 * c + 2] = v.z;
                    }
                    if (m > -INFINITY) e += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
                }
                g.lse_part[(int64_t)r * g.n_ct_total + n_blk] = make_float2(m, e);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// GEMM phase, LIGHT variant (<= 32 live rows: the forced item prefix, where every user has ONE distinct beam).  The phase is
// pure weight streaming, so it is organised for memory-level parallelism instead of reuse: a tile is (all live rows) x 16
// output columns, the 8 warps of the CTA split the reduction dimension (64-wide k-blocks round-robin), every warp loads its
// A / W fragments straight from global memory with 16-byte loads (permuted contraction index, as in dattn_dev.cuh) and the
// 8 partial accumulators are summed through shared memory before the same fused epilogues.
// ------------------------------------------------------------------------------------------------------------
constexpr int L_TN = 16;
__device__ void gemm_light(const GemmDesc& g, const int* __restrict__ live, int n_live, uint8_t* smem) {
    float* red = reinterpret_cast<float*>(smem);               // [8 warps][32 rows][16 cols]
    const int n_ct = (g.N + L_TN - 1) / L_TN, KB = g.K / 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    for (int tile = blockIdx.x; tile < n_ct; tile += gridDim.x) {
        const int col0 = tile * L_TN;
        float acc[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
        const bf16* arow[2][2];
        bool aok[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hi = 0; hi < 2; ++hi) {
                const int i = 16 * mt + 8 * hi + gq;
                aok[mt][hi] = i < n_live;
                arow[mt][hi] = g.A + (int64_t)(aok[mt][hi] ? live[i] : 0) * g.lda + 16 * t;
            }
        const bf16* brow[2];
        bool bok[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int n = col0 + 8 * nt + gq;
            bok[nt] = n < g.N;
            brow[nt] = g.W + (int64_t)(bok[nt] ? n : 0) * g.K + 16 * t;
        }
#pragma unroll 2
        for (int kb = warp; kb < KB; kb += PD_THREADS / 32) {
            uint32_t a[2][2][8], b[2][8];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) ld_row16(b[nt], brow[nt] + kb * 64, bok[nt]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hi = 0; hi < 2; ++hi) ld_row16(a[mt][hi], arow[mt][hi] + kb * 64, aok[mt][hi]);
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
                        mma16816(acc[mt][nt], a[mt][0][2 * s4], a[mt][1][2 * s4], a[mt][0][2 * s4 + 1], a[mt][1][2 * s4 + 1], b[nt][2 * s4],
                                 b[nt][2 * s4 + 1]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float* p = red + (warp * 32 + 16 * mt + gq) * L_TN + 8 * nt + 2 * t;
                *reinterpret_cast<float2*>(p) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<float2*>(p + 8 * L_TN) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
            }
        __syncthreads();
        // ---------------- reduce over the 8 warps + epilogue: thread = (row, column pair)
        const int row = threadIdx.x >> 3, cp = threadIdx.x & 7, col = col0 + 2 * cp;
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int w = 0; w < PD_THREADS / 32; ++w) {
            const float2 x = *reinterpret_cast<const float2*>(red + (w * 32 + row) * L_TN + 2 * cp);
            v0 += x.x; v1 += x.y;
        }
        const bool ok = row < n_live;
        const int r = ok ? live[row] : 0;
        if (g.mode == 0) {
            if (ok && col < g.N) {
                const float sc = row_scale(g, r);
                v0 *= sc; v1 *= sc;
                if (g.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                st_bf2(g.out16 + (int64_t)r * g.ldo + col, v0, v1);
            }
        } else if (g.mode == 1) {
            float ss = 0.f;
            if (ok && col < g.N) {
                const float2 ln = *reinterpret_cast<const float2*>(g.ln_next + col);
                float2* yp = reinterpret_cast<float2*>(g.y + (int64_t)r * g.d + col);
                float2 v = __ldcg(yp);
                v.x += v0; v.y += v1;
                *yp = v;
                st_bf2(g.y16 + (int64_t)r * g.d + col, v.x * ln.x, v.y * ln.y);
                ss = v.x * v.x + v.y * v.y;
            }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 4);
            if (cp == 0 && ok) atomicAdd(g.rowss_out + r, ss);
        } else {
            const float sc = ok ? row_scale(g, r) : 0.f;
            v0 *= sc; v1 *= sc;
            const bool c0 = col < g.V, c1 = col + 1 < g.V;
            if (!c0) v0 = -INFINITY;
            if (!c1) v1 = -INFINITY;
            if (ok) { if (c0) g.logits[(int64_t)r * g.ldl + col] = v0; if (c1) g.logits[(int64_t)r * g.ldl + col + 1] = v1; }
            float m = fmaxf(v0, v1);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            float e = (m > -INFINITY) ? __expf(v0 - m) + __expf(v1 - m) : 0.f;
            e += __shfl_xor_sync(0xffffffffu, e, 1); e += __shfl_xor_sync(0xffffffffu, e, 2); e += __shfl_xor_sync(0xffffffffu, e, 4);
            if (cp == 0 && ok) g.lse_part[(int64_t)r * g.n_ct_total + tile] = make_float2(m, e);
        }
        __syncthreads();     // `red` is rewritten by the next tile
    }
}

// ------------------------------------------------------------------------------------------------------------
// decoder self-attention of ONE new position per live row, with KV append and row indirection: warp per (row, head)
// (HF:modeling_t5.py:253-344 with the decoder's unidirectional relative bias; unscaled scores)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld16_bf(float (&f)[16], const bf16* p) {      // 16 consecutive bf16 (L2 path) -> fp32
    const uint4 x = __ldcg(reinterpret_cast<const uint4*>(p)), y = __ldcg(reinterpret_cast<const uint4*>(p + 8));
    const uint32_t u[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[i]));
        f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
}

// warp per (live row, head).  lane = (position group jg = lane / 4, dim quarter dq = lane % 4): the <= 8 cached positions of
// the usual item depth are scored IN PARALLEL (one round of independent loads instead of a dependent chain per position);
// longer prefixes loop in blocks of 8 with an online softmax per lane, merged across the groups with shuffles at the end.
__device__ void self_attn_phase(const PdParams& P, const PdLayer& L, const int* __restrict__ live, int n_live, const int* __restrict__ src,
                                int pos) {
    // Two (row, head) tasks per warp and iteration: their index -> K | V load chains are independent, so the second one rides
    // in the latency shadow of the first (the phase is a handful of dependent L2 round trips per task, not bandwidth).
    constexpr int U = 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, jg = lane >> 2, dq = lane & 3;
    const int A = P.A, T = P.T, H = P.H, n_tasks = n_live * H, n_warps = gridDim.x * (PD_THREADS / 32);
    for (int task0 = blockIdx.x * (PD_THREADS / 32) + warp; task0 < n_tasks; task0 += U * n_warps) {
#ifdef P5_CA_STAMPS
        const bool stamp = blockIdx.x == 0 && threadIdx.x == 0 && P.prof_fine;
        uint64_t ts0 = 0, ts1 = 0;
        if (stamp) ts0 = pd_timer();
#endif
        int r[U], h[U];
        bool has[U];
        const bf16* row[U];
        float q[U][16], m[U], l[U], acc[U][16];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int task = task0 + u * n_warps;
            has[u] = task < n_tasks;                               // warp-uniform
            const int tk = has[u] ? task : task0;
            const int i = tk / H;
            h[u] = tk - i * H; r[u] = live[i];
            row[u] = P.qkv + (int64_t)r[u] * 3 * A + h[u] * 64 + 16 * dq;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ld16_bf(q[u], row[u]);
            m[u] = -INFINITY; l[u] = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[u][c] = 0.f;
        }
        for (int j0 = 0; j0 <= pos; j0 += 8) {           // warp-uniform trip count: the shuffles below need every lane
            const int j = j0 + jg;
            const bool valid = j <= pos;
            const bf16 *kp[U], *vp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                kp[u] = row[u] + A; vp[u] = row[u] + 2 * A;
                if (valid && j < pos) {
                    const int64_t o = ((int64_t)__ldcg(src + r[u] * T + j) * T + j) * A + h[u] * 64 + 16 * dq;
                    kp[u] = L.Kc + o; vp[u] = L.Vc + o;
                }
            }
            float k[U][16], v[U][16];
#pragma unroll
            for (int u = 0; u < U; ++u) { ld16_bf(k[u], kp[u]); ld16_bf(v[u], vp[u]); }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float sc = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) sc = fmaf(q[u][c], k[u][c], sc);
                sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                sc += __shfl_xor_sync(0xffffffffu, sc, 2);
                if (valid) {
                    int di = j - pos + P.bias_off;
                    di = di < 0 ? 0 : (di >= P.n_delta ? P.n_delta - 1 : di);
                    sc += P.bias_dec[h[u] * P.n_delta + di];
                    const float mn = fmaxf(m[u], sc);
                    const float scale = __expf(m[u] - mn), p = __expf(sc - mn);
                    l[u] = l[u] * scale + p;
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[u][c] = fmaf(acc[u][c], scale, p * v[u][c]);
                    m[u] = mn;
                }
            }
        }
#ifdef P5_CA_STAMPS
        if (stamp) { ts1 = pd_timer(); P.prof[28] += ts1 - ts0; }
#endif
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // merge the 8 position groups (lanes differing in bits 2..4 hold the same dims)
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, m[u], o), ol = __shfl_xor_sync(0xffffffffu, l[u], o);
                const float mn = fmaxf(m[u], om);
                const float s0 = (m[u] > -INFINITY) ? __expf(m[u] - mn) : 0.f, s1 = (om > -INFINITY) ? __expf(om - mn) : 0.f;
                l[u] = l[u] * s0 + ol * s1;
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[u][c] = acc[u][c] * s0 + __shfl_xor_sync(0xffffffffu, acc[u][c], o) * s1;
                m[u] = mn;
            }
            if (jg == 1 && has[u]) {      // KV append: this row owns position `pos` (after the loop: the cache loads above
                                          // must not queue behind these stores)
                const int64_t o = ((int64_t)r[u] * T + pos) * A + h[u] * 64 + 16 * dq;
                *reinterpret_cast<uint4*>(L.Kc + o) = __ldcg(reinterpret_cast<const uint4*>(row[u] + A));
                *reinterpret_cast<uint4*>(L.Kc + o + 8) = __ldcg(reinterpret_cast<const uint4*>(row[u] + A + 8));
                *reinterpret_cast<uint4*>(L.Vc + o) = __ldcg(reinterpret_cast<const uint4*>(row[u] + 2 * A));
                *reinterpret_cast<uint4*>(L.Vc + o + 8) = __ldcg(reinterpret_cast<const uint4*>(row[u] + 2 * A + 8));
            }
            if (jg == 0 && has[u]) {
                const float inv = 1.f / l[u];
                const float* a = acc[u];
                bf16* out = P.ctx + (int64_t)r[u] * A + h[u] * 64 + 16 * dq;
                uint4 w0, w1;
                w0.x = pack_bf16(a[0] * inv, a[1] * inv); w0.y = pack_bf16(a[2] * inv, a[3] * inv);
                w0.z = pack_bf16(a[4] * inv, a[5] * inv); w0.w = pack_bf16(a[6] * inv, a[7] * inv);
                w1.x = pack_bf16(a[8] * inv, a[9] * inv); w1.y = pack_bf16(a[10] * inv, a[11] * inv);
                w1.z = pack_bf16(a[12] * inv, a[13] * inv); w1.w = pack_bf16(a[14] * inv, a[15] * inv);
                *reinterpret_cast<uint4*>(out) = w0;
                *reinterpret_cast<uint4*>(out + 8) = w1;
            }
        }
#ifdef P5_CA_STAMPS
        if (stamp) P.prof[30] += pd_timer() - ts1;
#endif
    }
}

// Self-attention of the PREFILL pass: the forced item prefix (positions 0 .. p of every user, known before decoding starts
// because the trie has a single child per node there) is decoded in one pass like a training sequence.  Item (b, t) lives in
// row b * K + t; its keys / values for positions j <= t are the k / v projections of the sibling rows b * K + j of the same
// pass (causal), and its own k / v go to position t of the user's representative KV-cache row b * K.
__device__ void self_attn_prefill(const PdParams& P, const PdLayer& L, int p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, jg = lane >> 2, dq = lane & 3;
    const int A = P.A, T = P.T, H = P.H, K = P.K, n_items = P.B * (p + 1);
    for (int task = blockIdx.x * (PD_THREADS / 32) + warp; task < n_items * H; task += gridDim.x * (PD_THREADS / 32)) {
        const int it = task / H, h = task - it * H, b = it / (p + 1), t = it - b * (p + 1), r = b * K + t;
        const bf16* row = P.qkv + (int64_t)r * 3 * A + h * 64 + 16 * dq;
        float q[16];
        ld16_bf(q, row);
        if (jg == 0) {
            const int64_t o = ((int64_t)(b * K) * T + t) * A + h * 64 + 16 * dq;
            *reinterpret_cast<uint4*>(L.Kc + o) = __ldcg(reinterpret_cast<const uint4*>(row + A));
            *reinterpret_cast<uint4*>(L.Kc + o + 8) = __ldcg(reinterpret_cast<const uint4*>(row + A + 8));
            *reinterpret_cast<uint4*>(L.Vc + o) = __ldcg(reinterpret_cast<const uint4*>(row + 2 * A));
            *reinterpret_cast<uint4*>(L.Vc + o + 8) = __ldcg(reinterpret_cast<const uint4*>(row + 2 * A + 8));
        }
        float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.f;
        for (int j0 = 0; j0 <= t; j0 += 8) {
            const int j = j0 + jg;
            const bool valid = j <= t;
            const bf16* kv = P.qkv + (int64_t)(b * K + (valid ? j : t)) * 3 * A + h * 64 + 16 * dq;
            float k[16], v[16];
            ld16_bf(k, kv + A);
            ld16_bf(v, kv + 2 * A);
            float sc = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) sc = fmaf(q[c], k[c], sc);
            sc += __shfl_xor_sync(0xffffffffu, sc, 1);
            sc += __shfl_xor_sync(0xffffffffu, sc, 2);
            if (valid) {
                int di = j - t + P.bias_off;
                di = di < 0 ? 0 : (di >= P.n_delta ? P.n_delta - 1 : di);
                sc += P.bias_dec[h * P.n_delta + di];
                const float mn = fmaxf(m, sc);
                const float scale = __expf(m - mn), pr = __expf(sc - mn);
                l = l * scale + pr;
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[c] = fmaf(acc[c], scale, pr * v[c]);
                m = mn;
            }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m, o), ol = __shfl_xor_sync(0xffffffffu, l, o);
            const float mn = fmaxf(m, om);
            const float s0 = (m > -INFINITY) ? __expf(m - mn) : 0.f, s1 = (om > -INFINITY) ? __expf(om - mn) : 0.f;
            l = l * s0 + ol * s1;
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = acc[c] * s0 + __shfl_xor_sync(0xffffffffu, acc[c], o) * s1;
            m = mn;
        }
        if (jg == 0) {
            const float inv = 1.f / l;
            bf16* out = P.ctx + (int64_t)r * A + h * 64 + 16 * dq;
            uint4 w0, w1;
            w0.x = pack_bf16(acc[0] * inv, acc[1] * inv); w0.y = pack_bf16(acc[2] * inv, acc[3] * inv);
            w0.z = pack_bf16(acc[4] * inv, acc[5] * inv); w0.w = pack_bf16(acc[6] * inv, acc[7] * inv);
            w1.x = pack_bf16(acc[8] * inv, acc[9] * inv); w1.y = pack_bf16(acc[10] * inv, acc[11] * inv);
            w1.z = pack_bf16(acc[12] * inv, acc[13] * inv); w1.w = pack_bf16(acc[14] * inv, acc[15] * inv);
            *reinterpret_cast<uint4*>(out) = w0;
            *reinterpret_cast<uint4*>(out + 8) = w1;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// cross-attention: CTA per (user, head); the user's live beams are the query rows against that user's K | V
// ------------------------------------------------------------------------------------------------------------
template <int NT>
__device__ void cross_attn_phase(const PdParams& P, const PdLayer& L, uint8_t* smem) {
    const int B = P.B, H = P.H, K = P.K;
    for (int task = blockIdx.x; task < B * H; task += gridDim.x) {
        const int b = task / H, h = task - b * H;
        const int nq = __ldcg(P.n_u + b);
        if (nq <= 0) continue;           // uniform per CTA
        DAttnDev a;
        a.B = B; a.H = H; a.Lq = nq; a.Lk = P.Le;
        a.q = P.cq; a.k = L.ck; a.v = L.cv;
        a.q_ld = P.A; a.q_bs = 0; a.k_ld = P.ckv_ld; a.k_bs = (int64_t)P.Le * P.ckv_ld; a.v_ld = P.ckv_ld; a.v_bs = a.k_bs;
        a.bias_rel = nullptr; a.bias_off = 0; a.n_delta = 0; a.key_mask = P.mask_e; a.causal = 0;
        a.kv_off = nullptr; a.kv_len = nullptr;
        __syncthreads();                 // the previous task's partial sums / staged V rows are no longer read
        dattn_fwd32_body<NT, PD_THREADS / 32>(a, P.ctx, P.A, 0, b, h, smem, P.live_u + b * K);
    }
}

// Le <= 256: TWO (user, head) pairs per CTA, four warps each, on their own named barrier and shared-memory half: the
// B * H = 240 pairs of the BASELINE eval shape then fit the 148 CTAs in one round instead of two.
__device__ void cross_attn_phase_x2(const PdParams& P, const PdLayer& L, uint8_t* smem) {
    const int B = P.B, H = P.H, K = P.K;
    const int grp = threadIdx.x >> 7, tid = threadIdx.x & 127;
    constexpr int HALF = DCfg<8, 4>::TILE > 4 * 32 * 64 * 4 ? DCfg<8, 4>::TILE : 4 * 32 * 64 * 4;       // V rows | partial sums
    uint8_t* my = smem + grp * (HALF + 4 * 32 * 2 * 4 + 128 + DCfg<8, 4>::TILE);                         // ... | softmax scratch | K rows
    for (int task = 2 * blockIdx.x + grp; task < B * H; task += 2 * gridDim.x) {
        const int b = task / H, h = task - b * H;
        const int nq = __ldcg(P.n_u + b);
        if (nq <= 0) continue;           // uniform per 4-warp group
        DAttnDev a;
        a.B = B; a.H = H; a.Lq = nq; a.Lk = P.Le;
        a.q = P.cq; a.k = L.ck; a.v = L.cv;
        a.q_ld = P.A; a.q_bs = 0; a.k_ld = P.ckv_ld; a.k_bs = (int64_t)P.Le * P.ckv_ld; a.v_ld = P.ckv_ld; a.v_bs = a.k_bs;
        a.bias_rel = nullptr; a.bias_off = 0; a.n_delta = 0; a.key_mask = P.mask_e; a.causal = 0;
        a.kv_off = nullptr; a.kv_len = nullptr;
        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");     // this group's previous pair is done with its half
        dattn_fwd32_body<8, 4, true>(a, P.ctx, P.A, 0, b, h, my, P.live_u + b * K, tid, 2 + grp);
    }
}

// Cross-attention when every user has only a few distinct beams (the forced item prefix: ONE).  A 32-row mma tile per
// (user, head) would be 97 % padding and the phase is pure K | V streaming, so it is organised for memory parallelism:
// CTA per (user, head), THREAD per key (Le <= 256).  Every thread loads its key row and its share of the value rows up
// front (16-byte loads, all independent), dots the key with the query held in shared memory, the block softmaxes, and the
// P.V product is reduced through shared memory (thread = (8-key group, 8-dim chunk)).  Up to 8 query rows per user reuse
// the K / V registers.
__device__ void cross_attn_light(const PdParams& P, const PdLayer& L, uint8_t* smem) {
    const int B = P.B, H = P.H, K = P.K, Le = P.Le, A = P.A;
    float* s_q = reinterpret_cast<float*>(smem);                 // [8][64] queries
    float* s_p = s_q + 8 * 64;                                   // [256] probabilities of the current query
    float* s_red = s_p + 256;                                    // [32] block reductions
    float* s_o = s_red + 32;                                     // [32 key groups][64 dims] partial outputs
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid >> 3, ch = tid & 7;                      // P.V mapping: keys 8 grp .. 8 grp + 7, dims 8 ch .. 8 ch + 7
    for (int task = blockIdx.x; task < B * H; task += gridDim.x) {
        const int b = task / H, h = task - b * H;
        const int nq = __ldcg(P.n_u + b);
        if (nq <= 0) continue;
        __syncthreads();                                         // the previous task's shared arrays are no longer read
        const bf16* kb = L.ck + (int64_t)b * Le * P.ckv_ld + h * 64;
        const bf16* vb = L.cv + (int64_t)b * Le * P.ckv_ld + h * 64;
        // ---- issue every global load first: key row `tid`, value chunk of the 8 keys of `grp`
        const bool kvalid = tid < Le && P.mask_e[(int64_t)b * Le + tid] != 0;
        uint4 kr[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) kr[c] = kvalid ? reinterpret_cast<const uint4*>(kb + (int64_t)tid * P.ckv_ld)[c] : make_uint4(0, 0, 0, 0);
        uint4 vr[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = 8 * grp + jj;
            vr[jj] = j < Le ? reinterpret_cast<const uint4*>(vb + (int64_t)j * P.ckv_ld)[ch] : make_uint4(0, 0, 0, 0);
        }
        for (int e = tid; e < nq * 64; e += PD_THREADS) {
            const int i = e >> 6, c = e & 63;
            s_q[e] = __bfloat162float(P.cq[(int64_t)__ldcg(P.live_u + b * K + i) * A + h * 64 + c]);
        }
        __syncthreads();
        for (int i = 0; i < nq; ++i) {
            const int r = __ldcg(P.live_u + b * K + i);
            float sc = -INFINITY;
            if (kvalid) {
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t u[4] = {kr[c].x, kr[c].y, kr[c].z, kr[c].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[e]));
                        d = fmaf(s_q[i * 64 + 8 * c + 2 * e], kv.x, d);
                        d = fmaf(s_q[i * 64 + 8 * c + 2 * e + 1], kv.y, d);
                    }
                }
                sc = d;
            }
            float mx = warp_max(sc);
            if (lane == 0) s_red[warp] = mx;
            __syncthreads();
            mx = s_red[0];
#pragma unroll
            for (int w2 = 1; w2 < PD_THREADS / 32; ++w2) mx = fmaxf(mx, s_red[w2]);
            const float p = (sc > -INFINITY) ? __expf(sc - mx) : 0.f;
            float sum = warp_sum(p);
            if (lane == 0) s_red[8 + warp] = sum;
            __syncthreads();
            sum = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < PD_THREADS / 32; ++w2) sum += s_red[8 + w2];
            const float inv = sum > 0.f ? 1.f / sum : 0.f;
            s_p[tid] = __bfloat162float(__float2bfloat16_rn(p * inv));        // P is a bf16 operand of P.V in the tiled kernels
            __syncthreads();
            float acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float pj = s_p[8 * grp + jj];
                const uint32_t u[4] = {vr[jj].x, vr[jj].y, vr[jj].z, vr[jj].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[e]));
                    acc[2 * e] = fmaf(pj, vv.x, acc[2 * e]);
                    acc[2 * e + 1] = fmaf(pj, vv.y, acc[2 * e + 1]);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) s_o[grp * 64 + 8 * ch + c] = acc[c];
            __syncthreads();
            if (tid < 64) {
                float o = 0.f;
#pragma unroll
                for (int g2 = 0; g2 < 32; ++g2) o += s_o[g2 * 64 + tid];
                P.ctx[(int64_t)r * A + h * 64 + tid] = __float2bfloat16_rn(o);
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// per user: log-softmax normaliser from the LM-head partials, trie scoring + top-2K, HF beam bookkeeping, duplicate
// detection, embedding of the next input token
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pd_trie_child(const int* off, const int* tok, const int* node, int n, int t) {
    if (n < 0) return -1;
    int lo = off[n], hi = off[n + 1] - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int v = tok[mid];
        if (v == t) return node[mid];
        if (v < t) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

struct UserSmem {
    float rowmax[PD_MAXK], logsum[PD_MAXK];
    int s_off[PD_MAXK + 1], s_nd[PD_MAXK], s_rep[PD_MAXK], s_cnt[PD_MAXK];
    float s_rs[PD_MAXK];
    float s_sc[1024]; int s_fl[1024];
    float s_best[8]; int s_besti[8], s_bestf[8];
    int run_sel[PD_MAXK], fin_sel[PD_MAXK], s_cb[2 * PD_MAXK], s_ct[2 * PD_MAXK], s_fin[PD_MAXK], rep_new[PD_MAXK], node_new[PD_MAXK];
    float s_lp[2 * PD_MAXK], s_rv[2 * PD_MAXK], s_fv[3 * PD_MAXK], s_misc[2];
    int n_new;
};

// embed the input token of row r (decoder input = shared embedding, HF T5Stack) and prime layer 0's norm: y, y16, rowss[0]
__device__ __forceinline__ void embed_row(const PdParams& P, int r, int tok, int lane) {
    tok = tok < 0 ? 0 : (tok >= P.V ? P.V - 1 : tok);
    const float* e = P.E + (int64_t)tok * P.d;
    const float* ln = P.layer[0].ln0;
    float ss = 0.f;
    for (int c = 4 * lane; c < P.d; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(e + c), w = *reinterpret_cast<const float4*>(ln + c);
        *reinterpret_cast<float4*>(P.y + (int64_t)r * P.d + c) = v;
        *reinterpret_cast<uint2*>(P.y16 + (int64_t)r * P.d + c) = make_uint2(pack_bf16(v.x * w.x, v.y * w.y), pack_bf16(v.z * w.z, v.w * w.w));
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const int n_sites = 3 * P.ND + 1;
    for (int s = lane; s < n_sites; s += 32) P.rowss[(int64_t)s * P.R + r] = (s == 0) ? ss : 0.f;
}

// log-softmax normaliser of one row from the LM-head tile partials (warp): S.rowmax / S.logsum [slot]
__device__ __forceinline__ void combine_lse(const PdParams& P, UserSmem& S, int r, int slot, int n_parts, int lane) {
    const float2* part = P.lse_part + (int64_t)r * n_parts;
    float m = -INFINITY, s = 0.f;
    for (int c = lane; c < n_parts; c += 32) {
        const float2 p = __ldcg(part + c);
        if (p.x > m) { s = s * __expf(m - p.x) + p.y; m = p.x; }
        else if (p.x > -INFINITY) s += p.y * __expf(p.x - m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
        const float M = fmaxf(m, om);
        s = (m > -INFINITY ? s * __expf(m - M) : 0.f) + (om > -INFINITY ? os * __expf(om - M) : 0.f);
        m = M;
    }
    if (lane == 0) { S.rowmax[slot] = m; S.logsum[slot] = logf(s); }
}

// After the prefill pass: the beam state HF's loop has after the p forced positions.  Beam 0 carries the accumulated
// log-probability of the forced tokens (added position by position in fp32, as the loop does); the K-1 other beams are the
// "-1e9" duplicates of beam 0 (HF:generation/utils.py:3200-3215: -1e9 plus a log-probability rounds back to -1e9 in fp32);
// every beam's trie node is the node after the prefix, its KV rows are the representative row b*K, and the logits of the
// next position sit in row b*K + p.
__device__ void prefill_user_init(const PdParams& P, UserSmem& S, int b, int cur, int p, int n_parts) {
    const int K = P.K, T = P.T, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = warp; t <= p; t += PD_THREADS / 32) combine_lse(P, S, b * K + t, t, n_parts, lane);
    __syncthreads();
    if (threadIdx.x == 0) {
        float sc = 0.f;
        for (int t = 0; t < p; ++t) {
            const float lp = (__ldcg(P.logits + (int64_t)(b * K + t) * P.Vpad + P.forced[t]) - S.rowmax[t]) - S.logsum[t];
            sc = lp + sc;
        }
        S.s_misc[0] = sc;
        P.live_u[b * K] = b * K + p;
        P.n_u[b] = 1;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += PD_THREADS) {
        const int r = b * K + k;
        P.run_score[cur][r] = (k == 0) ? S.s_misc[0] : PD_NEG_BIG;
        P.node[cur][r] = P.node_forced;
        P.rep[cur][r] = p;
    }
    for (int idx = threadIdx.x; idx < K * T; idx += PD_THREADS) {
        const int t = idx % T;
        P.seq[cur][b * K * T + idx] = (t >= 1 && t <= p) ? P.forced[t - 1] : 0;
    }
    __syncthreads();
}

__device__ void user_phase(const PdParams& P, UserSmem& S, int b, int cur, int cur_len, int step, int n_parts) {
    const int K = P.K, T = P.T, V = P.V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nxt = cur ^ 1;
    const int* node_in = P.node[cur];
    const int* rep_in = P.rep[cur];
    // ---- (1) log-softmax normaliser of the user's live rows from the LM-head tile partials
    const int nq = __ldcg(P.n_u + b);
    for (int i = warp; i < nq; i += PD_THREADS / 32) {
        const int r = __ldcg(P.live_u + b * K + i);
        combine_lse(P, S, r, r - b * K, n_parts, lane);
    }
    if (threadIdx.x < K) { S.s_nd[threadIdx.x] = __ldcg(node_in + b * K + threadIdx.x); S.s_rep[threadIdx.x] = __ldcg(rep_in + b * K + threadIdx.x); }
    __syncthreads();
    // ---- (2) score the trie children of every running beam, keep the best 2K (score desc, flat index asc)  [beam.cu topk]
    if (threadIdx.x < K) {            // children per beam, in parallel (the trie offsets are two dependent-free global loads each)
        const int nd = S.s_nd[threadIdx.x];
        S.s_cnt[threadIdx.x] = nd >= 0 ? P.t_off[nd + 1] - P.t_off[nd] : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < K; ++k) { S.s_off[k] = acc; acc += S.s_cnt[k]; }
        S.s_off[K] = acc;
    }
    __syncthreads();
    float* gsc = P.scr_score + (int64_t)b * P.scr_cap;
    int* gfl = P.scr_flat + (int64_t)b * P.scr_cap;
    const int n = min(S.s_off[K], P.scr_cap);
    const bool in_smem = n <= 1024;
    float* vsc = in_smem ? S.s_sc : gsc;
    int* vfl = in_smem ? S.s_fl : gfl;
    if (threadIdx.x < K) S.s_rs[threadIdx.x] = __ldcg(P.run_score[cur] + b * K + threadIdx.x);
    __syncthreads();
    // every candidate slot is one independent (trie token -> logit) load chain: all slots of the user in flight at once
    for (int slot = threadIdx.x; slot < n; slot += PD_THREADS) {
        int k = 0;
        while (k + 1 < K && S.s_off[k + 1] <= slot) ++k;           // beam that owns the slot (K <= 32)
        const int nd = S.s_nd[k], rk = S.s_rep[k];                 // rk: the row that owns this beam's logits
        const int tok = P.t_tok[P.t_off[nd] + (slot - S.s_off[k])];
        if (tok >= 0 && tok < V) {
            const float lp = (__ldcg(P.logits + (int64_t)(b * K + rk) * P.Vpad + tok) - S.rowmax[rk]) - S.logsum[rk];
            vsc[slot] = lp + S.s_rs[k];
            vfl[slot] = k * V + tok;
        } else {
            vsc[slot] = -INFINITY; vfl[slot] = 0x7fffffff;
        }
    }
    for (int sel = threadIdx.x; sel < 2 * K; sel += PD_THREADS) { S.s_lp[sel] = -INFINITY; S.s_cb[sel] = 0; S.s_ct[sel] = 0; }
    __syncthreads();
    if (in_smem) {
        for (int i = threadIdx.x; i < n; i += PD_THREADS) {
            const float v = S.s_sc[i];
            const int f = S.s_fl[i];
            if (!(v > -INFINITY)) continue;
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const float vj = S.s_sc[j];
                rank += (vj > v || (vj == v && S.s_fl[j] < f)) ? 1 : 0;
            }
            if (rank < 2 * K) { S.s_lp[rank] = v; S.s_cb[rank] = f / V; S.s_ct[rank] = f % V; }
        }
    } else {
        for (int sel = 0; sel < 2 * K; ++sel) {
            float best = -INFINITY; int bi = -1, bf = 0x7fffffff;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const float v = gsc[i]; const int f = gfl[i];
                if (v > best || (v == best && v > -INFINITY && f < bf)) { best = v; bi = i; bf = f; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int of = __shfl_xor_sync(0xffffffffu, bf, o);
                if (ov > best || (ov == best && ov > -INFINITY && of < bf)) { best = ov; bi = oi; bf = of; }
            }
            if (lane == 0) { S.s_best[warp] = best; S.s_besti[warp] = bi; S.s_bestf[warp] = bf; }
            __syncthreads();
            if (threadIdx.x == 0) {
                float bb = -INFINITY; int ii = -1, ff = 0x7fffffff;
                for (int w = 0; w < PD_THREADS / 32; ++w)
                    if (S.s_best[w] > bb || (S.s_best[w] == bb && bb > -INFINITY && S.s_bestf[w] < ff)) { bb = S.s_best[w]; ii = S.s_besti[w]; ff = S.s_bestf[w]; }
                if (ii >= 0 && bb > -INFINITY) { S.s_lp[sel] = bb; S.s_cb[sel] = ff / V; S.s_ct[sel] = ff % V; gsc[ii] = -INFINITY; }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // ---- (3) HF:generation/utils.py:2999-3073 running / finished beams + :2876-2921 early-stop heuristic  [beam.cu beam_update]
    const bool at_max = (cur_len + 1 >= P.max_len_eff);
    const bool us = __ldcg(P.unsat + b) != 0;
    const float denom_fin = powf((float)cur_len, P.length_penalty), denom_next = denom_fin;
    for (int c = threadIdx.x; c < 2 * K; c += PD_THREADS) {
        const float l = S.s_lp[c];
        const bool hit = (S.s_ct[c] == 1) || at_max;
        S.s_rv[c] = l + (hit ? PD_NEG_BIG : -0.0f);
        float v = l / denom_fin;
        v += us ? -0.0f : PD_NEG_BIG;
        v += (hit && c < K) ? -0.0f : PD_NEG_BIG;
        S.s_fv[K + c] = v;
    }
    for (int m = threadIdx.x; m < K; m += PD_THREADS) S.s_fv[m] = __ldcg(P.fin_score[cur] + b * K + m);
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * K; c += PD_THREADS) {
        const float v = S.s_rv[c];
        int rank = 0;
        for (int j = 0; j < 2 * K; ++j) rank += (S.s_rv[j] > v || (S.s_rv[j] == v && j < c)) ? 1 : 0;
        if (rank < K) {
            S.run_sel[rank] = c;
            P.run_score[nxt][b * K + rank] = v;
            if (rank == 0) S.s_misc[0] = v;
        }
    }
    for (int m = threadIdx.x; m < 3 * K; m += PD_THREADS) {
        const float v = S.s_fv[m];
        int rank = 0;
        for (int j = 0; j < 3 * K; ++j) rank += (S.s_fv[j] > v || (S.s_fv[j] == v && j < m)) ? 1 : 0;
        if (rank < K) {
            S.fin_sel[rank] = m;
            P.fin_score[nxt][b * K + rank] = v;
            int fin, gl;
            if (m < K) { fin = __ldcg(P.is_fin[cur] + b * K + m); gl = __ldcg(P.gen_len[cur] + b * K + m); }
            else {
                const int c = m - K;
                const bool hit = (S.s_ct[c] == 1) || at_max;
                fin = (hit && c < K) ? 1 : 0;
                gl = cur_len;
            }
            P.is_fin[nxt][b * K + rank] = fin; P.gen_len[nxt][b * K + rank] = gl;
            S.s_fin[rank] = fin;
            if (rank == K - 1) S.s_misc[1] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float best_running = S.s_misc[0] / denom_next, min_fs = S.s_misc[1];
        bool any = false;
        for (int k = 0; k < K; ++k) {
            const float worst = S.s_fin[k] ? min_fs : PD_NEG_BIG;
            if (best_running > worst) any = true;
        }
        P.unsat[b] = (us && any) ? 1 : 0;
    }
    // ---- (4) new trie nodes and DUPLICATE detection: beams with the same (parent representative, token) share one state
    if (threadIdx.x < K) {
        const int k = threadIdx.x, c = S.run_sel[k];
        const int parent = S.s_cb[c];
        S.node_new[k] = (S.s_lp[c] > -INFINITY) ? pd_trie_child(P.t_off, P.t_tok, P.t_node, S.s_nd[parent], S.s_ct[c]) : -1;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        const int k = threadIdx.x, c = S.run_sel[k];
        const int prep = S.s_rep[S.s_cb[c]], tok = S.s_ct[c];
        int first = k;
        for (int k2 = 0; k2 < k; ++k2) {
            const int c2 = S.run_sel[k2];
            if (S.s_rep[S.s_cb[c2]] == prep && S.s_ct[c2] == tok) { first = k2; break; }
        }
        S.rep_new[k] = first;
        P.rep[nxt][b * K + k] = first;
        P.node[nxt][b * K + k] = S.node_new[k];
        P.cur_tok[b * K + k] = tok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int cnt = 0;
        for (int k = 0; k < K; ++k)
            if (S.rep_new[k] == k && S.node_new[k] >= 0) P.live_u[b * K + cnt++] = b * K + k;    // rows that get computed next step
        P.n_u[b] = cnt;
        S.n_new = cnt;
    }
    // ---- (5) materialise the selected rows (sequences, finished sequences, KV row indirection)
    for (int idx = threadIdx.x; idx < K * T; idx += PD_THREADS) {
        const int k = idx / T, t = idx - k * T;
        const int c = S.run_sel[k];
        const int parent = b * K + S.s_cb[c];
        const int r = b * K + k;
        int v = __ldcg(P.seq[cur] + parent * T + t);
        if (t == cur_len) v = S.s_ct[c];
        P.seq[nxt][r * T + t] = v;
        int sr = __ldcg(P.src[cur] + parent * T + t);
        if (t >= cur_len) sr = b * K + S.rep_new[k];                // own positions: the representative row holds the K/V
        P.src[nxt][r * T + t] = sr;
        const int m = S.fin_sel[k];
        int fv;
        if (m < K) fv = __ldcg(P.fin_seq[cur] + (b * K + m) * T + t);
        else {
            const int c2 = m - K;
            fv = __ldcg(P.seq[cur] + (b * K + S.s_cb[c2]) * T + t);
            if (t == cur_len) fv = S.s_ct[c2];
        }
        P.fin_seq[nxt][r * T + t] = fv;
    }
    __syncthreads();
    // ---- (6) decoder input of the next position for the live rows
    if (step + 1 < P.n_steps) {
        for (int i = warp; i < S.n_new; i += PD_THREADS / 32) {
            const int r = P.live_u[b * K + i];
            embed_row(P, r, S.s_ct[S.run_sel[r - b * K]], lane);
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// Optional (P5_DECODE_PREFETCH=1) L2 prefetch of the users' cross K | V of a layer, requested four phases ahead by a warp that
// has no role in the GEMM phase; each CTA asks for its 1 / gridDim share of the encoder rows.  (Prefetching the weights of
// the next phases the same way was measured in round 2: no phase got shorter, the issuing phase got longer.)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ void prefetch_cross_kv(const PdParams& P, const PdLayer& L) {          // K | V of one layer: 2A contiguous bf16 per encoder row
    const int rows = P.B * P.Le;
    for (int r = blockIdx.x + gridDim.x * (threadIdx.x & 31); r < rows; r += gridDim.x * 32)
        if (__ldg(P.mask_e + r)) l2_prefetch(L.ck + (int64_t)r * P.ckv_ld, (uint32_t)(2 * P.A * 2));
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PD_THREADS, 1) decode_persistent_kernel(const PdParams* __restrict__ Pp) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ UserSmem US;
    __shared__ int s_nlive;
    __shared__ int s_uoff[65];
    const PdParams& P = *Pp;
    int* s_live = reinterpret_cast<int*>(smem);                       // [R] live rows of the current position
    uint8_t* work = smem + ((P.R * 4 + 127) & ~127);                  // GEMM ring / attention staging
    // ---- tcgen05 pipeline of the heavy GEMM phases: ring (1024-aligned, SWIZZLE_128B), mbarriers, TMEM accumulators
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_tcflag;
    const uint32_t ring = (smem_u32(work) + 1023u) & ~1023u;
    const uint32_t bars = ring + T_STAGES * T_STAGE_BYTES;
    if (threadIdx.x == 32) {
        for (int s2 = 0; s2 < T_STAGES; ++s2) { mbar_init(bars + 8u * s2, 1); mbar_init(bars + 8u * (T_STAGES + s2), 1); }
        for (int s2 = 0; s2 < 2; ++s2) { mbar_init(bars + 8u * (2 * T_STAGES + s2), 1); mbar_init(bars + 8u * (2 * T_STAGES + 2 + s2), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if ((threadIdx.x >> 5) == 2) tmem_alloc(smem_u32(&s_tmem_base), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    TcState tcs;
    unsigned bar_target = 0;
    const int B = P.B, K = P.K, R = P.R, T = P.T, d = P.d, A = P.A, ff = P.ff;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- init (HF:generation/utils.py:3190-3215): only beam 0 of every user is live, the K-1 others are its duplicates
    for (int r = blockIdx.x * PD_THREADS + threadIdx.x; r < R; r += gridDim.x * PD_THREADS) {
        const int b = r / K, k = r - b * K;
        for (int t = 0; t < T; ++t) { P.seq[0][r * T + t] = 0; P.fin_seq[0][r * T + t] = 0; P.src[0][r * T + t] = b * K; }
        P.node[0][r] = P.root_child;
        P.rep[0][r] = 0;
        P.run_score[0][r] = (k == 0) ? 0.f : PD_NEG_BIG;
        P.fin_score[0][r] = PD_NEG_BIG;
        P.is_fin[0][r] = 0; P.gen_len[0][r] = 0; P.cur_tok[r] = 0;
        if (k == 0) { P.unsat[b] = 1; P.n_u[b] = 1 + P.n_forced; }
        if (k <= P.n_forced) P.live_u[r] = r;          // prefill: item (b, t) = row b*K + t, t = 0 .. n_forced
    }
    {   // decoder inputs: the start token, and (prefill) the forced prefix tokens as the inputs of positions 1 .. p
        const int per = 1 + P.n_forced;
        for (int e = blockIdx.x * (PD_THREADS / 32) + warp; e < B * per; e += gridDim.x * (PD_THREADS / 32)) {
            const int b = e / per, t = e - b * per;
            embed_row(P, b * K + t, t == 0 ? 0 : P.forced[t - 1], lane);
        }
    }
    grid_barrier(P.bar, bar_target);

    uint64_t t_last = pd_timer();
    bool prof_heavy = false;
    auto mark = [&](int kind) {     // CTA 0 attributes the time since the previous mark to `kind`
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const uint64_t now = pd_timer();
            P.prof[(prof_heavy ? 16 : 0) + kind] += now - t_last;
            t_last = now;
        }
    };
    // heavy positions only: CTA 0's own work time of the two attention phases (slots 27 / 29), its wait in the grid barrier
    // behind them (28 / 30), and one back-to-back barrier per position as the cost of a bare barrier (31)
    auto mark_sub = [&](int slot) {
        if (prof_heavy && blockIdx.x == 0 && threadIdx.x == 0) { const uint64_t now = pd_timer(); P.prof[slot] += now - t_last; }
    };
    int cur = 0;
    const int n_forced = P.n_forced;
    // iteration -1 (only with a forced prefix): the PREFILL pass over positions 0 .. p of every user; it produces the logits of
    // position p, so the regular loop continues at position p + 1
    for (int it = (n_forced > 0 ? -1 : 0); it < P.n_steps; it = (it < 0 ? n_forced + 1 : it + 1)) {
        const bool prefill = it < 0;
        const int step = prefill ? n_forced : it;
        const int cur_len = step + 1, pos = step;
        // ---- live rows of this position (every CTA builds the same list)
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int b = 0; b < B; ++b) { s_uoff[b] = acc; acc += __ldcg(P.n_u + b); }
            s_uoff[B] = acc;
            s_nlive = acc;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < B * K; e += PD_THREADS) {
            const int b = e / K, i = e - b * K;
            if (i < s_uoff[b + 1] - s_uoff[b]) s_live[s_uoff[b] + i] = __ldcg(P.live_u + b * K + i);
        }
        __syncthreads();
        const int n_live = s_nlive;
        int max_nq = 0;
        for (int b = 0; b < B; ++b) max_nq = max(max_nq, s_uoff[b + 1] - s_uoff[b]);
        const float inv_d = 1.f / (float)d;
        // <= 32 live rows (one distinct beam per user through the forced item prefix): weight-streaming variant
        const bool light = n_live <= 32 && !P.no_light;
        const int n_parts = light ? (P.V + L_TN - 1) / L_TN : P.n_ct_head;
        // heavy positions: every beam row through the tcgen05 tiles (rows that are not live compute values nobody reads);
        // the mma.sync tile loop (gemm_phase) remains as the P5_DECODE_NO_TC=1 cross-check
        auto run_gemm = [&](const GemmDesc& gd, int wmap, int amap) {
            if (light) gemm_light(gd, s_live, n_live, work);
            else if (P.no_tc) gemm_phase(gd, s_live, n_live, work);
            else gemm_tc_phase(gd, P.tmaps + amap, P.tmaps + wmap, R, ring, bars, tmem_base, tcs, P.tile_cnt, &s_tcflag);
        };
        const int TM_HEAD = TM_PER_LAYER * P.ND, TM_Y16 = TM_HEAD + 1, TM_CTX = TM_HEAD + 2, TM_H = TM_HEAD + 3;
        prof_heavy = !light;
        mark(PH_LIST);

        for (int l = 0; l < P.ND; ++l) {
            const PdLayer& L = P.layer[l];
            float* ss0 = P.rowss + (int64_t)(3 * l) * R;
            float* ss1 = ss0 + R;
            float* ss2 = ss1 + R;
            float* ss_next = ss2 + R;                                   // site 3 (l + 1): next layer's ln0, or the final norm
            const float* ln_after = (l + 1 < P.ND) ? P.layer[l + 1].ln0 : P.ln_final;
            GemmDesc g;
            if ((threadIdx.x >> 5) == 2 && !P.no_prefetch) prefetch_cross_kv(P, L);     // needed four phases from here
            // (a) q | k | v = RMSNorm(y) . Wqkv^T
            g = GemmDesc{P.y16, d, d, L.wqkv, 3 * A, 0, 1.f, ss0, inv_d, P.eps, 0, P.qkv, 3 * A, nullptr, nullptr, nullptr, nullptr, d,
                         nullptr, 0, nullptr, 0, 0};
            run_gemm(g, TM_PER_LAYER * l + TM_QKV, TM_Y16);
            grid_barrier(P.bar, bar_target);
            mark(PH_QKV);
            // (b) self-attention over the cached positions (+ KV append)
            if (prefill) self_attn_prefill(P, L, n_forced);
            else self_attn_phase(P, L, s_live, n_live, P.src[cur], pos);
            __syncthreads();
            mark_sub(29);
            grid_barrier(P.bar, bar_target);
            mark(PH_SA);
            // (c) y += ctx . Wo^T      -> y16 = bf16(y * ln1), rowss1
            g = GemmDesc{P.ctx, A, A, L.wo, d, 1, 1.f, nullptr, inv_d, P.eps, 0, nullptr, 0, P.y, P.y16, L.ln1, ss1, d, nullptr, 0, nullptr, 0, 0};
            run_gemm(g, TM_PER_LAYER * l + TM_O, TM_CTX);
            grid_barrier(P.bar, bar_target);
            mark(PH_O);
            // (d) cross-attention query
            g = GemmDesc{P.y16, d, d, L.wcq, A, 0, 1.f, ss1, inv_d, P.eps, 0, P.cq, A, nullptr, nullptr, nullptr, nullptr, d, nullptr, 0, nullptr, 0, 0};
            run_gemm(g, TM_PER_LAYER * l + TM_CQ, TM_Y16);
            grid_barrier(P.bar, bar_target);
            mark(PH_CQ);
            // (e) cross-attention over the user's encoder K | V (zero position bias + encoder padding mask)
            if (max_nq <= 8 && P.Le <= 256) cross_attn_light(P, L, work);
            else if (P.Le <= 256) cross_attn_phase_x2(P, L, work);
            else cross_attn_phase<8>(P, L, work);
            __syncthreads();
            mark_sub(27);
#ifdef P5_CA_STAMPS
            if (prof_heavy && blockIdx.x == 0 && threadIdx.x == 0)
                for (int i = 0; i < 5; ++i) { const int a0[5] = {0, 2, 3, 5, 6}, a1[5] = {2, 3, 5, 6, 7}; P.prof[11 + i] += g_ca_stamp[a1[i]] - g_ca_stamp[a0[i]]; }
#endif
            grid_barrier(P.bar, bar_target);
            mark(PH_CA);
            // (f) y += ctx . Wco^T     -> y16 = bf16(y * ln2), rowss2
            g = GemmDesc{P.ctx, A, A, L.wco, d, 1, 1.f, nullptr, inv_d, P.eps, 0, nullptr, 0, P.y, P.y16, L.ln2, ss2, d, nullptr, 0, nullptr, 0, 0};
            run_gemm(g, TM_PER_LAYER * l + TM_CO, TM_CTX);
            grid_barrier(P.bar, bar_target);
            mark(PH_CO);
            // (g) h = relu(RMSNorm(y) . Wi^T)
            g = GemmDesc{P.y16, d, d, L.wi, ff, 0, 1.f, ss2, inv_d, P.eps, 1, P.h, ff, nullptr, nullptr, nullptr, nullptr, d, nullptr, 0, nullptr, 0, 0};
            run_gemm(g, TM_PER_LAYER * l + TM_WI, TM_Y16);
            grid_barrier(P.bar, bar_target);
            mark(PH_WI);
            // (h) y += h . Wo^T        -> y16 = bf16(y * ln0 of the next block / final norm), rowss of that site
            g = GemmDesc{P.h, ff, ff, L.wwo, d, 1, 1.f, nullptr, inv_d, P.eps, 0, nullptr, 0, P.y, P.y16, ln_after, ss_next, d, nullptr, 0, nullptr, 0, 0};
            run_gemm(g, TM_PER_LAYER * l + TM_WO, TM_H);
            grid_barrier(P.bar, bar_target);
            mark(PH_WO);
        }
        // ---- tied LM head: logits = (RMSNorm(y) * d^-0.5) . E^T  (+ per-tile log-sum-exp partials)   (P5_T5.py:357-361)
        {
            GemmDesc g{P.y16, d, d, P.E16, P.V, 2, P.hs, P.rowss + (int64_t)(3 * P.ND) * R, inv_d, P.eps, 0, nullptr, 0, nullptr, nullptr,
                       nullptr, nullptr, d, P.logits, P.Vpad, P.lse_part, n_parts, P.V};
            run_gemm(g, TM_HEAD, TM_Y16);
        }
        grid_barrier(P.bar, bar_target);
        mark(PH_LM);
        // ---- per user: normaliser, constrained top-2K, beam update, next input embedding
        for (int b = blockIdx.x; b < B; b += gridDim.x) {
            if (prefill) prefill_user_init(P, US, b, cur, n_forced, n_parts);
            user_phase(P, US, b, cur, cur_len, step, n_parts);
        }
        grid_barrier(P.bar, bar_target);
        mark(PH_USER);
        if (P.prof_fine) { grid_barrier(P.bar, bar_target); mark_sub(31); mark(PH_LIST); }
        cur ^= 1;
    }
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
    // ---- finalize (HF:generation/utils.py:3380-3400): the n_ret best finished hypotheses per user, cropped length
    if (blockIdx.x == 0) {
        __shared__ int s_max;
        if (threadIdx.x == 0) s_max = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < B * P.n_ret; i += PD_THREADS) {
            const int b = i / P.n_ret, k = i - b * P.n_ret;
            const int r = b * K + k;
            P.out_scores[i] = __ldcg(P.fin_score[cur] + r);
            for (int t = 0; t < P.max_len; ++t) P.out_seqs[(int64_t)i * P.max_len + t] = __ldcg(P.fin_seq[cur] + r * T + t);
            if (__ldcg(P.is_fin[cur] + r)) atomicMax(&s_max, __ldcg(P.gen_len[cur] + r));
        }
        __syncthreads();
        if (threadIdx.x == 0) P.out_len[0] = 1 + s_max;
    }
}

struct PersistWs {
    int R = 0, T = 0, K = 0, B = 0, ND = 0, d = 0, A = 0, ff = 0, Vpad = 0, scr_cap = 0;
    std::vector<void*> allocs;
    PdParams host;
    PdParams* dev = nullptr;
    CUtensorMap* tmaps_dev = nullptr;
    int smem_bytes = 0, grid = 0;
};
PersistWs* g_pws = nullptr;
cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;     // around the last persistent launch (bench.py eval roofline)
double g_last_bytes = 0.0;
int g_last_steps = 0;

}  // namespace

void free_persist_ws() {
    if (!g_pws) return;
    for (void* p : g_pws->allocs) cudaFree(p);
    delete g_pws;
    g_pws = nullptr;
}

// timing of the last persistent launch: milliseconds (CUDA events on the launch stream), algorithmic bytes, positions
int decode_last_launch(float* ms, double* bytes, int* steps) {
    if (!g_ev0) return 1;
    if (cudaEventSynchronize(g_ev1) != cudaSuccess) return 2;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_ev0, g_ev1) != cudaSuccess) return 3;
    if (ms) *ms = t;
    if (bytes) *bytes = g_last_bytes;
    if (steps) *steps = g_last_steps;
    return 0;
}

// per-phase nanoseconds of the last launch as CTA 0 saw them: out[0..15] positions run by the light GEMM variant,
// out[16..31] heavy positions; index = PdPhase (qkv, self-attn, o, cq, cross-attn, co, wi, wo, lm-head, user, list)
int decode_phase_ns(unsigned long long* out32) {
    if (!g_pws || !g_ev1) return 1;
    if (cudaEventSynchronize(g_ev1) != cudaSuccess) return 2;
    return cudaMemcpy(out32, g_pws->host.prof, 32 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 3;
}

bool decode_persistent_supported(const Engine* e, int B, int K, int max_len, int Le) {
    static const bool off = getenv("P5_DECODE_KERNELS") != nullptr;       // A/B switch: the per-position kernel chain of beam.cu
    if (off) return false;
    return e->dt == DT_BF16 && !e->gated && K >= 1 && K <= PD_MAXK && B >= 1 && B <= 64 && e->ND >= 1 && e->ND <= PD_MAXL &&
           e->d % 64 == 0 && e->ff % 64 == 0 && e->A % 64 == 0 && Le <= 512 && max_len <= 256;
}

// the search itself; the caller (beam.cu generate) has run the encoder, projected the cross K|V and built the decoder bias
const int* generate_persistent(Engine* e, const int* t_off, const int* t_tok, const int* t_node, int root_child, int max_depth,
                               int max_fanout, int B, int K, int Rret, int max_len, float length_penalty, int32_t* seqs, float* scores,
                               const int* forced_host, int n_forced, int node_forced) {
    cudaStream_t st = e->st;
    const int R = B * K, T = max_len, d = e->d, A = e->A, ff = e->ff, ND = e->ND, Vpad = e->Vpad;
    const int cand_cap = K * (max_fanout > 0 ? max_fanout : 1);
    PersistWs* w = g_pws;
    if (!w || w->R < R || w->T < T || w->K != K || w->B < B || w->ND != ND || w->d != d || w->A != A || w->ff != ff || w->Vpad != Vpad ||
        w->scr_cap < cand_cap) {
        P5_CUDA(cudaStreamSynchronize(st));
        free_persist_ws();
        w = g_pws = new PersistWs();
        w->R = R; w->T = T; w->K = K; w->B = B; w->ND = ND; w->d = d; w->A = A; w->ff = ff; w->Vpad = Vpad; w->scr_cap = cand_cap;
        auto al = [&](size_t bytes) { void* p = nullptr; P5_CUDA(cudaMalloc(&p, bytes ? bytes : 256)); P5_CUDA(cudaMemsetAsync(p, 0, bytes ? bytes : 256, st)); w->allocs.push_back(p); return p; };
        PdParams& H = w->host;
        memset(&H, 0, sizeof(H));
        H.y = (float*)al((size_t)R * d * 4); H.y16 = (bf16*)al((size_t)R * d * 2);
        H.rowss = (float*)al((size_t)(3 * ND + 1) * R * 4);
        H.qkv = (bf16*)al((size_t)R * 3 * A * 2); H.ctx = (bf16*)al((size_t)R * A * 2); H.cq = (bf16*)al((size_t)R * A * 2);
        H.h = (bf16*)al((size_t)R * ff * 2);
        H.logits = (float*)al((size_t)R * Vpad * 4);
        H.n_ct_head = (int)cdiv(e->V, G_TN);
        H.lse_part = (float2*)al((size_t)R * cdiv(e->V, L_TN) * 8);          // sized for the 16-column tiles of the light variant
        for (int l = 0; l < ND; ++l) { H.layer[l].Kc = (bf16*)al((size_t)R * T * A * 2); H.layer[l].Vc = (bf16*)al((size_t)R * T * A * 2); }
        for (int i = 0; i < 2; ++i) {
            H.seq[i] = (int*)al((size_t)R * T * 4); H.fin_seq[i] = (int*)al((size_t)R * T * 4); H.src[i] = (int*)al((size_t)R * T * 4);
            H.node[i] = (int*)al((size_t)R * 4); H.rep[i] = (int*)al((size_t)R * 4); H.is_fin[i] = (int*)al((size_t)R * 4);
            H.gen_len[i] = (int*)al((size_t)R * 4); H.run_score[i] = (float*)al((size_t)R * 4); H.fin_score[i] = (float*)al((size_t)R * 4);
        }
        H.cur_tok = (int*)al((size_t)R * 4); H.unsat = (int*)al((size_t)B * 4); H.live_u = (int*)al((size_t)R * 4); H.n_u = (int*)al((size_t)B * 4);
        H.scr_score = (float*)al((size_t)B * cand_cap * 4); H.scr_flat = (int*)al((size_t)B * cand_cap * 4); H.scr_cap = cand_cap;
        H.bar = (unsigned*)al(256);
        H.prof = (unsigned long long*)al(32 * 8);
        H.tile_cnt = (int*)al(1024 * 4);
        H.out_len = (int*)al(16);
        w->dev = (PdParams*)al(sizeof(PdParams));
        w->tmaps_dev = (CUtensorMap*)al(sizeof(CUtensorMap) * (TM_PER_LAYER * ND + 4));
        // launch geometry: one CTA per SM, all co-resident (cooperative launch)
        const int work = std::max(std::max(G_RING_BYTES, T_RING_BYTES), std::max(DCfg<8, PD_THREADS / 32>::TILE, PD_THREADS * 64 * 4) + PD_THREADS * 2 * 4);
        w->smem_bytes = (int)round_up((int64_t)R * 4, 128) + work;
        P5_CUDA(cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, w->smem_bytes));
        int sms = 0, per_sm = 0;
        P5_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
        P5_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_persistent_kernel, PD_THREADS, w->smem_bytes));
        P5_CHECK(per_sm >= 1, "persistent decode kernel does not fit on an SM");
        w->grid = sms;
    }
    PdParams& H = w->host;
    H.B = B; H.K = K; H.R = R; H.T = T; H.Le = e->Le; H.d = d; H.A = A; H.H = e->H; H.ff = ff; H.V = e->V; H.Vpad = Vpad; H.ND = ND;
    H.n_steps = std::min(max_len - 1, max_depth - 1);
    H.max_len = max_len; H.max_len_eff = std::min(max_len, H.n_steps + 1); H.n_ret = Rret; H.root_child = root_child;
    H.eps = e->cfg.ln_eps; H.hs = 1.f / sqrtf((float)d); H.length_penalty = length_penalty;
    static const bool no_light = getenv("P5_DECODE_NO_LIGHT") != nullptr;
    H.no_light = no_light ? 1 : 0;
    static const bool want_prefetch = getenv("P5_DECODE_PREFETCH") != nullptr;
    H.no_prefetch = want_prefetch ? 0 : 1;      // measured r02: requesting the next phases' bytes ahead does not shorten them
    static const bool prof_fine = getenv("P5_DECODE_PROF_FINE") != nullptr;
    H.prof_fine = prof_fine ? 1 : 0;
    // forced item prefix (single-child trie nodes from the root): decoded in ONE prefill pass instead of one position at a
    // time.  Needs the items (b, t) of a user to fit its K beam rows and the last position to stay in the regular loop.
    static const bool no_prefill = getenv("P5_DECODE_NO_PREFILL") != nullptr;
    if (no_prefill || n_forced + 1 > K || n_forced > 31 || n_forced + 1 > H.n_steps) n_forced = 0;
    H.n_forced = n_forced; H.node_forced = n_forced > 0 ? node_forced : root_child;
    for (int i = 0; i < 32; ++i) H.forced[i] = i < n_forced ? forced_host[i] : 0;
    for (int l = 0; l < ND; ++l) {
        const DecLayerOff& o = e->dec[l];
        PdLayer& L = H.layer[l];
        L.wqkv = e->P16 + o.sa.q; L.wo = e->P16 + o.sa.o; L.wcq = e->P16 + o.ca.q; L.wco = e->P16 + o.ca.o;
        L.wi = e->P16 + o.ff.wi; L.wwo = e->P16 + o.ff.wo;
        L.ln0 = e->P + o.ln0; L.ln1 = e->P + o.ln1; L.ln2 = e->P + o.ln2;
        L.ck = (const bf16*)e->ckv[l]; L.cv = L.ck + A;
    }
    H.E = e->P + e->off_shared; H.E16 = e->P16 + e->off_shared; H.ln_final = e->P + e->off_dec_final;
    H.ckv_ld = e->ckv_ld; H.mask_e = e->mask_e;
    H.bias_dec = e->bias_dec; H.n_delta = 2 * T - 1; H.bias_off = T - 1;
    H.t_off = t_off; H.t_tok = t_tok; H.t_node = t_node;
    H.out_seqs = seqs; H.out_scores = scores;
    {   // tensor maps of the tcgen05 phases: K-major bf16, SWIZZLE_128B, boxes 64 (k) x 128 (A rows) / 64 (W rows)
        std::vector<CUtensorMap> maps(TM_PER_LAYER * ND + 4);
        auto mk = [&](const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
            const uint64_t dims[4] = {(uint64_t)K, (uint64_t)rows, 1, 1};
            const uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)rows * ld * 2, (uint64_t)rows * ld * 2};
            const uint32_t box[4] = {(uint32_t)T_KB, (uint32_t)box_rows, 1, 1};
            return tmap_bf16_4d(ptr, dims, strides, box);
        };
        for (int l = 0; l < ND; ++l) {
            const PdLayer& L = H.layer[l];
            maps[TM_PER_LAYER * l + TM_QKV] = mk(L.wqkv, 3 * A, d, d, T_TN);
            maps[TM_PER_LAYER * l + TM_O] = mk(L.wo, d, A, A, T_TN);
            maps[TM_PER_LAYER * l + TM_CQ] = mk(L.wcq, A, d, d, T_TN);
            maps[TM_PER_LAYER * l + TM_CO] = mk(L.wco, d, A, A, T_TN);
            maps[TM_PER_LAYER * l + TM_WI] = mk(L.wi, ff, d, d, T_TN);
            maps[TM_PER_LAYER * l + TM_WO] = mk(L.wwo, d, ff, ff, T_TN);
        }
        maps[TM_PER_LAYER * ND] = mk(H.E16, e->V, d, d, T_TN);
        maps[TM_PER_LAYER * ND + 1] = mk(H.y16, R, d, d, T_TM);
        maps[TM_PER_LAYER * ND + 2] = mk(H.ctx, R, A, A, T_TM);
        maps[TM_PER_LAYER * ND + 3] = mk(H.h, R, ff, ff, T_TM);
        P5_CUDA(cudaMemcpyAsync(w->tmaps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
        H.tmaps = w->tmaps_dev;
        static const bool no_tc = getenv("P5_DECODE_NO_TC") != nullptr;
        H.no_tc = (no_tc || d % 64 != 0) ? 1 : 0;
    }
    P5_CUDA(cudaMemcpyAsync(w->dev, &H, sizeof(PdParams), cudaMemcpyHostToDevice, st));
    P5_CUDA(cudaMemsetAsync(H.bar, 0, 256, st));
    P5_CUDA(cudaMemsetAsync(H.prof, 0, 32 * 8, st));
    const PdParams* dp = w->dev;
    void* args[] = {(void*)&dp};
    if (!g_ev0) { P5_CUDA(cudaEventCreate(&g_ev0)); P5_CUDA(cudaEventCreate(&g_ev1)); }
    P5_CUDA(cudaEventRecord(g_ev0, st));
    P5_CUDA(cudaLaunchCooperativeKernel((const void*)decode_persistent_kernel, dim3(w->grid), dim3(PD_THREADS), args, (size_t)w->smem_bytes, st));
    P5_CUDA(cudaEventRecord(g_ev1, st));
    ++g_launches;
    // algorithmic bytes of the launch: per position the decoder-block weights it multiplies with and the
    // tied LM head once (bf16), plus every user's cross K|V once (bf16)
    // (a PASS is one trip through the decoder: the prefill covers positions 0 .. n_forced in one, so the weights stream
    // n_steps - n_forced times, not n_steps times)
    g_last_steps = H.n_steps;
    const int passes = H.n_forced > 0 ? H.n_steps - H.n_forced : H.n_steps;
    g_last_bytes = (double)passes * (2.0 * ((double)ND * ((double)3 * A * d + 3.0 * (double)A * d + 2.0 * (double)ff * d) + (double)e->V * d) +
                                        2.0 * (double)B * e->Le * ND * 2.0 * A);
    return H.out_len;      // device: 1 + longest returned hypothesis
}

}  // namespace p5
