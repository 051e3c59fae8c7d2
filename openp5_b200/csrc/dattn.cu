// Decoder attention of the TRAIN step in bf16 mode (HF:models/t5/modeling_t5.py:253-344, T5Attention called from the
// decoder T5LayerSelfAttention / T5LayerCrossAttention): Lq = Ld <= 16 query rows per sequence, Lk <= 512 keys.
//
// Why its own kernel: with 8 target tokens per user a (b, h) pair has 8 x Lk scores.  A tcgen05 tile needs 128 query
// rows (16x padding, three launches with S round-tripped through HBM), and the fp32 SIMT kernel spent 53 us forward /
// 110 us backward per layer on the cross-attention issuing scalar FMAs.  Here one CTA owns a (b, h) pair and runs the
// four contractions on m16n8k16 mma.sync tiles (the query block is exactly one M=16 tile), everything else in
// registers:
//   * Q / K / dO / V fragments are loaded straight from global memory with 2 x 16-byte loads per row.  The d_kv = 64
//     contraction index is PERMUTED (thread t of a quad owns d in [16t, 16t+16)), which is legal because A and B use
//     the same permutation and a dot product does not care about the order of its terms;
//   * each warp owns a contiguous key range; softmax statistics are combined across warps through shared memory;
//   * P (forward) and dS (backward) go from the accumulator layout directly into the A operand of P.V / dS.K;
//   * V (forward) / K (backward) are staged once in shared memory with cp.async (128-byte rows, 16-byte chunks XOR-
//     swizzled by row) and read with ldmatrix.trans as B operands; dV = Pd^T dO and dK = dS^T Q read Pd / dS back from
//     a per-warp shared tile with ldmatrix.trans (transposed A operand);
//   * dropout uses the same counter hash and element index as attn_simt_* so forward and backward agree by
//     construction; gradients are written as bf16 in place (no fp32 staging buffer, no cast, no memset).
#include "dattn_dev.cuh"

namespace p5 {
extern int g_launches;
#define LAUNCHED() do { P5_CUDA(cudaGetLastError()); ++g_launches; } while (0)

namespace {

// ------------------------------------------------------------------------------------------------------------
// forward: O[b, i, h*64 + c] (bf16) and LSE[b, h, i]
// ------------------------------------------------------------------------------------------------------------
template <int NT, int NW, bool LQ16>
__global__ void __launch_bounds__(NW * 32)
dattn_fwd_kernel(DAttnDev a, bf16* __restrict__ O, int64_t ld_o, int64_t bs_o, float* __restrict__ lse) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    using C = DCfg<NT, NW>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* Vs = smem;                                              // later aliased by the fp32 partial outputs
    float* red = reinterpret_cast<float*>(smem + (C::TILE > C::ORED ? C::TILE : C::ORED));   // [2][NW][16]
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int Lk = a.kv_len ? a.kv_len[b] : a.Lk;
    const int64_t k_boff = a.kv_off ? (int64_t)a.kv_off[b] * a.k_ld : (int64_t)b * a.k_bs;
    const int64_t v_boff = a.kv_off ? (int64_t)a.kv_off[b] * a.v_ld : (int64_t)b * a.v_bs;
    const int kpw = min(C::KW, (((Lk + NW - 1) / NW) + 15) & ~15);   // keys per warp (multiple of 16)
    const int key0 = warp * kpw, ntw = kpw >> 3;
    const bf16* kb = a.k + k_boff + h * 64;
    const bf16* vb = a.v + v_boff + h * 64;

    stage_rows(Vs, warp * C::KW, vb, a.v_ld, key0, kpw, Lk, lane);     // V rows of this warp -> smem (async)

    uint32_t qlo[8], qhi[8];
    const bf16* qb = a.q + (int64_t)b * a.q_bs + h * 64 + 16 * t;
    ld_row16(qlo, qb + (int64_t)g * a.q_ld, g < a.Lq);
    if (LQ16) ld_row16(qhi, qb + (int64_t)(g + 8) * a.q_ld, g + 8 < a.Lq);

    float s[NT][4];
    qk_tiles<NT, LQ16>(s, qlo, qhi, kb, a.k_ld, key0, ntw, Lk, g, t);

    ScoreCtx sc{a.Lq, Lk, a.Lk, a.causal, a.bias_off, a.n_delta, a.bias_rel ? a.bias_rel + h * a.n_delta : nullptr,
                a.key_mask ? a.key_mask + (int64_t)b * a.Lk : nullptr};
    float mx[2] = {-FLT_MAX, -FLT_MAX};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r = g + 8 * (e >> 1), j = key0 + 8 * nt + 2 * t + (e & 1);
            float v = -FLT_MAX;
            if (nt < ntw && (LQ16 || e < 2) && score_valid(sc, r, j)) v = s[nt][e] + (sc.bias ? sc.bias[bias_index(sc, r, j)] : 0.f);
            s[nt][e] = v;
            mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], 1));
        mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], 2));
    }
    if (NW > 1) {
        if (t == 0) { red[warp * 16 + g] = mx[0]; red[warp * 16 + g + 8] = mx[1]; }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NW; ++w) { mx[0] = fmaxf(mx[0], red[w * 16 + g]); mx[1] = fmaxf(mx[1], red[w * 16 + g + 8]); }
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float p = (s[nt][e] == -FLT_MAX) ? 0.f : __expf(s[nt][e] - mx[e >> 1]);
            s[nt][e] = p;
            sum[e >> 1] += p;
        }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        sum[q] += __shfl_xor_sync(0xffffffffu, sum[q], 1);
        sum[q] += __shfl_xor_sync(0xffffffffu, sum[q], 2);
    }
    if (NW > 1) {
        float* red2 = red + NW * 16;
        if (t == 0) { red2[warp * 16 + g] = sum[0]; red2[warp * 16 + g + 8] = sum[1]; }
        __syncthreads();
        sum[0] = sum[1] = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { sum[0] += red2[w * 16 + g]; sum[1] += red2[w * 16 + g + 8]; }
    }
    if (warp == 0 && t == 0 && lse) {
        if (g < a.Lq) lse[((int64_t)b * a.H + h) * a.Lq + g] = mx[0] + logf(sum[0]);
        if (LQ16 && g + 8 < a.Lq) lse[((int64_t)b * a.H + h) * a.Lq + g + 8] = mx[1] + logf(sum[1]);
    }
    const float inv[2] = {sum[0] > 0.f ? 1.f / sum[0] : 0.f, sum[1] > 0.f ? 1.f / sum[1] : 0.f};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r = g + 8 * (e >> 1), j = key0 + 8 * nt + 2 * t + (e & 1);
            float p = s[nt][e] * inv[e >> 1];
            if (a.drop.thr && p != 0.f) {
                const uint64_t idx = (((uint64_t)b * a.H + h) * a.Lq + r) * (uint64_t)Lk + j;
                p = drop_keep(a.drop.seed, a.drop.site, idx, a.drop.thr) ? p * a.drop.inv_keep : 0.f;
            }
            s[nt][e] = p;
        }
    // ---- O = P V
    float o[8][4];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c][0] = o[c][1] = o[c][2] = o[c][3] = 0.f;
    cp_async_wait_all();
    __syncwarp();
    pv_tiles<NT>(o, s, Vs, warp * C::KW, ntw, lane);
    if (NW > 1) {
        __syncthreads();                       // every warp is done with its V rows: reuse the tile for the partial sums
        float* part = reinterpret_cast<float*>(smem);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<float2*>(part + (warp * 16 + g) * 64 + 8 * c + 2 * t) = make_float2(o[c][0], o[c][1]);
            *reinterpret_cast<float2*>(part + (warp * 16 + g + 8) * 64 + 8 * c + 2 * t) = make_float2(o[c][2], o[c][3]);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 16 * 8; e += NW * 32) {      // (row, 8-column chunk)
            const int r = e >> 3, c8 = (e & 7) * 8;
            if (r >= a.Lq) continue;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w)
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] += part[(w * 16 + r) * 64 + c8 + q];
            uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            *reinterpret_cast<uint4*>(O + (int64_t)b * bs_o + (int64_t)r * ld_o + h * 64 + c8) = pk;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            bf16* op = O + (int64_t)b * bs_o + h * 64 + 8 * c + 2 * t;
            if (g < a.Lq) *reinterpret_cast<uint32_t*>(op + (int64_t)g * ld_o) = pack_bf16(o[c][0], o[c][1]);
            if (LQ16 && g + 8 < a.Lq) *reinterpret_cast<uint32_t*>(op + (int64_t)(g + 8) * ld_o) = pack_bf16(o[c][2], o[c][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// forward for 17..32 query rows (two m16 tiles): the decode-step cross-attention, where the num_beams (20) running
// beams of a user are the query rows against that user's cross K/V.  Inference only: no dropout, no LSE.
// K fragments are loaded once and feed both row tiles.
// ------------------------------------------------------------------------------------------------------------
template <int NT, int NW>
__global__ void __launch_bounds__(NW * 32)
dattn_fwd32_kernel(DAttnDev a, bf16* __restrict__ O, int64_t ld_o, int64_t bs_o) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    extern __shared__ __align__(128) uint8_t smem[];
    dattn_fwd32_body<NT, NW>(a, O, ld_o, bs_o, blockIdx.x / a.H, blockIdx.x % a.H, smem);
}

// ------------------------------------------------------------------------------------------------------------
// backward: P recomputed from LSE; dQ, dK, dV written as bf16; d(bias_rel) accumulated atomically (self-attention)
// ------------------------------------------------------------------------------------------------------------
template <int NT, int NW, bool LQ16>
__global__ void __launch_bounds__(NW * 32)
dattn_bwd_kernel(DAttnDev a, const bf16* __restrict__ dO, int64_t ld_do, int64_t bs_do, const float* __restrict__ lse,
                 bf16* __restrict__ dQ, int64_t ld_dq, int64_t bs_dq, bf16* __restrict__ dK, bf16* __restrict__ dV, int64_t ld_dkv,
                 int64_t bs_dkv, float* __restrict__ dbias_rel) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    using C = DCfg<NT, NW>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* Ks = smem;                                               // [NW*KW][64] swizzled
    float* part = reinterpret_cast<float*>(smem + C::TILE);           // [NW][16][64] dQ partials
    uint8_t* Qs = smem + C::TILE + C::ORED;                           // [16][64] swizzled
    uint8_t* dOs = Qs + 16 * 128;                                     // [16][64] swizzled
    uint8_t* PdS = dOs + 16 * 128;                                    // per warp: Pd [16][KW+8], dS [16][KW+8] (bf16)
    float* red = reinterpret_cast<float*>(PdS + NW * 2 * 16 * C::PSTRIDE);   // [NW][16]
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int Lk = a.kv_len ? a.kv_len[b] : a.Lk;
    const int64_t k_boff = a.kv_off ? (int64_t)a.kv_off[b] * a.k_ld : (int64_t)b * a.k_bs;
    const int64_t v_boff = a.kv_off ? (int64_t)a.kv_off[b] * a.v_ld : (int64_t)b * a.v_bs;
    const int64_t dkv_boff = a.kv_off ? (int64_t)a.kv_off[b] * ld_dkv : (int64_t)b * bs_dkv;
    const int kpw = min(C::KW, (((Lk + NW - 1) / NW) + 15) & ~15);
    const int key0 = warp * kpw, ntw = kpw >> 3;
    const bf16* kb = a.k + k_boff + h * 64;
    const bf16* vb = a.v + v_boff + h * 64;
    const bf16* qb = a.q + (int64_t)b * a.q_bs + h * 64;
    const bf16* gb = dO + (int64_t)b * bs_do + h * 64;

    stage_rows(Ks, warp * C::KW, kb, a.k_ld, key0, kpw, Lk, lane);      // K rows of this warp (B operand of dQ = dS K)
    for (int e = threadIdx.x; e < 2 * 16 * 8; e += NW * 32) {             // Q and dO tiles (B operands of dK, dV)
        const int which = e >> 7, r = (e >> 3) & 15, c = e & 7;
        uint8_t* dst = (which ? dOs : Qs) + swz(r, c);
        const bf16* src = which ? gb + (int64_t)r * ld_do : qb + (int64_t)r * a.q_ld;
        if (r < a.Lq) cp_async16(smem_u32(dst), src + c * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }

    uint32_t alo[8], ahi[8];
    float s[NT][4], dp[NT][4];
    ld_row16(alo, qb + (int64_t)g * a.q_ld + 16 * t, g < a.Lq);
    if (LQ16) ld_row16(ahi, qb + (int64_t)(g + 8) * a.q_ld + 16 * t, g + 8 < a.Lq);
    qk_tiles<NT, LQ16>(s, alo, ahi, kb, a.k_ld, key0, ntw, Lk, g, t);               // S = Q K^T
    ld_row16(alo, gb + (int64_t)g * ld_do + 16 * t, g < a.Lq);
    if (LQ16) ld_row16(ahi, gb + (int64_t)(g + 8) * ld_do + 16 * t, g + 8 < a.Lq);
    qk_tiles<NT, LQ16>(dp, alo, ahi, vb, a.v_ld, key0, ntw, Lk, g, t);              // dPd = dO V^T

    ScoreCtx sc{a.Lq, Lk, a.Lk, a.causal, a.bias_off, a.n_delta, a.bias_rel ? a.bias_rel + h * a.n_delta : nullptr,
                a.key_mask ? a.key_mask + (int64_t)b * a.Lk : nullptr};
    float l[2] = {0.f, 0.f};
    if (g < a.Lq) l[0] = lse[((int64_t)b * a.H + h) * a.Lq + g];
    if (LQ16 && g + 8 < a.Lq) l[1] = lse[((int64_t)b * a.H + h) * a.Lq + g + 8];
    // p, pd (-> dp slot keeps dP), delta
    float dlt[2] = {0.f, 0.f};
    uint8_t* Pd_w = PdS + warp * 2 * 16 * C::PSTRIDE;
    uint8_t* dS_w = Pd_w + 16 * C::PSTRIDE;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        float pd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r = g + 8 * (e >> 1), j = key0 + 8 * nt + 2 * t + (e & 1);
            float p = 0.f, gdp = 0.f;
            pd[e] = 0.f;
            if (nt < ntw && (LQ16 || e < 2) && score_valid(sc, r, j)) {
                p = __expf(s[nt][e] + (sc.bias ? sc.bias[bias_index(sc, r, j)] : 0.f) - l[e >> 1]);
                gdp = dp[nt][e];
                pd[e] = p;
                if (a.drop.thr) {
                    const uint64_t idx = (((uint64_t)b * a.H + h) * a.Lq + r) * (uint64_t)Lk + j;
                    const bool keep = drop_keep(a.drop.seed, a.drop.site, idx, a.drop.thr);
                    pd[e] = keep ? p * a.drop.inv_keep : 0.f;
                    gdp = keep ? gdp * a.drop.inv_keep : 0.f;
                }
            }
            s[nt][e] = p;
            dp[nt][e] = gdp;
            dlt[e >> 1] += p * gdp;
        }
        if (nt < ntw) {     // dropped probabilities -> per-warp tile (A operand of dV = Pd^T dO)
            *reinterpret_cast<uint32_t*>(Pd_w + g * C::PSTRIDE + (8 * nt + 2 * t) * 2) = pack_bf16(pd[0], pd[1]);
            *reinterpret_cast<uint32_t*>(Pd_w + (g + 8) * C::PSTRIDE + (8 * nt + 2 * t) * 2) = pack_bf16(pd[2], pd[3]);
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        dlt[q] += __shfl_xor_sync(0xffffffffu, dlt[q], 1);
        dlt[q] += __shfl_xor_sync(0xffffffffu, dlt[q], 2);
    }
    if (NW > 1) {
        if (t == 0) { red[warp * 16 + g] = dlt[0]; red[warp * 16 + g + 8] = dlt[1]; }
        cp_async_wait_all();
        __syncthreads();               // also publishes the Q / dO tiles
        dlt[0] = dlt[1] = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { dlt[0] += red[w * 16 + g]; dlt[1] += red[w * 16 + g + 8]; }
    } else {
        cp_async_wait_all();
        __syncwarp();
    }
    // dS = P o (dP - delta)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float ds = s[nt][e] * (dp[nt][e] - dlt[e >> 1]);
            s[nt][e] = ds;
            if (dbias_rel && ds != 0.f) {
                const int r = g + 8 * (e >> 1), j = key0 + 8 * nt + 2 * t + (e & 1);
                atomicAdd(dbias_rel + h * a.n_delta + bias_index(sc, r, j), ds);
            }
        }
        if (nt < ntw) {
            *reinterpret_cast<uint32_t*>(dS_w + g * C::PSTRIDE + (8 * nt + 2 * t) * 2) = pack_bf16(s[nt][0], s[nt][1]);
            *reinterpret_cast<uint32_t*>(dS_w + (g + 8) * C::PSTRIDE + (8 * nt + 2 * t) * 2) = pack_bf16(s[nt][2], s[nt][3]);
        }
    }
    __syncwarp();
    // ---- dQ partial = dS K (this warp's keys)
    {
        float o[8][4];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c][0] = o[c][1] = o[c][2] = o[c][3] = 0.f;
        pv_tiles<NT>(o, s, Ks, warp * C::KW, ntw, lane);
        if (NW > 1) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                *reinterpret_cast<float2*>(part + (warp * 16 + g) * 64 + 8 * c + 2 * t) = make_float2(o[c][0], o[c][1]);
                *reinterpret_cast<float2*>(part + (warp * 16 + g + 8) * 64 + 8 * c + 2 * t) = make_float2(o[c][2], o[c][3]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                bf16* op = dQ + (int64_t)b * bs_dq + h * 64 + 8 * c + 2 * t;
                if (g < a.Lq) *reinterpret_cast<uint32_t*>(op + (int64_t)g * ld_dq) = pack_bf16(o[c][0], o[c][1]);
                if (LQ16 && g + 8 < a.Lq) *reinterpret_cast<uint32_t*>(op + (int64_t)(g + 8) * ld_dq) = pack_bf16(o[c][2], o[c][3]);
            }
        }
    }
    // ---- dV = Pd^T dO, dK = dS^T Q for this warp's keys: M = keys (16 per tile), K = 16 query rows, N = 64
    {
        uint32_t bq[8][2], bg[8][2];      // B fragments of Q and dO for the 8 d-tiles (k = query row)
        const int qrow = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) {
            ldsm_x4_t(bq[2 * c2][0], bq[2 * c2][1], bq[2 * c2 + 1][0], bq[2 * c2 + 1][1], smem_u32(Qs + swz(qrow, 2 * c2 + (lane >> 4))));
            ldsm_x4_t(bg[2 * c2][0], bg[2 * c2][1], bg[2 * c2 + 1][0], bg[2 * c2 + 1][1], smem_u32(dOs + swz(qrow, 2 * c2 + (lane >> 4))));
        }
        const int arow = (lane & 7) + ((lane >> 4) & 1) * 8;         // query row of the transposed A tiles
        const int acol = ((lane >> 3) & 1) * 8;                      // key offset inside the 16-key tile
#pragma unroll
        for (int mt = 0; mt < NT / 2; ++mt) {
            if (2 * mt >= ntw) continue;
#pragma unroll
            for (int which = 0; which < 2; ++which) {               // 0: dV from Pd / dO, 1: dK from dS / Q
                const uint8_t* At = which ? dS_w : Pd_w;
                uint32_t a0, a1, a2, a3;
                ldsm_x4_t(a0, a1, a2, a3, smem_u32(At + arow * C::PSTRIDE + (16 * mt + acol) * 2));
                float o[8][4];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    o[c][0] = o[c][1] = o[c][2] = o[c][3] = 0.f;
                    if (which) mma16816(o[c], a0, a1, a2, a3, bq[c][0], bq[c][1]);
                    else mma16816(o[c], a0, a1, a2, a3, bg[c][0], bg[c][1]);
                }
                bf16* out = (which ? dK : dV) + dkv_boff + h * 64;
                const int j0 = key0 + 16 * mt + g;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (j0 < Lk) *reinterpret_cast<uint32_t*>(out + (int64_t)j0 * ld_dkv + 8 * c + 2 * t) = pack_bf16(o[c][0], o[c][1]);
                    if (j0 + 8 < Lk) *reinterpret_cast<uint32_t*>(out + (int64_t)(j0 + 8) * ld_dkv + 8 * c + 2 * t) = pack_bf16(o[c][2], o[c][3]);
                }
            }
        }
    }
    if (NW > 1) {
        __syncthreads();
        for (int e = threadIdx.x; e < 16 * 8; e += NW * 32) {
            const int r = e >> 3, c8 = (e & 7) * 8;
            if (r >= a.Lq) continue;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w)
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] += part[(w * 16 + r) * 64 + c8 + q];
            uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            *reinterpret_cast<uint4*>(dQ + (int64_t)b * bs_dq + (int64_t)r * ld_dq + h * 64 + c8) = pk;
        }
    }
}

DAttnDev to_dev(const AttnArgs& a) {
    DAttnDev d;
    d.B = a.B; d.H = a.H; d.Lq = a.Lq; d.Lk = a.Lk;
    d.q = (const bf16*)a.q.ptr; d.k = (const bf16*)a.k.ptr; d.v = (const bf16*)a.v.ptr;
    d.q_ld = a.q.ld; d.q_bs = a.q.bs; d.k_ld = a.k.ld; d.k_bs = a.k.bs; d.v_ld = a.v.ld; d.v_bs = a.v.bs;
    d.bias_rel = a.bias_rel; d.bias_off = a.bias_off; d.n_delta = a.n_delta;
    d.key_mask = a.key_mask; d.causal = a.causal; d.kv_off = a.kv_off; d.kv_len = a.kv_len; d.drop = a.drop;
    return d;
}

template <int NT, int NW, bool LQ16>
void launch_fwd(const AttnArgs& a, void* O, int64_t ld_o, int64_t bs_o, float* lse, cudaStream_t st) {
    constexpr int sm = DCfg<NT, NW>::FWD_SMEM;
    static bool set = false;
    if (!set) { P5_CUDA(cudaFuncSetAttribute(dattn_fwd_kernel<NT, NW, LQ16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); set = true; }
    launch_k(dattn_fwd_kernel<NT, NW, LQ16>, (unsigned)(a.B * a.H), NW * 32, sm, st, to_dev(a), (bf16*)O, ld_o, bs_o, lse);
    LAUNCHED();
}
template <int NT, int NW, bool LQ16>
void launch_bwd(const AttnArgs& a, const void* dO, int64_t ld_do, int64_t bs_do, const float* lse, void* dQ, int64_t ld_dq,
                int64_t bs_dq, void* dK, void* dV, int64_t ld_dkv, int64_t bs_dkv, float* dbias_rel, cudaStream_t st) {
    constexpr int sm = DCfg<NT, NW>::BWD_SMEM;
    static bool set = false;
    if (!set) { P5_CUDA(cudaFuncSetAttribute(dattn_bwd_kernel<NT, NW, LQ16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); set = true; }
    launch_k(dattn_bwd_kernel<NT, NW, LQ16>, (unsigned)(a.B * a.H), NW * 32, sm, st, to_dev(a), (const bf16*)dO, ld_do, bs_do, lse, (bf16*)dQ,
                                                                              ld_dq, bs_dq, (bf16*)dK, (bf16*)dV, ld_dkv, bs_dkv, dbias_rel);
    LAUNCHED();
}

template <int NT, int NW>
void launch_fwd32(const AttnArgs& a, void* O, int64_t ld_o, int64_t bs_o, cudaStream_t st) {
    constexpr int part = NW * 32 * 64 * 4;
    constexpr int sm = (DCfg<NT, NW>::TILE > part ? DCfg<NT, NW>::TILE : part) + NW * 32 * 2 * 4;
    static bool set = false;
    if (!set) { P5_CUDA(cudaFuncSetAttribute(dattn_fwd32_kernel<NT, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); set = true; }
    launch_k(dattn_fwd32_kernel<NT, NW>, (unsigned)(a.B * a.H), NW * 32, sm, st, to_dev(a), (bf16*)O, ld_o, bs_o);
    LAUNCHED();
}

}  // namespace

bool dattn_supported(const AttnArgs& a) {
    static const bool off = getenv("P5_NO_DATTN") != nullptr;
    if (off) return false;
    auto al = [](const AttnView& v) { return v.dtype == DT_BF16 && (v.ld % 8) == 0 && (v.bs % 8) == 0 && ((uintptr_t)v.ptr % 16) == 0; };
    return a.Lq >= 1 && a.Lq <= 16 && a.Lk >= 1 && a.Lk <= 512 && al(a.q) && al(a.k) && al(a.v) && a.row_map == nullptr &&
           a.q_pos_offset == 0;
}

// Lk <= 256: eight warps with 32 keys each instead of four with 64 (shorter per-warp chain, A/B switch P5_DATTN_NW4=1)
static bool dattn_wide() {
    static const bool narrow = getenv("P5_DATTN_NW4") != nullptr;
    return !narrow;
}
#define P5_DATTN_DISPATCH(FN, ...)                                                              \
    do {                                                                                        \
        const bool hi = a.Lq > 8;                                                               \
        if (a.Lk <= 16) { if (hi) FN<2, 1, true>(__VA_ARGS__); else FN<2, 1, false>(__VA_ARGS__); }          \
        else if (a.Lk <= 64) { if (hi) FN<2, 4, true>(__VA_ARGS__); else FN<2, 4, false>(__VA_ARGS__); }     \
        else if (a.Lk <= 256 && dattn_wide()) { if (hi) FN<4, 8, true>(__VA_ARGS__); else FN<4, 8, false>(__VA_ARGS__); }  \
        else if (a.Lk <= 256) { if (hi) FN<8, 4, true>(__VA_ARGS__); else FN<8, 4, false>(__VA_ARGS__); }    \
        else { if (hi) FN<8, 8, true>(__VA_ARGS__); else FN<8, 8, false>(__VA_ARGS__); }                     \
    } while (0)

bool dattn_infer_supported(const AttnArgs& a) {
    if (a.Lq <= 16) return dattn_supported(a) && a.drop.thr == 0;
    AttnArgs c = a;
    c.Lq = 16;
    return a.Lq <= 32 && dattn_supported(c) && a.drop.thr == 0;
}

void dattn_fwd(const AttnArgs& a, void* O, int64_t ld_o, int64_t bs_o, float* lse, cudaStream_t st) {
    if (a.B <= 0) return;
    if (a.Lq > 16) {     // inference-only two-tile variant
        P5_CHECK(dattn_infer_supported(a) && lse == nullptr, "dattn_fwd: 17..32 query rows are supported without dropout / LSE only");
        P5_CHECK((ld_o % 8) == 0 && (bs_o % 8) == 0, "dattn_fwd: output rows must be 16-byte aligned");
        if (a.Lk <= 64) launch_fwd32<2, 4>(a, O, ld_o, bs_o, st);
        else if (a.Lk <= 256 && dattn_wide()) launch_fwd32<4, 8>(a, O, ld_o, bs_o, st);
        else if (a.Lk <= 256) launch_fwd32<8, 4>(a, O, ld_o, bs_o, st);
        else launch_fwd32<8, 8>(a, O, ld_o, bs_o, st);
        return;
    }
    P5_CHECK(dattn_supported(a), "dattn_fwd: unsupported geometry");
    P5_CHECK((ld_o % 8) == 0 && (bs_o % 8) == 0, "dattn_fwd: output rows must be 16-byte aligned");
    P5_DATTN_DISPATCH(launch_fwd, a, O, ld_o, bs_o, lse, st);
}

void dattn_bwd(const AttnArgs& a, const void* dO, int64_t ld_do, int64_t bs_do, const float* lse, void* dQ, int64_t ld_dq,
               int64_t bs_dq, void* dK, void* dV, int64_t ld_dkv, int64_t bs_dkv, float* dbias_rel, cudaStream_t st) {
    if (a.B <= 0) return;
    P5_CHECK(dattn_supported(a), "dattn_bwd: unsupported geometry");
    P5_CHECK((ld_do % 8) == 0 && (bs_do % 8) == 0 && (ld_dq % 8) == 0 && (bs_dq % 8) == 0 && (ld_dkv % 8) == 0 && (bs_dkv % 8) == 0,
             "dattn_bwd: rows must be 16-byte aligned");
    P5_DATTN_DISPATCH(launch_bwd, a, dO, ld_do, bs_do, lse, dQ, ld_dq, bs_dq, dK, dV, ld_dkv, bs_dkv, dbias_rel, st);
}

}  // namespace p5
