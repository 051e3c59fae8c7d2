// Fused encoder self-attention BACKWARD on tcgen05 (bf16 operands, fp32 softmax math): neither S, P, dP nor dS ever
// touch HBM.   HF:models/t5/modeling_t5.py:308-344 differentiated by hand:
//     S = Q K^T + bias + mask,  P = softmax(S),  Pd = dropout(P),  O = Pd V
//     dPd = dO V^T,  dP = dropout'(dPd),  dS = P o (dP - delta),  delta_i = sum_c dO_ic O_ic
//     dV = Pd^T dO,  dK = dS^T Q,  dQ = dS K,  d(bias)[h, j-i] += dS
// It replaces, per layer, unpack x2 + (dP GEMM -> fp32 S buffer) + softmax_bwd + three batched GEMMs + pack
// (265 us at T5-base B=64 Le=256) and makes the forward's P_save write unnecessary (P is recomputed from the row
// statistic lse2 = m2 + log2(l) that the forward stores: 4 B per row instead of 2 B per score).
//
// One persistent CTA per SM walks (batch, head) pairs.  Sequence length <= 256 (BIG = false): two 128-row query tiles,
// two 128-key blocks; Q, K, V, dO of the pair are TMA-loaded once into 128B-swizzled smem (128 KB) and serve as K-major
// AND MN-major UMMA operands in place (the same bytes are A of S = Q K^T and B of dK = dS^T Q).
// 256 < length <= 512 (BIG = true, T5-large / Yelp configs): tensor memory holds dQ of only two query tiles next to
// S, dPd, dV_j, dK_j, so a pair is processed as up to two UNITS of two query tiles each; a unit keeps its Q / dO tiles
// resident and streams the key blocks K_j, V_j through a two-slot ring; the second unit of a pair adds its dK / dV
// contribution onto the bf16 rows the first unit wrote (same CTA, same thread, so plain read-modify-write).
//   warp 0    : TMA producer
//   warp 1    : MMA issuer.  per (key block j, query tile i):  S = Q_i K_j^T, dPd = dO_i V_j^T  (128x128x64 each)
//               then dV_j += Pd^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j  (128x64x128 each, Pd / dS read from smem)
//   warp 2    : TMEM allocator: S 128 + dPd 128 + dV 64 + dK 64 + dQ 2 x 64 = 512 columns
//   warp 3    : bias (log2 domain, four shifted copies for LDS.128) and key-mask tables per pair
//   warps 4-19: four softmax warpgroups; thread = (query row, 32 of the 128 key columns).  P = 2^(s2 - lse2), dropout
//               mask regenerated from the counter hash, dS, bf16 Pd / dS tiles written to swizzled smem (the same
//               physical tile is the MN-major A of dV / dK and the K-major A of dQ); the relative-bias gradient is
//               summed along diagonals with a systolic shuffle (one accumulator per lane slides down the rows, so
//               j - i stays constant) and lands in shared memory with 2 atomics per lane per 32x32 block.
#include "kernels.cuh"
#include "tc_ptx.cuh"
#include <float.h>
#include <type_traits>

namespace p5 {
extern int g_launches;

namespace {

constexpr int NWG = 4;
constexpr int FB_THREADS = 128 + NWG * 128;
constexpr int QT = 128, KB = 128;
constexpr float LOG2E_F = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

struct FbParams {
    int B, H, L;
    const float* bias_rel;     // [H, 2L-1]
    const int* key_mask;       // [B, L] or null
    const float* row_lse2;     // [B, H, L]  m2 + log2(l) of the forward (log2 domain)
    const bf16* ctx;           // forward output O, rows as qkv
    const bf16* dctx;          // dO
    int64_t ld_ctx, ld_dctx;
    bf16* dqkv;                // dQ | dK | dV column blocks (A wide each), rows as qkv
    int64_t ld_dqkv;
    int A;
    float* dbias_rel;          // [H, 2L-1] accumulated atomically, may be null
    const int* offs;           // packed rows: first row of batch b; null = padded [B, L]
    const int* lens;
    uint32_t sQ, sdO, sK, sV, sPd, sdS, sBias, sMask, sStat, sDb, sScr, sBar, bias_cs;
    DropCfg drop;
};

template <bool BIG, bool DROP>
__global__ void __launch_bounds__(FB_THREADS, 1)
fattn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ FbParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base + P.sQ, sdO = base + P.sdO, sK = base + P.sK, sV = base + P.sV, sPd = base + P.sPd,
                   sdS = base + P.sdS, bar = base + P.sBar;
    float* bias_s = reinterpret_cast<float*>(gbase + P.sBias);
    float* mask_s = reinterpret_cast<float*>(gbase + P.sMask);
    float* rows_s = reinterpret_cast<float*>(gbase + P.sStat);     // [2 buffers][delta 256 | lse2 256]
    float* sdb = reinterpret_cast<float*>(gbase + P.sDb);          // [2L] relative-bias gradient of the current pair
    float* scr_all = reinterpret_cast<float*>(gbase + P.sScr);     // [16 warps][32] parked diagonal sums
    const uint32_t ld_full = bar, ld_empty = bar + 8, sdp_full = bar + 32, sdp_free = bar + 40, pds_full = bar + 48,
                   pds_free = bar + 56, dvk_full = bar + 64, dvk_free = bar + 72, dq_full = bar + 80, dq_free = bar + 88,
                   tmem_holder = bar + 128;
    auto bm_full = [&](int i) { return bar + 96 + 8 * i; };      // two table buffers: warp 3 works one pair ahead
    auto bm_empty = [&](int i) { return bar + 112 + 8 * i; };
    auto kv_full = [&](int i) { return bar + 144 + 8 * i; };     // BIG: two-slot ring of (K_j, V_j) key blocks
    auto kv_empty = [&](int i) { return bar + 160 + 8 * i; };
    constexpr int MASK_FLOATS = (BIG ? 4 : 2) * KB;              // key-mask table of one pair
    volatile uint32_t* tmem_holder_ptr = reinterpret_cast<volatile uint32_t*>(gbase + P.sBar + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_pairs = P.B * P.H, L = P.L;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmdO); }
    if (warp == 1 && lane == 0) {
        mbar_init(ld_full, 1); mbar_init(ld_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bm_full(i), 2); mbar_init(bm_empty(i), 4 * NWG); }   // bm_full: warps 2 and 3
        mbar_init(sdp_full, 1); mbar_init(sdp_free, 4 * NWG);
        mbar_init(pds_full, 4 * NWG); mbar_init(pds_free, 1);
        mbar_init(dvk_full, 1); mbar_init(dvk_free, 4 * NWG);
        mbar_init(dq_full, 1); mbar_init(dq_free, 4 * NWG);
        if constexpr (BIG)
            for (int i = 0; i < 2; ++i) { mbar_init(kv_full(i), 1); mbar_init(kv_empty(i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_holder, 512);
    if (warp >= 4)
        for (int e = threadIdx.x - 128; e < 2 * L; e += NWG * 128) sdb[e] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder_ptr;
    pdl_wait();               // PDL: the set-up above overlaps the previous kernel's tail (see common.cuh)
    pdl_launch_dependents();
    const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 320;
    auto tdQ = [&](int i) { return tmem + 384u + 64u * (uint32_t)i; };

    if (warp == 0) {
        // ========================= TMA producer =========================
        if (lane == 0) {
            uint32_t ph = 0;
            [[maybe_unused]] uint32_t kvc = 0;      // BIG: key blocks loaded so far (slot = kvc & 1, phase = (kvc >> 1) & 1)
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                const int b = pair / P.H, h = pair % P.H;
                const int len = P.lens ? P.lens[b] : L;
                const int row0 = P.offs ? P.offs[b] : 0, bc = P.offs ? 0 : b;
                const int nt = (len + QT - 1) / QT;
                if constexpr (!BIG) {
                    mbar_wait(ld_empty, ph ^ 1);
                    mbar_expect_tx(ld_full, (uint32_t)(4 * nt * QT * 128));
                    for (int t = 0; t < nt; ++t) {
                        tma_load_4d(sQ + t * (QT * 128), &tmQ, ld_full, 0, row0 + t * QT, h, bc);
                        tma_load_4d(sdO + t * (QT * 128), &tmdO, ld_full, 0, row0 + t * QT, h, bc);
                        tma_load_4d(sK + t * (KB * 128), &tmK, ld_full, 0, row0 + t * KB, h, bc);
                        tma_load_4d(sV + t * (KB * 128), &tmV, ld_full, 0, row0 + t * KB, h, bc);
                    }
                    ph ^= 1;
                } else {
                    for (int i0 = 0; i0 < nt; i0 += 2) {          // unit = two query tiles of the pair
                        const int ni = min(2, nt - i0);
                        mbar_wait(ld_empty, ph ^ 1);
                        mbar_expect_tx(ld_full, (uint32_t)(2 * ni * QT * 128));
                        for (int t = 0; t < ni; ++t) {
                            tma_load_4d(sQ + t * (QT * 128), &tmQ, ld_full, 0, row0 + (i0 + t) * QT, h, bc);
                            tma_load_4d(sdO + t * (QT * 128), &tmdO, ld_full, 0, row0 + (i0 + t) * QT, h, bc);
                        }
                        ph ^= 1;
                        for (int j = 0; j < nt; ++j) {            // every key block of the pair streams past the unit
                            const uint32_t slot = kvc & 1u;
                            mbar_wait(kv_empty(slot), ((kvc >> 1) & 1u) ^ 1u);
                            mbar_expect_tx(kv_full(slot), (uint32_t)(2 * KB * 128));
                            tma_load_4d(sK + slot * (KB * 128), &tmK, kv_full(slot), 0, row0 + j * KB, h, bc);
                            tma_load_4d(sV + slot * (KB * 128), &tmV, kv_full(slot), 0, row0 + j * KB, h, bc);
                            ++kvc;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ========================= MMA issuer =========================
        if (lane == 0) {
            const uint32_t f32bf16 = (1u << 4) | (1u << 7) | (1u << 10);
            const uint32_t idesc_s = f32bf16 | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);                       // K-major x K-major
            const uint32_t idesc_g = f32bf16 | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // MN x MN
            const uint32_t idesc_q = f32bf16 | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);              // K x MN
            uint32_t ld_ph = 0, sdpf_ph = 0, pdsf_ph = 0, dvkf_ph = 0, dqf_ph = 0;
            // BIG: key blocks whose S / dPd MMAs have been started (kv_i) and whose last MMA has been issued (kv_u); block n
            // of the stream sits in ring slot n & 1 (the producer's kvc counts the same sequence)
            [[maybe_unused]] uint32_t kv_i = 0, kv_u = 0;
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                const int len = P.lens ? P.lens[pair / P.H] : L;
                const int nt = (len + QT - 1) / QT;
                // unit = the query tiles [i0, i0 + ni) of the pair against all nt key blocks (one unit when !BIG)
                for (int i0 = 0; i0 < (BIG ? nt : 1); i0 += 2) {
                    const int ni = BIG ? min(2, nt - i0) : nt, T = nt * ni;
                    mbar_wait(ld_full, ld_ph);
                    ld_ph ^= 1;
                    tc_fence_after();
                    auto issue_sdp = [&](int t) {
                        const int j = t / ni, i = t % ni;
                        uint32_t ks = (uint32_t)j;              // smem tile of K_j / V_j
                        if constexpr (BIG) {
                            ks = kv_i & 1u;
                            if (i == 0) mbar_wait(kv_full(ks), (kv_i >> 1) & 1u);      // K_j, V_j have landed
                        }
                        mbar_wait(sdp_free, sdpf_ph ^ 1);       // the softmax warps have read the previous S / dPd out of TMEM
                        sdpf_ph ^= 1;
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tS, make_smem_desc(sQ + i * (QT * 128) + k * 32, 16, 1024),
                                      make_smem_desc(sK + ks * (KB * 128) + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tdP, make_smem_desc(sdO + i * (QT * 128) + k * 32, 16, 1024),
                                      make_smem_desc(sV + ks * (KB * 128) + k * 32, 16, 1024), idesc_s, k != 0);
                        umma_commit(sdp_full);
                        if constexpr (BIG) { if (i == ni - 1) ++kv_i; }
                    };
                    issue_sdp(0);
                    for (int t = 0; t < T; ++t) {
                        const int j = t / ni, i = t % ni;
                        if (t + 1 < T) issue_sdp(t + 1);
                        mbar_wait(pds_full, pdsf_ph);           // Pd / dS tiles of (j, i) are in smem
                        pdsf_ph ^= 1;
                        if (i == 0) { mbar_wait(dvk_free, dvkf_ph ^ 1); dvkf_ph ^= 1; }      // previous dV / dK read out
                        if (t == 0) { mbar_wait(dq_free, dqf_ph ^ 1); dqf_ph ^= 1; }         // previous unit's dQ read out
                        tc_fence_after();
                        const uint32_t ks = BIG ? (kv_u & 1u) : (uint32_t)j;
#pragma unroll
                        for (int k = 0; k < 8; ++k)       // dV_j += Pd^T dO_i : A = Pd tile MN-major (k = query row), B = dO_i MN-major
                            umma_bf16(tdV, make_smem_desc(sPd + k * 2048, 16384, 1024),
                                      make_smem_desc(sdO + i * (QT * 128) + k * 2048, 8192, 1024), idesc_g, (i | k) != 0);
#pragma unroll
                        for (int k = 0; k < 8; ++k)       // dK_j += dS^T Q_i
                            umma_bf16(tdK, make_smem_desc(sdS + k * 2048, 16384, 1024),
                                      make_smem_desc(sQ + i * (QT * 128) + k * 2048, 8192, 1024), idesc_g, (i | k) != 0);
                        // dV_j / dK_j are complete with the MMAs above: signal the softmax warps (which wait for them before the
                        // read-out) without the dQ MMAs in between
                        if (i == ni - 1) umma_commit(dvk_full);
#pragma unroll
                        for (int k = 0; k < 8; ++k)       // dQ_i += dS K_j : A = dS tile K-major (k = key), B = K_j MN-major
                            umma_bf16(tdQ(i), make_smem_desc(sdS + (k >> 2) * (QT * 128) + (k & 3) * 32, 16, 1024),
                                      make_smem_desc(sK + ks * (KB * 128) + k * 2048, 8192, 1024), idesc_q, (j | k) != 0);
                        umma_commit(pds_free);
                        if (i == ni - 1) {
                            if constexpr (BIG) {          // every MMA that reads K_j / V_j has been issued: the slot is
                                umma_commit(kv_empty(kv_u & 1u));   // refilled once they retire
                                ++kv_u;
                            }
                        }
                    }
                    umma_commit(dq_full);
                    umma_commit(ld_empty);
                }
            }
        }
        __syncwarp();
    } else if (warp == 2 || warp == 3) {
        // ========================= per-pair tables, one pair AHEAD of the softmax warps (double buffered) ==========
        // warp 3: bias copies + key mask; warp 2 (idle once tensor memory is allocated): the delta / lse rows.  Both are
        // chains of global-load round trips; split over two warps the tables of a CTA's FIRST pair (which nothing hides)
        // arrive in half the time.
        //  * bias (log2 domain, four shifted copies for LDS.128) and key mask: see fattn.cu
        //  * rows_s[0][i] = delta_i = sum_c dO_ic O_ic, rows_s[1][i] = lse2_i (+inf for rows past the sequence end, so
        //    that P = 2^(x - lse2) is an exact 0 there without a select).  The global-load latency is off the softmax
        //    warps' critical path.
        uint32_t bm_ph = 0;
        const int n_delta = 2 * L - 1;
        const int cs = (int)P.bias_cs;
        int buf = 0;
        for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const int b = pair / P.H, h = pair % P.H;
            const int len = P.lens ? P.lens[b] : L;
            const int nt = (len + KB - 1) / KB;
            const int64_t row0 = P.offs ? (int64_t)P.offs[b] : (int64_t)b * L;
            // BIG: one table buffer (shared memory), rebuilt per unit = two query tiles starting at tile u0
            for (int u0 = 0; u0 < (BIG ? nt : 1); u0 += 2) {
            const int nu = BIG ? min(2, nt - u0) : nt;
            float* bias_b = bias_s + buf * 4 * cs;
            float* mask_b = mask_s + buf * MASK_FLOATS;
            float* rows_b = rows_s + buf * 4 * QT;
            mbar_wait(bm_empty(buf), ((bm_ph >> buf) & 1u) ^ 1u);
            if (warp == 3) {
#pragma unroll 4
            for (int e = lane; e < L + nt * KB + 4; e += 32) {
                const float v = (P.bias_rel && e < n_delta) ? P.bias_rel[h * n_delta + e] * LOG2E_F : 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (e >= c) bias_b[c * cs + e - c] = v;
            }
            for (int j = lane; j < nt * KB; j += 32)
                mask_b[j] = (j < len && (!P.key_mask || P.key_mask[b * L + j] != 0)) ? 0.f : -INFINITY;
            } else {
            // one row per lane (128 contiguous bytes of O and of dO each): 8 independent 16-byte loads in flight per
            // tensor and lane, 32 rows per step -> the whole pair costs a handful of memory round trips
            for (int i0 = 0; i0 < nu * QT; i0 += 32) {
                const int li = i0 + lane;              // row inside the unit
                const int gi = u0 * QT + li;           // query position
                float part = 0.f, lse_v = INFINITY;
                if (gi < len) {
                    const uint4* po = reinterpret_cast<const uint4*>(P.ctx + (row0 + gi) * P.ld_ctx + h * 64);
                    const uint4* pg = reinterpret_cast<const uint4*>(P.dctx + (row0 + gi) * P.ld_dctx + h * 64);
                    lse_v = P.row_lse2[((int64_t)b * P.H + h) * L + gi];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint4 o[4], g[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) { o[q] = po[4 * hh + q]; g[q] = pg[4 * hh + q]; }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            part += bf_lo(o[q].x) * bf_lo(g[q].x) + bf_hi(o[q].x) * bf_hi(g[q].x) + bf_lo(o[q].y) * bf_lo(g[q].y) +
                                    bf_hi(o[q].y) * bf_hi(g[q].y) + bf_lo(o[q].z) * bf_lo(g[q].z) + bf_hi(o[q].z) * bf_hi(g[q].z) +
                                    bf_lo(o[q].w) * bf_lo(g[q].w) + bf_hi(o[q].w) * bf_hi(g[q].w);
                    }
                }
                rows_b[li] = part;
                rows_b[2 * QT + li] = lse_v;
            }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bm_full(buf));
            bm_ph ^= 1u << buf;
            if constexpr (!BIG) buf ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ========================= softmax warps =========================
        const int wg = (warp - 4) >> 2, sw = warp & 3;
        const int r = sw * 32 + lane;                     // row inside a 128-row tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(sw * 32) << 16;
        const int st_tid = threadIdx.x - 128;             // 0 .. 511
        float* scr = scr_all + (warp - 4) * 32;
        uint32_t bm_ph = 0, sdp_ph = 0, pdsf_ph = 0, dvk_ph = 0, dq_ph = 0;
        const uint32_t thr_hi = P.drop.thr & 0xffff0000u;       // keep iff 16-bit field >= thr16, compared in place
        const float ik = P.drop.inv_keep;
        const DropKey dkey = drop_key(P.drop.seed, P.drop.site);   // once per kernel, not per hash
        const int cs = (int)P.bias_cs;
        const int n_delta = 2 * L - 1;
        int buf = 0;
        // sequence length / first row of the NEXT pair are fetched one pair ahead (a global-load round trip otherwise opens
        // every pair of every softmax warp)
        int len_nx = L;
        int64_t row0_nx = 0;
        auto fetch_geom = [&](int pr) {
            if (pr < n_pairs) {
                const int bb = pr / P.H;
                len_nx = P.lens ? P.lens[bb] : L;
                row0_nx = P.offs ? (int64_t)P.offs[bb] : (int64_t)bb * L;
            }
        };
        fetch_geom((int)blockIdx.x);
        for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const int b = pair / P.H, h = pair % P.H;
            const int len = len_nx;
            const int nt = (len + QT - 1) / QT;
            const int64_t row0 = row0_nx;
            fetch_geom(pair + (int)gridDim.x);
            // unit = the query tiles [u0, u0 + nu) of the pair against all nt key blocks (one unit when !BIG)
            for (int u0 = 0; u0 < (BIG ? nt : 1); u0 += 2) {
            const int nu = BIG ? min(2, nt - u0) : nt;
            const float* bias_b = bias_s + buf * 4 * cs;
            const float* mask_b = mask_s + buf * MASK_FLOATS;
            const float* rows_b = rows_s + buf * 4 * QT;
            mbar_wait(bm_full(buf), (bm_ph >> buf) & 1u);
            bm_ph ^= 1u << buf;
            const float delta0 = rows_b[r], delta1 = rows_b[QT + r], lse0 = rows_b[2 * QT + r], lse1 = rows_b[3 * QT + r];

            for (int j = 0; j < nt; ++j) {
                for (int i = 0; i < nu; ++i) {
                    const int gi = (u0 + i) * QT + r;
                    const bool row_ok = gi < len;
                    const int j0 = j * KB + wg * 32;
                    const int o = (row_ok ? (L - 1 - gi) : 0) + j0;
                    const float4* b4 = reinterpret_cast<const float4*>(bias_b + (o & 3) * cs + (o & ~3));
                    const float4* m4 = reinterpret_cast<const float4*>(mask_b + j0);
                    const bool full = (j0 + 32 <= len) && !P.key_mask;       // warp-uniform
                    const uint32_t pair0 = (uint32_t)((uint64_t)((((int64_t)b * P.H + h) * L + gi) * L + j0) >> 1);
                    const float my_lse = i ? lse1 : lse0, my_delta = i ? delta1 : delta0;
                    // diagonal index of (row of lane 0, column j0): entry + (t - lane) is the bias slot of element (lane, t)
                    const int diag0 = (j0 - ((u0 + i) * QT + sw * 32)) + (L - 1);
                    float dacc[2] = {0.f, 0.f};
                    mbar_wait(sdp_full, sdp_ph);
                    sdp_ph ^= 1;
                    tc_fence_after();
                    // two straight-line copies of the block (FULL: no key-mask loads / adds) instead of predicated-off instructions
                    auto block = [&](auto full_c) {
                    constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t vs[16], vd[16];
                        tmem_ld16(tS + lane_off + wg * 32 + hf * 16, vs);
                        tmem_ld16(tdP + lane_off + wg * 32 + hf * 16, vd);
                        tmem_ld_wait();
                        if (hf == 1) {           // S / dPd fully read: the MMA warp may overwrite them
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(sdp_free);
                        }
                        uint32_t pk[2][8];       // [8-column group][Pd x4 | dS x4] packed bf16 pairs
#pragma unroll
                        for (int q2 = 0; q2 < 2; ++q2) {
                            const int q = 2 * hf + q2;               // 8-column group of this thread's 32 columns
                            float ds[8];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const float4 bb = b4[2 * q + u];
                                float x[4];
                                x[0] = fmaf(__uint_as_float(vs[8 * q2 + 4 * u]), LOG2E_F, bb.x);
                                x[1] = fmaf(__uint_as_float(vs[8 * q2 + 4 * u + 1]), LOG2E_F, bb.y);
                                x[2] = fmaf(__uint_as_float(vs[8 * q2 + 4 * u + 2]), LOG2E_F, bb.z);
                                x[3] = fmaf(__uint_as_float(vs[8 * q2 + 4 * u + 3]), LOG2E_F, bb.w);
                                if constexpr (!FULL) {
                                    const float4 mm = m4[2 * q + u];
                                    x[0] += mm.x; x[1] += mm.y; x[2] += mm.z; x[3] += mm.w;
                                }
#pragma unroll
                                for (int e2 = 0; e2 < 2; ++e2) {       // one hash per aligned pair of columns
                                    const float p0 = ex2_approx(x[2 * e2] - my_lse), p1 = ex2_approx(x[2 * e2 + 1] - my_lse);
                                    const float g0 = __uint_as_float(vd[8 * q2 + 4 * u + 2 * e2]), g1 = __uint_as_float(vd[8 * q2 + 4 * u + 2 * e2 + 1]);
                                    float m0 = 1.f, m1 = 1.f;          // dropout multiplier: 1 / keep or 0
                                    if constexpr (DROP) {
                                        const uint32_t hsh = drop_hash_k(dkey, pair0 + 4 * q + 2 * u + e2);
                                        m0 = (hsh << 16) >= thr_hi ? ik : 0.f;
                                        m1 = hsh >= thr_hi ? ik : 0.f;
                                    }
                                    const float d0 = p0 * m0, d1 = p1 * m1;
                                    ds[4 * u + 2 * e2] = p0 * fmaf(g0, m0, -my_delta);
                                    ds[4 * u + 2 * e2 + 1] = p1 * fmaf(g1, m1, -my_delta);
                                    pk[q2][2 * u + e2] = pack2(d0, d1);
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) pk[q2][4 + u] = pack2(ds[2 * u], ds[2 * u + 1]);
                            if (P.dbias_rel) {
                                // systolic diagonal sum over this half's 32 rows x 16 columns: the accumulator that sits
                                // on lane l after local column t holds diagonal t - l.  What falls off lane 31 is parked in
                                // a per-warp scratch row with a plain store (no atomics on the shuffle chain).
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const int t = 8 * q2 + e;                   // column inside the half
                                    if (t > 0 && lane == 31) scr[hf * 16 + t - 1] = dacc[hf];   // diagonal (t - 1) - 31
                                    dacc[hf] = __shfl_up_sync(0xffffffffu, dacc[hf], 1);
                                    if (lane == 0) dacc[hf] = 0.f;
                                    dacc[hf] += ds[e];
                                }
                            }
                        }
                        if (hf == 0) {           // all the arithmetic above overlapped the MMAs that still read the previous tiles
                            mbar_wait(pds_free, pdsf_ph ^ 1);
                            pdsf_ph ^= 1;
                        }
                        // bf16 tiles: [64-key chunk (wg >> 1)][128 rows][128 B], 16-byte units XOR (row & 7)
#pragma unroll
                        for (int q2 = 0; q2 < 2; ++q2) {
                            const int q = 2 * hf + q2;
                            const uint32_t off = (uint32_t)((wg >> 1) * (QT * 128) + r * 128 + ((((wg & 1) * 4 + q) ^ (r & 7)) << 4));
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPd + off), "r"(pk[q2][0]), "r"(pk[q2][1]),
                                         "r"(pk[q2][2]), "r"(pk[q2][3]) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdS + off), "r"(pk[q2][4]), "r"(pk[q2][5]),
                                         "r"(pk[q2][6]), "r"(pk[q2][7]) : "memory");
                        }
                    }
                    };
                    if (full) block(std::true_type{}); else block(std::false_type{});
                    // Pd / dS tiles complete: make the generic-proxy smem writes visible to the tensor core
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pds_full);
                    if (P.dbias_rel) {
                        // half hf covers columns 16 hf .. 16 hf + 15: lane l ends on diagonal 15 - l (+ 16 hf); the parked
                        // values are diagonals (t - 1) - 31 for t = 1 .. 15
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            if (dacc[hf] != 0.f) atomicAdd(&sdb[diag0 + 16 * hf + 15 - lane], dacc[hf]);
                            if (lane < 15) {
                                const float v = scr[hf * 16 + lane];
                                if (v != 0.f) atomicAdd(&sdb[diag0 + 16 * hf + lane - 31], v);
                            }
                        }
                        __syncwarp();
                    }
                    if (i == nu - 1) {
                        // ---- dV_j, dK_j complete: rows = keys of block j, this warpgroup writes 16 of the 64 columns
                        //      (BIG: complete over this unit's query tiles; the pair's second unit adds onto the first's rows)
                        mbar_wait(dvk_full, dvk_ph);
                        dvk_ph ^= 1;
                        tc_fence_after();
                        uint32_t v[16], w[16];
                        tmem_ld16(tdV + lane_off + wg * 16, v);
                        tmem_ld16(tdK + lane_off + wg * 16, w);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(dvk_free);
                        const int gj = j * KB + r;
                        if (gj < len) {
                            bf16* dst = P.dqkv + (row0 + gj) * P.ld_dqkv + h * 64 + wg * 16;
                            if constexpr (BIG) {
                                if (u0 > 0) {      // written by this very thread while it processed the pair's first unit
#pragma unroll
                                    for (int q = 0; q < 2; ++q) {
                                        const uint4 pa = *reinterpret_cast<const uint4*>(dst + 2 * P.A + 8 * q);
                                        const uint4 pc = *reinterpret_cast<const uint4*>(dst + P.A + 8 * q);
                                        const uint32_t pav[4] = {pa.x, pa.y, pa.z, pa.w}, pcv[4] = {pc.x, pc.y, pc.z, pc.w};
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            v[8 * q + 2 * e] = __float_as_uint(__uint_as_float(v[8 * q + 2 * e]) + bf_lo(pav[e]));
                                            v[8 * q + 2 * e + 1] = __float_as_uint(__uint_as_float(v[8 * q + 2 * e + 1]) + bf_hi(pav[e]));
                                            w[8 * q + 2 * e] = __float_as_uint(__uint_as_float(w[8 * q + 2 * e]) + bf_lo(pcv[e]));
                                            w[8 * q + 2 * e + 1] = __float_as_uint(__uint_as_float(w[8 * q + 2 * e + 1]) + bf_hi(pcv[e]));
                                        }
                                    }
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                uint4 a, c;
                                a.x = pack2(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1]));
                                a.y = pack2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
                                a.z = pack2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
                                a.w = pack2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
                                c.x = pack2(__uint_as_float(w[8 * q]), __uint_as_float(w[8 * q + 1]));
                                c.y = pack2(__uint_as_float(w[8 * q + 2]), __uint_as_float(w[8 * q + 3]));
                                c.z = pack2(__uint_as_float(w[8 * q + 4]), __uint_as_float(w[8 * q + 5]));
                                c.w = pack2(__uint_as_float(w[8 * q + 6]), __uint_as_float(w[8 * q + 7]));
                                *reinterpret_cast<uint4*>(dst + 2 * P.A + 8 * q) = a;     // dV
                                *reinterpret_cast<uint4*>(dst + P.A + 8 * q) = c;         // dK
                            }
                        }
                    }
                }
            }
            // ---- dQ tiles of the unit
            mbar_wait(dq_full, dq_ph);
            dq_ph ^= 1;
            tc_fence_after();
            for (int i = 0; i < nu; ++i) {
                uint32_t v[16];
                tmem_ld16(tdQ(i) + lane_off + wg * 16, v);
                tmem_ld_wait();
                const int gi = (u0 + i) * QT + r;
                if (gi < len) {
                    bf16* dst = P.dqkv + (row0 + gi) * P.ld_dqkv + h * 64 + wg * 16;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint4 a;
                        a.x = pack2(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1]));
                        a.y = pack2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
                        a.z = pack2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
                        a.w = pack2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
                        *reinterpret_cast<uint4*>(dst + 8 * q) = a;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(dq_free); mbar_arrive(bm_empty(buf)); }
            if constexpr (!BIG) buf ^= 1;
            // ---- relative-bias gradient of this unit -> global, table cleared for the next one
            if (P.dbias_rel) {
                asm volatile("bar.sync 1, %0;" ::"n"(NWG * 128) : "memory");
                for (int e = st_tid; e < n_delta; e += NWG * 128) {
                    const float v = sdb[e];
                    if (v != 0.f) { atomicAdd(P.dbias_rel + h * n_delta + e, v); sdb[e] = 0.f; }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NWG * 128) : "memory");
            }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace

// the fused backward covers this encoder length (the forward then saves row statistics instead of probabilities)
bool fattn_bwd_supported(int L) {
    static const bool off = getenv("P5_NO_FATTN_BWD") != nullptr;
    static const bool no_big = getenv("P5_NO_FATTN_BWD_BIG") != nullptr;     // 256 < L <= 512 back on the materialised path
    return !off && L % 8 == 0 && L <= (no_big ? 256 : 512);
}

bool fattn_bwd(const void* qkv, int64_t ld_qkv, int A, int B, int H, int L, const float* bias_rel, const int* key_mask,
               const float* row_lse2, const void* ctx, int64_t ld_ctx, const void* dctx, int64_t ld_dctx, void* dqkv,
               int64_t ld_dqkv, float* dbias_rel, DropCfg drop, cudaStream_t st, const int* offs, const int* lens,
               int64_t packed_rows) {
    if (!fattn_bwd_supported(L) || ld_qkv % 8 != 0 || ld_ctx % 8 != 0 || ld_dctx % 8 != 0 || ld_dqkv % 8 != 0 || !row_lse2)
        return false;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        P5_CUDA(cudaGetDevice(&dev));
        P5_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const bool big = L > 256;
    FbParams P;
    P.B = B; P.H = H; P.L = L;
    const uint32_t tile = QT * 128;                      // 16 KB: 128 rows x 64 bf16
    // big: sQ / sdO hold the unit's two query tiles, sK / sV are the two slots of the key-block ring
    P.sQ = 0; P.sdO = 2 * tile; P.sK = 4 * tile; P.sV = 6 * tile; P.sPd = 8 * tile; P.sdS = 10 * tile;
    P.sBias = 12 * tile;
    P.bias_cs = (uint32_t)(((2 * L + 4 + 31) & ~31) + 8);
    const uint32_t ntbl = big ? 1 : 2;                   // table buffers (bias copies + key mask): one is all that fits at L > 256
    P.sMask = P.sBias + ntbl * (uint32_t)round_up(4 * P.bias_cs * 4, 16);
    P.sStat = P.sMask + ntbl * (big ? 4 : 2) * KB * 4;
    P.sDb = P.sStat + 2 * 4 * QT * 4;
    P.sScr = P.sDb + (uint32_t)round_up(2 * L * 4, 16);
    P.sBar = P.sScr + NWG * 4 * 32 * 4;
    const size_t smem = P.sBar + 256 + 1024;
    P5_CHECK(smem <= 232448, "fattn_bwd: shared memory budget exceeded");
    P.bias_rel = bias_rel; P.key_mask = offs ? nullptr : key_mask; P.row_lse2 = row_lse2;
    P.ctx = (const bf16*)ctx; P.dctx = (const bf16*)dctx; P.ld_ctx = ld_ctx; P.ld_dctx = ld_dctx;
    P.dqkv = (bf16*)dqkv; P.ld_dqkv = ld_dqkv; P.A = A; P.dbias_rel = dbias_rel;
    P.offs = offs; P.lens = lens; P.drop = drop;
    // compile-time variants: {Le <= 256, 256 < Le <= 512} x {dropout off, on}
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FbParams);
    static const KernelFn variants[4] = {fattn_bwd_kernel<false, false>, fattn_bwd_kernel<false, true>,
                                         fattn_bwd_kernel<true, false>, fattn_bwd_kernel<true, true>};
    const int vi = (big ? 2 : 0) + (drop.thr ? 1 : 0);
    static size_t max_set[4] = {0, 0, 0, 0};
    if (smem > max_set[vi]) {
        P5_CUDA(cudaFuncSetAttribute(variants[vi], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set[vi] = smem;
    }
    const uint64_t rows = (uint64_t)(offs ? packed_rows : L);
    const uint64_t dims[4] = {64, rows, (uint64_t)H, (uint64_t)(offs ? 1 : B)};
    const uint32_t box[4] = {64, 128, 1, 1};
    const uint64_t strides[3] = {(uint64_t)ld_qkv * 2, 128, rows * (uint64_t)ld_qkv * 2};
    const uint64_t strides_g[3] = {(uint64_t)ld_dctx * 2, 128, rows * (uint64_t)ld_dctx * 2};
    const bf16* base = (const bf16*)qkv;
    CUtensorMap tmQ = tmap_bf16_4d(base, dims, strides, box);
    CUtensorMap tmK = tmap_bf16_4d(base + A, dims, strides, box);
    CUtensorMap tmV = tmap_bf16_4d(base + 2 * A, dims, strides, box);
    CUtensorMap tmdO = tmap_bf16_4d((const bf16*)dctx, dims, strides_g, box);
    const int n_pairs = B * H;
    const int budget = sm_budget() < num_sms ? sm_budget() : num_sms;     // leaves SMs to a concurrent NCCL all-reduce (common.cuh)
    const int grid = n_pairs < budget ? n_pairs : budget;
    launch_k(variants[vi], grid, FB_THREADS, smem, st, tmQ, tmK, tmV, tmdO, P);
    P5_CUDA(cudaGetLastError());
    ++g_launches;
    return true;
}

}  // namespace p5
