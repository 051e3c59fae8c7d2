// Fused encoder self-attention forward on tcgen05 (bf16 operands, fp32 softmax): S = Q K^T is never written to HBM.
//   HF:models/t5/modeling_t5.py:308-334 — scores (unscaled) + shared relative position bias + additive padding mask,
//   softmax in fp32, dropout on the probabilities, P V.
//
// One persistent CTA per SM walks (batch, head) pairs; K and V of the pair ([Lk, 64] each, Lk <= 512) are loaded once
// into 128B-swizzled smem by TMA and reused by every 128-row query tile.
//   warp 0    : TMA producer (K, V per pair; Q per tile)
//   warp 1    : MMA issuer.  pass 1: S_blk = Q K_blk^T (128x128x64) per 128-key block -> TMEM (three score buffers)
//               pass 2: S_blk again, then O += P_blk V_blk (128x64x128) with P_blk read from smem
//   warp 2    : TMEM allocator (3 x 128 columns S + 64 columns O)
//   warp 3    : per-pair tables, one pair ahead (double buffered): relative bias x log2(e) in four shifted copies (every
//               thread reads its 32 entries with LDS.128), key mask
//   warps 4-19: softmax, four warpgroups.  thread = (query row, 32 of the 128 key columns of a block); four warps per
//               scheduler hide the ALU / MUFU / LDS latencies of the softmax arithmetic.  Log2 domain throughout:
//               pass 1: row max only (FFMA + FMNMX per score; the four column slices are combined through smem);
//               pass 2: p~ = 2^(s2 - m2) UN-normalised (FFMA + EX2), row sum, dropout mask -> bf16 -> swizzled smem A
//               tile; O is scaled by 1 / l (and by dropout's 1 / (1 - p)) when it is read out of TMEM.
//               The hot loops are compile-time variants (dropout on / off, probabilities saved or not, full or partial
//               key slice): a profile of the runtime-flag form showed a third of its issue slots going to the dropout
//               key re-derived per iteration and to predicated-off key-mask instructions.
// Two passes over the key blocks (QK^T is recomputed: K = 64, cheap) avoid rescaling the O accumulator in TMEM.  For the
// backward the kernel stores lse2 = m2 + log2(l) per row (fattn_bwd recomputes P, nothing of size L^2 is written), or —
// only when the fused backward is switched off — the un-normalised P plus 1 / l for the materialised GEMM chain.
#include "kernels.cuh"
#include "tc_ptx.cuh"
#include <float.h>
#include <type_traits>

namespace p5 {
extern int g_launches;
#define LOG2E 1.4426950408889634f
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

static constexpr int NWG = 4;              // softmax warpgroups: each owns 128 / NWG = 32 key columns of every block
static constexpr int FA_THREADS = 128 + NWG * 128;   // 4 control warps + NWG softmax warpgroups
static constexpr int QT = 128;      // query rows per tile
static constexpr int KB = 128;      // keys per block
static constexpr int NSB = 3;       // S = Q K^T buffers in TMEM (3 x 128 columns + 64 columns of O <= 512)

struct FaParams {
    int B, H, Lq, Lk, nkb, nqt, nq_buf, np_buf;   // nq_buf / np_buf: Q-tile and P-tile buffers that fit next to K and V
    const float* bias_rel;   // [H, Lq + Lk - 1]
    const int* key_mask;     // [B, Lk]
    bf16* P_save;            // [B, H, Lq, Lk] UN-normalised probabilities 2^(s2 - m2) (un-dropped)
    float* row_scale;        // [B, H, Lq] 1 / row sum: P = P_save * row_scale
    float* row_lse2;         // [B, H, Lq] m2 + log2(row sum): P = 2^(s2 - lse2), all the fused backward needs
    uint32_t bias_cs;        // stride (floats) between the four shifted copies of the bias row
    uint32_t ntbl, tbl_stride;   // table buffers (2 when they fit: the table warp runs one pair ahead), floats per bias buffer
    bf16* ctx;               // [B*Lq, ld_ctx]
    int64_t ld_ctx;
    uint32_t sK, sV, sQ, sP, sBias, sMask, sStat, sBar;   // smem offsets from the 1024-aligned base
    const int* offs;         // packed (padding-free) token layout: first row of batch b in qkv / ctx; null = padded [B, L]
    const int* lens;         // packed: tokens of batch b (queries == keys); rows/keys >= len are skipped / masked
    DropCfg drop;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// DROP: dropout on the probabilities; SAVEP: the un-normalised probabilities are written for the materialised backward
template <bool DROP, bool SAVEP>
__global__ void __launch_bounds__(FA_THREADS, 1)
fattn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FaParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sK = base + P.sK, sV = base + P.sV, sQ = base + P.sQ, sP = base + P.sP, bar = base + P.sBar;
    float* bias_all = reinterpret_cast<float*>(gbase + P.sBias);    // [ntbl][4 copies][bias_cs]
    float* mask_all = reinterpret_cast<float*>(gbase + P.sMask);    // [ntbl][nkb * KB]
    float* statm_s = reinterpret_cast<float*>(gbase + P.sStat);    // [NWG][128] row max per warpgroup (log2 domain)
    float* statl_s = statm_s + NWG * QT;                           // [NWG][128] row sum per warpgroup
    // barriers (8 bytes each)
    const uint32_t kv_full = bar, kv_empty = bar + 8;
    auto q_full = [&](int i) { return bar + 16 + 8 * i; };
    auto q_empty = [&](int i) { return bar + 32 + 8 * i; };
    auto s_full = [&](int i) { return bar + 48 + 8 * i; };     // NSB score buffers in TMEM (deeper MMA prefetch)
    auto s_empty = [&](int i) { return bar + 72 + 8 * i; };
    auto p_full = [&](int i) { return bar + 96 + 8 * i; };
    auto p_empty = [&](int i) { return bar + 112 + 8 * i; };
    const uint32_t o_full = bar + 128, o_empty = bar + 136;
    auto bm_full = [&](int i) { return bar + 144 + 8 * i; };     // table buffers: the table warp works one pair ahead
    auto bm_empty = [&](int i) { return bar + 160 + 8 * i; };
    const uint32_t tmem_holder = bar + 176;
    volatile uint32_t* tmem_holder_ptr = reinterpret_cast<volatile uint32_t*>(gbase + P.sBar + 176);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_pairs = P.B * P.H;
    const int Lk = P.Lk, Lq = P.Lq, nkb = P.nkb, nqt = P.nqt;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); }
    if (warp == 1 && lane == 0) {
        mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(q_full(i), 1); mbar_init(q_empty(i), 1);
            mbar_init(p_full(i), 4 * NWG); mbar_init(p_empty(i), 1);
        }
        for (int i = 0; i < NSB; ++i) { mbar_init(s_full(i), 1); mbar_init(s_empty(i), 4 * NWG); }
        mbar_init(o_full, 1); mbar_init(o_empty, 4 * NWG);
        for (int i = 0; i < 2; ++i) { mbar_init(bm_full(i), 1); mbar_init(bm_empty(i), 4 * NWG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder_ptr;
    pdl_wait();               // PDL: the set-up above overlaps the previous kernel's tail (see common.cuh)
    pdl_launch_dependents();
    auto tS = [&](int i) { return tmem + (uint32_t)i * 128u; };
    const uint32_t tO = tmem + 384;

    if (warp == 0) {
        // ========================= TMA producer =========================
        if (lane == 0) {
            uint32_t kv_ph = 0, q_ph = 0;   // phase bits, one per buffer (register-resident: no dynamically indexed arrays)
            int qb = 0;
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                const int b = pair / P.H, h = pair % P.H;
                // packed layout: the pair's rows start at offs[b] of a [Mtot, .] matrix (batch coordinate 0) and only
                // ceil(len / 128) key blocks / query tiles exist
                const int len = P.lens ? P.lens[b] : Lk;
                const int row0 = P.offs ? P.offs[b] : 0, bc = P.offs ? 0 : b;
                const int nkb = (len + KB - 1) / KB, nqt = (len + QT - 1) / QT;
                mbar_wait(kv_empty, kv_ph ^ 1);
                mbar_expect_tx(kv_full, (uint32_t)(2 * nkb * KB * 128));
                for (int kb = 0; kb < nkb; ++kb) {
                    tma_load_4d(sK + kb * (KB * 128), &tmK, kv_full, 0, row0 + kb * KB, h, bc);
                    tma_load_4d(sV + kb * (KB * 128), &tmV, kv_full, 0, row0 + kb * KB, h, bc);
                }
                kv_ph ^= 1;
                for (int qt = 0; qt < nqt; ++qt) {
                    mbar_wait(q_empty(qb), ((q_ph >> qb) & 1u) ^ 1u);
                    mbar_expect_tx(q_full(qb), QT * 128);
                    tma_load_4d(sQ + qb * (QT * 128), &tmQ, q_full(qb), 0, row0 + qt * QT, h, bc);
                    q_ph ^= 1u << qb;
                    qb = (qb + 1) % P.nq_buf;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ========================= MMA issuer =========================
        if (lane == 0) {
            // S: M=128, N=128, A (Q) K-major, B (K) K-major.   PV: M=128, N=64, A (P) K-major, B (V) MN-major.
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                                     ((uint32_t)(128 >> 4) << 24);
            uint32_t kv_ph = 0, q_ph = 0, se_ph = 0, pf_ph = 0, oe_ph = 0;   // phase bits per buffer
            int qb = 0, sb = 0, pb = 0;
            auto issue_s = [&](int kb, uint32_t q_addr) {
                mbar_wait(s_empty(sb), ((se_ph >> sb) & 1u) ^ 1u);
                se_ph ^= 1u << sb;
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tS(sb), make_smem_desc(q_addr + k * 32, 16, 1024),
                              make_smem_desc(sK + kb * (KB * 128) + k * 32, 16, 1024), idesc_s, k != 0);
                umma_commit(s_full(sb));
                sb = (sb + 1) % NSB;
            };
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                const int len = P.lens ? P.lens[pair / P.H] : Lk;
                const int nkb = (len + KB - 1) / KB, nqt = (len + QT - 1) / QT;
                mbar_wait(kv_full, kv_ph);
                kv_ph ^= 1;
                for (int qt = 0; qt < nqt; ++qt) {
                    mbar_wait(q_full(qb), (q_ph >> qb) & 1u);
                    q_ph ^= 1u << qb;
                    tc_fence_after();
                    const uint32_t q_addr = sQ + qb * (QT * 128);
                    // pass 1: statistics
                    for (int kb = 0; kb < nkb; ++kb) issue_s(kb, q_addr);
                    // pass 2: probabilities and O
                    mbar_wait(o_empty, oe_ph ^ 1);     // the previous tile's O has been read out
                    oe_ph ^= 1;
                    issue_s(0, q_addr);
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (kb + 1 < nkb) issue_s(kb + 1, q_addr);
                        else umma_commit(q_empty(qb));   // all QK^T of this tile issued: Q slot may be refilled once they retire
                        mbar_wait(p_full(pb), (pf_ph >> pb) & 1u);
                        pf_ph ^= 1u << pb;
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < KB / 16; ++k) {
                            const uint64_t da = make_smem_desc(sP + pb * (QT * KB * 2) + (k >> 2) * (QT * 128) + (k & 3) * 32, 16, 1024);
                            const uint64_t db = make_smem_desc(sV + kb * (KB * 128) + k * 2048, 8192, 1024);
                            umma_bf16(tO, da, db, idesc_o, (kb | k) != 0);
                        }
                        umma_commit(p_empty(pb));
                        pb = (pb + 1 == P.np_buf) ? 0 : pb + 1;
                    }
                    umma_commit(o_full);
                    qb = (qb + 1) % P.nq_buf;
                }
                umma_commit(kv_empty);   // K/V of this pair may be overwritten once every MMA above has retired
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ========================= bias / mask tables per (batch, head) =========================
        // Everything is kept in the log2 domain (bias * log2(e)) so that a probability is one FFMA + one EX2.
        // bias4_s holds FOUR copies of the bias row, copy c shifted by c entries: a thread whose first entry is
        // o = (Lq-1-i) + j0 reads copy (o & 3) at the 16-byte aligned index o - (o & 3) with LDS.128.
        uint32_t bm_ph = 0;
        const int n_delta = Lq + Lk - 1;
        const int cs = (int)P.bias_cs;
        int buf = 0;
        for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const int b = pair / P.H, h = pair % P.H;
            const int len = P.lens ? P.lens[b] : Lk;
            const int nkb = (len + KB - 1) / KB;
            float* bias_s = bias_all + buf * P.tbl_stride;
            float* mask_s = mask_all + buf * (P.nkb * KB);
            mbar_wait(bm_empty(buf), ((bm_ph >> buf) & 1u) ^ 1u);
            // entries past n_delta are only touched for padded keys (masked with -inf): keep them finite
            // (unrolled: four global loads in flight per lane — the table of a CTA's first pair is on the critical path)
#pragma unroll 4
            for (int e = lane; e < Lq + nkb * KB + 4; e += 32) {
                const float v = (P.bias_rel && e < n_delta) ? P.bias_rel[h * n_delta + e] * LOG2E : 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (e >= c) bias_s[c * cs + e - c] = v;
            }
            for (int j = lane; j < nkb * KB; j += 32)
                mask_s[j] = (j < len && (!P.key_mask || P.key_mask[b * Lk + j] != 0)) ? 0.f : -INFINITY;
            __syncwarp();
            if (lane == 0) mbar_arrive(bm_full(buf));
            bm_ph ^= 1u << buf;
            buf = (buf + 1 == (int)P.ntbl) ? 0 : buf + 1;
        }
    } else if (warp >= 4) {
        // ========================= softmax warps =========================
        const int wg = (warp - 4) >> 2;                   // warpgroup: key columns [wg*32, wg*32+32) of every block
        const int sw = warp & 3;
        const int r = sw * 32 + lane;                     // row of the query tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(sw * 32) << 16;
        uint32_t sf_ph = 0, pe_ph = 0, of_ph = 0, bm_ph = 0;   // phase bits per buffer
        int sb = 0, pb = 0, buf = 0;
        const uint32_t thr_hi = P.drop.thr & 0xffff0000u;    // keep iff the element's 16-bit field >= thr16, compared in place
        const DropKey dkey = drop_key(P.drop.seed, P.drop.site);   // once per kernel, not per hash
        const int cs = (int)P.bias_cs;
        // sequence length / first row of the NEXT pair are fetched one pair ahead (a global-load round trip otherwise opens
        // every pair of every softmax warp)
        int len_nx = Lq;
        int64_t row0_nx = 0;
        auto fetch_geom = [&](int pr) {
            if (pr < n_pairs) {
                const int bb = pr / P.H;
                len_nx = P.lens ? P.lens[bb] : Lq;
                row0_nx = P.offs ? (int64_t)P.offs[bb] : (int64_t)bb * Lq;
            }
        };
        fetch_geom((int)blockIdx.x);
        for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const int b = pair / P.H, h = pair % P.H;
            const int len = len_nx;
            const int nkb = (len + KB - 1) / KB, nqt = (len + QT - 1) / QT;
            const int64_t ctx_row0 = row0_nx;
            fetch_geom(pair + (int)gridDim.x);
            const float* bias_s = bias_all + buf * P.tbl_stride;
            const float* mask_s = mask_all + buf * (P.nkb * KB);
            mbar_wait(bm_full(buf), (bm_ph >> buf) & 1u);
            bm_ph ^= 1u << buf;
            for (int qt = 0; qt < nqt; ++qt) {
                const int i = qt * QT + r;                // query position
                const bool row_ok = i < len;
                const int o_row = row_ok ? (Lq - 1 - i) : 0;     // bias entry of key j is o_row + j
                const int64_t grow = (((int64_t)b * P.H + h) * Lq + i) * Lk;
                // ---------------- pass 1: m2 = max_j (s_ij + bias) * log2(e) over this warpgroup's columns
                float m2 = -INFINITY;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(s_full(sb), (sf_ph >> sb) & 1u);
                    sf_ph ^= 1u << sb;
                    tc_fence_after();
                    {
                        const int j0 = kb * KB + wg * 32;
                        const int o = o_row + j0;
                        const float4* b4 = reinterpret_cast<const float4*>(bias_s + (o & 3) * cs + (o & ~3));
                        const float4* m4 = reinterpret_cast<const float4*>(mask_s + j0);
                        const bool full = (j0 + 32 <= len) && !P.key_mask;       // warp-uniform
                        // two straight-line copies (FULL: no key-mask loads / adds) instead of predicated-off instructions
                        auto slice_max = [&](auto full_c) {
                            constexpr bool FULL = decltype(full_c)::value;
                            float cm = -INFINITY;
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {         // 16 columns at a time keeps the register count down
                                uint32_t v[16];
                                tmem_ld16(tS(sb) + lane_off + wg * 32 + hf * 16, v);
                                tmem_ld_wait();
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float4 bb = b4[4 * hf + q];
                                    float x0 = fmaf(__uint_as_float(v[4 * q]), LOG2E, bb.x), x1 = fmaf(__uint_as_float(v[4 * q + 1]), LOG2E, bb.y);
                                    float x2 = fmaf(__uint_as_float(v[4 * q + 2]), LOG2E, bb.z), x3 = fmaf(__uint_as_float(v[4 * q + 3]), LOG2E, bb.w);
                                    if constexpr (!FULL) {
                                        const float4 mm = m4[4 * hf + q];
                                        x0 += mm.x; x1 += mm.y; x2 += mm.z; x3 += mm.w;
                                    }
                                    cm = fmaxf(cm, fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)));
                                }
                            }
                            return cm;
                        };
                        const float cm = full ? slice_max(std::true_type{}) : slice_max(std::false_type{});
                        m2 = fmaxf(m2, cm);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_empty(sb));
                    sb = (sb + 1) % NSB;
                }
                // combine the column slices of the row
                statm_s[wg * QT + r] = m2;
                asm volatile("bar.sync 1, %0;" ::"n"(NWG * 128) : "memory");
#pragma unroll
                for (int g2 = 0; g2 < NWG; ++g2) m2 = fmaxf(m2, statm_s[g2 * QT + r]);
                if (!(m2 > -INFINITY)) m2 = 0.f;            // rows past the sequence end (never stored)
                // ---------------- pass 2: p~ = 2^(s2 - m2) (UN-normalised), dropout, smem A tile; l = sum p~
                float l = 0.f;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(s_full(sb), (sf_ph >> sb) & 1u);
                    sf_ph ^= 1u << sb;
                    tc_fence_after();
                    // this warpgroup's 32 keys live in 64-key chunk (wg >> 1), 16-byte units (wg & 1) * 4 .. + 3
                    const uint32_t p_chunk = sP + pb * (QT * KB * 2) + (wg >> 1) * (QT * 128) + r * 128;
                    {
                        const int j0 = kb * KB + wg * 32;
                        const int o = o_row + j0;
                        const float4* b4 = reinterpret_cast<const float4*>(bias_s + (o & 3) * cs + (o & ~3));
                        const float4* m4 = reinterpret_cast<const float4*>(mask_s + j0);
                        const bool full = (j0 + 32 <= len) && !P.key_mask;       // warp-uniform
                        const uint32_t pair0 = (uint32_t)((uint64_t)(grow + j0) >> 1);
                        mbar_wait(p_empty(pb), ((pe_ph >> pb) & 1u) ^ 1u);   // the MMA that read this P buffer has retired
                        pe_ph ^= 1u << pb;
                        auto slice_p = [&](auto full_c) {
                            constexpr bool FULL = decltype(full_c)::value;
                            float ls = 0.f;
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {         // 16 columns at a time keeps the register count down
                                uint32_t v[16];
                                tmem_ld16(tS(sb) + lane_off + wg * 32 + hf * 16, v);
                                tmem_ld_wait();
#pragma unroll
                                for (int q2 = 0; q2 < 2; ++q2) {     // 8 keys per step: one 16-byte store each way
                                    const int q = 2 * hf + q2;
                                    float pr[8];
#pragma unroll
                                    for (int u = 0; u < 2; ++u) {
                                        const float4 bb = b4[2 * q + u];
                                        float x0 = fmaf(__uint_as_float(v[8 * q2 + 4 * u]), LOG2E, bb.x);
                                        float x1 = fmaf(__uint_as_float(v[8 * q2 + 4 * u + 1]), LOG2E, bb.y);
                                        float x2 = fmaf(__uint_as_float(v[8 * q2 + 4 * u + 2]), LOG2E, bb.z);
                                        float x3 = fmaf(__uint_as_float(v[8 * q2 + 4 * u + 3]), LOG2E, bb.w);
                                        if constexpr (!FULL) {
                                            const float4 mm = m4[2 * q + u];
                                            x0 += mm.x; x1 += mm.y; x2 += mm.z; x3 += mm.w;
                                        }
                                        pr[4 * u] = ex2_approx(x0 - m2); pr[4 * u + 1] = ex2_approx(x1 - m2);
                                        pr[4 * u + 2] = ex2_approx(x2 - m2); pr[4 * u + 3] = ex2_approx(x3 - m2);
                                    }
                                    ls += ((pr[0] + pr[1]) + (pr[2] + pr[3])) + ((pr[4] + pr[5]) + (pr[6] + pr[7]));
                                    uint32_t pd[4];
                                    if constexpr (DROP) {
                                        // the 1 / (1 - p) factor is applied to O once per row (inv_l below), not per probability
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            const uint32_t hsh = drop_hash_k(dkey, pair0 + 4 * q + u);
                                            pd[u] = pack_bf16x2((hsh << 16) >= thr_hi ? pr[2 * u] : 0.f, hsh >= thr_hi ? pr[2 * u + 1] : 0.f);
                                        }
                                    }
                                    if constexpr (SAVEP || !DROP) {
                                        uint32_t pk[4];
#pragma unroll
                                        for (int u = 0; u < 4; ++u) pk[u] = pack_bf16x2(pr[2 * u], pr[2 * u + 1]);
                                        if constexpr (!DROP) {
#pragma unroll
                                            for (int u = 0; u < 4; ++u) pd[u] = pk[u];
                                        }
                                        // global P_save (un-normalised, un-dropped; the backward multiplies by row_scale)
                                        if constexpr (SAVEP) {
                                            if (row_ok) {
                                                const int j = j0 + 8 * q;
                                                if (j + 8 <= Lk) {
                                                    *reinterpret_cast<uint4*>(P.P_save + grow + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                                } else {
#pragma unroll
                                                    for (int t = 0; t < 4; ++t)
                                                        if (j + 2 * t < Lk) *reinterpret_cast<uint32_t*>(P.P_save + grow + j + 2 * t) = pk[t];
                                                }
                                            }
                                        }
                                    }
                                    // smem A tile, K-major SWIZZLE_128B: [128 rows][128 B] per 64-key chunk, 16-byte units XOR (row & 7)
                                    const uint32_t addr = p_chunk + (uint32_t)((((wg & 1) * 4 + q) ^ (r & 7)) << 4);
                                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pd[0]), "r"(pd[1]), "r"(pd[2]),
                                                 "r"(pd[3]) : "memory");
                                }
                            }
                            return ls;
                        };
                        l += full ? slice_p(std::true_type{}) : slice_p(std::false_type{});
                    }
                    // S buffer free; P tile complete: make the generic-proxy smem writes visible to the tensor core
                    tc_fence_before();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(s_empty(sb)); mbar_arrive(p_full(pb)); }
                    sb = (sb + 1) % NSB;
                    pb = (pb + 1 == P.np_buf) ? 0 : pb + 1;
                }
                // row sum over the four column slices -> 1 / l normalises O here and P_save in the backward
                statl_s[wg * QT + r] = l;
                asm volatile("bar.sync 1, %0;" ::"n"(NWG * 128) : "memory");
                l = 0.f;
#pragma unroll
                for (int g2 = 0; g2 < NWG; ++g2) l += statl_s[g2 * QT + r];
                const float inv_l = l > 0.f ? 1.f / l : 0.f;
                if (wg == 0 && row_ok && P.row_scale) P.row_scale[((int64_t)b * P.H + h) * Lq + i] = inv_l;
                if (wg == 1 && row_ok && P.row_lse2) P.row_lse2[((int64_t)b * P.H + h) * Lq + i] = m2 + __log2f(l);
                const float o_scale = DROP ? inv_l * P.drop.inv_keep : inv_l;     // dropout's 1 / (1 - p), once per row
                // ---------------- O -> ctx (this warpgroup writes 16 of the 64 head columns)
                mbar_wait(o_full, of_ph);
                of_ph ^= 1;
                tc_fence_after();
                {
                    uint32_t o[16];
                    tmem_ld16(tO + lane_off + wg * 16, o);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(o_empty);
                    if (row_ok) {
                        bf16* dst = P.ctx + (ctx_row0 + i) * P.ld_ctx + h * 64 + wg * 16;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            uint4 w;
                            w.x = pack_bf16x2(__uint_as_float(o[8 * q]) * o_scale, __uint_as_float(o[8 * q + 1]) * o_scale);
                            w.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * o_scale, __uint_as_float(o[8 * q + 3]) * o_scale);
                            w.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * o_scale, __uint_as_float(o[8 * q + 5]) * o_scale);
                            w.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * o_scale, __uint_as_float(o[8 * q + 7]) * o_scale);
                            *reinterpret_cast<uint4*>(dst + 8 * q) = w;
                        }
                    }
                }
            }
            // bias/mask tables of this pair no longer needed
            __syncwarp();
            if (lane == 0) mbar_arrive(bm_empty(buf));
            buf = (buf + 1 == (int)P.ntbl) ? 0 : buf + 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// qkv: [B*L, 3A] bf16 (q | k | v column blocks, heads 64 wide).  Returns false if the shape is not supported.
bool fattn_fwd(const void* qkv, int64_t ld_qkv, int A, int B, int H, int L, const float* bias_rel, const int* key_mask,
               void* P_save, float* row_scale, float* row_lse2, void* ctx, int64_t ld_ctx, DropCfg drop, cudaStream_t st,
               const int* offs, const int* lens, int64_t packed_rows) {
    if (L > 512 || L % 8 != 0 || ld_qkv % 8 != 0) return false;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        P5_CUDA(cudaGetDevice(&dev));
        P5_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    FaParams P;
    P.B = B; P.H = H; P.Lq = L; P.Lk = L;
    P.nkb = (int)cdiv(L, KB); P.nqt = (int)cdiv(L, QT);
    const uint32_t kv_bytes = (uint32_t)P.nkb * KB * 128;
    P.nq_buf = (kv_bytes <= 48 * 1024) ? 2 : 1;
    P.sK = 0; P.sV = kv_bytes; P.sQ = 2 * kv_bytes; P.sP = P.sQ + P.nq_buf * QT * 128;
    P.np_buf = (kv_bytes <= 48 * 1024) ? 2 : 1;       // L = 512: K + V take 128 KB, one P tile is all that fits
    P.sBias = P.sP + P.np_buf * QT * KB * 2;
    P.bias_cs = (uint32_t)(((L + P.nkb * KB + 4 + 31) & ~31) + 8);   // copies land in different bank groups
    P.tbl_stride = 4 * P.bias_cs;
    {   // two table buffers when shared memory allows (L <= 384)
        const size_t one = (size_t)P.sBias + P.tbl_stride * 4 + (size_t)P.nkb * KB * 4 + NWG * QT * 8 + 256 + 1024;
        P.ntbl = (one + P.tbl_stride * 4 + (size_t)P.nkb * KB * 4 <= 232448) ? 2 : 1;
    }
    P.sMask = P.sBias + P.ntbl * P.tbl_stride * 4;
    P.sStat = P.sMask + P.ntbl * (uint32_t)(P.nkb * KB * 4);
    P.sBar = P.sStat + NWG * QT * 8;
    const size_t smem = P.sBar + 256 + 1024;
    P5_CHECK(smem <= 232448, "fattn_fwd: shared memory budget exceeded");
    P.bias_rel = bias_rel; P.key_mask = offs ? nullptr : key_mask; P.P_save = (bf16*)P_save; P.row_scale = row_scale; P.row_lse2 = row_lse2; P.ctx = (bf16*)ctx; P.ld_ctx = ld_ctx;
    P.offs = offs; P.lens = lens;
    P.drop = drop;
    // compile-time variants: dropout on / off, probabilities saved (materialised backward) or not
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const FaParams);
    static const KernelFn variants[4] = {fattn_fwd_kernel<false, false>, fattn_fwd_kernel<false, true>,
                                         fattn_fwd_kernel<true, false>, fattn_fwd_kernel<true, true>};
    const int vi = (drop.thr ? 2 : 0) + (P_save ? 1 : 0);
    static size_t max_set[4] = {0, 0, 0, 0};
    if (smem > max_set[vi]) {
        P5_CUDA(cudaFuncSetAttribute(variants[vi], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set[vi] = smem;
    }
    // 4-D views (c, pos, h, b): row stride ld_qkv, head stride 64 elements, batch stride L*ld_qkv
    // padded: (c, pos, h, b) with batch stride L*ld.  packed: one [packed_rows, .] matrix, batch dim 1 (TMA zero-fills
    // rows past packed_rows; rows of the NEXT sequence that fall into a box are masked by `lens`)
    const uint64_t dims[4] = {64, (uint64_t)(offs ? packed_rows : L), (uint64_t)H, (uint64_t)(offs ? 1 : B)};
    const uint64_t strides[3] = {(uint64_t)ld_qkv * 2, 128, (uint64_t)(offs ? packed_rows : L) * ld_qkv * 2};
    const uint32_t box[4] = {64, 128, 1, 1};
    const bf16* base = (const bf16*)qkv;
    CUtensorMap tmQ = tmap_bf16_4d(base, dims, strides, box);
    CUtensorMap tmK = tmap_bf16_4d(base + A, dims, strides, box);
    CUtensorMap tmV = tmap_bf16_4d(base + 2 * A, dims, strides, box);
    const int n_pairs = B * H;
    const int budget = sm_budget() < num_sms ? sm_budget() : num_sms;     // leaves SMs to a concurrent NCCL all-reduce (common.cuh)
    const int grid = n_pairs < budget ? n_pairs : budget;
    launch_k(variants[vi], grid, FA_THREADS, smem, st, tmQ, tmK, tmV, P);
    P5_CUDA(cudaGetLastError());
    ++g_launches;
    return true;
}

}  // namespace p5
