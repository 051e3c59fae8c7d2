// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[b][M,N] = epilogue( A[b][M,K] * B[b][N,K]^T )
//
// This single kernel carries every dense contraction of the T5 block
// (HF:models/t5/modeling_t5.py:277,298-299,338 q/k/v/o; :93-102 wi/wo; P5_T5.py:361 lm_head) and their
// dgrad / wgrad forms.  bf16 operands, fp32 accumulation in tensor memory.
//
// Structure (persistent, warp specialised, one CTA per SM):
//   warp 0   : TMA producer      cp.async.bulk.tensor.4d -> 128B-swizzled smem ring (STAGES deep)
//   warp 1   : MMA issuer        one lane issues tcgen05.mma.cta_group::1.kind::f16, 128 x BLOCK_N x 16
//   warp 2   : TMEM allocator    2 accumulator stages x BLOCK_N fp32 columns
//   warps 4-11: epilogue         two warps per TMEM lane quarter, each owning half of the tile's columns (one warp per
//                                scheduler cannot hide its own ALU / LDS latencies; the K = 768 shapes were epilogue-bound):
//                                tcgen05.ld 32x32b -> registers -> per-warp swizzled smem transpose -> fused epilogue (one
//                                compile-time variant per flag set, operands prefetched 16 row pairs ahead) -> row-segment stores
// Two mbarrier pipelines: smem full/empty (TMA<->MMA), tmem full/empty (MMA<->epilogue).  Tile widths 64 / 128 / 192 / 256;
// an opt-in CTA-pair variant (cta_group::2, cluster of two, 256 x 256 tile per pair); batch operands may be broadcast;
// programmatic dependent launch with the wait at the start or (for GEMMs independent of their predecessor) at the end.
//
// Operands may be K-major (reduction index contiguous) or MN-major (row index contiguous); the latter lets
// dgrad/wgrad read the forward tensors in place (no transposes): TMA loads [64k x 64mn] boxes and the smem
// descriptor uses the MN-major SWIZZLE_128B canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) (uint128 units).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <unordered_map>
#include <vector>
#include <mutex>
#include <algorithm>
#include <string>
#include <string.h>
#include <stdlib.h>

namespace p5 {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row
static constexpr int UMMA_K = 16;
static constexpr int EPI_WARPS = 8;          // two warps per TMEM lane quarter: each takes half of the tile's columns
static constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;
static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
struct TcParams {
    int M, N, K;
    int nb1, nb2;
    int a_major, b_major;
    int dbg;   // perf-debug only (P5_GEMM_DBG): 1 = no global stores, 2 = no smem staging either, 4 = no TMEM loads
    int late_wait;   // GemmProblem::indep_of_prev: griddepcontrol.wait at the end instead of the start
    int a_m1, a_m2, b_m1, b_m2;   // batch-coordinate multipliers (0 = the operand is broadcast over that batch dimension)
    GemmEpilogue epi;
};

template <int BLOCK_N, bool CTA2 = false>
struct TcCfg {
    // CTA2: a CTA pair shares one 256 x BLOCK_N tile; each CTA stages its 128 rows of A and HALF of B
    static constexpr int B_ROWS = CTA2 ? BLOCK_N / 2 : BLOCK_N;
    static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = CTA2 ? 6 : ((BLOCK_N >= 192) ? 4 : (BLOCK_N == 128 ? 6 : 8));
    static constexpr int TMEM_COLS = 2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);   // power of two >= 2 accumulators
    static constexpr int EPI_STRIDE = 32;                       // floats per staged row: 32 columns, column pairs XOR-swizzled by the row
    static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_STRIDE * 4;   // one 32x32 fp32 tile per epilogue warp
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES;
};

// 8 consecutive output columns of one row: fused epilogue + (vectorised) store
__device__ __forceinline__ void epi_store8(const GemmEpilogue& e, const float* acc, int64_t idx, int ncols_valid) {
    const bool vec = (ncols_valid == 8) && ((idx & 7) == 0);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[j] * e.alpha;
    if (e.flags & EPI_RELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (e.flags & EPI_MULPOS) {
        if (vec && e.aux_dtype == DT_BF16) {
            uint4 a = *reinterpret_cast<const uint4*>((const bf16*)e.aux + idx);
            const bf16* ab = reinterpret_cast<const bf16*>(&a);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __bfloat162float(ab[j]) > 0.f ? v[j] : 0.f;
        } else {
            for (int j = 0; j < ncols_valid; ++j) v[j] = ld_as_f32(e.aux, e.aux_dtype, idx + j) > 0.f ? v[j] : 0.f;
        }
    }
    if (e.flags & EPI_DROPOUT) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            v[j] = drop_keep(e.seed, e.site, (uint64_t)(idx + j), e.drop_thr) ? v[j] * e.inv_keep : 0.f;
    }
    if (e.flags & EPI_ADD_RESID) {
        if (vec) {
            float4 r0 = *reinterpret_cast<const float4*>(e.resid + idx);
            float4 r1 = *reinterpret_cast<const float4*>(e.resid + idx + 4);
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        } else {
            for (int j = 0; j < ncols_valid; ++j) v[j] += e.resid[idx + j];
        }
    }
    if (e.flags & EPI_ACCUM) {
        const float* c = (const float*)e.C;
        if (vec) {
            float4 r0 = *reinterpret_cast<const float4*>(c + idx);
            float4 r1 = *reinterpret_cast<const float4*>(c + idx + 4);
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        } else {
            for (int j = 0; j < ncols_valid; ++j) v[j] += c[idx + j];
        }
    }
    if (e.flags & EPI_ATOMIC) {
        float* c = (float*)e.C;
        for (int j = 0; j < ncols_valid; ++j) atomicAdd(c + idx + j, v[j]);
    } else if (e.c_dtype == DT_F32) {
        float* c = (float*)e.C;
        if (vec) {
            *reinterpret_cast<float4*>(c + idx) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(c + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
            for (int j = 0; j < ncols_valid; ++j) c[idx + j] = v[j];
        }
    } else {
        bf16* c = (bf16*)e.C;
        if (vec) {
            uint4 o;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            *reinterpret_cast<uint4*>(c + idx) = o;
        } else {
            for (int j = 0; j < ncols_valid; ++j) c[idx + j] = __float2bfloat16_rn(v[j]);
        }
    }
}

// 2 consecutive output columns of one row (the staged epilogue: lanes own adjacent column pairs, so every warp-wide
// access below is one contiguous 128-byte (bf16) or 256-byte (fp32) row segment)
// EPI >= 0: the flag set (bits 0-5) and the output dtype (bit 6 = fp32) are compile-time constants, so each
// specialisation is straight-line code (the runtime-flag form made the epilogue instruction-fetch bound);
// EPI = -1 keeps the generic runtime form for uncommon combinations.
static constexpr int EPI_OUT_F32 = 64;
template <int EPI>
__device__ __forceinline__ void epi_store2(const GemmEpilogue& e, float a0, float a1, int64_t idx, bool two) {
    const int flags = EPI >= 0 ? (EPI & 63) : e.flags;
    const bool out_f32 = EPI >= 0 ? ((EPI & EPI_OUT_F32) != 0) : (e.c_dtype == DT_F32);
    const bool vec = two && ((idx & 1) == 0);
    float v0 = a0 * e.alpha, v1 = a1 * e.alpha;
    if (flags & EPI_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
    if (flags & EPI_MULPOS) {
        if (vec && e.aux_dtype == DT_BF16) {
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>((const bf16*)e.aux + idx);
            const float2 f = __bfloat1622float2(a);
            v0 = f.x > 0.f ? v0 : 0.f; v1 = f.y > 0.f ? v1 : 0.f;
        } else {
            v0 = ld_as_f32(e.aux, e.aux_dtype, idx) > 0.f ? v0 : 0.f;
            if (two) v1 = ld_as_f32(e.aux, e.aux_dtype, idx + 1) > 0.f ? v1 : 0.f;
        }
    }
    if (flags & EPI_DROPOUT) {
        v0 = drop_keep(e.seed, e.site, (uint64_t)idx, e.drop_thr) ? v0 * e.inv_keep : 0.f;
        v1 = drop_keep(e.seed, e.site, (uint64_t)(idx + 1), e.drop_thr) ? v1 * e.inv_keep : 0.f;
    }
    if (flags & EPI_ADD_RESID) {
        if (vec) { const float2 r = *reinterpret_cast<const float2*>(e.resid + idx); v0 += r.x; v1 += r.y; }
        else { v0 += e.resid[idx]; if (two) v1 += e.resid[idx + 1]; }
    }
    if (flags & EPI_ACCUM) {
        const float* c = (const float*)e.C;
        if (vec) { const float2 r = *reinterpret_cast<const float2*>(c + idx); v0 += r.x; v1 += r.y; }
        else { v0 += c[idx]; if (two) v1 += c[idx + 1]; }
    }
    if (flags & EPI_ATOMIC) {
        float* c = (float*)e.C;
        if (vec) {   // one 8-byte vector reduction instead of two scalar atomics (sm_90+)
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(c + idx), "f"(v0), "f"(v1) : "memory");
        } else {
            atomicAdd(c + idx, v0);
            if (two) atomicAdd(c + idx + 1, v1);
        }
    } else if (out_f32) {
        float* c = (float*)e.C;
        if (vec) *reinterpret_cast<float2*>(c + idx) = make_float2(v0, v1);
        else { c[idx] = v0; if (two) c[idx + 1] = v1; }
    } else {
        bf16* c = (bf16*)e.C;
        if (vec) *reinterpret_cast<__nv_bfloat162*>(c + idx) = __floats2bfloat162_rn(v0, v1);
        else { c[idx] = __float2bfloat16_rn(v0); if (two) c[idx + 1] = __float2bfloat16_rn(v1); }
    }
}

template <int BLOCK_N, int EPI, bool CTA2 = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ TcParams P) {
    using Cfg = TcCfg<BLOCK_N, CTA2>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_holder = bar_base + 8u * (2 * STAGES + 4);
    float* epi_stage = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
    volatile uint32_t* tmem_holder_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_holder - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // CTA2: the scheduling unit is the CTA pair (cluster of 2); it owns 256 rows of which this CTA takes 128
    const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const long long unit = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;
    const long long unit_stride = CTA2 ? (gridDim.x >> 1) : gridDim.x;
    constexpr int UNIT_M = CTA2 ? 2 * BLOCK_M : BLOCK_M;
    const int m_tiles = (P.M + UNIT_M - 1) / UNIT_M;
    const int n_tiles = (P.N + BLOCK_N - 1) / BLOCK_N;
    const int k_blocks = (P.K + BLOCK_K - 1) / BLOCK_K;
    const long long total_tiles = (long long)m_tiles * n_tiles * P.nb1 * P.nb2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), (CTA2 ? 2 : 1) * EPI_WARPS);  // one arrive per epilogue warp (of both CTAs of a pair)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CTA2) tmem_alloc_2cta(tmem_holder, Cfg::TMEM_COLS);
        else tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder_ptr;
    // PDL: barrier init, TMEM allocation and tensor-map prefetch above overlap the tail of the previous kernel; nothing
    // below touches global memory before the previous grid has completed
    if (!P.late_wait) pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = unit; t < total_tiles; t += unit_stride) {
                const int n_blk = (int)(t % n_tiles);
                const long long t2 = t / n_tiles;
                const int m_blk = (int)(t2 % m_tiles) * (CTA2 ? 2 : 1) + (int)cta_rank;   // 128-row block of this CTA
                const int b = (int)(t2 / m_tiles);
                const int b1 = b % P.nb1, b2 = b / P.nb1;
                const int n0 = n_blk * BLOCK_N + (int)cta_rank * Cfg::B_ROWS;            // first B row staged by this CTA
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + A_STAGE_BYTES;
                    // pair mode: only the leader arms its barrier, with the bytes of BOTH CTAs' loads
                    if (!CTA2 || leader) mbar_expect_tx(full_bar(stage), (CTA2 ? 2 : 1) * Cfg::STAGE_BYTES);
                    auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1) {
                        const bool is_a = tm == &tmA;
                        const int c2 = b1 * (is_a ? P.a_m1 : P.b_m1), c3 = b2 * (is_a ? P.a_m2 : P.b_m2);
                        if constexpr (CTA2) tma_load_4d_2cta(dst, tm, full_bar(stage), c0, c1, c2, c3);
                        else tma_load_4d(dst, tm, full_bar(stage), c0, c1, c2, c3);
                    };
                    if (P.a_major == MAJOR_K) {
                        load(sa, &tmA, kb * BLOCK_K, m_blk * BLOCK_M);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_M / 64; ++c) load(sa + c * 8192, &tmA, m_blk * BLOCK_M + c * 64, kb * BLOCK_K);
                    }
                    if (P.b_major == MAJOR_K) {
                        load(sb, &tmB, kb * BLOCK_K, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < Cfg::B_ROWS / 64; ++c) load(sb + c * 8192, &tmB, n0 + c * 64, kb * BLOCK_K);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && leader) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
            // a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)   (pair mode: M = 256 across the two CTAs)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.a_major & 1) << 15) |
                                   ((uint32_t)(P.b_major & 1) << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                   ((uint32_t)(UNIT_M >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long t = unit; t < total_tiles; t += unit_stride) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // K-major : advance 16 elements = 32 bytes inside the 128B swizzle row; SBO = 8 rows = 1024 B
                        // MN-major: advance 16 k-rows   = 2048 bytes; LBO = next 64-wide MN chunk (8192 B); SBO = 1024 B
                        const uint64_t da = (P.a_major == MAJOR_K) ? make_smem_desc(sa + k * 32, 16, 1024)
                                                                   : make_smem_desc(sa + k * 2048, 8192, 1024);
                        const uint64_t db = (P.b_major == MAJOR_K) ? make_smem_desc(sb + k * 32, 16, 1024)
                                                                   : make_smem_desc(sb + k * 2048, 8192, 1024);
                        if constexpr (CTA2) umma_bf16_2cta(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // frees the smem slot (in both CTAs of a pair) once these MMAs retire
                    if constexpr (CTA2) umma_commit_2cta(empty_bar(stage)); else umma_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue (of both CTAs)
                if constexpr (CTA2) umma_commit_2cta(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp & 3;            // TMEM lane quarter owned by this warp
        const int half = (warp - 4) >> 2;   // which half of the tile's 32-column chunks
        constexpr int NCH = BLOCK_N / 64;   // 32-column chunks per warp
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long t = unit; t < total_tiles; t += unit_stride) {
            const int n_blk = (int)(t % n_tiles);
            const long long t2 = t / n_tiles;
            const int m_blk = (int)(t2 % m_tiles) * (CTA2 ? 2 : 1) + (int)cta_rank;
            const int b = (int)(t2 / m_tiles);
            const int b1 = b % P.nb1, b2 = b / P.nb1;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int row = m_blk * BLOCK_M + ew * 32 + lane;
            const bool row_ok = row < P.M;
            const int64_t row_off = (int64_t)b1 * P.epi.cs1 + (int64_t)b2 * P.epi.cs2 + (int64_t)row * P.epi.ldc;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            auto release_acc = [&]() {      // all TMEM reads of this warp are done: hand the accumulator back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (CTA2 && !leader) mbar_arrive_cluster(tempty_bar(acc), 0); else mbar_arrive(tempty_bar(acc)); }
            };
            if constexpr (EPI != -2) {
                // TMEM -> registers (thread = row) -> per-warp smem tile -> registers (half-warp = 16 column pairs of a row,
                // two rows per step) -> global: every access of the fused epilogue (aux / residual / C) is a row segment
                float* tile = epi_stage + (warp - 4) * (32 * Cfg::EPI_STRIDE);
                const int64_t boff = (int64_t)b1 * P.epi.cs1 + (int64_t)b2 * P.epi.cs2;
                const int row0 = m_blk * BLOCK_M + ew * 32;
                const int cp = lane & 15, rpar = lane >> 4;
                // inputs of the fused epilogue (ReLU-mask operand, fp32 residual, fp32 accumulate target) are PREFETCHED
                // into registers for all 32 rows of a chunk before the TMEM load: independent loads in flight instead of
                // one dependent load per row
                constexpr int F = EPI >= 0 ? (EPI & 63) : 0;
                constexpr bool PF_AUX = (F & EPI_MULPOS) != 0, PF_RES = (F & EPI_ADD_RESID) != 0, PF_ACC = (F & EPI_ACCUM) != 0;
                constexpr bool PF = PF_AUX || PF_RES || PF_ACC;
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    const int cc = half * NCH + c;
                    const int col = n_blk * BLOCK_N + cc * 32 + 2 * cp;
                    const bool col_ok = col < P.N, two = col + 1 < P.N;
                    const int nrows = min(32, P.M - row0);
                    const int64_t idx0 = boff + (int64_t)(row0 + rpar) * P.epi.ldc + col;    // this lane's first row
                    // fast path (every specialised epilogue): whole column pair in range and 4/8-byte aligned -> hoisted
                    // address arithmetic, straight-line fully unrolled row loop
                    // (full 32-row tiles only, so that the unrolled row loop is branch-free and the rows interleave)
                    const bool fast = (EPI >= 0) && col_ok && two && nrows == 32 && (((idx0 | P.epi.ldc) & 1) == 0) &&
                                      (!PF_AUX || P.epi.aux_dtype == DT_BF16);
                    uint32_t pa[PF_AUX ? 16 : 1];
                    float2 pr[PF_RES ? 16 : 1], pc[PF_ACC ? 16 : 1];
                    if (PF && fast) {
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            const int64_t idx = idx0 + (int64_t)(2 * rr) * P.epi.ldc;
                            if constexpr (PF_AUX) pa[rr] = *reinterpret_cast<const uint32_t*>((const bf16*)P.epi.aux + idx);
                            if constexpr (PF_RES) pr[rr] = *reinterpret_cast<const float2*>(P.epi.resid + idx);
                            if constexpr (PF_ACC) pc[rr] = *reinterpret_cast<const float2*>((const float*)P.epi.C + idx);
                        }
                    }
                    uint32_t r[32];
                    if (!(P.dbg & 4)) {
                        tmem_ld32(taddr + cc * 32, r);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = j;
                    }
                    if (c == NCH - 1) release_acc();
                    if (P.dbg & 2) {
                        uint32_t x = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) x ^= r[j];
                        if (x == 0x12345u) P.epi.alpha == 0.f ? (void)0 : (void)atomicAdd((int*)P.epi.C, 1);
                        continue;
                    }
                    // staged row `lane`: column pair j lives at float2 slot j ^ (lane & 15) (conflict-free both ways)
                    float* myrow = tile + lane * Cfg::EPI_STRIDE;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        *reinterpret_cast<float2*>(myrow + 2 * (j ^ cp)) = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                    __syncwarp();
                    if (col_ok) {
                        if (P.dbg & 1) {
                            float acc2 = 0.f;
                            for (int rr = 0; rr < 16; ++rr) {
                                const int R = 2 * rr + rpar;
                                const float2 v = *reinterpret_cast<const float2*>(tile + R * Cfg::EPI_STRIDE + 2 * (cp ^ (R & 15)));
                                acc2 += v.x + v.y;
                            }
                            if (acc2 == 1.2345e30f) ((float*)P.epi.C)[0] = acc2;
                        } else if (fast) {
                            const GemmEpilogue& e = P.epi;
                            const float alpha = e.alpha;
                            const int64_t ldc = e.ldc;
                            char* cptr = (char*)e.C + idx0 * ((EPI & EPI_OUT_F32) ? 4 : 2);
                            const int64_t cstep = 2 * ldc * ((EPI & EPI_OUT_F32) ? 4 : 2);
                            const uint64_t seed = e.seed;
                            const uint32_t site = e.site, thr = e.drop_thr;
                            const float inv_keep = e.inv_keep;
#pragma unroll
                            for (int rr = 0; rr < 16; ++rr) {
                                const int R = 2 * rr + rpar;
                                const float2 v = *reinterpret_cast<const float2*>(tile + R * Cfg::EPI_STRIDE + 2 * (cp ^ (R & 15)));
                                float v0 = v.x * alpha, v1 = v.y * alpha;
                                if constexpr ((F & EPI_RELU) != 0) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                                if constexpr (PF_AUX) {
                                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pa[rr]));
                                    v0 = f.x > 0.f ? v0 : 0.f; v1 = f.y > 0.f ? v1 : 0.f;
                                }
                                if constexpr ((F & EPI_DROPOUT) != 0) {
                                    bool k0, k1;   // idx is even on the fast path: one hash for the pair
                                    drop_keep2(seed, site, (uint64_t)(idx0 + (int64_t)(2 * rr) * ldc), thr, k0, k1);
                                    v0 = k0 ? v0 * inv_keep : 0.f;
                                    v1 = k1 ? v1 * inv_keep : 0.f;
                                }
                                if constexpr (PF_RES) { v0 += pr[rr].x; v1 += pr[rr].y; }
                                if constexpr (PF_ACC) { v0 += pc[rr].x; v1 += pc[rr].y; }
                                char* dst = cptr + rr * cstep;
                                if constexpr ((F & EPI_ATOMIC) != 0)
                                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(v0), "f"(v1) : "memory");
                                else if constexpr ((EPI & EPI_OUT_F32) != 0)
                                    *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                                else
                                    *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(v0, v1);
                            }
                        } else {
#pragma unroll 4
                            for (int rr = 0; rr < 16; ++rr) {
                                const int R = 2 * rr + rpar;
                                if (R >= nrows) continue;
                                const float2 v = *reinterpret_cast<const float2*>(tile + R * Cfg::EPI_STRIDE + 2 * (cp ^ (R & 15)));
                                epi_store2<EPI>(P.epi, v.x, v.y, boff + (int64_t)(row0 + R) * P.epi.ldc + col, two);
                            }
                        }
                    }
                    __syncwarp();
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                const int cc = half * NCH + c;
                uint32_t r[32];
                tmem_ld32(taddr + cc * 32, r);
                tmem_ld_wait();
                if (c == NCH - 1) release_acc();
                const int col0 = n_blk * BLOCK_N + cc * 32;
                if (row_ok && col0 < P.N) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int col = col0 + g * 8;
                        int nv = P.N - col;
                        if (nv > 0) {
                            if (nv > 8) nv = 8;
                            float a[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) a[j] = __uint_as_float(r[g * 8 + j]);
                            epi_store8(P.epi, a, row_off + col, nv);
                        }
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all();     // neither CTA frees TMEM / exits while its peer still reads or signals it
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CTA2) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
    // independent-of-predecessor GEMM: the dependency wait happens here, so that "this grid completed" still implies
    // "everything before it on the stream completed" for the kernels that follow
    if (P.late_wait) pdl_wait();
}

// ------------------------------------------------------------------------------------------
// host side: tensor-map cache + launcher
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        P5_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        P5_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

struct TmapKey {
    const void* ptr;
    uint64_t dims[4];
    uint64_t strides[3];
    uint32_t box[4];
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
        size_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
        return h;
    }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mu;
static int g_tc_launches = 0;

void gemm_tc_clear_cache() {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    g_tmap_cache.clear();
}
int gemm_tc_launch_count() { return g_tc_launches; }

// cached 4-D bf16 SWIZZLE_128B tensor map (dims innermost first, strides in bytes for dims 1..3)
CUtensorMap tmap_bf16_4d(const void* ptr, const uint64_t dims[4], const uint64_t strides[3], const uint32_t box[4]) {
    TmapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = ptr;
    for (int i = 0; i < 4; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
    for (int i = 0; i < 3; ++i) key.strides[i] = strides[i];
    {
        std::lock_guard<std::mutex> g(g_tmap_mu);
        auto it = g_tmap_cache.find(key);
        if (it != g_tmap_cache.end()) return it->second;
    }
    CUtensorMap m;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), key.dims,
                                 key.strides, key.box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[512];
        snprintf(b, sizeof(b),
                 "cuTensorMapEncodeTiled failed (%d): ptr=%p dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] "
                 "box=[%u,%u,%u,%u]",
                 (int)r, ptr, (unsigned long long)key.dims[0], (unsigned long long)key.dims[1],
                 (unsigned long long)key.dims[2], (unsigned long long)key.dims[3], (unsigned long long)key.strides[0],
                 (unsigned long long)key.strides[1], (unsigned long long)key.strides[2], key.box[0], key.box[1],
                 key.box[2], key.box[3]);
        throw P5Error(3, b);
    }
    std::lock_guard<std::mutex> g(g_tmap_mu);
    // bounded: pad-to-longest batches change Le (hence the activation shapes) from step to step, so an epoch can create
    // thousands of distinct keys; an encode costs ~1 us, so the cache is simply restarted when it is full
    if (g_tmap_cache.size() >= 4096) g_tmap_cache.clear();
    g_tmap_cache[key] = m;
    return m;
}

static CUtensorMap make_tmap(const GemmOperand& op, int rows, int K, int nb1, int nb2, int box_rows) {
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    const uint64_t ld_b = (uint64_t)op.ld * 2;
    if (op.major == MAJOR_K) {
        dims[0] = (uint64_t)K; dims[1] = (uint64_t)rows;
        box[0] = BLOCK_K; box[1] = (uint32_t)box_rows;
    } else {
        dims[0] = (uint64_t)rows; dims[1] = (uint64_t)K;
        box[0] = 64; box[1] = BLOCK_K;
    }
    if (op.bcast1) nb1 = 1;     // broadcast operand: the kernel always passes coordinate 0 for that dimension
    if (op.bcast2) nb2 = 1;
    dims[2] = (uint64_t)nb1; dims[3] = (uint64_t)nb2;
    box[2] = 1; box[3] = 1;
    strides[0] = ld_b;
    strides[1] = nb1 > 1 ? (uint64_t)op.bs1 * 2 : dims[1] * ld_b;
    strides[2] = nb2 > 1 ? (uint64_t)op.bs2 * 2 : (nb1 > 1 ? (uint64_t)op.bs1 * 2 * nb1 : dims[1] * ld_b);
    return tmap_bf16_4d(op.ptr, dims, strides, box);
}

static bool operand_ok(const GemmOperand& o, int nb1, int nb2, bool allow_mn) {
    if (o.dtype != DT_BF16) return false;
    if (o.major == MAJOR_MN && !allow_mn) return false;
    if (((uintptr_t)o.ptr & 15) != 0) return false;
    if (o.ld % 8 != 0 || o.ld <= 0) return false;
    if (nb1 > 1 && !o.bcast1 && (o.bs1 % 8 != 0 || o.bs1 <= 0)) return false;
    if (nb2 > 1 && !o.bcast2 && (o.bs2 % 8 != 0 || o.bs2 <= 0)) return false;
    return true;
}

bool gemm_tc_supported(const GemmProblem& p, bool allow_mn_major) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0) return false;
    if (!operand_ok(p.A, p.nb1, p.nb2, allow_mn_major)) return false;
    if (!operand_ok(p.B, p.nb1, p.nb2, allow_mn_major)) return false;
    // TMA zero-fills out-of-bounds boxes, so M, N, K need no padding or alignment (only ld / base address do)
    if ((p.epi.flags & (EPI_ACCUM | EPI_ATOMIC)) && p.epi.c_dtype != DT_F32) return false;
    return true;
}

static int g_num_sms = 0;
static int g_sm_reserved = 0;
static int g_force_block_n = 0;  // test hook

void sm_reserve(int n_sms) { g_sm_reserved = n_sms < 0 ? 0 : n_sms; g_num_sms = 0; }
int sm_budget() {
    static int total = 0;
    if (!total) {
        int dev = 0;
        P5_CUDA(cudaGetDevice(&dev));
        P5_CUDA(cudaDeviceGetAttribute(&total, cudaDevAttrMultiProcessorCount, dev));
    }
    const int n = total - g_sm_reserved;
    return n < 8 ? 8 : n;
}

// ---- optional per-launch timing (bench.py roofline leg): CUDA events on the launching stream around every
//      tcgen05 GEMM launch, with the algorithmic FLOPs of the problem
struct ProfRec { cudaEvent_t e0, e1; double flops; int bn; int M, N, K, nb, flags, am, bm; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
void gemm_tc_prof_enable(bool on) {
    g_prof_on = on;
    if (on) {
        for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        g_prof.clear();
    }
}
// returns JSON: per BLOCK_N class {launches, ms, flops}; synchronises the device
std::string gemm_tc_prof_summary() {
    cudaDeviceSynchronize();
    double ms[3] = {0, 0, 0}, fl[3] = {0, 0, 0};
    long n[3] = {0, 0, 0};
    for (auto& r : g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) continue;
        const int c = r.bn >= 192 ? 0 : (r.bn == 128 ? 1 : 2);
        ms[c] += t; fl[c] += r.flops; n[c] += 1;
    }
    char b[512];
    snprintf(b, sizeof(b),
             "{\"bn256\": {\"launches\": %ld, \"ms\": %.6f, \"flops\": %.6e}, \"bn128\": {\"launches\": %ld, \"ms\": %.6f, "
             "\"flops\": %.6e}, \"bn64\": {\"launches\": %ld, \"ms\": %.6f, \"flops\": %.6e}}",
             n[0], ms[0], fl[0], n[1], ms[1], fl[1], n[2], ms[2], fl[2]);
    return std::string(b);
}
// per-shape table of the profiled launches: "M N K batches flags a_major b_major bn : launches total_ms TFLOP/s" lines
std::string gemm_tc_prof_shapes() {
    cudaDeviceSynchronize();
    struct Acc { long n = 0; double ms = 0, fl = 0; };
    std::unordered_map<std::string, Acc> tab;
    for (auto& r : g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) continue;
        char k[128];
        snprintf(k, sizeof(k), "%6d %6d %6d nb%-3d epi%-3d maj%d%d bn%d", r.M, r.N, r.K, r.nb, r.flags, r.am, r.bm, r.bn);
        Acc& a = tab[k];
        a.n += 1; a.ms += t; a.fl += r.flops;
    }
    std::vector<std::pair<double, std::string>> rows;
    for (auto& kv : tab) {
        char b[256];
        snprintf(b, sizeof(b), "%s : %4ld x %8.2f us  %7.1f TFLOP/s  total %8.3f ms\n", kv.first.c_str(), kv.second.n, 1e3 * kv.second.ms / kv.second.n,
                 kv.second.ms > 0 ? kv.second.fl / (kv.second.ms * 1e-3) / 1e12 : 0.0, kv.second.ms);
        rows.push_back({-kv.second.ms, b});
    }
    std::sort(rows.begin(), rows.end());
    std::string out;
    for (auto& r : rows) out += r.second;
    return out;
}
void gemm_tc_force_block_n(int bn) { g_force_block_n = bn; }

static int g_epi_mode = -1;   // -1: read P5_GEMM_EPI; 0 = direct register->global generic epilogue, 1 = staged generic, 2 = staged specialised
void gemm_tc_set_epilogue(int mode) { g_epi_mode = mode; }

template <int BN, int EPI>
static void launch_tc(const GemmProblem& p, cudaStream_t stream) {
    using Cfg = TcCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        P5_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    CUtensorMap tmA = make_tmap(p.A, p.M, p.K, p.nb1, p.nb2, BLOCK_M);
    CUtensorMap tmB = make_tmap(p.B, p.N, p.K, p.nb1, p.nb2, BN);
    TcParams P;
    P.M = p.M; P.N = p.N; P.K = p.K; P.nb1 = p.nb1; P.nb2 = p.nb2;
    P.a_major = p.A.major; P.b_major = p.B.major;
    P.epi = p.epi;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("P5_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
    P.dbg = dbg;
    P.late_wait = (p.indep_of_prev && pdl_enabled()) ? 1 : 0;
    P.a_m1 = p.A.bcast1 ? 0 : 1; P.a_m2 = p.A.bcast2 ? 0 : 1; P.b_m1 = p.B.bcast1 ? 0 : 1; P.b_m2 = p.B.bcast2 ? 0 : 1;
    const long long tiles = (long long)cdiv(p.M, BLOCK_M) * cdiv(p.N, BN) * p.nb1 * p.nb2;
    const int grid = (int)(tiles < g_num_sms ? tiles : g_num_sms);
    ProfRec rec;
    if (g_prof_on) {
        cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1);
        rec.flops = 2.0 * p.M * p.N * (double)p.K * p.nb1 * p.nb2; rec.bn = BN;
        rec.M = p.M; rec.N = p.N; rec.K = p.K; rec.nb = p.nb1 * p.nb2; rec.flags = p.epi.flags | (p.epi.c_dtype == DT_F32 ? 64 : 0);
        rec.am = p.A.major; rec.bm = p.B.major;
        cudaEventRecord(rec.e0, stream);
    }
    launch_k(gemm_tc_kernel<BN, EPI>, grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream, tmA, tmB, P);
    P5_CUDA(cudaGetLastError());
    if (g_prof_on) { cudaEventRecord(rec.e1, stream); g_prof.push_back(rec); }
    ++g_tc_launches;
}

// CTA-pair (cta_group::2) launch of the 256-wide tile: clusters of two CTAs, each pair owns a 256 x 256 tile.  Per
// k-block a CTA pulls 16 KB of A + 16 KB of B from L2 instead of 16 + 32 KB: the single-CTA kernel moves ~94 B/clk/SM
// at full tensor rate, which is what caps it below the cuBLAS rate.
template <int EPI>
static void launch_tc_pair(const GemmProblem& p, cudaStream_t stream) {
    constexpr int BN = 256;
    using Cfg = TcCfg<BN, true>;
    static bool attr_set = false;
    if (!attr_set) {
        P5_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    CUtensorMap tmA = make_tmap(p.A, p.M, p.K, p.nb1, p.nb2, BLOCK_M);
    CUtensorMap tmB = make_tmap(p.B, p.N, p.K, p.nb1, p.nb2, Cfg::B_ROWS);
    TcParams P;
    P.M = p.M; P.N = p.N; P.K = p.K; P.nb1 = p.nb1; P.nb2 = p.nb2;
    P.a_major = p.A.major; P.b_major = p.B.major;
    P.epi = p.epi;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("P5_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
    P.dbg = dbg;
    P.late_wait = (p.indep_of_prev && pdl_enabled()) ? 1 : 0;
    P.a_m1 = p.A.bcast1 ? 0 : 1; P.a_m2 = p.A.bcast2 ? 0 : 1; P.b_m1 = p.B.bcast1 ? 0 : 1; P.b_m2 = p.B.bcast2 ? 0 : 1;
    const long long units = (long long)cdiv(p.M, 2 * BLOCK_M) * cdiv(p.N, BN) * p.nb1 * p.nb2;
    const int pairs = (int)(units < g_num_sms / 2 ? units : g_num_sms / 2);
    ProfRec rec;
    if (g_prof_on) {
        cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1);
        rec.flops = 2.0 * p.M * p.N * (double)p.K * p.nb1 * p.nb2; rec.bn = BN;
        cudaEventRecord(rec.e0, stream);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    P5_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, true>, tmA, tmB, P));
    P5_CUDA(cudaGetLastError());
    if (g_prof_on) { cudaEventRecord(rec.e1, stream); g_prof.push_back(rec); }
    ++g_tc_launches;
}

// Pair mode is OPT-IN (P5_GEMM_PAIR=1): measured on B200 it is parity-green but not faster (T5-base step 16.91 vs
// 16.86 ms; 128x256x16 MMAs retire every ~245 cycles in both modes), i.e. the kernel is bound by operand delivery
// from shared memory into the tensor core (SS mode: 12 KB per MMA), not by L2 -> SM traffic, which pair mode cuts
// by a third.  The next step for the GEMM is therefore A-from-TMEM (TS mode), not wider operand sharing.
static bool use_pair(const GemmProblem& p) {
    static const int mode = [] { const char* e = getenv("P5_GEMM_PAIR"); return e ? atoi(e) : 0; }();
    if (!mode) return false;
    const long long units = (long long)cdiv(p.M, 2 * BLOCK_M) * cdiv(p.N, 256) * p.nb1 * p.nb2;
    return units >= g_num_sms / 2;
}

template <int EPI>
static void launch_bn(int bn, const GemmProblem& p, cudaStream_t stream) {
    if (bn == 256 && use_pair(p)) launch_tc_pair<EPI>(p, stream);
    else if (bn == 256) launch_tc<256, EPI>(p, stream);
    else if (bn == 192) launch_tc<192, EPI>(p, stream);
    else if (bn == 128) launch_tc<128, EPI>(p, stream);
    else launch_tc<64, EPI>(p, stream);
}

// tile width (64 / 128 / 192 / 256) of an M x N output (x batches) on `sms` persistent CTAs: the candidate with the lowest
// rounds(bn) x (128 + bn + 24); a narrower tile has to be 3 % better to displace a wider one.  Pure host arithmetic
// (p5_gemm_tile_width exposes it to the CPU tests).
int gemm_tc_tile_width(int M, int N, int batches, int sms) {
    const long long mt = cdiv(M, BLOCK_M) * (long long)batches;
    int bn = 0;
    double best = 0;
    for (int cand : {256, 192, 128, 64}) {
        if (cand > 64 && N <= cand / 2) continue;          // more than half of the tile would be padding
        const double c = (double)cdiv(mt * cdiv(N, cand), sms) * (128 + cand + 24);
        if (!bn || c < best * 0.97) { bn = cand; best = c; }
    }
    return bn;
}

void gemm_tc(const GemmProblem& p, cudaStream_t stream) {
    P5_CHECK(gemm_tc_supported(p, true), "gemm_tc: unsupported problem");
    if (!g_num_sms) g_num_sms = sm_budget();
    int bn = g_force_block_n ? g_force_block_n : p.prefer_bn;
    // One cost model over all four tile widths: time ~ rounds(bn) x (bytes a CTA pulls per k-block ~ 128 + bn, plus a fixed
    // part); ties go to the wider tile.  It differs from the older case analysis below (P5_TILE_MODEL=0) for M ~ 4-5 k rows
    // (the eval encoder: N = 768 in ONE round of 192-wide tiles instead of two rounds of 128-wide ones) and for the
    // decoder's M = 512 x N = 3072 (one round of 128-wide tiles).  Measured on one B200: eval batch 7.54 -> 7.41 ms, train
    // step 15.95 -> 15.92 ms.
    static const int tile_model = [] { const char* e = getenv("P5_TILE_MODEL"); return e ? atoi(e) : 1; }();
    if (!bn && tile_model && !p.tail_filled) bn = gemm_tc_tile_width(p.M, p.N, p.nb1 * p.nb2, g_num_sms);
    if (!bn) {
        const long long mt = cdiv(p.M, BLOCK_M) * (long long)p.nb1 * p.nb2;
        // largest tile that still gives every SM a tile; narrow outputs use a narrow tile
        if (p.N > 128 && mt * cdiv(p.N, 256) >= g_num_sms) {
            // persistent waves: time ~ rounds x (bytes a CTA pulls per k-block ~ 128 + bn, plus a fixed part).  A 192-wide
            // tile wins when it removes a mostly empty last wave (N = 768 with ~100 row tiles: 300 tiles = 2.03 waves
            // of 256-wide tiles, but 400 tiles = 2.7 waves of 192-wide ones)
            const double c256 = (double)cdiv(mt * cdiv(p.N, 256), g_num_sms) * (128 + 256 + 24);
            const double c192 = (double)cdiv(mt * cdiv(p.N, 192), g_num_sms) * (128 + 192 + 24);
            // (the backward's dgrads pass prefer_bn = 256: their partial last wave is filled by the wgrad that follows)
            static const bool use192 = getenv("P5_NO_BN192") == nullptr;
            bn = (c192 < 0.95 * c256 && use192 && !p.tail_filled) ? 192 : 256;
        }
        else if (p.N > 64 && mt * cdiv(p.N, 128) >= g_num_sms) bn = 128;
        else bn = 64;
        if (p.N <= 64) bn = 64;
    }
    if (g_epi_mode < 0) {
        const char* e = getenv("P5_GEMM_EPI");
        g_epi_mode = (e && strcmp(e, "direct") == 0) ? 0 : ((e && strcmp(e, "generic") == 0) ? 1 : 2);
    }
    if (g_epi_mode == 0) { launch_bn<-2>(bn, p, stream); return; }
    if (g_epi_mode == 1) { launch_bn<-1>(bn, p, stream); return; }
    // compile-time specialisations of the flag sets the engine uses; anything else takes the generic kernel
    const int key = (p.epi.flags & 63) | (p.epi.c_dtype == DT_F32 ? EPI_OUT_F32 : 0);
    switch (key) {
        case 0: launch_bn<0>(bn, p, stream); break;
        case EPI_OUT_F32: launch_bn<EPI_OUT_F32>(bn, p, stream); break;
        case EPI_RELU: launch_bn<EPI_RELU>(bn, p, stream); break;
        case EPI_RELU | EPI_DROPOUT: launch_bn<(EPI_RELU | EPI_DROPOUT)>(bn, p, stream); break;
        case EPI_ADD_RESID | EPI_OUT_F32: launch_bn<(EPI_ADD_RESID | EPI_OUT_F32)>(bn, p, stream); break;
        case EPI_ADD_RESID | EPI_DROPOUT | EPI_OUT_F32: launch_bn<(EPI_ADD_RESID | EPI_DROPOUT | EPI_OUT_F32)>(bn, p, stream); break;
        case EPI_MULPOS: launch_bn<EPI_MULPOS>(bn, p, stream); break;
        case EPI_ACCUM | EPI_OUT_F32: launch_bn<(EPI_ACCUM | EPI_OUT_F32)>(bn, p, stream); break;
        case EPI_ATOMIC | EPI_OUT_F32: launch_bn<(EPI_ATOMIC | EPI_OUT_F32)>(bn, p, stream); break;
        default: launch_bn<-1>(bn, p, stream); break;
    }
}

}  // namespace p5
