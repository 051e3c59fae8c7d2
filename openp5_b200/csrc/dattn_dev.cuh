// dattn_dev.cuh — device-side building blocks of the mma.sync decoder attention (dattn.cu): fragment loads with the
// permuted contraction index, swizzled K/V staging, the score / P.V tile loops, and the inference forward of one
// (user, head) pair, shared by the stand-alone kernels (dattn.cu) and the persistent decode kernel (decode_persist.cu).
#pragma once
#include "kernels.cuh"
#include "tc_ptx.cuh"
#include <float.h>

namespace p5 {
namespace {


#ifdef P5_CA_STAMPS      // debugging aid of decode_persist.cu: where one (user, head) pair of the decode cross-attention spends its time
__device__ unsigned long long g_ca_stamp[8];
#define CA_STAMP(i) do { if (blockIdx.x == 0 && tid == 0 && bar_id == 2) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_ca_stamp[i] = t_; } } while (0)
#else
#define CA_STAMP(i)
#endif

struct DAttnDev {
    int B, H, Lq, Lk;
    const bf16 *q, *k, *v;
    int64_t q_ld, q_bs, k_ld, k_bs, v_ld, v_bs;
    const float* bias_rel; int bias_off, n_delta;
    const int* key_mask;
    int causal;
    const int* kv_off; const int* kv_len;
    DropCfg drop;
};

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(saddr), "l"(g));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// 16 consecutive bf16 (32 bytes) of one row as 8 packed pairs; zeros when !ok
__device__ __forceinline__ void ld_row16(uint32_t (&r)[8], const bf16* p, bool ok) {
    uint4 x = make_uint4(0, 0, 0, 0), y = make_uint4(0, 0, 0, 0);
    if (ok) {
        x = *reinterpret_cast<const uint4*>(p);
        y = *reinterpret_cast<const uint4*>(p + 8);
    }
    r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w; r[4] = y.x; r[5] = y.y; r[6] = y.z; r[7] = y.w;
}

// swizzled [rows][64] bf16 tile: byte offset of 16-byte chunk c of row r
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// stage rows [key0, key0 + nrows) of a [*, 64]-column slice into a swizzled tile at local rows [lrow0, ...); rows >= Lk are zero
__device__ __forceinline__ void stage_rows(uint8_t* tile, int lrow0, const bf16* src, int64_t ld, int key0, int nrows, int Lk,
                                           int lane) {
    for (int e = lane; e < nrows * 8; e += 32) {
        const int r = e >> 3, c = e & 7;
        uint8_t* dst = tile + swz(lrow0 + r, c);
        if (key0 + r < Lk) cp_async16(smem_u32(dst), src + (int64_t)(key0 + r) * ld + c * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
}

struct ScoreCtx {
    int Lq, Lk, LkMask, causal, bias_off, n_delta;
    const float* bias;      // bias_rel + h * n_delta or null
    const int* mask;        // key_mask + b * a.Lk or null
};
__device__ __forceinline__ bool score_valid(const ScoreCtx& c, int r, int j) {
    return r < c.Lq && j < c.Lk && (!c.causal || j <= r) && (!c.mask || c.mask[j] != 0);
}
__device__ __forceinline__ int bias_index(const ScoreCtx& c, int r, int j) {
    int di = j - r + c.bias_off;
    return di < 0 ? 0 : (di >= c.n_delta ? c.n_delta - 1 : di);
}

// S (or dPd) tiles of this warp: acc[nt] += A(rows of the query block, fragments lo/hi) . B(rows key0 + 8 nt + g)^T
// The B fragments of up to four tiles are loaded BEFORE the first mma (predicated loads, no branch): the kernel is a
// chain of global-load round trips, and with one load -> mma dependency per tile the loads were not in flight together.
template <int NT, bool LQ16>
__device__ __forceinline__ void qk_tiles(float (&acc)[NT][4], const uint32_t (&alo)[8], const uint32_t (&ahi)[8], const bf16* kbase,
                                         int64_t ld, int key0, int ntw, int Lk, int g, int t) {
    constexpr int G = NT < 4 ? NT : 4;
#pragma unroll
    for (int n0 = 0; n0 < NT; n0 += G) {
        uint32_t kr[G][8];
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const int j = key0 + 8 * (n0 + u) + g;
            ld_row16(kr[u], kbase + (int64_t)j * ld + 16 * t, (n0 + u) < ntw && j < Lk);   // zeros outside the warp's keys
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
            acc[n0 + u][0] = acc[n0 + u][1] = acc[n0 + u][2] = acc[n0 + u][3] = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                mma16816(acc[n0 + u], alo[2 * s], LQ16 ? ahi[2 * s] : 0u, alo[2 * s + 1], LQ16 ? ahi[2 * s + 1] : 0u, kr[u][2 * s],
                         kr[u][2 * s + 1]);
        }
    }
}

// acc[8 d-tiles] += A(P or dS in accumulator layout, this warp's keys) . B(tile rows = keys, via ldmatrix.trans)
template <int NT>
__device__ __forceinline__ void pv_tiles(float (&o)[8][4], const float (&p)[NT][4], const uint8_t* tile, int lrow0, int ntw, int lane) {
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
        if (2 * kk < ntw) {
            const uint32_t a0 = pack_bf16(p[2 * kk][0], p[2 * kk][1]), a1 = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
            const uint32_t a2 = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]), a3 = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
            const int row = lrow0 + 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(b0, b1, b2, b3, smem_u32(tile + swz(row, 2 * c2 + (lane >> 4))));
                mma16816(o[2 * c2], a0, a1, a2, a3, b0, b1);
                mma16816(o[2 * c2 + 1], a0, a1, a2, a3, b2, b3);
            }
        }
    }
}

template <int NT, int NW> struct DCfg {
    static constexpr int KW = NT * 8;                          // max keys per warp
    static constexpr int TILE = NW * KW * 128;                 // staged K or V rows (bytes)
    static constexpr int ORED = NW * 16 * 64 * 4;              // cross-warp partial outputs (fp32)
    static constexpr int PSTRIDE = (KW + 8) * 2;               // bytes per query row of the per-warp Pd / dS tiles
    static constexpr int FWD_SMEM = (TILE > ORED ? TILE : ORED) + NW * 16 * 2 * 4;
    static constexpr int BWD_SMEM = TILE + ORED + 2 * 16 * 128 + NW * 2 * 16 * PSTRIDE + NW * 16 * 4;
};

// One (batch b, head h) pair on NW warps: up to 32 query rows (the beams of a user at a decode step) against that user's
// keys.  `qrows` (optional): the i-th query / output row lives at absolute row qrows[i] of q / O (row stride q_ld / ld_o,
// no batch stride) — the persistent decode kernel passes the user's live beam rows; null = rows b*bs + i*ld.
// `tid` / `bar_id`: the NW warps may be a SUB-GROUP of the CTA (two (user, head) pairs side by side in the persistent decode
// kernel): tid = thread index inside the group, bar_id = the named barrier the group synchronises on (0 = the whole CTA).
// KSMEM: the K rows go through shared memory too (cp.async, every row of the pair in flight at once; the tile follows the
// V region and the softmax scratch) instead of eight dependent rounds of register loads — for K | V that come from HBM.
template <int NT, int NW, bool KSMEM = false>
__device__ __forceinline__ void dattn_fwd32_body(const DAttnDev& a, bf16* __restrict__ O, int64_t ld_o, int64_t bs_o, int b, int h,
                                                 uint8_t* smem, const int* qrows = nullptr, int tid = -1, int bar_id = 0) {
    if (tid < 0) tid = threadIdx.x;
    auto group_sync = [&]() {
        if (bar_id == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(NW * 32) : "memory");
    };
    CA_STAMP(0);
    using C = DCfg<NT, NW>;
    constexpr int PART = NW * 32 * 64 * 4;
    constexpr int REGION = C::TILE > PART ? C::TILE : PART;
    uint8_t* Vs = smem;
    float* red = reinterpret_cast<float*>(smem + REGION);            // [2][NW][32]
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int Lk = a.kv_len ? a.kv_len[b] : a.Lk;
    const int64_t k_boff = a.kv_off ? (int64_t)a.kv_off[b] * a.k_ld : (int64_t)b * a.k_bs;
    const int64_t v_boff = a.kv_off ? (int64_t)a.kv_off[b] * a.v_ld : (int64_t)b * a.v_bs;
    const int kpw = min(C::KW, (((Lk + NW - 1) / NW) + 15) & ~15);
    const int key0 = warp * kpw, ntw = kpw >> 3;
    const bf16* kb = a.k + k_boff + h * 64;
    const bf16* vb = a.v + v_boff + h * 64;
    uint8_t* Ks = smem + REGION + 2 * NW * 32 * 4 + 128;
    if (KSMEM) {
        stage_rows(Ks, warp * C::KW, kb, a.k_ld, key0, kpw, Lk, lane);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    stage_rows(Vs, warp * C::KW, vb, a.v_ld, key0, kpw, Lk, lane);
    asm volatile("cp.async.commit_group;\n" ::: "memory");

    // validity of this lane's 2 NT key columns (length + padding mask), read once and early: the score loop below touches
    // every column 4 times and used to go back to the mask in global memory for each of them
    uint32_t kbits = 0;
    {
        const int* km = a.key_mask ? a.key_mask + (int64_t)b * a.Lk : nullptr;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = key0 + 8 * nt + 2 * t + e;
                if (nt < ntw && j < Lk && (!km || km[j] != 0)) kbits |= 1u << (2 * nt + e);
            }
    }
    uint32_t q[2][2][8];
    const bf16* qb = a.q + (qrows ? 0 : (int64_t)b * a.q_bs) + h * 64 + 16 * t;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
            const int r = 16 * mt + 8 * hi + g;
            ld_row16(q[mt][hi], qb + (int64_t)((qrows && r < a.Lq) ? qrows[r] : r) * a.q_ld, r < a.Lq);
        }
    CA_STAMP(1);
    if (KSMEM) {
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
    }
    CA_STAMP(2);
    float s[2][NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
        if (nt < ntw) {
            const int j = key0 + 8 * nt + g;
            uint32_t kr[8];
            if (KSMEM) {
                const int lr = warp * C::KW + 8 * nt + g;
                const uint4 x = *reinterpret_cast<const uint4*>(Ks + swz(lr, 2 * t)), y = *reinterpret_cast<const uint4*>(Ks + swz(lr, 2 * t + 1));
                kr[0] = x.x; kr[1] = x.y; kr[2] = x.z; kr[3] = x.w; kr[4] = y.x; kr[5] = y.y; kr[6] = y.z; kr[7] = y.w;
            } else {
                ld_row16(kr, kb + (int64_t)j * a.k_ld + 16 * t, j < Lk);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    mma16816(s[mt][nt], q[mt][0][2 * k4], q[mt][1][2 * k4], q[mt][0][2 * k4 + 1], q[mt][1][2 * k4 + 1], kr[2 * k4],
                             kr[2 * k4 + 1]);
        }
    }
    CA_STAMP(3);
    ScoreCtx sc{a.Lq, Lk, a.Lk, a.causal, a.bias_off, a.n_delta, a.bias_rel ? a.bias_rel + h * a.n_delta : nullptr,
                a.key_mask ? a.key_mask + (int64_t)b * a.Lk : nullptr};
    float mx[2][2], sum[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        mx[mt][0] = mx[mt][1] = -FLT_MAX;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = 16 * mt + g + 8 * (e >> 1), j = key0 + 8 * nt + 2 * t + (e & 1);
                float v = -FLT_MAX;
                if (((kbits >> (2 * nt + (e & 1))) & 1u) && r < sc.Lq && (!sc.causal || j <= r))
                    v = s[mt][nt][e] + (sc.bias ? sc.bias[bias_index(sc, r, j)] : 0.f);
                s[mt][nt][e] = v;
                mx[mt][e >> 1] = fmaxf(mx[mt][e >> 1], v);
            }
#pragma unroll
        for (int q2 = 0; q2 < 2; ++q2) {
            mx[mt][q2] = fmaxf(mx[mt][q2], __shfl_xor_sync(0xffffffffu, mx[mt][q2], 1));
            mx[mt][q2] = fmaxf(mx[mt][q2], __shfl_xor_sync(0xffffffffu, mx[mt][q2], 2));
        }
    }
    if (t == 0)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { red[warp * 32 + 16 * mt + g] = mx[mt][0]; red[warp * 32 + 16 * mt + g + 8] = mx[mt][1]; }
    group_sync();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            mx[mt][0] = fmaxf(mx[mt][0], red[w * 32 + 16 * mt + g]);
            mx[mt][1] = fmaxf(mx[mt][1], red[w * 32 + 16 * mt + g + 8]);
        }
    float* red2 = red + NW * 32;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        sum[mt][0] = sum[mt][1] = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = (s[mt][nt][e] == -FLT_MAX) ? 0.f : __expf(s[mt][nt][e] - mx[mt][e >> 1]);
                s[mt][nt][e] = p;
                sum[mt][e >> 1] += p;
            }
#pragma unroll
        for (int q2 = 0; q2 < 2; ++q2) {
            sum[mt][q2] += __shfl_xor_sync(0xffffffffu, sum[mt][q2], 1);
            sum[mt][q2] += __shfl_xor_sync(0xffffffffu, sum[mt][q2], 2);
        }
        if (t == 0) { red2[warp * 32 + 16 * mt + g] = sum[mt][0]; red2[warp * 32 + 16 * mt + g + 8] = sum[mt][1]; }
    }
    group_sync();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        float tot[2] = {0.f, 0.f};
#pragma unroll
        for (int w = 0; w < NW; ++w) { tot[0] += red2[w * 32 + 16 * mt + g]; tot[1] += red2[w * 32 + 16 * mt + g + 8]; }
        const float inv0 = tot[0] > 0.f ? 1.f / tot[0] : 0.f, inv1 = tot[1] > 0.f ? 1.f / tot[1] : 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { s[mt][nt][0] *= inv0; s[mt][nt][1] *= inv0; s[mt][nt][2] *= inv1; s[mt][nt][3] *= inv1; }
    }
    CA_STAMP(4);
    cp_async_wait_all();
    __syncwarp();
    CA_STAMP(5);
    float o[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int c = 0; c < 8; ++c) o[mt][c][0] = o[mt][c][1] = o[mt][c][2] = o[mt][c][3] = 0.f;
        pv_tiles<NT>(o[mt], s[mt], Vs, warp * C::KW, ntw, lane);
    }
    CA_STAMP(6);
    group_sync();                       // every warp is done with its V rows: reuse the region for the partial sums
    float* part = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<float2*>(part + (warp * 32 + 16 * mt + g) * 64 + 8 * c + 2 * t) = make_float2(o[mt][c][0], o[mt][c][1]);
            *reinterpret_cast<float2*>(part + (warp * 32 + 16 * mt + g + 8) * 64 + 8 * c + 2 * t) = make_float2(o[mt][c][2], o[mt][c][3]);
        }
    group_sync();
    for (int e = tid; e < 32 * 8; e += NW * 32) {      // (row, 8-column chunk)
        const int r = e >> 3, c8 = (e & 7) * 8;
        if (r >= a.Lq) continue;
        float v[8];
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2) v[q2] = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int q2 = 0; q2 < 8; ++q2) v[q2] += part[(w * 32 + r) * 64 + c8 + q2];
        uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        *reinterpret_cast<uint4*>(O + (qrows ? (int64_t)qrows[r] * ld_o : (int64_t)b * bs_o + (int64_t)r * ld_o) + h * 64 + c8) = pk;
    }
    CA_STAMP(7);
}


}  // namespace
}  // namespace p5
