// Attention kernels (HF:models/t5/modeling_t5.py:253-344): scores are UNSCALED q.k, plus the shared relative
// position bias and the additive finfo(fp32).min padding / causal mask, softmax in fp32, dropout on the
// probabilities, then P.V.
//
//  * attn_simt_fwd / attn_simt_bwd : fused fp32 kernels (no S materialised).  They are the exact-arithmetic path
//    of the parity mode and serve the decoder's short query blocks (Lq = Ld ~ 8, decode Lq = 1) in bf16 mode.
//  * softmax_fwd / softmax_bwd     : row kernels between the batched tcgen05 GEMMs (S = QK^T, O = PV, dP = dO V^T,
//    dQ/dK/dV) of the bf16 encoder path.
#include "kernels.cuh"
#include <float.h>

namespace p5 {
extern int g_launches;
#define LAUNCHED() do { P5_CUDA(cudaGetLastError()); ++g_launches; } while (0)

// query rows per CTA = 8 warps x RPW rows per warp (RPW = 1 for short decoder blocks, 2 default, 4 for beam rows)
static constexpr int KT = 64;    // keys per smem tile
static constexpr int DK = 64;    // d_kv
static constexpr int KS = DK + 4;  // smem row stride of K/V tiles: 16-byte aligned rows, conflict-free LDS.128 across keys
#define MASK_MIN (-FLT_MAX)      // torch.finfo(torch.float32).min

struct AttnDev {
    int B, H, Lq, Lk;
    const void *q, *k, *v;
    int q_dt, k_dt, v_dt;
    int64_t q_ld, q_bs, k_ld, k_bs, v_ld, v_bs;
    const float* bias_rel; int bias_off, n_delta;
    const int* key_mask;
    int causal, q_pos_offset;
    const int* row_map;
    const int* kv_off; const int* kv_len;
    DropCfg drop;
};

__device__ __forceinline__ float score_bias(const AttnDev& a, int h, int kb, int i_pos, int j) {
    float s = 0.f;
    if (a.bias_rel) {
        int di = j - i_pos + a.bias_off;
        di = di < 0 ? 0 : (di >= a.n_delta ? a.n_delta - 1 : di);
        s += a.bias_rel[h * a.n_delta + di];
    }
    float m = 0.f;
    if (a.key_mask && a.key_mask[(int64_t)kb * a.Lk + j] == 0) m = MASK_MIN;
    if (a.causal && j > i_pos) m = MASK_MIN;
    return s + m;
}

// 64-wide dot product of two smem rows with 128-bit shared loads (a: broadcast row, b: per-lane row)
__device__ __forceinline__ float dot64(const float* a, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < DK; c += 4) {
        const float4 x = *reinterpret_cast<const float4*>(a + c);
        const float4 y = *reinterpret_cast<const float4*>(b + c);
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    return s;
}

// cooperative load of a [KT x 64] tile (rows = keys) into smem as fp32, zero beyond Lk.  16-byte global loads
// (8 bf16 / 4 fp32 per thread per request); the views are 16-byte aligned (ld, bs, h*64 multiples of 8 elements).
__device__ __forceinline__ void load_kv_tile(float (*dst)[KS], const void* base, int dt, int64_t ld, int64_t boff,
                                             int h, int j0, int Lk) {
    if (dt == DT_BF16) {
        const bf16* p = (const bf16*)base;
        for (int e = threadIdx.x; e < KT * DK / 8; e += blockDim.x) {
            const int j = e >> 3, c = (e & 7) * 8;
            float v[8];
            if (j0 + j < Lk) {
                const uint4 t = *reinterpret_cast<const uint4*>(p + boff + (int64_t)(j0 + j) * ld + h * DK + c);
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
                for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(hh[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[j][c + q] = v[q];
        }
    } else {
        const float* p = (const float*)base;
        for (int e = threadIdx.x; e < KT * DK / 4; e += blockDim.x) {
            const int j = e >> 4, c = (e & 15) * 4;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + j < Lk) t = *reinterpret_cast<const float4*>(p + boff + (int64_t)(j0 + j) * ld + h * DK + c);
            dst[j][c] = t.x; dst[j][c + 1] = t.y; dst[j][c + 2] = t.z; dst[j][c + 3] = t.w;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int RPW>
__global__ void __launch_bounds__(256)
attn_simt_fwd_kernel(AttnDev a, void* O, int o_dt, int64_t ld_o, int64_t bs_o, float* lse) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    constexpr int QB = 8 * RPW;
    extern __shared__ __align__(16) float smem[];
    float (*Qs)[DK] = reinterpret_cast<float (*)[DK]>(smem);                       // [QB][64]
    float (*KVs)[KS] = reinterpret_cast<float (*)[KS]>(smem + QB * DK);    // [KT][65]
    float* Ss = smem + QB * DK + KT * (KS);                                     // [QB][Lk]
    const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * QB;
    const int kb = a.row_map ? a.row_map[b] : b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Lk = a.kv_len ? a.kv_len[kb] : a.Lk;                                   // packed K/V: per-batch key count
    const int64_t k_boff = a.kv_off ? (int64_t)a.kv_off[kb] * a.k_ld : (int64_t)kb * a.k_bs;
    const int64_t v_boff = a.kv_off ? (int64_t)a.kv_off[kb] * a.v_ld : (int64_t)kb * a.v_bs;

    for (int e = threadIdx.x; e < QB * DK; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        float v = 0.f;
        if (i0 + r < a.Lq) v = ld_as_f32(a.q, a.q_dt, (int64_t)b * a.q_bs + (int64_t)(i0 + r) * a.q_ld + h * DK + c);
        Qs[r][c] = v;
    }
    // ---- scores
    for (int j0 = 0; j0 < Lk; j0 += KT) {
        __syncthreads();
        load_kv_tile(KVs, a.k, a.k_dt, a.k_ld, k_boff, h, j0, Lk);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr;
            if (i0 + r >= a.Lq) continue;          // rows past Lq (short decoder blocks) do no work
            const int i_pos = a.q_pos_offset + i0 + r;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int jl = lane + 32 * jj, j = j0 + jl;
                if (j < Lk) {
                    const float s = dot64(Qs[r], KVs[jl]);
                    Ss[r * Lk + j] = s + score_bias(a, h, kb, i_pos, j);
                }
            }
        }
    }
    __syncwarp();
    // ---- softmax (each warp owns its two rows; Ss rows are private to the warp)
    float inv_sum[RPW];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const int r = warp * RPW + rr;
        if (i0 + r >= a.Lq) continue;
        float mx = -INFINITY;
        for (int j = lane; j < Lk; j += 32) mx = fmaxf(mx, Ss[r * Lk + j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < Lk; j += 32) {
            const float p = __expf(Ss[r * Lk + j] - mx);
            Ss[r * Lk + j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        inv_sum[rr] = 1.f / sum;
        const int i = i0 + r;
        if (lse && lane == 0 && i < a.Lq) lse[((int64_t)b * a.H + h) * a.Lq + i] = mx + logf(sum);
        for (int j = lane; j < Lk; j += 32) {
            float p = Ss[r * Lk + j] * inv_sum[rr];
            if (a.drop.thr) {
                const uint64_t idx = (((uint64_t)b * a.H + h) * a.Lq + i) * (uint64_t)Lk + j;
                p = drop_keep(a.drop.seed, a.drop.site, idx, a.drop.thr) ? p * a.drop.inv_keep : 0.f;
            }
            Ss[r * Lk + j] = p;
        }
    }
    // ---- O = P V
    float acc[RPW][2];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) acc[rr][0] = acc[rr][1] = 0.f;
    for (int j0 = 0; j0 < Lk; j0 += KT) {
        __syncthreads();
        load_kv_tile(KVs, a.v, a.v_dt, a.v_ld, v_boff, h, j0, Lk);
        __syncthreads();
        const int jn = min(KT, Lk - j0);
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr;
            if (i0 + r >= a.Lq) continue;
            for (int j = 0; j < jn; ++j) {
                const float p = Ss[r * Lk + j0 + j];
                acc[rr][0] = fmaf(p, KVs[j][lane], acc[rr][0]);
                acc[rr][1] = fmaf(p, KVs[j][lane + 32], acc[rr][1]);
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const int i = i0 + warp * RPW + rr;
        if (i < a.Lq) {
            const int64_t o = (int64_t)b * bs_o + (int64_t)i * ld_o + h * DK;
            st_from_f32(O, o_dt, o + lane, acc[rr][0]);
            st_from_f32(O, o_dt, o + lane + 32, acc[rr][1]);
        }
    }
}

static AttnDev to_dev(const AttnArgs& a) {
    const int bias_off = a.bias_off, n_delta = a.n_delta;
    AttnDev d;
    d.B = a.B; d.H = a.H; d.Lq = a.Lq; d.Lk = a.Lk;
    d.q = a.q.ptr; d.k = a.k.ptr; d.v = a.v.ptr;
    d.q_dt = a.q.dtype; d.k_dt = a.k.dtype; d.v_dt = a.v.dtype;
    d.q_ld = a.q.ld; d.q_bs = a.q.bs; d.k_ld = a.k.ld; d.k_bs = a.k.bs; d.v_ld = a.v.ld; d.v_bs = a.v.bs;
    d.bias_rel = a.bias_rel; d.bias_off = bias_off; d.n_delta = n_delta;
    d.key_mask = a.key_mask; d.causal = a.causal; d.q_pos_offset = a.q_pos_offset; d.row_map = a.row_map;
    d.kv_off = a.kv_off; d.kv_len = a.kv_len;
    d.drop = a.drop;
    return d;
}

static size_t fwd_smem(int QB, int Lk) { return (size_t)(QB * DK + KT * (KS) + QB * Lk) * sizeof(float); }
static size_t bwd_smem(int QB, int Lk, int n_delta) {
    return (size_t)(2 * QB * DK + KT * (KS) + QB * Lk + QB * KT + QB + n_delta) * sizeof(float);
}
static int pick_rpw(int Lq) { return Lq <= 8 ? 1 : ((Lq > 16 && Lq <= 32) ? 4 : 2); }

template <int RPW>
static void launch_attn_fwd(const AttnArgs& a, void* O, int o_dtype, int64_t ld_o, int64_t bs_o, float* lse, cudaStream_t st) {
    constexpr int QB = 8 * RPW;
    static size_t max_set = 0;
    const size_t sm = fwd_smem(QB, a.Lk);
    if (sm > max_set) {
        P5_CUDA(cudaFuncSetAttribute(attn_simt_fwd_kernel<RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        max_set = sm;
    }
    AttnDev d = to_dev(a);
    dim3 grid((unsigned)cdiv(a.Lq, QB), (unsigned)a.H, (unsigned)a.B);
    launch_k(attn_simt_fwd_kernel<RPW>, grid, 256, sm, st, d, O, o_dtype, ld_o, bs_o, lse);
    LAUNCHED();
}

void attn_simt_fwd(const AttnArgs& a, void* O, int o_dtype, int64_t ld_o, int64_t bs_o, float* lse, cudaStream_t st) {
    if (a.B <= 0 || a.Lq <= 0) return;
    P5_CHECK(a.Lk >= 1 && a.Lk <= 1024, "attn_simt_fwd: Lk out of range");
    switch (pick_rpw(a.Lq)) {
        case 1: launch_attn_fwd<1>(a, O, o_dtype, ld_o, bs_o, lse, st); break;
        case 4: launch_attn_fwd<4>(a, O, o_dtype, ld_o, bs_o, lse, st); break;
        default: launch_attn_fwd<2>(a, O, o_dtype, ld_o, bs_o, lse, st); break;
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward (recomputes P from the saved log-sum-exp)
// ------------------------------------------------------------------------------------------------------------
template <int RPW>
__global__ void __launch_bounds__(256)
attn_simt_bwd_kernel(AttnDev a, const void* O, const void* dO, int o_dt, int64_t ld_o, int64_t bs_o,
                     const float* __restrict__ lse, float* dQ, int64_t ld_dq, int64_t bs_dq, float* dK, float* dV,
                     int64_t ld_dkv, int64_t bs_dkv, float* dbias_rel, int atomic_kv) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    constexpr int QB = 8 * RPW;
    extern __shared__ __align__(16) float smem[];
    float (*Qs)[DK] = reinterpret_cast<float (*)[DK]>(smem);
    float (*dOs)[DK] = reinterpret_cast<float (*)[DK]>(smem + QB * DK);
    float (*KVs)[KS] = reinterpret_cast<float (*)[KS]>(smem + 2 * QB * DK);
    float* Ss = smem + 2 * QB * DK + KT * (KS);         // [QB][Lk]  p, then ds
    float* Pt = Ss + QB * a.Lk;                              // [QB][KT]  dropped p of the current tile
    float* dlt = Pt + QB * KT;                               // [QB]
    float* sdb = dlt + QB;                                   // [n_delta]
    const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * QB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Lk = a.kv_len ? a.kv_len[b] : a.Lk;
    const int64_t k_boff = a.kv_off ? (int64_t)a.kv_off[b] * a.k_ld : (int64_t)b * a.k_bs;
    const int64_t v_boff = a.kv_off ? (int64_t)a.kv_off[b] * a.v_ld : (int64_t)b * a.v_bs;
    const int64_t dkv_boff = a.kv_off ? (int64_t)a.kv_off[b] * ld_dkv : (int64_t)b * bs_dkv;
    const int nvalid = min(QB, a.Lq - i0);     // query rows of this block that exist

    for (int e = threadIdx.x; e < QB * DK; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        float qv = 0.f, gv = 0.f;
        if (i0 + r < a.Lq) {
            qv = ld_as_f32(a.q, a.q_dt, (int64_t)b * a.q_bs + (int64_t)(i0 + r) * a.q_ld + h * DK + c);
            gv = ld_as_f32(dO, o_dt, (int64_t)b * bs_o + (int64_t)(i0 + r) * ld_o + h * DK + c);
        }
        Qs[r][c] = qv;
        dOs[r][c] = gv;
    }
    if (dbias_rel)
        for (int e = threadIdx.x; e < a.n_delta; e += blockDim.x) sdb[e] = 0.f;
    __syncthreads();
    // delta_i = sum_c dO_ic * O_ic
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const int r = warp * RPW + rr, i = i0 + r;
        float dsum = 0.f;
        if (i < a.Lq) {
            const int64_t o = (int64_t)b * bs_o + (int64_t)i * ld_o + h * DK;
            dsum = dOs[r][lane] * ld_as_f32(O, o_dt, o + lane) + dOs[r][lane + 32] * ld_as_f32(O, o_dt, o + lane + 32);
        }
        dsum = warp_sum(dsum);
        if (lane == 0) dlt[r] = dsum;
    }
    // ---- recompute p_ij = exp(s_ij - lse_i)
    for (int j0 = 0; j0 < Lk; j0 += KT) {
        __syncthreads();
        load_kv_tile(KVs, a.k, a.k_dt, a.k_ld, k_boff, h, j0, Lk);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr, i = i0 + r;
            const int i_pos = a.q_pos_offset + i;
            const float l = (i < a.Lq) ? lse[((int64_t)b * a.H + h) * a.Lq + i] : 0.f;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int jl = lane + 32 * jj, j = j0 + jl;
                if (j < Lk && i >= a.Lq) Ss[r * Lk + j] = 0.f;
                if (j < Lk && i < a.Lq) {
                    float s = dot64(Qs[r], KVs[jl]);
                    s += score_bias(a, h, b, i_pos, j);
                    Ss[r * Lk + j] = (i < a.Lq) ? __expf(s - l) : 0.f;
                } 
            }
        }
    }
    // ---- dP, dS, dV
    for (int j0 = 0; j0 < Lk; j0 += KT) {
        __syncthreads();
        load_kv_tile(KVs, a.v, a.v_dt, a.v_ld, v_boff, h, j0, Lk);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr, i = i0 + r;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int jl = lane + 32 * jj, j = j0 + jl;
                float pd = 0.f;
                if (j < Lk && i < a.Lq) {
                    const float dpd = dot64(dOs[r], KVs[jl]);
                    const float p = Ss[r * Lk + j];
                    float dp = dpd;
                    pd = p;
                    if (a.drop.thr) {
                        const uint64_t idx = (((uint64_t)b * a.H + h) * a.Lq + i) * (uint64_t)Lk + j;
                        const bool keep = drop_keep(a.drop.seed, a.drop.site, idx, a.drop.thr);
                        dp = keep ? dpd * a.drop.inv_keep : 0.f;
                        pd = keep ? p * a.drop.inv_keep : 0.f;
                    }
                    const float ds = p * (dp - dlt[r]);
                    Ss[r * Lk + j] = ds;
                    if (dbias_rel) {
                        int di = j - (a.q_pos_offset + i) + a.bias_off;
                        di = di < 0 ? 0 : (di >= a.n_delta ? a.n_delta - 1 : di);
                        atomicAdd(&sdb[di], ds);
                    }
                }
                Pt[r * KT + jl] = pd;
            }
        }
        __syncthreads();
        // dV[j, c..c+3] += sum_r pd[r][j] * dO[r][c..c+3]   (valid rows only; one LDS + one LDS.128 per 4 FMAs)
        for (int e = threadIdx.x; e < KT * DK / 4; e += blockDim.x) {
            const int j = e >> 4, c = (e & 15) * 4;
            if (j0 + j < Lk) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < nvalid; ++r) {
                    const float p = Pt[r * KT + j];
                    const float4 g = *reinterpret_cast<const float4*>(&dOs[r][c]);
                    v.x = fmaf(p, g.x, v.x); v.y = fmaf(p, g.y, v.y); v.z = fmaf(p, g.z, v.z); v.w = fmaf(p, g.w, v.w);
                }
                float* dst = dV + dkv_boff + (int64_t)(j0 + j) * ld_dkv + h * DK + c;
                if (!atomic_kv) *reinterpret_cast<float4*>(dst) = v;   // one q-block per (b, h): this CTA owns the element
                else { atomicAdd(dst, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w); }
            }
        }
    }
    // ---- dQ, dK
    float acc[RPW][2];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) acc[rr][0] = acc[rr][1] = 0.f;
    for (int j0 = 0; j0 < Lk; j0 += KT) {
        __syncthreads();
        load_kv_tile(KVs, a.k, a.k_dt, a.k_ld, k_boff, h, j0, Lk);
        __syncthreads();
        const int jn = min(KT, Lk - j0);
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr;
            if (r >= nvalid) continue;
            for (int j = 0; j < jn; ++j) {
                const float ds = Ss[r * Lk + j0 + j];
                acc[rr][0] = fmaf(ds, KVs[j][lane], acc[rr][0]);
                acc[rr][1] = fmaf(ds, KVs[j][lane + 32], acc[rr][1]);
            }
        }
        for (int e = threadIdx.x; e < KT * DK / 4; e += blockDim.x) {
            const int j = e >> 4, c = (e & 15) * 4;
            if (j0 + j < Lk) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < nvalid; ++r) {
                    const float ds = Ss[r * Lk + j0 + j];
                    const float4 q = *reinterpret_cast<const float4*>(&Qs[r][c]);
                    v.x = fmaf(ds, q.x, v.x); v.y = fmaf(ds, q.y, v.y); v.z = fmaf(ds, q.z, v.z); v.w = fmaf(ds, q.w, v.w);
                }
                float* dst = dK + dkv_boff + (int64_t)(j0 + j) * ld_dkv + h * DK + c;
                if (!atomic_kv) *reinterpret_cast<float4*>(dst) = v;
                else { atomicAdd(dst, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w); }
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const int i = i0 + warp * RPW + rr;
        if (i < a.Lq) {
            const int64_t o = (int64_t)b * bs_dq + (int64_t)i * ld_dq + h * DK;
            dQ[o + lane] = acc[rr][0];
            dQ[o + lane + 32] = acc[rr][1];
        }
    }
    if (dbias_rel) {
        __syncthreads();
        for (int e = threadIdx.x; e < a.n_delta; e += blockDim.x) {
            const float v = sdb[e];
            if (v != 0.f) atomicAdd(dbias_rel + h * a.n_delta + e, v);
        }
    }
}

template <int RPW>
static void launch_attn_bwd(const AttnArgs& a, const void* O, const void* dO, int o_dtype, int64_t ld_o, int64_t bs_o,
                            const float* lse, float* dQ, int64_t ld_dq, int64_t bs_dq, float* dK, float* dV, int64_t ld_dkv,
                            int64_t bs_dkv, float* dbias_rel, cudaStream_t st) {
    constexpr int QB = 8 * RPW;
    static size_t max_set = 0;
    const size_t sm = bwd_smem(QB, a.Lk, a.n_delta);
    if (sm > max_set) {
        P5_CUDA(cudaFuncSetAttribute(attn_simt_bwd_kernel<RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        max_set = sm;
    }
    AttnDev d = to_dev(a);
    dim3 grid((unsigned)cdiv(a.Lq, QB), (unsigned)a.H, (unsigned)a.B);
    // with a single q-block per (b, h) every dK/dV element has exactly one producer: plain stores, no memset needed
    launch_k(attn_simt_bwd_kernel<RPW>, grid, 256, sm, st, d, O, dO, o_dtype, ld_o, bs_o, lse, dQ, ld_dq, bs_dq, dK, dV, ld_dkv,
                                                     bs_dkv, dbias_rel, grid.x > 1 ? 1 : 0);
    LAUNCHED();
}

void attn_simt_bwd(const AttnArgs& a, const void* O, const void* dO, int o_dtype, int64_t ld_o, int64_t bs_o,
                   const float* lse, float* dQ, int64_t ld_dq, int64_t bs_dq, float* dK, float* dV, int64_t ld_dkv,
                   int64_t bs_dkv, float* dbias_rel, cudaStream_t st) {
    if (a.B <= 0 || a.Lq <= 0) return;
    P5_CHECK(a.Lk >= 1 && a.Lk <= 1024, "attn_simt_bwd: Lk out of range");
    P5_CHECK(a.row_map == nullptr, "attn_simt_bwd: row_map is inference-only");
    if (pick_rpw(a.Lq) == 1)
        launch_attn_bwd<1>(a, O, dO, o_dtype, ld_o, bs_o, lse, dQ, ld_dq, bs_dq, dK, dV, ld_dkv, bs_dkv, dbias_rel, st);
    else
        launch_attn_bwd<2>(a, O, dO, o_dtype, ld_o, bs_o, lse, dQ, ld_dq, bs_dq, dK, dV, ld_dkv, bs_dkv, dbias_rel, st);
}

// ------------------------------------------------------------------------------------------------------------
// materialised softmax between the batched tensor-core GEMMs.  One CTA per (b, h) [x row slice]; each warp walks
// rows; each lane owns adjacent column PAIRS (8-byte loads of S / dP, 4-byte bf16x2 stores), Lk even and <= 512.
// Dropout uses one counter hash per pair (16 bits per element), regenerated identically in backward.
// ------------------------------------------------------------------------------------------------------------
static constexpr int SM_MAXP = 8;  // pairs per lane: Lk <= 512

__device__ __forceinline__ void drop_pair(const DropCfg& d, uint64_t pair_idx, bool& k0, bool& k1) {
    drop_keep2(d.seed, d.site, pair_idx << 1, d.thr, k0, k1);
}
template <typename T> __device__ __forceinline__ void st_pair(T* p, float a, float b);
template <> __device__ __forceinline__ void st_pair<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void st_pair<bf16>(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
template <typename T> __device__ __forceinline__ float2 ld_pair(const T* p);
template <> __device__ __forceinline__ float2 ld_pair<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 ld_pair<bf16>(const bf16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

template <typename T>
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const float* __restrict__ S, const float* __restrict__ bias_rel, const int* __restrict__ key_mask,
                   T* __restrict__ P_save, T* __restrict__ Pd, int H, int Lq, int Lk, int causal, DropCfg drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    extern __shared__ float sh_mask[];   // [Lk] additive key mask of this batch row (0 or finfo.min)
    const int bh = blockIdx.x, b = bh / H, h = bh % H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n_delta = Lq + Lk - 1;
    for (int j = threadIdx.x; j < Lk; j += blockDim.x) sh_mask[j] = (key_mask && key_mask[b * Lk + j] == 0) ? MASK_MIN : 0.f;
    __syncthreads();
    const float* brow = bias_rel ? bias_rel + h * n_delta : nullptr;
    for (int i = blockIdx.y * nw + warp; i < Lq; i += nw * gridDim.y) {
        const int64_t row = ((int64_t)bh * Lq + i) * Lk;
        float v[SM_MAXP][2];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < SM_MAXP; ++k) {
            const int j = 2 * (lane + 32 * k);
            v[k][0] = v[k][1] = -INFINITY;
            if (j < Lk) {
                const float2 s = *reinterpret_cast<const float2*>(S + row + j);
                const float2 m = *reinterpret_cast<const float2*>(sh_mask + j);
                float a0 = s.x, a1 = s.y;
                if (brow) { a0 += brow[j - i + Lq - 1]; a1 += brow[j + 1 - i + Lq - 1]; }
                float m0 = m.x, m1 = m.y;
                if (causal) { if (j > i) m0 = MASK_MIN; if (j + 1 > i) m1 = MASK_MIN; }
                v[k][0] = a0 + m0; v[k][1] = a1 + m1;
                mx = fmaxf(mx, fmaxf(v[k][0], v[k][1]));
            }
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < SM_MAXP; ++k) {
            const int j = 2 * (lane + 32 * k);
            if (j < Lk) { v[k][0] = __expf(v[k][0] - mx); v[k][1] = __expf(v[k][1] - mx); sum += v[k][0] + v[k][1]; }
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 0; k < SM_MAXP; ++k) {
            const int j = 2 * (lane + 32 * k);
            if (j < Lk) {
                const float p0 = v[k][0] * inv, p1 = v[k][1] * inv;
                st_pair<T>(P_save + row + j, p0, p1);
                if (drop.thr) {
                    bool k0, k1;
                    drop_pair(drop, (uint64_t)(row + j) >> 1, k0, k1);
                    st_pair<T>(Pd + row + j, k0 ? p0 * drop.inv_keep : 0.f, k1 ? p1 * drop.inv_keep : 0.f);
                }
            }
        }
    }
}

void softmax_fwd(const float* S, const float* bias_rel, const int* key_mask, void* P_save, void* Pd, int dtype, int B,
                 int H, int Lq, int Lk, int causal, DropCfg drop, cudaStream_t st) {
    if (B <= 0) return;
    P5_CHECK(Lk <= 64 * SM_MAXP && (Lk % 2) == 0, "softmax_fwd: Lk must be even and <= 512");
    P5_CHECK(!drop.thr || Pd != nullptr, "softmax_fwd: dropout needs a Pd buffer");
    dim3 grid((unsigned)(B * H), (unsigned)(Lq >= 128 ? 2 : 1));
    const size_t sm = (size_t)Lk * sizeof(float);
    if (dtype == DT_F32)
        launch_k(softmax_fwd_kernel<float>, grid, 256, sm, st, S, bias_rel, key_mask, (float*)P_save, (float*)Pd, H, Lq, Lk, causal, drop);
    else
        launch_k(softmax_fwd_kernel<bf16>, grid, 256, sm, st, S, bias_rel, key_mask, (bf16*)P_save, (bf16*)Pd, H, Lq, Lk, causal, drop);
    LAUNCHED();
}

// NP = column pairs per lane (Lk <= 64 * NP); each warp works on TWO rows at a time so that twice as many
// independent loads are in flight (the kernel is latency-, not bandwidth-limited at one row per warp)
template <typename T, int NP>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ dPd, const T* __restrict__ P, T* __restrict__ dS, T* __restrict__ Pd_out,
                   float* __restrict__ dbias_rel, int H, int Lq, int Lk, DropCfg drop, const int* __restrict__ lens,
                   const float* __restrict__ row_scale) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    extern __shared__ float sdb[];  // [n_delta]
    const int bh = blockIdx.x, h = bh % H;
    // packed training: only the first len rows / columns of this (b, h) hold probabilities; everything else is
    // written as exact zeros so that the dV / dQ / dK GEMMs over the padded geometry stay clean
    const int len = lens ? lens[bh / H] : max(Lq, Lk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n_delta = Lq + Lk - 1;
    if (dbias_rel) {
        for (int e = threadIdx.x; e < n_delta; e += blockDim.x) sdb[e] = 0.f;
        __syncthreads();
    }
    for (int i0 = 2 * (blockIdx.y * nw + warp); i0 < Lq; i0 += 2 * nw * gridDim.y) {
        float p[2][NP][2], dp[2][NP][2];
        float dot[2] = {0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int i = i0 + rr;
            const int64_t row = ((int64_t)bh * Lq + i) * Lk;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int j = 2 * (lane + 32 * k);
                p[rr][k][0] = p[rr][k][1] = dp[rr][k][0] = dp[rr][k][1] = 0.f;
                if (i < Lq && j < Lk) {
                    if (i < len && j < len) {
                        float2 pp = ld_pair<T>(P + row + j);
                        if (row_scale) { const float rs = row_scale[(int64_t)bh * Lq + i]; pp.x *= rs; pp.y *= rs; }
                        if (j + 1 >= len) pp.y = 0.f;             // pair straddling the sequence end
                        const float2 g = *reinterpret_cast<const float2*>(dPd + row + j);
                        float g0 = g.x, g1 = g.y;
                        if (drop.thr) {
                            bool k0, k1;
                            drop_pair(drop, (uint64_t)(row + j) >> 1, k0, k1);
                            g0 = k0 ? g0 * drop.inv_keep : 0.f;
                            g1 = k1 ? g1 * drop.inv_keep : 0.f;
                            if (Pd_out) st_pair<T>(Pd_out + row + j, k0 ? pp.x * drop.inv_keep : 0.f, k1 ? pp.y * drop.inv_keep : 0.f);
                        } else if (Pd_out) {
                            st_pair<T>(Pd_out + row + j, pp.x, pp.y);   // packed, no dropout: clean copy of P
                        }
                        if (pp.y == 0.f) g1 = 0.f;
                        p[rr][k][0] = pp.x; p[rr][k][1] = pp.y; dp[rr][k][0] = g0; dp[rr][k][1] = g1;
                        dot[rr] += g0 * pp.x + g1 * pp.y;
                    } else if (Pd_out) {
                        st_pair<T>(Pd_out + row + j, 0.f, 0.f);   // outside the sequence: exact zeros
                    }
                }
            }
        }
        dot[0] = warp_sum(dot[0]);
        dot[1] = warp_sum(dot[1]);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int i = i0 + rr;
            const int64_t row = ((int64_t)bh * Lq + i) * Lk;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int j = 2 * (lane + 32 * k);
                if (i < Lq && j < Lk) {
                    const float d0 = p[rr][k][0] * (dp[rr][k][0] - dot[rr]), d1 = p[rr][k][1] * (dp[rr][k][1] - dot[rr]);
                    st_pair<T>(dS + row + j, d0, d1);
                    if (dbias_rel) { atomicAdd(&sdb[j - i + Lq - 1], d0); atomicAdd(&sdb[j + 1 - i + Lq - 1], d1); }
                }
            }
        }
    }
    if (dbias_rel) {
        __syncthreads();
        for (int e = threadIdx.x; e < n_delta; e += blockDim.x) {
            const float v = sdb[e];
            if (v != 0.f) atomicAdd(dbias_rel + h * n_delta + e, v);
        }
    }
}

// raw register form of a loaded probability pair: bf16 pairs stay packed (one register) until they are used
template <typename T> struct PairRaw;
template <> struct PairRaw<float> {
    float2 v;
    __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float2*>(p); }
    __device__ __forceinline__ void keep(bool c0, bool c1) { v.x = c0 ? v.x : 0.f; v.y = c1 ? v.y : 0.f; }
    __device__ __forceinline__ float2 get() const { return v; }
};
template <> struct PairRaw<bf16> {
    uint32_t v;
    __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint32_t*>(p); }
    __device__ __forceinline__ void keep(bool c0, bool c1) { v &= (c0 ? 0x0000ffffu : 0u) | (c1 ? 0xffff0000u : 0u); }
    __device__ __forceinline__ float2 get() const { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
};

// Square (Lq == Lk == 64 * NP) fast path of softmax_bwd for the encoder self-attention.  Two things made the generic
// kernel slow (ncu: 0.33 IPC, 78 regs, 16 ATOMS.CAST.SPIN loops per lane per row pair): nested bounds branches kept
// the loads from being batched, and the relative-bias gradient went through shared-memory float atomics.  Here
//  * every load of a row pair is issued up front, validity is applied with selects (no branches in the row loop);
//  * lane slots are SKEWED by the row: slot q of row pair m covers column pair (q + m) mod L/2, so j - i of a slot
//    is constant (2q-1, 2q, 2q+1) while the warp walks down the rows and the diagonal sums stay in registers; they
//    are flushed to shared memory only when a slot wraps around (once per slot) and at the end.
template <typename T, int NP>
__global__ void __launch_bounds__(256, (NP <= 4 ? 3 : 2))
softmax_bwd_sq_kernel(const float* __restrict__ dPd, const T* __restrict__ P, T* __restrict__ dS, T* __restrict__ Pd_out,
                      float* __restrict__ dbias_rel, int H, int L, DropCfg drop, const int* __restrict__ lens,
                      const float* __restrict__ row_scale) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    extern __shared__ float sdb[];  // [2L - 1]
    const int bh = blockIdx.x, h = bh % H;
    const int len = lens ? lens[bh / H] : L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int HALF = 32 * NP;   // column pairs per row
    const int n_delta = 2 * L - 1;
    if (dbias_rel) {
        for (int e = threadIdx.x; e < n_delta; e += blockDim.x) sdb[e] = 0.f;
        __syncthreads();
    }
    float acc[NP][3];
    bool wprev[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = 0.f; wprev[k] = false; }
    auto flush = [&](int k) {
        const int base = 2 * (lane + 32 * k) - 1 - (wprev[k] ? L : 0) + L - 1;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = acc[k][c];
            const int dl = base + c;
            if (v != 0.f && dl >= 0 && dl < n_delta) atomicAdd(&sdb[dl], v);
            acc[k][c] = 0.f;
        }
    };
    for (int m = blockIdx.y * nw + warp; m < HALF; m += nw * gridDim.y) {
        const int i0 = 2 * m;
        const int64_t row0 = ((int64_t)bh * L + i0) * L;
        if (i0 >= len) {   // warp-uniform: both rows lie outside the sequence -> exact zeros
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    const int64_t off = row0 + (int64_t)rr * L + 2 * (lane + 32 * k);
                    st_pair<T>(dS + off, 0.f, 0.f);
                    if (Pd_out) st_pair<T>(Pd_out + off, 0.f, 0.f);
                }
            continue;
        }
        const T* __restrict__ Pr = P + row0;
        const float* __restrict__ Gr = dPd + row0;
        T* __restrict__ dSr = dS + row0;
        T* __restrict__ Pdr = Pd_out ? Pd_out + row0 : nullptr;
        const uint32_t pair0 = (uint32_t)((uint64_t)row0 >> 1);   // pair index of the row start (B*H*L*L/2 < 2^32)
        PairRaw<T> pp[2][NP];
        float2 g[2][NP];
        float rs[2] = {1.f, 1.f};      // P = P_raw * row_scale (un-normalised probabilities saved by fattn_fwd)
        if (row_scale) {
            rs[0] = row_scale[(int64_t)bh * L + i0];
            rs[1] = i0 + 1 < len ? row_scale[(int64_t)bh * L + i0 + 1] : 0.f;
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int o = rr * L + 2 * ((lane + 32 * k + m) & (HALF - 1));
                pp[rr][k].load(Pr + o);
                g[rr][k] = *reinterpret_cast<const float2*>(Gr + o);
            }
        float dot[2] = {0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const bool rv = i0 + rr < len;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int j = 2 * ((lane + 32 * k + m) & (HALF - 1));
                const int o = rr * L + j;
                const bool c0 = rv && j < len, c1 = rv && j + 1 < len;
                pp[rr][k].keep(c0, c1);
                const float2 pv = pp[rr][k].get();
                const float p0 = pv.x * rs[rr], p1 = pv.y * rs[rr];
                float g0 = c0 ? g[rr][k].x : 0.f, g1 = c1 ? g[rr][k].y : 0.f;
                if (drop.thr) {
                    bool k0, k1;
                    drop_pair(drop, (uint64_t)(pair0 + (uint32_t)(o >> 1)), k0, k1);
                    g0 = k0 ? g0 * drop.inv_keep : 0.f;
                    g1 = k1 ? g1 * drop.inv_keep : 0.f;
                    if (Pdr) st_pair<T>(Pdr + o, k0 ? p0 * drop.inv_keep : 0.f, k1 ? p1 * drop.inv_keep : 0.f);
                } else if (Pdr) {
                    st_pair<T>(Pdr + o, p0, p1);
                }
                g[rr][k] = make_float2(g0, g1);
                dot[rr] += g0 * p0 + g1 * p1;
            }
        }
        dot[0] = warp_sum(dot[0]);
        dot[1] = warp_sum(dot[1]);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int qm = lane + 32 * k + m;
            const int j = 2 * (qm & (HALF - 1));
            const bool w = qm >= HALF;
            if (dbias_rel && w != wprev[k]) { flush(k); wprev[k] = w; }
            const float2 pe = pp[0][k].get(), po = pp[1][k].get();
            const float e0 = pe.x * rs[0] * (g[0][k].x - dot[0]), e1 = pe.y * rs[0] * (g[0][k].y - dot[0]);   // row 2m
            const float o0 = po.x * rs[1] * (g[1][k].x - dot[1]), o1 = po.y * rs[1] * (g[1][k].y - dot[1]);   // row 2m+1
            st_pair<T>(dSr + j, e0, e1);
            st_pair<T>(dSr + L + j, o0, o1);
            acc[k][0] += o0;          // j - i = 2q - 1
            acc[k][1] += e0 + o1;     //         2q
            acc[k][2] += e1;          //         2q + 1
        }
    }
    if (dbias_rel) {
#pragma unroll
        for (int k = 0; k < NP; ++k) flush(k);
        __syncthreads();
        for (int e = threadIdx.x; e < n_delta; e += blockDim.x) {
            const float v = sdb[e];
            if (v != 0.f) atomicAdd(dbias_rel + h * n_delta + e, v);
        }
    }
}

void softmax_bwd(const float* dPd, const void* P, void* dS, void* Pd_out, int dtype, float* dbias_rel, int B, int H,
                 int Lq, int Lk, DropCfg drop, cudaStream_t st, const int* lens, const float* row_scale) {
    if (B <= 0) return;
    P5_CHECK(Lk <= 64 * SM_MAXP && (Lk % 2) == 0, "softmax_bwd: Lk must be even and <= 512");
    const size_t sm = (size_t)(Lq + Lk) * sizeof(float);
    static const bool generic_only = getenv("P5_SMBWD_GENERIC") != nullptr;
    if (Lq == Lk && (Lk == 64 || Lk == 128 || Lk == 256 || Lk == 512) && !generic_only) {
        dim3 g2((unsigned)(B * H), (unsigned)(Lk >= 256 ? 4 : (Lk >= 128 ? 2 : 1)));
#define P5_SMBWD_SQ(TT, NPV) launch_k(softmax_bwd_sq_kernel<TT, NPV>, g2, 256, sm, st, dPd, (const TT*)P, (TT*)dS, (TT*)Pd_out, dbias_rel, H, Lk, drop, lens, row_scale)
        if (dtype == DT_F32) {
            if (Lk == 64) P5_SMBWD_SQ(float, 1); else if (Lk == 128) P5_SMBWD_SQ(float, 2); else if (Lk == 256) P5_SMBWD_SQ(float, 4); else P5_SMBWD_SQ(float, 8);
        } else {
            if (Lk == 64) P5_SMBWD_SQ(bf16, 1); else if (Lk == 128) P5_SMBWD_SQ(bf16, 2); else if (Lk == 256) P5_SMBWD_SQ(bf16, 4); else P5_SMBWD_SQ(bf16, 8);
        }
#undef P5_SMBWD_SQ
        LAUNCHED();
        return;
    }
    dim3 grid((unsigned)(B * H), (unsigned)(Lq >= 128 ? 2 : 1));
    if (dtype == DT_F32) {
        if (Lk <= 256) launch_k(softmax_bwd_kernel<float, 4>, grid, 256, sm, st, dPd, (const float*)P, (float*)dS, (float*)Pd_out, dbias_rel, H, Lq, Lk, drop, lens, row_scale);
        else launch_k(softmax_bwd_kernel<float, 8>, grid, 256, sm, st, dPd, (const float*)P, (float*)dS, (float*)Pd_out, dbias_rel, H, Lq, Lk, drop, lens, row_scale);
    } else {
        if (Lk <= 256) launch_k(softmax_bwd_kernel<bf16, 4>, grid, 256, sm, st, dPd, (const bf16*)P, (bf16*)dS, (bf16*)Pd_out, dbias_rel, H, Lq, Lk, drop, lens, row_scale);
        else launch_k(softmax_bwd_kernel<bf16, 8>, grid, 256, sm, st, dPd, (const bf16*)P, (bf16*)dS, (bf16*)Pd_out, dbias_rel, H, Lq, Lk, drop, lens, row_scale);
    }
    LAUNCHED();
}

}  // namespace p5
