// HBM-bound kernels of the T5 block: embedding gather, RMSNorm fwd/bwd, dropout-cast, gated-GELU, cross-entropy,
// runner loss, relative-position-bias tables.  Coalesced / vectorised; statistics in fp32.
#include "kernels.cuh"
#include <algorithm>

namespace p5 {
extern int g_launches;
#define LAUNCHED() do { P5_CUDA(cudaGetLastError()); ++g_launches; } while (0)

// ---- small vector helpers ------------------------------------------------------------------------------------
template <int VEC> __device__ __forceinline__ void ldv(const float* p, float* v) {
    if constexpr (VEC == 4) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else v[0] = p[0];
}
template <int VEC> __device__ __forceinline__ void ldv(const bf16* p, float* v) {
    if constexpr (VEC == 4) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
        float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else v[0] = __bfloat162float(p[0]);
}
template <int VEC> __device__ __forceinline__ void stv(float* p, const float* v) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else p[0] = v[0];
}
template <int VEC> __device__ __forceinline__ void stv(bf16* p, const float* v) {
    if constexpr (VEC == 4) {
        uint2 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
        h[0] = __floats2bfloat162_rn(v[0], v[1]);
        h[1] = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = t;
    } else p[0] = __float2bfloat16_rn(v[0]);
}

// =================================================================================================================
// embeddings
// =================================================================================================================
template <int VEC>
__global__ void embed_fwd_kernel(const float* __restrict__ E, const float* __restrict__ W, const int* __restrict__ ids,
                                 const int* __restrict__ ww, float* __restrict__ x, int M, int d, int vocab,
                                 int ww_rows, DropCfg drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    int id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    int w = ww ? ww[row] : -1;
    if (w >= ww_rows) w = ww_rows - 1;
    const float* e = E + (int64_t)id * d;
    const float* wr = (w >= 0) ? W + (int64_t)w * d : nullptr;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float v[VEC], u[VEC];
        ldv<VEC>(e + c, v);
        if (wr) {
            ldv<VEC>(wr + c, u);
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] += u[j];
        }
        if (drop.thr) {
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                v[j] = drop_keep(drop.seed, drop.site, (uint64_t)row * d + c + j, drop.thr) ? v[j] * drop.inv_keep : 0.f;
        }
        stv<VEC>(x + (int64_t)row * d + c, v);
    }
}

void embed_fwd(const float* E, const float* Wword, const int* ids, const int* ww, float* x, int M, int d, int vocab,
               int ww_rows, DropCfg drop, cudaStream_t st) {
    if (M <= 0) return;
    const int rows_per_block = 8;
    dim3 grid((unsigned)cdiv(M, rows_per_block));
    if (d % 128 == 0)
        launch_k(embed_fwd_kernel<4>, grid, rows_per_block * 32, 0, st, E, Wword, ids, ww, x, M, d, vocab, ww_rows, drop);
    else
        launch_k(embed_fwd_kernel<1>, grid, rows_per_block * 32, 0, st, E, Wword, ids, ww, x, M, d, vocab, ww_rows, drop);
    LAUNCHED();
}

__global__ void embed_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ ids, const int* __restrict__ ww,
                                 float* __restrict__ dE, float* __restrict__ dW, int M, int d, int vocab, int ww_rows,
                                 DropCfg drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    int id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    int w = ww ? ww[row] : -1;
    if (w >= ww_rows) w = ww_rows - 1;
    for (int c = lane; c < d; c += 32) {
        float g = dx[(int64_t)row * d + c];
        if (drop.thr) g = drop_keep(drop.seed, drop.site, (uint64_t)row * d + c, drop.thr) ? g * drop.inv_keep : 0.f;
        if (g != 0.f) {
            atomicAdd(dE + (int64_t)id * d + c, g);
            if (w >= 0) atomicAdd(dW + (int64_t)w * d + c, g);
        }
    }
}

void embed_bwd(const float* dx, const int* ids, const int* ww, float* dE, float* dWword, int M, int d, int vocab,
               int ww_rows, DropCfg drop, cudaStream_t st) {
    if (M <= 0) return;
    launch_k(embed_bwd_kernel, (unsigned)cdiv(M, 8), 256, 0, st, dx, ids, ww, dE, dWword, M, d, vocab, ww_rows, drop);
    LAUNCHED();
}

// =================================================================================================================
// RMSNorm
// =================================================================================================================
template <typename T, int VEC>
__global__ void rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ n,
                                   float* __restrict__ rstd, int M, int d, float eps, DropCfg drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + (int64_t)row * d;
    float ss = 0.f;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float v[VEC];
        ldv<VEC>(xr + c, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) ss += v[j] * v[j];
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / (float)d + eps);
    if (rstd && lane == 0) rstd[row] = r;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float v[VEC], g[VEC];
        ldv<VEC>(xr + c, v);
        ldv<VEC>(w + c, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            v[j] = g[j] * (v[j] * r);
            if (drop.thr)
                v[j] = drop_keep(drop.seed, drop.site, (uint64_t)row * d + c + j, drop.thr) ? v[j] * drop.inv_keep : 0.f;
        }
        stv<VEC>(n + (int64_t)row * d + c, v);
    }
}

void rmsnorm_fwd(const float* x, const float* w, void* n, int n_dtype, float* rstd, int M, int d, float eps,
                 DropCfg drop, cudaStream_t st) {
    if (M <= 0) return;
    dim3 grid((unsigned)cdiv(M, 8));
    const bool v4 = d % 128 == 0;
    if (n_dtype == DT_F32) {
        if (v4) launch_k(rmsnorm_fwd_kernel<float, 4>, grid, 256, 0, st, x, w, (float*)n, rstd, M, d, eps, drop);
        else launch_k(rmsnorm_fwd_kernel<float, 1>, grid, 256, 0, st, x, w, (float*)n, rstd, M, d, eps, drop);
    } else {
        if (v4) launch_k(rmsnorm_fwd_kernel<bf16, 4>, grid, 256, 0, st, x, w, (bf16*)n, rstd, M, d, eps, drop);
        else launch_k(rmsnorm_fwd_kernel<bf16, 1>, grid, 256, 0, st, x, w, (bf16*)n, rstd, M, d, eps, drop);
    }
    LAUNCHED();
}

// One CTA = 8 warps handles RB rows; each warp owns rows (warp, warp+8, ...).  dw partials are reduced through
// shared memory and flushed with one atomicAdd per column per CTA.
static constexpr int RMS_BWD_ROWS = 32;
template <typename T>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const T* __restrict__ dn, const float* __restrict__ x, const float* __restrict__ rstd,
                   const float* __restrict__ w, const float* dres, float* dx,   // dres may alias dx (in-place)
                   float* __restrict__ dw, int M, int d, DropCfg drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    extern __shared__ float sdw[];  // [d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < d; c += blockDim.x) sdw[c] = 0.f;
    __syncthreads();
    const int r0 = blockIdx.x * RMS_BWD_ROWS;
    float acc[32];  // per-lane dw partials for columns lane + 32*k (d <= 1024)
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    for (int rr = warp; rr < RMS_BWD_ROWS; rr += 8) {
        const int row = r0 + rr;
        if (row >= M) break;
        const float* xr = x + (int64_t)row * d;
        const T* dr = dn + (int64_t)row * d;
        const float r = rstd[row];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const int c = lane + 32 * k;
            if (c < d) {
                float g = to_f32(dr[c]);
                if (drop.thr) g = drop_keep(drop.seed, drop.site, (uint64_t)row * d + c, drop.thr) ? g * drop.inv_keep : 0.f;
                const float xh = xr[c] * r;
                dot += g * w[c] * xh;
                acc[k] += g * xh;
            }
        }
        dot = warp_sum(dot) / (float)d;
        for (int c = lane; c < d; c += 32) {
            float g = to_f32(dr[c]);
            if (drop.thr) g = drop_keep(drop.seed, drop.site, (uint64_t)row * d + c, drop.thr) ? g * drop.inv_keep : 0.f;
            const float xh = xr[c] * r;
            float v = r * (g * w[c] - xh * dot);
            if (dres) v += dres[(int64_t)row * d + c];
            dx[(int64_t)row * d + c] = v;
        }
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int c = lane + 32 * k;
        if (c < d && acc[k] != 0.f) atomicAdd(&sdw[c], acc[k]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float v = sdw[c];
        if (v != 0.f) atomicAdd(dw + c, v);
    }
}

// vectorised form for d = NV * 128: each lane owns NV float4 column groups, the whole row lives in registers
// (x, dn, dres are each read exactly once: 2+4+4 B in, 4 B out per element), dw partials stay in registers per warp
template <typename T, int NV>
__global__ void __launch_bounds__(256, 2)
rmsnorm_bwd_vec_kernel(const T* __restrict__ dn, const float* __restrict__ x, const float* __restrict__ rstd,
                       const float* __restrict__ w, const float* dres, float* dx, float* __restrict__ dw, int M,
                       DropCfg drop, bf16* __restrict__ dx_cast, DropCfg cast_drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    constexpr int d = NV * 128;
    __shared__ __align__(16) float sdw[8][d];     // per-warp dw partials (summed once at the end; shared fp32 atomics are CAS loops)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const DropKey dkey = drop_key(drop.seed, drop.site);     // once per kernel, not per element
    const uint32_t thr_hi = drop.thr & 0xffff0000u;
    float acc[NV][4];     // the norm weight (3 KB) is re-read from L1 per row instead of pinning NV*4 registers
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
    // persistent rows: the grid is sized to the resident CTAs (2 per SM) and every warp strides over the rows, so the
    // last wave is as full as the first (400 x 32-row CTAs on 148 SMs were 2.7 waves)
    for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
        const int64_t base = (int64_t)row * d;
        float xv[NV][4], gv[NV][4], rv[NV][4];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = 4 * (lane + 32 * k);
            ldv<4>(x + base + c, xv[k]);
            ldv<4>(dn + base + c, gv[k]);
            if (dres) ldv<4>(dres + base + c, rv[k]);
        }
        const float r = rstd[row];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = 4 * (lane + 32 * k);
            float wk[4];
            ldv<4>(w + c, wk);
            if (drop.thr) {     // one hash per aligned pair of elements (base + c is a multiple of 4): same mask as drop_keep
#pragma unroll
                for (int j2 = 0; j2 < 2; ++j2) {
                    const uint32_t hsh = drop_hash_k(dkey, (uint32_t)((uint64_t)(base + c + 2 * j2) >> 1));
                    gv[k][2 * j2] = (hsh << 16) >= thr_hi ? gv[k][2 * j2] * drop.inv_keep : 0.f;
                    gv[k][2 * j2 + 1] = hsh >= thr_hi ? gv[k][2 * j2 + 1] * drop.inv_keep : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float g = gv[k][j];
                xv[k][j] *= r;                  // xhat
                dot += g * wk[j] * xv[k][j];
                acc[k][j] += g * xv[k][j];
            }
        }
        dot = warp_sum(dot) * (1.f / (float)d);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = 4 * (lane + 32 * k);
            float o[4], wk[4];
            ldv<4>(w + c, wk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = r * (gv[k][j] * wk[j] - xv[k][j] * dot);
                if (dres) o[j] += rv[k][j];
            }
            stv<4>(dx + base + c, o);
            if (dx_cast) {   // bf16 copy of dx with the NEXT backward op's dropout mask applied (saves a drop_cast pass)
                if (cast_drop.thr) {
                    bool kk[4];
                    drop_keep2(cast_drop.seed, cast_drop.site, (uint64_t)(base + c), cast_drop.thr, kk[0], kk[1]);
                    drop_keep2(cast_drop.seed, cast_drop.site, (uint64_t)(base + c) + 2, cast_drop.thr, kk[2], kk[3]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = kk[j] ? o[j] * cast_drop.inv_keep : 0.f;
                }
                stv<4>(dx_cast + base + c, o);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k)
        *reinterpret_cast<float4*>(&sdw[warp][4 * (lane + 32 * k)]) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += 256) {
        float v = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) v += sdw[w8][c];
        if (v != 0.f) atomicAdd(dw + c, v);
    }
}

template <typename T>
static bool launch_rms_bwd_vec(const void* dn, const float* x, const float* rstd, const float* w, const float* dres,
                               float* dx, float* dw, int M, int d, DropCfg drop, bf16* dx_cast, DropCfg cast_drop,
                               cudaStream_t st) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int64_t want = cdiv(M, 8);
    dim3 grid((unsigned)(want < 2 * sms ? want : 2 * sms));
    switch (d) {
        case 512: launch_k(rmsnorm_bwd_vec_kernel<T, 4>, grid, 256, 0, st, (const T*)dn, x, rstd, w, dres, dx, dw, M, drop, dx_cast, cast_drop); return true;
        case 768: launch_k(rmsnorm_bwd_vec_kernel<T, 6>, grid, 256, 0, st, (const T*)dn, x, rstd, w, dres, dx, dw, M, drop, dx_cast, cast_drop); return true;
        case 1024: launch_k(rmsnorm_bwd_vec_kernel<T, 8>, grid, 256, 0, st, (const T*)dn, x, rstd, w, dres, dx, dw, M, drop, dx_cast, cast_drop); return true;
        default: return false;
    }
}

void rmsnorm_bwd(const void* dn, int dn_dtype, const float* x, const float* rstd, const float* w, const float* dres,
                 float* dx, float* dw, int M, int d, DropCfg drop, cudaStream_t st, void* dx_cast, int cast_dtype,
                 DropCfg cast_drop) {
    if (M <= 0) return;
    bf16* fused_cast = (dx_cast && cast_dtype == DT_BF16 && (d == 512 || d == 768 || d == 1024)) ? (bf16*)dx_cast : nullptr;
    if (dn_dtype == DT_F32 ? launch_rms_bwd_vec<float>(dn, x, rstd, w, dres, dx, dw, M, d, drop, fused_cast, cast_drop, st)
                           : launch_rms_bwd_vec<bf16>(dn, x, rstd, w, dres, dx, dw, M, d, drop, fused_cast, cast_drop, st)) {
        LAUNCHED();
        if (dx_cast && !fused_cast) drop_cast(dx, dx_cast, cast_dtype, (int64_t)M * d, cast_drop, st);
        return;
    }
    dim3 grid((unsigned)cdiv(M, RMS_BWD_ROWS));
    const size_t sm = (size_t)d * sizeof(float);
    if (dn_dtype == DT_F32)
        launch_k(rmsnorm_bwd_kernel<float>, grid, 256, sm, st, (const float*)dn, x, rstd, w, dres, dx, dw, M, d, drop);
    else
        launch_k(rmsnorm_bwd_kernel<bf16>, grid, 256, sm, st, (const bf16*)dn, x, rstd, w, dres, dx, dw, M, d, drop);
    LAUNCHED();
    if (dx_cast) drop_cast(dx, dx_cast, cast_dtype, (int64_t)M * d, cast_drop, st);
}

// =================================================================================================================
// casts / dropout-cast
// =================================================================================================================
template <typename T>
__global__ void drop_cast_kernel(const float* __restrict__ in, T* __restrict__ out, int64_t n, DropCfg drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            float v[4];
            ldv<4>(in + i, v);
            if (drop.thr) {
                bool k[4];
                drop_keep2(drop.seed, drop.site, (uint64_t)i, drop.thr, k[0], k[1]);
                drop_keep2(drop.seed, drop.site, (uint64_t)i + 2, drop.thr, k[2], k[3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = k[j] ? v[j] * drop.inv_keep : 0.f;
            }
            stv<4>(out + i, v);
        } else {
            for (int64_t k = i; k < n; ++k) {
                float v = in[k];
                if (drop.thr) v = drop_keep(drop.seed, drop.site, (uint64_t)k, drop.thr) ? v * drop.inv_keep : 0.f;
                out[k] = from_f32<T>(v);
            }
        }
    }
}
static inline unsigned ew_grid(int64_t n, int per_thread) {
    int64_t b = cdiv(n, (int64_t)256 * per_thread);
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (unsigned)b;
}
void drop_cast(const float* in, void* out, int out_dtype, int64_t n, DropCfg drop, cudaStream_t st) {
    if (n <= 0) return;
    if (out_dtype == DT_F32) launch_k(drop_cast_kernel<float>, ew_grid(n, 4), 256, 0, st, in, (float*)out, n, drop);
    else launch_k(drop_cast_kernel<bf16>, ew_grid(n, 4), 256, 0, st, in, (bf16*)out, n, drop);
    LAUNCHED();
}
void cast_f32_to(const float* in, void* out, int out_dtype, int64_t n, cudaStream_t st) {
    DropCfg none;
    drop_cast(in, out, out_dtype, n, none, st);
}
__global__ void cast_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, int64_t n) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}
void cast_to_f32(const void* in, int in_dtype, float* out, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    if (in_dtype == DT_F32) {
        P5_CUDA(cudaMemcpyAsync(out, in, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        return;
    }
    launch_k(cast_to_f32_kernel, ew_grid(n, 1), 256, 0, st, (const bf16*)in, out, n);
    LAUNCHED();
}
template <typename T>
__global__ void cast_block_kernel(const float* __restrict__ in, int64_t ld_in, T* __restrict__ out, int64_t ld_out,
                                  int rows, int cols) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int64_t n = (int64_t)rows * cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        out[(int64_t)r * ld_out + c] = from_f32<T>(in[(int64_t)r * ld_in + c]);
    }
}
void cast_block_f32_to(const float* in, int64_t ld_in, void* out, int out_dtype, int64_t ld_out, int rows, int cols,
                       cudaStream_t st) {
    const int64_t n = (int64_t)rows * cols;
    if (n <= 0) return;
    if (out_dtype == DT_F32) launch_k(cast_block_kernel<float>, ew_grid(n, 1), 256, 0, st, in, ld_in, (float*)out, ld_out, rows, cols);
    else launch_k(cast_block_kernel<bf16>, ew_grid(n, 1), 256, 0, st, in, ld_in, (bf16*)out, ld_out, rows, cols);
    LAUNCHED();
}
__global__ void add_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}
void add_f32(float* dst, const float* src, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    launch_k(add_f32_kernel, ew_grid(n, 1), 256, 0, st, dst, src, n);
    LAUNCHED();
}

// =================================================================================================================
// packed (variable-length) <-> padded row layouts.  offs[b] = first packed row of sequence b, lens[b] = valid tokens
// =================================================================================================================
// dst_packed[offs[b] + i, :] = src_padded[b*L + i, :]  (i < lens[b]); W elements per row, 16-byte chunks
__global__ void pack_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ offs,
                                 const int* __restrict__ lens, int L, int chunks) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int len = lens[b];
    const int64_t n = (int64_t)len * chunks;
    const uint4* s = src + (int64_t)b * L * chunks;
    uint4* d = dst + (int64_t)offs[b] * chunks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = s[i];
}
// dst_padded[b*L + i, :] = i < lens[b] ? src_packed[offs[b] + i, :] : 0
__global__ void unpack_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ offs,
                                   const int* __restrict__ lens, int L, int chunks) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int len = lens[b];
    const int64_t n = (int64_t)L * chunks, nv = (int64_t)len * chunks;
    const uint4* s = src + (int64_t)offs[b] * chunks;
    uint4* d = dst + (int64_t)b * L * chunks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        d[i] = i < nv ? s[i] : make_uint4(0, 0, 0, 0);
}
void pack_rows(const void* src_padded, void* dst_packed, const int* offs, const int* lens, int B, int L, int64_t row_bytes,
               cudaStream_t st) {
    if (B <= 0) return;
    P5_CHECK(row_bytes % 16 == 0, "pack_rows: row size must be a multiple of 16 bytes");
    const int chunks = (int)(row_bytes / 16);
    dim3 grid((unsigned)std::min<int64_t>(64, cdiv((int64_t)L * chunks, 256)), (unsigned)B);
    launch_k(pack_rows_kernel, grid, 256, 0, st, (const uint4*)src_padded, (uint4*)dst_packed, offs, lens, L, chunks);
    LAUNCHED();
}
void unpack_rows(const void* src_packed, void* dst_padded, const int* offs, const int* lens, int B, int L, int64_t row_bytes,
                 cudaStream_t st) {
    if (B <= 0) return;
    P5_CHECK(row_bytes % 16 == 0, "unpack_rows: row size must be a multiple of 16 bytes");
    const int chunks = (int)(row_bytes / 16);
    dim3 grid((unsigned)std::min<int64_t>(64, cdiv((int64_t)L * chunks, 256)), (unsigned)B);
    launch_k(unpack_rows_kernel, grid, 256, 0, st, (const uint4*)src_packed, (uint4*)dst_padded, offs, lens, L, chunks);
    LAUNCHED();
}
// packed int arrays from padded [B, L] (token ids / whole-word ids)
__global__ void pack_ints_kernel(const int* __restrict__ src, int* __restrict__ dst, const int* __restrict__ offs,
                                 const int* __restrict__ lens, int L) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < lens[b]; i += blockDim.x) dst[offs[b] + i] = src[b * L + i];
}
void pack_ints(const int* src_padded, int* dst_packed, const int* offs, const int* lens, int B, int L, cudaStream_t st) {
    if (B <= 0) return;
    launch_k(pack_ints_kernel, B, 256, 0, st, src_padded, dst_packed, offs, lens, L);
    LAUNCHED();
}

// =================================================================================================================
// gated GELU (T5 v1.1 FFN)
// =================================================================================================================
__device__ __forceinline__ float gelu_new_f(float x) {
    const float k = 0.7978845608028654f;  // sqrt(2/pi)
    return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float gelu_new_grad(float x) {
    const float k = 0.7978845608028654f;
    const float u = k * (x + 0.044715f * x * x * x);
    const float t = tanhf(u);
    return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k * (1.f + 3.f * 0.044715f * x * x);
}
template <typename T>
__global__ void gated_gelu_fwd_kernel(const T* __restrict__ z, T* __restrict__ h, int M, int ff, DropCfg drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int64_t n = (int64_t)M * ff;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / ff, c = i % ff;
        float v = gelu_new_f(to_f32(z[r * 2 * ff + c])) * to_f32(z[r * 2 * ff + ff + c]);
        if (drop.thr) v = drop_keep(drop.seed, drop.site, (uint64_t)i, drop.thr) ? v * drop.inv_keep : 0.f;
        h[i] = from_f32<T>(v);
    }
}
template <typename T>
__global__ void gated_gelu_bwd_kernel(const T* __restrict__ z, const T* __restrict__ dh, T* __restrict__ dz, int M,
                                      int ff, DropCfg drop) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int64_t n = (int64_t)M * ff;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / ff, c = i % ff;
        float g = to_f32(dh[i]);
        if (drop.thr) g = drop_keep(drop.seed, drop.site, (uint64_t)i, drop.thr) ? g * drop.inv_keep : 0.f;
        const float z0 = to_f32(z[r * 2 * ff + c]), z1 = to_f32(z[r * 2 * ff + ff + c]);
        dz[r * 2 * ff + c] = from_f32<T>(g * z1 * gelu_new_grad(z0));
        dz[r * 2 * ff + ff + c] = from_f32<T>(g * gelu_new_f(z0));
    }
}
void gated_gelu_fwd(const void* z, void* h, int dtype, int M, int ff, DropCfg drop, cudaStream_t st) {
    const int64_t n = (int64_t)M * ff;
    if (n <= 0) return;
    if (dtype == DT_F32) launch_k(gated_gelu_fwd_kernel<float>, ew_grid(n, 1), 256, 0, st, (const float*)z, (float*)h, M, ff, drop);
    else launch_k(gated_gelu_fwd_kernel<bf16>, ew_grid(n, 1), 256, 0, st, (const bf16*)z, (bf16*)h, M, ff, drop);
    LAUNCHED();
}
void gated_gelu_bwd(const void* z, const void* dh, void* dz, int dtype, int M, int ff, DropCfg drop, cudaStream_t st) {
    const int64_t n = (int64_t)M * ff;
    if (n <= 0) return;
    if (dtype == DT_F32)
        launch_k(gated_gelu_bwd_kernel<float>, ew_grid(n, 1), 256, 0, st, (const float*)z, (const float*)dh, (float*)dz, M, ff, drop);
    else
        launch_k(gated_gelu_bwd_kernel<bf16>, ew_grid(n, 1), 256, 0, st, (const bf16*)z, (const bf16*)dh, (bf16*)dz, M, ff, drop);
    LAUNCHED();
}

// =================================================================================================================
// cross entropy over the item-token vocabulary (one CTA per decoder position)
// =================================================================================================================
// one CTA per target position: ONE pass over the row with 16-byte loads and an online (max, sum) per thread
__global__ void __launch_bounds__(512)
ce_fwd_kernel(const float* __restrict__ logits, int64_t ld, const int* __restrict__ labels, float* __restrict__ loss,
              float* __restrict__ lse, int V) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    __shared__ float sh[32];
    const int row = blockIdx.x;
    const float* l = logits + (int64_t)row * ld;
    float mx = -INFINITY, s = 0.f;
    const int V4 = V >> 2;
    const float4* l4 = reinterpret_cast<const float4*>(l);       // rows start 16-byte aligned (ld % 4 == 0)
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
    if (threadIdx.x == 0) {
        const float z = bm + logf(s);
        lse[row] = z;
        const int y = labels[row];
        loss[row] = (y >= 0 && y < V) ? z - l[y] : 0.f;  // ignore_index = -100 (never produced by the collator)
    }
}
void ce_fwd(const float* logits, int64_t ld, const int* labels, float* loss_tok, float* lse, int M, int V,
            cudaStream_t st) {
    if (M <= 0) return;
    P5_CHECK(ld % 4 == 0, "ce_fwd: logits rows must be 16-byte aligned");
    launch_k(ce_fwd_kernel, M, 512, 0, st, logits, ld, labels, loss_tok, lse, V);
    LAUNCHED();
}
template <typename T>
__global__ void __launch_bounds__(512)
ce_bwd_kernel(const float* __restrict__ logits, int64_t ld, const float* __restrict__ lse,
              const int* __restrict__ labels, const float* __restrict__ dloss, T* __restrict__ dlogits, int V,
              int Vpad) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int row = blockIdx.x;
    const float* l = logits + (int64_t)row * ld;
    T* o = dlogits + (int64_t)row * Vpad;
    const int y = labels[row];
    const float g = (y >= 0 && y < V) ? dloss[row] : 0.f;
    const float z = lse[row];
    // four columns per thread and iteration: 16-byte loads, 8- / 16-byte stores (Vpad % 4 == 0, rows 16-byte aligned)
    for (int c = 4 * threadIdx.x; c < Vpad; c += 4 * blockDim.x) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c + 4 <= V) {
            const float4 x = *reinterpret_cast<const float4*>(l + c);
            v[0] = g * __expf(x.x - z); v[1] = g * __expf(x.y - z); v[2] = g * __expf(x.z - z); v[3] = g * __expf(x.w - z);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + j < V) v[j] = g * __expf(l[c + j] - z);
        }
        if (y >= c && y < c + 4) v[y - c] -= g;
        stv<4>(o + c, v);
    }
}
void ce_bwd(const float* logits, int64_t ld, const float* lse, const int* labels, const float* dloss, void* dlogits,
            int d_dtype, int M, int V, int Vpad, cudaStream_t st) {
    if (M <= 0) return;
    P5_CHECK(ld % 4 == 0 && Vpad % 4 == 0, "ce_bwd: rows must be 16-byte aligned");
    if (d_dtype == DT_F32) launch_k(ce_bwd_kernel<float>, M, 512, 0, st, logits, ld, lse, labels, dloss, (float*)dlogits, V, Vpad);
    else launch_k(ce_bwd_kernel<bf16>, M, 512, 0, st, logits, ld, lse, labels, dloss, (bf16*)dlogits, V, Vpad);
    LAUNCHED();
}

// one block; B*Ld is small (<= 64K)
__global__ void runner_loss_kernel(const float* __restrict__ loss_tok, const int* __restrict__ mask, int B, int Ld,
                                   float* __restrict__ loss_out, float* __restrict__ dloss) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    __shared__ float sh[32];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float s = 0.f, cnt = 0.f;
        for (int t = 0; t < Ld; ++t) {
            const float m = mask[b * Ld + t] != 0 ? 1.f : 0.f;
            s += loss_tok[b * Ld + t] * m;
            cnt += m;
        }
        const float den = fmaxf(cnt, 1.f);
        acc += s / den;
        if (dloss)
            for (int t = 0; t < Ld; ++t) dloss[b * Ld + t] = (mask[b * Ld + t] != 0 ? 1.f : 0.f) / den / (float)B;
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) loss_out[0] = acc / (float)B;
}
void runner_loss_fwd_bwd(const float* loss_tok, const int* labels_mask, int B, int Ld, float* loss_out,
                         float* dloss_tok, cudaStream_t st) {
    launch_k(runner_loss_kernel, 1, 256, 0, st, loss_tok, labels_mask, B, Ld, loss_out, dloss_tok);
    LAUNCHED();
}

// =================================================================================================================
// relative position bias tables
// =================================================================================================================
__global__ void relbias_build_kernel(const float* __restrict__ table, const int* __restrict__ lut,
                                     float* __restrict__ bias_rel, int H, int n_delta) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * n_delta) return;
    const int h = i / n_delta, dlt = i % n_delta;
    bias_rel[i] = table[lut[dlt] * H + h];
}
void relbias_build(const float* table, const int* bucket_lut, float* bias_rel, int H, int n_delta, cudaStream_t st) {
    launch_k(relbias_build_kernel, (unsigned)cdiv(H * n_delta, 256), 256, 0, st, table, bucket_lut, bias_rel, H, n_delta);
    LAUNCHED();
}
__global__ void relbias_scatter_kernel(const float* __restrict__ dbias_rel, const int* __restrict__ lut,
                                       float* __restrict__ dtable, int H, int n_delta) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * n_delta) return;
    const int h = i / n_delta, dlt = i % n_delta;
    const float g = dbias_rel[i];
    if (g != 0.f) atomicAdd(dtable + lut[dlt] * H + h, g);
}
void relbias_scatter_grad(const float* dbias_rel, const int* bucket_lut, float* dtable, int H, int n_delta,
                          cudaStream_t st) {
    launch_k(relbias_scatter_kernel, (unsigned)cdiv(H * n_delta, 256), 256, 0, st, dbias_rel, bucket_lut, dtable, H, n_delta);
    LAUNCHED();
}

}  // namespace p5
