// Optimiser-side kernels over the FLAT parameter / gradient buffers (all parameters of the model live in one
// contiguous fp32 allocation, so clip_grad_norm_ + AdamW are two coalesced passes instead of ~260 small ones).
//   - sumsq_norm : global L2 norm of the gradients          (ref DistributedRunner.py:81, clip_grad_norm_)
//   - adamw_flat : transformers==4.26 AdamW.step semantics   (ref SingleRunner.py:191-214; formula SURVEY §8a-11)
//                  + optional fused clip coefficient + bf16 shadow write for the tensor-core GEMMs
#include "kernels.cuh"

namespace p5 {
extern int g_launches;
#define LAUNCHED() do { P5_CUDA(cudaGetLastError()); ++g_launches; } while (0)

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    __shared__ float sh[32];
    float s = 0.f;
    const int64_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = g4[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += g[i] * g[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partial, int np, float* __restrict__ out) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    __shared__ float sh[32];
    // accumulate the per-block partials in double: 1e3 partials of ~1e5 elements each
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) s += (double)partial[i];
    float f = (float)s;
    // block reduce (float is enough after the per-thread double accumulation of <= 4 partials)
    f = block_sum(f, sh);
    if (threadIdx.x == 0) out[0] = sqrtf(f);
}
void sumsq_partial(const float* g, int64_t n, float* partial, int nblocks, cudaStream_t st) {
    launch_k(sumsq_partial_kernel, nblocks, 256, 0, st, g, n, partial);
    LAUNCHED();
}
void sumsq_final(const float* partial, int np, float* out_norm, cudaStream_t st) {
    launch_k(sumsq_final_kernel, 1, 256, 0, st, partial, np, out_norm);
    LAUNCHED();
}
void sumsq_norm(const float* g, int64_t n, float* partial, float* out_norm, cudaStream_t st) {
    const int np = 1024;
    sumsq_partial(g, n, partial, np, st);
    sumsq_final(partial, np, out_norm, st);
}

__global__ void scale_kernel(float* __restrict__ g, int64_t n, float s) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] *= s;
}
void scale_f32(float* g, int64_t n, float s, cudaStream_t st) {
    if (n <= 0) return;
    launch_k(scale_kernel, 148 * 8, 256, 0, st, g, n, s);
    LAUNCHED();
}

// p, g, m, v fp32 (16 B read + 12 B write per parameter) + 2 B bf16 shadow write (+ 4 B when the gradient is
// cleared in the same pass: the fused zero_grad replaces a separate 4 B/param memset).
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p16, int64_t n, float lr, float b1, float b2, float eps, float wd, float step_size,
             float clip, const float* __restrict__ norm_ptr, float grad_div, int zero_g, NoDecay nd) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    float gs = grad_div;
    if (clip > 0.f && norm_ptr) {
        // clip_grad_norm_: coef = clip / (norm + 1e-6), clamped to 1; the norm is of the (already averaged) gradient
        const float coef = clip / (norm_ptr[0] * grad_div + 1e-6f);
        gs *= fminf(coef, 1.f);
    }
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
        // parameter groups of ref SingleRunner.py:186-205: names containing "bias" (= the two relative_attention_bias
        // tables) take weight_decay 0; the ranges are 4-aligned, so a float4 never straddles a boundary
        const int64_t e0 = i << 2;
        const float lwd = ((e0 >= nd.lo0 && e0 < nd.hi0) || (e0 >= nd.lo1 && e0 < nd.hi1)) ? 0.f : lr * wd;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = ga[j] * gs;
            ma[j] = ma[j] * b1 + gj * (1.f - b1);
            va[j] = va[j] * b2 + gj * gj * (1.f - b2);
            const float denom = sqrtf(va[j]) + eps;         // eps OUTSIDE the bias correction (HF 4.26)
            pa[j] = pa[j] - step_size * (ma[j] / denom);
            pa[j] = pa[j] - lwd * pa[j];                    // decoupled decay AFTER the update, un-corrected lr
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        if (zero_g) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // fused model.zero_grad()
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (p16) {
            uint2 t;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
            h[0] = __floats2bfloat162_rn(pa[0], pa[1]);
            h[1] = __floats2bfloat162_rn(pa[2], pa[3]);
            reinterpret_cast<uint2*>(p16)[i] = t;
        }
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gj = g[i] * gs;
        const float mj = m[i] * b1 + gj * (1.f - b1);
        const float vj = v[i] * b2 + gj * gj * (1.f - b2);
        float pj = p[i] - step_size * (mj / (sqrtf(vj) + eps));
        const float lwd = ((i >= nd.lo0 && i < nd.hi0) || (i >= nd.lo1 && i < nd.hi1)) ? 0.f : lr * wd;
        pj = pj - lwd * pj;
        p[i] = pj; m[i] = mj; v[i] = vj;
        if (zero_g) g[i] = 0.f;
        if (p16) p16[i] = __float2bfloat16_rn(pj);
    }
}

void adamw_flat(float* p, float* g, float* m, float* v, bf16* p16, int64_t n, float lr, float b1, float b2,
                float eps, float wd, int step, float clip, const float* norm_ptr, float grad_div, cudaStream_t st,
                bool zero_grad_after, NoDecay nd) {
    if (n <= 0) return;
    // step_size = lr * sqrt(1 - b2^t) / (1 - b1^t)   (correct_bias=True)
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
    launch_k(adamw_kernel, 148 * 8, 256, 0, st, p, g, m, v, p16, n, lr, b1, b2, eps, wd, step_size, clip, norm_ptr, grad_div, zero_grad_after ? 1 : 0, nd);
    LAUNCHED();
}

}  // namespace p5
