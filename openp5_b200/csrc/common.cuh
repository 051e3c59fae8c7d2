// openp5_b200 — shared device/host helpers for the sm_100a hot path.
// Everything in csrc/ is compiled with -gencode arch=compute_100a,code=sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <utility>
#include <cstdlib>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

namespace p5 {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error handling: the C-ABI converts P5Error into a return code + p5_last_error() string
// ---------------------------------------------------------------------------------------------
struct P5Error : public std::runtime_error {
    int code;
    P5Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define P5_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            char _b[512];                                                                          \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                     __LINE__, cudaGetErrorString(_e));                                            \
            throw ::p5::P5Error(2, _b);                                                            \
        }                                                                                          \
    } while (0)

#define P5_CHECK(cond, msg)                                                               \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            char _b[512];                                                                 \
            snprintf(_b, sizeof(_b), "%s (%s) at %s:%d", msg, #cond, __FILE__, __LINE__); \
            throw ::p5::P5Error(1, _b);                                                   \
        }                                                                                 \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------------
enum DType : int { DT_F32 = 0, DT_BF16 = 1 };
static inline size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = DT_F32; };
template <> struct dtype_of<bf16> { static constexpr int value = DT_BF16; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float ld_as_f32(const void* p, int dt, int64_t i) {
    return dt == DT_F32 ? ((const float*)p)[i] : __bfloat162float(((const bf16*)p)[i]);
}
__device__ __forceinline__ void st_from_f32(void* p, int dt, int64_t i, float v) {
    if (dt == DT_F32) ((float*)p)[i] = v;
    else ((bf16*)p)[i] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------
// counter-based dropout RNG. keep(seed, site, idx) is a pure function so the backward pass
// regenerates the forward mask instead of storing it.  (Reference semantics: nn.Dropout(p) in
// HF:models/t5/modeling_t5.py:94,125,331,375; bit-parity with torch's Philox stream is not
// possible nor required — tests check keep-rate, scaling and fwd/bwd mask agreement.)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// The (seed, site) pair is folded into a 64-bit key (k0, k1) by two full mixing rounds; the key depends only on
// kernel-uniform values, so the compiler hoists it out of every element loop.  Per pair of elements the cost is one
// mixing round: h = mix32(pair ^ k0) + k1.  (mix32 is a bijection with full avalanche, so for a fixed key the
// counter stream is equidistributed; the additive k1 decorrelates the 16-bit halves between sites.  The pair
// counter is truncated to 32 bits: masks repeat after 2^33 elements, far beyond any tensor of this engine.)
__host__ __device__ __forceinline__ uint32_t drop_hash(uint64_t seed, uint32_t site, uint64_t idx) {
    const uint32_t s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32);
    const uint32_t k0 = mix32(mix32(s0 ^ 0x9e3779b9U) + 0x9e3779b9U * (site + 1u) + s1);
    const uint32_t k1 = mix32(k0 ^ 0x85ebca6bU);
    return mix32((uint32_t)idx ^ k0) + k1;
}
// The two halves of drop_hash for kernels that hash inside a hot loop: the key once per kernel (the compiler otherwise
// re-derives it on the uniform datapath in every iteration: ~7 issue slots per hash), the counter round per pair.
struct DropKey { uint32_t k0, k1; };
__host__ __device__ __forceinline__ DropKey drop_key(uint64_t seed, uint32_t site) {
    const uint32_t s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32);
    DropKey k;
    k.k0 = mix32(mix32(s0 ^ 0x9e3779b9U) + 0x9e3779b9U * (site + 1u) + s1);
    k.k1 = mix32(k.k0 ^ 0x85ebca6bU);
    return k;
}
__host__ __device__ __forceinline__ uint32_t drop_hash_k(DropKey k, uint32_t pair_idx) { return mix32(pair_idx ^ k.k0) + k.k1; }
// One 32-bit hash serves an aligned PAIR of elements (2k, 2k+1): the low 16 bits decide the even element, the high
// 16 bits the odd one (keep iff bits >= thr16, thr16 = round(p * 2^16)).  Kernels that own both elements of a pair
// (GEMM epilogue, softmax, dropout-cast) hash once per pair; element-wise kernels call drop_keep and get the same mask.
__host__ __device__ __forceinline__ bool drop_keep(uint64_t seed, uint32_t site, uint64_t idx, uint32_t thr) {
    const uint32_t h = drop_hash(seed, site, idx >> 1);
    return ((idx & 1) ? (h >> 16) : (h & 0xffffu)) >= (thr >> 16);
}
__host__ __device__ __forceinline__ void drop_keep2(uint64_t seed, uint32_t site, uint64_t even_idx, uint32_t thr, bool& k0,
                                                    bool& k1) {
    const uint32_t h = drop_hash(seed, site, even_idx >> 1);
    k0 = (h & 0xffffu) >= (thr >> 16);
    k1 = (h >> 16) >= (thr >> 16);
}
static inline uint32_t drop_threshold(float p) {
    double t = (double)p * 4294967296.0;
    if (t <= 0) return 0u;
    if (t >= 4294967295.0) return 4294967295u;
    return (uint32_t)t;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A train step is ~700 short kernels on one stream and a decode position ~170;
// with plain stream order every boundary costs grid drain + launch latency + the next kernel's prologue.  Every kernel
// of this library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with
// griddepcontrol.wait (blocks until the previous grid has COMPLETED and flushed its memory) followed by
// griddepcontrol.launch_dependents (lets the next grid's CTAs be scheduled as soon as every CTA of this grid has
// started, i.e. into SM slots as they drain).  Data hazards are impossible by construction: no kernel touches global
// memory before its wait, and since every kernel waits, completion order along the stream is transitive.  What
// overlaps is CTA scheduling, barrier/TMEM/tensor-map set-up (placed before the wait in the tcgen05 kernels) and the
// launch latency itself.  P5_NO_PDL=1 restores plain stream serialisation (the device instructions become no-ops).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = getenv("P5_NO_PDL") == nullptr;
    return on;
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    P5_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum; `sh` must hold >= 32 floats; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = (lane < nw) ? sh[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = (lane < nw) ? sh[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// ---------------------------------------------------------------------------------------------
// GEMM interface shared by the tcgen05 kernel (gemm_tc.cu) and the SIMT kernel (gemm_simt.cu)
// ---------------------------------------------------------------------------------------------
enum Major : int { MAJOR_K = 0, MAJOR_MN = 1 };

// A logical [rows, K] operand (rows = M for A, N for B), optionally batched over (b1, b2).
//   MAJOR_K : element (r, k) at ptr[r*ld + k]     (the reduction index is contiguous)
//   MAJOR_MN: element (r, k) at ptr[k*ld + r]     (the row index is contiguous)
struct GemmOperand {
    const void* ptr = nullptr;
    int dtype = DT_BF16;
    int major = MAJOR_K;
    int64_t ld = 0;
    int64_t bs1 = 0, bs2 = 0;  // batch strides in elements
    bool bcast1 = false, bcast2 = false;   // the operand is shared by every index of batch dimension 1 / 2 (stride 0)
};

enum EpiFlags : int {
    EPI_RELU = 1,        // v = max(v, 0)
    EPI_MULPOS = 2,      // v = aux[idx] > 0 ? v : 0        (ReLU backward from the saved activation)
    EPI_DROPOUT = 4,     // v = keep(seed, site, idx) ? v * inv_keep : 0
    EPI_ADD_RESID = 8,   // v += resid_f32[idx]
    EPI_ACCUM = 16,      // v += C_old[idx]                  (gradient accumulation, C must be fp32)
    EPI_ATOMIC = 32,     // atomicAdd(C[idx], v)             (split-K weight gradients; C must be fp32)
};

struct GemmEpilogue {
    void* C = nullptr;          // output, logical [M, N] row-major, ldc
    int c_dtype = DT_BF16;
    int64_t ldc = 0;
    int64_t cs1 = 0, cs2 = 0;   // batch strides (elements)
    float alpha = 1.f;
    int flags = 0;
    const void* aux = nullptr;  // same layout (ldc, cs1, cs2) as C, dtype aux_dtype
    int aux_dtype = DT_BF16;
    const float* resid = nullptr;  // fp32, same layout as C
    uint64_t seed = 0;
    uint32_t site = 0;
    uint32_t drop_thr = 0;
    float inv_keep = 1.f;
};

struct GemmProblem {
    int M = 0, N = 0, K = 0;
    int nb1 = 1, nb2 = 1;
    int prefer_bn = 0;   // tcgen05 path: 0 = heuristic, else force BLOCK_N (256 / 128 / 64)
    // The kernel launched immediately before this one on the stream neither produces an input of this GEMM nor reads
    // its output: the PDL wait moves to the END of the kernel, so this GEMM's CTAs fill the SMs that idle during the
    // partial last wave of its predecessor (completion order along the stream stays transitive).
    bool indep_of_prev = false;
    bool tail_filled = false;   // an independent kernel follows and fills this GEMM's partial last wave: keep the 256-wide tile
    GemmOperand A, B;
    GemmEpilogue epi;
};

// shared epilogue math: acc -> value to store; idx is the element offset within C (incl. batch)
__device__ __forceinline__ float epilogue_apply(const GemmEpilogue& e, float acc, int64_t idx) {
    float v = acc * e.alpha;
    if (e.flags & EPI_RELU) v = fmaxf(v, 0.f);
    if (e.flags & EPI_MULPOS) v = ld_as_f32(e.aux, e.aux_dtype, idx) > 0.f ? v : 0.f;
    if (e.flags & EPI_DROPOUT) v = drop_keep(e.seed, e.site, (uint64_t)idx, e.drop_thr) ? v * e.inv_keep : 0.f;
    if (e.flags & EPI_ADD_RESID) v += e.resid[idx];
    if (e.flags & EPI_ACCUM) v += ((const float*)e.C)[idx];
    return v;
}

// SMs the persistent kernels (tcgen05 GEMM, fused attention, RMSNorm backward) size their grids to.  During a
// data-parallel backward the NCCL all-reduce kernels (P5_COMM_CTAS CTAs, comm.cu) run concurrently on a side stream; a
// persistent grid of one 200 KB-smem CTA per SM cannot share an SM with them, and with a static tile stride the CTAs that
// find their SM occupied make the whole GEMM wait.  While a communicator is active the persistent grids therefore leave
// that many SMs to NCCL.  (gemm_tc.cu)
int sm_budget();
void sm_reserve(int n_sms);

// launchers (throw P5Error)
void gemm_simt(const GemmProblem& p, cudaStream_t stream);
bool gemm_tc_supported(const GemmProblem& p, bool allow_mn_major);
void gemm_tc(const GemmProblem& p, cudaStream_t stream);
void gemm_tc_clear_cache();
int  gemm_tc_launch_count();
int  gemm_tc_tile_width(int M, int N, int batches, int sms);   // the tile-width cost model (host arithmetic only)

}  // namespace p5
