// bf16x3 — the tensor-core PARITY GEMM: fp32 operands through the SAME tcgen05 bf16 kernel (gemm_tc.cu) at fp32-class
// accuracy, so that the 1e-3 logits / loss gate of the north star is checked on the tensor-core path itself and not
// only on the SIMT fp32 path (SURVEY.md §7 "hard parts" 1).
//
// Every fp32 value x is split into two bf16 numbers  hi = bf16(x),  lo = bf16(x - hi)  (x = hi + lo up to 2^-17 |x|).
//   A.B^T = hi_a.hi_b + hi_a.lo_b + lo_a.hi_b + O(2^-16)
// The three products are ONE tcgen05 GEMM over a K-concatenated pair of bf16 operands
//   A' = [ hi_a | hi_a | lo_a ]   ([M, 3K]),    B' = [ hi_b | lo_b | hi_b ]   ([N, 3K])
// accumulated in fp32 in tensor memory: same kernel, same TMA / UMMA descriptors, same fused epilogues as the bf16
// training path; only the operands are prepared by the split kernel below.  MN-major operands ([K, rows] storage: the
// in-place dgrad / wgrad forms) are concatenated along their leading (k) dimension instead.
#include "common.cuh"
#include <vector>

namespace p5 {
extern int g_launches;

// in: fp32 [rows, K] (K-major, ld) or [K, rows] (MN-major, ld).  out: bf16, three K-segments, pattern chooses which
// segment holds hi / lo: pattern 0 (A operand) = hi, hi, lo;  pattern 1 (B operand) = hi, lo, hi.
__global__ void __launch_bounds__(256)
split_x3_kernel(const float* __restrict__ in, int64_t ld_in, bf16* __restrict__ out, int64_t ld_out, int rows, int K, int mn_major,
                int pattern) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t total = (int64_t)rows * K;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int r, k;
        int64_t src;
        if (!mn_major) { r = (int)(i / K); k = (int)(i - (int64_t)r * K); src = (int64_t)r * ld_in + k; }
        else { k = (int)(i / rows); r = (int)(i - (int64_t)k * rows); src = (int64_t)k * ld_in + r; }
        const float x = in[src];
        const bf16 hi = __float2bfloat16_rn(x);
        const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
        const bf16 s1 = pattern == 0 ? hi : lo, s2 = pattern == 0 ? lo : hi;
        if (!mn_major) {
            bf16* o = out + (int64_t)r * ld_out + k;
            o[0] = hi; o[K] = s1; o[2 * (int64_t)K] = s2;
        } else {
            out[(int64_t)k * ld_out + r] = hi;
            out[((int64_t)K + k) * ld_out + r] = s1;
            out[(2 * (int64_t)K + k) * ld_out + r] = s2;
        }
    }
}

struct X3Scratch { bf16* a = nullptr; bf16* b = nullptr; size_t a_cap = 0, b_cap = 0; };
static X3Scratch g_x3;

static bf16* grow(bf16*& p, size_t& cap, size_t elems, cudaStream_t st) {
    if (elems > cap) {
        if (p) { P5_CUDA(cudaStreamSynchronize(st)); P5_CUDA(cudaFree(p)); }
        cap = elems + elems / 4;
        P5_CUDA(cudaMalloc(&p, cap * sizeof(bf16)));
    }
    return p;
}
void gemm_x3_release() {
    if (g_x3.a) cudaFree(g_x3.a);
    if (g_x3.b) cudaFree(g_x3.b);
    g_x3 = X3Scratch();
}

bool gemm_x3_supported(const GemmProblem& p) {
    return p.nb1 == 1 && p.nb2 == 1 && p.A.dtype == DT_F32 && p.B.dtype == DT_F32 && p.M > 0 && p.N > 0 && p.K > 0;
}

static GemmOperand split_operand(const GemmOperand& op, int rows, int K, int pattern, bf16* dst, cudaStream_t st) {
    GemmOperand o = op;
    o.dtype = DT_BF16;
    o.ptr = dst;
    // zero fill: the K-major leading dimension is padded to a multiple of 8 elements (TMA stride alignment)
    if (op.major == MAJOR_K) {
        o.ld = round_up(3 * (int64_t)K, 8);
        if (o.ld != 3 * (int64_t)K) P5_CUDA(cudaMemsetAsync(dst, 0, (size_t)rows * o.ld * sizeof(bf16), st));
    } else {
        o.ld = round_up(rows, 8);
        if (o.ld != rows) P5_CUDA(cudaMemsetAsync(dst, 0, (size_t)3 * K * o.ld * sizeof(bf16), st));
    }
    const int64_t total = (int64_t)rows * K;
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv(total, 256), 148 * 16);
    launch_k(split_x3_kernel, grid, 256, 0, st, (const float*)op.ptr, op.ld, dst, o.ld, rows, K, op.major == MAJOR_MN ? 1 : 0, pattern);
    P5_CUDA(cudaGetLastError());
    ++g_launches;
    return o;
}

void gemm_bf16x3(const GemmProblem& p, cudaStream_t st) {
    P5_CHECK(gemm_x3_supported(p), "gemm_bf16x3: fp32, non-batched operands only");
    const size_t a_elems = p.A.major == MAJOR_K ? (size_t)p.M * round_up(3 * (int64_t)p.K, 8) : (size_t)3 * p.K * round_up(p.M, 8);
    const size_t b_elems = p.B.major == MAJOR_K ? (size_t)p.N * round_up(3 * (int64_t)p.K, 8) : (size_t)3 * p.K * round_up(p.N, 8);
    bf16* a = grow(g_x3.a, g_x3.a_cap, a_elems, st);
    bf16* b = grow(g_x3.b, g_x3.b_cap, b_elems, st);
    GemmProblem q = p;
    q.A = split_operand(p.A, p.M, p.K, 0, a, st);
    q.B = split_operand(p.B, p.N, p.K, 1, b, st);
    q.K = 3 * p.K;
    q.indep_of_prev = false;   // consumes the split kernels launched just before it
    P5_CHECK(gemm_tc_supported(q, true), "gemm_bf16x3: the split operands do not qualify for the tcgen05 kernel");
    gemm_tc(q, st);
}

}  // namespace p5
